"""Synthetic, seeded .pbrt scenes for BASELINE.json's configs.

Every generator returns .pbrt TEXT: the same text goes through the product's front end (pt_parse_string) and
the resulting pt_scene_desc is what both the CUDA path and the CPU oracle consume.  Floats are printed with
'%.9g' so that strtof() recovers the exact f32.
"""
import numpy as np


def _fmt(a):
    return " ".join("%.9g" % v for v in np.asarray(a, dtype=np.float32).ravel())


def _quad(p, idx=(0, 1, 2, 0, 2, 3)):
    return 'Shape "trianglemesh" "integer indices" [%s] "point P" [%s]\n' % (" ".join(map(str, idx)), _fmt(p))


_CORNELL_HEAD = """LookAt 278 273 -800  278 273 0  0 1 0
Camera "perspective" "float fov" [39.3077]
Film "image" "integer xresolution" [{res}] "integer yresolution" [{res}] "string filename" ["cornell.exr"]
Sampler "{sampler}" "integer pixelsamples" [{spp}]
Integrator "path" "integer maxdepth" [{maxdepth}]
Accelerator "bvh"
WorldBegin
"""


def cornell_box(res=256, spp=16, maxdepth=5, sampler="halton", extra_lights=False, materials="matte"):
    """Config 1 / 4: the classic Cornell box, 32 matte triangles + a 2-triangle DiffuseAreaLight."""
    s = _CORNELL_HEAD.format(res=res, spp=spp, maxdepth=maxdepth, sampler=sampler)
    white = 'Material "matte" "rgb Kd" [0.725 0.71 0.68]\n'
    red = 'Material "matte" "rgb Kd" [0.63 0.065 0.05]\n'
    green = 'Material "matte" "rgb Kd" [0.14 0.45 0.091]\n'
    if materials == "mixed":  # exercises plastic / metal / glass / mirror on the two blocks
        short_m = 'Material "glass" "float eta" [1.5]\n'
        tall_m = 'Material "metal" "rgb eta" [0.2 0.92 1.1] "rgb k" [3.9 2.45 2.14] "float roughness" [0.05]\n'
        floor_m = 'Material "plastic" "rgb Kd" [0.4 0.4 0.4] "rgb Ks" [0.3 0.3 0.3] "float roughness" [0.1]\n'
        back_m = 'Material "mirror" "rgb Kr" [0.8 0.8 0.8]\n'
    else:
        short_m = tall_m = floor_m = back_m = white
    s += "AttributeBegin\n" + floor_m + _quad([552.8, 0, 0, 0, 0, 0, 0, 0, 559.2, 549.6, 0, 559.2]) + "AttributeEnd\n"
    s += "AttributeBegin\n" + white + _quad([556, 548.8, 0, 556, 548.8, 559.2, 0, 548.8, 559.2, 0, 548.8, 0]) + "AttributeEnd\n"
    s += "AttributeBegin\n" + back_m + _quad([549.6, 0, 559.2, 0, 0, 559.2, 0, 548.8, 559.2, 556, 548.8, 559.2]) + "AttributeEnd\n"
    s += "AttributeBegin\n" + green + _quad([0, 0, 559.2, 0, 0, 0, 0, 548.8, 0, 0, 548.8, 559.2]) + "AttributeEnd\n"
    s += "AttributeBegin\n" + red + _quad([552.8, 0, 0, 549.6, 0, 559.2, 556, 548.8, 559.2, 556, 548.8, 0]) + "AttributeEnd\n"
    short = [
        [130, 165, 65, 82, 165, 225, 240, 165, 272, 290, 165, 114],
        [290, 0, 114, 290, 165, 114, 240, 165, 272, 240, 0, 272],
        [130, 0, 65, 130, 165, 65, 290, 165, 114, 290, 0, 114],
        [82, 0, 225, 82, 165, 225, 130, 165, 65, 130, 0, 65],
        [240, 0, 272, 240, 165, 272, 82, 165, 225, 82, 0, 225],
    ]
    tall = [
        [423, 330, 247, 265, 330, 296, 314, 330, 456, 472, 330, 406],
        [423, 0, 247, 423, 330, 247, 472, 330, 406, 472, 0, 406],
        [472, 0, 406, 472, 330, 406, 314, 330, 456, 314, 0, 456],
        [314, 0, 456, 314, 330, 456, 265, 330, 296, 265, 0, 296],
        [265, 0, 296, 265, 330, 296, 423, 330, 247, 423, 0, 247],
    ]
    s += "AttributeBegin\n" + short_m + "".join(_quad(q) for q in short) + "AttributeEnd\n"
    s += "AttributeBegin\n" + tall_m + "".join(_quad(q) for q in tall) + "AttributeEnd\n"
    s += "AttributeBegin\n" + 'AreaLightSource "diffuse" "rgb L" [17 12 4]\n' + 'Material "matte" "rgb Kd" [0 0 0]\n'
    s += _quad([343, 548.7, 227, 343, 548.7, 332, 213, 548.7, 332, 213, 548.7, 227]) + "AttributeEnd\n"
    if extra_lights:
        s += 'AttributeBegin\nLightSource "point" "rgb I" [40000 40000 60000] "point from" [100 400 100]\nAttributeEnd\n'
        s += 'AttributeBegin\nLightSource "distant" "rgb L" [0.8 0.7 0.6] "point from" [0 1 -1] "point to" [0 0 0]\nAttributeEnd\n'
    s += "WorldEnd\n"
    return s


def _mesh_text(P, idx, extra=""):
    return 'Shape "trianglemesh" %s "integer indices" [%s] "point P" [%s]\n' % (
        extra, " ".join(map(str, np.asarray(idx).ravel())), _fmt(P))


def displaced_grid(n=708, res=1024, spp=64, maxdepth=8, seed=7, sampler="halton", normals=False):
    """Config 2: a displaced n x n vertex grid (2(n-1)^2 triangles; n=708 -> 999,698) split in three material
    bands (plastic / copper metal / glass), a matte ground plane, one quad area light and one distant light."""
    rng = np.random.RandomState(seed)
    xs = np.linspace(-10.0, 10.0, n, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    Y = np.zeros_like(X)
    for _ in range(6):
        fx, fz = rng.uniform(0.2, 1.6, 2)
        ph = rng.uniform(0, 2 * np.pi, 2)
        amp = rng.uniform(0.15, 0.5)
        Y += (amp * np.sin(fx * X + ph[0]) * np.cos(fz * Z + ph[1])).astype(np.float32)
    Y += 1.0
    P = np.stack([X, Y, Z], axis=-1).astype(np.float32)  # [row z][col x]
    s = """LookAt 0 9 -16  0 0.5 0  0 1 0
Camera "perspective" "float fov" [45]
Film "image" "integer xresolution" [%d] "integer yresolution" [%d] "string filename" ["grid.exr"]
Sampler "%s" "integer pixelsamples" [%d]
Integrator "path" "integer maxdepth" [%d]
Accelerator "bvh"
WorldBegin
""" % (res, res, sampler, spp, maxdepth)
    s += 'AttributeBegin\nLightSource "distant" "rgb L" [1.2 1.1 1.0] "point from" [3 8 -4] "point to" [0 0 0]\nAttributeEnd\n'
    s += 'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [30 28 24]\nMaterial "matte" "rgb Kd" [0 0 0]\n'
    s += _quad([-3, 9, -3, 3, 9, -3, 3, 9, 3, -3, 9, 3], idx=(0, 1, 2, 0, 2, 3)) + "AttributeEnd\n"
    s += 'AttributeBegin\nMaterial "matte" "rgb Kd" [0.5 0.5 0.5]\n' + _quad([-40, -0.5, -40, -40, -0.5, 40, 40, -0.5, 40, 40, -0.5, -40]) + "AttributeEnd\n"
    mats = [
        'Material "plastic" "rgb Kd" [0.6 0.2 0.15] "rgb Ks" [0.3 0.3 0.3] "float roughness" [0.1]\n',
        'Material "metal" "float roughness" [0.01]\n',
        'Material "glass" "float eta" [1.5]\n',
    ]
    bands = [0, (n - 1) // 3, 2 * (n - 1) // 3, n - 1]
    for b in range(3):
        r0, r1 = bands[b], bands[b + 1]
        rows = r1 - r0 + 1
        Pb = P[r0:r1 + 1].reshape(-1, 3)
        ii, jj = np.meshgrid(np.arange(rows - 1), np.arange(n - 1), indexing="ij")
        v00 = ii * n + jj
        tri = np.stack([v00, v00 + n, v00 + n + 1, v00, v00 + n + 1, v00 + 1], axis=-1).reshape(-1, 3)
        extra = ""
        if normals:
            # smooth vertex normals from central differences of the height field
            dydx = np.gradient(Y, xs, axis=1)
            dydz = np.gradient(Y, xs, axis=0)
            N = np.stack([-dydx, np.ones_like(Y), -dydz], axis=-1)[r0:r1 + 1].reshape(-1, 3)
            extra = '"normal N" [%s]' % _fmt(N)
        s += "AttributeBegin\n" + mats[b] + _mesh_text(Pb, tri, extra) + "AttributeEnd\n"
    s += "WorldEnd\n"
    return s


def _blob(n_sub, seed):
    """Seeded noisy icosphere: 20 * 4^n_sub triangles."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    V = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    F = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7],
                  [9, 8, 1]], dtype=np.int64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    for _ in range(n_sub):
        cache = {}
        V = list(V)
        newF = []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = (V[a] + V[b]) / 2.0
                V.append(m / np.linalg.norm(m))
                cache[k] = len(V) - 1
            return cache[k]

        for a, b, c in F:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            newF += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        V = np.array(V)
        F = np.array(newF, dtype=np.int64)
    rng = np.random.RandomState(seed)
    k = rng.uniform(1.0, 4.0, (4, 3))
    ph = rng.uniform(0, 2 * np.pi, 4)
    r = 1.0 + sum(0.08 * np.sin(V @ k[i] + ph[i]) for i in range(4))
    return (V * r[:, None]).astype(np.float32), F


def instanced_blobs(n_instances=1000, n_sub=4, res=(1920, 1080), spp=256, maxdepth=8, n_lights=256, seed=11, sampler="halton"):
    """Config 3 / 5: one noisy icosphere (20*4^n_sub triangles; n_sub=4 -> 5120, use n_sub=5 for 20480) instanced
    n_instances times on a jittered grid, plus n_lights small emissive quads (2 area-light triangles each)."""
    V, F = _blob(n_sub, seed)
    rng = np.random.RandomState(seed + 1)
    s = """LookAt 0 14 -34  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [%d] "integer yresolution" [%d] "string filename" ["blobs.exr"]
Sampler "%s" "integer pixelsamples" [%d]
Integrator "path" "integer maxdepth" [%d]
Accelerator "bvh"
WorldBegin
""" % (res[0], res[1], sampler, spp, maxdepth)
    mats = ['Material "plastic" "rgb Kd" [0.2 0.3 0.7] "rgb Ks" [0.4 0.4 0.4] "float roughness" [0.05]\n',
            'Material "metal" "float roughness" [0.05]\n', 'Material "glass"\n', 'Material "matte" "rgb Kd" [0.7 0.6 0.3]\n']
    for m in range(4):
        s += 'ObjectBegin "blob%d"\n' % m + mats[m] + _mesh_text(V, F) + "ObjectEnd\n"
    side = int(np.ceil(n_instances ** (1.0 / 3.0) * 1.6))
    cells = [(i, j, k) for i in range(side) for k in range(side) for j in range(max(1, side // 4))]
    rng.shuffle(cells)
    cells = cells[:n_instances]
    spacing = 40.0 / side
    for c, (i, j, k) in enumerate(cells):
        pos = (np.array([i, j, k], dtype=np.float64) - np.array([side / 2, 0, side / 2])) * spacing + rng.uniform(-0.3, 0.3, 3) * spacing
        ax = rng.normal(size=3)
        ang = rng.uniform(0, 360)
        sc = rng.uniform(0.35, 0.6) * spacing
        s += "AttributeBegin\nTranslate %s\nRotate %.9g %s\nScale %.9g %.9g %.9g\nObjectInstance \"blob%d\"\nAttributeEnd\n" % (
            _fmt(pos), ang, _fmt(ax), sc, sc, sc, c % 4)
    s += 'AttributeBegin\nMaterial "matte" "rgb Kd" [0.45 0.45 0.45]\n' + _quad([-60, -2, -60, -60, -2, 60, 60, -2, 60, 60, -2, -60]) + "AttributeEnd\n"
    for _ in range(n_lights):
        c = np.array([rng.uniform(-20, 20), rng.uniform(8, 14), rng.uniform(-20, 20)])
        h = rng.uniform(0.2, 0.5)
        col = rng.uniform(20, 60, 3)
        q = [c[0] - h, c[1], c[2] - h, c[0] + h, c[1], c[2] - h, c[0] + h, c[1], c[2] + h, c[0] - h, c[1], c[2] + h]
        s += 'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [%s]\nMaterial "matte" "rgb Kd" [0 0 0]\n' % _fmt(col) + _quad(q) + "AttributeEnd\n"
    s += "WorldEnd\n"
    return s
