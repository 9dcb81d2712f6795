"""Pins the oracle's leaf arithmetic against the reference's OWN known-answer tests:
tests/sampling.rs:14-22,24-64,110-135,254-349, tests/shapes.rs:17-122,156-198,480-504, tests/bsdfs.rs:482-605.
The oracle is a CPU restatement (the Rust reference cannot be built here); image-level parity stays "unpinned"."""
import ctypes as C
import math

import numpy as np
import pytest


def rev32(n):
    return int("{:032b}".format(n)[::-1], 2)


def f32(x):
    return np.float32(x)


def test_radical_inverse_base2_exact(oracle):  # tests/sampling.rs:14-22
    for a in range(1024):
        assert f32(oracle.orc_radical_inverse(0, a)) == f32(f32(rev32(a)) * f32(2.3283064365386963e-10))


def _pcg(seq=None):
    state, inc = 0x853C49E6748FEA9B, 0xDA3E39CB94B95BDB
    M = (1 << 64) - 1

    def step():
        nonlocal state
        old = state
        state = (old * 0x5851F42D4C957F2D + inc) & M
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF
    if seq is not None:
        state = 0
        inc = ((seq << 1) | 1) & M
        step()
        state = (state + 0x853C49E6748FEA9B) & M
        step()
    return step


def test_rng_matches_pcg32(oracle):  # core/rng.rs:3-66 vs an independent PCG32 (XSH-RR) written from the published algorithm
    for seq in (None, 0, 1, 7, 12345):
        step = _pcg(seq)
        out = np.zeros(64, np.uint32)
        oracle.orc_rng_sequence(seq or 0, 0 if seq is None else 1, 64, out.ctypes.data, None)
        assert [int(v) for v in out] == [step() for _ in range(64)]


def test_scrambled_radical_inverse_vs_pbrt_v2(oracle):  # tests/sampling.rs:24-64
    for dim in range(128):
        step = _pcg(dim)
        base = int(oracle.orc_prime(dim))
        perm = [base - 1 - i for i in range(base)]
        # shuffle_array(perm, base, 1, rng) with uniform_uint32_threshold (sampling.rs:4-15, rng.rs:48-56)
        for i in range(base):
            b = base - i
            thr = ((1 << 32) - b) % b
            while True:
                r = step()
                if r >= thr:
                    break
            other = i + r % b
            perm[i], perm[other] = perm[other], perm[i]
        parr = np.array(perm, np.uint16)
        for index in (0, 1, 2, 1151, 32351, 4363211, 681122):
            val = f32(0.0)
            inv_base = f32(1.0) / f32(base)
            inv_bi = inv_base
            n = index
            while n > 0:
                val = f32(val + f32(perm[n % base]) * inv_bi)
                n = int(f32(f32(n) * inv_base))
                inv_bi = f32(inv_bi * inv_base)
            val = f32(val + f32(perm[0]) * f32(base) / f32(base - 1.0) * inv_bi)
            got = oracle.orc_scrambled_radical_inverse(dim, index, parr.ctypes.data)
            assert abs(float(val) - got) < 1e-5, (dim, index, val, got)


def test_sobol_first_dimension_is_van_der_corput(oracle):  # tests/sampling.rs:126-134
    for i in range(8192):
        assert f32(oracle.orc_sobol_sample(i, 0, 0)) == f32(f32(rev32(i)) * f32(2.3283064365386963e-10))


def test_sobol_interval_index_covers_pixels(oracle):
    # sobol_interval_to_index maps (pixel, frame) to an index whose first two dimensions fall in that pixel
    m = 4
    res = 1 << m
    for px, py, frame in [(0, 0, 0), (3, 5, 1), (15, 15, 7), (9, 2, 3)]:
        idx = oracle.orc_sobol_interval_to_index(m, frame, px, py)
        x = oracle.orc_sobol_sample(idx, 0, 0) * res
        y = oracle.orc_sobol_sample(idx, 1, 0) * res
        assert int(x) == px and int(y) == py, (px, py, frame, x, y)


def test_distribution_1d_discrete(oracle):  # tests/sampling.rs:254-321
    func = np.array([0.0, 1.0, 0.0, 3.0], np.float32)
    pdf = C.c_float(0)

    def sd(u):
        off = oracle.orc_distribution1d_sample_discrete(func.ctypes.data, 4, C.c_float(u), C.addressof(pdf))
        return off, pdf.value
    assert sd(0.0) == (1, 0.25)
    assert sd(0.125) == (1, 0.25)
    assert sd(0.24999) == (1, 0.25)
    assert sd(0.250001) == (3, 0.75)
    assert sd(0.625) == (3, 0.75)
    assert sd(0.99999994) == (3, 0.75)
    assert sd(1.0) == (3, 0.75)
    u = f32(0.25)
    umax = f32(0.25)
    for _ in range(20):
        u = np.nextafter(u, f32(-1))
        umax = np.nextafter(umax, f32(2))
    while u < umax:
        if sd(float(u))[0] == 3:
            break
        assert sd(float(u))[0] == 1
        u = np.nextafter(u, f32(2))
    assert u < umax
    while u <= umax:
        assert sd(float(u))[0] == 3
        u = np.nextafter(u, f32(2))


def test_distribution_1d_continuous(oracle):  # tests/sampling.rs:323-349
    func = np.array([1.0, 1.0, 2.0, 4.0, 8.0], np.float32)
    pdf, off = C.c_float(0), C.c_uint32(0)

    def sc(u):
        v = oracle.orc_distribution1d_sample_continuous(func.ctypes.data, 5, C.c_float(u), C.addressof(pdf), C.addressof(off))
        return v, pdf.value, off.value
    assert sc(0.0) == (0.0, 5 * 1.0 / 16.0, 0)
    assert abs(sc(0.5)[0] - 0.8) < 1e-6
    v, p, o = sc(0.75)
    assert abs(v - 0.9) < 1e-6 and p == 5 * 8.0 / 16.0 and o == 4
    assert sc(1.0)[0] == 1.0


def test_triangle_badcase_is_dropped(P):
    """tests/shapes.rs:480-504: the bad-case triangle is collinear; create_triangle_mesh drops it (area <= 1e-16,
    shapes/triangle.rs:726) so the reference's assertion is vacuous.  The front end must drop it too."""
    txt = ('Camera "perspective"\nWorldBegin\nShape "trianglemesh" "integer indices" [0 1 2] "point P" '
           '[-1113.45459 -79.049614 -56.2431908 -1113.45459 -87.0922699 -56.2431908 -1113.45459 -79.2090149 -56.2431908]\nWorldEnd\n')
    hs = P.HostScene.parse_string(txt)
    assert hs.desc.n_tris == 0


def _sphere_mesh(rng, n_theta=16, n_phi=16):
    # tests/shapes.rs:22-66: a closed, randomly perturbed sphere mesh
    verts = []
    for t in range(n_theta):
        theta = math.pi * t / (n_theta - 1)
        for p in range(n_phi):
            phi = 2 * math.pi * p / (n_phi - 1)
            x, y, z = math.sin(theta) * math.cos(phi), math.sin(theta) * math.sin(phi), math.cos(theta)
            verts.append([x, y, z])
    verts = np.array(verts, np.float32)
    # make the seam / poles watertight by sharing vertices, then perturb
    idx = np.arange(n_theta * n_phi).reshape(n_theta, n_phi)
    idx[:, -1] = idx[:, 0]
    idx[0, :] = idx[0, 0]
    idx[-1, :] = idx[-1, 0]
    verts = verts * (1.0 + 0.3 * (rng.rand(len(verts), 1).astype(np.float32) - 0.5)) * 5.0
    tris = []
    for t in range(n_theta - 1):
        for p in range(n_phi - 1):
            a, b, c, d = idx[t, p], idx[t + 1, p], idx[t + 1, p + 1], idx[t, p + 1]
            tris.append([a, b, c])
            tris.append([a, c, d])
    return verts, np.array(tris)


def test_triangle_watertight(oracle):  # tests/shapes.rs:17-122
    rng = np.random.RandomState(3)
    verts, tris = _sphere_mesh(rng)
    P = verts[tris].reshape(-1, 9).astype(np.float32)
    n_rays = 2000
    for i in range(n_rays):
        if i % 3 == 0:  # aim exactly at a vertex
            target = verts[rng.randint(len(verts))]
            d = target
        else:
            d = rng.normal(size=3)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        ray = np.array([0, 0, 0, np.inf, d[0], d[1], d[2], 0], np.float32)
        hit = False
        for k in range(P.shape[0]):
            if oracle.orc_tri_intersect_raw(P[k].ctypes.data, 4, ray.ctypes.data, None, None):
                hit = True
                break
        assert hit, (i, d)


def test_triangle_reintersect(oracle):  # tests/shapes.rs:156-198: a ray spawned from a hit never re-hits the triangle
    rng = np.random.RandomState(5)
    t = C.c_float(0)
    bary = np.zeros(3, np.float32)
    out = np.zeros(8, np.float32)
    misses = 0
    for i in range(2000):
        scale = 10.0 ** rng.uniform(-3, 3)
        p = (rng.uniform(-1, 1, 9) * scale).astype(np.float32)
        o = (rng.uniform(-1, 1, 3) * scale * 4).astype(np.float32)
        b = rng.dirichlet([1, 1, 1])
        target = b[0] * p[0:3] + b[1] * p[3:6] + b[2] * p[6:9]
        d = (target - o).astype(np.float32)
        ray = np.array([o[0], o[1], o[2], np.inf, d[0], d[1], d[2], 0], np.float32)
        if not oracle.orc_tri_intersect_raw(p.ctypes.data, 4, ray.ctypes.data, C.addressof(t), bary.ctypes.data):
            misses += 1
            continue
        for _ in range(8):
            w = rng.normal(size=3).astype(np.float32)
            oracle.orc_spawn_ray_from_hit(p.ctypes.data, 4, ray.ctypes.data, bary.ctypes.data, w.ctypes.data, out.ctypes.data)
            assert oracle.orc_tri_intersect_raw(p.ctypes.data, 4, out.ctypes.data, None, None) == 0
    assert misses < 400


@pytest.mark.parametrize("kind,alpha,vis", [(0, 0.0, 1), (1, 0.5, 1), (1, 0.1, 1), (1, 0.5, 0), (1, 0.3, 0)])
def test_bxdf_sampling_matches_pdf(oracle, kind, alpha, vis):
    """tests/bsdfs.rs:482-605 in spirit: sample_f's empirical distribution must integrate like pdf (chi^2-style
    bucket comparison on a coarse (cos theta, phi) grid)."""
    rng = np.random.RandomState(11)
    wo = np.array([0.3, 0.2, 0.9327], np.float32)
    wo /= np.linalg.norm(wo)
    nt, nph = 8, 16
    counts = np.zeros((nt, nph))
    n = 40000
    wi = np.zeros(3, np.float32)
    pdf = C.c_float(0)
    got = 0
    for i in range(n):
        u = rng.rand(2).astype(np.float32)
        if not oracle.orc_bxdf_sample(kind, C.c_float(alpha), vis, wo.ctypes.data, u.ctypes.data, wi.ctypes.data, C.addressof(pdf)):
            continue
        if wi[2] <= 0:
            continue
        got += 1
        ti = min(int(wi[2] * nt), nt - 1)
        pi = min(int((math.atan2(wi[1], wi[0]) + math.pi) / (2 * math.pi) * nph), nph - 1)
        counts[ti, pi] += 1
    # expected counts from pdf by midpoint quadrature (16x16 sub-samples per cell)
    expected = np.zeros((nt, nph))
    sub = 12
    for ti in range(nt):
        for pi in range(nph):
            acc = 0.0
            for a in range(sub):
                for b in range(sub):
                    ct = (ti + (a + 0.5) / sub) / nt
                    ph = (pi + (b + 0.5) / sub) / nph * 2 * math.pi - math.pi
                    st = math.sqrt(max(0.0, 1 - ct * ct))
                    w = np.array([st * math.cos(ph), st * math.sin(ph), ct], np.float32)
                    acc += oracle.orc_bxdf_pdf(kind, C.c_float(alpha), vis, wo.ctypes.data, w.ctypes.data)
            expected[ti, pi] = acc / (sub * sub) * (1.0 / nt) * (2 * math.pi / nph)
    expected *= n
    mask = expected > 50
    rel = np.abs(counts[mask] - expected[mask]) / np.sqrt(expected[mask])
    # well-populated cells agree within ~5 sigma (+ quadrature error for peaky lobes)
    assert np.median(rel) < 1.5 and np.percentile(rel, 95) < 6.0, (np.median(rel), rel.max())
    assert abs(counts.sum() - expected.sum()) / n < 0.05
