"""The committed golden fixtures (tools/make_golden.py) against the oracle on CPU, and against the CUDA path on GPU."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_golden(P):
    import oracle_api
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=64, spp=4))
    orc = oracle_api.OracleScene(hs)
    z = np.load(os.path.join(G, "cornell_64_streams.npz"))
    assert np.array_equal(orc.sampler_values(z["pix"], z["sn"], 0, 32).view(np.uint32), z["halton"].view(np.uint32))
    assert np.array_equal(orc.camera_rays(z["pix"], z["sn"])[0].view(np.uint32), z["rays"].view(np.uint32))
    rgb, st = orc.render(threads=1)
    assert np.array_equal(rgb, np.load(os.path.join(G, "cornell_64_4spp.npy")))
    assert [st[k] for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested")] == list(z["counters"])
    # tile-parallel rendering only changes the film merge order
    rgb8, _ = orc.render(threads=8)
    assert np.sqrt(((rgb8 - rgb) ** 2).mean()) / rgb.mean() < 1e-6
    hs2 = P.HostScene.parse_string(P.scenes.cornell_box(res=64, spp=4, sampler="sobol"))
    orc2 = oracle_api.OracleScene(hs2)
    assert np.array_equal(orc2.sampler_values(z["pix"], z["sn"], 0, 32).view(np.uint32), np.load(os.path.join(G, "cornell_64_sobol_values.npy")).view(np.uint32))


def _widening_scenes(P):
    from test_hostcheck import aov_scene, spot_scene
    return {"spot": spot_scene(P, res=32, spp=4), "aov_ns": aov_scene(P, "ns", 0.75, res=32, spp=2),
            "lbvh": P.scenes.cornell_box(res=32, spp=2).replace('Accelerator "bvh"', 'Accelerator "bvh" "string layout" "lbvh"')}


def test_oracle_reproduces_golden_widening(P):
    """Spot lights + mixed materials, the aov integrator on a mesh with normals / tangents, the LBVH layout."""
    import oracle_api
    z = np.load(os.path.join(G, "widening_32.npz"))
    K = ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested")
    for name, txt in _widening_scenes(P).items():
        rgb, st = oracle_api.OracleScene(P.HostScene.parse_string(txt)).render(threads=1)
        assert np.array_equal(rgb, z[name]), name
        if name + "_counters" in z.files:
            assert [st[k] for k in K] == list(z[name + "_counters"]), name


def round2_scenes(P):
    import tempfile
    from test_hostcheck import ao_scene, tree_scene, env_scene, extra_materials_scene
    return {"ao": ao_scene(P, nsamples=4, res=24, spp=2), "direct_all": tree_scene(P, '"directlighting"', res=24, spp=2, nsamples=2),
            "whitted": tree_scene(P, '"whitted"', res=24, spp=2), "env_path": env_scene(P, tempfile.gettempdir(), mapped=False, res=24, spp=2),
            "materials_path": extra_materials_scene(P, res=24, spp=2)}


R2_KEYS = ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices")


def test_oracle_reproduces_golden_round2(P):
    """ambientocclusion, directlighting (sample arrays), whitted, the infinite light, uber / substrate / translucent."""
    import oracle_api
    z = np.load(os.path.join(G, "round2_24.npz"))
    for name, txt in round2_scenes(P).items():
        rgb, st = oracle_api.OracleScene(P.HostScene.parse_string(txt)).render(threads=1)
        assert np.array_equal(rgb, z[name]), name
        assert [st[k] for k in R2_KEYS] == list(z[name + "_counters"]), name


@pytest.mark.gpu
def test_gpu_reproduces_golden_round2(P):
    z = np.load(os.path.join(G, "round2_24.npz"))
    for name, txt in round2_scenes(P).items():
        sc = P.Scene(P.HostScene.parse_string(txt), 0)
        rgb, st = sc.render(collect_counters=1)
        assert np.sqrt(((rgb - z[name]) ** 2).mean()) / z[name].mean() <= 1e-3, name
        assert [st[k] for k in R2_KEYS] == list(z[name + "_counters"]), name
        sc.close()


def test_oracle_furnace(P):
    """Analytic check the reference has no test for: a closed diffuse box lit by a two-sided emitter in the middle.
    With albedo rho and uniform emission, radiance converges to a finite positive value and energy is conserved:
    increasing maxdepth never decreases the image and the increments shrink geometrically (ratio <= rho)."""
    import oracle_api
    def scene(depth):
        q = lambda p: 'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [%s]\n' % " ".join(map(str, p))
        s = 'LookAt 0 0 -0.9  0 0 0  0 1 0\nCamera "perspective" "float fov" [90]\nFilm "image" "integer xresolution" [16] "integer yresolution" [16]\n'
        s += 'Sampler "halton" "integer pixelsamples" [64]\nIntegrator "path" "integer maxdepth" [%d] "float rrthreshold" [0]\nWorldBegin\nMaterial "matte" "rgb Kd" [0.5 0.5 0.5]\n' % depth
        c = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
        for f in [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]:
            s += q([v for i in f for v in c[i]])
        s += 'AreaLightSource "diffuse" "rgb L" [1 1 1] "bool twosided" ["true"]\n' + q([-0.2, 0.5, -0.2, 0.2, 0.5, -0.2, 0.2, 0.5, 0.2, -0.2, 0.5, 0.2]) + "WorldEnd\n"
        return s
    means = []
    for depth in (1, 2, 3, 4):
        hs = P.HostScene.parse_string(scene(depth))
        means.append(float(oracle_api.OracleScene(hs).render(threads=4)[0].mean()))
    inc = np.diff(means)
    assert (inc > 0).all() and inc[1] / inc[0] < 0.6 and inc[2] / inc[1] < 0.6, (means, inc)


@pytest.mark.gpu
def test_gpu_reproduces_golden(P):
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=64, spp=4))
    sc = P.Scene(hs, 0)
    z = np.load(os.path.join(G, "cornell_64_streams.npz"))
    assert np.array_equal(sc.sampler_values(z["pix"], z["sn"], 0, 32).view(np.uint32), z["halton"].view(np.uint32))
    assert np.array_equal(sc.camera_rays(z["pix"], z["sn"])[0].view(np.uint32), z["rays"].view(np.uint32))
    rgb, st = sc.render(collect_counters=1)
    ref = np.load(os.path.join(G, "cornell_64_4spp.npy"))
    assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() <= 1e-3
    assert [st[k] for k in ("camera_rays", "regular_rays", "shadow_rays")] == list(z["counters"][:3])
    hs2 = P.HostScene.parse_string(P.scenes.cornell_box(res=64, spp=4, sampler="sobol"))
    sc2 = P.Scene(hs2, 0)
    assert np.array_equal(sc2.sampler_values(z["pix"], z["sn"], 0, 32).view(np.uint32), np.load(os.path.join(G, "cornell_64_sobol_values.npy")).view(np.uint32))


@pytest.mark.gpu
def test_gpu_reproduces_golden_widening(P):
    z = np.load(os.path.join(G, "widening_32.npz"))
    K = ("camera_rays", "regular_rays", "shadow_rays")
    for name, txt in _widening_scenes(P).items():
        sc = P.Scene(P.HostScene.parse_string(txt), 0)
        rgb, st = sc.render(collect_counters=1)
        ref = z[name]
        assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() <= 1e-3, name
        if name + "_counters" in z.files:
            assert [st[k] for k in K] == list(z[name + "_counters"][:3]), name
            if name == "lbvh":
                assert [st[k] for k in ("nodes_visited", "tris_tested")] == list(z[name + "_counters"][3:]), name
        sc.close()
