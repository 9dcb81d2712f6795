"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI (ctypes) with host buffers
and is compared with the CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star): sampler values and
primary-hit primitive ids bit-exact, hit t within 1e-5 relative, image per-pixel relative RMSE <= 1e-3 at equal spp."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

T_REL_TOL = 1e-5       # hit distance, relative
RMSE_TOL = 1e-3        # image relative RMSE


def rel_rmse(a, b):
    return float(np.sqrt(((a.astype(np.float64) - b) ** 2).mean()) / max(float(b.mean()), 1e-20))


@pytest.fixture(scope="module")
def pair(P):
    import oracle_api
    cache = {}

    def get(name, text, spectrum="rgb"):
        if name not in cache:
            hs = P.HostScene.parse_string(text, spectrum=spectrum)
            cache[name] = (hs, P.Scene(hs, 0), oracle_api.OracleScene(hs))
        return cache[name]
    yield get
    for hs, sc, orc in cache.values():
        sc.close()
        orc.close()


def _samples(hs, n, spp, seed=1):
    sb = hs.desc.film.sample_bounds
    rng = np.random.RandomState(seed)
    pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
    return pix, rng.randint(0, spp, n).astype(np.int32)


@pytest.mark.parametrize("sampler", ["halton", "sobol"])
def test_sampler_stream_bit_exact(P, pair, sampler):
    hs, sc, orc = pair("cornell_" + sampler, P.scenes.cornell_box(res=200, spp=64, sampler=sampler))
    pix, sn = _samples(hs, 6000, 64)
    # include the negative sample-bounds pixels and the corners (halton.rs:22-29 math_mod)
    pix[:4] = [[-1, -1], [200, 200], [-1, 200], [0, 0]]
    for dim0, nd in [(0, 40), (200, 8), (1000, 24)]:
        a, b = sc.sampler_values(pix, sn, dim0, nd), orc.sampler_values(pix, sn, dim0, nd)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (sampler, dim0)
        assert (a >= 0).all() and (a < 1).all()


def test_camera_rays_bit_exact(P, pair):
    hs, sc, orc = pair("cornell_halton", P.scenes.cornell_box(res=200, spp=64))
    pix, sn = _samples(hs, 5000, 64)
    ra, pfa = sc.camera_rays(pix, sn)
    rb, pfb = orc.camera_rays(pix, sn)
    assert np.array_equal(ra.view(np.uint32), rb.view(np.uint32))
    assert np.array_equal(pfa.view(np.uint32), pfb.view(np.uint32))


def _check_intersections(sc, orc, rays):
    pa, ta, ba, na, tra = sc.intersect(rays, counters=True)
    pb, tb, bb, nb, trb = orc.intersect(rays)
    assert np.array_equal(pa, pb), "%d primitive ids differ" % int((pa != pb).sum())
    hit = pb >= 0
    assert np.all(np.abs(ta[hit] - tb[hit]) <= T_REL_TOL * np.abs(tb[hit]))
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32))            # in fact bit-exact
    assert np.array_equal(ba.view(np.uint32), bb.view(np.uint32))
    assert (na, tra) == (nb, trb), "traversal visits differ: nodes %d/%d tris %d/%d" % (na, nb, tra, trb)
    return hit.mean()


def _random_rays(orc, n, seed, tmax=np.inf):
    wb = orc.world_bound()
    rng = np.random.RandomState(seed)
    o = rng.uniform(wb[:3] - 0.1 * (wb[3:] - wb[:3]), wb[3:] + 0.1 * (wb[3:] - wb[:3]), (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # axis-aligned and zero-component directions exercise the inf/NaN slab paths
    d[:64] = np.eye(3)[rng.randint(0, 3, 64)] * rng.choice([-1.0, 1.0], (64, 1))
    d[64:128, 1] = 0.0
    t = np.full((n, 1), tmax) if np.isscalar(tmax) else tmax.reshape(n, 1)
    return np.concatenate([o, t, d, np.zeros((n, 1))], 1).astype(np.float32)


@pytest.mark.parametrize("name", ["cornell", "grid", "grid_normals", "instances"])
def test_closest_hit_bit_exact(P, pair, name):
    texts = {"cornell": lambda: P.scenes.cornell_box(res=200, spp=64),
             "grid": lambda: P.scenes.displaced_grid(n=160, res=128, spp=4),
             "grid_normals": lambda: P.scenes.displaced_grid(n=96, res=128, spp=4, normals=True),
             "instances": lambda: P.scenes.instanced_blobs(n_instances=60, n_sub=2, res=(96, 64), spp=4, n_lights=8)}
    hs, sc, orc = pair(name, texts[name]())
    pix, sn = _samples(hs, 20000, 4)
    rays, _ = orc.camera_rays(pix, sn)
    frac = _check_intersections(sc, orc, rays)
    assert frac > 0.3
    _check_intersections(sc, orc, _random_rays(orc, 20000, 7))
    # any-hit agrees with the oracle, and with closest-hit where tmax is far beyond the hit
    r = _random_rays(orc, 20000, 9, tmax=np.random.RandomState(3).uniform(0.01, 50.0, 20000))
    assert np.array_equal(sc.intersect_p(r), orc.intersect_p(r))


LBVH = lambda t: t.replace('Accelerator "bvh"', 'Accelerator "bvh" "string layout" "lbvh"')


@pytest.mark.parametrize("name", ["cornell", "grid", "instances"])
def test_lbvh_layout_bit_exact(P, pair, name):
    """the binary LinearBVHNode layout (accel/lbvh.rs) next to the 4-wide one: primitive ids, t,
    barycentrics and the node / triangle visit counters equal the oracle's LBVHAccel restatement; the closest hits are
    the ones the 4-wide layout finds (same t; ids may differ only on exact ties, which the visit order decides)."""
    texts = {"cornell": lambda: P.scenes.cornell_box(res=200, spp=64),
             "grid": lambda: P.scenes.displaced_grid(n=160, res=128, spp=4),
             "instances": lambda: P.scenes.instanced_blobs(n_instances=60, n_sub=2, res=(96, 64), spp=4, n_lights=8)}
    txt = texts[name]()
    assert 'Accelerator "bvh"' in txt
    hs, sc, orc = pair("lbvh_" + name, LBVH(txt))
    assert hs.desc.accel.type == P.capi.PT_ACCEL_LBVH
    pix, sn = _samples(hs, 20000, 4)
    rays, _ = orc.camera_rays(pix, sn)
    assert _check_intersections(sc, orc, rays) > 0.3
    _check_intersections(sc, orc, _random_rays(orc, 20000, 7))
    r = _random_rays(orc, 20000, 9, tmax=np.random.RandomState(3).uniform(0.01, 50.0, 20000))
    assert np.array_equal(sc.intersect_p(r), orc.intersect_p(r))
    # against the 4-wide layout of the same scene
    hq, sq, oq = pair(name, txt)
    pa, ta, _ = sc.intersect(rays)
    pq, tq, _ = sq.intersect(rays)
    assert np.array_equal(ta.view(np.uint32), tq.view(np.uint32))
    assert (pa != pq).mean() < 1e-3
    assert np.array_equal(sc.intersect_p(r), sq.intersect_p(r))


def test_lbvh_layout_image(P, pair):
    hs, sc, orc = pair("lbvh_mixed", LBVH(P.scenes.cornell_box(res=96, spp=16, materials="mixed", extra_lights=True, maxdepth=8)))
    rgb, st = sc.render(collect_counters=1)
    ref, so = orc.render()
    assert rel_rmse(rgb, ref) <= RMSE_TOL
    for k in ("camera_rays", "regular_rays", "shadow_rays"):
        assert st[k] == so[k], (k, st[k], so[k])
    assert abs(st["nodes_visited"] - so["nodes_visited"]) <= 1e-3 * so["nodes_visited"]
    assert st["extend_bytes"] == 32 * st["nodes_visited"] + 48 * st["tris_tested"] + 48 * st["extend_rays"]


def test_mesh_tangents(P, pair):
    """Mesh "S" tangents (triangle.rs:386-391) with anisotropic metal / plastic: radiance per sample vs the oracle."""
    from test_hostcheck import tangent_scene
    hs, sc, orc = pair("tangents", tangent_scene(P, res=64, spp=8))
    pix, sn = _samples(hs, 20000, 8, seed=6)
    la = sc.li(pix, sn)
    lb, _, _, _ = orc.li(pix, sn)
    close = (np.abs(la - lb) <= 1e-3 * np.maximum(np.abs(lb), 1e-2)).all(1)
    assert close.mean() > 0.999, close.mean()
    rgb, _ = sc.render()
    ref, _ = orc.render()
    assert rel_rmse(rgb, ref) <= RMSE_TOL


def test_spot_lights(P, pair):
    """LightSource "spot" (lights/spot.rs): radiance per sample and image vs the oracle, spatial light sampling."""
    from test_hostcheck import spot_scene
    hs, sc, orc = pair("spots", spot_scene(P, res=64, spp=8))
    pix, sn = _samples(hs, 20000, 8, seed=8)
    la = sc.li(pix, sn)
    lb, _, _, _ = orc.li(pix, sn)
    close = (np.abs(la - lb) <= 1e-3 * np.maximum(np.abs(lb), 1e-2)).all(1)
    assert close.mean() > 0.999, close.mean()
    rgb, _ = sc.render()
    ref, _ = orc.render()
    assert rel_rmse(rgb, ref) <= RMSE_TOL
    wb = orc.world_bound()
    pts = np.random.RandomState(2).uniform(wb[:3], wb[3:], (100, 3)).astype(np.float32)
    assert np.allclose(sc.light_distribution(pts), orc.light_distribution(pts), rtol=2e-6, atol=0)


@pytest.mark.parametrize("target", ["distance", "depth", "n", "ns", "uv", "dpdus"])
def test_aov_integrator(P, pair, target):
    """Integrator "aov" (integrators/aov.rs): per-sample values bit-exact vs the oracle (no transcendental functions on
    this path), image RMSE, on a scene with normals / tangents and on an instanced one."""
    from test_hostcheck import aov_scene
    for instanced in (False, True):
        hs, sc, orc = pair("aov_%s_%d" % (target, instanced), aov_scene(P, target, 0.002 if target in ("distance", "depth") else 0.75, res=72, spp=4, instanced=instanced))
        pix, sn = _samples(hs, 20000, 4, seed=12)
        la = sc.li(pix, sn)
        lb = orc.li(pix, sn)[0]
        assert np.array_equal(la.view(np.uint32), lb.view(np.uint32)), np.abs(la - lb).max()
        rgb, st = sc.render()
        ref, _ = orc.render()
        assert rel_rmse(rgb, ref) <= 1e-5
        assert st["shadow_rays"] == 0 and st["regular_rays"] == st["camera_rays"]


def test_hlbvh_split_method(P, pair):
    """Accelerator "bvh" "string splitmethod" ["hlbvh"] (build/hlbvh.rs): hits bit-exact and visit counters equal to the oracle's
    on the Morton-treelet tree, with instances."""
    hl = lambda t: t.replace('Accelerator "bvh"', 'Accelerator "bvh" "string splitmethod" ["hlbvh"]')
    for name, txt in (("hl_grid", P.scenes.displaced_grid(n=120, res=64, spp=4)), ("hl_inst", P.scenes.instanced_blobs(n_instances=60, n_sub=2, res=(96, 64), spp=4, n_lights=8))):
        hs, sc, orc = pair(name, hl(txt))
        rays = _random_rays(orc, 30000, 31)
        _check_intersections(sc, orc, rays)
        rgb, st = sc.render(collect_counters=1)
        ref, so = orc.render()
        assert rel_rmse(rgb, ref) <= RMSE_TOL
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested"):
            assert st[k] == so[k], (name, k)


def test_ao_integrator(P):
    """Integrator "ambientocclusion" (integrators/ao.rs): per-sample values bit for bit vs the oracle (cosine and uniform
    hemisphere, halton and sobol sample arrays, instances, a surface without a material), image, ray / visit counters, and
    the pipelined waves against one wave at a time."""
    from test_hostcheck import check_ao, check_ao_random
    check_ao(P, n=20000, res=72, spp=4, wave_paths=3000, threads=0)
    check_ao_random(P, res=72, spp=3, threads=0)


def test_tree_integrators(P):
    """Integrators "directlighting" (one / all with sample arrays / arrays used up) and "whitted" (core/integrator/
    sampler.rs:37-160 specular_reflect / specular_transmit as a depth-first tree walk): per-sample radiance within 2e-5 of
    the oracle's recursion, equal ray / BVH-visit / vertex counters, image, pipelined waves against one wave at a time."""
    from test_hostcheck import check_tree, check_tree_random
    check_tree(P, n=20000, res=72, spp=4, wave_paths=3000, threads=0)
    check_tree_random(P, res=72, spp=2, threads=0)


def test_infinite_light(P, tmp_path):
    """LightSource "infinite" (lights/infinite.rs:108-185; Distribution2D, MIPMap level 0): constant and mapped environment,
    path / directlighting / whitted, spatial and power light distributions.  Same rays as the oracle (equal counters); acos /
    atan2 are CUDA's on the device, so the lookup coordinates and MIS pdfs differ by a few ulp: per-sample rtol 1e-4."""
    from test_hostcheck import check_env
    check_env(P, tmp_path, n=20000, res=72, spp=4, wave_paths=3000, threads=0, rtol=1e-4, frac=0.9995)


def test_extra_materials(P):
    """Materials uber / substrate / translucent (their own shade class with five-lobe BSDFs; FresnelBlend and
    LambertianTransmission lobes) under path / directlighting / whitted: radiance within 2e-5 on >= 99.9 % of the samples,
    equal ray / visit / vertex counters."""
    from test_hostcheck import check_extra_materials
    check_extra_materials(P, n=20000, res=72, spp=4, wave_paths=3000, threads=0)


def test_mix_material(P):
    """Material "mix" (ScaledBxDF lobes of two named materials in the generic five-lobe class) under path / directlighting /
    whitted: radiance within 2e-5 on >= 99.9 % of the samples, equal ray / visit / vertex counters."""
    from test_hostcheck import check_mix
    check_mix(P, n=20000, res=72, spp=4, wave_paths=3000, threads=0)


def test_stratified_sampler(P):
    """Sampler "stratified": per-pixel shuffled tables, the tile's PCG32 stream after them, Latin-hypercube sample arrays
    (path / ambientocclusion / directlighting all / whitted): image and counters equal to the oracle's."""
    from test_hostcheck import check_stratified
    check_stratified(P, res=72, threads=0)


def test_zerotwo_sampler(P):
    """Sampler "02sequence" / "lowdiscrepancy": scrambled (0,2)-sequence tables and sample arrays drawn on the device in the
    reference's order: image and counters equal to the oracle's."""
    from test_hostcheck import check_zerotwo
    check_zerotwo(P, res=72, threads=0)


def test_map_lights(P, tmp_path):
    """LightSource "projection" / "goniometric": per-sample radiance against the oracle (acos / atan2 are CUDA's: 99.5 % of the
    samples within 1e-4), equal counters, image."""
    from test_hostcheck import check_map_lights
    check_map_lights(P, tmp_path, n=20000, res=72, spp=4, wave_paths=3000, threads=0, rtol=1e-4, frac=0.995)


def test_cameras(P):
    """Camera "orthographic" (with and without a lens) and "environment": camera rays bit for bit vs the oracle, per-sample
    radiance, image and counters (halton / sobol / random samplers)."""
    from test_hostcheck import check_cameras
    check_cameras(P, n=20000, res=72, spp=4, threads=0)


def test_one_sided_triangles_bit_exact(P, pair):
    """Back-face culling of one-sided meshes (triangle.rs:236-246), with and without ReverseOrientation."""
    from test_hostcheck import one_sided_scene
    hs, sc, orc = pair("one_sided", one_sided_scene(P, res=64, spp=4))
    rays = _random_rays(orc, 30000, 21)
    _check_intersections(sc, orc, rays)
    r = _random_rays(orc, 30000, 22, tmax=np.random.RandomState(5).uniform(1.0, 600.0, 30000))
    assert np.array_equal(sc.intersect_p(r), orc.intersect_p(r))
    rgb, _ = sc.render()
    ref, _ = orc.render()
    assert rel_rmse(rgb, ref) <= RMSE_TOL


def test_light_distribution_bit_exact(P, pair):
    for name, txt in [("cornell_halton", P.scenes.cornell_box(res=200, spp=64)), ("cornell_lights", P.scenes.cornell_box(res=64, spp=4, extra_lights=True))]:
        hs, sc, orc = pair(name, txt)
        wb = orc.world_bound()
        pts = np.random.RandomState(2).uniform(wb[:3], wb[3:], (200, 3)).astype(np.float32)
        a, b = sc.light_distribution(pts), orc.light_distribution(pts)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (np.abs(a - b) / np.abs(b)).max()


@pytest.mark.parametrize("name", ["cornell", "mixed", "grid", "instances"])
def test_li_matches_oracle(P, pair, name):
    texts = {"cornell": lambda: P.scenes.cornell_box(res=200, spp=64),
             "mixed": lambda: P.scenes.cornell_box(res=96, spp=16, materials="mixed", extra_lights=True, maxdepth=8),
             "grid": lambda: P.scenes.displaced_grid(n=160, res=128, spp=4),
             "instances": lambda: P.scenes.instanced_blobs(n_instances=60, n_sub=2, res=(96, 64), spp=4, n_lights=8)}
    hs, sc, orc = pair(name if name != "cornell" else "cornell_halton", texts[name]())
    pix, sn = _samples(hs, 20000, hs.desc.sampler.samples_per_pixel, seed=4)
    la = sc.li(pix, sn)
    lb, _, _, _ = orc.li(pix, sn)
    close = (np.abs(la - lb) <= 1e-3 * np.maximum(np.abs(lb), 1e-2)).all(1)
    # transcendental ulps (CUDA vs libm) may flip a discrete decision on a rare path; the estimator must agree otherwise
    assert close.mean() > 0.999, close.mean()
    assert abs(la.sum() - lb.sum()) <= 2e-3 * lb.sum()


def test_image_rmse_cornell_config1(P, pair):
    """BASELINE.json configs[0]: Cornell box 256x256, maxdepth 5, 16 spp Halton, matte + DiffuseAreaLight, BVH."""
    hs, sc, orc = pair("config1", P.scenes.cornell_box(res=256, spp=16))
    rgb, st = sc.render(collect_counters=1)
    ref, so = orc.render()
    assert rel_rmse(rgb, ref) <= RMSE_TOL
    for k in ("camera_rays", "regular_rays", "shadow_rays"):
        assert st[k] == so[k], (k, st[k], so[k])
    assert abs(st["nodes_visited"] - so["nodes_visited"]) <= 1e-4 * so["nodes_visited"]
    assert st["kernel_launches"] > 0
    golden = os.path.join(os.path.dirname(__file__), "golden", "cornell_64_4spp.npy")
    if os.path.exists(golden):
        hs2, sc2, _ = pair("golden", P.scenes.cornell_box(res=64, spp=4))
        assert rel_rmse(sc2.render()[0], np.load(golden)) <= RMSE_TOL


@pytest.mark.parametrize("name", ["mixed", "grid", "sobol", "gaussian", "instances"])
def test_image_rmse_other_scenes(P, pair, name):
    texts = {"mixed": lambda: P.scenes.cornell_box(res=96, spp=16, materials="mixed", extra_lights=True, maxdepth=8),
             "grid": lambda: P.scenes.displaced_grid(n=160, res=128, spp=4),
             "sobol": lambda: P.scenes.cornell_box(res=100, spp=8, sampler="sobol"),
             "gaussian": lambda: P.scenes.cornell_box(res=80, spp=4).replace('Sampler', 'PixelFilter "gaussian" "float xwidth" [1.5] "float ywidth" [1.5]\nSampler'),
             "instances": lambda: P.scenes.instanced_blobs(n_instances=60, n_sub=2, res=(96, 64), spp=4, n_lights=8)}
    hs, sc, orc = pair(name, texts[name]())
    rgb, st = sc.render()
    ref, so = orc.render()
    assert np.isfinite(rgb).all()
    assert rel_rmse(rgb, ref) <= RMSE_TOL, rel_rmse(rgb, ref)


def test_image_rmse_spectral(P, pair):
    """configs[3] at test size: 60-bin SampledSpectrum Cornell box."""
    hs, sc, orc = pair("spectral", P.scenes.cornell_box(res=48, spp=4), spectrum="sampled")
    assert hs.desc.n_channels == 60
    rgb, _ = sc.render()
    ref, _ = orc.render()
    assert rel_rmse(rgb, ref) <= RMSE_TOL


def test_sharding_properties(P, pair):
    """Size-independent properties: rendering is deterministic; spp shards and tile shards partition the work, so the
    sum of the partial films equals the full film (the multi-GPU reduce), at any wave size."""
    hs, sc, orc = pair("config1", P.scenes.cornell_box(res=256, spp=16))
    full, st = sc.render()
    again, _ = sc.render(wave_paths=100000)
    assert np.array_equal(full, again) or rel_rmse(again, full) < 1e-5   # wave size only changes the float summation order
    import ctypes as C
    parts = []
    for r in range(4):
        ptr, n, _ = sc.render_partial(spp_begin=4 * r, spp_end=4 * r + 4)
        import torch
        from pbrt_r3_b200 import dist as D
        parts.append(D.film_tensor(ptr, n, "cuda:0").clone())
    total = sum(parts)
    rgb = sc.film_resolve(total.data_ptr())
    assert rel_rmse(rgb, full) < 1e-5
    parts = []
    for r in range(3):
        ptr, n, s = sc.render_partial(shard_index=r, shard_count=3)
        parts.append(D.film_tensor(ptr, n, "cuda:0").clone())
    rgb = sc.film_resolve(sum(parts).data_ptr())
    assert rel_rmse(rgb, full) < 1e-5


def test_wave_pipelining_is_invisible(P, pair):
    """render_waves runs up to 8 waves at once in segments of the path state (pt_render_opts.wave_segments): paths do not
    interact and the film gathers the waves in order, so the image is bit-identical and the ray / node / triangle
    counters are equal for any number of waves in flight (mixed materials: paths of very different lengths)."""
    hs, sc, orc = pair("mixed_pipe", P.scenes.cornell_box(res=96, spp=8, materials="mixed", extra_lights=True, maxdepth=8))
    base, st1 = sc.render(wave_paths=20000, wave_segments=1, collect_counters=1)
    for k in (2, 4, 8):
        rgb, st = sc.render(wave_paths=20000, wave_segments=k, collect_counters=1)
        assert np.array_equal(rgb, base), k
        for key in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert st[key] == st1[key], (k, key)
        assert st["extend_launches"] < st1["extend_launches"]
    ref, _ = orc.render()
    assert rel_rmse(base, ref) <= RMSE_TOL


def test_full_size_config2_properties(P):
    """configs[1] geometry at full size (1M triangles): BVH build on the host, primary hits vs the oracle on a sample of
    rays, closest-hit/any-hit consistency and determinism on the whole primary wave."""
    import oracle_api
    hs = P.HostScene.parse_string(P.scenes.displaced_grid(n=708, res=256, spp=1))
    assert hs.desc.n_tris > 990000
    sc = P.Scene(hs, 0)
    orc = oracle_api.OracleScene(hs)
    pix, sn = _samples(hs, 50000, 1)
    rays, _ = orc.camera_rays(pix, sn)
    _check_intersections(sc, orc, rays)
    _check_intersections(sc, orc, _random_rays(orc, 50000, 11))
    # property: a ray stopped just beyond its closest hit is occluded; stopped well before it is not
    p, t, _ = sc.intersect(rays)
    hit = p >= 0
    r2 = rays[hit].copy()
    r2[:, 3] = t[hit] * 1.001
    assert sc.intersect_p(r2).all()
    r2[:, 3] = t[hit] * 0.5
    occ = sc.intersect_p(r2)
    p3, t3, _ = sc.intersect(r2)
    assert np.array_equal(occ.astype(bool), p3 >= 0)
    a, _ = sc.render()
    b, _ = sc.render()
    assert np.array_equal(a, b)
    sc.close()
    orc.close()


def test_full_size_config3_properties(P):
    """configs[2] geometry at full size: 1000 instances of a 20 480-triangle blob (20.5 M instanced triangles), 256 emissive
    quads, two-level 4-wide BVH.  Primary hits vs the oracle on a sample of rays (ids, t, barycentrics, visit counters),
    closest-hit / any-hit consistency, per-sample radiance vs the oracle, determinism."""
    import oracle_api
    hs = P.HostScene.parse_string(P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(240, 135), spp=2, maxdepth=8, n_lights=256))
    assert hs.desc.n_instances == 1000 and hs.desc.n_lights == 512
    sc = P.Scene(hs, 0)
    orc = oracle_api.OracleScene(hs)
    pix, sn = _samples(hs, 30000, 2)
    rays, _ = orc.camera_rays(pix, sn)
    assert _check_intersections(sc, orc, rays) > 0.2
    _check_intersections(sc, orc, _random_rays(orc, 30000, 13))
    p, t, _ = sc.intersect(rays)
    hit = p >= 0
    r2 = rays[hit].copy()
    r2[:, 3] = t[hit] * 1.001
    assert sc.intersect_p(r2).all()
    r2[:, 3] = t[hit] * 0.5
    assert np.array_equal(sc.intersect_p(r2).astype(bool), sc.intersect(r2)[0] >= 0)
    la = sc.li(pix[:8000], sn[:8000])
    lb = orc.li(pix[:8000], sn[:8000])[0]
    close = (np.abs(la - lb) <= 1e-3 * np.maximum(np.abs(lb), 1e-2)).all(1)
    assert close.mean() > 0.999, close.mean()
    a, _ = sc.render()
    assert np.array_equal(a, sc.render(wave_segments=1)[0])
    sc.close()
    orc.close()


def test_edge_cases(P):
    # empty scene (no primitives, no lights) renders black; scene without lights renders black; zero rays is fine
    hs = P.HostScene.parse_string('Camera "perspective"\nFilm "image" "integer xresolution" [8] "integer yresolution" [8]\nWorldBegin\nWorldEnd\n')
    sc = P.Scene(hs, 0)
    rgb, _ = sc.render()
    assert rgb.shape == (8, 8, 3) and not rgb.any()
    assert sc.intersect(np.zeros((0, 8), np.float32))[0].shape == (0,)
    r = np.array([[0, 0, 0, np.inf, 0, 0, 1, 0]], np.float32)
    assert sc.intersect(r)[0][0] == -1 and sc.intersect_p(r)[0] == 0
    sc.close()
    txt = P.scenes.cornell_box(res=16, spp=1)
    txt = txt.replace('AreaLightSource "diffuse" "rgb L" [17 12 4]\n', "")
    hs = P.HostScene.parse_string(txt)
    sc = P.Scene(hs, 0)
    assert not sc.render()[0].any()
    sc.close()
    with pytest.raises(P.PbrtError):   # the random sampler has no (pixel, sample number) -> value function
        P.Scene(P.HostScene.parse_string(P.scenes.cornell_box(res=16, spp=1, sampler="random")), 0).li(np.zeros((1, 2), np.int32), np.zeros(1, np.int32))


def test_image_random_sampler(P, pair):
    """Sampler "random" (samplers/random.rs): per-tile sequential PCG32 streams reproduced; image and ray counts vs the oracle."""
    hs, sc, orc = pair("random", P.scenes.cornell_box(res=72, spp=4, sampler="random"))
    rgb, st = sc.render(collect_counters=1)
    ref, so = orc.render()
    assert rel_rmse(rgb, ref) <= RMSE_TOL, rel_rmse(rgb, ref)
    for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
        assert st[k] == so[k], (k, st[k], so[k])      # the streams are reproduced: every path draws the same numbers
    assert np.array_equal(rgb, sc.render()[0])
