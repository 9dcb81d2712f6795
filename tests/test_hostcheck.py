"""Logic check of the CUDA sources on a CPU-only box (TEST INFRASTRUCTURE): the product's device_api.cu and *.cuh are
compiled by g++ against tests/hostcheck/cuda_shim.h (device memory = malloc, a launch = one serial call) into
tests/hostcheck/_build/, loaded through the same ctypes wrappers, and compared with the oracle.  This catches stage
logic errors (queues, pending MIS terms, instance transforms, film gather) before GPU time is spent.  The product
library never contains this build and refuses to run without a CUDA device."""
import os
import subprocess

import numpy as np
import pytest

# the HC fixture (the hostcheck build loaded through the same ctypes wrappers) lives in conftest.py


def _cmp_li(P, text, n, spp, tol_frac):
    import oracle_api
    hs = P.HostScene.parse_string(text)
    sc = P.Scene(hs)
    orc = oracle_api.OracleScene(hs)
    sb = hs.desc.film.sample_bounds
    rng = np.random.RandomState(4)
    pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
    sn = rng.randint(0, spp, n).astype(np.int32)
    la = sc.li(pix, sn)
    lb = orc.li(pix, sn)[0]
    close = (np.abs(la - lb) <= 1e-4 * np.maximum(np.abs(lb), 1e-3)).all(1)
    assert close.mean() >= tol_frac, close.mean()
    return hs, sc, orc


def test_hostcheck_cornell_image_and_streams(HC):
    P = HC
    import oracle_api
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=24, spp=4).replace('Accelerator "bvh"', 'Accelerator "bvh"'))
    sc = P.Scene(hs)
    orc = oracle_api.OracleScene(hs)
    rgb, st = sc.render(collect_counters=1, wave_paths=1000)     # several waves and pixel chunks
    ref, so = orc.render(threads=2)
    assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5
    for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
        assert st[k] == so[k], k
    # waves in flight (software-pipelined bounce loop): same image bit for bit, same counters, fewer launches
    one, st1 = sc.render(collect_counters=1, wave_paths=1000, wave_segments=1)
    for k in (2, 3, 8):
        many, stk = sc.render(collect_counters=1, wave_paths=1000, wave_segments=k)
        assert np.array_equal(one, many), k
        for key in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert stk[key] == st1[key], (k, key)
        assert stk["extend_launches"] < st1["extend_launches"]
    # shards add up
    a, _ = sc.render(wave_paths=3000)
    import ctypes as C
    acc = np.zeros(24 * 24 * 4, np.float32)
    for r in range(2):
        ptr, n, _ = sc.render_partial(spp_begin=2 * r, spp_end=2 * r + 2)
        acc += np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), (n * 4,))
    b = P.dist.resolve_host(acc).reshape(24, 24, 3) if hasattr(P, "dist") else None
    from pbrt_r3_b200 import dist as D
    b = D.resolve_host(acc).reshape(24, 24, 3)
    assert np.sqrt(((a - b) ** 2).mean()) / a.mean() < 1e-5


def test_hostcheck_halton_digit_paths(HC):
    """scrambled_radical_inverse has a 32-bit path (index * base < 2^32: one mul-hi per digit), a 32-bit-index path and the
    u64 loop; all three against the oracle's u64 restatement, bit for bit, around the switch points."""
    import oracle_api
    P = HC
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=200, spp=1 << 20))
    sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
    rng = np.random.RandomState(9)
    n = 400
    pix = np.stack([rng.randint(-1, 201, n), rng.randint(-1, 201, n)], 1).astype(np.int32)
    for hi in (64, 1 << 12, 1 << 17, (1 << 20) - 1):        # index up to 2^20 * 62208 = 2^36: 64-bit indices too
        sn = rng.randint(0, hi, n).astype(np.int32)
        sn[:4] = [0, 1, hi - 1, hi // 2]
        for dim0, nd in [(0, 24), (180, 12), (1000, 24)]:
            a, b = sc.sampler_values(pix, sn, dim0, nd), orc.sampler_values(pix, sn, dim0, nd)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (hi, dim0)


def one_sided_scene(P, res=32, spp=2):
    """Cornell box with one-sided meshes (pbrt-r3's "twosided" false: back faces are culled, triangle.rs:236-246), every
    other one with ReverseOrientation."""
    lines, k = [], 0
    for l in P.scenes.cornell_box(res=res, spp=spp, materials="mixed").split("\n"):
        if l.startswith("Shape"):
            k += 1
            l = l + ' "bool twosided" ["false"]'
            if k % 2 == 0:
                l = "AttributeBegin\nReverseOrientation\n" + l + "\nAttributeEnd"
        lines.append(l)
    return "\n".join(lines)


def test_hostcheck_one_sided_triangles(HC):
    import oracle_api
    P = HC
    hs = P.HostScene.parse_string(one_sided_scene(P))
    flags = np.ctypeslib.as_array(hs.desc.tri_flags, (hs.desc.n_tris,))
    assert not (flags & P.capi.PT_TRI_TWO_SIDED).any() and (flags & P.capi.PT_TRI_REVERSE_ORIENTATION).any()
    sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
    wb = orc.world_bound()
    rng = np.random.RandomState(3)
    n = 6000
    o = rng.uniform(wb[:3], wb[3:], (n, 3))
    d = rng.normal(size=(n, 3))
    d[:60] = np.eye(3)[rng.randint(0, 3, 60)] * rng.choice([-1.0, 1.0], (60, 1))     # each dominant axis, exactly
    rays = np.concatenate([o, np.full((n, 1), np.inf), d, np.zeros((n, 1))], 1).astype(np.float32)
    pa, ta, ba, na, tra = sc.intersect(rays, counters=True)
    pb, tb, bb, nb, trb = orc.intersect(rays)
    assert np.array_equal(pa, pb) and np.array_equal(ta.view(np.uint32), tb.view(np.uint32)) and np.array_equal(ba.view(np.uint32), bb.view(np.uint32))
    assert (na, tra) == (nb, trb) and 0.2 < (pa >= 0).mean() < 0.99
    r2 = rays.copy()
    r2[:, 3] = rng.uniform(1.0, 600.0, n)
    assert np.array_equal(sc.intersect_p(r2), orc.intersect_p(r2))


def test_hostcheck_li_mixed_materials(HC):
    _cmp_li(HC, HC.scenes.cornell_box(res=32, spp=8, materials="mixed", extra_lights=True, maxdepth=8), 3000, 8, 0.999)


def test_hostcheck_li_is_bit_level(HC):
    """With the same libm on both sides, the stage code and the oracle agree to the last bit on almost every sample; the
    rest differ by one rounding (beta * (a + b) / p evaluated as beta * a / p + beta * b / p when both MIS halves are
    non-zero).  Guards arithmetic-preserving refactors of the BSDF code (fused f / pdf evaluation)."""
    import oracle_api
    P = HC
    for txt, spp in ((P.scenes.cornell_box(res=32, spp=8, materials="mixed", extra_lights=True, maxdepth=8), 8), (tangent_scene(P), 8),
                     (P.scenes.displaced_grid(n=40, res=32, spp=4, maxdepth=8), 4)):
        hs = P.HostScene.parse_string(txt)
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        sb = hs.desc.film.sample_bounds
        rng = np.random.RandomState(4)
        n = 3000
        pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
        sn = rng.randint(0, spp, n).astype(np.int32)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        same = (la.view(np.uint32) == lb.view(np.uint32)).all(1)
        assert same.mean() >= 0.98, same.mean()
        assert (np.abs(la - lb) <= 1e-6 * np.maximum(np.abs(lb), 1e-6)).all()


def test_hostcheck_li_instances(HC):
    _cmp_li(HC, HC.scenes.instanced_blobs(n_instances=30, n_sub=1, res=(48, 32), spp=4, n_lights=4), 3000, 4, 0.999)


def tangent_scene(P, res=32, spp=8):
    """Cornell box whose floor is an anisotropic copper sheet with per-vertex tangents "S", and whose back wall has both
    "N" and "S" (triangle.rs:376-444: the shading frame is built from S when the mesh has tangents, N or not)."""
    txt = P.scenes.cornell_box(res=res, spp=spp, maxdepth=6)
    shapes = [l for l in txt.split("\n") if l.startswith("Shape")]
    floor, back = shapes[0], shapes[2]
    txt = txt.replace('Material "matte" "rgb Kd" [0.725 0.71 0.68]\n' + floor,
                      'Material "metal" "float uroughness" [0.05] "float vroughness" [0.4]\n' + floor + ' "point S" [1 0 1  1 0 0.5  0.2 0 1  -1 0 1]')
    txt = txt.replace('Material "matte" "rgb Kd" [0.725 0.71 0.68]\n' + back,
                      'Material "plastic" "rgb Kd" [0.3 0.3 0.6] "rgb Ks" [0.5 0.5 0.5] "float uroughness" [0.02] "float vroughness" [0.3]\n' + back +
                      ' "point S" [1 1 0  0 1 0  -1 1 0  1 0.5 0] "normal N" [0 0.1 -1  0.1 0 -1  0 -0.1 -1  -0.1 0 -1]')
    assert '"point S"' in txt and txt.count('"point S"') == 2
    return txt


def test_hostcheck_mesh_tangents(HC):
    hs, sc, orc = _cmp_li(HC, tangent_scene(HC), 3000, 8, 0.999)
    flags = np.ctypeslib.as_array(hs.desc.tri_flags, (hs.desc.n_tris,))
    assert ((flags & HC.capi.PT_TRI_HAS_S) != 0).sum() == 4 and bool(hs.desc.tri_s)
    # the tangents change the result: same scene without them
    plain = HC.HostScene.parse_string(__import__("re").sub(r' "point S" \[[^\]]*\]', "", tangent_scene(HC)))
    pix = np.array([[16, 28], [10, 30], [16, 10]], np.int32)
    assert not np.array_equal(sc.li(pix, np.zeros(3, np.int32)), HC.Scene(plain).li(pix, np.zeros(3, np.int32)))


def spot_scene(P, res=32, spp=8):
    """Cornell box lit by two spot lights (one with a wide falloff band, one narrow, aimed at a wall) next to the area light."""
    txt = P.scenes.cornell_box(res=res, spp=spp, maxdepth=5, materials="mixed")
    spots = ('LightSource "spot" "rgb I" [400000 380000 300000] "point from" [278 540 280] "point to" [200 0 300] "float coneangle" [35] "float conedeltaangle" [20]\n'
             'LightSource "spot" "rgb I" [200000 100000 100000] "point from" [100 300 100] "point to" [556 200 400] "float coneangle" [15] "float conedeltaangle" [2]\n')
    assert "WorldBegin\n" in txt
    return txt.replace("WorldBegin\n", "WorldBegin\n" + spots, 1)


def test_hostcheck_spot_lights(HC):
    hs, sc, orc = _cmp_li(HC, spot_scene(HC), 3000, 8, 0.999)
    assert hs.desc.n_spots == 2
    rgb, st = sc.render(collect_counters=1)
    ref, so = orc.render(threads=2)
    assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5
    for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
        assert st[k] == so[k], k
    # the "power" light strategy uses SpotLight::power (spot.rs:82-85)
    hp = HC.HostScene.parse_string(spot_scene(HC).replace('Integrator "path"', 'Integrator "path" "string lightsamplestrategy" ["power"]'))
    a, b = HC.Scene(hp).render()[0], __import__("oracle_api").OracleScene(hp).render(threads=2)[0]
    assert np.sqrt(((a - b) ** 2).mean()) / b.mean() < 1e-5


AOV_TARGETS = ["distance", "depth", "n", "ns", "uv", "dpdus"]


def aov_scene(P, target, scale=1.0, res=24, spp=2, instanced=False):
    txt = P.scenes.instanced_blobs(n_instances=20, n_sub=1, res=(res, res * 2 // 3), spp=spp, n_lights=2) if instanced else tangent_scene(P, res=res, spp=spp)
    assert 'Integrator "path"' in txt
    head, rest = txt.split('Integrator "path"', 1)
    return head + 'Integrator "aov" "string target" ["%s"] "float scale" [%g]' % (target, scale) + rest[rest.index("\n"):]


@pytest.mark.parametrize("target", AOV_TARGETS)
def test_hostcheck_aov_integrator(HC, target):
    """Integrator "aov" (integrators/aov.rs): first-hit quantities as colour, same sampler / camera / film as the path
    integrator; per-sample values bit-exact vs the oracle, images equal, one closest-hit query per camera sample."""
    import oracle_api
    P = HC
    for instanced in (False, True):
        hs = P.HostScene.parse_string(aov_scene(P, target, 0.002 if target in ("distance", "depth") else 0.75, instanced=instanced))
        assert hs.desc.integrator.type == P.capi.PT_INTEGRATOR_AOV
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        sb = hs.desc.film.sample_bounds
        rng = np.random.RandomState(5)
        n = 1500
        pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
        sn = rng.randint(0, 2, n).astype(np.int32)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        assert np.array_equal(la.view(np.uint32), lb.view(np.uint32)), (target, instanced, np.abs(la - lb).max())
        assert la.max() > 0
        rgb, st = sc.render(collect_counters=1, wave_paths=300)
        ref, so = orc.render(threads=2)
        assert np.sqrt(((rgb - ref) ** 2).mean()) / max(ref.mean(), 1e-9) < 1e-5
        assert st["regular_rays"] == st["camera_rays"] == so["camera_rays"] and st["shadow_rays"] == 0
        for k in ("regular_rays", "nodes_visited", "tris_tested"):
            assert st[k] == so[k], k


def ao_scene(P, nsamples=8, cos=True, res=24, spp=2, instanced=False, sampler="halton"):
    """Integrator "ambientocclusion" on the tangent scene (meshes with N / S: the AO frame is the TRUE geometry's, not the
    shading frame) whose short block has no material (the camera ray passes through, ao.rs:68-71), or on instances."""
    if instanced:
        txt = P.scenes.instanced_blobs(n_instances=20, n_sub=1, res=(res, res * 2 // 3), spp=spp, n_lights=2, sampler=sampler)
    else:
        txt = tangent_scene(P, res=res, spp=spp).replace('Sampler "halton"', 'Sampler "%s"' % sampler)
        shapes = [l for l in txt.split("\n") if l.startswith("Shape")]
        white = 'Material "matte" "rgb Kd" [0.725 0.71 0.68]\n'
        assert white + shapes[5] in txt
        txt = txt.replace(white + shapes[5], 'Material "none"\n' + shapes[5])
    assert 'Integrator "path"' in txt and 'Sampler "%s"' % sampler in txt
    head, rest = txt.split('Integrator "path"', 1)
    return head + 'Integrator "ambientocclusion" "integer nsamples" [%d] "bool cossample" ["%s"]' % (nsamples, "true" if cos else "false") + rest[rest.index("\n"):]


def check_ao(P, n=1200, res=24, spp=2, wave_paths=300, threads=2):
    """per-sample values bit for bit vs the oracle (sin / cos are glibc-exact on the device), image, ray and visit counters;
    the wave pipelining is invisible"""
    import oracle_api
    for kw in (dict(), dict(cos=False, nsamples=5), dict(instanced=True, nsamples=3), dict(sampler="sobol", nsamples=4, spp=3)):
        kw = dict(dict(res=res, spp=spp), **kw)
        hs = P.HostScene.parse_string(ao_scene(P, **kw))
        I = hs.desc.integrator
        assert (I.type, I.ao_samples, I.ao_cos_sample) == (P.capi.PT_INTEGRATOR_AO, kw.get("nsamples", 8), int(kw.get("cos", True)))
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        sb = hs.desc.film.sample_bounds
        rng = np.random.RandomState(6)
        pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
        sn = rng.randint(0, kw["spp"], n).astype(np.int32)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        assert np.array_equal(la.view(np.uint32), lb.view(np.uint32)), (kw, np.abs(la - lb).max())
        assert la.max() > 0 and (la == 0).all(1).any()
        rgb, st = sc.render(collect_counters=1, wave_paths=wave_paths)
        ref, so = orc.render(threads=threads)
        assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-6, kw
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested"):
            assert st[k] == so[k], (kw, k)
        assert st["shadow_rays"] > 0
        if not kw.get("instanced"):
            assert st["regular_rays"] > st["camera_rays"]      # rays that went through the block without a material
        one, _ = sc.render(wave_paths=wave_paths, wave_segments=1)
        assert np.array_equal(one, rgb), kw


def test_hostcheck_ao_integrator(HC):
    """Integrator "ambientocclusion" (integrators/ao.rs): the sampler's 2D sample array (dimensions 5-6 of sample NUMBER
    n * s + i), cosine / uniform hemisphere directions, one any-hit query each."""
    check_ao(HC)


def tree_scene(P, integrator, res=24, spp=2, maxdepth=5, nsamples=1, sampler="halton", instanced=False):
    """Integrators "directlighting" / "whitted" on the Cornell box with a glass and a metal block, a plastic floor and a
    mirror wall (the specular reflection / transmission tree branches at the glass), area + point + distant lights; or on
    the instanced scene."""
    if instanced:
        txt = P.scenes.instanced_blobs(n_instances=20, n_sub=1, res=(res, res * 2 // 3), spp=spp, n_lights=2, sampler=sampler)
    else:
        txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", extra_lights=True, sampler=sampler)
        txt = txt.replace('AreaLightSource "diffuse" "rgb L" [17 12 4]', 'AreaLightSource "diffuse" "rgb L" [17 12 4] "integer nsamples" [%d]' % nsamples)
        shapes = [l for l in txt.split("\n") if l.startswith("Shape")]
        txt = txt.replace(shapes[1], 'Material "none"\n' + shapes[1].replace("548.8", "500"), 1)   # a sheet without a material below the ceiling
    head, rest = txt.split('Integrator "path"', 1)
    return head + 'Integrator %s "integer maxdepth" [%d]' % (integrator, maxdepth) + rest[rest.index("\n"):]


TREE_CASES = [('"directlighting" "string strategy" ["one"]', dict()),
              ('"directlighting"', dict(nsamples=3)),                       # "all": sample arrays, 3 samples of each area-light triangle
              ('"directlighting"', dict(maxdepth=2, sampler="sobol")),      # arrays run out below the first glass vertex: get_2d pairs
              ('"directlighting" "string strategy" ["one"]', dict(instanced=True)),
              ('"whitted"', dict()),
              ('"whitted"', dict(instanced=True, maxdepth=3, sampler="sobol"))]


def check_tree(P, n=1500, res=24, spp=2, wave_paths=300, threads=2, cases=None):
    """per-sample radiance within a few ulp of the oracle's recursion (same rays, same samples: the ray and BVH visit
    counters are EQUAL; the nested f * li * w products are associated differently), image, pipelined waves invisible"""
    import oracle_api
    for integ, kw in (cases or TREE_CASES):
        kw = dict(dict(res=res, spp=spp), **kw)
        hs = P.HostScene.parse_string(tree_scene(P, integ, **kw))
        whitted = "whitted" in integ
        assert hs.desc.integrator.type == (P.capi.PT_INTEGRATOR_WHITTED if whitted else P.capi.PT_INTEGRATOR_DIRECT)
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        sb = hs.desc.film.sample_bounds
        rng = np.random.RandomState(8)
        pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
        sn = rng.randint(0, kw["spp"], n).astype(np.int32)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        bad = np.abs(la - lb) > 2e-5 * np.maximum(np.abs(lb), 1e-3)
        assert not bad.any(), (integ, kw, int(bad.any(1).sum()), la[bad.any(1)][:3], lb[bad.any(1)][:3])
        assert la.max() > 0
        rgb, st = sc.render(collect_counters=1, wave_paths=wave_paths)
        ref, so = orc.render(threads=threads)
        assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5, (integ, kw)
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert st[k] == so[k], (integ, kw, k, st[k], so[k])
        if not kw.get("instanced") and kw.get("maxdepth", 5) >= 5:
            assert st["path_vertices"] > 1.02 * st["camera_rays"]      # the tree has more than one vertex for some samples
        one, _ = sc.render(wave_paths=wave_paths, wave_segments=1)
        assert np.array_equal(one, rgb), (integ, kw)


def check_ao_random(P, res=33, spp=3, threads=2):
    """ambientocclusion under Sampler "random": the pixel's nsamples x spp array is drawn by start_pixel, before its first
    camera sample (image and counters equal to the oracle's)"""
    import oracle_api
    for kw in (dict(nsamples=4), dict(nsamples=3, cos=False, instanced=True)):
        hs = P.HostScene.parse_string(ao_scene(P, res=res, spp=spp, sampler="random", **kw))
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        rgb, st = sc.render(collect_counters=1)
        ref, so = orc.render(threads=threads)
        assert ref.mean() > 0 and np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-6, kw
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested"):
            assert st[k] == so[k], (kw, k, st[k], so[k])


def test_hostcheck_ao_random_sampler(HC):
    check_ao_random(HC)


def stratified_variant(txt, strata=2, dims=4, jitter=True):
    """the scene text with Sampler "stratified" (strata x strata samples per pixel) in place of its sampler"""
    lines = txt.split("\n")
    k = [i for i, l in enumerate(lines) if l.startswith("Sampler ")][0]
    lines[k] = 'Sampler "stratified" "integer xsamples" [%d] "integer ysamples" [%d] "integer dimensions" [%d] "bool jitter" ["%s"]' % (strata, strata, dims, "true" if jitter else "false")
    return "\n".join(lines)


def check_stratified(P, res=33, threads=2):
    """Sampler "stratified" (samplers/stratified.rs, core/sampler/pixel_sampler.rs): the pixel's shuffled 1D / 2D tables for the
    first `dimensions` requests of each kind, the tile's PCG32 stream afterwards, Latin-hypercube sample arrays for
    ambientocclusion and directlighting "all" - drawn on the device in the reference's order: image and counters equal to
    the oracle's (a 33-pixel image has partial tiles)"""
    import oracle_api
    power = 'Integrator "path" "string lightsamplestrategy" ["power"]'
    cases = [(P.scenes.cornell_box(res=res, spp=4, materials="mixed", extra_lights=True).replace('Integrator "path"', power), dict(strata=2)),
             (P.scenes.cornell_box(res=res, spp=4, materials="mixed", extra_lights=True).replace('Integrator "path"', power), dict(strata=3, dims=1, jitter=False)),
             (P.scenes.cornell_box(res=res, spp=4, materials="mixed").replace('Integrator "path"', power), dict(strata=1, dims=0)),
             (ao_scene(P, res=res, spp=4, nsamples=5), dict(strata=2)),
             (tree_scene(P, '"directlighting"', res=res, spp=4, nsamples=3), dict(strata=2, dims=2)),
             (tree_scene(P, '"whitted"', res=res, spp=4), dict(strata=2))]
    for txt, kw in cases:
        hs = P.HostScene.parse_string(stratified_variant(txt, **kw))
        S = hs.desc.sampler
        assert (S.type, S.samples_per_pixel, S.strata) == (P.capi.PT_SAMPLER_STRATIFIED, kw["strata"] ** 2, kw["strata"])
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        rgb, st = sc.render(collect_counters=1)
        ref, so = orc.render(threads=threads)
        assert ref.mean() > 0 and np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5, kw
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested"):
            assert st[k] == so[k], (kw, k, st[k], so[k])
        assert st["camera_rays"] >= res * res * kw["strata"] ** 2


def test_hostcheck_stratified_sampler(HC):
    check_stratified(HC)


def zerotwo_variant(txt, spp=4, dims=4, name="02sequence"):
    lines = txt.split("\n")
    k = [i for i, l in enumerate(lines) if l.startswith("Sampler ")][0]
    lines[k] = 'Sampler "%s" "integer pixelsamples" [%d] "integer dimensions" [%d]' % (name, spp, dims)
    return "\n".join(lines)


def check_zerotwo(P, res=33, threads=2):
    """Sampler "02sequence" / "lowdiscrepancy" (samplers/zerotwosequence.rs): scrambled van der Corput / Sobol' tables per
    pixel (Gray-code order, shuffled within and across the pixel samples), (0,2)-sequence sample arrays for
    ambientocclusion and directlighting "all", 3 requested samples rounded up to 4: image and counters equal to the oracle's"""
    import oracle_api
    power = 'Integrator "path" "string lightsamplestrategy" ["power"]'
    cases = [(P.scenes.cornell_box(res=res, spp=4, materials="mixed", extra_lights=True).replace('Integrator "path"', power), dict(spp=4)),
             (P.scenes.cornell_box(res=res, spp=4, materials="mixed").replace('Integrator "path"', power), dict(spp=3, dims=2, name="lowdiscrepancy")),
             (ao_scene(P, res=res, spp=4, nsamples=5), dict(spp=2)),
             (tree_scene(P, '"directlighting"', res=res, spp=4, nsamples=3), dict(spp=4, dims=1)),
             (tree_scene(P, '"whitted"', res=res, spp=4), dict(spp=1, dims=0))]
    for txt, kw in cases:
        hs = P.HostScene.parse_string(zerotwo_variant(txt, **kw))
        assert (hs.desc.sampler.type, hs.desc.sampler.samples_per_pixel) == (P.capi.PT_SAMPLER_ZEROTWO, kw["spp"])
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        rgb, st = sc.render(collect_counters=1)
        ref, so = orc.render(threads=threads)
        assert ref.mean() > 0 and np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5, kw
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested"):
            assert st[k] == so[k], (kw, k, st[k], so[k])
        assert st["camera_rays"] >= res * res * (1 << (kw["spp"] - 1).bit_length())


def test_hostcheck_zerotwo_sampler(HC):
    check_zerotwo(HC)


def test_zerotwo_tables_are_stratified(oracle):
    """the defining property of the (0,2)-sequence the oracle's tables are built from: every power-of-two prefix of the Gray-code
    walk puts exactly one point into every elementary interval of that size (checked through an image: a 1 spp and a 4 spp
    render of a furnace-like scene differ, while the generator matrices are pinned by their recurrence)"""
    c1 = [1 << 31]
    for i in range(31):
        c1.append(c1[-1] ^ (c1[-1] >> 1))
    assert [hex(v) for v in c1[:8]] == ["0x80000000", "0xc0000000", "0xa0000000", "0xf0000000", "0x88000000", "0xcc000000", "0xaa000000", "0xff000000"] and c1[31] == 0xffffffff
    c0 = [1 << (31 - i) for i in range(32)]
    pts, v = [], [0, 0]
    for i in range(16):
        pts.append((v[0] / 2 ** 32, v[1] / 2 ** 32))
        z = ((i + 1) & -(i + 1)).bit_length() - 1
        v = [v[0] ^ c0[z], v[1] ^ c1[z]]
    for a, b in ((16, 1), (8, 2), (4, 4), (2, 8), (1, 16)):       # all elementary intervals of area 1 / 16
        assert len({(int(x * a), int(y * b)) for x, y in pts}) == 16


def check_tree_random(P, res=33, spp=2, threads=2):
    """Sampler "random" under whitted / directlighting: the depth-first walk draws from the tile's PCG32 stream in the
    recursion's own order (same image, same counters as the oracle's recursion); strategy "all" reads the sample arrays
    RandomSampler::start_pixel draws ahead of every pixel (separate u_light / u_scattering arrays, random.rs:46-60), then
    get_2d pairs once they run out (maxdepth 2: the glass vertex and its reflection use them up)"""
    import oracle_api
    for integ, kw in (('"whitted"', {}), ('"directlighting" "string strategy" ["one"]', {}), ('"directlighting"', dict(nsamples=3)), ('"directlighting"', dict(nsamples=2, maxdepth=2))):
        hs = P.HostScene.parse_string(tree_scene(P, integ, res=res, spp=spp, sampler="random", **kw))
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        rgb, st = sc.render(collect_counters=1)
        ref, so = orc.render(threads=threads)
        assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5, integ
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert st[k] == so[k], (integ, k, st[k], so[k])


def test_hostcheck_tree_integrators_random_sampler(HC):
    check_tree_random(HC)


def test_hostcheck_tree_integrators(HC):
    """Integrators "directlighting" (strategies one / all, sample arrays) and "whitted": the specular reflection /
    transmission recursion as a depth-first tree walk per path slot."""
    check_tree(HC)


def write_pfm(path, img):
    """RGB float image, row 0 = top, as a little-endian PFM (rows bottom to top in the file)"""
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (w, h))
        f.write(np.ascontiguousarray(img[::-1], "<f4").tobytes())


def env_scene(P, tmpdir, mapped, integrator='"path"', res=24, spp=4, sampler="halton", strategy=None, map_shape=(8, 16), constant=None):
    """Cornell box with mixed materials (open towards the camera) inside an environment light: constant, or a 16 x 8
    lat-long map with one bright texel region, rotated; the box's own area light stays."""
    txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", sampler=sampler, maxdepth=5)
    light = 'AttributeBegin\nRotate -70 1 0 0\nRotate 25 0 0 1\nLightSource "infinite" "rgb L" [0.5 0.6 0.9] "rgb scale" [1.5 1 1]%s\nAttributeEnd\n'
    if mapped:
        rng = np.random.RandomState(3)
        img = rng.uniform(0.05, 0.6, tuple(map_shape) + (3,)).astype(np.float32)
        img[1:3, 4:7] = [9.0, 7.0, 4.0]       # a sun
        img[map_shape[0] - 2, 2] = -1.0       # negative texels are clamped (infinite.rs:54-55)
        if constant is not None:
            img[:] = constant
        name = "env_%dx%d%s.pfm" % (map_shape[1], map_shape[0], "" if constant is None else "_c")
        write_pfm(os.path.join(str(tmpdir), name), img)
        light = light % (' "string mapname" ["%s"]' % os.path.join(str(tmpdir), name))
    else:
        light = light % ""
    txt = txt.replace("WorldEnd", light + "WorldEnd")
    head, rest = txt.split('Integrator "path"', 1)
    extra = ' "string lightsamplestrategy" ["%s"]' % strategy if strategy else ""
    return head + "Integrator " + integrator + extra + rest


ENV_CASES = [dict(mapped=False), dict(mapped=True), dict(mapped=True, sampler="sobol", strategy="power"),
             dict(mapped=True, integrator='"directlighting" "string strategy" ["one"]'), dict(mapped=True, integrator='"whitted"'),
             dict(mapped=False, integrator='"directlighting"'),
             dict(mapped=True, map_shape=(7, 12)), dict(mapped=True, map_shape=(9, 16), sampler="sobol")]   # resampled to 16 x 8 / 16 x 16 (mipmap.rs:316-404)


def check_env(P, tmpdir, n=1500, res=24, spp=4, wave_paths=300, threads=2, rtol=1e-5, frac=1.0, cases=None):
    """LightSource "infinite" (lights/infinite.rs): Le of escaped rays, Distribution2D importance sampling + MIS against the
    oracle's own MIPMap / Distribution2D restatement; same rays (equal counters).  On the GPU acos / atan2 are CUDA's (a few
    ulp in the lookup coordinates and the MIS pdf): rtol / frac are the caller's."""
    import oracle_api
    for kw in (cases or ENV_CASES):
        kw = dict(dict(res=res, spp=spp), **kw)
        hs = P.HostScene.parse_string(env_scene(P, tmpdir, **kw))
        assert hs.desc.n_envmaps == 1 and hs.desc.lights[hs.desc.n_lights - 1].type == P.capi.PT_LIGHT_INFINITE
        assert (hs.desc.envmaps[0].width, hs.desc.envmaps[0].height) == (tuple(kw.get("map_shape", (8, 16)))[::-1] if kw["mapped"] else (1, 1))
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        sb = hs.desc.film.sample_bounds
        rng = np.random.RandomState(10)
        pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
        sn = rng.randint(0, kw["spp"], n).astype(np.int32)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        ok = (np.abs(la - lb) <= rtol * np.maximum(np.abs(lb), 1e-3)).all(1)
        assert ok.mean() >= frac, (kw, ok.mean(), la[~ok][:3], lb[~ok][:3])
        assert (lb > 0).all(1).mean() > 0.5          # the environment lights most samples
        rgb, st = sc.render(collect_counters=1, wave_paths=wave_paths)
        ref, so = orc.render(threads=threads)
        assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-4, kw
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert st[k] == so[k], (kw, k, st[k], so[k])
        one, _ = sc.render(wave_paths=wave_paths, wave_segments=1)
        assert np.array_equal(one, rgb), kw


def test_hostcheck_infinite_light(HC, tmp_path):
    check_env(HC, tmp_path, cases=[ENV_CASES[k] for k in (0, 2, 3, 4, 6, 7)])      # all eight on the GPU (test_gpu_parity.py)


def map_light_scene(P, tmpdir, kind, integrator='"path"', res=24, spp=4, sampler="halton", strategy=None, mapped=True):
    """Cornell box (mixed materials) lit by a LightSource "projection" (a slide projected from the camera side onto the
    blocks and the back wall) or "goniometric" (near the ceiling, translated: the reference transforms the direction as a
    point) in place of the box's area light's company: the area light stays, dimmed by nothing."""
    txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", sampler=sampler, maxdepth=5)
    mapname = ""
    if mapped:
        rng = np.random.RandomState(5)
        shape = (6, 10) if kind == "projection" else (8, 16)      # the slide is resampled to 16 x 8 (aspect 10 / 6 > 1)
        img = rng.uniform(0.0, 1.0, shape + (3,)).astype(np.float32)
        img[2:4, 3:6] = [6.0, 1.0, 0.5]
        img[0, 0] = -2.0                                          # clamped at zero by make_mipmap
        name = os.path.join(str(tmpdir), "%s.pfm" % kind)
        write_pfm(name, img)
        mapname = ' "string mapname" ["%s"]' % name
    if kind == "projection":
        light = ('AttributeBegin\nTranslate 278 273 -300\nRotate 8 0 1 0\nLightSource "projection" "rgb I" [90000 80000 70000] "rgb scale" [2 2 1] "float fov" [50]%s\nAttributeEnd\n' % mapname)
    else:
        light = ('AttributeBegin\nTranslate 278 400 280\nRotate 90 1 0 0\nScale 2 2 2\nLightSource "goniometric" "rgb L" [50000 50000 60000]%s\nAttributeEnd\n' % mapname)
    txt = txt.replace("WorldEnd", light + "WorldEnd")
    head, rest = txt.split('Integrator "path"', 1)
    extra = ' "string lightsamplestrategy" ["%s"]' % strategy if strategy else ""
    return head + "Integrator " + integrator + extra + rest


MAP_LIGHT_CASES = [dict(kind="projection"), dict(kind="goniometric"), dict(kind="projection", mapped=False, sampler="sobol", strategy="power"),
                   dict(kind="goniometric", strategy="power"), dict(kind="projection", integrator='"directlighting"'),
                   dict(kind="goniometric", integrator='"whitted"')]


def check_map_lights(P, tmpdir, n=1500, res=24, spp=4, wave_paths=300, threads=2, rtol=1e-5, frac=1.0):
    """LightSource "projection" / "goniometric" (lights/projection.rs, goniometric.rs) against the oracle: same rays (equal
    counters); the goniometric light's acos / atan2 are CUDA's on the GPU (rtol / frac are the caller's)"""
    import oracle_api
    for kw in MAP_LIGHT_CASES:
        kw = dict(dict(res=res, spp=spp), **kw)
        hs = P.HostScene.parse_string(map_light_scene(P, tmpdir, **kw))
        L = hs.desc.lights[hs.desc.n_lights - 1]
        assert L.type == (P.capi.PT_LIGHT_PROJECTION if kw["kind"] == "projection" else P.capi.PT_LIGHT_GONIOMETRIC) and hs.desc.n_envmaps == 1
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        sb = hs.desc.film.sample_bounds
        rng = np.random.RandomState(16)
        pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
        sn = rng.randint(0, kw["spp"], n).astype(np.int32)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        ok = (np.abs(la - lb) <= rtol * np.maximum(np.abs(lb), 1e-3)).all(1)
        assert ok.mean() >= frac, (kw, ok.mean(), la[~ok][:3], lb[~ok][:3])
        # the light matters: without it the same samples are darker
        base = oracle_api.OracleScene(P.HostScene.parse_string(map_light_scene(P, tmpdir, **kw).replace(' "rgb scale" [2 2 1]', "").replace('LightSource "%s"' % kw["kind"], 'LightSource "%s" "rgb scale" [0 0 0]' % kw["kind"]))).li(pix, sn)[0]
        assert lb.sum() > 1.05 * base.sum(), (kw, lb.sum(), base.sum())
        rgb, st = sc.render(collect_counters=1, wave_paths=wave_paths)
        ref, so = orc.render(threads=threads)
        assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-4, kw
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert st[k] == so[k], (kw, k, st[k], so[k])


def test_hostcheck_map_lights(HC, tmp_path):
    check_map_lights(HC, tmp_path)


def test_hostcheck_resampled_constant_map(HC, tmp_path):
    """A constant 5 x 3 map resampled to 8 x 4 stays constant up to rounding (the Lanczos weights are normalised): the render
    equals that of a constant 8 x 4 map."""
    a = HC.Scene(HC.HostScene.parse_string(env_scene(HC, tmp_path, True, map_shape=(3, 5), constant=1.0, res=16, spp=2, strategy="power"))).render(wave_paths=300)[0]
    b = HC.Scene(HC.HostScene.parse_string(env_scene(HC, tmp_path, True, map_shape=(4, 8), constant=1.0, res=16, spp=2, strategy="power"))).render(wave_paths=300)[0]
    assert b.mean() > 0 and np.sqrt(((a - b) ** 2).mean()) / b.mean() < 1e-3


def extra_materials_scene(P, integrator='"path"', res=24, spp=4, sampler="halton", maxdepth=6):
    """Cornell box with an uber short block (opacity < 1: the pass-through SpecularTransmission lobe, all five lobes), a
    substrate tall block (FresnelBlend, anisotropic), a translucent back wall, a second uber floor (Kr only) and a
    translucent ceiling with reflect = transmit = 0 (no BSDF at all: rays pass through)."""
    txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", extra_lights=True, sampler=sampler, maxdepth=maxdepth)
    rep = {'Material "glass" "float eta" [1.5]\n':
           'Material "uber" "rgb Kd" [0.3 0.2 0.1] "rgb Ks" [0.3 0.3 0.3] "rgb Kr" [0.2 0.2 0.3] "rgb Kt" [0.3 0.4 0.3] "rgb opacity" [0.8 0.7 0.9] "float uroughness" [0.05] "float vroughness" [0.2] "float index" [1.4]\n',
           'Material "metal" "rgb eta" [0.2 0.92 1.1] "rgb k" [3.9 2.45 2.14] "float roughness" [0.05]\n':
           'Material "substrate" "rgb Kd" [0.5 0.3 0.2] "rgb Ks" [0.1 0.1 0.15] "float uroughness" [0.02] "float vroughness" [0.3]\n',
           'Material "mirror" "rgb Kr" [0.8 0.8 0.8]\n':
           'Material "translucent" "rgb Kd" [0.5 0.5 0.4] "rgb Ks" [0.3 0.3 0.3] "rgb reflect" [0.6 0.5 0.5] "rgb transmit" [0.4 0.5 0.5] "float roughness" [0.2]\n',
           'Material "plastic" "rgb Kd" [0.4 0.4 0.4] "rgb Ks" [0.3 0.3 0.3] "float roughness" [0.1]\n':
           'Material "uber" "rgb Kd" [0.4 0.4 0.4] "rgb Ks" [0 0 0] "rgb Kr" [0.3 0.3 0.3]\n'}
    for a, b in rep.items():
        assert a in txt
        txt = txt.replace(a, b)
    shapes = [l for l in txt.split("\n") if l.startswith("Shape")]
    white = 'Material "matte" "rgb Kd" [0.725 0.71 0.68]\n'
    assert white + shapes[1] in txt
    txt = txt.replace(white + shapes[1], 'Material "translucent" "rgb reflect" [0 0 0] "rgb transmit" [0 0 0]\n' + shapes[1].replace("548.8", "500"))
    head, rest = txt.split('Integrator "path"', 1)
    return head + "Integrator " + integrator + rest


def check_extra_materials(P, n=1500, res=24, spp=4, wave_paths=300, threads=2):
    """uber / substrate / translucent (materials/{uber,substrate,translucent}.rs): per-sample radiance against the oracle
    (FresnelBlend's powf(x, 5) is formed in double on the device: a last-place difference in rare cases), equal counters"""
    import oracle_api
    for integ, kw in [('"path"', dict()), ('"path"', dict(sampler="sobol", maxdepth=9)), ('"directlighting" "string strategy" ["one"]', dict()), ('"whitted"', dict())]:
        hs = P.HostScene.parse_string(extra_materials_scene(P, integ, res=res, spp=spp, **kw))
        types = sorted(set(hs.desc.materials[i].type for i in range(hs.desc.n_materials)))
        assert types == [P.capi.PT_MAT_MATTE, P.capi.PT_MAT_UBER, P.capi.PT_MAT_SUBSTRATE, P.capi.PT_MAT_TRANSLUCENT], types
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        sb = hs.desc.film.sample_bounds
        rng = np.random.RandomState(12)
        pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
        sn = rng.randint(0, spp, n).astype(np.int32)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        ok = (np.abs(la - lb) <= 2e-5 * np.maximum(np.abs(lb), 1e-3)).all(1)
        assert ok.mean() >= 0.999, (integ, kw, ok.mean(), la[~ok][:3], lb[~ok][:3])
        assert la.max() > 0
        rgb, st = sc.render(collect_counters=1, wave_paths=wave_paths)
        ref, so = orc.render(threads=threads)
        assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-4, (integ, kw)
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert st[k] == so[k], (integ, kw, k, st[k], so[k])
        one, _ = sc.render(wave_paths=wave_paths, wave_segments=1)
        assert np.array_equal(one, rgb), (integ, kw)


def test_hostcheck_extra_materials(HC):
    check_extra_materials(HC)


def mix_scene(P, integrator='"path"', res=24, spp=4, sampler="halton", maxdepth=6):
    """Cornell box whose short block is a mix of plastic and smooth glass (amount per channel), the tall block a mix of
    substrate and rough metal, the floor a mix of OrenNayar matte and mirror with a negative / above-one amount (clamped),
    the back wall a mix of a known named material with an unknown one (matte from the mix's own parameters)."""
    txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", extra_lights=True, sampler=sampler, maxdepth=maxdepth)
    named = ('MakeNamedMaterial "pl" "string type" ["plastic"] "rgb Kd" [0.5 0.3 0.2] "rgb Ks" [0.4 0.4 0.4] "float roughness" [0.08]\n'
             'MakeNamedMaterial "gl" "string type" ["glass"] "float eta" [1.45]\n'
             'MakeNamedMaterial "sub" "string type" ["substrate"] "rgb Kd" [0.2 0.5 0.3] "rgb Ks" [0.3 0.3 0.3] "float uroughness" [0.05] "float vroughness" [0.2]\n'
             'MakeNamedMaterial "met" "string type" ["metal"] "float roughness" [0.1]\n'
             'MakeNamedMaterial "on" "string type" ["matte"] "rgb Kd" [0.6 0.6 0.5] "float sigma" [25]\n'
             'MakeNamedMaterial "mir" "string type" ["mirror"] "rgb Kr" [0.85 0.85 0.9]\n')
    rep = {'Material "glass" "float eta" [1.5]\n': 'Material "mix" "string namedmaterial1" ["pl"] "string namedmaterial2" ["gl"] "rgb amount" [0.7 0.4 0.55]\n',
           'Material "metal" "rgb eta" [0.2 0.92 1.1] "rgb k" [3.9 2.45 2.14] "float roughness" [0.05]\n': 'Material "mix" "string namedmaterial1" ["sub"] "string namedmaterial2" ["met"] "float amount" [0.35]\n',
           'Material "plastic" "rgb Kd" [0.4 0.4 0.4] "rgb Ks" [0.3 0.3 0.3] "float roughness" [0.1]\n': 'Material "mix" "string namedmaterial1" ["on"] "string namedmaterial2" ["mir"] "rgb amount" [-0.5 0.6 1.5]\n',
           'Material "mirror" "rgb Kr" [0.8 0.8 0.8]\n': 'Material "mix" "string namedmaterial1" ["pl"] "string namedmaterial2" ["nope"] "rgb Kd" [0.1 0.2 0.6]\n'}
    for a, b in rep.items():
        assert a in txt, a
        txt = txt.replace(a, b)
    txt = txt.replace("WorldBegin\n", "WorldBegin\n" + named, 1)
    head, rest = txt.split('Integrator "path"', 1)
    return head + 'Integrator %s "integer maxdepth" [%d]' % (integrator, maxdepth) + rest[rest.index("\n"):]


def check_mix(P, n=1500, res=24, spp=4, wave_paths=300, threads=2):
    """Material "mix" (materials/mix.rs, reflection/scaled.rs): per-sample radiance against the oracle's ScaledBxDF
    restatement under path (FresnelSpecular glass) and the tree integrators (split specular lobes), equal counters"""
    import oracle_api
    power = '"path" "string lightsamplestrategy" ["power"]'      # no voxel table: 10 s per scene in the CPU build
    for integ, kw in [(power, dict()), (power, dict(sampler="sobol", maxdepth=9)), ('"directlighting" "string strategy" ["one"]', dict()), ('"whitted"', dict())]:
        hs = P.HostScene.parse_string(mix_scene(P, integ, res=res, spp=spp, **kw))
        mats = [hs.desc.materials[i] for i in range(hs.desc.n_materials)]
        mixes = [m for m in mats if m.type == P.capi.PT_MAT_MIX]
        assert len(mixes) == 4 and all(mats[m.kd].type != P.capi.PT_MAT_MIX and mats[m.ks].type != P.capi.PT_MAT_MIX for m in mixes)
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        sb = hs.desc.film.sample_bounds
        rng = np.random.RandomState(18)
        pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
        sn = rng.randint(0, spp, n).astype(np.int32)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        ok = (np.abs(la - lb) <= 2e-5 * np.maximum(np.abs(lb), 1e-3)).all(1)
        assert ok.mean() >= 0.999, (integ, kw, ok.mean(), la[~ok][:3], lb[~ok][:3])
        assert la.max() > 0
        rgb, st = sc.render(collect_counters=1, wave_paths=wave_paths)
        ref, so = orc.render(threads=threads)
        assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-4, (integ, kw)
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert st[k] == so[k], (integ, kw, k, st[k], so[k])


def test_hostcheck_mix_material(HC):
    check_mix(HC)


def test_hostcheck_mix_limits(HC):
    """amount 1 = the first material alone (s2 = 0 scales the second one's lobes to black): same image as that material;
    more than five lobes in total, and a mix of mixes, are refused"""
    base = HC.scenes.cornell_box(res=16, spp=2, materials="mixed")
    named = ('MakeNamedMaterial "a" "string type" ["plastic"] "rgb Kd" [0.5 0.3 0.2]\nMakeNamedMaterial "b" "string type" ["matte"]\n'
             'MakeNamedMaterial "u" "string type" ["uber"] "rgb Kr" [0.2 0.2 0.2] "rgb Kt" [0.2 0.2 0.2] "rgb opacity" [0.5 0.5 0.5]\n'
             'MakeNamedMaterial "mm" "string type" ["mix"] "string namedmaterial1" ["a"] "string namedmaterial2" ["b"]\n')
    gl = 'Material "glass" "float eta" [1.5]\n'
    one = base.replace("WorldBegin\n", "WorldBegin\n" + named, 1).replace(gl, 'Material "mix" "string namedmaterial1" ["a"] "string namedmaterial2" ["b"] "float amount" [1]\n')
    alone = base.replace(gl, 'Material "plastic" "rgb Kd" [0.5 0.3 0.2]\n')
    a = HC.Scene(HC.HostScene.parse_string(one)).render()[0]
    b = HC.Scene(HC.HostScene.parse_string(alone)).render()[0]
    assert b.mean() > 0 and np.sqrt(((a - b) ** 2).mean()) / b.mean() < 0.05      # same estimator up to the lobe choice among 3 instead of 2
    with pytest.raises(HC.PbrtError):
        HC.Scene(HC.HostScene.parse_string(base.replace("WorldBegin\n", "WorldBegin\n" + named, 1).replace(gl, 'Material "mix" "string namedmaterial1" ["u"] "string namedmaterial2" ["a"]\n')))
    with pytest.raises(HC.PbrtError):
        HC.HostScene.parse_string(base.replace("WorldBegin\n", "WorldBegin\n" + named, 1).replace(gl, 'Material "mix" "string namedmaterial1" ["mm"] "string namedmaterial2" ["a"]\n'))


CAMERA_CASES = [('Camera "orthographic" "float screenwindow" [-285 285 -280 280]', None),
                ('Camera "orthographic" "float screenwindow" [-285 285 -280 280] "float lensradius" [6] "float focaldistance" [1000] "float shutteropen" [0.2]', None),
                ('Camera "environment"', "LookAt 278 273 250  278 273 500  0 1 0")]


def camera_scene(P, cam, lookat=None, res=24, spp=4, sampler="halton"):
    txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", extra_lights=True, sampler=sampler)
    assert 'Camera "perspective" "float fov" [39.3077]' in txt
    txt = txt.replace('Camera "perspective" "float fov" [39.3077]', cam)
    txt = txt.replace('Integrator "path"', 'Integrator "path" "string lightsamplestrategy" ["power"]')     # no voxel table: 10 s per scene in the CPU build
    if lookat:
        txt = txt.replace("LookAt 278 273 -800  278 273 0  0 1 0", lookat)
        txt = txt.replace('"integer yresolution" [%d]' % res, '"integer yresolution" [%d]' % (res // 2))   # lat-long: 2 x 1
    return txt


def check_cameras(P, n=1500, res=24, spp=4, threads=2):
    """Camera "orthographic" (cameras/orthographic.rs:60-90, with and without a lens) and "environment" (environment.rs:36-58,
    from inside the box): camera rays bit for bit (sin / cos are glibc-exact on the device), per-sample radiance, image, counters"""
    import oracle_api
    for (cam, lookat), sampler in zip(CAMERA_CASES, ("halton", "sobol", "random")):
        hs = P.HostScene.parse_string(camera_scene(P, cam, lookat, res=res, spp=spp, sampler=sampler))
        want = P.capi.PT_CAMERA_ENVIRONMENT if "environment" in cam else P.capi.PT_CAMERA_ORTHOGRAPHIC
        assert hs.desc.camera.type == want
        sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
        if sampler != "random":
            sb = hs.desc.film.sample_bounds
            rng = np.random.RandomState(14)
            pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
            sn = rng.randint(0, spp, n).astype(np.int32)
            (ra, pa), (rb, pb) = sc.camera_rays(pix, sn), orc.camera_rays(pix, sn)
            assert np.array_equal(ra.view(np.uint32), rb.view(np.uint32)) and np.array_equal(pa, pb), cam
            if want == P.capi.PT_CAMERA_ORTHOGRAPHIC and "lensradius" not in cam:
                assert np.ptp(ra[:, 4:7], axis=0).max() == 0 and np.ptp(ra[:, 0]) > 400      # rays = (o, tmax, d, time): parallel, from a plane
            la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
            ok = (np.abs(la - lb) <= 1e-5 * np.maximum(np.abs(lb), 1e-3)).all(1)
            assert ok.mean() >= 0.999 and la.max() > 0, (cam, ok.mean())
        rgb, st = sc.render(collect_counters=1)
        ref, so = orc.render(threads=threads)
        assert ref.mean() > 0 and np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-4, cam
        for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
            assert st[k] == so[k], (cam, k, st[k], so[k])
        if want == P.capi.PT_CAMERA_ENVIRONMENT:
            assert (rgb.reshape(-1, 3).sum(1) > 0).mean() > 0.7       # from inside the box (open towards -z) most directions see a wall


def test_hostcheck_cameras(HC):
    check_cameras(HC)


def test_hostcheck_hlbvh_split(HC):
    """splitmethod "hlbvh": the same stage logic on the Morton-treelet tree (visit counters equal the oracle's)."""
    hl = lambda t: t.replace('Accelerator "bvh"', 'Accelerator "bvh" "string splitmethod" ["hlbvh"]')
    hs, sc, orc = _cmp_li(HC, hl(HC.scenes.instanced_blobs(n_instances=30, n_sub=1, res=(48, 32), spp=4, n_lights=4)), 2000, 4, 0.999)
    rgb, st = sc.render(collect_counters=1)
    ref, so = orc.render(threads=2)
    assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5
    for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
        assert st[k] == so[k], k


def test_hostcheck_lbvh_layout(HC):
    """Accelerator layout "lbvh" (accel/lbvh.rs): same stage logic on the binary node layout, instances included."""
    lb = lambda t: t.replace('Accelerator "bvh"', 'Accelerator "bvh" "string layout" "lbvh"')
    hs, sc, orc = _cmp_li(HC, lb(HC.scenes.instanced_blobs(n_instances=30, n_sub=1, res=(48, 32), spp=4, n_lights=4)), 2000, 4, 0.999)
    assert hs.desc.accel.type == HC.capi.PT_ACCEL_LBVH
    rgb, st = sc.render(collect_counters=1)
    ref, so = orc.render(threads=2)
    assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5
    for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
        assert st[k] == so[k], k


def test_hostcheck_gaussian_filter_film(HC):
    P = HC
    import oracle_api
    txt = P.scenes.cornell_box(res=20, spp=2).replace('Sampler', 'PixelFilter "gaussian" "float xwidth" [1.5] "float ywidth" [1.5]\nSampler')
    hs = P.HostScene.parse_string(txt)
    rgb, _ = P.Scene(hs).render(wave_paths=500)
    ref, _ = oracle_api.OracleScene(hs).render(threads=2)
    assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5


def test_hostcheck_random_sampler_tile_streams(HC):
    """Sampler "random": one PCG32 stream per 16x16 render tile, drawn from in pixel / sample / path order
    (samplers/random.rs, sampler.rs:285).  Same image and the same ray / node / triangle counts as the oracle's
    tile-sequential loop; tile groups (small waves) and tile shards do not change the streams."""
    P = HC
    import oracle_api
    import ctypes as C
    from pbrt_r3_b200 import dist as D
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=33, spp=2, sampler="random"))
    sc = P.Scene(hs)
    rgb, st = sc.render(collect_counters=1)
    ref, so = oracle_api.OracleScene(hs).render(threads=2)
    assert np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean() < 1e-5
    for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"):
        assert st[k] == so[k], k
    assert np.array_equal(rgb, sc.render(wave_paths=600)[0])
    acc = np.zeros(33 * 33 * 4, np.float32)
    for r in range(3):
        ptr, n, _ = sc.render_partial(shard_index=r, shard_count=3)
        acc += np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), (n * 4,))
    assert np.sqrt(((D.resolve_host(acc).reshape(33, 33, 3) - rgb) ** 2).mean()) / rgb.mean() < 1e-6
    with pytest.raises(P.PbrtError):
        sc.render_partial(spp_begin=0, spp_end=1)      # a sample range alone cannot reproduce the sequential streams
    with pytest.raises(P.PbrtError):
        sc.li(np.zeros((1, 2), np.int32), np.zeros(1, np.int32))
