"""Branches of the hot path that BASELINE's five configs never reach: OrenNayar, rough glass
(MicrofacetReflection + MicrofacetTransmission), glass with black Kr / Kt, thin lens, maxsampleluminance, the "power" /
"uniform" light strategies on the device and validate_radiance_result.  Scenes: tests/branch_scenes.py.

Two legs per branch:
  * CPU (`-m "not gpu"`): the product's stage code compiled for the host (tests/hostcheck) vs the oracle, per-sample
    radiance at bit level (same libm on both sides) + image + ray counters;
  * GPU (`-m gpu`): the CUDA library through the C ABI vs the oracle, per-sample radiance and image RMSE <= 1e-3,
    equal camera / regular / shadow ray counts."""
import numpy as np
import pytest

import branch_scenes as BS

RMSE_TOL = 1e-3
COUNTERS = ("camera_rays", "regular_rays", "shadow_rays")


def rel_rmse(a, b):
    return float(np.sqrt(((a.astype(np.float64) - b) ** 2).mean()) / max(float(np.abs(b).mean()), 1e-20))


def _samples(hs, n, spp, seed):
    sb = hs.desc.film.sample_bounds
    rng = np.random.RandomState(seed)
    pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
    return pix, rng.randint(0, spp, n).astype(np.int32)


def _lobe_census(P, hs):
    """which materials the parsed scene really holds (guards against a variant silently falling back to the base scene)"""
    d = hs.desc
    return [(d.materials[i].type, d.materials[i].sigma, d.materials[i].uroughness, d.materials[i].vroughness) for i in range(d.n_materials)]


def _expect(P, name, hs):
    d = hs.desc
    mats = _lobe_census(P, hs)
    if name == "oren_nayar":
        assert sum(1 for m in mats if m[0] == P.capi.PT_MAT_MATTE and m[1] > 0) == 3
    if name.startswith("rough_glass"):
        assert any(m[0] == P.capi.PT_MAT_GLASS and m[2] > 0 and m[3] > 0 for m in mats)
    if name == "thin_lens":
        assert d.camera.lens_radius == 12.0 and d.camera.focal_distance == 950.0
    if name == "max_sample_luminance":
        assert d.film.max_sample_luminance == 1.5
    if name == "power":
        assert d.integrator.light_strategy == 2 and d.n_lights == 4
        return
    if name == "uniform":
        assert d.integrator.light_strategy == 1 and d.n_lights == 1


# ---- CPU leg: hostcheck build vs oracle ----------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(BS.ALL))
def test_hostcheck_branch(HC, name):
    import oracle_api
    P = HC
    spp = 4
    txt = BS.ALL[name](P, res=24, spp=spp)
    if "lightsamplestrategy" not in txt and name != "oren_nayar":
        # the serial host build of the 64^3-voxel light table costs 20 s per scene and does not matter to these branches
        # (the spatial strategy is covered by test_hostcheck.py and by oren_nayar here)
        txt = txt.replace('Integrator "path"', 'Integrator "path" "string lightsamplestrategy" ["power"]')
    hs = P.HostScene.parse_string(txt)
    _expect(P, name, hs)
    sc, orc = P.Scene(hs), oracle_api.OracleScene(hs)
    pix, sn = _samples(hs, 2500, spp, seed=17)
    la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
    if name == "thin_lens":      # the lens sample really moves the ray origin off the pinhole
        ra, _ = sc.camera_rays(pix, sn)
        rb, _ = orc.camera_rays(pix, sn)
        assert np.array_equal(ra.view(np.uint32), rb.view(np.uint32))
        assert len(np.unique(ra[:, :3], axis=0)) > 1000
    # same libm on both sides: bit-identical on almost every sample, the rest one rounding apart (beta*(a+b)/p vs
    # beta*a/p + beta*b/p when both MIS halves are non-zero)
    if name == "invalid_radiance":
        # orc_li is PathIntegrator::li itself; pt_li applies validate_radiance_result (k_finish) like render_tile does
        y = lb @ np.array([0.212671, 0.715160, 0.072169], np.float32)
        invalid = ~np.isfinite(lb).all(1) | (y < -1e-5) | np.isinf(y)
        assert invalid.sum() > 100 and np.isnan(lb).any() and np.isinf(lb).any() and (y[np.isfinite(y)] < -1e-5).any()
        assert not la[invalid].any() and np.isfinite(la).all()
        la, lb = la[~invalid], lb[~invalid]
    same = (la.view(np.uint32) == lb.view(np.uint32)).all(1)
    assert same.mean() >= 0.97, (name, same.mean())
    assert (np.abs(la - lb) <= 1e-5 * np.maximum(np.abs(lb), 1e-5)).all(), name
    rgb, st = sc.render(collect_counters=1, wave_paths=700)
    ref, so = orc.render(threads=2)
    assert np.isfinite(rgb).all()
    assert rel_rmse(rgb, ref) < 1e-5, (name, rel_rmse(rgb, ref))
    for k in COUNTERS + ("nodes_visited", "tris_tested", "path_vertices"):
        assert st[k] == so[k], (name, k, st[k], so[k])
    if name == "max_sample_luminance":
        base = P.HostScene.parse_string(txt.replace(' "float maxsampleluminance" [1.5]', ""))
        r0 = oracle_api.OracleScene(base).render(threads=2)[0]
        assert rel_rmse(r0, ref) > 1e-2            # the clamp changes the image


# ---- GPU leg ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(BS.ALL))
def test_gpu_branch(P, name):
    import oracle_api
    spp = 8
    hs = P.HostScene.parse_string(BS.ALL[name](P, res=96, spp=spp))
    _expect(P, name, hs)
    sc, orc = P.Scene(hs, 0), oracle_api.OracleScene(hs)
    try:
        pix, sn = _samples(hs, 20000, spp, seed=23)
        la, lb = sc.li(pix, sn), orc.li(pix, sn)[0]
        if name == "invalid_radiance":     # pt_li validates (k_finish), orc_li is the raw li()
            y = lb @ np.array([0.212671, 0.715160, 0.072169], np.float32)
            invalid = ~np.isfinite(lb).all(1) | (y < -1e-5) | np.isinf(y)
            assert invalid.sum() > 500 and not la[invalid].any() and np.isfinite(la).all()
            la, lb = la[~invalid], lb[~invalid]
        close = (np.abs(la - lb) <= 1e-3 * np.maximum(np.abs(lb), 1e-2)).all(1)
        # rounding of beta * (a + b) / p vs beta * a / p + beta * b / p may flip a discrete decision on a rare path
        assert close.mean() > 0.999, (name, close.mean())
        if name == "thin_lens":
            ra, _ = sc.camera_rays(pix, sn)
            rb, _ = orc.camera_rays(pix, sn)
            assert np.array_equal(ra.view(np.uint32), rb.view(np.uint32))
        rgb, st = sc.render()
        ref, so = orc.render()
        assert np.isfinite(rgb).all()
        assert rel_rmse(rgb, ref) <= RMSE_TOL, (name, rel_rmse(rgb, ref))
        for k in COUNTERS:
            assert st[k] == so[k], (name, k, st[k], so[k])
    finally:
        sc.close()
        orc.close()
