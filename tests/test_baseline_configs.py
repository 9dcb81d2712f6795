"""BASELINE.json configs 2-5 at their stated resolution and geometry (VERDICT round 1, row g): image relative RMSE <= 1e-3
vs the oracle at equal spp AND equal camera / regular / shadow ray counts.  The spp is reduced (the oracle has to finish
in seconds); resolution, triangle counts, instance counts, light counts, maxdepth and spectrum width are the configs'.
Halton samples are pure functions of (pixel, sample number), so sample numbers 0..k-1 of the full-size render are exactly
what the k-spp render takes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RMSE_TOL = 1e-3


def rel_rmse(a, b):
    return float(np.sqrt(((a.astype(np.float64) - b) ** 2).mean()) / max(float(b.mean()), 1e-20))


def _compare(P, text, spectrum="rgb", expect=None, **opts):
    import oracle_api
    hs = P.HostScene.parse_string(text, spectrum=spectrum)
    if expect:
        expect(hs.desc)
    sc, orc = P.Scene(hs, 0), oracle_api.OracleScene(hs)
    try:
        rgb, st = sc.render(collect_counters=1, **opts)
        ref, so = orc.render()
        assert np.isfinite(rgb).all()
        e = rel_rmse(rgb, ref)
        assert e <= RMSE_TOL, e
        for k in ("camera_rays", "regular_rays", "shadow_rays"):
            assert st[k] == so[k], (k, st[k], so[k])
        # the device's traversal counters are the reference traversal's (the roofline's byte count rests on them)
        for k in ("nodes_visited", "tris_tested", "path_vertices"):
            assert st[k] == so[k], (k, st[k], so[k])
        return rgb, ref, st
    finally:
        sc.close()
        orc.close()


def test_config2_full_resolution(P):
    """configs[1]: 1M-triangle displaced mesh, 1024x1024, plastic / metal / glass, maxdepth 8 - 4 of the 64 spp."""
    def expect(d):
        assert d.n_tris > 990000 and tuple(d.film.full_resolution) == (1024, 1024) and d.integrator.max_depth == 8
    _compare(P, P.scenes.displaced_grid(n=708, res=1024, spp=4, maxdepth=8), expect=expect)


def test_config3_full_resolution(P):
    """configs[2]: 1000 instances of a 20 480-triangle blob (20.5 M instanced triangles), 256 emissive quads, 1920x1080,
    two-level 4-wide BVH, spatial light sampling, maxdepth 8 - 1 of the 256 spp."""
    def expect(d):
        assert d.n_instances == 1000 and d.n_lights == 512 and tuple(d.film.full_resolution) == (1920, 1080)
    _compare(P, P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(1920, 1080), spp=1, maxdepth=8, n_lights=256), expect=expect)


def test_config4_full_resolution(P):
    """configs[3]: 60-bin SampledSpectrum Cornell box, 512x512 - 8 of the 128 spp."""
    def expect(d):
        assert d.n_channels == 60 and tuple(d.film.full_resolution) == (512, 512)
    _compare(P, P.scenes.cornell_box(res=512, spp=8), spectrum="sampled", expect=expect)


def test_config5_crop_of_tiles(P):
    """configs[4]: the instanced scene at 3840x2160 - a 768x432 crop window of the 4K film (the sampler's Halton scales and
    every pixel's sample indices are those of the full film), 1 of the 1024 spp; then the same crop rendered as 4
    interleaved-tile shards whose film partials must add up to it (the multi-GPU split of that config)."""
    txt = P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(3840, 2160), spp=1, maxdepth=8, n_lights=256)
    key = '"string filename" ["blobs.exr"]'
    assert key in txt
    txt = txt.replace(key, key + ' "float cropwindow" [0.4 0.6 0.4 0.6]')

    def expect(d):
        b = d.film.cropped_pixel_bounds
        assert tuple(d.film.full_resolution) == (3840, 2160) and (b[2] - b[0], b[3] - b[1]) == (768, 432)
    rgb, ref, st = _compare(P, txt, expect=expect)
    assert rgb.shape == (432, 768, 3)
    from pbrt_r3_b200 import dist as D
    hs = P.HostScene.parse_string(txt)
    sc = P.Scene(hs, 0)
    try:
        parts, rays = [], 0
        for r in range(4):
            ptr, n, s = sc.render_partial(shard_index=r, shard_count=4)
            parts.append(D.film_tensor(ptr, n, "cuda:0").clone())
            rays += s["regular_rays"] + s["shadow_rays"]
        assert rays == st["regular_rays"] + st["shadow_rays"]
        both = sc.film_resolve(sum(parts).data_ptr())
        assert rel_rmse(both, rgb) < 1e-6
    finally:
        sc.close()
