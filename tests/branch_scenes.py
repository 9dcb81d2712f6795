"""Scenes that reach the branches of the hot path the BASELINE configs do not (TEST INFRASTRUCTURE): OrenNayar
(oren_nayar.rs:26-69), rough glass = MicrofacetReflection + MicrofacetTransmission (microfacet.rs:142-263,
glass.rs:47-110), glass with Kr = 0 / Kt = 0, the thin-lens camera (perspective.rs:141-166), Film
maxsampleluminance (film_tile.rs:86-89), the "power" / "uniform" light strategies (lightdistrib/{power,uniform}.rs) and
validate_radiance_result (sampler.rs:151-176).  Shared by the hostcheck (CPU) and the -m gpu parity tests."""

WHITE = 'Material "matte" "rgb Kd" [0.725 0.71 0.68]'
RED = 'Material "matte" "rgb Kd" [0.63 0.065 0.05]'
GREEN = 'Material "matte" "rgb Kd" [0.14 0.45 0.091]'
GLASS = 'Material "glass" "float eta" [1.5]'


def _base(P, res, spp, maxdepth=8):
    txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", extra_lights=True, maxdepth=maxdepth)
    for key in (WHITE, RED, GREEN, GLASS):
        assert key in txt
    return txt


def oren_nayar(P, res=32, spp=8):
    """matte sigma 20 / 60 / 95 (clamped to 90): the OrenNayar lobe instead of Lambert on the three matte walls."""
    txt = _base(P, res, spp)
    txt = txt.replace(WHITE, WHITE + ' "float sigma" [20]').replace(RED, RED + ' "float sigma" [60]').replace(GREEN, GREEN + ' "float sigma" [95]')
    return txt


def rough_glass(P, res=32, spp=8, kr=None, kt=None, ur=0.1, vr=0.3, remap=True):
    """the short block as rough (anisotropic) glass: MicrofacetReflection + MicrofacetTransmission, optionally with
    Kr = 0 (transmission lobe only) or Kt = 0 (reflection lobe only)."""
    m = GLASS + ' "float uroughness" [%g] "float vroughness" [%g]' % (ur, vr)
    if kr is not None:
        m += ' "rgb Kr" [%g %g %g]' % tuple(kr)
    if kt is not None:
        m += ' "rgb Kt" [%g %g %g]' % tuple(kt)
    if not remap:
        m += ' "bool remaproughness" ["false"]'
    return _base(P, res, spp).replace(GLASS, m)


def smooth_glass_partial(P, res=32, spp=8, kr=(0, 0, 0), kt=(0.9, 0.8, 0.7)):
    """smooth glass (FresnelSpecular, fresnel.rs:150-216) with a black or tinted Kr / Kt."""
    m = GLASS + ' "rgb Kr" [%g %g %g] "rgb Kt" [%g %g %g]' % (tuple(kr) + tuple(kt))
    return _base(P, res, spp).replace(GLASS, m)


def thin_lens(P, res=32, spp=8):
    txt = _base(P, res, spp)
    cam = 'Camera "perspective" "float fov" [39.3077]'
    assert cam in txt
    return txt.replace(cam, cam + ' "float lensradius" [12] "float focaldistance" [950]')


def max_sample_luminance(P, res=32, spp=8, value=1.5):
    txt = _base(P, res, spp)
    key = '"string filename" ["cornell.exr"]'
    assert key in txt
    return txt.replace(key, key + ' "float maxsampleluminance" [%g]' % value)


def light_strategy(P, name, res=32, spp=8):
    """"power": one distribution over the area-light triangles, the point and the distant light.  "uniform" is only used
    by the reference when the scene has exactly ONE light (create_light_sample_distribution.rs:23-29), so that variant
    keeps the point light only."""
    txt = _base(P, res, spp)
    if name == "uniform":
        txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", extra_lights=False, maxdepth=8)
        txt = txt.replace('AreaLightSource "diffuse" "rgb L" [17 12 4]\n', "")
        txt = txt.replace("WorldBegin\n", 'WorldBegin\nAttributeBegin\nLightSource "point" "rgb I" [400000 400000 600000] "point from" [278 400 200]\nAttributeEnd\n', 1)
    assert 'Integrator "path"' in txt
    return txt.replace('Integrator "path"', 'Integrator "path" "string lightsamplestrategy" ["%s"]' % name)


def invalid_radiance(P, res=32, spp=4):
    """validate_radiance_result (sampler.rs:151-176): the green wall's reflectance overflows to +inf in one channel (paths
    through it carry inf, and inf * 0 = NaN against the light's zero blue channel), and a second area light emits negative
    radiance (samples that see it have luminance < -1e-5): all of these samples must turn black, the others are untouched.
    (The light distribution only sees finite, clamped contributions, so the reference's asserts stay silent.)"""
    txt = _base(P, res, spp)
    assert '"rgb L" [17 12 4]' in txt
    txt = txt.replace('"rgb L" [17 12 4]', '"rgb L" [17 12 0]').replace(GREEN, 'Material "matte" "rgb Kd" [0.14 1e39 0.091]')
    bad = ('AttributeBegin\nAreaLightSource "diffuse" "rgb L" [0.5 -50 4]\nMaterial "matte" "rgb Kd" [0 0 0]\n'
           'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [100 548.0 100 160 548.0 100 160 548.0 160 100 548.0 160]\nAttributeEnd\n')
    return txt.replace("WorldEnd\n", bad + "WorldEnd\n")


ALL = {
    "oren_nayar": oren_nayar,
    "rough_glass": rough_glass,
    "rough_glass_noremap": lambda P, res=32, spp=8: rough_glass(P, res, spp, ur=0.05, vr=0.05, remap=False),
    "rough_glass_kr0": lambda P, res=32, spp=8: rough_glass(P, res, spp, kr=(0, 0, 0)),
    "rough_glass_kt0": lambda P, res=32, spp=8: rough_glass(P, res, spp, kt=(0, 0, 0)),
    "smooth_glass_kr0": smooth_glass_partial,
    "smooth_glass_kt0": lambda P, res=32, spp=8: smooth_glass_partial(P, res, spp, kr=(0.9, 0.9, 0.9), kt=(0, 0, 0)),
    "thin_lens": thin_lens,
    "max_sample_luminance": max_sample_luminance,
    "power": lambda P, res=32, spp=8: light_strategy(P, "power", res, spp),
    "uniform": lambda P, res=32, spp=8: light_strategy(P, "uniform", res, spp),
    "invalid_radiance": invalid_radiance,
}
