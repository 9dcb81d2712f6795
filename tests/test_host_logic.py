"""CPU-side checks (no GPU): the C-ABI library loads and exports every declared symbol, the ctypes mirror matches the
header, the .pbrt front end reproduces the reference's scene-description rules, and the product's host BVH build
equals the oracle's restatement of the reference build (node by node, primitive order included)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(P):
    lib = P.capi.load()
    hdr = open(os.path.join(ROOT, "include", "pbrt_r3_b200.h")).read()
    declared = set(re.findall(r"\b(pt_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert set(P.capi.EXPORTS) == declared
    assert lib.pt_abi_version() == P.capi.PT_ABI_VERSION


def test_ctypes_mirror_matches_header(P):
    names = ["pt_material", "pt_light", "pt_spot", "pt_options", "pt_instance", "pt_camera", "pt_film", "pt_sampler", "pt_integrator", "pt_accel", "pt_scene_desc",
             "pt_render_opts", "pt_stats"]
    src = '#include "pbrt_r3_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){\n' + "".join(
        'printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in names) + 'printf("off_lights %zu\\n", offsetof(pt_scene_desc, lights));\n' \
        'printf("off_accel %zu\\n", offsetof(pt_scene_desc, accel));\nreturn 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = dict(l.split() for l in subprocess.check_output([exe]).decode().splitlines())
    for n in names:
        assert C.sizeof(getattr(P.capi, n)) == int(out[n]), n
    assert P.capi.pt_scene_desc.lights.offset == int(out["off_lights"])
    assert P.capi.pt_scene_desc.accel.offset == int(out["off_accel"])


def test_front_end_defaults_and_film_rules(P):
    hs = P.HostScene.parse_string('Camera "perspective"\nWorldBegin\nWorldEnd\n')
    d = hs.desc
    # render_options.rs:61-88 / film.rs:538-571 / filters/box_filter.rs:26-30 / halton.rs:279 / path.rs:257-262
    assert list(d.film.full_resolution) == [1280, 720]
    assert list(d.film.cropped_pixel_bounds) == [0, 0, 1280, 720]
    assert list(d.film.sample_bounds) == [-1, -1, 1281, 721]          # floor(0 - 0.5) .. ceil(res + 0.5) (film.rs:166-179)
    assert list(d.film.filter_radius) == [0.5, 0.5] and all(v == 1.0 for v in d.film.filter_table)
    assert (d.sampler.type, d.sampler.samples_per_pixel) == (P.capi.PT_SAMPLER_HALTON, 16)
    assert (d.integrator.max_depth, d.integrator.rr_threshold, d.integrator.light_strategy) == (5, 1.0, 0)
    assert (d.accel.type, d.accel.split_method, d.accel.max_prims_in_node) == (0, 0, 4)
    assert hs.filename == "pbrt.exr" and d.film.scale == 1.0 and np.isinf(d.film.max_sample_luminance)
    assert d.camera.lens_radius == 0.0 and d.camera.focal_distance == np.float32(1e6)


def test_front_end_cornell(P):
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=256, spp=16))
    d = hs.desc
    assert d.n_tris == 32 and d.n_lights == 2 and d.n_objects == 1 and d.n_instances == 0
    flags = np.ctypeslib.as_array(d.tri_flags, (32,))
    assert (flags & P.capi.PT_TRI_TWO_SIDED).all()              # twosided defaults to true (triangle.rs:707)
    lights = [d.lights[i] for i in range(2)]
    assert all(l.type == P.capi.PT_LIGHT_AREA and l.two_sided == 0 for l in lights)   # diffuse.rs:178
    tl = np.ctypeslib.as_array(d.tri_light, (32,))
    assert sorted(tl[tl >= 0]) == [0, 1] and [lights[0].tri, lights[1].tri] == [30, 31]
    # LookAt: right = up x dir (matrix4x4.rs:130-197) -> camera_to_world column 3 is the eye
    c2w = np.array(list(d.camera.camera_to_world)).reshape(4, 4)
    assert np.allclose(c2w[:3, 3], [278, 273, -800]) and np.allclose(c2w[:3, 2], [0, 0, 1])


def test_front_end_quirks(P):
    # degenerate triangles dropped (triangle.rs:726); material params looked up before shape params for spectra
    # (texture_params.rs:75-83); shape float params override material floats; point light translate order (point.rs:121)
    txt = '''Camera "perspective"
WorldBegin
Material "plastic" "rgb Kd" [0.1 0.2 0.3] "float roughness" [0.2]
Shape "trianglemesh" "integer indices" [0 1 2 0 0 1] "point P" [0 0 0 1 0 0 0 1 0] "rgb Kd" [0.9 0.9 0.9] "float roughness" [0.05]
Translate 1 2 3
Scale 2 2 2
LightSource "point" "point from" [1 1 1]
WorldEnd
'''
    hs = P.HostScene.parse_string(txt)
    d = hs.desc
    assert d.n_tris == 1
    m = d.materials[d.tri_material[0]]
    assert m.type == P.capi.PT_MAT_PLASTIC and abs(m.uroughness - 0.05) < 1e-9
    kd = [d.spectra[m.kd * 3 + k] for k in range(3)]
    assert np.allclose(kd, [0.1, 0.2, 0.3])
    l = d.lights[0]
    # l2w = translate(from) * ctm applied to the origin: from + ctm translation = (1,1,1) + (1,2,3)
    assert l.type == P.capi.PT_LIGHT_POINT and np.allclose(list(l.v), [2, 3, 4])


def test_front_end_plymesh(P, tmp_path):
    """Shape "plymesh" (shapes/plymesh.rs): ascii and binary PLY, normals + uv, a quad split as (0,1,2) (3,0,2), file name
    relative to the including scene file; the flattened scene equals the same mesh given as a trianglemesh."""
    import struct
    verts = [(0, 0, 0, 0, 0, 1, 0, 0), (1, 0, 0, 0, 0, 1, 1, 0), (1, 1, 0.25, 0, 0.1, 1, 1, 1), (0, 1, 0, 0, 0, 1, 0, 1), (2, 0, 0, 0, 0, 1, 0.5, 0.5)]
    faces = [(0, 1, 2, 3), (1, 4, 2)]
    hdr = "ply\nformat %s 1.0\ncomment test\nelement vertex 5\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\n" \
          "property float ny\nproperty float nz\nproperty float u\nproperty float v\nelement face 2\nproperty list uchar int vertex_indices\nend_header\n"
    (tmp_path / "geo").mkdir()
    with open(tmp_path / "geo" / "a.ply", "w") as f:
        f.write(hdr % "ascii")
        for v in verts:
            f.write(" ".join(repr(float(x)) for x in v) + "\n")
        for fc in faces:
            f.write("%d %s\n" % (len(fc), " ".join(map(str, fc))))
    for name, fmt, e in (("l.ply", "binary_little_endian", "<"), ("b.ply", "binary_big_endian", ">")):
        with open(tmp_path / "geo" / name, "wb") as f:
            f.write((hdr % fmt).encode())
            for v in verts:
                f.write(struct.pack(e + "8f", *v))
            for fc in faces:
                f.write(struct.pack(e + "B%di" % len(fc), len(fc), *fc))
    head = 'Camera "perspective"\nWorldBegin\nMaterial "matte"\nTranslate 0 0 3\nRotate 20 0 1 0\n'
    tm = head + 'Shape "trianglemesh" "integer indices" [0 1 2 3 0 2 1 4 2] "point P" [%s] "normal N" [%s] "float uv" [%s]\nWorldEnd\n' % (
        " ".join(repr(float(x)) for v in verts for x in v[:3]), " ".join(repr(float(x)) for v in verts for x in v[3:6]),
        " ".join(repr(float(x)) for v in verts for x in v[6:]))
    ref_hs = P.HostScene.parse_string(tm)       # keep the owner alive: desc points into it
    ref = ref_hs.desc
    assert ref.n_tris == 3
    (tmp_path / "inc.pbrt").write_text('Shape "plymesh" "string filename" ["geo/a.ply"]\n')
    for name in ("a.ply", "l.ply", "b.ply"):
        scene = tmp_path / ("s_" + name + ".pbrt")
        scene.write_text(head + 'Shape "plymesh" "string filename" ["geo/%s"]\nWorldEnd\n' % name)
        hs = P.HostScene.parse_file(str(scene))
        d = hs.desc
        assert d.n_tris == 3
        for fld, k in (("tri_p", 9), ("tri_n", 9), ("tri_uv", 6)):
            a, b = np.ctypeslib.as_array(getattr(d, fld), (3 * k,)), np.ctypeslib.as_array(getattr(ref, fld), (3 * k,))
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, fld)
        assert list(np.ctypeslib.as_array(d.tri_flags, (3,))) == list(np.ctypeslib.as_array(ref.tri_flags, (3,)))
    (tmp_path / "s_inc.pbrt").write_text(head + 'Include "inc.pbrt"\nWorldEnd\n')
    assert P.HostScene.parse_file(str(tmp_path / "s_inc.pbrt")).desc.n_tris == 3
    with open(tmp_path / "geo" / "bad.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n"
                "0 0 0\n1 0 0\n0 1 0\n5 0 1 2 0 1\n")
    (tmp_path / "s_bad.pbrt").write_text(head + 'Shape "plymesh" "string filename" ["geo/bad.ply"]\nWorldEnd\n')
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_file(str(tmp_path / "s_bad.pbrt"))
    (tmp_path / "s_missing.pbrt").write_text(head + 'Shape "plymesh" "string filename" ["geo/none.ply"]\nWorldEnd\n')
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_file(str(tmp_path / "s_missing.pbrt"))


def test_front_end_constant_foldable_textures(P):
    # textures/{constant,scale,mix}.rs: value1 * value2 ; t1 * (1 - amt) + t2 * amt, folded on the host
    txt = ('Camera "perspective"\nWorldBegin\n'
           'Texture "a" "spectrum" "constant" "rgb value" [0.2 0.4 0.8]\nTexture "f" "float" "constant" "float value" [0.25]\n'
           'Texture "s" "spectrum" "scale" "texture tex1" "a" "rgb tex2" [0.5 0.5 0.25]\n'
           'Texture "m" "spectrum" "mix" "texture tex1" "a" "texture tex2" "s" "texture amount" "f"\n'
           'Texture "r" "float" "scale" "texture tex1" "f" "float tex2" [0.4]\n'
           'Material "plastic" "texture Kd" "m" "texture Ks" "s" "texture roughness" "r"\n'
           'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]\nWorldEnd\n')
    hs = P.HostScene.parse_string(txt)
    d = hs.desc
    m = d.materials[0]
    sp = np.ctypeslib.as_array(d.spectra, (d.n_spectra * 3,)).reshape(-1, 3)
    f = np.float32
    a, t2 = np.array([0.2, 0.4, 0.8], f), np.array([0.5, 0.5, 0.25], f)
    s_ = a * t2
    mix = a * (f(1.0) - f(0.25)) + s_ * f(0.25)
    assert np.array_equal(sp[m.kd], mix) and np.array_equal(sp[m.ks], s_)
    assert m.uroughness == f(0.25) * f(0.4)
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_string(txt.replace('"spectrum" "scale"', '"spectrum" "imagemap"'))


def test_front_end_map_lights(P, tmp_path):
    """projection.rs:21-62, 176-195 / goniometric.rs:143-159: p_light, I * scale, the slide's screen bounds from its aspect,
    Transform::perspective(fov, 1e-3, 1e30), cos_total_width from the corner direction"""
    from test_hostcheck import write_pfm
    write_pfm(str(tmp_path / "slide.pfm"), np.full((4, 10, 3), 0.5, np.float32))
    base = 'Camera "perspective"\nWorldBegin\nTranslate 1 2 3\n%s\nWorldEnd\n'
    hs = P.HostScene.parse_string(base % ('LightSource "projection" "rgb I" [2 3 4] "rgb scale" [2 1 1] "float fov" [60] "string mapname" ["%s"]' % (tmp_path / "slide.pfm")))
    L, E = hs.desc.lights[0], hs.desc.envmaps[0]
    assert L.type == P.capi.PT_LIGHT_PROJECTION and list(L.v) == [1, 2, 3] and L.tri == 0
    assert list(hs.desc.spectra[3 * L.spectrum:3 * L.spectrum + 3]) == [4, 3, 4]
    assert (E.width, E.height) == (10, 4) and list(E.screen_bounds) == [-2.5, -1, 2.5, 1] and E.hitcher == np.float32(1e-3)
    t = np.tan(np.radians(30.0))
    corner = np.array([2.5 * t, t, 1.0])
    assert abs(E.cos_total_width - 1.0 / np.linalg.norm(corner)) < 1e-6
    m = np.array(list(E.xf), np.float64).reshape(4, 4)
    q = m @ np.array([corner[0], corner[1], 1.0, 1.0])
    assert np.allclose(q[:2] / q[3], [2.5, 1.0], atol=1e-5)            # the corner direction projects onto the screen corner
    hs = P.HostScene.parse_string(base % 'LightSource "projection"')    # no map: one white texel, aspect 1, fov 45
    E = hs.desc.envmaps[0]
    assert (E.width, E.height) == (1, 1) and list(E.screen_bounds) == [-1, -1, 1, 1] and list(E.texels[0:3]) == [1, 1, 1]
    hs = P.HostScene.parse_string(base % 'LightSource "goniometric" "rgb L" [5 6 7]')
    L, E = hs.desc.lights[0], hs.desc.envmaps[0]
    assert L.type == P.capi.PT_LIGHT_GONIOMETRIC and list(L.v) == [1, 2, 3]
    m = np.array(list(E.xf), np.float32).reshape(4, 4)
    assert np.array_equal(m, np.array([[1, 0, 0, -1], [0, 1, 0, -2], [0, 0, 1, -3], [0, 0, 0, 1]], np.float32))   # world_to_light, translation included
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_string(base % 'LightSource "projection" "string mapname" ["missing.pfm"]')


def test_front_end_mix_material(P):
    """scene_context.rs:446-451, mix.rs:96-103: two named materials (matte from the mix's own parameters for an unknown
    name), "amount" default 0.5; identical sub-materials are shared"""
    base = ('Camera "perspective"\nWorldBegin\nMakeNamedMaterial "a" "string type" ["plastic"] "rgb Kd" [0.5 0.3 0.2]\n'
            'MakeNamedMaterial "b" "string type" ["mirror"]\n%s\nShape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]\nWorldEnd\n')
    hs = P.HostScene.parse_string(base % 'Material "mix" "string namedmaterial1" ["a"] "string namedmaterial2" ["b"] "rgb amount" [0.2 0.4 0.6]')
    d = hs.desc
    M = d.materials[d.tri_material[0]]
    assert M.type == P.capi.PT_MAT_MIX and d.materials[M.kd].type == P.capi.PT_MAT_PLASTIC and d.materials[M.ks].type == P.capi.PT_MAT_MIRROR
    assert np.allclose(list(d.spectra[3 * M.kr:3 * M.kr + 3]), [0.2, 0.4, 0.6])
    hs = P.HostScene.parse_string(base % 'Material "mix" "string namedmaterial1" ["zzz"] "string namedmaterial2" ["b"] "rgb Kd" [0.1 0.2 0.3]')
    d = hs.desc
    M = d.materials[d.tri_material[0]]
    sub = d.materials[M.kd]
    assert sub.type == P.capi.PT_MAT_MATTE and np.allclose(list(d.spectra[3 * sub.kd:3 * sub.kd + 3]), [0.1, 0.2, 0.3])
    assert np.allclose(list(d.spectra[3 * M.kr:3 * M.kr + 3]), [0.5, 0.5, 0.5])
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_string(base % 'MakeNamedMaterial "n" "string type" ["none"]\nMaterial "mix" "string namedmaterial1" ["n"] "string namedmaterial2" ["b"]')


def test_front_end_stratified_sampler(P):
    """create_stratified_sampler (stratified.rs:130-150): defaults 4 x 4, jitter, 4 dimensions; --quick = 1 x 1"""
    base = 'Camera "perspective"\n%s\nWorldBegin\nWorldEnd\n'
    S = P.HostScene.parse_string(base % 'Sampler "stratified"').desc.sampler
    assert (S.type, S.samples_per_pixel, S.strata, S.jitter, S.n_sampled_dimensions) == (P.capi.PT_SAMPLER_STRATIFIED, 16, 4, 1, 4)
    S = P.HostScene.parse_string(base % 'Sampler "stratified" "integer xsamples" [3] "integer ysamples" [3] "bool jitter" ["false"] "integer dimensions" [7]').desc.sampler
    assert (S.samples_per_pixel, S.strata, S.jitter, S.n_sampled_dimensions) == (9, 3, 0, 7)
    Z = P.HostScene.parse_string(base % 'Sampler "02sequence"').desc.sampler                      # zerotwosequence.rs:107-125
    assert (Z.type, Z.samples_per_pixel, Z.n_sampled_dimensions) == (P.capi.PT_SAMPLER_ZEROTWO, 16, 4)
    Z = P.HostScene.parse_string(base % 'Sampler "lowdiscrepancy" "integer pixelsamples" [6] "integer dimensions" [2]').desc.sampler
    assert (Z.type, Z.samples_per_pixel, Z.n_sampled_dimensions) == (P.capi.PT_SAMPLER_ZEROTWO, 6, 2)   # rounded up to 8 by the sampler itself
    with pytest.raises(P.PbrtError):       # the reference's 2D stratification loops x over ysamples: only the square case is reproduced
        P.HostScene.parse_string(base % 'Sampler "stratified" "integer xsamples" [2] "integer ysamples" [3]')
    lib = P.capi.load()
    lib.pt_set_options(P.capi.pt_options(1, 0))
    try:
        assert P.HostScene.parse_string(base % 'Sampler "stratified"').desc.sampler.samples_per_pixel == 1
    finally:
        lib.pt_set_options(P.capi.pt_options(0, 0))


def test_front_end_cameras(P):
    """orthographic.rs:193-253 / environment.rs:90-151: same parameters as the perspective camera; Transform::orthographic(0, 1)
    is the identity, so raster_to_camera maps the raster onto the screen window at z = 0"""
    base = '%s\nFilm "image" "integer xresolution" [40] "integer yresolution" [20]\nWorldBegin\nWorldEnd\n'
    d = P.HostScene.parse_string(base % 'Camera "orthographic"').desc.camera
    assert d.type == P.capi.PT_CAMERA_ORTHOGRAPHIC and d.lens_radius == 0.0 and d.focal_distance == np.float32(1e6)
    m = np.array(list(d.raster_to_camera), np.float32).reshape(4, 4)
    corner = lambda x, y: (m @ np.array([x, y, 0, 1], np.float32))[:3]
    assert np.allclose(corner(0, 0), [-2, 1, 0], atol=1e-6) and np.allclose(corner(40, 20), [2, -1, 0], atol=1e-6)   # frame = 2: window (-2, 2) x (-1, 1), y flipped
    d = P.HostScene.parse_string(base % 'Camera "orthographic" "float screenwindow" [-3 5 -1 7] "float lensradius" [0.5]').desc.camera
    m = np.array(list(d.raster_to_camera), np.float32).reshape(4, 4)
    assert np.allclose(corner(0, 0), [-3, 7, 0], atol=1e-6) and np.allclose(corner(40, 20), [5, -1, 0], atol=1e-6) and d.lens_radius == 0.5
    d = P.HostScene.parse_string(base % 'Camera "environment" "float lensradius" [0.5] "float shutterclose" [0.25]').desc.camera
    assert d.type == P.capi.PT_CAMERA_ENVIRONMENT and d.lens_radius == 0.0 and d.shutter_close == 0.25
    assert P.HostScene.parse_string(base % 'Camera "perspective"').desc.camera.type == P.capi.PT_CAMERA_PERSPECTIVE
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_string(base % 'Camera "realistic"')


def test_front_end_ao_integrator(P):
    base = 'Camera "perspective"\n%sIntegrator "ambientocclusion" %s\nWorldBegin\nWorldEnd\n'
    d = P.HostScene.parse_string(base % ("", "")).desc.integrator
    assert (d.type, d.ao_samples, d.ao_cos_sample) == (P.capi.PT_INTEGRATOR_AO, 64, 1)                      # ao.rs:124-125 defaults
    d = P.HostScene.parse_string(base % ("", '"integer nsamples" [7] "bool cossample" ["false"]')).desc.integrator
    assert (d.ao_samples, d.ao_cos_sample) == (7, 0)
    assert P.HostScene.parse_string(base % ('Sampler "random"\n', "")).desc.sampler.type == P.capi.PT_SAMPLER_RANDOM
    lib = P.capi.load()
    lib.pt_set_options(P.capi.pt_options(1, 0))
    try:
        assert P.HostScene.parse_string(base % ("", '"integer nsamples" [7]')).desc.integrator.ao_samples == 1  # ao.rs:128-130
    finally:
        lib.pt_set_options(P.capi.pt_options(0, 0))


def test_front_end_tree_integrators(P):
    base = 'Camera "perspective"\n%sIntegrator %s\nWorldBegin\nAttributeBegin\nAreaLightSource "diffuse" %s\nShape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]\nAttributeEnd\nLightSource "point"\nWorldEnd\n'
    hs = P.HostScene.parse_string(base % ("", '"directlighting"', '"integer nsamples" [6]'))
    d = hs.desc.integrator
    assert (d.type, d.direct_strategy, d.max_depth) == (P.capi.PT_INTEGRATOR_DIRECT, P.capi.PT_DIRECT_ALL, 5)   # directlighting.rs:140-141
    assert [hs.desc.lights[i].n_samples for i in range(2)] == [6, 1]                                            # diffuse.rs:177, light.rs:76-78
    d = P.HostScene.parse_string(base % ("", '"directlighting" "string strategy" ["one"] "integer maxdepth" [3]', "")).desc.integrator
    assert (d.direct_strategy, d.max_depth) == (P.capi.PT_DIRECT_ONE, 3)
    d = P.HostScene.parse_string(base % ("", '"whitted"', "")).desc.integrator
    assert (d.type, d.max_depth) == (P.capi.PT_INTEGRATOR_WHITTED, 5)
    for bad in ('"whitted" "integer pixelbounds" [0 1 0 1]',):
        with pytest.raises(P.PbrtError):
            P.HostScene.parse_string(base % ("", bad, ""))
    for ok in ('"whitted"', '"directlighting" "string strategy" ["one"]', '"directlighting"', '"ambientocclusion"'):
        assert P.HostScene.parse_string(base % ('Sampler "random"\n', ok, "")).desc.sampler.type == P.capi.PT_SAMPLER_RANDOM
    lib = P.capi.load()
    lib.pt_set_options(P.capi.pt_options(1, 0))
    try:
        assert P.HostScene.parse_string(base % ("", '"directlighting"', '"integer nsamples" [6]')).desc.lights[0].n_samples == 1   # diffuse.rs:181-183
    finally:
        lib.pt_set_options(P.capi.pt_options(0, 0))


def _loop_reference(P0, F0, levels):
    """Independent Loop subdivision (closed or open manifold triangle meshes; pbrt-v3 rules: beta = 3/16 for valence 3, else
    3 / (8 valence); regular interior vertices 1/16; boundary 1/8; odd vertices 3/8 3/8 1/8 1/8 or 1/2 1/2), then the limit
    positions IN PLACE in vertex order like pbrt-r3 (loopsubdiv.rs:609-619).  float64, dictionary adjacency."""
    V, F = [np.array(p, np.float64) for p in P0], [tuple(f) for f in F0]
    beta = lambda n: 3.0 / 16.0 if n == 3 else 3.0 / (8.0 * n)

    def adjacency(V, F):
        nb, edge_faces = [set() for _ in V], {}
        for fi, f in enumerate(F):
            for k in range(3):
                a, b = f[k], f[(k + 1) % 3]
                nb[a].add(b); nb[b].add(a)
                edge_faces.setdefault((min(a, b), max(a, b)), []).append(fi)
        bverts = [set() for _ in V]
        for (a, b), fs in edge_faces.items():
            if len(fs) == 1:
                bverts[a].add(b); bverts[b].add(a)
        return nb, edge_faces, bverts
    for _ in range(levels):
        nb, edge_faces, bverts = adjacency(V, F)
        newV = []
        for i, p in enumerate(V):
            if bverts[i]:
                q = list(bverts[i])
                newV.append((1 - 2 / 8.0) * p + (V[q[0]] + V[q[1]]) / 8.0)
            else:
                n = len(nb[i])
                b = 1.0 / 16.0 if n == 6 else beta(n)
                newV.append((1 - n * b) * p + b * sum(V[j] for j in nb[i]))
        emap = {}
        for (a, b), fs in edge_faces.items():
            if len(fs) == 1:
                q = 0.5 * V[a] + 0.5 * V[b]
            else:
                o = [next(v for v in F[fi] if v != a and v != b) for fi in fs]
                q = 3 / 8.0 * (V[a] + V[b]) + 1 / 8.0 * (V[o[0]] + V[o[1]])
            emap[(a, b)] = len(newV)
            newV.append(q)
        newF = []
        for f in F:
            e = [emap[(min(f[k], f[(k + 1) % 3]), max(f[k], f[(k + 1) % 3]))] for k in range(3)]
            newF += [(f[0], e[0], e[2]), (e[0], f[1], e[1]), (e[2], e[1], f[2]), (e[0], e[1], e[2])]
        V, F = newV, newF
    nb, edge_faces, bverts = adjacency(V, F)
    return V, F, nb, bverts


def test_front_end_heightfield_and_loopsubdiv(P):
    """Shapes that are triangle meshes by construction (shapes/heightfield.rs, shapes/loopsubdiv.rs)."""
    base = 'Camera "perspective"\nWorldBegin\n%s\nWorldEnd\n'
    tri = lambda hs: np.ctypeslib.as_array(hs.desc.tri_p, (hs.desc.n_tris * 9,)).reshape(-1, 3, 3).copy()
    hs = P.HostScene.parse_string(base % 'Shape "heightfield" "integer nu" [3] "integer nv" [2] "float Pz" [0 1 0 0.5 0.5 0.5]')
    assert hs.desc.n_tris == 4 and bool(hs.desc.tri_uv)
    assert np.array_equal(tri(hs)[0], [[0, 0, 0], [0.5, 0, 1], [0.5, 1, 0.5]]) and np.array_equal(tri(hs)[1], [[0, 0, 0], [0.5, 1, 0.5], [0, 1, 0.5]])   # heightfield.rs:51-58
    uv = np.ctypeslib.as_array(hs.desc.tri_uv, (4 * 6,)).reshape(-1, 3, 2)
    assert np.array_equal(uv[0], [[0, 0], [0.5, 0], [0.5, 1]])
    for bad in ('Shape "heightfield" "integer nu" [3] "float Pz" [0 1 0]', 'Shape "heightfield" "integer nu" [3] "integer nv" [2] "float Pz" [0 1 0]'):
        with pytest.raises(P.PbrtError):
            P.HostScene.parse_string(base % bad)
    # Loop subdivision: octahedron (closed, valence 4), a fan with a valence-3 and a valence-6 vertex... and an open square (boundary rules)
    oct_p = [(0, 0, 1), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (0, -1, 0), (0, 0, -1)]
    oct_f = [(0, 1, 2), (0, 2, 3), (0, 3, 4), (0, 4, 1), (5, 2, 1), (5, 3, 2), (5, 4, 3), (5, 1, 4)]
    tet_p = [(1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)]
    tet_f = [(0, 1, 2), (0, 3, 1), (0, 2, 3), (1, 3, 2)]
    sq_p = [(0, 0, 0), (1, 0, 0.2), (1, 1, 0), (0, 1, 0.3), (0.5, 0.5, 0.5)]
    sq_f = [(0, 1, 4), (1, 2, 4), (2, 3, 4), (3, 0, 4)]
    for name, pts, faces in (("octahedron", oct_p, oct_f), ("tetrahedron", tet_p, tet_f), ("open pyramid", sq_p, sq_f)):
        for levels in (0, 1, 3):
            txt = 'Shape "loopsubdiv" "integer levels" [%d] "integer indices" [%s] "point P" [%s]' % (
                levels, " ".join(str(i) for f in faces for i in f), " ".join("%g" % c for p in pts for c in p))
            hs = P.HostScene.parse_string(base % txt)
            assert hs.desc.n_tris == len(faces) * 4 ** levels and bool(hs.desc.tri_n), (name, levels)
            T = tri(hs)
            V, F, nb, bverts = _loop_reference(pts, faces, levels)
            # the refined connectivity, then pbrt-r3's in-place limit positions in vertex order: compare as point SETS (the
            # vertex numbering of the two implementations differs)
            # the two implementations number vertices and faces alike (children of the old vertices, then the edge vertices in
            # first-seen order; four children per face): pbrt-r3's IN-PLACE limit positions (loopsubdiv.rs:609-619), vertex by vertex
            L = [v.copy() for v in V]
            for i in range(len(L)):
                if bverts[i]:
                    q = list(bverts[i])
                    L[i] = (1 - 2 / 5.0) * L[i] + (L[q[0]] + L[q[1]]) / 5.0
                else:
                    nv = len(nb[i])
                    b = 3.0 / 16.0 if nv == 3 else 3.0 / (8.0 * nv)
                    gam = 1.0 / (nv + 3.0 / (8.0 * b))
                    L[i] = (1 - nv * gam) * L[i] + gam * sum(L[j] for j in nb[i])
            want = np.array([[L[i] for i in f] for f in F])
            assert np.allclose(T, want, atol=2e-6), (name, levels, np.abs(T - want).max())
            lo, hi = np.array(pts).min(0), np.array(pts).max(0)
            assert (T.reshape(-1, 3) >= lo - 1e-6).all() and (T.reshape(-1, 3) <= hi + 1e-6).all()     # convex-hull property of the limit surface
            # watertightness is kept: every edge of a closed mesh is shared by exactly two triangles
            key = lambda p: tuple(np.round(p, 6))
            edges = {}
            for t in T:
                for k in range(3):
                    e = tuple(sorted((key(t[k]), key(t[(k + 1) % 3]))))
                    edges[e] = edges.get(e, 0) + 1
            counts = sorted(set(edges.values()))
            assert counts == ([2] if name != "open pyramid" else [1, 2]), (name, levels, counts)
            n = np.ctypeslib.as_array(hs.desc.tri_n, (hs.desc.n_tris * 9,)).reshape(-1, 3)
            assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    # level 0 of the octahedron = the limit positions of the control vertices, applied in place in vertex order (loopsubdiv.rs:609-619)
    hs = P.HostScene.parse_string(base % ('Shape "loopsubdiv" "integer levels" [0] "integer indices" [%s] "point P" [%s]' % (
        " ".join(str(i) for f in oct_f for i in f), " ".join("%g" % c for p in oct_p for c in p))))
    g = 1.0 / (4 + 3.0 / (8.0 * (3.0 / 32.0)))
    V = [np.array(p, np.float64) for p in oct_p]
    nbr = {0: [1, 2, 3, 4], 1: [0, 2, 4, 5], 2: [0, 1, 3, 5], 3: [0, 2, 4, 5], 4: [0, 1, 3, 5], 5: [1, 2, 3, 4]}
    for i in range(6):
        V[i] = (1 - 4 * g) * V[i] + g * sum(V[j] for j in nbr[i])
    want = np.array([[V[i] for i in f] for f in oct_f])
    assert np.allclose(tri(hs), want, atol=1e-6)
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_string(base % 'Shape "loopsubdiv" "point P" [0 0 0 1 0 0 0 1 0]')


def test_front_end_nurbs(P):
    """Shape "nurbs" (shapes/nurbs.rs): a rational tensor-product B-spline surface diced into diceu x dicev vertices; points
    against scipy's B-spline basis, normals against finite differences of the same surface."""
    from scipy.interpolate import BSpline
    base = 'Camera "perspective"\nWorldBegin\n%s\nWorldEnd\n'
    rng = np.random.RandomState(2)
    nu, nv, uo, vo = 5, 4, 4, 3
    uk, vk = np.array([0, 0, 0, 0, 1, 2, 2, 2, 2], np.float64), np.array([0, 0, 0, 1, 2, 2, 2], np.float64)
    cp, w = rng.uniform(-1, 1, (nv, nu, 3)), rng.uniform(0.5, 2, (nv, nu, 1))
    pw = np.concatenate([cp * w, w], 2).astype(np.float32)
    du, dv = 7, 6
    txt = 'Shape "nurbs" "integer nu" [%d] "integer uorder" [%d] "float uknots" [%s] "integer nv" [%d] "integer vorder" [%d] "float vknots" [%s] "point4 Pw" [%s] "integer diceu" [%d] "integer dicev" [%d]' % (
        nu, uo, " ".join(map(str, uk)), nv, vo, " ".join(map(str, vk)), " ".join("%.9g" % x for x in pw.reshape(-1)), du, dv)
    hs = P.HostScene.parse_string(base % txt)
    assert hs.desc.n_tris == 2 * (du - 1) * (dv - 1) and bool(hs.desc.tri_n) and bool(hs.desc.tri_uv)
    T = np.ctypeslib.as_array(hs.desc.tri_p, (hs.desc.n_tris * 9,)).reshape(-1, 3, 3)
    N = np.ctypeslib.as_array(hs.desc.tri_n, (hs.desc.n_tris * 9,)).reshape(-1, 3, 3)
    UV = np.ctypeslib.as_array(hs.desc.tri_uv, (hs.desc.n_tris * 6,)).reshape(-1, 3, 2)

    def basis(k, order, n, t):
        t = min(max(t, k[order - 1] + 1e-9), k[n] - 1e-9)
        return np.nan_to_num(np.array([BSpline.basis_element(k[i:i + order + 1], extrapolate=False)(t) if k[i] != k[i + order] else 0.0 for i in range(n)]))

    def surf(u, v):
        h = np.einsum("j,i,jik->k", basis(vk, vo, nv, v), basis(uk, uo, nu, u), pw.astype(np.float64))
        return h[:3] / h[3]
    for t in range(0, hs.desc.n_tris, 3):
        for k in range(3):
            u, v = UV[t, k]
            assert np.allclose(T[t, k], surf(u, v), atol=2e-6), (t, k)
            if 0.05 < u < 1.95 and 0.05 < v < 1.95:
                e = 1e-4
                n = np.cross(surf(u + e, v) - surf(u - e, v), surf(u, v + e) - surf(u, v - e))
                assert np.dot(n / np.linalg.norm(n), N[t, k]) > 0.9999, (t, k)
    assert UV.min() == 0.0 and UV.max() == 2.0           # u0 / u1 default to the knot range (nurbs.rs:357-365)
    # non-rational "P", a sub-range, the default 30 x 30 dicing
    flat = 'Shape "nurbs" "integer nu" [2] "integer uorder" [2] "float uknots" [0 0 1 1] "integer nv" [2] "integer vorder" [2] "float vknots" [0 0 1 1] "point P" [0 0 0 1 0 0 0 1 0 1 1 1] "float u0" [0.25] "float u1" [2]'
    hs = P.HostScene.parse_string(base % flat)
    assert hs.desc.n_tris == 2 * 29 * 29
    UV = np.ctypeslib.as_array(hs.desc.tri_uv, (hs.desc.n_tris * 6,)).reshape(-1, 2)
    assert UV[:, 0].min() == 0.25 and UV[:, 0].max() == 1.0
    Tp = np.ctypeslib.as_array(hs.desc.tri_p, (hs.desc.n_tris * 9,)).reshape(-1, 3)
    assert np.allclose(Tp[:, 2], Tp[:, 0] * Tp[:, 1], atol=1e-6)            # the bilinear patch z = x y
    for bad in ('Shape "nurbs" "integer nu" [2]', flat.replace('"float uknots" [0 0 1 1]', '"float uknots" [0 0 1]')):
        with pytest.raises(P.PbrtError):
            P.HostScene.parse_string(base % bad)


def test_front_end_targz_and_gz_scenes(P, tmp_path):
    """pbrt_parse_file on a ".tar.gz" scene (parser.rs:32-63: unpacked into a temporary directory, first .pbrt file) whose
    main file Includes a gzipped file (read_file.rs:38-54); same scene as the plain text."""
    import gzip
    import tarfile
    txt = P.scenes.cornell_box(res=16, spp=1)
    head, rest = txt.split("WorldBegin\n", 1)
    os.makedirs(tmp_path / "scene" / "geo")
    (tmp_path / "scene" / "main.pbrt").write_text(head + 'WorldBegin\nInclude "geo/world.pbrt.gz"\n')
    with gzip.open(tmp_path / "scene" / "geo" / "world.pbrt.gz", "wt") as f:
        f.write(rest)
    with tarfile.open(tmp_path / "scene.tar.gz", "w:gz") as t:
        t.add(tmp_path / "scene", arcname="a_rather_long_directory_name_" * 5 + "/scene")     # GNU long names
    a, b, c = P.HostScene.parse_string(txt), P.HostScene.parse_file(str(tmp_path / "scene.tar.gz")), P.HostScene.parse_file(str(tmp_path / "scene" / "main.pbrt"))
    for x in (b, c):
        assert x.desc.n_tris == a.desc.n_tris == 32 and x.desc.n_lights == a.desc.n_lights
        assert np.array_equal(np.ctypeslib.as_array(x.desc.tri_p, (32 * 9,)), np.ctypeslib.as_array(a.desc.tri_p, (32 * 9,)))
    assert not [n for n in os.listdir("/tmp") if n.startswith("pbrt_r3_b200_")]          # the temporary directory is gone
    (tmp_path / "empty.tar.gz").write_bytes(gzip.compress(b"\0" * 1024))
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_file(str(tmp_path / "empty.tar.gz"))


def test_front_end_blackbody_and_spectrum_files(P, tmp_path):
    """ "blackbody" parameters (scene_context.rs:763-778 -> SampledSpectrum::from_blackbody, sampled.rs:241-251, blackbody.rs)
    against a float64 restatement; "spectrum" FILE parameters keep the parameter's default, like the reference (which
    files the loaded spectrum under the file's name, scene_context.rs:744-758)."""
    base = 'Camera "perspective"\nWorldBegin\nAttributeBegin\nAreaLightSource "diffuse" %s\nShape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]\nAttributeEnd\nWorldEnd\n'

    def light_spectrum(hs):
        d, nc = hs.desc, hs.desc.n_channels
        return np.array([d.spectra[d.lights[0].spectrum * nc + k] for k in range(nc)], np.float64)

    def planck(lam_nm, t):
        c, h, kb = 299792458.0, 6.62606957e-34, 1.3806488e-23
        l = lam_nm * 1e-9
        return (2 * h * c * c) / (l ** 5 * (np.exp((h * c) / (l * kb * t)) - 1))

    def bins60(pairs):
        lam = np.arange(360, 831, dtype=np.float64)
        out = np.zeros(60)
        for t, sc in pairs:
            v = planck(lam, t) / planck(2.8977721e-3 / t * 1e9, t)
            fine = np.linspace(400, 700, 60 * 200 + 1)
            vf = np.interp(fine, lam, v)
            out += np.array([np.trapezoid(vf[i * 200:(i + 1) * 200 + 1], fine[i * 200:(i + 1) * 200 + 1]) / 5.0 for i in range(60)]) * sc
        return out
    for pairs in ([(6500.0, 1.0)], [(3000.0, 2.5)], [(6500.0, 1.0), (3000.0, 0.5)]):
        txt = base % ('"blackbody L" [%s]' % " ".join("%g %g" % p for p in pairs))
        ref = bins60(pairs)
        got = light_spectrum(P.HostScene.parse_string(txt, spectrum="sampled"))
        assert np.allclose(got, ref, rtol=2e-4), (pairs, np.abs(got / ref - 1).max())
        rgb = light_spectrum(P.HostScene.parse_string(txt))
        # RGB build: SampledSpectrum::to_rgb of the same bins; the 60-bin X/Y/Z tables are the library's own data: compare through Y
        y = 0.212671 * rgb[0] + 0.715160 * rgb[1] + 0.072169 * rgb[2]
        assert 0.5 * ref.mean() < y < 1.5 * ref.max() and rgb.min() > 0
    assert light_spectrum(P.HostScene.parse_string(base % '"blackbody L" [3000 1]'))[0] > 2 * light_spectrum(P.HostScene.parse_string(base % '"blackbody L" [3000 1]'))[2]   # warm light
    (tmp_path / "x.spd").write_text("400 0.1\n700 0.9\n")
    hs = P.HostScene.parse_string(base % ('"spectrum L" ["%s"]' % (tmp_path / "x.spd")))
    assert list(light_spectrum(hs)) == [1.0, 1.0, 1.0]            # the default of "L" (diffuse.rs:175)
    assert any("spectrum file" in w for w in hs.warnings)
    hs = P.HostScene.parse_string(base % '"spectrum L" ["missing.spd"]')
    assert any("Unable to read SPD file" in w for w in hs.warnings)
    hs = P.HostScene.parse_string(base % '"spectrum L" [400 2 700 2]')        # inline samples do take effect (param_set.rs add_sampled_spectrum)
    v = light_spectrum(hs)
    assert abs(0.212671 * v[0] + 0.715160 * v[1] + 0.072169 * v[2] - 2.0) < 0.02      # an equal-energy spectrum of 2: Y = 2


def test_front_end_extra_materials(P):
    base = 'Camera "perspective"\nWorldBegin\nMaterial %s\nShape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]\nWorldEnd\n'
    sp = lambda hs, i: [hs.desc.spectra[3 * i + k] for k in range(3)]
    hs = P.HostScene.parse_string(base % '"uber"')
    m = hs.desc.materials[0]                                                                            # uber.rs:142-153 defaults
    assert m.type == P.capi.PT_MAT_UBER and (m.uroughness, m.vroughness, m.eta, m.remap_roughness) == (np.float32(0.1), np.float32(0.1), 1.5, 1)
    assert [sp(hs, i) for i in (m.kd, m.ks, m.kr, m.kt, m.opacity)] == [[0.25] * 3, [0.25] * 3, [0.0] * 3, [0.0] * 3, [1.0] * 3]
    m = P.HostScene.parse_string(base % '"uber" "float roughness" [0.3] "float vroughness" [0.2] "float index" [1.33]').desc.materials[0]
    assert (m.uroughness, m.vroughness, m.eta) == (np.float32(0.3), np.float32(0.2), np.float32(1.33))
    hs = P.HostScene.parse_string(base % '"substrate"')
    m = hs.desc.materials[0]                                                                            # substrate.rs:62-68
    assert m.type == P.capi.PT_MAT_SUBSTRATE and sp(hs, m.kd) == [0.5] * 3 and sp(hs, m.ks) == [0.5] * 3 and (m.uroughness, m.vroughness) == (np.float32(0.1),) * 2
    hs = P.HostScene.parse_string(base % '"translucent" "rgb transmit" [0.1 0.2 0.3]')
    m = hs.desc.materials[0]                                                                            # translucent.rs:102-109
    assert m.type == P.capi.PT_MAT_TRANSLUCENT and sp(hs, m.kd) == [0.25] * 3 and sp(hs, m.kr) == [0.5] * 3
    assert np.allclose(sp(hs, m.kt), [0.1, 0.2, 0.3]) and m.eta == 1.5
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_string(base % '"uber"', spectrum="sampled")


def test_front_end_infinite_light(P, tmp_path):
    from test_hostcheck import write_pfm
    img = np.arange(4 * 8 * 3, dtype=np.float32).reshape(4, 8, 3) - 5.0
    write_pfm(str(tmp_path / "e.pfm"), img)
    base = 'Camera "perspective"\nWorldBegin\nRotate 90 0 0 1\nLightSource "infinite" "rgb L" [2 1 0.5] %s\nWorldEnd\n'
    hs = P.HostScene.parse_string(base % '"integer samples" [8]')
    d = hs.desc
    assert (d.n_lights, d.n_envmaps, d.lights[0].type, d.lights[0].tri, d.lights[0].n_samples) == (1, 1, P.capi.PT_LIGHT_INFINITE, 0, 8)   # infinite.rs:279
    e = d.envmaps[0]
    assert (e.width, e.height) == (1, 1) and [e.texels[k] for k in range(3)] == [2.0, 1.0, 0.5]                      # make_mipmap without a map (:59-64)
    assert np.allclose(np.array(e.l2w[:9]).reshape(3, 3), [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-7)             # the CTM's rotation
    assert np.allclose(np.array(e.l2w[:9]).reshape(3, 3).T, np.array(e.w2l[:9]).reshape(3, 3), atol=1e-7)
    hs = P.HostScene.parse_string(base % ('"string mapname" ["%s"]' % (tmp_path / "e.pfm")))
    e = hs.desc.envmaps[0]
    assert (e.width, e.height) == (8, 4)
    tex = np.array([e.texels[k] for k in range(4 * 8 * 3)], np.float32).reshape(4, 8, 3)
    assert np.array_equal(tex, np.maximum(img, 0) * np.array([2, 1, 0.5], np.float32))                               # clamped at 0, times L (:52-56); row 0 = top
    write_pfm(str(tmp_path / "npot.pfm"), np.ones((3, 8, 3), np.float32))
    e = P.HostScene.parse_string(base % ('"string mapname" ["%s"]' % (tmp_path / "npot.pfm"))).desc.envmaps[0]
    assert (e.width, e.height) == (8, 3)       # handed over as read: the MIPMap resamples to 8 x 4 (mipmap.rs:316-404)
    for bad in ('"string mapname" ["%s"]' % (tmp_path / "missing.pfm"), '"string mapname" ["x.exr"]'):
        with pytest.raises(P.PbrtError):
            P.HostScene.parse_string(base % bad)
    with pytest.raises(P.PbrtError):
        P.HostScene.parse_string(base % "", spectrum="sampled")


def test_front_end_aov_integrator(P):
    base = 'Camera "perspective"\nIntegrator "aov" %s\nWorldBegin\nWorldEnd\n'
    d = P.HostScene.parse_string(base % "").desc.integrator
    assert (d.type, d.aov_target, d.aov_scale) == (P.capi.PT_INTEGRATOR_AOV, P.capi.PT_AOV_UV, 1.0)       # aov.rs:197-198 defaults
    d = P.HostScene.parse_string(base % '"string target" ["shading.n"] "float scale" [0.5]').desc.integrator
    assert (d.aov_target, d.aov_scale) == (P.capi.PT_AOV_NS, 0.5)
    for bad in ("dpdx", "nonsense"):
        with pytest.raises(P.PbrtError):
            P.HostScene.parse_string(base % ('"string target" ["%s"]' % bad))
    assert P.HostScene.parse_string('Camera "perspective"\nWorldBegin\nWorldEnd\n').desc.integrator.type == P.capi.PT_INTEGRATOR_PATH


def test_front_end_spot_light(P):
    # lights/spot.rs:136-166: light_to_world = CTM * translate(from) * inverse(dir_to_z); cosines with the host libm
    txt = ('Camera "perspective"\nWorldBegin\nTranslate 1 2 3\nLightSource "spot" "rgb I" [10 20 30] "rgb scale" [2 2 2] "point from" [0 5 0] "point to" [0 0 0] '
           '"float coneangle" [40] "float conedeltaangle" [10]\nWorldEnd\n')
    hs = P.HostScene.parse_string(txt)
    d = hs.desc
    assert d.n_lights == 1 and d.n_spots == 1
    l, sp = d.lights[0], d.spots[0]
    assert l.type == P.capi.PT_LIGHT_SPOT and l.tri == 0 and np.allclose(list(l.v), [1, 7, 3])
    assert np.allclose(np.ctypeslib.as_array(d.spectra, (d.n_spectra * 3,))[3 * l.spectrum: 3 * l.spectrum + 3], [20, 40, 60])
    assert sp.cos_total_width == np.float32(np.cos(np.float32(40.0) * np.float32(np.pi / 180))) or abs(sp.cos_total_width - np.cos(np.radians(40.0))) < 1e-7
    assert abs(sp.cos_falloff_start - np.cos(np.radians(30.0))) < 1e-7
    w2l = np.array(list(sp.w2l)).reshape(3, 3)
    assert np.allclose(w2l @ np.array([0, -1, 0.0]), [0, 0, 1], atol=1e-6)       # the axis from -> to is +z of the light
    assert np.allclose(w2l @ w2l.T, np.eye(3), atol=1e-6)


def test_front_end_instances(P):
    hs = P.HostScene.parse_string(P.scenes.instanced_blobs(n_instances=5, n_sub=1, res=(32, 24), spp=1, n_lights=2))
    d = hs.desc
    assert d.n_objects == 5 and d.n_instances == 5 and d.n_lights == 4
    prims = np.ctypeslib.as_array(d.prims, (d.obj_first[d.n_objects],))
    top = prims[d.obj_first[0]:d.obj_first[1]]
    assert (top < 0).sum() == 5 and (top >= 0).sum() == 2 + 4


def test_front_end_errors(P):
    for bad in ['Shape "sphere"', 'WorldBegin\nLightSource "laser"', 'Sampler "maxmindist"\nWorldBegin\nWorldEnd', 'Bogus 1 2 3']:
        with pytest.raises(P.PbrtError):
            P.HostScene.parse_string('Camera "perspective"\n' + bad + "\n")


def _compare_bvh(P, hs, objects):
    import oracle_api
    orc = oracle_api.OracleScene(hs)
    hb = P.HostBvh(hs)
    for k in objects:
        n_nodes_o, ordered_o, depth_o = orc.bvh_info(k)
        info = hb.info(k)
        assert np.array_equal(info["ordered"], ordered_o), "object %d: primitive order differs" % k
        assert info["n_nodes"] == n_nodes_o, (k, info["n_nodes"], n_nodes_o)
        step = max(1, n_nodes_o // 3000)
        for node in list(range(0, n_nodes_o, step)) + [n_nodes_o - 1] if n_nodes_o else []:
            bo, co, ao, lo = orc.bvh_node(k, node)
            bp, cp, ap, lp = hb.node(k, node)
            assert lo == lp and np.array_equal(co, cp), (k, node)
            if not lo:
                assert np.array_equal(bo.view(np.uint32), bp.view(np.uint32)) and np.array_equal(ao, ap), (k, node)
    wb = orc.world_bound()
    assert np.array_equal(wb.view(np.uint32), hb.info(0)["bounds"].view(np.uint32))
    orc.close()
    hb.close()


def test_host_bvh_equals_oracle_cornell(P):
    _compare_bvh(P, P.HostScene.parse_string(P.scenes.cornell_box()), [0])


def test_host_bvh_equals_oracle_grid(P):
    _compare_bvh(P, P.HostScene.parse_string(P.scenes.displaced_grid(n=96, res=32, spp=1)), [0])


@pytest.mark.parametrize("split", ["middle", "equal"])
def test_host_bvh_equals_oracle_other_splits(P, split):
    txt = P.scenes.displaced_grid(n=40, res=32, spp=1).replace('Accelerator "bvh"', 'Accelerator "bvh" "string splitmethod" ["%s"] "integer maxnodeprims" [3]' % split)
    _compare_bvh(P, P.HostScene.parse_string(txt), [0])


@pytest.mark.parametrize("maxprims", [4, 1])
def test_host_bvh_equals_oracle_hlbvh(P, maxprims):
    """splitmethod "hlbvh" (build/hlbvh.rs): Morton treelets + SAH over the treelets; host build == oracle restatement, and the
    tree finds what a brute-force search finds."""
    import oracle_api
    hl = 'Accelerator "bvh" "string splitmethod" ["hlbvh"] "integer maxnodeprims" [%d]' % maxprims
    for txt, objects in ((P.scenes.displaced_grid(n=96, res=32, spp=1), [0]), (P.scenes.cornell_box(), [0]),
                         (P.scenes.instanced_blobs(n_instances=40, n_sub=2, res=(32, 24), spp=1, n_lights=4), [0, 1, 2])):
        hs = P.HostScene.parse_string(txt.replace('Accelerator "bvh"', hl))
        assert hs.desc.accel.split_method == P.capi.PT_SPLIT_HLBVH
        _compare_bvh(P, hs, objects)
        sah = oracle_api.OracleScene(P.HostScene.parse_string(txt))
        orc = oracle_api.OracleScene(hs)
        assert not np.array_equal(orc.bvh_info(0)[1], sah.bvh_info(0)[1])          # a different tree than the SAH build's
        wb = orc.world_bound()
        rng = np.random.RandomState(3)
        o = rng.uniform(wb[:3] - 1, wb[3:] + 1, (3000, 3)).astype(np.float32)
        t = rng.uniform(wb[:3], wb[3:], (3000, 3)).astype(np.float32)
        rays = np.concatenate([o, np.full((3000, 1), np.inf, np.float32), (t - o), np.zeros((3000, 1), np.float32)], 1).astype(np.float32)
        pa, ta = orc.intersect(rays)[:2]
        pb, tb = orc.intersect_exhaustive(rays)[:2]
        assert np.array_equal(pa, pb) and np.array_equal(ta.view(np.uint32), tb.view(np.uint32))
        orc.close(); sah.close()


def test_host_bvh_equals_oracle_instances(P):
    hs = P.HostScene.parse_string(P.scenes.instanced_blobs(n_instances=40, n_sub=2, res=(32, 24), spp=1, n_lights=4))
    _compare_bvh(P, hs, [0, 1, 2, 3, 4])


def test_host_bvh_coincident_centroids_make_big_leaf(P):
    # many identical triangles: centroid bounds are degenerate -> one leaf with all primitives (node.rs:55-62)
    tri = "0 0 0 1 0 0 0 1 0 "
    n = 12
    txt = 'Camera "perspective"\nWorldBegin\nShape "trianglemesh" "integer indices" [%s] "point P" [%s]\nWorldEnd\n' % (
        " ".join(str(i) for i in range(3 * n)), tri * n)
    hs = P.HostScene.parse_string(txt)
    _compare_bvh(P, hs, [0])
    hb = P.HostBvh(hs)
    assert hb.info(0)["n_nodes"] == 1 and hb.node(0, 0)[3] == 1


@pytest.mark.parametrize("scene", ["cornell", "grid", "instances"])
def test_oracle_lbvh_layout_finds_the_same_hits(P, scene):
    """Accelerator layout "lbvh" (accel/lbvh.rs) vs the 4-wide layout in the oracle: the same binary tree walked two ways.
    Closest-hit t is identical (same triangle test, same shrinking t_max), ids differ at most on exact ties, any-hit agrees,
    and the binary walk visits more (smaller) nodes."""
    import oracle_api
    txt = {"cornell": lambda: P.scenes.cornell_box(res=64, spp=4), "grid": lambda: P.scenes.displaced_grid(n=64, res=64, spp=1),
           "instances": lambda: P.scenes.instanced_blobs(n_instances=20, n_sub=1, res=(48, 32), spp=1, n_lights=4)}[scene]()
    hq = P.HostScene.parse_string(txt)
    hl = P.HostScene.parse_string(txt.replace('Accelerator "bvh"', 'Accelerator "bvh" "string layout" "lbvh"'))
    assert (hq.desc.accel.type, hl.desc.accel.type) == (P.capi.PT_ACCEL_QBVH, P.capi.PT_ACCEL_LBVH)
    oq, ol = oracle_api.OracleScene(hq), oracle_api.OracleScene(hl)
    sb = hq.desc.film.sample_bounds
    rng = np.random.RandomState(11)
    n = 4000
    pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
    rays, _ = oq.camera_rays(pix, np.zeros(n, np.int32))
    wb = oq.world_bound()
    o = rng.uniform(wb[:3], wb[3:], (n, 3))
    d = rng.normal(size=(n, 3))
    d[:30] = np.eye(3)[rng.randint(0, 3, 30)]                      # axis-aligned: inf / NaN slab products
    rnd = np.concatenate([o, np.full((n, 1), np.inf), d, np.zeros((n, 1))], 1).astype(np.float32)
    for r in (rays, rnd):
        pq, tq, bq, nq, trq = oq.intersect(r)
        pl, tl, bl, nl, trl = ol.intersect(r)
        assert np.array_equal(tq.view(np.uint32), tl.view(np.uint32))
        assert (pq != pl).mean() < 2e-3
        assert nl > nq
    r = rnd.copy()
    r[:, 3] = rng.uniform(0.01, 30.0, n)
    assert np.array_equal(oq.intersect_p(r), ol.intersect_p(r))
    oq.close(); ol.close()


def test_device_sincos_equals_libm(tmp_path):
    """The device's sin/cos (dmath.cuh pt_sinf/pt_cosf, f64 polynomial) must equal the host libm bit for bit: the
    reference's f32::sin/cos are libm calls, and ulp differences get amplified by specular chains."""
    exe = str(tmp_path / "check_sincos")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tools", "check_sincos.cpp"), "-o", exe])
    out = subprocess.check_output([exe, "5000000"]).decode().split()
    assert out[0] == "0" and out[1] == "0", out


def test_push_position_formula_matches_order_table():
    """k_trace stores hit child k at stack position popc(mask & before_k) (dtrav2.cuh, trav_node_step) instead of walking
    the reference's ORDER_TABLE nibble list (qbvh_x86.rs:187-204).  Check the closed form against all 128 entries."""
    import re
    src = open(os.path.join(ROOT, "pbrt-r3_b200", "csrc", "cuda", "dtrav.cuh")).read()
    body = src[src.index("kOrderTable[128] = {"):]
    body = body[:body.index("};")]
    table = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", body)]
    assert len(table) == 128
    for mask in range(16):
        for s in range(8):
            order, lst = table[mask * 8 + s], []
            while (order & 4) == 0:
                lst.append(order & 3)
                order >>= 4
            bl = 0 if (s & 4) else 0xC
            br = 0x3 if (s & 4) else 0
            before = [bl | (0 if (s & 2) else 2), bl | (1 if (s & 2) else 0), br | (0 if (s & 1) else 8), br | (4 if (s & 1) else 0)]
            mine = sorted([k for k in range(4) if mask & (1 << k)], key=lambda k: bin(mask & before[k]).count("1"))
            assert mine == lst, (mask, s, mine, lst)
            assert sorted(bin(mask & before[k]).count("1") for k in mine) == list(range(len(lst)))
