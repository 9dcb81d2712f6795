"""ctypes mirror of include/pbrt_r3_b200.h (the C ABI of the B200 PathIntegrator core).

The structures below must stay byte-compatible with the header; tests/test_abi.py checks the sizes
against a compiled probe.  Nothing here computes: it only declares types and loads the library.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PBRT_R3_B200_LIB") or os.path.join(HERE, "libpbrt_r3_b200.so")  # env override: kernel A/B experiments
DATA_DIR = os.path.join(HERE, "data")

PT_ABI_VERSION = 3
PT_MAT_NONE, PT_MAT_MATTE, PT_MAT_PLASTIC, PT_MAT_GLASS, PT_MAT_METAL, PT_MAT_MIRROR = -1, 0, 1, 2, 3, 4
PT_MAT_UBER, PT_MAT_SUBSTRATE, PT_MAT_TRANSLUCENT, PT_MAT_MIX = 5, 6, 7, 8
PT_LIGHT_AREA, PT_LIGHT_POINT, PT_LIGHT_DISTANT, PT_LIGHT_SPOT, PT_LIGHT_INFINITE, PT_LIGHT_PROJECTION, PT_LIGHT_GONIOMETRIC = 0, 1, 2, 3, 4, 5, 6
PT_SAMPLER_HALTON, PT_SAMPLER_SOBOL, PT_SAMPLER_RANDOM, PT_SAMPLER_STRATIFIED, PT_SAMPLER_ZEROTWO = 0, 1, 2, 3, 4
PT_ACCEL_QBVH, PT_ACCEL_LBVH = 0, 1
PT_SPLIT_SAH, PT_SPLIT_HLBVH, PT_SPLIT_MIDDLE, PT_SPLIT_EQUAL = 0, 1, 2, 3
PT_CAMERA_PERSPECTIVE, PT_CAMERA_ORTHOGRAPHIC, PT_CAMERA_ENVIRONMENT = 0, 1, 2
PT_INTEGRATOR_PATH, PT_INTEGRATOR_AOV, PT_INTEGRATOR_AO, PT_INTEGRATOR_DIRECT, PT_INTEGRATOR_WHITTED = 0, 1, 2, 3, 4
PT_DIRECT_ALL, PT_DIRECT_ONE = 0, 1
PT_AOV_DISTANCE, PT_AOV_DEPTH, PT_AOV_N, PT_AOV_NS, PT_AOV_UV, PT_AOV_DPDUS = 0, 1, 2, 3, 4, 5
PT_SPECTRUM_RGB, PT_SPECTRUM_SAMPLED = 0, 1
PT_TRI_REVERSE_ORIENTATION, PT_TRI_SWAPS_HANDEDNESS, PT_TRI_TWO_SIDED, PT_TRI_HAS_N, PT_TRI_HAS_UV, PT_TRI_HAS_S = 1, 2, 4, 8, 16, 32

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)


class pt_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("kd", C.c_int32), ("ks", C.c_int32), ("kr", C.c_int32), ("sigma", C.c_float),
                ("uroughness", C.c_float), ("vroughness", C.c_float), ("eta", C.c_float), ("remap_roughness", C.c_int32),
                ("kt", C.c_int32), ("opacity", C.c_int32)]


class pt_light(C.Structure):
    _fields_ = [("type", C.c_int32), ("spectrum", C.c_int32), ("tri", C.c_int32), ("two_sided", C.c_int32),
                ("v", C.c_float * 3), ("n_samples", C.c_int32)]


class pt_envmap(C.Structure):
    _fields_ = [("l2w", C.c_float * 9), ("w2l", C.c_float * 9), ("width", C.c_int32), ("height", C.c_int32), ("texels", C.POINTER(C.c_float)),
                ("xf", C.c_float * 16), ("screen_bounds", C.c_float * 4), ("hitcher", C.c_float), ("cos_total_width", C.c_float)]


class pt_spot(C.Structure):
    _fields_ = [("w2l", C.c_float * 9), ("cos_total_width", C.c_float), ("cos_falloff_start", C.c_float), ("pad", C.c_float)]


class pt_instance(C.Structure):
    _fields_ = [("object", C.c_int32), ("m", C.c_float * 16), ("minv", C.c_float * 16)]


class pt_camera(C.Structure):
    _fields_ = [("raster_to_camera", C.c_float * 16), ("camera_to_world", C.c_float * 16), ("lens_radius", C.c_float),
                ("focal_distance", C.c_float), ("shutter_open", C.c_float), ("shutter_close", C.c_float), ("type", C.c_int32), ("pad", C.c_int32)]


class pt_film(C.Structure):
    _fields_ = [("full_resolution", C.c_int32 * 2), ("cropped_pixel_bounds", C.c_int32 * 4), ("sample_bounds", C.c_int32 * 4),
                ("filter_radius", C.c_float * 2), ("filter_table", C.c_float * 256), ("scale", C.c_float),
                ("max_sample_luminance", C.c_float)]


class pt_sampler(C.Structure):
    _fields_ = [("type", C.c_int32), ("samples_per_pixel", C.c_int32), ("sample_at_pixel_center", C.c_int32), ("strata", C.c_int32),
                ("jitter", C.c_int32), ("n_sampled_dimensions", C.c_int32)]


class pt_integrator(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("rr_threshold", C.c_float), ("light_strategy", C.c_int32), ("type", C.c_int32),
                ("aov_target", C.c_int32), ("aov_scale", C.c_float), ("ao_samples", C.c_int32), ("ao_cos_sample", C.c_int32), ("direct_strategy", C.c_int32)]


class pt_accel(C.Structure):
    _fields_ = [("type", C.c_int32), ("split_method", C.c_int32), ("max_prims_in_node", C.c_int32), ("pad", C.c_int32)]


class pt_scene_desc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("spectrum_kind", C.c_uint32), ("n_channels", C.c_uint32), ("n_spectra", C.c_uint32),
                ("spectra", f32p), ("xyz_weights", C.c_float * 180), ("xyz_scale", C.c_float),
                ("n_tris", C.c_uint32), ("tri_p", f32p), ("tri_n", f32p), ("tri_uv", f32p), ("tri_s", f32p), ("tri_flags", u32p),
                ("tri_material", i32p), ("tri_light", i32p),
                ("n_objects", C.c_uint32), ("obj_first", u32p), ("prims", i32p), ("n_instances", C.c_uint32),
                ("instances", C.POINTER(pt_instance)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(pt_material)),
                ("n_lights", C.c_uint32), ("n_emitters_only", C.c_uint32), ("lights", C.POINTER(pt_light)), ("n_spots", C.c_uint32), ("spots", C.POINTER(pt_spot)),
                ("n_envmaps", C.c_uint32), ("envmaps", C.POINTER(pt_envmap)),
                ("camera", pt_camera), ("film", pt_film), ("sampler", pt_sampler), ("integrator", pt_integrator), ("accel", pt_accel)]


class pt_render_opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("shard_index", C.c_int32), ("shard_count", C.c_int32), ("tile_size", C.c_int32),
                ("spp_begin", C.c_int32), ("spp_end", C.c_int32), ("wave_paths", C.c_int32), ("collect_counters", C.c_int32), ("wave_segments", C.c_int32)]


class pt_stats(C.Structure):
    _fields_ = [("camera_rays", C.c_uint64), ("regular_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("nodes_visited", C.c_uint64),
                ("tris_tested", C.c_uint64), ("path_vertices", C.c_uint64), ("zero_radiance_samples", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_ms", C.c_double), ("extend_ms", C.c_double), ("shadow_ms", C.c_double),
                ("shade_ms", C.c_double), ("extend_launches", C.c_uint64), ("extend_rays", C.c_uint64), ("extend_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class pt_options(C.Structure):
    _fields_ = [("quick_render", C.c_int32), ("quick_render_full_resolution", C.c_int32)]


# every symbol include/pbrt_r3_b200.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = [
    "pt_scene_create", "pt_scene_destroy", "pt_scene_reupload", "pt_last_error", "pt_abi_version", "pt_render", "pt_render_partial", "pt_film_resolve",
    "pt_film_clear", "pt_intersect", "pt_intersect_p", "pt_intersect_device", "pt_sampler_values", "pt_camera_rays", "pt_li",
    "pt_bvh_info", "pt_bvh_node", "pt_light_distribution", "pt_parse_file", "pt_parse_string", "pt_host_scene_desc",
    "pt_host_scene_filename", "pt_host_scene_warning_count", "pt_host_scene_warning", "pt_host_scene_destroy", "pt_host_bvh_build", "pt_host_bvh_info", "pt_host_bvh_node",
    "pt_host_bvh_destroy", "pt_set_options", "pt_get_options",
]

_lib = None


def load(path=None):
    """Load libpbrt_r3_b200.so (built by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first "
                           "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.pt_last_error.restype = C.c_char_p
    lib.pt_abi_version.restype = C.c_int
    lib.pt_parse_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    lib.pt_parse_string.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    lib.pt_host_scene_desc.argtypes = [vp]
    lib.pt_host_scene_desc.restype = C.POINTER(pt_scene_desc)
    lib.pt_host_scene_filename.argtypes = [vp]
    lib.pt_host_scene_filename.restype = C.c_char_p
    lib.pt_host_scene_warning_count.argtypes = [vp]
    lib.pt_host_scene_warning_count.restype = C.c_uint32
    lib.pt_host_scene_warning.argtypes = [vp, C.c_uint32]
    lib.pt_host_scene_warning.restype = C.c_char_p
    lib.pt_host_scene_destroy.argtypes = [vp]
    lib.pt_host_scene_destroy.restype = None
    lib.pt_host_bvh_build.argtypes = [C.POINTER(pt_scene_desc), C.POINTER(vp)]
    lib.pt_host_bvh_info.argtypes = [vp, C.c_uint32, u32p, u32p, i32p, f32p, u32p, u32p]
    lib.pt_host_bvh_node.argtypes = [vp, C.c_uint32, C.c_uint32, f32p, C.POINTER(C.c_int64), i32p, i32p]
    lib.pt_host_bvh_destroy.argtypes = [vp]
    lib.pt_host_bvh_destroy.restype = None
    lib.pt_set_options.argtypes = [C.POINTER(pt_options)]
    lib.pt_set_options.restype = None
    lib.pt_get_options.argtypes = [C.POINTER(pt_options)]
    lib.pt_get_options.restype = None
    if hasattr(lib, "pt_scene_create"):
        lib.pt_scene_create.argtypes = [C.POINTER(pt_scene_desc), C.c_int, C.POINTER(vp)]
        lib.pt_scene_destroy.argtypes = [vp]
        lib.pt_scene_destroy.restype = None
        lib.pt_scene_reupload.argtypes = [vp, C.POINTER(C.c_uint64)]
        lib.pt_render.argtypes = [vp, C.POINTER(pt_render_opts), f32p, C.POINTER(pt_stats)]
        lib.pt_render_partial.argtypes = [vp, C.POINTER(pt_render_opts), C.POINTER(vp), C.POINTER(pt_stats)]
        lib.pt_film_resolve.argtypes = [vp, vp, f32p]
        lib.pt_film_clear.argtypes = [vp]
        lib.pt_intersect.argtypes = [vp, f32p, C.c_uint32, i32p, f32p, f32p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.pt_intersect_p.argtypes = [vp, f32p, C.c_uint32, C.POINTER(C.c_uint8)]
        lib.pt_intersect_device.argtypes = [vp, vp, vp, C.c_uint32, vp, C.c_int, f32p]
        lib.pt_sampler_values.argtypes = [vp, i32p, i32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p]
        lib.pt_camera_rays.argtypes = [vp, i32p, i32p, C.c_uint32, f32p, f32p]
        lib.pt_li.argtypes = [vp, i32p, i32p, C.c_uint32, f32p]
        lib.pt_bvh_info.argtypes = [vp, C.c_uint32, u32p, u32p, i32p]
        lib.pt_bvh_node.argtypes = [vp, C.c_uint32, C.c_uint32, f32p, C.POINTER(C.c_int64), i32p, i32p]
        lib.pt_light_distribution.argtypes = [vp, f32p, C.c_uint32, f32p]
    if path is None:
        _lib = lib
    return lib
