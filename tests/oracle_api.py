"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (the CPU restatement of pbrt-r3's path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
DATA = os.path.join(ROOT, "pbrt-r3_b200", "data")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        build()
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_scene_create.restype = vp
    lib.orc_scene_create.argtypes = [vp]
    lib.orc_scene_destroy.argtypes = [vp]
    lib.orc_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.orc_render_shard.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.orc_intersect.argtypes = [vp, vp, C.c_uint32, vp, vp, vp, vp, vp]
    lib.orc_intersect_p.argtypes = [vp, vp, C.c_uint32, vp]
    lib.orc_intersect_exhaustive.argtypes = [vp, vp, C.c_uint32, C.c_int, vp, vp, vp]
    lib.orc_sampler_values.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    lib.orc_camera_rays.argtypes = [vp, vp, vp, C.c_uint32, vp, vp]
    lib.orc_li.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, vp, vp]
    lib.orc_bvh_info.argtypes = [vp, C.c_uint32, vp, vp, vp, vp]
    lib.orc_bvh_node.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    lib.orc_world_bound.argtypes = [vp, vp]
    lib.orc_light_distribution.argtypes = [vp, vp, C.c_uint32, vp]
    lib.orc_radical_inverse.restype = C.c_float
    lib.orc_radical_inverse.argtypes = [C.c_uint32, C.c_uint64]
    lib.orc_scrambled_radical_inverse.restype = C.c_float
    lib.orc_scrambled_radical_inverse.argtypes = [C.c_uint32, C.c_uint64, vp]
    lib.orc_scrambled_radical_inverse_default.restype = C.c_float
    lib.orc_scrambled_radical_inverse_default.argtypes = [C.c_uint32, C.c_uint64]
    lib.orc_inverse_radical_inverse.restype = C.c_uint64
    lib.orc_inverse_radical_inverse.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    lib.orc_prime.restype = C.c_uint64
    lib.orc_prime.argtypes = [C.c_uint32]
    lib.orc_prime_sum.restype = C.c_uint64
    lib.orc_prime_sum.argtypes = [C.c_uint32]
    lib.orc_perm.restype = C.c_uint32
    lib.orc_perm.argtypes = [C.c_uint64]
    lib.orc_sobol_sample.restype = C.c_float
    lib.orc_sobol_sample.argtypes = [C.c_int64, C.c_uint32, C.c_uint32]
    lib.orc_sobol_interval_to_index.restype = C.c_uint64
    lib.orc_sobol_interval_to_index.argtypes = [C.c_uint32, C.c_uint64, C.c_int32, C.c_int32]
    lib.orc_rng_sequence.argtypes = [C.c_uint64, C.c_int, C.c_uint32, vp, vp]
    lib.orc_distribution1d.argtypes = [vp, C.c_uint32, vp, vp]
    lib.orc_distribution1d_sample_discrete.restype = C.c_uint32
    lib.orc_distribution1d_sample_discrete.argtypes = [vp, C.c_uint32, C.c_float, vp]
    lib.orc_distribution1d_sample_continuous.restype = C.c_float
    lib.orc_distribution1d_sample_continuous.argtypes = [vp, C.c_uint32, C.c_float, vp, vp]
    lib.orc_tri_intersect_raw.argtypes = [vp, C.c_uint32, vp, vp, vp]
    lib.orc_spawn_ray_from_hit.argtypes = [vp, C.c_uint32, vp, vp, vp, vp]
    lib.orc_bxdf_sample.argtypes = [C.c_int, C.c_float, C.c_int, vp, vp, vp, vp]
    lib.orc_bxdf_pdf.restype = C.c_float
    lib.orc_bxdf_pdf.argtypes = [C.c_int, C.c_float, C.c_int, vp, vp]
    if lib.orc_set_data_dir(DATA.encode()) != 0:
        raise RuntimeError(lib.orc_last_error().decode())
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data if a is not None else None


class OracleScene:
    """The oracle's view of a parsed scene (same pt_scene_desc the CUDA path gets)."""

    def __init__(self, host_scene):
        self.lib = load()
        self.host = host_scene  # keeps the desc alive
        self.h = self.lib.orc_scene_create(C.cast(host_scene.desc_ptr, C.c_void_p))
        if not self.h:
            raise RuntimeError(self.lib.orc_last_error().decode())
        self.nc = host_scene.desc.n_channels

    def close(self):
        if self.h:
            self.lib.orc_scene_destroy(self.h)
            self.h = None

    def render(self, threads=None, s0=0, s1=0, want_xyzw=False, shard_index=0, shard_count=1):
        H, W, _ = self.host.image_shape
        rgb = np.zeros((H, W, 3), np.float32)
        xyzw = np.zeros((H, W, 4), np.float32) if want_xyzw else None
        cnt = np.zeros(6, np.uint64)
        self.lib.orc_render_shard(self.h, threads or os.cpu_count() or 1, s0, s1, shard_index, shard_count, _ptr(rgb), _ptr(xyzw), _ptr(cnt))
        stats = dict(zip(["camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices"], [int(v) for v in cnt]))
        return (rgb, stats, xyzw) if want_xyzw else (rgb, stats)

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        bary = np.zeros((n, 3), np.float32)
        nodes, tris = C.c_uint64(0), C.c_uint64(0)
        self.lib.orc_intersect(self.h, _ptr(rays), n, _ptr(prim), _ptr(t), _ptr(bary), C.addressof(nodes), C.addressof(tris))
        return prim, t, bary, nodes.value, tris.value

    def intersect_p(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], np.uint8)
        self.lib.orc_intersect_p(self.h, _ptr(rays), rays.shape[0], _ptr(occ))
        return occ

    def intersect_exhaustive(self, rays, threads=None):
        """ExhaustiveAccel: every primitive in scene order, no BVH; returns (prim, t, instance)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        prim, t, inst = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.orc_intersect_exhaustive(self.h, _ptr(rays), n, threads or os.cpu_count() or 1, _ptr(prim), _ptr(t), _ptr(inst))
        return prim, t, inst

    def sampler_values(self, pixel_xy, sample_num, dim0, ndims):
        pixel_xy = np.ascontiguousarray(pixel_xy, np.int32).reshape(-1, 2)
        sample_num = np.ascontiguousarray(sample_num, np.int32)
        out = np.zeros((pixel_xy.shape[0], ndims), np.float32)
        rc = self.lib.orc_sampler_values(self.h, _ptr(pixel_xy), _ptr(sample_num), pixel_xy.shape[0], dim0, ndims, _ptr(out))
        assert rc == 0
        return out

    def camera_rays(self, pixel_xy, sample_num):
        pixel_xy = np.ascontiguousarray(pixel_xy, np.int32).reshape(-1, 2)
        sample_num = np.ascontiguousarray(sample_num, np.int32)
        n = pixel_xy.shape[0]
        rays = np.zeros((n, 8), np.float32)
        pf = np.zeros((n, 2), np.float32)
        self.lib.orc_camera_rays(self.h, _ptr(pixel_xy), _ptr(sample_num), n, _ptr(rays), _ptr(pf))
        return rays, pf

    def li(self, pixel_xy, sample_num):
        pixel_xy = np.ascontiguousarray(pixel_xy, np.int32).reshape(-1, 2)
        sample_num = np.ascontiguousarray(sample_num, np.int32)
        n = pixel_xy.shape[0]
        out = np.zeros((n, self.nc), np.float32)
        tri = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        cnt = np.zeros(6, np.uint64)
        self.lib.orc_li(self.h, _ptr(pixel_xy), _ptr(sample_num), n, _ptr(out), _ptr(tri), _ptr(t), _ptr(cnt))
        return out, tri, t, cnt

    def bvh_info(self, obj=0):
        nn, npr, md = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        self.lib.orc_bvh_info(self.h, obj, C.addressof(nn), C.addressof(npr), None, C.addressof(md))
        ordered = np.zeros(npr.value, np.int32)
        self.lib.orc_bvh_info(self.h, obj, None, None, _ptr(ordered), None)
        return nn.value, ordered, md.value

    def bvh_node(self, obj, node):
        boxes = np.zeros((2, 3, 4), np.float32)
        ch = np.zeros(4, np.int64)
        ax = np.zeros(3, np.int32)
        leaf = C.c_int32(0)
        rc = self.lib.orc_bvh_node(self.h, obj, node, _ptr(boxes), _ptr(ch), _ptr(ax), C.addressof(leaf))
        assert rc == 0
        return boxes, ch, ax, leaf.value

    def world_bound(self):
        b = np.zeros(6, np.float32)
        self.lib.orc_world_bound(self.h, _ptr(b))
        return b

    def light_distribution(self, points):
        points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        nl = self.host.desc.n_lights
        out = np.zeros((points.shape[0], nl), np.float32)
        self.lib.orc_light_distribution(self.h, _ptr(points), points.shape[0], _ptr(out))
        return out
