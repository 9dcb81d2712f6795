"""pbrt-r3_b200 — B200-native PathIntegrator core behind the reference's plugin names.

Python host mirror used by the tests and bench.py (the reference's host is Rust; rustc is absent here).  Everything
that computes goes through the C ABI of include/pbrt_r3_b200.h (ctypes, host buffers in / host buffers out); torch is
only used by pbrt_r3_b200.dist for the NCCL film reduce.  There is NO CPU fallback: `Scene.create` raises when the CUDA
library is missing or no device is visible.

Reference call chain mirrored here:
    pbrt_parse_file -> SceneContext -> make_scene / make_integrator -> integrator.render(scene) -> film.write_image
    src/bin/pbrt.rs:219-356, src/core/integrator/sampler.rs:259-325
"""
import ctypes as C

import numpy as np

from . import capi, scenes  # noqa: F401
from .capi import pt_render_opts, pt_scene_desc, pt_stats


class PbrtError(RuntimeError):
    """Mirror of PbrtError (reference src/core/error/error.rs:4-30)."""


def _check(lib, rc, what):
    if rc != 0:
        raise PbrtError("%s: %s" % (what, lib.pt_last_error().decode()))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class HostScene:
    """Result of parsing a .pbrt scene (pbrt_parse_file + SceneContext): owns the flattened pt_scene_desc."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib

    @classmethod
    def parse_string(cls, text, pixelsamples=-1, spectrum="rgb"):
        lib = capi.load()
        h = C.c_void_p()
        kind = capi.PT_SPECTRUM_SAMPLED if spectrum in ("sampled", 60, "60") else capi.PT_SPECTRUM_RGB
        _check(lib, lib.pt_parse_string(text.encode(), int(pixelsamples), kind, C.byref(h)), "pt_parse_string")
        return cls(h, lib)

    @classmethod
    def parse_file(cls, path, pixelsamples=-1, spectrum="rgb"):
        lib = capi.load()
        h = C.c_void_p()
        kind = capi.PT_SPECTRUM_SAMPLED if spectrum in ("sampled", 60, "60") else capi.PT_SPECTRUM_RGB
        _check(lib, lib.pt_parse_file(str(path).encode(), int(pixelsamples), kind, C.byref(h)), "pt_parse_file")
        return cls(h, lib)

    @property
    def desc_ptr(self):
        return self._lib.pt_host_scene_desc(self._h)

    @property
    def desc(self):
        return self.desc_ptr.contents

    @property
    def filename(self):
        return self._lib.pt_host_scene_filename(self._h).decode()

    @property
    def warnings(self):
        """what the reference would log with warn! while parsing"""
        return [self._lib.pt_host_scene_warning(self._h, i).decode() for i in range(self._lib.pt_host_scene_warning_count(self._h))]

    @property
    def image_shape(self):
        b = self.desc.film.cropped_pixel_bounds
        return (b[3] - b[1], b[2] - b[0], 3)

    def close(self):
        if self._h:
            self._lib.pt_host_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostBvh:
    """Host-only accelerator build (same code pt_scene_create runs), for build-parity checks without a GPU."""

    def __init__(self, host_scene):
        self._lib = capi.load()
        self._h = C.c_void_p()
        self.host = host_scene
        _check(self._lib, self._lib.pt_host_bvh_build(host_scene.desc_ptr, C.byref(self._h)), "pt_host_bvh_build")

    def info(self, obj=0):
        nn, npr, gn, qd = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        b = np.zeros(6, np.float32)
        _check(self._lib, self._lib.pt_host_bvh_info(self._h, obj, C.byref(nn), C.byref(npr), None, _p(b, C.c_float), C.byref(gn), C.byref(qd)), "pt_host_bvh_info")
        ordered = np.zeros(npr.value, np.int32)
        _check(self._lib, self._lib.pt_host_bvh_info(self._h, obj, None, None, _p(ordered, C.c_int32), None, None, None), "pt_host_bvh_info")
        return dict(n_nodes=nn.value, ordered=ordered, bounds=b, gpu_nodes=gn.value, qdepth=qd.value)

    def node(self, obj, node):
        boxes = np.zeros((2, 3, 4), np.float32)
        ch = np.zeros(4, np.int64)
        ax = np.zeros(3, np.int32)
        leaf = C.c_int32(0)
        _check(self._lib, self._lib.pt_host_bvh_node(self._h, obj, node, _p(boxes, C.c_float), _p(ch, C.c_int64), _p(ax, C.c_int32), C.byref(leaf)), "pt_host_bvh_node")
        return boxes, ch, ax, leaf.value

    def close(self):
        if self._h:
            self._lib.pt_host_bvh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Device-resident scene + integrator (SceneContext::make_scene + make_integrator; QBVHAccel::new)."""

    def __init__(self, host_scene, device=0):
        self.host = host_scene
        self._lib = capi.load()
        if not hasattr(self._lib, "pt_scene_create"):
            raise PbrtError("libpbrt_r3_b200.so was built without the CUDA kernels; there is no CPU fallback")
        self._h = C.c_void_p()
        _check(self._lib, self._lib.pt_scene_create(host_scene.desc_ptr, int(device), C.byref(self._h)), "pt_scene_create")
        self.n_channels = host_scene.desc.n_channels

    def close(self):
        if self._h:
            self._lib.pt_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reupload(self):
        """Re-send the flattened scene from pinned host memory to the device; returns the bytes copied."""
        b = C.c_uint64(0)
        _check(self._lib, self._lib.pt_scene_reupload(self._h, C.byref(b)), "pt_scene_reupload")
        return b.value

    @staticmethod
    def _opts(**kw):
        o = pt_render_opts()
        for k, v in kw.items():
            setattr(o, k, int(v))
        return o

    def _out(self, out):
        """The host RGB buffer pt_render / pt_film_resolve write: a fresh array, or the caller's own [H,W,3] float32 C-contiguous
        one (the reference's Film keeps one RGB buffer for its life; a page-locked buffer makes the read-back a plain DMA)."""
        H, W, _ = self.host.image_shape
        if out is None:
            return np.zeros((H, W, 3), np.float32)
        if out.dtype != np.float32 or out.shape != (H, W, 3) or not out.flags["C_CONTIGUOUS"]:
            raise PbrtError("out must be a C-contiguous float32 array of shape %r" % ((H, W, 3),))
        return out

    def render(self, out=None, **opts):
        """Integrator::render: returns (rgb[H,W,3] float32 host array, stats dict)."""
        rgb = self._out(out)
        st = pt_stats()
        o = self._opts(**opts)
        _check(self._lib, self._lib.pt_render(self._h, C.byref(o), _p(rgb, C.c_float), C.byref(st)), "pt_render")
        return rgb, st.as_dict()

    def render_partial(self, **opts):
        """Leaves the partial film {X,Y,Z,sum_w} on the device; returns (device pointer, n_pixels, stats)."""
        ptr = C.c_void_p()
        st = pt_stats()
        o = self._opts(**opts)
        _check(self._lib, self._lib.pt_render_partial(self._h, C.byref(o), C.byref(ptr), C.byref(st)), "pt_render_partial")
        H, W, _ = self.host.image_shape
        return ptr.value, H * W, st.as_dict()

    def film_resolve(self, dev_ptr=None, out=None):
        rgb = self._out(out)
        _check(self._lib, self._lib.pt_film_resolve(self._h, C.c_void_p(dev_ptr or 0), _p(rgb, C.c_float)), "pt_film_resolve")
        return rgb

    def intersect(self, rays, counters=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        bary = np.zeros((n, 3), np.float32)
        nodes, tris = C.c_uint64(0), C.c_uint64(0)
        _check(self._lib, self._lib.pt_intersect(self._h, _p(rays, C.c_float), n, _p(prim, C.c_int32), _p(t, C.c_float), _p(bary, C.c_float),
                                                 C.byref(nodes) if counters else None, C.byref(tris) if counters else None), "pt_intersect")
        if counters:
            return prim, t, bary, nodes.value, tris.value
        return prim, t, bary

    def intersect_p(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], np.uint8)
        _check(self._lib, self._lib.pt_intersect_p(self._h, _p(rays, C.c_float), rays.shape[0], _p(occ, C.c_uint8)), "pt_intersect_p")
        return occ

    def sampler_values(self, pixel_xy, sample_num, dim0, ndims):
        pixel_xy = np.ascontiguousarray(pixel_xy, np.int32).reshape(-1, 2)
        sample_num = np.ascontiguousarray(sample_num, np.int32)
        out = np.zeros((pixel_xy.shape[0], ndims), np.float32)
        _check(self._lib, self._lib.pt_sampler_values(self._h, _p(pixel_xy, C.c_int32), _p(sample_num, C.c_int32), pixel_xy.shape[0], dim0, ndims,
                                                      _p(out, C.c_float)), "pt_sampler_values")
        return out

    def camera_rays(self, pixel_xy, sample_num):
        pixel_xy = np.ascontiguousarray(pixel_xy, np.int32).reshape(-1, 2)
        sample_num = np.ascontiguousarray(sample_num, np.int32)
        n = pixel_xy.shape[0]
        rays = np.zeros((n, 8), np.float32)
        pf = np.zeros((n, 2), np.float32)
        _check(self._lib, self._lib.pt_camera_rays(self._h, _p(pixel_xy, C.c_int32), _p(sample_num, C.c_int32), n, _p(rays, C.c_float), _p(pf, C.c_float)),
               "pt_camera_rays")
        return rays, pf

    def li(self, pixel_xy, sample_num):
        pixel_xy = np.ascontiguousarray(pixel_xy, np.int32).reshape(-1, 2)
        sample_num = np.ascontiguousarray(sample_num, np.int32)
        n = pixel_xy.shape[0]
        out = np.zeros((n, self.n_channels), np.float32)
        _check(self._lib, self._lib.pt_li(self._h, _p(pixel_xy, C.c_int32), _p(sample_num, C.c_int32), n, _p(out, C.c_float)), "pt_li")
        return out

    def bvh_info(self, obj=0):
        nn, npr = C.c_uint32(0), C.c_uint32(0)
        _check(self._lib, self._lib.pt_bvh_info(self._h, obj, C.byref(nn), C.byref(npr), None), "pt_bvh_info")
        ordered = np.zeros(npr.value, np.int32)
        _check(self._lib, self._lib.pt_bvh_info(self._h, obj, None, None, _p(ordered, C.c_int32)), "pt_bvh_info")
        return nn.value, ordered

    def bvh_node(self, obj, node):
        boxes = np.zeros((2, 3, 4), np.float32)
        ch = np.zeros(4, np.int64)
        ax = np.zeros(3, np.int32)
        leaf = C.c_int32(0)
        _check(self._lib, self._lib.pt_bvh_node(self._h, obj, node, _p(boxes, C.c_float), _p(ch, C.c_int64), _p(ax, C.c_int32), C.byref(leaf)), "pt_bvh_node")
        return boxes, ch, ax, leaf.value

    def light_distribution(self, points):
        points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        nl = self.host.desc.n_lights
        out = np.zeros((points.shape[0], nl), np.float32)
        _check(self._lib, self._lib.pt_light_distribution(self._h, _p(points, C.c_float), points.shape[0], _p(out, C.c_float)), "pt_light_distribution")
        return out


def render_pbrt(text=None, path=None, device=0, pixelsamples=-1, spectrum="rgb", **opts):
    """`pbrt-r3 scene.pbrt` equivalent: parse, build, render; returns (rgb, stats, HostScene)."""
    hs = HostScene.parse_string(text, pixelsamples, spectrum) if text is not None else HostScene.parse_file(path, pixelsamples, spectrum)
    sc = Scene(hs, device)
    try:
        rgb, st = sc.render(**opts)
    finally:
        sc.close()
    return rgb, st, hs
