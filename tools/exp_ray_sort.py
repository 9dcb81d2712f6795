"""Experiment: does ordering the rays of a launch by (origin cell, direction octant) speed up k_trace on the instanced scene?
Secondary rays of BASELINE configs[2]: origins = primary hit points of ~2 M camera samples (tile order, like the wavefront's
queues), directions = random, flipped to the camera's side of the surface.  pt_intersect_device (the same persistent-warp
kernel the bounce loop launches) on: the queue order, a random permutation, and the rays sorted by several keys."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pbrt_r3_b200 as P
from pbrt_r3_b200 import capi
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
if cfg == 3:
    text = P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(1920, 1080), spp=4, maxdepth=8, n_lights=256)
    W, H = 1920, 1080
else:
    text = P.scenes.displaced_grid(n=708, res=1024, spp=4, maxdepth=8)
    W, H = 1024, 1024
hs = P.HostScene.parse_string(text)
sc = P.Scene(hs, 0)
lib = capi.load()
# camera samples in 16x16 tile order (the order of the wavefront's pixel list)
T = 16
pix = np.array([[tx * T + x, ty * T + y] for ty in range(H // T) for tx in range(W // T) for y in range(T) for x in range(T)], np.int32)
sn = np.zeros(len(pix), np.int32)
rays, _ = sc.camera_rays(pix, sn)
prim, t, _ = sc.intersect(rays)
hit = prim >= 0
o = rays[hit, 0:3] + rays[hit, 4:7] * t[hit, None]
d0 = rays[hit, 4:7]
rng = np.random.RandomState(1)
d = rng.normal(size=o.shape).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
flip = (d * d0).sum(1) > 0
d[flip] *= -1
o = o + d * 1e-3 * np.abs(o).max()
n = len(o)
ro = np.concatenate([o, np.full((n, 1), np.inf, np.float32)], 1).astype(np.float32)
rd = np.concatenate([d, np.zeros((n, 1), np.float32)], 1).astype(np.float32)

def morton(q):
    q = q.astype(np.uint64)
    r = np.zeros(len(q), np.uint64)
    for b in range(10):
        for a in range(3):
            r |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return r
lo, hi = o.min(0), o.max(0)
cell = np.clip(((o - lo) / (hi - lo) * 1023), 0, 1023).astype(np.uint32)
m = morton(cell)
octant = ((d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] < 0).astype(np.uint64) << np.uint64(2)))
orders = {"queue order (tile / pixel)": np.arange(n), "random permutation": rng.permutation(n),
          "morton(origin)": np.argsort(m, kind="stable"), "octant, morton(origin)": np.argsort((octant << np.uint64(30)) | m, kind="stable"),
          "morton(origin) >> 9, octant": np.argsort(((m >> np.uint64(9)) << np.uint64(3)) | octant, kind="stable"),
          "morton(origin) >> 15, octant": np.argsort(((m >> np.uint64(15)) << np.uint64(3)) | octant, kind="stable")}
hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
print("cfg", cfg, "secondary rays", n)
for name, idx in orders.items():
    a = torch.from_numpy(np.ascontiguousarray(ro[idx])).cuda()
    b = torch.from_numpy(np.ascontiguousarray(rd[idx])).cuda()
    torch.cuda.synchronize()
    ms = C.c_float(0)
    for rep in (3, 10):
        rc = lib.pt_intersect_device(sc._h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), n, C.c_void_p(hits.data_ptr()), rep, C.byref(ms))
        assert rc == 0
    print("%-32s %.3f ms  %.0f Mrays/s" % (name, ms.value, n / ms.value / 1e3), flush=True)
