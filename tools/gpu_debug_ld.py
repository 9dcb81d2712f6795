import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbrt_r3_b200 as P
import oracle_api
txt = P.scenes.instanced_blobs(n_instances=60, n_sub=2, res=(96, 64), spp=4, n_lights=8).replace('"integer maxdepth" [8]', '"integer maxdepth" [1]')
hs = P.HostScene.parse_string(txt)
sc = P.Scene(hs); orc = oracle_api.OracleScene(hs)
wb = orc.world_bound(); print("world bound", wb)
pts = np.random.RandomState(2).uniform(wb[:3], wb[3:], (4000, 3)).astype(np.float32)
a = sc.light_distribution(pts); b = orc.light_distribution(pts)
rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
print("light dist: max rel", rel.max(), "rows differing >1e-5:", (rel.max(1) > 1e-5).sum(), "of", len(pts))
bad = np.where(rel.max(1) > 1e-5)[0][:5]
for i in bad: print(pts[i], a[i], b[i])
pix = np.array([[89, 51], [93, 51], [68, 40], [66, 40], [67, 40], [69, 40], [50, 50]], np.int32); sn = np.array([0, 2, 2, 1, 2, 0, 0], np.int32)
rays, _ = orc.camera_rays(pix, sn)
p, t, bary = sc.intersect(rays)
hitp = rays[:, 0:3] + rays[:, 4:7] * t[:, None]
print("hit prims", p, "t", t)
a = sc.light_distribution(hitp); b = orc.light_distribution(hitp)
for i in range(len(pix)): print(pix[i], "gpu", a[i][:6], "cpu", b[i][:6])
print("li gpu", sc.li(pix, sn)); print("li cpu", orc.li(pix, sn)[0])
