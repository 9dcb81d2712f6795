"""Per-sample radiance of the GPU path vs the oracle on the 'mixed' Cornell scene: how many samples differ and by how much."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api
import pbrt_r3_b200 as P
hs = P.HostScene.parse_string(P.scenes.cornell_box(res=96, spp=16, materials="mixed", extra_lights=True, maxdepth=8))
sc = P.Scene(hs, 0); orc = oracle_api.OracleScene(hs)
sb = hs.desc.film.sample_bounds
rng = np.random.RandomState(4)
n = 20000
pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
sn = rng.randint(0, 16, n).astype(np.int32)
la = sc.li(pix, sn); lb = orc.li(pix, sn)[0]
rel = np.abs(la - lb).max(1) / np.maximum(np.abs(lb).max(1), 1e-2)
for tol in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1):
    print("rel diff > %g: %d of %d" % (tol, (rel > tol).sum(), n))
print("sum", la.sum(), lb.sum(), abs(la.sum() - lb.sum()) / lb.sum())
bad = np.argsort(-rel)[:8]
for i in bad: print(pix[i], sn[i], la[i], lb[i])
rgb, _ = sc.render(); ref, _ = orc.render()
print("image rel rmse", float(np.sqrt(((rgb - ref) ** 2).mean()) / ref.mean()))
