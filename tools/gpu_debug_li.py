"""Bring-up helper: which samples of pt_li differ from the oracle, grouped by the material of the primary hit."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbrt_r3_b200 as P
import oracle_api
txt = P.scenes.instanced_blobs(n_instances=60, n_sub=2, res=(96, 64), spp=4, n_lights=8)
for depth in (1, 2, 8):
    t2 = txt.replace('"integer maxdepth" [8]', '"integer maxdepth" [%d]' % depth)
    hs = P.HostScene.parse_string(t2)
    sc = P.Scene(hs); orc = oracle_api.OracleScene(hs)
    sb = hs.desc.film.sample_bounds
    rng = np.random.RandomState(4)
    n = 20000
    pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
    sn = rng.randint(0, 4, n).astype(np.int32)
    la = sc.li(pix, sn); lb, tri, t, _ = orc.li(pix, sn)
    close = (np.abs(la - lb) <= 1e-3 * np.maximum(np.abs(lb), 1e-2)).all(1)
    mats = np.ctypeslib.as_array(hs.desc.tri_material, (hs.desc.n_tris,))
    mtype = np.array([hs.desc.materials[i].type for i in range(hs.desc.n_materials)])
    first_mat = np.where(tri >= 0, mtype[mats[np.maximum(tri, 0)]], -1)
    print("maxdepth", depth, "close", close.mean())
    for m in np.unique(first_mat):
        sel = first_mat == m
        print("   first-hit material type", m, "n", sel.sum(), "bad", (~close & sel).sum())
    bad = np.where(~close)[0][:6]
    for i in bad: print("   bad", pix[i], sn[i], "tri", tri[i], "mat", first_mat[i], la[i], lb[i])
