"""First-contact GPU script: runs each stage of the CUDA path on the Cornell box against the oracle and prints
where they diverge.  Used during bring-up; the real checks live in tests/ (-m gpu)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbrt_r3_b200 as P
import oracle_api

def main():
    res = int(os.environ.get("RES", "128")); spp = int(os.environ.get("SPP", "16"))
    kind = os.environ.get("SCENE", "cornell")
    if kind == "cornell": txt = P.scenes.cornell_box(res=res, spp=spp)
    elif kind == "mixed": txt = P.scenes.cornell_box(res=res, spp=spp, materials="mixed", extra_lights=True, maxdepth=8)
    else: txt = P.scenes.displaced_grid(n=int(os.environ.get("GRID", "128")), res=res, spp=spp)
    hs = P.HostScene.parse_string(txt)
    t = time.time(); sc = P.Scene(hs); print("scene create s", time.time() - t)
    orc = oracle_api.OracleScene(hs)
    sb = hs.desc.film.sample_bounds
    rng = np.random.RandomState(1)
    n = 4096
    pix = np.stack([rng.randint(sb[0], sb[2], n), rng.randint(sb[1], sb[3], n)], 1).astype(np.int32)
    sn = rng.randint(0, spp, n).astype(np.int32)
    a = sc.sampler_values(pix, sn, 0, 24); b = orc.sampler_values(pix, sn, 0, 24)
    print("sampler bit-exact:", np.array_equal(a.view(np.uint32), b.view(np.uint32)), "max abs diff", np.abs(a - b).max())
    ra, pfa = sc.camera_rays(pix, sn); rb, pfb = orc.camera_rays(pix, sn)
    print("camera rays bit-exact:", np.array_equal(ra.view(np.uint32), rb.view(np.uint32)), "max abs diff", np.abs(ra - rb)[np.isfinite(rb)].max())
    pa, ta, ba, nodes_a, tris_a = sc.intersect(rb, counters=True); pb, tb, bb, nodes_b, tris_b = orc.intersect(rb)
    print("primary prim ids equal:", np.array_equal(pa, pb), "mismatch", int((pa != pb).sum()), "t bit-exact", np.array_equal(ta.view(np.uint32), tb.view(np.uint32)),
          "nodes", nodes_a, nodes_b, "tris", tris_a, tris_b)
    # secondary rays: random origins inside the scene
    wb = orc.world_bound()
    o = rng.uniform(wb[:3], wb[3:], (n, 3)); d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, np.full((n, 1), np.inf), d, np.zeros((n, 1))], 1).astype(np.float32)
    pa, ta, ba, nodes_a, tris_a = sc.intersect(rays, counters=True); pb, tb, bb, nodes_b, tris_b = orc.intersect(rays)
    print("random rays prim ids equal:", np.array_equal(pa, pb), "mismatch", int((pa != pb).sum()), "t bit-exact", np.array_equal(ta.view(np.uint32), tb.view(np.uint32)),
          "nodes", nodes_a, nodes_b, "tris", tris_a, tris_b)
    rays[:, 3] = rng.uniform(1, 300, n)
    print("any-hit equal:", np.array_equal(sc.intersect_p(rays), orc.intersect_p(rays)))
    pts = rng.uniform(wb[:3], wb[3:], (64, 3)).astype(np.float32)
    fa = sc.light_distribution(pts); fb = orc.light_distribution(pts)
    print("light distribution max rel diff", np.abs(fa - fb).max() / max(np.abs(fb).max(), 1e-20), "bit-exact", np.array_equal(fa.view(np.uint32), fb.view(np.uint32)))
    la = sc.li(pix, sn); lb, tri_b, t_b, _ = orc.li(pix, sn)
    d = np.abs(la - lb); rel = d.sum() / max(lb.sum(), 1e-20)
    print("li: exact fraction", float((la.view(np.uint32) == lb.view(np.uint32)).all(1).mean()), "close(1e-4) fraction", float((d <= 1e-4 * np.maximum(np.abs(lb), 1e-3)).all(1).mean()),
          "sum rel diff", rel, "mean", la.mean(), lb.mean())
    bad = np.where(~(d <= 1e-3 * np.maximum(np.abs(lb), 1e-3)).all(1))[0][:8]
    for i in bad: print("  bad", pix[i], sn[i], la[i], lb[i])
    t = time.time(); rgb, st = sc.render(collect_counters=1); tg = time.time() - t
    t = time.time(); ref, so = orc.render(); tc = time.time() - t
    diff = rgb - ref
    rmse = np.sqrt((diff ** 2).mean()) / max(ref.mean(), 1e-20)
    print("render gpu s", tg, "cpu s", tc, "rel RMSE", rmse, "max abs", np.abs(diff).max(), "mean", rgb.mean(), ref.mean())
    print("gpu stats", {k: v for k, v in st.items()})
    print("cpu stats", so)
    np.save(os.path.join(ROOT, "gpurun_out", "gpu_%s.npy" % kind), rgb); np.save(os.path.join(ROOT, "gpurun_out", "cpu_%s.npy" % kind), ref)

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
