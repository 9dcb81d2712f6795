"""Throughput of the BASELINE.json configs that are not the bench line (configs[2] instanced 10 M triangles,
configs[3] 60-bin spectral Cornell, configs[4] = the instanced scene at 3840x2160: pixel-chunked waves), at reduced spp so that a run takes seconds.  Prints one JSON line per config:
Mrays/s, Msamples/s, k_trace algorithmic-bytes fraction of the measured HBM peak (same definitions as bench.py).
Usage: python tools/bench_configs.py [--spp3 N] [--spp4 N] [--steps K]"""
import argparse
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbrt_r3_b200 as P

ap = argparse.ArgumentParser()
ap.add_argument("--spp3", type=int, default=8)
ap.add_argument("--spp4", type=int, default=16)
ap.add_argument("--spp5", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--only", default="")
args = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6544.3)


def run(name, text, spp, spectrum="rgb"):
    t0 = time.time()
    hs = P.HostScene.parse_string(text, spectrum=spectrum)
    t1 = time.time()
    sc = P.Scene(hs, 0)
    t2 = time.time()
    _, _, cst = sc.render_partial(spp_begin=0, spp_end=spp, collect_counters=1)
    for _ in range(3):
        sc.render_partial(spp_begin=0, spp_end=spp)
    dev = tr = sh = 0.0
    for _ in range(args.steps):
        _, _, st = sc.render_partial(spp_begin=0, spp_end=spp)
        dev += st["device_ms"]; tr += st["extend_ms"]; sh += st["shade_ms"]
    rays = cst["regular_rays"] + cst["shadow_rays"]
    line = {"config": name, "spp": spp, "n_tris": int(hs.desc.n_tris), "n_instances": int(hs.desc.n_instances), "n_channels": int(hs.desc.n_channels),
            "mrays_per_s": rays * args.steps / dev / 1e3, "msamples_per_s": cst["camera_rays"] * args.steps / dev / 1e3,
            "ms_per_step": dev / args.steps, "trace_ms": tr / args.steps, "shade_ms": sh / args.steps,
            "bytes_per_ray": cst["extend_bytes"] / max(cst["extend_rays"], 1), "nodes_per_ray": cst["nodes_visited"] / max(cst["extend_rays"], 1),
            "tris_per_ray": cst["tris_tested"] / max(cst["extend_rays"], 1),
            "trace_frac_of_hbm_peak": cst["extend_bytes"] / (tr / args.steps / 1e3) / 1e9 / peak,
            "parse_s": t1 - t0, "scene_create_s": t2 - t1}
    print(json.dumps(line), flush=True)
    sc.close()


if args.only in ("", "3"):
    run("configs[2]: 1000 instances of a 20480-triangle blob (20.5 M instanced triangles), 256 area-light quads, 1920x1080, maxdepth 8",
        P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(1920, 1080), spp=args.spp3, maxdepth=8, n_lights=256), args.spp3)
if args.only in ("", "4"):
    run("configs[3]: 60-bin SampledSpectrum Cornell box 512x512", P.scenes.cornell_box(res=512, spp=args.spp4), args.spp4, spectrum="sampled")
if args.only in ("", "5"):
    run("configs[4]: the instanced scene at 3840x2160 (8.3 M sample pixels: waves are pixel chunks), one GPU's share at reduced spp",
        P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(3840, 2160), spp=args.spp5, maxdepth=8, n_lights=256), args.spp5)
