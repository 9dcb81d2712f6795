"""Throughput of the integrators beside "path" (ambientocclusion, directlighting, whitted) on the geometry of
BASELINE configs[1] (1 M-triangle displaced mesh, 1024^2) and configs[2] (instanced, 1920x1080), at a few spp.  One JSON
line per case: Mrays/s on the GPU (pt_render_partial, film left on the device), k_trace's algorithmic-bytes fraction of
the HBM peak, and the oracle's rate on the host cores over a bounded sample (every R-th 16x16 tile, one sample number).
Usage: python tools/bench_integrators.py [--steps K] [--only NAME] [--no-cpu]"""
import argparse
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbrt_r3_b200 as P

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--only", default="")
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6544.3)
except Exception:
    peak = 6544.3


def with_integrator(text, spec):
    head, rest = text.split('Integrator "path"', 1)
    return head + spec + rest[rest.index("\n"):]


def run(name, text, spp):
    hs = P.HostScene.parse_string(text)
    sc = P.Scene(hs, 0)
    _, _, cst = sc.render_partial(spp_begin=0, spp_end=spp, collect_counters=1)
    for _ in range(3):
        sc.render_partial(spp_begin=0, spp_end=spp)
    dev = tr = sh = 0.0
    for _ in range(args.steps):
        _, _, st = sc.render_partial(spp_begin=0, spp_end=spp)
        dev += st["device_ms"]; tr += st["extend_ms"]; sh += st["shade_ms"]
    rays = cst["regular_rays"] + cst["shadow_rays"]
    line = {"case": name, "spp": spp, "n_tris": int(hs.desc.n_tris), "n_instances": int(hs.desc.n_instances),
            "mrays_per_s": round(rays * args.steps / dev / 1e3, 1), "msamples_per_s": round(cst["camera_rays"] * args.steps / dev / 1e3, 2),
            "rays_per_step": rays, "ms_per_step": round(dev / args.steps, 2), "trace_ms": round(tr / args.steps, 2), "stage_ms": round(sh / args.steps, 2),
            "kernel_launches": st["kernel_launches"],
            "trace_frac_of_hbm_peak": round(cst["extend_bytes"] / (tr / args.steps / 1e3) / 1e9 / peak, 3)}
    sc.close()
    if not args.no_cpu:
        import oracle_api
        sb = hs.desc.film.sample_bounds
        npix = (sb[2] - sb[0]) * (sb[3] - sb[1])
        R = max(1, int(round(npix * (rays / max(cst["camera_rays"], 1)) / 8.0e6)))   # about 8 M rays of CPU work
        orc = oracle_api.OracleScene(hs)
        t0 = time.time()
        _, so = orc.render(s0=0, s1=1, shard_index=0, shard_count=R)
        dt = time.time() - t0
        orc.close()
        line["cpu_port_mrays_per_s"] = round((so["regular_rays"] + so["shadow_rays"]) / dt / 1e6, 2)
        line["cpu_cores"] = os.cpu_count()
        line["cpu_sample"] = "sample number 0 of every %d-th 16x16 tile" % R
    print(json.dumps(line), flush=True)


CASES = {
    "direct_one_grid": lambda: (with_integrator(P.scenes.displaced_grid(n=708, res=1024, spp=16, maxdepth=8), 'Integrator "directlighting" "string strategy" ["one"] "integer maxdepth" [5]'), 16),
    "direct_all_grid": lambda: (with_integrator(P.scenes.displaced_grid(n=708, res=1024, spp=16, maxdepth=8), 'Integrator "directlighting" "integer maxdepth" [5]'), 16),
    "whitted_grid": lambda: (with_integrator(P.scenes.displaced_grid(n=708, res=1024, spp=16, maxdepth=8), 'Integrator "whitted" "integer maxdepth" [5]'), 16),
    "direct_one_instanced": lambda: (with_integrator(P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(1920, 1080), spp=4, maxdepth=8, n_lights=256),
                                                     'Integrator "directlighting" "string strategy" ["one"] "integer maxdepth" [5]'), 4),
    "ao_grid": lambda: (with_integrator(P.scenes.displaced_grid(n=708, res=1024, spp=4, maxdepth=8), 'Integrator "ambientocclusion" "integer nsamples" [64]'), 4),
    "ao_instanced": lambda: (with_integrator(P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(1920, 1080), spp=1, maxdepth=8, n_lights=256),
                                             'Integrator "ambientocclusion" "integer nsamples" [64]'), 1),
}
for k, f in CASES.items():
    if args.only and args.only not in k:
        continue
    text, spp = f()
    run(k, text, spp)
