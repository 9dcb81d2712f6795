#!/usr/bin/env python
"""bench.py — PathIntegrator bounce loop throughput on B200 (BASELINE.json metric: Mrays/s & Msamples/s).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU restatement of pbrt-r3's path on the host cores)

Workload = BASELINE.json configs[1]: synthetic 1M-triangle displaced mesh, 1024x1024, 64 spp Halton, plastic/metal/
glass mix, maxdepth 8, spatial light sampling.  One step = one full render pass of that scene (per GPU: 64 spp of every
pixel of the sample bounds; with N GPUs rank r renders sample numbers [64r, 64(r+1)) and the film partials are summed
onto rank 0 with one NCCL reduce, i.e. weak scaling in spp).

`value`  : rays (Scene::intersect + Scene::intersect_p queries, the reference's own counters, scene.rs:11-12) per
           second with the scene resident in HBM and the film left on the device (pt_render_partial [+ NCCL reduce]).
`e2e`    : the same metric through the reference-facing C-ABI call with HOST buffers: per step the flattened scene is
           re-sent from pinned host memory (pt_scene_reupload), rendered, and the RGB image read back (pt_render).
`roofline`: k_trace (closest-hit + any-hit traversal), algorithmic bytes 128*N_node + 48*N_tri + 48 per ray
           (N_node/N_tri from an untimed counting pass of the same step) / its CUDA-event time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = {"grid_n": 708, "res": 1024, "spp": 64, "maxdepth": 8}


def workload_name(c):
    return "synthetic %d-triangle displaced-mesh scene %dx%d, %d spp halton, plastic/metal/glass, maxdepth %d" % (
        2 * (c["grid_n"] - 1) ** 2 + 4, c["res"], c["res"], c["spp"], c["maxdepth"])


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6544.3, "fallback (B200_PROFILING.md measured copy bandwidth)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.sm, self.mx, self.reasons = gpu, False, [], 0, set()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.hw_thermal_slowdown," \
            "clocks_throttle_reasons.sw_thermal_slowdown,clocks_throttle_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                              timeout=5).decode().strip().split(",")
                self.sm.append(float(out[0]))
                self.mx = max(self.mx, float(out[1]))
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        self.stop_flag = True
        self.join(timeout=3)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx or None, "reasons": sorted(self.reasons)}


def oracle_rate(hs, threads, s0, s1):
    """Times the CPU oracle on sample numbers [s0, s1) of the workload; returns (Mrays/s, Msamples/s, seconds, stats)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api
    orc = oracle_api.OracleScene(hs)
    t = time.time()
    _, st = orc.render(threads=threads, s0=s0, s1=s1)
    dt = time.time() - t
    orc.close()
    rays = st["regular_rays"] + st["shadow_rays"]
    return rays / dt / 1e6, st["camera_rays"] / dt / 1e6, dt, st


def run_reference(args, cfg, out=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build()
    import pbrt_r3_b200 as P
    hs = P.HostScene.parse_string(P.scenes.displaced_grid(n=cfg["grid_n"], res=cfg["res"], spp=cfg["spp"], maxdepth=cfg["maxdepth"]))
    cores = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api
    orc = oracle_api.OracleScene(hs)     # BVH build outside the timed region, like the reference's make_scene
    sample = "1 of %d spp of every pixel per step (sample numbers cycle 0,1,2,..), all %d host threads" % (cfg["spp"], cores)
    times, rays, samples = [], 0, 0
    for it in range(args.warmup + args.steps):
        s = it % cfg["spp"]
        t = time.time()
        _, st = orc.render(threads=cores, s0=s, s1=s + 1)
        dt = time.time() - t
        if it >= args.warmup:
            times.append(dt)
            rays += st["regular_rays"] + st["shadow_rays"]
            samples += st["camera_rays"]
    tot = sum(times)
    v = rays / tot / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / max(len(times), 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "msamples_per_s": samples / tot / 1e6,
            "config": {"workload": workload_name(cfg), "impl": "C++ restatement of pbrt-r3's CPU path (oracle/; the Rust reference cannot be built: no rustc)"},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out.write(json.dumps(line) + "\n")
    out.flush()
    return 0


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner on the first collective,
    make / nvcc notes from build()): route fd 1 to stderr for the whole run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _quiet_stdout()
    try:
        return _main(out)
    finally:
        out.flush()


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--grid", type=int, default=CONFIG["grid_n"])
    ap.add_argument("--res", type=int, default=CONFIG["res"])
    ap.add_argument("--spp", type=int, default=CONFIG["spp"])
    ap.add_argument("--maxdepth", type=int, default=CONFIG["maxdepth"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wave-paths", type=int, default=0, help="path slots per wave (0 = library default)")
    ap.add_argument("--wave-segments", type=int, default=0, help="waves in flight (0 = library default)")
    args = ap.parse_args()
    cfg = {"grid_n": args.grid, "res": args.res, "spp": args.spp, "maxdepth": args.maxdepth}
    if args.impl == "reference":
        return run_reference(args, cfg, out)

    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        import __graft_entry__ as g
        g.build()
    if world > 1:
        dist.barrier()
    import pbrt_r3_b200 as P
    from pbrt_r3_b200 import dist as D

    spp_total = cfg["spp"] * world          # weak scaling in spp: every GPU renders cfg["spp"] samples of every pixel
    hs = P.HostScene.parse_string(P.scenes.displaced_grid(n=cfg["grid_n"], res=cfg["res"], spp=spp_total, maxdepth=cfg["maxdepth"]))
    sc = P.Scene(hs, local)                 # host SAH build + upload: outside the timed region (the reference's make_scene)
    s0, s1 = D.spp_shard(spp_total, rank, world)
    dev = "cuda:%d" % local
    H, W, _ = hs.image_shape

    wopts = {"wave_paths": args.wave_paths, "wave_segments": args.wave_segments}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local)

    # untimed counting pass of one step: N_node / N_tri of the reference traversal (deterministic)
    ptr, npx, cst = sc.render_partial(spp_begin=s0, spp_end=s1, collect_counters=1, **wopts)
    rays_per_step = cst["regular_rays"] + cst["shadow_rays"]
    samples_per_step = cst["camera_rays"]
    alg_bytes = cst["extend_bytes"]

    def step_device():
        """One step; returns (stats, device seconds): CUDA events on the library's launch stream bracket the render
        (pt_stats.device_ms), torch events on torch's stream bracket the NCCL reduce."""
        ptr, npx, st = sc.render_partial(spp_begin=s0, spp_end=s1, **wopts)
        dev_ms = st["device_ms"]
        if world > 1:
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            D.reduce_film(D.film_tensor(ptr, npx, dev), dst=0)
            r1.record()
            r1.synchronize()
            dev_ms += r0.elapsed_time(r1)
        return st, dev_ms / 1e3

    def step_e2e():
        nb = sc.reupload()
        if world == 1:
            rgb, st = sc.render(spp_begin=s0, spp_end=s1, **wopts)
            return st, nb, rgb.nbytes
        ptr, npx, st = sc.render_partial(spp_begin=s0, spp_end=s1, **wopts)
        film = D.film_tensor(ptr, npx, dev)
        D.reduce_film(film, dst=0)
        d2h = 0
        if rank == 0:
            rgb = sc.film_resolve(film.data_ptr())
            d2h = rgb.nbytes
        return st, nb, d2h

    for _ in range(max(args.warmup, 3)):
        step_device()
    clk = ClockSampler(local)
    clk.start()
    barrier()
    t0 = time.time()
    trace_ms, shade_ms, launches, trace_launches, dev_s = 0.0, 0.0, 0, 0, 0.0
    for _ in range(args.steps):
        st, ds = step_device()
        dev_s += ds
        trace_ms += st["extend_ms"]
        shade_ms += st["shade_ms"]
        launches += st["kernel_launches"]
        trace_launches += st["extend_launches"]
    barrier()
    wall = time.time() - t0
    # device time of the K steps (CUDA events on the launching streams), max over ranks; wall time is reported beside it
    tt = torch.tensor([dev_s, wall], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_steps, wall = float(tt[0].item()), float(tt[1].item())
    clocks = clk.result()

    # end-to-end: host buffers in, host image out, every step
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.time()
    h2d = d2h = 0
    for _ in range(args.steps):
        _, a, b = step_e2e()
        h2d, d2h = a, b
    barrier()
    te = torch.tensor([time.time() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())

    tot_rays = torch.tensor([rays_per_step, samples_per_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot_rays, op=dist.ReduceOp.SUM)
    rays_all, samples_all = float(tot_rays[0].item()), float(tot_rays[1].item())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = alg_bytes * args.steps / (trace_ms / 1e3) / 1e9 if trace_ms > 0 else 0.0
        npix = samples_per_step / max(cfg["spp"], 1)
        wave_m = npix * max(1, min(cfg["spp"], int((args.wave_paths or 6400000) // max(npix, 1)))) / 1e6   # paths per wave, millions
        nseg = args.wave_segments or 6
        value = rays_all * args.steps / t_steps / 1e6
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * t_steps / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "msamples_per_s": samples_all * args.steps / t_steps / 1e6,
            "config": {"workload": workload_name(cfg), "parallelism": "spp-sharded x%d, scene replicated, NCCL film reduce" % world,
                       "spp_per_gpu": cfg["spp"], "waves_in_flight": nseg,
                       "l2_policy": "inputs larger than L2 (126 MB): every bounce streams the path state of the waves in flight (%.1f GB: %d waves x %.2f M paths x ~270 B) "
                                    "and walks %.0f MB of BVH nodes + triangle records" % (
                                        nseg * wave_m * 270 / 1e3, nseg, wave_m, (2 * (cfg["grid_n"] - 1) ** 2) * (48 + 48 + 32) / 1e6)},
            "e2e": {"value": rays_all * args.steps / t_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "k_trace", "peak_source": peak_src, "bytes_per_ray": alg_bytes / max(rays_per_step, 1),
                         "trace_ms_per_step": trace_ms / args.steps, "shade_ms_per_step": shade_ms / args.steps,
                         "trace_launches_per_step": trace_launches / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps},
        }
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                line["roofline"]["traffic"] = json.load(open(tp)).get("k_trace_dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ncpu = max(1, min(cfg["spp"], 4))     # about 20-30 s of CPU work on a 16-core host
            mr, ms, dt, _ = oracle_rate(hs, cores, 0, ncpu)
            line["cpu_baseline"] = {"value": mr, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                    "sample": "sample numbers 0..%d of every pixel (%d/%d of the step), %d threads, %.1f s wall; C++ restatement of pbrt-r3's CPU path" % (ncpu - 1, ncpu, cfg["spp"], cores, dt),
                                    "msamples_per_s": ms}
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
