#!/usr/bin/env python
"""bench.py — PathIntegrator bounce loop throughput on B200 (BASELINE.json metric: Mrays/s & Msamples/s at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W [--config C]   (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                         (the CPU restatement of pbrt-r3's path on the host cores)

Default workload = BASELINE.json configs[4], the configuration the metric's "1/2/4/8 B200" split is quoted on: the
synthetic instanced scene (1000 instances of a 20 480-triangle blob = 20.5 M instanced triangles, 256 emissive quads =
512 area lights, two-level 4-wide BVH, plastic / metal / glass / matte, maxdepth 8) at 3840x2160, 1024 spp Halton.
One step = sample numbers [0, S) of every pixel of that render (S = 8: 66.5 M camera samples, ~0.57 G rays) — a FIXED
job: with N GPUs rank r renders the sample numbers [S r / N, S (r + 1) / N) of every pixel (scene replicated) and the
film partials {X, Y, Z, sum w} are summed onto rank 0 with one NCCL reduce, i.e. STRONG scaling.  `--config 2` is the
round-1 line (1 M-triangle displaced mesh, 1024^2, 64 spp per step); `--config 1|3|4` the other BASELINE configs.

`value`  : rays (Scene::intersect + Scene::intersect_p queries, the reference's own counters, scene.rs:11-12) of the
           whole job per second, scene resident in HBM, film left on the device (pt_render_partial [+ NCCL reduce]).
`e2e`    : the same metric through the reference-facing C-ABI call with HOST buffers: per step the flattened scene is
           re-sent from pinned host memory (pt_scene_reupload), rendered, and the RGB image read back (pt_render).
`roofline`: the dominant kernel, k_trace: algorithmic bytes 128 N_node + 48 N_tri + 48 per ray (N_node /
           N_tri of the reference traversal, counted by an untimed pass of the same step) / its CUDA-event time.
`rooflines`: the same for k_shade (B_shade x path vertices / its time) and for the whole step (all bytes / step time).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# step_spp: sample numbers per step (the fixed job that --gpus N splits); spp: the config's own samples per pixel
CONFIGS = {
    1: {"name": "configs[0]", "scene": "cornell", "res": (256, 256), "spp": 16, "step_spp": 16, "maxdepth": 5, "spectrum": "rgb"},
    2: {"name": "configs[1]", "scene": "grid", "grid_n": 708, "res": (1024, 1024), "spp": 64, "step_spp": 64, "maxdepth": 8, "spectrum": "rgb"},
    3: {"name": "configs[2]", "scene": "instanced", "res": (1920, 1080), "spp": 256, "step_spp": 16, "maxdepth": 8, "spectrum": "rgb"},
    4: {"name": "configs[3]", "scene": "cornell", "res": (512, 512), "spp": 128, "step_spp": 32, "maxdepth": 5, "spectrum": "sampled"},
    5: {"name": "configs[4]", "scene": "instanced", "res": (3840, 2160), "spp": 1024, "step_spp": 8, "maxdepth": 8, "spectrum": "rgb"},
}
DEFAULT_CONFIG = 5


def scene_text(P, c):
    if c["scene"] == "cornell":
        return P.scenes.cornell_box(res=c["res"][0], spp=c["spp"], maxdepth=c["maxdepth"])
    if c["scene"] == "grid":
        return P.scenes.displaced_grid(n=c["grid_n"], res=c["res"][0], spp=c["spp"], maxdepth=c["maxdepth"])
    return P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=c["res"], spp=c["spp"], maxdepth=c["maxdepth"], n_lights=256)


def workload_name(c):
    what = {"cornell": "Cornell box (34 triangles, matte + DiffuseAreaLight)",
            "grid": "synthetic %d-triangle displaced-mesh scene, plastic/metal/glass" % (2 * (c.get("grid_n", 2) - 1) ** 2 + 4),
            "instanced": "synthetic instanced scene (1000 x 20480-triangle blob = 20.5 M instanced triangles, 512 area lights, two-level QBVH), plastic/metal/glass/matte"}[c["scene"]]
    return "%s: %s, %dx%d, %d spp halton%s, maxdepth %d; one step = sample numbers [0, %d) of every pixel (fixed job, split over the GPUs)" % (
        c["name"], what, c["res"][0], c["res"][1], c["spp"], ", 60-bin SampledSpectrum" if c["spectrum"] == "sampled" else "", c["maxdepth"], c["step_spp"])


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6544.3, "fallback (B200_PROFILING.md measured copy bandwidth)"


def shade_bytes_per_vertex(desc):
    """B_shade of SURVEY.md 8d: hit 16 + triangle 48 [+ 72 normals / uv] + material 64 (60-bin: + 3 x 240) + light CDF
    binary search 2 x 4 x ceil(log2(n_lights + 1)) + light record 64 + three rays written 96 + path state read + write
    2 x (24 + 2 x 4 x C)."""
    C = int(desc.n_channels)
    b = 16 + 48 + (72 if (bool(desc.tri_n) or bool(desc.tri_uv)) else 0) + 64 + (3 * 240 if C != 3 else 0)
    b += 2 * 4 * int(math.ceil(math.log2(int(desc.n_lights) + 1))) + 64 + 3 * 32 + 2 * (24 + 2 * 4 * C)
    return b


def sample_bytes_per_camera_sample(desc):
    """B_samp of SURVEY.md 8d: ray write 32 + half the path state + 16 F (F = pixels of the filter footprint)."""
    C = int(desc.n_channels)
    rx, ry = desc.film.filter_radius[0], desc.film.filter_radius[1]
    F = max(1, int(math.ceil(2 * rx)) * int(math.ceil(2 * ry)))
    return 32 + (24 + 2 * 4 * C) + 16 * F


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


def cpu_sample_plan(hs, c):
    """A bounded sample of the step for the CPU arm: ONE sample number of every R-th 16x16 render tile (interleaved, so the
    sample covers the whole image), R chosen so that it is about one million camera samples."""
    sb = hs.desc.film.sample_bounds
    npix = (sb[2] - sb[0]) * (sb[3] - sb[1])
    R = max(1, int(round(npix / 1.0e6)))
    return R, "sample number s (cycling 0, 1, 2, ..) of every %s16x16 tile of the %dx%d sample bounds (%.2f M of the step's %.1f M camera samples)" % (
        ("%d-th " % R) if R > 1 else "", sb[2] - sb[0], sb[3] - sb[1], npix / R / 1e6, npix * c["step_spp"] / 1e6)


def run_reference(args, c, out):
    """--impl reference: the CPU implementation of the path on the host cores.  The Rust reference cannot be built here (no
    rustc / cargo, crates not vendored), so this is its C++ restatement (oracle/, `kind: port`), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build()
    import pbrt_r3_b200 as P
    hs = P.HostScene.parse_string(scene_text(P, c), spectrum=c["spectrum"])
    cores = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api
    orc = oracle_api.OracleScene(hs)     # BVH build outside the timed region, like the reference's make_scene
    R, sample = cpu_sample_plan(hs, c)
    times, rays, samples = [], 0, 0
    for it in range(args.warmup + args.steps):
        s = it % c["step_spp"]
        t = time.time()
        _, st = orc.render(threads=cores, s0=s, s1=s + 1, shard_index=it % R, shard_count=R)
        dt = time.time() - t
        if it >= args.warmup:
            times.append(dt)
            rays += st["regular_rays"] + st["shadow_rays"]
            samples += st["camera_rays"]
    tot = sum(times)
    v = rays / tot / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / max(len(times), 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "msamples_per_s": samples / tot / 1e6,
            "config": {"workload": workload_name(c)},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": "per step: " + sample + ", %d host threads; C++ restatement of pbrt-r3's CPU path (oracle/): no rustc in the image" % cores},
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
    ap.add_argument("--config", type=int, default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--step-spp", type=int, default=0, help="sample numbers per step (0 = the config's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wave-paths", type=int, default=0, help="path slots per wave (0 = library default)")
    ap.add_argument("--wave-segments", type=int, default=0, help="waves in flight (0 = library default)")
    args = ap.parse_args()
    c = dict(CONFIGS[args.config])
    if args.step_spp > 0:
        c["step_spp"] = min(args.step_spp, c["spp"])
    if args.impl == "reference":
        return run_reference(args, c, out)

    import numpy as np  # noqa: F401
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

    t0 = time.time()
    hs = P.HostScene.parse_string(scene_text(P, c), spectrum=c["spectrum"])
    parse_s = time.time() - t0
    t0 = time.time()
    torch.zeros(1, device="cuda:%d" % local)   # CUDA context creation, reported on its own
    torch.cuda.synchronize(local)
    cuda_init_s = time.time() - t0
    t0 = time.time()
    sc = P.Scene(hs, local)                 # host SAH build + flatten + upload + light table: outside the timed region (the reference's make_scene)
    scene_create_s = time.time() - t0
    s0, s1 = D.spp_shard(c["step_spp"], rank, world)      # the fixed job's sample numbers, split over the ranks
    dev = "cuda:%d" % local

    wopts = {"wave_paths": args.wave_paths, "wave_segments": args.wave_segments}

    def render_partial(**kw):
        # more ranks than sample numbers: this rank's range [s0, s0) is empty (s0 = step_spp > 0) and pt_render_partial
        # leaves an all-zero partial film
        return sc.render_partial(spp_begin=s0, spp_end=s1, **wopts, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local)

    # untimed counting pass of one step: N_node / N_tri of the reference traversal and the path vertices (deterministic)
    ptr, npx, cst = render_partial(collect_counters=1)
    rays_per_step = cst["regular_rays"] + cst["shadow_rays"]
    samples_per_step = cst["camera_rays"]

    def step_device():
        """One step; returns (stats, device seconds): CUDA events on the library's launch stream bracket the render
        (pt_stats.device_ms), torch events on torch's stream bracket the NCCL reduce."""
        ptr, npx, st = render_partial()
        dev_ms = st["device_ms"]
        if world > 1:
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            D.reduce_film(D.film_tensor(ptr, npx, dev), dst=0)
            r1.record()
            r1.synchronize()
            dev_ms += r0.elapsed_time(r1)
        return st, dev_ms / 1e3

    # the host image every e2e step is read back into: allocated once, page-locked (like the pinned staging copy of the
    # scene that pt_scene_reupload sends); only rank 0 resolves the film
    H_img, W_img, _ = hs.image_shape
    rgb_host = torch.empty((H_img, W_img, 3), dtype=torch.float32, pin_memory=True).numpy() if rank == 0 else None

    def step_e2e():
        nb = sc.reupload()
        if world == 1:
            rgb, st = sc.render(out=rgb_host, spp_begin=s0, spp_end=s1, **wopts)
            return st, nb, rgb.nbytes
        ptr, npx, st = render_partial()
        film = D.film_tensor(ptr, npx, dev)
        D.reduce_film(film, dst=0)
        d2h = 0
        if rank == 0:
            rgb = sc.film_resolve(film.data_ptr(), out=rgb_host)
            d2h = rgb.nbytes
        return st, nb, d2h

    W = max(args.warmup, 3)
    for _ in range(W):
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

    # whole-job totals (all ranks) and the slowest rank's kernel times
    tot = torch.tensor([rays_per_step, samples_per_step, cst["extend_bytes"], cst["path_vertices"], launches], device=dev, dtype=torch.float64)
    mx = torch.tensor([trace_ms, shade_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    rays_all, samples_all, ext_bytes_all, vertices_all, launches_all = [float(v) for v in tot.tolist()]

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        K = args.steps
        # rank 0's own kernels against rank 0's own bytes (every rank runs the same kernels on 1/N of the job)
        b_shade, b_samp = shade_bytes_per_vertex(hs.desc), sample_bytes_per_camera_sample(hs.desc)
        tr_bytes, sh_bytes, sm_bytes = cst["extend_bytes"], b_shade * cst["path_vertices"], b_samp * samples_per_step
        tr_ach = tr_bytes * K / (trace_ms / 1e3) / 1e9 if trace_ms > 0 else 0.0
        sh_ach = sh_bytes * K / (shade_ms / 1e3) / 1e9 if shade_ms > 0 else 0.0
        step_bytes_all = ext_bytes_all + b_shade * vertices_all + b_samp * samples_all
        st_ach = step_bytes_all * K / t_steps / 1e9 / max(world, 1)        # per GPU
        value = rays_all * K / t_steps / 1e6
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp)).get("config%d" % args.config)
                if tj:
                    traffic, traffic_src = tj.get("k_trace_dram_bytes_per_launch"), tj.get("source")
            except Exception:
                pass
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_steps / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "msamples_per_s": samples_all * K / t_steps / 1e6,
            "config": {"workload": workload_name(c),
                       "parallelism": "sample numbers of the step split over %d GPU%s (rank r: [%d r / N, %d (r + 1) / N)), scene replicated, one NCCL film reduce" % (
                           world, "" if world == 1 else "s", c["step_spp"], c["step_spp"]),
                       "step_spp": c["step_spp"], "spp_per_gpu_per_step": s1 - s0,
                       "l2_policy": "inputs larger than L2 (126 MB): every bounce streams the path state of the waves in flight (%.1f M camera samples per step per GPU x ~270 B) "
                                    "and walks the BVH (%d triangles, %d instances)" % (samples_per_step / 1e6, int(hs.desc.n_tris), int(hs.desc.n_instances))},
            "e2e": {"value": rays_all * K / t_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "rays_per_step": int(rays_all), "camera_samples_per_step": int(samples_all), "path_vertices": int(vertices_all),
            "parse_s": parse_s, "scene_create_s": scene_create_s, "cuda_init_s": cuda_init_s,
            "roofline": {"bound": "hbm", "achieved": tr_ach, "peak": peak, "unit": "GB/s", "frac": tr_ach / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "kernel": "k_trace", "peak_source": peak_src, "bytes_per_ray": tr_bytes / max(cst["extend_rays"], 1),
                         "nodes_per_ray": cst["nodes_visited"] / max(cst["extend_rays"], 1), "tris_per_ray": cst["tris_tested"] / max(cst["extend_rays"], 1),
                         "trace_ms_per_step": trace_ms / K, "shade_ms_per_step": shade_ms / K,
                         "trace_launches_per_step": trace_launches / K, "wall_ms_per_step": 1e3 * wall / K},
            "rooflines": {
                "k_trace": {"achieved": tr_ach, "frac": tr_ach / peak, "ms_per_step": trace_ms / K, "bytes_per_step": tr_bytes,
                            "definition": "sum over rays of 128 N_node + 48 N_tri + 48 (SURVEY 8d) / CUDA-event time of the step's k_trace launches (rank 0)"},
                "k_shade": {"achieved": sh_ach, "frac": sh_ach / peak, "ms_per_step": shade_ms / K, "bytes_per_step": sh_bytes, "bytes_per_vertex": b_shade,
                            "definition": "B_shade (SURVEY 8d) x path vertices / CUDA-event time of the step's k_shade launches (rank 0)"},
                "step": {"achieved": st_ach, "frac": st_ach / peak, "ms_per_step": 1e3 * t_steps / K, "bytes_per_step": step_bytes_all,
                         "definition": "(k_trace bytes + B_shade x vertices + B_samp x camera samples, all ranks) / step time (max over ranks) / N GPUs"},
                "unit": "GB/s", "peak": peak},
        }
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_api
            cores = os.cpu_count() or 1
            R, sample = cpu_sample_plan(hs, c)
            orc = oracle_api.OracleScene(hs)
            # one untimed pass first, like the reference arm's warm-up: the spatial light distribution is filled lazily, voxel by
            # voxel under a lock (spatial.rs:200-261), and the first pass over the image pays for most of it (x3 on config 5)
            orc.render(threads=cores, s0=0, s1=1, shard_index=R - 1, shard_count=R)
            crays = csamples = 0
            t = time.time()
            nrep = 0
            while nrep < 4 or (time.time() - t < 10.0 and nrep < 64):      # about 10-20 s of CPU work
                _, so = orc.render(threads=cores, s0=nrep % c["step_spp"], s1=nrep % c["step_spp"] + 1, shard_index=nrep % R, shard_count=R)
                crays += so["regular_rays"] + so["shadow_rays"]
                csamples += so["camera_rays"]
                nrep += 1
            dt = time.time() - t
            orc.close()
            line["cpu_baseline"] = {"value": crays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                    "sample": "%d x (%s) after one untimed pass, %d threads, %.1f s wall; C++ restatement of pbrt-r3's CPU path (oracle/)" % (nrep, sample, cores, dt),
                                    "msamples_per_s": csamples / dt / 1e6}
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
