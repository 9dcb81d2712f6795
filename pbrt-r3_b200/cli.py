"""`pbrt-r3` command line on the GPU path (reference src/bin/pbrt.rs:44-135 options, :219-256 create_scene_and_integrator,
:258-356 render_scene, :358-402 main): same options, same overrides (--pixelsamples replaces the Sampler parameter,
--outfile the Film filename, --quick / --quick-full-resolution set PbrtOptions), same exit codes (-1 on error), and the
image is written by the film's file name extension like Film::write_image.

    python -m pbrt_r3_b200 scene.pbrt [-o out.exr] [-s 64] [--quick] [--quiet] [--stats] [--gpus N]

Options of the reference that only make sense for its CPU renderer or its other subsystems are accepted and ignored
(--nthreads, --logdir, --logtostderr, --minloglevel, --profile, --display-server, --sequential-display, and --cropwindow,
which the reference parses but never applies: PbrtOptions.crop_window is commented out, pbrt_options.rs:9); --cat / --toply
(scene reformatting) are not part of the render path and exit with an error.  --gpus N (extension) renders with N
processes, one per GPU, sample-sharded, with the film partials summed by NCCL (pbrt_r3_b200.dist)."""
import argparse
import os
import sys
import time


def build_parser():
    ap = argparse.ArgumentParser(prog="pbrt-r3", add_help=False, description="pbrt-r3 PathIntegrator on B200")
    ap.add_argument("-i", "--infile", metavar="filename")
    ap.add_argument("-o", "--outfile", metavar="filename")
    ap.add_argument("--cropwindow", metavar="x0,x1,y0,y1")
    ap.add_argument("-h", "--help", action="help")
    ap.add_argument("-j", "--nthreads", type=int, metavar="num")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--logdir", metavar="dir")
    ap.add_argument("--logtostderr", action="store_true")
    ap.add_argument("--minloglevel", type=int, metavar="num")
    ap.add_argument("-c", "--cat", action="store_true")
    ap.add_argument("-t", "--toply", action="store_true")
    ap.add_argument("--display-server", metavar="url")
    ap.add_argument("-s", "--pixelsamples", type=int, metavar="num")
    ap.add_argument("--quick-full-resolution", "--quick_full_resolution", dest="quick_full_resolution", action="store_true")
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("-k", "--sequential-display", metavar="dir")
    ap.add_argument("--spectrum", default="rgb", choices=["rgb", "sampled"], help="the reference's `sampled-spectrum` feature build (60 bins)")
    ap.add_argument("--gpus", type=int, default=1, help="GPUs of this node (one process each)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("pbrtfile", nargs="*", metavar="filename.pbrt")
    return ap


def print_stats(st, seconds, out=sys.stdout):
    rays = st["regular_rays"] + st["shadow_rays"]
    out.write("Statistics:\n")
    out.write("  Integrator\n    Camera rays traced %30d\n" % st["camera_rays"])
    out.write("  Intersections\n    Regular ray intersection tests %18d\n    Shadow ray intersection tests %19d\n" % (st["regular_rays"], st["shadow_rays"]))
    out.write("  GPU\n    Kernel launches %33d\n    Device time (ms) %32.1f\n    Mrays/s %41.1f\n    Msamples/s %38.1f\n" % (
        st["kernel_launches"], st["device_ms"], rays / max(st["device_ms"], 1e-9) / 1e3, st["camera_rays"] / max(st["device_ms"], 1e-9) / 1e3))
    out.write("  Wall time (s) %37.2f\n" % seconds)


def render_scene(path, opts):
    import pbrt_r3_b200 as P
    from . import capi, imageio
    lib = capi.load()
    o = capi.pt_options(1 if (opts.quick or opts.quick_full_resolution) else 0, 1 if opts.quick_full_resolution else 0)
    lib.pt_set_options(o)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(opts.device)))
    if world > 1:
        import torch
        local = local % max(torch.cuda.device_count(), 1)   # more ranks than GPUs (tests on a 1-GPU box): ranks share devices
    if not opts.quiet and rank == 0:
        print("pbrt-r3 (B200 PathIntegrator core) [%d GPU process%s]" % (world, "" if world == 1 else "es"))
    t0 = time.time()
    ps = max(1, opts.pixelsamples) if opts.pixelsamples is not None else -1
    hs = P.HostScene.parse_file(path, ps, opts.spectrum)
    name = opts.outfile if opts.outfile else hs.filename
    if rank == 0 and not opts.quiet:
        for w in hs.warnings:
            sys.stderr.write("WARNING: %s\n" % w)
    sc = P.Scene(hs, local)
    try:
        if world == 1:
            rgb, st = sc.render()
        else:
            from . import dist as D
            rgb, st = D.render_distributed(sc, hs, rank, world, local)
    finally:
        sc.close()
    if rank == 0:
        # Film displays (pbrt.rs:296-321): --display-server host:port (tev), -k dir (numbered EXR frames); a display that
        # cannot be reached is a warning, not an error
        from . import display
        shows = []
        if opts.display_server:
            try:
                shows.append(display.TevDisplay(opts.display_server))
            except Exception as e:
                sys.stderr.write("WARNING: %s\n" % e)
        if opts.sequential_display:
            shows.append(display.SequentialDisplay(opts.sequential_display))
        try:
            display.show(shows, hs, rgb)
        except Exception as e:
            sys.stderr.write("WARNING: %s\n" % e)
        for d in shows:
            d.close()
        imageio.write_image(name, rgb)
        if not opts.quiet:
            print("\n")
            if opts.stats:
                print_stats(st, time.time() - t0)
    return 0


def main(argv=None):
    opts = build_parser().parse_args(argv)
    path = opts.pbrtfile[0] if opts.pbrtfile else opts.infile
    if not path or not os.path.exists(path):
        print(build_parser().format_usage())
        return -1
    if opts.cat or opts.toply:
        sys.stderr.write("--cat / --toply reformat the scene file and do not render: not part of the GPU path\n")
        return -1
    if opts.gpus > 1 and "RANK" not in os.environ:
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(opts.gpus), "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29533"), "-m", "pbrt_r3_b200"] + (argv if argv is not None else sys.argv[1:])
        # the ranks keep the caller's working directory (relative scene / output paths mean the same thing there); the
        # package is found through PYTHONPATH
        env = dict(os.environ)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env["PYTHONPATH"] = root + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
        return subprocess.call(cmd, env=env)
    try:
        return render_scene(os.path.realpath(path), opts)
    except Exception as e:  # errors are logged and the exit code is -1 (pbrt.rs:334-338)
        sys.stderr.write("ERROR: %s\n" % e)
        return -1


if __name__ == "__main__":
    sys.exit(main())
