"""Experiment: does running two independent wave pipelines CONCURRENTLY on one GPU (trace of one overlapping shade of the
other) beat one pipeline?  Two Scene objects of the same scene (own streams, own path state), each rendering half of the
sample numbers from its own host thread, against one Scene rendering all of them.  Wall time around a device sync."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pbrt_r3_b200 as P
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if cfg == 3:
    text = P.scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(1920, 1080), spp=256, maxdepth=8, n_lights=256)
else:
    text = P.scenes.displaced_grid(n=708, res=1024, spp=64, maxdepth=8)
hs = P.HostScene.parse_string(text)
a, b = P.Scene(hs, 0), P.Scene(hs, 0)

def run(sc, s0, s1, **kw):
    sc.render_partial(spp_begin=s0, spp_end=s1, **kw)

def timed(fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / n * 1e3

def both(**kw):
    t = [threading.Thread(target=run, args=(a, 0, spp // 2), kwargs=kw), threading.Thread(target=run, args=(b, spp // 2, spp), kwargs=kw)]
    for x in t: x.start()
    for x in t: x.join()

print("cfg", cfg, "spp", spp)
print("one pipeline          : %.2f ms" % timed(lambda: run(a, 0, spp)))
print("two sequential halves : %.2f ms" % timed(lambda: (run(a, 0, spp // 2), run(b, spp // 2, spp))))
print("two concurrent halves : %.2f ms" % timed(both))
for seg in (3, 4):
    print("two concurrent halves, %d waves in flight each: %.2f ms" % (seg, timed(lambda: both(wave_segments=seg))))
