"""One small render of the bench scene for profiling: a single 3 spp wave (1026^2 x 3 = 3.2 M paths), so that the
launch sequence is raygen, trace #0 (camera rays), shade x classes, trace #1 (continuation + shadow + MIS rays), ...
Usage: python tools/prof_one.py [spp]   (PBRT_R3_B200_LIB selects a variant library)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbrt_r3_b200 as P

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 3
hs = P.HostScene.parse_string(P.scenes.displaced_grid(n=708, res=1024, spp=spp, maxdepth=8))
sc = P.Scene(hs, 0)
for _ in range(2):
    ptr, npx, st = sc.render_partial(spp_begin=0, spp_end=spp)
print("rays %d  trace %.2f ms  shade %.2f ms  device %.2f ms  launches %d" % (st["regular_rays"] + st["shadow_rays"], st["extend_ms"], st["shade_ms"], st["device_ms"], st["kernel_launches"]))
