"""Runs the command line with --gpus N and with one GPU on the same scene and compares the images (needs N GPUs).
Usage: python tools/check_cli_gpus.py [N]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbrt_r3_b200 as P
from pbrt_r3_b200 import imageio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
td = tempfile.mkdtemp()
for name, txt in (("halton", P.scenes.cornell_box(res=96, spp=8, materials="mixed", extra_lights=True, maxdepth=6)),
                  ("random", P.scenes.cornell_box(res=96, spp=4, sampler="random"))):
    scene = os.path.join(td, name + ".pbrt")
    open(scene, "w").write(txt)
    outs = []
    for g in (1, n):
        out = os.path.join(td, "%s_%d.exr" % (name, g))
        r = subprocess.run([sys.executable, "-m", "pbrt_r3_b200", scene, "-o", out, "--gpus", str(g), "--quiet"], cwd=ROOT, capture_output=True, text=True)
        assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
        outs.append(imageio.read_exr(out))
    a, b = outs
    rel = float(np.sqrt(((a - b) ** 2).mean()) / a.mean())
    print("%s: %d GPUs vs 1 GPU relative RMSE %.3g (film partial sums in a different order)" % (name, n, rel))
    assert rel < 1e-5
print("ok")
