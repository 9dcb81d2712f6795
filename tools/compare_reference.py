"""Second half of tools/run_reference.sh: the reference's EXR renders vs the oracle (and the CUDA library when a GPU is
visible) on the same .pbrt files; keeps the reference images as tests/golden/ref_<scene>.npy.
Usage: python tools/compare_reference.py <dir with *.pbrt and *.exr>"""
import glob
import os
import sys

import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api
import pbrt_r3_b200 as P
from pbrt_r3_b200 import imageio


def rel_rmse(a, b):
    return float(np.sqrt(((a.astype(np.float64) - b) ** 2).mean()) / max(float(b.mean()), 1e-20))


d = sys.argv[1]
bad = 0
for f in sorted(glob.glob(os.path.join(d, "*.pbrt"))):
    name = os.path.basename(f)[:-5]
    exr = os.path.join(d, name + ".exr")
    if not os.path.exists(exr):
        print("%-40s no reference image" % name)
        continue
    ref = imageio.read_exr(exr)
    np.save(os.path.join(ROOT, "tests", "golden", "ref_%s.npy" % name), ref)
    hs = P.HostScene.parse_file(f, spectrum="sampled" if "spectral" in name else "rgb")
    orc = oracle_api.OracleScene(hs).render()[0]
    line = "%-40s oracle vs reference rel RMSE %.3g" % (name, rel_rmse(orc, ref))
    ok = rel_rmse(orc, ref) <= 1e-3
    try:
        gpu = P.Scene(hs, 0).render()[0]
        line += "   CUDA vs reference %.3g" % rel_rmse(gpu, ref)
        ok = ok and rel_rmse(gpu, ref) <= 1e-3
    except P.PbrtError:
        line += "   (no CUDA device)"
    print(line, "" if ok else "  <-- above the 1e-3 bar")
    bad += 0 if ok else 1
sys.exit(1 if bad else 0)
