"""Writes the synthetic BASELINE scenes as .pbrt files for tools/run_reference.sh (sizes a CPU renders in minutes; the
geometry generators and seeds are the ones bench.py and the tests use).  Usage: python tools/make_reference_scenes.py <dir>"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbrt_r3_b200 import scenes

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
S = {
    "config1_cornell_256_16spp": scenes.cornell_box(res=256, spp=16, maxdepth=5),
    "config1_cornell_sobol": scenes.cornell_box(res=128, spp=16, maxdepth=5, sampler="sobol"),
    "config2_grid_1m_256_4spp": scenes.displaced_grid(n=708, res=256, spp=4, maxdepth=8),
    "config3_instanced_480x270_2spp": scenes.instanced_blobs(n_instances=1000, n_sub=5, res=(480, 270), spp=2, maxdepth=8, n_lights=256),
    "config4_spectral_cornell_128_8spp": scenes.cornell_box(res=128, spp=8, maxdepth=5),
    "mixed_materials_96_16spp": scenes.cornell_box(res=96, spp=16, materials="mixed", extra_lights=True, maxdepth=8),
}
for name, txt in S.items():
    open(os.path.join(out, name + ".pbrt"), "w").write(txt)
    print("wrote", name)
