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
# round-2 widening: the integrators beside "path", the infinite light, uber / substrate / translucent, hlbvh, loopsubdiv
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pbrt_r3_b200 as P
import test_hostcheck as T
S.update({
    "r2_ambientocclusion_64_4spp": T.ao_scene(P, nsamples=16, res=64, spp=4),
    "r2_directlighting_all_64_4spp": T.tree_scene(P, '"directlighting"', res=64, spp=4, nsamples=2),
    "r2_directlighting_one_64_4spp": T.tree_scene(P, '"directlighting" "string strategy" ["one"]', res=64, spp=4),
    "r2_whitted_64_4spp": T.tree_scene(P, '"whitted"', res=64, spp=4),
    "r2_infinite_constant_64_8spp": T.env_scene(P, out, mapped=False, res=64, spp=8),
    "r2_infinite_map_64_8spp": T.env_scene(P, out, mapped=True, res=64, spp=8),
    "r2_extra_materials_64_8spp": T.extra_materials_scene(P, res=64, spp=8),
    "r2_hlbvh_grid_64_4spp": scenes.displaced_grid(n=96, res=64, spp=4).replace('Accelerator "bvh"', 'Accelerator "bvh" "string splitmethod" ["hlbvh"]'),
})
for name, txt in S.items():
    open(os.path.join(out, name + ".pbrt"), "w").write(txt)
    print("wrote", name)
