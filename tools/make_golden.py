"""Generates tests/golden/*: small fixtures produced by the CPU oracle (the Rust reference cannot be built in this
image, so these pin the ORACLE's output, not the reference binary's; see DESIGN.md "parity unpinned").
Run from the repo root:  python tools/make_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
import pbrt_r3_b200 as P
import oracle_api
out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
hs = P.HostScene.parse_string(P.scenes.cornell_box(res=64, spp=4))
orc = oracle_api.OracleScene(hs)
rgb, st = orc.render(threads=1)
np.save(os.path.join(out, "cornell_64_4spp.npy"), rgb)
sb = hs.desc.film.sample_bounds
pix = np.array([[x, y] for y in (sb[1], 0, 17, sb[3] - 1) for x in (sb[0], 0, 31, sb[2] - 1)], np.int32)
sn = (np.arange(len(pix)) % 4).astype(np.int32)
np.savez(os.path.join(out, "cornell_64_streams.npz"), pix=pix, sn=sn, halton=orc.sampler_values(pix, sn, 0, 32), rays=orc.camera_rays(pix, sn)[0],
         counters=np.array([st[k] for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested")], np.int64))
hs2 = P.HostScene.parse_string(P.scenes.cornell_box(res=64, spp=4, sampler="sobol"))
orc2 = oracle_api.OracleScene(hs2)
np.save(os.path.join(out, "cornell_64_sobol_values.npy"), orc2.sampler_values(pix, sn, 0, 32))
# round-1 widening: spot lights + mixed materials (path), mesh tangents / normals (aov "ns"), the LBVH layout's visit counters
from test_hostcheck import aov_scene, spot_scene
hs3 = P.HostScene.parse_string(spot_scene(P, res=32, spp=4))
rgb3, st3 = oracle_api.OracleScene(hs3).render(threads=1)
hs4 = P.HostScene.parse_string(aov_scene(P, "ns", 0.75, res=32, spp=2))
rgb4, _ = oracle_api.OracleScene(hs4).render(threads=1)
hs5 = P.HostScene.parse_string(P.scenes.cornell_box(res=32, spp=2).replace('Accelerator "bvh"', 'Accelerator "bvh" "string layout" "lbvh"'))
rgb5, st5 = oracle_api.OracleScene(hs5).render(threads=1)
np.savez(os.path.join(out, "widening_32.npz"), spot=rgb3, aov_ns=rgb4, lbvh=rgb5,
         spot_counters=np.array([st3[k] for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested")], np.int64),
         lbvh_counters=np.array([st5[k] for k in ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested")], np.int64))
# round-2 widening: ambientocclusion / directlighting / whitted, the infinite light (constant; a map needs a file), uber /
# substrate / translucent
from test_golden import round2_scenes
K = ("camera_rays", "regular_rays", "shadow_rays", "nodes_visited", "tris_tested", "path_vertices")
r2 = {}
for name, txt in round2_scenes(P).items():
    rgb, st = oracle_api.OracleScene(P.HostScene.parse_string(txt)).render(threads=1)
    r2[name] = rgb
    r2[name + "_counters"] = np.array([st[k] for k in K], np.int64)
np.savez(os.path.join(out, "round2_24.npz"), **r2)
print("golden written:", os.listdir(out))
