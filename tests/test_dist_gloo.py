"""world_size-2 gloo test (CPU) of the multi-GPU host logic: spp shards partition the samples, the film partials are
summed onto rank 0 with one reduce, and the resolved image equals the single-process render.  The partial films come
from the ORACLE here (no GPU in this container); on the GPU box the same code path runs with NCCL (bench.py --gpus N)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_api
    import pbrt_r3_b200 as P
    from pbrt_r3_b200 import dist as D
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=32, spp=8))
    orc = oracle_api.OracleScene(hs)
    s0, s1 = D.spp_shard(8, rank, world)
    _, _, xyzw = orc.render(threads=2, s0=s0, s1=s1, want_xyzw=True)
    film = torch.from_numpy(xyzw.reshape(-1).copy())
    D.reduce_film(film, dst=0)
    if rank == 0:
        np.save(out_path, D.resolve_host(film.numpy(), hs.desc.film.scale).reshape(32, 32, 3))
    dist.barrier()
    dist.destroy_process_group()


def test_spp_shards_partition():
    from pbrt_r3_b200 import dist as D
    for spp in (1, 7, 16, 64, 1000):
        for world in (1, 2, 3, 8):
            r = [D.spp_shard(spp, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_shard_opts_by_sampler():
    from pbrt_r3_b200 import dist as D
    assert D.shard_opts(0, 64, 1, 4) == {"spp_begin": 16, "spp_end": 32}
    assert D.shard_opts(1, 64, 3, 4) == {"spp_begin": 48, "spp_end": 64}
    assert D.shard_opts(2, 64, 1, 4) == {"shard_index": 1, "shard_count": 4}


def test_two_rank_film_reduce_matches_single_process(P, tmp_path):
    import oracle_api
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, 29000 + os.getpid() % 2000, out), nprocs=2, join=True)
    hs = P.HostScene.parse_string(P.scenes.cornell_box(res=32, spp=8))
    ref, _ = oracle_api.OracleScene(hs).render(threads=2)
    got = np.load(out)
    assert np.sqrt(((got - ref) ** 2).mean()) / ref.mean() < 1e-5
