"""world_size-2 gloo tests (CPU) of the multi-GPU path: the shards partition the samples, the film partials are summed
onto rank 0 with one reduce, and the resolved image equals the single-process render.  The partial films are rendered by
the PRODUCT's own stage code and sharding logic (pt_render_partial with dist.shard_opts, compiled for the host by
tests/hostcheck - there is no GPU in this container); on the GPU box the same path runs with NCCL (bench.py --gpus N,
tests/test_cli.py::test_cli_two_ranks_match_one)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene_text(P, sampler, spp):
    txt = P.scenes.cornell_box(res=32, spp=spp, sampler="halton" if sampler == "stratified" else sampler)
    if sampler == "stratified":      # spp = strata^2; tiles are interleaved over the ranks like the random sampler's
        from test_hostcheck import stratified_variant
        txt = stratified_variant(txt, strata=int(round(spp ** 0.5)))
    # the serial host build of the 64^3-voxel light table is irrelevant here
    return txt.replace('Integrator "path"', 'Integrator "path" "string lightsamplestrategy" ["power"]')


def _worker(rank, world, port, out_path, sampler, spp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    import conftest
    import pbrt_r3_b200 as P
    from pbrt_r3_b200 import dist as D
    os.environ["PBRT_R3_B200_DATA"] = os.path.join(ROOT, "pbrt-r3_b200", "data")
    P.capi._lib = P.capi.load(conftest.HOSTCHECK_LIB)     # built by the parent process
    hs = P.HostScene.parse_string(_scene_text(P, sampler, spp))
    sc = P.Scene(hs)
    opts = D.shard_opts(hs.desc.sampler.type, hs.desc.sampler.samples_per_pixel, rank, world)
    ptr, n, st = sc.render_partial(**opts)
    film = torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), (n * 4,)).copy())
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
    assert D.shard_opts(3, 16, 2, 4) == {"shard_index": 2, "shard_count": 4}      # stratified: per-tile streams too
    assert D.shard_opts(4, 6, 0, 2) == {"shard_index": 0, "shard_count": 2} and D.effective_spp(4, 6) == 8   # 02sequence: tiles; spp rounds up
    # Sobol takes the next power of two (sobol.rs:24): 6 requested = 8 rendered, and all 8 are sharded
    assert D.effective_spp(1, 6) == 8 and D.effective_spp(0, 6) == 6
    assert [D.shard_opts(1, 6, r, 2) for r in range(2)] == [{"spp_begin": 0, "spp_end": 4}, {"spp_begin": 4, "spp_end": 8}]


import pytest


@pytest.mark.parametrize("sampler,spp", [("halton", 8), ("sobol", 6), ("random", 2), ("stratified", 4)])
def test_two_rank_film_reduce_matches_single_process(HC, tmp_path, sampler, spp):
    import oracle_api
    P = HC
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, 29000 + os.getpid() % 2000, out, sampler, spp), nprocs=2, join=True)
    hs = P.HostScene.parse_string(_scene_text(P, sampler, spp))
    one, _ = P.Scene(hs).render()                                   # the product, one process
    ref, _ = oracle_api.OracleScene(hs).render(threads=2)           # the oracle
    got = np.load(out)
    assert np.sqrt(((got - one) ** 2).mean()) / one.mean() < 1e-5
    assert np.sqrt(((got - ref) ** 2).mean()) / ref.mean() < 1e-5
