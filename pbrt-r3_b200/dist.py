"""Multi-GPU plumbing: one process per GPU, spp (or tile) shards, ONE exchange step — the film partials {X,Y,Z,sum_w}
are summed onto rank 0 (torch.distributed reduce: NCCL over NVLink on GPUs, gloo in the CPU tests) and resolved
there.  Mirrors Film::merge_film_tile's sum (reference src/core/film/film.rs:219-241) across processes; the scene is
replicated, nothing else is exchanged."""
import numpy as np


def spp_shard(spp_total, rank, world):
    """Contiguous sample-number range [begin, end) of this rank; every sample index is rendered exactly once."""
    base, rem = divmod(int(spp_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def effective_spp(sampler_type, spp):
    """Samples per pixel the sampler really takes: SobolSampler rounds the request up to a power of two
    (samplers/sobol.rs:24), as pt_scene_create does on the device."""
    spp = max(int(spp), 1)
    if int(sampler_type) in (1, 4):   # PT_SAMPLER_SOBOL, PT_SAMPLER_ZEROTWO
        return 1 << (spp - 1).bit_length()
    return spp


def shard_opts(sampler_type, spp_total, rank, world):
    """pt_render_partial options of this rank.  Halton / Sobol samples are pure functions of (pixel, sample number):
    shard the sample numbers (of the EFFECTIVE count, see effective_spp).  Sampler "random" draws from one sequential
    PCG32 stream per 16x16 render tile (samplers/random.rs), so only whole tiles can be split: interleave the tiles
    across ranks."""
    if int(sampler_type) >= 2:   # PT_SAMPLER_RANDOM, PT_SAMPLER_STRATIFIED
        return {"shard_index": int(rank), "shard_count": int(world)}
    b, e = spp_shard(effective_spp(sampler_type, spp_total), rank, world)
    return {"spp_begin": b, "spp_end": e}


class _CudaArray:
    """Zero-copy view of a device pointer for torch.as_tensor (__cuda_array_interface__ v2)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def film_tensor(dev_ptr, n_pixels, device):
    import torch
    return torch.as_tensor(_CudaArray(dev_ptr, n_pixels * 4), device=device)


def reduce_film(film, dst=0, group=None):
    """Sum the film partials of all ranks into rank `dst` (in place).  `film` is a float32 tensor (n_pixels*4)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
        if film.is_cuda:
            # NCCL is ordered on torch's current stream only; pt_film_resolve launches on the library's own stream
            import torch
            torch.cuda.current_stream(film.device).synchronize()
    return film


def resolve_host(xyzw, scale=1.0):
    """Film::write_image's arithmetic on a host array (used by the gloo tests to check the reduced film):
    rgb = xyz_to_rgb(xyz) / sum_w, clamped at 0, times scale (film.rs:447-472)."""
    p = np.asarray(xyzw, np.float32).reshape(-1, 4)
    x, y, z, w = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    f = np.float32
    c = np.stack([f(3.240479) * x - f(1.537150) * y - f(0.498535) * z, f(-0.969256) * x + f(1.875991) * y + f(0.041556) * z,
                  f(0.055648) * x - f(0.204043) * y + f(1.057311) * z], 1).astype(np.float32)
    pos = w > 0
    inv = np.where(pos, f(1.0) / np.where(pos, w, f(1.0)), f(1.0)).astype(np.float32)
    c = np.where(pos[:, None], np.maximum(f(0.0), c * inv[:, None]), c)
    return (c * f(scale)).astype(np.float32)


def render_distributed(scene, host_scene, rank, world, local_device, backend=None):
    """One Integrator::render over `world` processes (one per GPU): every rank renders its shard of the samples (tiles for
    Sampler "random") into a partial film, one reduce sums the partials onto rank 0, which resolves the image.
    backend "nccl": one GPU per rank, the partial films are reduced in place over NVLink.  backend "gloo" (chosen when the
    node has fewer GPUs than ranks, e.g. the 2-rank test on a 1-GPU box where the ranks share the device): the partials
    are staged through host memory for the reduce.  Returns (rgb or None on ranks > 0, stats of this rank)."""
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    if not dist.is_initialized():
        import os
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_device)
        if backend == "nccl":
            dist.init_process_group(backend, device_id=torch.device("cuda", local_device))
        else:
            dist.init_process_group(backend)
    d = host_scene.desc
    opts = shard_opts(d.sampler.type, d.sampler.samples_per_pixel, rank, world)
    ptr, npx, st = scene.render_partial(**opts)
    film = film_tensor(ptr, npx, "cuda:%d" % local_device)
    if dist.get_backend() == "nccl":
        reduce_film(film, dst=0)
    else:
        host = film.cpu()
        reduce_film(host, dst=0)
        if rank == 0:
            film.copy_(host)
            torch.cuda.synchronize(local_device)
    rgb = scene.film_resolve(film.data_ptr()) if rank == 0 else None
    dist.barrier()
    return rgb, st
