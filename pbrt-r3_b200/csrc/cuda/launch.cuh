// launch.cuh — launch macro and launch-shape constants shared by the translation units of the library (device_api.cu
// and shade_tu.cu, the shading stage; both are compiled with exact IEEE division / square root and without FMA
// contraction - `make FAST_SHADE=1` is an opt-in experiment, never used for a reported number).
#pragma once
#if defined(PT_HOSTCHECK)
// tests/hostcheck: the SAME sources compiled by g++ against a serial stand-in for the CUDA runtime, to check the
// stage logic on a CPU-only box.  Never built into the product library (which needs a CUDA device).
#include "cuda_shim.h"
#define PT_LAUNCH(grid, block, stream, ...) hostcheck::begin_launch(), __VA_ARGS__
#else
#include <cuda_runtime.h>
#define PT_LAUNCH(grid, block, stream, ...) __VA_ARGS__<<<(grid), (block), 0, (stream)>>>
#endif
// spectral channel count of the second template instantiation (tests/hostcheck may build RGB only to save compile time)
#ifndef PT_NCS
#define PT_NCS 60
#endif

#ifndef PT_SHADE_BLOCK
#define PT_SHADE_BLOCK 128
#endif
#ifndef PT_SHADE_MINBLOCKS
#define PT_SHADE_MINBLOCKS 4
#endif
// CTAs per SM of the shade kernels' launch.  A persistent grid (8 per SM) with a grid-stride loop balances badly here:
// vertices differ a lot in cost (lobe sampled, light type, MIS branch, termination) and every CTA is resident for the
// whole kernel; with about one vertex per thread the block scheduler does the balancing: k_shade 96.4 -> 80.5 ms per
// step on config 2 (profiles/r1_shade_fastmath.md)
#ifndef PT_SHADE_GRID_PER_SM
#define PT_SHADE_GRID_PER_SM 128
#endif
// CTAs per SM of the streaming kernels (raygen, finish, film gather)
#ifndef PT_STREAM_GRID_PER_SM
#define PT_STREAM_GRID_PER_SM 128
#endif
// 60-bin spectra: three live spectra are 180 registers; at the RGB kernels' 128-register cap the spectral shade kernels
// spill 2.2 KB per thread, so they trade occupancy for registers (255 registers, 2 CTAs / SM)
#ifndef PT_SHADE_MINBLOCKS_SPECTRAL
#define PT_SHADE_MINBLOCKS_SPECTRAL 2
#endif
