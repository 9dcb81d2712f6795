// shade_tu.cu — the shading stage (k_shade: li() loop body after the intersection, dshade.cuh / dbsdf.cuh) as its own
// translation unit, so that it can be compiled on its own (parallel build) and, for experiments, with other float flags
// (Makefile FAST_SHADE=1: approximate division / sqrt; measured in profiles/r1_shade_fastmath.md - 25 % faster shade
// kernels, but not parity-safe, so the product build keeps -fmad=false -prec-div=true -prec-sqrt=true here too).
// Sampler values (div_rn) and the light-distribution voxel are exact under either setting.
#if !defined(PT_HOSTCHECK)
#include "launch.cuh"
#include "dshade.cuh"
#endif

namespace ptd {

// One shade kernel per shading class; the bin size is read on the device (no host round trip between trace and shade).
// ENV = the scene has infinite lights (LightSource "infinite"): a second set of kernels, so that every other scene runs
// exactly the code it ran before the environment light existed.
// CTAs per SM the register allocation aims at.  The specular classes (glass, mirror) do no light sampling: little
// arithmetic between scattered loads, so they take more resident warps over registers (PT_SHADE_MINBLOCKS_SPEC).
#ifndef PT_SHADE_MINBLOCKS_SPEC
#define PT_SHADE_MINBLOCKS_SPEC 8
#endif
#ifndef PT_SHADE_MINBLOCKS_MATTE
#define PT_SHADE_MINBLOCKS_MATTE PT_SHADE_MINBLOCKS
#endif
#ifndef PT_SHADE_MINBLOCKS_PLASTIC
#define PT_SHADE_MINBLOCKS_PLASTIC PT_SHADE_MINBLOCKS
#endif
#ifndef PT_SHADE_MINBLOCKS_METAL
#define PT_SHADE_MINBLOCKS_METAL PT_SHADE_MINBLOCKS
#endif
constexpr int shade_minblocks(int nc, int cls) {
  return nc != 3 ? PT_SHADE_MINBLOCKS_SPECTRAL
                 : ((cls == CL_GLASS || cls == CL_MIRROR) ? PT_SHADE_MINBLOCKS_SPEC
                                                          : (cls == CL_MATTE ? PT_SHADE_MINBLOCKS_MATTE : (cls == CL_PLASTIC ? PT_SHADE_MINBLOCKS_PLASTIC : (cls == CL_METAL ? PT_SHADE_MINBLOCKS_METAL : PT_SHADE_MINBLOCKS))));
}
template <int NC, int CLS, bool ENV>
__global__ void __launch_bounds__(PT_SHADE_BLOCK, shade_minblocks(NC, CLS)) k_shade(const __grid_constant__ DScene S, WaveState W, Queues Q, unsigned long long* vertex_counter) {
  const uint32_t n = Q.counts[8 + CLS];
  const uint32_t* bin = Q.bins + (size_t)CLS * W.cap;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) shade_body<NC, CLS, ENV>(S, W, Q, bin[k], vertex_counter);
}


template <int NC, bool ENV>
static void launch_shade_nc(int cls, int grid, cudaStream_t stream, const DScene& S, const WaveState& W, const Queues& Q, unsigned long long* vc) {
  switch (cls) {
    case CL_NONE: PT_LAUNCH(grid, PT_SHADE_BLOCK, stream, k_shade<NC, CL_NONE, ENV>)(S, W, Q, vc); break;
    case CL_MATTE: PT_LAUNCH(grid, PT_SHADE_BLOCK, stream, k_shade<NC, CL_MATTE, ENV>)(S, W, Q, vc); break;
    case CL_PLASTIC: PT_LAUNCH(grid, PT_SHADE_BLOCK, stream, k_shade<NC, CL_PLASTIC, ENV>)(S, W, Q, vc); break;
    case CL_GLASS: PT_LAUNCH(grid, PT_SHADE_BLOCK, stream, k_shade<NC, CL_GLASS, ENV>)(S, W, Q, vc); break;
    case CL_METAL: PT_LAUNCH(grid, PT_SHADE_BLOCK, stream, k_shade<NC, CL_METAL, ENV>)(S, W, Q, vc); break;
    case CL_MIRROR: PT_LAUNCH(grid, PT_SHADE_BLOCK, stream, k_shade<NC, CL_MIRROR, ENV>)(S, W, Q, vc); break;
    case CL_EXTRA:  // uber / substrate / translucent: RGB only (front end)
      if constexpr (NC == 3) PT_LAUNCH(grid, PT_SHADE_BLOCK, stream, k_shade<NC, CL_EXTRA, ENV>)(S, W, Q, vc);
      break;
    default: break;
  }
}

}  // namespace ptd

void launch_shade(int n_channels, int cls, int grid, cudaStream_t stream, const ptd::DScene& S, const ptd::WaveState& W, const ptd::Queues& Q,
                  unsigned long long* vertex_counter) {
  if (n_channels == 3) {
    if (S.n_inf != 0u || S.has_map_lights) ptd::launch_shade_nc<3, true>(cls, grid, stream, S, W, Q, vertex_counter);
    else ptd::launch_shade_nc<3, false>(cls, grid, stream, S, W, Q, vertex_counter);
  } else ptd::launch_shade_nc<PT_NCS, false>(cls, grid, stream, S, W, Q, vertex_counter);  // infinite lights: RGB only (front end)
}
