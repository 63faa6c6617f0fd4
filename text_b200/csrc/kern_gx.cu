// kern_gx.cu — the single-pass step with a guessed cut (beam_gx.h), opt-in (FLT_GX=1)
#include "beam_core.h"
#include "beam_lf.h"
#include "beam_gx.h"
#include "kernels.h"
using namespace flt;
#if FLT_DEVICE_BUILD
// the single-pass step with a guessed cut (beam_gx.h), two-kernel path: token lists from flt_k_topm
__global__ void __launch_bounds__(256) flt_k_gx_lex(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  gxDecodeCta<true>(cta, c, a, smem);
}
__global__ void __launch_bounds__(256) flt_k_gx_lf(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  gxDecodeCta<false>(cta, c, a, smem);
}
__global__ void __launch_bounds__(512, 2) flt_k_gx512_lex(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  gxDecodeCta<true>(cta, c, a, smem);
}
__global__ void __launch_bounds__(512, 2) flt_k_gx512_lf(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  gxDecodeCta<false>(cta, c, a, smem);
}
__global__ void __launch_bounds__(256) flt_k_gx_gmem_lex(DecCfg c, BatchArgs a) {
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  gxDecodeCta<true>(cta, c, a, a.wsGlobal + (long long)blockIdx.x * a.wsStride);
}
__global__ void __launch_bounds__(256) flt_k_gx_gmem_lf(DecCfg c, BatchArgs a) {
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  gxDecodeCta<false>(cta, c, a, a.wsGlobal + (long long)blockIdx.x * a.wsStride);
}
#endif
