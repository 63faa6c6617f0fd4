// kern_step.cu — the generic beam step (beam_core.h decodeCta): one kernel per compilation, selected with
// -DFLT_KERNEL=<n> (see the Makefile), so that the seven variants compile in parallel instead of one after the
// other in one ptxas run. <false>: max-merge, word-level LM; <true> (`_wide`): with the full-expansion paths
// (logAdd merging, token-level LMs, unranked rows walking the token list), which slowed the others by 6 % when
// they shared a body.
#include "beam_core.h"
#include "beam_lf.h"
#include "kernels.h"
using namespace flt;
#if FLT_DEVICE_BUILD
#if FLT_KERNEL == 0 // flt_k_decode
// workspace in shared memory (the fast path: every access is an LDS/STS with constant-bank offsets)
__global__ void __launch_bounds__(256) flt_k_decode(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  decodeCta<false>(cta, c, a, smem);
}
#endif
#if FLT_KERNEL == 1 // flt_k_decode512
// same, 512 threads per utterance (two CTAs per SM): small batches leave SMs under-occupied
__global__ void __launch_bounds__(512, 2) flt_k_decode512(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  decodeCta<false>(cta, c, a, smem);
}
#endif
#if FLT_KERNEL == 2 // flt_k_decode1024
// 1024 threads per utterance: beams so wide (K = 500) that the small workspace region leaves room for one
// CTA per SM only — the items of a frame (thousands) are then spread over all 32 warps the SM can hold
__global__ void __launch_bounds__(1024, 1) flt_k_decode1024(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  decodeCta<false>(cta, c, a, smem);
}
#endif
#if FLT_KERNEL == 3 // flt_k_decode_gmem
// workspace in a global slab per CTA (beams / candidate sets too large for shared memory)
__global__ void __launch_bounds__(256) flt_k_decode_gmem(DecCfg c, BatchArgs a) {
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  decodeCta<false>(cta, c, a, a.wsGlobal + (long long)blockIdx.x * a.wsStride);
}
#endif
#if FLT_KERNEL == 4 // flt_k_decode_wide
// the same three with the full-expansion paths compiled in (DecCfg::wide: logAdd merging, token-level
// LMs, unranked rows walking the token list); kept out of the kernels above, which they slowed by 6 %
__global__ void __launch_bounds__(256) flt_k_decode_wide(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  decodeCta<true>(cta, c, a, smem);
}
#endif
#if FLT_KERNEL == 5 // flt_k_decode512_wide
__global__ void __launch_bounds__(512, 2) flt_k_decode512_wide(DecCfg c, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  decodeCta<true>(cta, c, a, smem);
}
#endif
#if FLT_KERNEL == 6 // flt_k_decode_gmem_wide
__global__ void __launch_bounds__(256) flt_k_decode_gmem_wide(DecCfg c, BatchArgs a) {
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  decodeCta<true>(cta, c, a, a.wsGlobal + (long long)blockIdx.x * a.wsStride);
}
#endif
#endif
