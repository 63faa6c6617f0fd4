// kernels.h — the beam-step kernels, defined in kern_step.cu (one kernel per object, -DFLT_KERNEL=n) and
// kern_gx.cu, launched from flt_abi.cu: the device-code units of the library compile in parallel (a single
// unit took nine minutes).
#pragma once
#include "beam_core.h"
#if FLT_DEVICE_BUILD
__global__ void flt_k_decode(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_decode512(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_decode1024(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_decode_gmem(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_decode_wide(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_decode512_wide(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_decode_gmem_wide(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_gx_lex(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_gx_lf(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_gx512_lex(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_gx512_lf(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_gx_gmem_lex(flt::DecCfg c, flt::BatchArgs a);
__global__ void flt_k_gx_gmem_lf(flt::DecCfg c, flt::BatchArgs a);
#endif
