// runtime.h — device-runtime shim used by flt_abi.cu. Under nvcc this is the CUDA runtime. Under
// -DFLT_HOST_MODEL (tests/model only, see spmd.h) "device" memory is host memory and a launch is a
// sequential loop over CTAs, so the kernel logic can be exercised without a GPU.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "spmd.h"

#if FLT_DEVICE_BUILD
#include <cuda_runtime.h>
#endif

namespace flt {
namespace rt {

#if FLT_DEVICE_BUILD
using Stream = cudaStream_t;
inline const char* check(cudaError_t e) { return e == cudaSuccess ? nullptr : cudaGetErrorString(e); }
#define FLT_RT_TRY(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) throw std::runtime_error(std::string(#expr) + ": " + \
                                                    cudaGetErrorString(_e));    \
  } while (0)

inline void* devAlloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  FLT_RT_TRY(cudaMalloc(&p, bytes));
  return p;
}
inline void devFree(void* p) {
  if (p) cudaFree(p);
}
inline void h2d(void* d, const void* h, size_t bytes, Stream s) {
  if (bytes) FLT_RT_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
}
inline void d2h(void* h, const void* d, size_t bytes, Stream s) {
  if (bytes) FLT_RT_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
}
inline void devZero(void* d, size_t bytes, Stream s) {
  if (bytes) FLT_RT_TRY(cudaMemsetAsync(d, 0, bytes, s));
}
inline void d2d(void* dst, const void* src, size_t bytes, Stream s) {
  if (bytes) FLT_RT_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
}
inline void sync(Stream s) { FLT_RT_TRY(cudaStreamSynchronize(s)); }
inline bool isDevicePtr(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
inline bool isPinnedPtr(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}
#else
using Stream = void*;
// filled with a garbage pattern: "device" memory is never zero for free, and a kernel that relies on
// it must fail in the logic harness too
inline void* devAlloc(size_t bytes) {
  void* p = malloc(bytes ? bytes : 16);
  if (p) memset(p, 0x5A, bytes ? bytes : 16);
  return p;
}
inline void devFree(void* p) { free(p); }
inline void h2d(void* d, const void* h, size_t bytes, Stream) { memcpy(d, h, bytes); }
inline void d2d(void* dst, const void* src, size_t bytes, Stream) { memcpy(dst, src, bytes); }
inline void d2h(void* h, const void* d, size_t bytes, Stream) { memcpy(h, d, bytes); }
inline void devZero(void* d, size_t bytes, Stream) { memset(d, 0, bytes); }
inline void sync(Stream) {}
inline bool isDevicePtr(const void*) { return true; } // one address space
inline bool isPinnedPtr(const void*) { return false; }
#endif

// growable device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    devFree(p);
    p = nullptr;
    cap = 0;
    p = devAlloc(bytes);
    cap = bytes;
  }
  void release() {
    devFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return (T*)p;
  }
};

} // namespace rt
} // namespace flt
