// spmd.h — the thin layer the decode kernels are written against.
//
// Kernels are written as CTA-wide SPMD programs: strided loops over items (`for i = tid; i < n;
// i += nthr`) separated by CTA barriers, plus a handful of atomics and CTA collectives. Compiled
// by nvcc this is plain CUDA for sm_100a (the product). The same source also compiles with g++
// under -DFLT_HOST_MODEL as a one-thread-per-CTA sequential program (barriers are no-ops, atomics
// are plain read-modify-writes): that build exists ONLY as a test harness for kernel logic in the
// GPU-less build container (tests/model/); it is never part of the shipped library, which has no
// CPU path and fails loudly without a CUDA device.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(FLT_HOST_MODEL)
#define FLT_DEVICE_BUILD 1
#define FLT_DEV __device__ __forceinline__
#define FLT_HD __host__ __device__ __forceinline__
#else
#define FLT_DEVICE_BUILD 0
#define FLT_DEV inline
#define FLT_HD inline
#include <cmath>
struct float4 {
  float x, y, z, w;
};
struct int2 {
  int x, y;
};
struct int4 {
  int x, y, z, w;
};
#endif

namespace flt {

struct Cta {
  int tid;  // thread index in the CTA
  int nthr; // threads per CTA
  int bid;  // CTA index
  int nblk; // CTAs in the grid
  int bar = 0; // 0 = the whole CTA (__syncthreads); else a named barrier over this role's nthr threads
  FLT_DEV void sync() const {
#if FLT_DEVICE_BUILD
    if (bar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory");
#endif
  }
};

/* ---- bit casts (memcpy compiles to a register move on both sides) ---- */
FLT_HD uint64_t f64Bits(double x) {
  uint64_t b;
  memcpy(&b, &x, 8);
  return b;
}
FLT_HD double bitsF64(uint64_t b) {
  double x;
  memcpy(&x, &b, 8);
  return x;
}
FLT_HD uint32_t f32Bits(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  return b;
}
FLT_HD float bitsF32(uint32_t b) {
  float x;
  memcpy(&x, &b, 4);
  return x;
}
FLT_HD bool isNegInf(float x) { return f32Bits(x) == 0xFF800000u; }
// Order-preserving maps: a > b  <=>  key(a) > key(b) (as unsigned); -0.0 < +0.0, NaNs excluded.
FLT_HD uint64_t orderedKey64(double s) {
  uint64_t b = f64Bits(s);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
FLT_HD uint32_t orderedKey32(float s) {
  uint32_t b = f32Bits(s);
  return (b >> 31) ? ~b : (b | 0x80000000u);
}
FLT_HD float orderedKey32Inv(uint32_t k) {
  uint32_t b = (k >> 31) ? (k & 0x7FFFFFFFu) : ~k;
  return bitsF32(b);
}

/* ---- atomics (CTA-private or global memory) ---- */
FLT_DEV int atomAdd(int* p, int v) {
#if FLT_DEVICE_BUILD
  return atomicAdd(p, v);
#else
  int o = *p;
  *p = o + v;
  return o;
#endif
}
FLT_DEV int atomCAS(int* p, int cmp, int val) {
#if FLT_DEVICE_BUILD
  return atomicCAS(p, cmp, val);
#else
  int o = *p;
  if (o == cmp) *p = val;
  return o;
#endif
}
FLT_DEV int atomMin(int* p, int v) {
#if FLT_DEVICE_BUILD
  return atomicMin(p, v);
#else
  int o = *p;
  if (v < o) *p = v;
  return o;
#endif
}
FLT_DEV int atomMax(int* p, int v) {
#if FLT_DEVICE_BUILD
  return atomicMax(p, v);
#else
  int o = *p;
  if (v > o) *p = v;
  return o;
#endif
}
FLT_DEV unsigned long long atomCAS64(unsigned long long* p, unsigned long long cmp,
                                     unsigned long long val) {
#if FLT_DEVICE_BUILD
  return atomicCAS(p, cmp, val);
#else
  unsigned long long o = *p;
  if (o == cmp) *p = val;
  return o;
#endif
}
FLT_DEV unsigned long long atomMax64(unsigned long long* p, unsigned long long v) {
#if FLT_DEVICE_BUILD
  return atomicMax(p, v);
#else
  unsigned long long o = *p;
  if (v > o) *p = v;
  return o;
#endif
}

/* ---- volatile-free CTA broadcast helpers are plain shared/global loads after cta.sync() ---- */

FLT_HD int nextPow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// 64-bit mixer (splitmix64 finaliser) used for the n-gram keys and hash-table slots.
FLT_HD uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

#if !FLT_DEVICE_BUILD
inline double flt_log1p(double x) { return std::log1p(x); }
inline double flt_exp(double x) { return std::exp(x); }
#else
FLT_DEV double flt_log1p(double x) { return log1p(x); }
FLT_DEV double flt_exp(double x) { return exp(x); }
#endif

} // namespace flt
