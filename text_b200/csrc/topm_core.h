// topm_core.h — K1, the per-frame token-beam select: the HBM-streaming kernel of the path.
//
// Replaces `std::iota` + `std::partial_sort` over one emission row per frame
// (decoder/LexiconFreeDecoder.cpp:39-51, decoder/LexiconDecoder.cpp:41-52). Every row of the
// [B*T, N] fp32 emission matrix is read from HBM exactly once, in coalesced 16-byte loads, staged
// in shared memory, and reduced to a short ranked list (token, value) — the only part of the row
// the beam kernel needs besides a handful of single-element gathers.
//
// Selection without sorting the row: the row is split into P interleaved chunks (P >= the number
// of entries wanted); the wanted-th largest chunk maximum is a lower bound of the wanted-th largest
// element, so filtering the row against it leaves ~wanted survivors, which are then sorted.
// Keys are (order-preserving fp32 bits << 32) | ~token, i.e. all distinct: value descending, then
// token ascending — a deterministic refinement of the reference's unspecified tie order.
//
// Modes
//   ranked by e[n]                     (lexicon-free decoder)
//   ranked by e[n] + bias[n]           (lexicon decoder, root rows: bias = lmWeight * smeared score
//                                       of the root child, -inf for tokens that are not expandable)
//   restricted to the beamSizeToken largest e[n] first when beamSizeToken < N.
#pragma once
#include "spmd.h"

namespace flt {

struct TopMCfg {
  int N;          // row length
  int M;          // list entries written per row
  int bst;        // token-set size (>= N: unrestricted)
  const float* bias; // [N] or null
  int P;          // chunks (pow2, >= nthr, >= wanted)
  int capS;       // survivor capacity (pow2)
  int stage;      // 1 = stage the row in shared memory
};

struct TopMArgs {
  const float* emis; // [rows, N]
  long long rows;
  int* outTok;    // [rows, M]   (-1 = no entry)
  float* outVal;  // [rows, M]   e[n] of the entry
  float* outThr;  // [rows] value of the bst-th largest e (null when unrestricted)
};

FLT_DEV unsigned long long topmKey(float v, int tok) {
  return ((unsigned long long)orderedKey32(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)tok);
}
FLT_DEV int topmKeyTok(unsigned long long k) { return (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull)); }
FLT_DEV float topmKeyVal(unsigned long long k) { return orderedKey32Inv((uint32_t)(k >> 32)); }

// descending bitonic sort of buf[0..P), P a power of two
FLT_DEV void ctaBitonicDesc(const Cta& cta, unsigned long long* buf, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = cta.tid; i < P; i += cta.nthr) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = buf[i], b = buf[l];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            buf[i] = b;
            buf[l] = a;
          }
        }
      }
      cta.sync();
    }
  }
}

constexpr int kTopMMaxQ = 8; // chunks per thread (P <= 8 * nthr)

struct TopMSmem {
  float* row;                  // [N] staged row (when cfg.stage)
  unsigned long long* sortBuf; // [max(P, capS)]
  int* cnt;                    // [4]
  unsigned long long* red;     // [64]
};

FLT_HD size_t carveTopM(char* base, const TopMCfg& c, TopMSmem& s) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off = (off + bytes + 15) / 16 * 16;
    return p;
  };
  const int nbuf = c.P > c.capS ? c.P : c.capS;
  s.sortBuf = (unsigned long long*)take(sizeof(unsigned long long) * nbuf);
  s.red = (unsigned long long*)take(sizeof(unsigned long long) * 64);
  s.cnt = (int*)take(sizeof(int) * 4);
  s.row = (float*)take(c.stage ? sizeof(float) * c.N : 0);
  return off;
}

FLT_DEV unsigned long long ctaMaxKey(const Cta& cta, unsigned long long v, unsigned long long* red) {
#if FLT_DEVICE_BUILD
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u > v ? u : v;
  }
  const int warp = cta.tid >> 5, lane = cta.tid & 31, nw = (cta.nthr + 31) >> 5;
  cta.sync();
  if (lane == 0) red[warp] = v;
  cta.sync();
  unsigned long long r = red[0];
  for (int i = 1; i < nw; ++i) r = red[i] > r ? red[i] : r;
  return r;
#else
  (void)cta;
  (void)red;
  return v;
#endif
}

// Select the `want` largest keys of {key(i) : i in [0,N), valid(i)} into sortBuf[0..want), sorted
// descending; entries beyond the number of valid elements are 0. keyOf(i) returns 0 for invalid.
template <class KeyOf>
FLT_DEV void topmSelect(const Cta& cta, const TopMCfg& c, TopMSmem& s, int N, int want, KeyOf keyOf) {
  unsigned long long* buf = s.sortBuf;
  if (N <= c.capS) { // small rows: sort everything
    const int P = nextPow2(N);
    for (int i = cta.tid; i < P; i += cta.nthr) buf[i] = i < N ? keyOf(i) : 0ull;
    cta.sync();
    ctaBitonicDesc(cta, buf, P);
    return;
  }
  // 1. chunk maxima: element i belongs to chunk i & (P-1)
  const int P = c.P;
#if !FLT_DEVICE_BUILD
  for (int i = 0; i < P; ++i) buf[i] = 0ull; // model: one thread owns every chunk
  for (int i = 0; i < N; ++i) {
    const unsigned long long key = keyOf(i);
    if (key > buf[i & (P - 1)]) buf[i & (P - 1)] = key;
  }
#else
  {
    unsigned long long best[kTopMMaxQ];
    const int q = P / cta.nthr; // chunks per thread (host guarantees 1 <= q <= kTopMMaxQ)
#pragma unroll
    for (int k = 0; k < kTopMMaxQ; ++k) best[k] = 0ull;
    int k = 0;
    for (int i = cta.tid; i < N; i += cta.nthr) {
      const unsigned long long key = keyOf(i);
#pragma unroll
      for (int z = 0; z < kTopMMaxQ; ++z)
        if (z == k) best[z] = key > best[z] ? key : best[z];
      k = (k + 1 == q) ? 0 : k + 1;
    }
#pragma unroll
    for (int z = 0; z < kTopMMaxQ; ++z)
      if (z < q) buf[cta.tid + z * cta.nthr] = best[z];
  }
#endif
  cta.sync();
  ctaBitonicDesc(cta, buf, P);
  const unsigned long long tau = buf[want - 1]; // lower bound of the want-th largest key
  cta.sync();
  // 2. filter the row against tau
  if (cta.tid == 0) s.cnt[0] = 0;
  cta.sync();
  for (int i = cta.tid; i < N; i += cta.nthr) {
    const unsigned long long key = keyOf(i);
    if (key >= tau && key != 0ull) {
      const int pos = atomAdd(&s.cnt[0], 1);
      if (pos < c.capS) buf[pos] = key;
    }
  }
  cta.sync();
  const int ns = s.cnt[0];
  cta.sync();
  if (ns <= c.capS) {
    const int P2 = nextPow2(ns > want ? ns : want);
    for (int i = ns + cta.tid; i < P2; i += cta.nthr) buf[i] = 0ull;
    cta.sync();
    ctaBitonicDesc(cta, buf, P2);
    return;
  }
  // 3. adversarial rows (more survivors than capS): extract the maxima one at a time
  unsigned long long last = ~0ull;
  for (int r = 0; r < want; ++r) {
    unsigned long long m = 0ull;
    for (int i = cta.tid; i < N; i += cta.nthr) {
      const unsigned long long key = keyOf(i);
      if (key < last && key > m) m = key;
    }
    m = ctaMaxKey(cta, m, s.red);
    if (cta.tid == 0) buf[r] = m;
    last = m ? m : 1ull;
    cta.sync();
  }
}

// One CTA handles rows bid, bid + nblk, ...
FLT_DEV void topmCta(const Cta& cta, const TopMCfg& c, const TopMArgs& a, char* smem) {
  TopMSmem s;
  carveTopM(smem, c, s);
  const int N = c.N;
  for (long long r = cta.bid; r < a.rows; r += cta.nblk) {
    const float* g = a.emis + r * N;
    const float* row = g;
    if (c.stage) {
      // coalesced 16-byte loads when the row is 16-byte aligned, scalar otherwise
      if ((((uintptr_t)g) & 15) == 0 && (N & 3) == 0) {
        const float4* g4 = (const float4*)g;
        float4* s4 = (float4*)s.row;
        for (int i = cta.tid; i < (N >> 2); i += cta.nthr) s4[i] = g4[i];
      } else {
        for (int i = cta.tid; i < N; i += cta.nthr) s.row[i] = g[i];
      }
      row = s.row;
      cta.sync();
    }
    int* ot = a.outTok + r * c.M;
    float* ov = a.outVal + r * c.M;
    const bool restricted = c.bst < N;
    if (!restricted) {
      if (!c.bias) {
        topmSelect(cta, c, s, N, c.M, [&](int i) { return topmKey(row[i], i); });
        for (int j = cta.tid; j < c.M; j += cta.nthr) {
          const unsigned long long k = s.sortBuf[j];
          ot[j] = k ? topmKeyTok(k) : -1;
          ov[j] = k ? topmKeyVal(k) : 0.0f;
        }
      } else {
        const float* bias = c.bias;
        topmSelect(cta, c, s, N, c.M, [&](int i) {
          const float b = bias[i];
          return isNegInf(b) ? 0ull : topmKey(row[i] + b, i);
        });
        for (int j = cta.tid; j < c.M; j += cta.nthr) {
          const unsigned long long k = s.sortBuf[j];
          const int tok = k ? topmKeyTok(k) : -1;
          ot[j] = tok;
          ov[j] = tok >= 0 ? row[tok] : 0.0f;
        }
      }
      cta.sync();
    } else {
      // token set = the bst largest e[n]; then (optionally) re-rank its eligible members by e + bias
      topmSelect(cta, c, s, N, c.bst, [&](int i) { return topmKey(row[i], i); });
      if (cta.tid == 0 && a.outThr) a.outThr[r] = topmKeyVal(s.sortBuf[c.bst - 1]);
      if (!c.bias) {
        for (int j = cta.tid; j < c.M; j += cta.nthr) {
          const unsigned long long k = j < c.bst ? s.sortBuf[j] : 0ull;
          ot[j] = k ? topmKeyTok(k) : -1;
          ov[j] = k ? topmKeyVal(k) : 0.0f;
        }
        cta.sync();
      } else {
        cta.sync();
        const int P = nextPow2(c.bst);
        for (int j = cta.tid; j < P; j += cta.nthr) {
          unsigned long long k = j < c.bst ? s.sortBuf[j] : 0ull;
          if (k) {
            const int tok = topmKeyTok(k);
            const float b = c.bias[tok];
            k = isNegInf(b) ? 0ull : topmKey(row[tok] + b, tok);
          }
          s.sortBuf[j] = k;
        }
        cta.sync();
        ctaBitonicDesc(cta, s.sortBuf, P);
        for (int j = cta.tid; j < c.M; j += cta.nthr) {
          const unsigned long long k = j < P ? s.sortBuf[j] : 0ull;
          const int tok = k ? topmKeyTok(k) : -1;
          ot[j] = tok;
          ov[j] = tok >= 0 ? row[tok] : 0.0f;
        }
        cta.sync();
      }
    }
  }
}

} // namespace flt
