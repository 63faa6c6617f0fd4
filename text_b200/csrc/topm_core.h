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
  float biasMax;     // largest finite bias (bound of the streaming kernel's raw-emission filter)
  int biasKeyed;     // streaming kernel: 1 = the one-pass filter runs on e + bias (bias spread wider than ~0.5),
                     // 0 = on raw emissions against bound - biasMax, survivors re-keyed
  int P;          // chunks (pow2, >= nthr, >= wanted)
  int capS;       // survivor capacity (pow2)
  int stage;      // 1 = stage the row in shared memory
  int fast;       // 1 = register-resident fast path applies (see fastSelect)
  int extra;      // extra ints behind rankCnt (scratch of the fused producer, fused_core.h)
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
  int* rankCnt;                // [capS] rank counters of the fast path (kept zero between rows)
};

FLT_HD size_t carveTopM(char* base, const TopMCfg& c, TopMSmem& s) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off = (off + bytes + 15) / 16 * 16;
    return p;
  };
  const int nbuf = 2 * (c.P > c.capS ? c.P : c.capS); // [0,capS) ranked, [capS,2capS) unordered
  s.sortBuf = (unsigned long long*)take(sizeof(unsigned long long) * nbuf);
  s.red = (unsigned long long*)take(sizeof(unsigned long long) * 64);
  s.cnt = (int*)take(sizeof(int) * 4);
  s.rankCnt = (int*)take(sizeof(int) * (c.capS + c.extra));
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

/* ------------------------------------------------------------------ fast path (device only) --- */
// Register-resident variant for the benchmark shapes: N <= 4*kFastVec*threads, 16-byte aligned
// rows, at most 256 entries wanted. Per row and thread: kFastVec 16-byte loads, one fp32 max per
// element, a 32-lane shuffle sort of the per-thread maxima, one fp32 compare per element, and a
// rank-by-counting of the ~1.5x`want` survivors. No per-element atomics, no block-wide sort.
constexpr int kFastVec = 10; // float4 per thread -> N <= 10240 at 256 threads

#if FLT_DEVICE_BUILD
FLT_DEV float warpSortDesc(float v, int lane) { // bitonic sort across the 32 lanes, descending
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = ((lane & k) == 0);        // this k-block sorts descending
      const bool lower = ((lane & j) == 0);     // lane holds the first element of the pair
      const float mx = fmaxf(v, o), mn = fminf(v, o);
      v = (up == lower) ? mx : mn;
    }
  }
  return v;
}

// Selects the `want` (<= 256) best of the row held in registers. keyv[] are the ranking values
// (fp32, -inf = invalid); on return s.sortBuf[0..ns) holds composite keys of the survivors ranked
// descending (ns >= min(want, #valid)), and *nsOut = ns. Returns false if the survivors overflowed
// capS (caller falls back to the generic path).
FLT_DEV bool fastSelect(const Cta& cta, const TopMCfg& c, TopMSmem& s, const float (&keyv)[4 * kFastVec],
                        int want, int minExpected, int* nsOut) {
  const int lane = cta.tid & 31, warp = cta.tid >> 5, nw = cta.nthr >> 5;
  const float ninf = bitsF32(0xFF800000u);
  float m = ninf;
#pragma unroll
  for (int z = 0; z < 4 * kFastVec; ++z) m = fmaxf(m, keyv[z]);
  const float sorted = warpSortDesc(m, lane);
  const int r = (want + nw - 1) / nw; // every warp certifies r elements >= its r-th largest maximum
  const float tw = __shfl_sync(0xffffffffu, sorted, r - 1);
  float* tauS = (float*)s.red;
  if (lane == 0) tauS[warp] = tw;
  if (cta.tid == 0) s.cnt[0] = 0;
  cta.sync();
  float tau = tauS[0];
  for (int i = 1; i < nw; ++i) tau = fminf(tau, tauS[i]);
  // -inf (padding, ineligible tokens) never passes: clamp the bound to the lowest finite float
  tau = fmaxf(tau, bitsF32(0xFF7FFFFFu));
  // count, warp-scan, one atomic per warp, then write
  int cntMine = 0;
#pragma unroll
  for (int z = 0; z < 4 * kFastVec; ++z) cntMine += keyv[z] >= tau ? 1 : 0;
  int incl = cntMine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  int base = 0;
  if (lane == 31) base = atomAdd(&s.cnt[0], incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  int pos = base + incl - cntMine;
  unsigned long long* src = s.sortBuf + c.capS;
  if (cntMine) {
#pragma unroll
    for (int z = 0; z < 4 * kFastVec; ++z) {
      if (keyv[z] >= tau) {
        // element index of register z: vector it = z/4 at float4 index it*nthr + tid
        const int idx = ((z >> 2) * cta.nthr + cta.tid) * 4 + (z & 3);
        if (pos < c.capS) src[pos] = topmKey(keyv[z], idx);
        ++pos;
      }
    }
  }
  cta.sync();
  const int ns = s.cnt[0];
  // overflow, or a row that needs -inf entries to fill the list: generic path
  if (ns > c.capS || ns < minExpected) return false;
  // rank by counting, spread over the CTA: thread (a, part) counts its slice of the survivors
  int* rankCnt = s.rankCnt; // zero on entry, re-zeroed below
  if (ns > 0) {
    const int parts = ns >= cta.nthr ? 1 : cta.nthr / ns;
    const int slice = (ns + parts - 1) / parts;
    for (int t = cta.tid; t < ns * parts; t += cta.nthr) {
      const int a = t % ns, part = t / ns;
      const int lo = part * slice, hi = lo + slice < ns ? lo + slice : ns;
      const unsigned long long ka = src[a];
      int cnt = 0;
      for (int b = lo; b < hi; ++b) cnt += src[b] > ka ? 1 : 0;
      if (cnt) atomAdd(&rankCnt[a], cnt);
    }
  }
  cta.sync();
  for (int a = cta.tid; a < ns; a += cta.nthr) {
    s.sortBuf[rankCnt[a]] = src[a];
    rankCnt[a] = 0;
  }
  cta.sync();
  *nsOut = ns;
  return true;
}
#endif

#if FLT_DEVICE_BUILD
// One row through the fast path. Returns false (nothing written) if the row must take the generic
// path (survivor overflow on adversarial data).
FLT_DEV bool topmRowFast(const Cta& cta, const TopMCfg& c, const TopMArgs& a, TopMSmem& s,
                         const float* g, long long r) {
  const int N = c.N, nvec = N >> 2;
  const float ninf = bitsF32(0xFF800000u);
  float keyv[4 * kFastVec];
  const float4* g4 = (const float4*)g;
#pragma unroll
  for (int it = 0; it < kFastVec; ++it) {
    const int v = it * cta.nthr + cta.tid;
    float4 x = make_float4(ninf, ninf, ninf, ninf);
    if (v < nvec) x = __ldcs(g4 + v); // streaming: each row is read exactly once
    keyv[4 * it + 0] = x.x;
    keyv[4 * it + 1] = x.y;
    keyv[4 * it + 2] = x.z;
    keyv[4 * it + 3] = x.w;
  }
  const bool restricted = c.bst < N;
  if (c.bias && !restricted) {
    const float4* b4 = (const float4*)c.bias;
#pragma unroll
    for (int it = 0; it < kFastVec; ++it) {
      const int v = it * cta.nthr + cta.tid;
      if (v < nvec) {
        const float4 b = __ldg(b4 + v);
        keyv[4 * it + 0] = isNegInf(b.x) ? ninf : keyv[4 * it + 0] + b.x;
        keyv[4 * it + 1] = isNegInf(b.y) ? ninf : keyv[4 * it + 1] + b.y;
        keyv[4 * it + 2] = isNegInf(b.z) ? ninf : keyv[4 * it + 2] + b.z;
        keyv[4 * it + 3] = isNegInf(b.w) ? ninf : keyv[4 * it + 3] + b.w;
      }
    }
  }
  int ns = 0;
  const int want = restricted ? c.bst : c.M;
  // raw emissions: fewer survivors than min(want, N) means -inf values are needed -> generic path
  const bool biased = c.bias && !restricted;
  const int minExpected = biased ? 0 : (want < N ? want : N);
  if (!fastSelect(cta, c, s, keyv, want, minExpected, &ns)) return false;
  int* ot = a.outTok + r * c.M;
  float* ov = a.outVal + r * c.M;
  if (!restricted) {
    for (int j = cta.tid; j < c.M; j += cta.nthr) {
      if (j < ns) {
        const unsigned long long k = s.sortBuf[j];
        const int tok = topmKeyTok(k);
        ot[j] = tok;
        ov[j] = c.bias ? g[tok] : topmKeyVal(k);
      } else {
        ot[j] = -1;
        ov[j] = 0.0f;
      }
    }
    cta.sync();
    return true;
  }
  // token set = the bst largest e[n]
  const int nset = ns < c.bst ? ns : c.bst;
  if (cta.tid == 0 && a.outThr) a.outThr[r] = nset > 0 ? topmKeyVal(s.sortBuf[nset - 1]) : ninf;
  if (!c.bias) {
    for (int j = cta.tid; j < c.M; j += cta.nthr) {
      const bool ok = j < nset;
      const unsigned long long k = ok ? s.sortBuf[j] : 0ull;
      ot[j] = ok ? topmKeyTok(k) : -1;
      ov[j] = ok ? topmKeyVal(k) : 0.0f;
    }
    cta.sync();
    return true;
  }
  // re-rank the eligible members of the set by e + bias
  unsigned long long* tmp = s.sortBuf + c.capS;
  for (int j = cta.tid; j < nset; j += cta.nthr) {
    const int tok = topmKeyTok(s.sortBuf[j]);
    const float b = c.bias[tok];
    tmp[j] = isNegInf(b) ? 0ull : topmKey(topmKeyVal(s.sortBuf[j]) + b, tok);
  }
  cta.sync();
  for (int j = cta.tid; j < c.M; j += cta.nthr) {
    ot[j] = -1;
    ov[j] = 0.0f;
  }
  cta.sync();
  for (int x = cta.tid; x < nset; x += cta.nthr) {
    const unsigned long long ka = tmp[x];
    if (!ka) continue;
    int rank = 0;
    for (int y = 0; y < nset; ++y) rank += tmp[y] > ka ? 1 : 0;
    if (rank < c.M) {
      const int tok = topmKeyTok(ka);
      ot[rank] = tok;
      ov[rank] = g[tok];
    }
  }
  cta.sync();
  return true;
}
#endif

// One CTA handles rows bid, bid + nblk, ...
FLT_DEV void topmCta(const Cta& cta, const TopMCfg& c, const TopMArgs& a, char* smem) {
  TopMSmem s;
  carveTopM(smem, c, s);
  const int N = c.N;
  for (int i = cta.tid; i < c.capS; i += cta.nthr) s.rankCnt[i] = 0;
  cta.sync();
  for (long long r = cta.bid; r < a.rows; r += cta.nblk) {
    const float* g = a.emis + r * N;
    const float* row = g;
#if FLT_DEVICE_BUILD
    if (c.fast && topmRowFast(cta, c, a, s, g, r)) continue;
#endif
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
