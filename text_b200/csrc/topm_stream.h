// topm_stream.h — the streaming form of the token-beam select (K1): the HBM-bound kernel of the path.
//
// Replaces `std::iota` + `std::partial_sort` over one emission row per frame
// (decoder/LexiconFreeDecoder.cpp:39-51, decoder/LexiconDecoder.cpp:41-52), like topm_core.h, for the
// shapes the benchmark configurations use (16-byte aligned rows, N % 4 == 0, up to ~380 entries wanted).
//
// One 128-thread CTA owns one shared-memory stage of N floats; four CTAs share an SM. Per row:
//   wait(rowFull: the TMA bulk copy of the row into the stage, cp.async.bulk + mbarrier complete_tx)
//   -> ONE pass over the stage against a running guess of the select bound (the previous row's exact
//      want-th value minus an adaptive margin), collecting the few elements that reach it
//   -> the stage is free: one thread issues the bulk copy of the CTA's NEXT row, which streams in
//      from HBM behind the rest of the select
//      (lexicon decoder: of the ranking keys e[n] + bias[n]; the bias row is read through L1 alongside)
//   -> the survivors are ranked exactly (128-bin histogram over [bound, maximum], suffix scan,
//      comparison inside bins) and the list is written.
// If at least `want` elements passed, the result is exact (every element >= the want-th largest was
// collected); else the row is selected again with an exact two-pass bound from L2, and after repeated
// misses (peaky rows whose level moves from frame to frame) the CTA switches to two passes over the
// stage for a spell of rows. Every row is read from HBM exactly once: 4N bytes in, 8M bytes out.
// Sixteen warps per SM keep four 40 KB bulk copies in flight per SM — 24 MB over the chip, several
// times the latency-bandwidth product of HBM3e — so the kernel runs at the memory's pace.
#pragma once
#include "fused_core.h"
#include "topm_core.h"

namespace flt {

constexpr int kStreamThreads = 128;
constexpr int kStreamSPT = 4;    // survivors a thread ranks: up to 512 per row
                                 // (survivor capacity TopMCfg::capS = kStreamSPT x threads: 512, or 1024 for long lists)

#if FLT_DEVICE_BUILD
// prodFilter (fused_core.h) on the ranking keys e[n] + bias[n] of the lexicon decoder: the bias row (4N bytes,
// the same for every row and every CTA) is read through L1 next to the staged emissions, so the one pass
// collects exactly the keys that reach `bound` — filtering raw emissions against bound - max(bias) lets almost
// the whole row through when the bias spreads wider than the emissions do (a smeared n-gram LM).
FLT_DEV float4 streamKeyed(const float4 x, const float4 b) {
  const float ninf = bitsF32(0xFF800000u);
  float4 k;
  k.x = isNegInf(b.x) ? ninf : x.x + b.x;
  k.y = isNegInf(b.y) ? ninf : x.y + b.y;
  k.z = isNegInf(b.z) ? ninf : x.z + b.z;
  k.w = isNegInf(b.w) ? ninf : x.w + b.w;
  return k;
}
FLT_DEV float prodFilterBiased(const Cta& p, const TopMCfg& c, TopMSmem& s, const float4* r4, const float4* b4,
                               float bound, unsigned long long* sv) {
  const int nvec = c.N >> 2;
  float top = bitsF32(0xFF800000u);
  for (int v0 = p.tid; v0 < nvec; v0 += 32 * p.nthr) {
    unsigned hits = 0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const int v = v0 + k * p.nthr;
      if (v < nvec) {
        const float4 x = streamKeyed(r4[v], __ldg(b4 + v));
        const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
        top = fmaxf(top, mx);
        hits |= (mx >= bound ? 1u : 0u) << k;
      }
    }
    while (hits) {
      const int k = __ffs(hits) - 1;
      hits &= hits - 1;
      const int v = v0 + k * p.nthr;
      const float4 x = streamKeyed(r4[v], __ldg(b4 + v));
      const float xs[4] = {x.x, x.y, x.z, x.w};
      const unsigned m4 = (x.x >= bound ? 1u : 0u) | (x.y >= bound ? 2u : 0u) | (x.z >= bound ? 4u : 0u) |
                          (x.w >= bound ? 8u : 0u);
      int pos = atomAdd(&s.cnt[0], __popc(m4));
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (m4 & (1u << q)) {
          if (pos < c.capS) sv[pos] = topmKey(xs[q], v * 4 + q);
          ++pos;
        }
    }
  }
  return top;
}
// survivors of the raw-emission filter -> ranking keys e + bias that reach `bound` (compacted in place);
// returns their number and this thread's largest key value. Used when the bias is (nearly) constant over the
// tokens that start a word (ZeroLM: all zero), where filtering raw emissions against bound - max(bias) is as
// tight as the keyed filter and saves reading the bias row.
FLT_DEV int streamRekey(const Cta& p, const TopMCfg& c, TopMSmem& s, int n1, float bound, float& top) {
  unsigned long long* sv = s.sortBuf + c.capS;
  unsigned long long mine[kStreamSPT];
  const float ninf = bitsF32(0xFF800000u);
  top = ninf;
#pragma unroll
  for (int z = 0; z < kStreamSPT; ++z) {
    const int a = p.tid + z * p.nthr;
    mine[z] = 0ull;
    if (a < n1) {
      const unsigned long long k = sv[a];
      const int tok = topmKeyTok(k);
      const float b = __ldg(c.bias + tok);
      if (!isNegInf(b)) {
        const float kv = topmKeyVal(k) + b;
        if (kv >= bound) {
          mine[z] = topmKey(kv, tok);
          top = fmaxf(top, kv);
        }
      }
    }
  }
  p.sync(); // every entry has been read
  if (p.tid == 0) s.cnt[0] = 0;
  p.sync();
#pragma unroll
  for (int z = 0; z < kStreamSPT; ++z)
    if (mine[z]) sv[atomAdd(&s.cnt[0], 1)] = mine[z];
  p.sync();
  return s.cnt[0];
}
#endif

// One row. `row` = the shared-memory stage, `grow` = the same row in global memory (L2).
template <class StageFree>
FLT_DEV void streamRow(const Cta& p, const TopMCfg& c, TopMSmem& s, ProdGuess& pg, const float* row,
                       const float* grow, int* outTok, float* outVal, float* outThr, StageFree stageFree) {
  const int N = c.N;
  const bool restricted = c.bst < N;
  const int want = restricted ? c.bst : c.M;
  const bool biased = c.bias != nullptr && !restricted;
#if FLT_DEVICE_BUILD
  {
    const int lane = p.tid & 31, warp = p.tid >> 5, nw = p.nthr >> 5;
    float* bnd = (float*)s.red;          // [nw] per-warp bound, [32 + nw] per-warp maximum
    int* hist = s.rankCnt;               // [kProdBins] zero on entry, re-zeroed below
    int* above = s.rankCnt + kProdBins;  // [kProdBins]
    unsigned long long* sv = s.sortBuf + c.capS;
    const int minExpected = biased ? 1 : (want < N ? want : N);
    const float ninf = bitsF32(0xFF800000u);
    auto okCount = [&](int n) { return n >= minExpected && n <= c.capS && n <= kStreamSPT * p.nthr; };
    const float4* b4 = (const float4*)c.bias;
    float bound = pg.g, top = ninf;
    int ns = 0;
    // exact bound from the per-thread maxima of the ranking keys of the row at r4 (pass 1), then the
    // filter (pass 2); with a bias both passes read it for the whole row (L2)
    auto keyed = [&](const float4* r4, int v, bool global) -> float4 {
      float4 x = global ? __ldg(r4 + v) : r4[v];
      if (biased) {
        const float4 b = __ldg(b4 + v);
        x.x = isNegInf(b.x) ? ninf : x.x + b.x;
        x.y = isNegInf(b.y) ? ninf : x.y + b.y;
        x.z = isNegInf(b.z) ? ninf : x.z + b.z;
        x.w = isNegInf(b.w) ? ninf : x.w + b.w;
      }
      return x;
    };
    auto exactSelect = [&](const float4* r4, bool global) {
      const int nvec = N >> 2;
      float m = ninf;
#pragma unroll 4
      for (int v = p.tid; v < nvec; v += p.nthr) {
        const float4 x = keyed(r4, v, global);
        m = fmaxf(m, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
      }
      const float sorted = warpSortDesc(m, lane);
      const int r = (want + nw - 1) / nw; // every warp certifies r elements >= its r-th largest maximum
      const float tw = r <= 32 ? __shfl_sync(0xffffffffu, sorted, r - 1) : ninf;
      if (lane == 0) bnd[warp] = tw;
      if (p.tid == 0) s.cnt[0] = 0;
      p.sync();
      float bd = bnd[0];
      for (int i = 1; i < nw; ++i) bd = fminf(bd, bnd[i]);
      bd = fmaxf(bd, bitsF32(0xFF7FFFFFu)); // -inf never passes
      p.sync(); // bnd[] is reused below
      float tp = ninf;
#pragma unroll 2
      for (int v = p.tid; v < nvec; v += p.nthr) {
        const float4 x = keyed(r4, v, global);
        const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
        tp = fmaxf(tp, mx);
        if (mx >= bd) {
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (xs[q] >= bd) {
              const int pos = atomAdd(&s.cnt[0], 1);
              if (pos < c.capS) sv[pos] = topmKey(xs[q], v * 4 + q);
            }
        }
      }
      p.sync();
      bound = bd;
      top = tp;
      ns = s.cnt[0];
    };
    // (the per-warp certificate of the exact bound covers 32 elements per warp: longer lists take the
    // generic select below when a guess misses)
    const bool canExact = want <= 32 * nw;
    bool generic = false;
    if (pg.exactRows > 0) {
      // ---- exact mode: both passes read the stage, which is released after the second
      --pg.exactRows;
      if (canExact) exactSelect((const float4*)row, false);
      else generic = true;
      stageFree();
    } else {
      // ---- guess mode: one pass over the stage against the running guess, stage released at once
      // c.biasKeyed: the bias spreads over the tokens (a smeared n-gram LM) -> filter the keys e + bias;
      // else (constant bias, ZeroLM) filter the raw emissions against bound - max(bias) and re-key the survivors
      const bool useKeyed = biased && c.biasKeyed;
      top = useKeyed ? prodFilterBiased(p, c, s, (const float4*)row, b4, bound, sv)
                  : prodFilter(p, c, s, (const float4*)row, biased ? bound - c.biasMax : bound, sv);
      p.sync(); // every thread is done reading the stage
      stageFree();
      ns = s.cnt[0];
      bool miss = ns > c.capS || ns > kStreamSPT * p.nthr;
      if (!miss && biased && !useKeyed) ns = streamRekey(p, c, s, ns, bound, top);
      miss = miss || !okCount(ns) || (biased && ns < (want < N ? want : N) && bound > bitsF32(0xFF7FFFFFu));
      pg.missRate = 0.9f * pg.missRate + (miss ? 0.1f : 0.0f);
      if (miss) {
        // too many survivors: guess closer to the want-th value next time; too few: further below it
        pg.margin = fminf(fmaxf(pg.margin * (ns > want ? 0.7f : 1.4f), 0.02f), 4.0f);
        p.sync(); // everyone has read cnt[0]
        if (canExact) {
          exactSelect((const float4*)grow, true); // exact two-pass select of this row from L2
        } else {
          // long lists: filter the row again from L2 with the bound moved towards the side that missed (a
          // bracket once both sides are known); exact as soon as want <= survivors <= capacity
          generic = true;
          const float lowest = bitsF32(0xFF7FFFFFu);
          const int wantN = want < N ? want : N;
          float bLo = ninf, bHi = bitsF32(0x7F800000u); // too many at bLo, too few at bHi
          auto ctaTop = [&](float t) { // row maximum, uniform over the CTA (prodFilter returns per-thread maxima)
            const unsigned tk = __reduce_max_sync(0xffffffffu, orderedKey32(t));
            if (lane == 0) bnd[32 + warp] = orderedKey32Inv(tk);
            p.sync();
            float m = bnd[32];
            for (int i = 1; i < nw; ++i) m = fmaxf(m, bnd[32 + i]);
            p.sync();
            return m;
          };
          top = ctaTop(top);
          for (int tries = 0; tries < 6 && generic; ++tries) {
            const bool tooFew = ns < wantN;
            float nb;
            if (!(bound < 3.0e38f)) { // no guess yet (first row of the CTA): start just below the maximum
              nb = top - 0.5f;
            } else {
              if (tooFew) bHi = bound;
              else bLo = bound;
              const float span = fmaxf(top - bound, 1e-3f);
              nb = tooFew ? bound - 1.5f * span : bound + 0.4f * span;
              if (!isNegInf(bLo) && bHi < 3.0e38f) nb = 0.5f * (bLo + bHi);
              if (tooFew && tries >= 4) nb = lowest; // (biased) fewer expandable tokens than wanted: take them all
              if (!(nb < bHi) || !(nb > bLo)) break;
            }
            if (!(nb > lowest)) nb = lowest;
            bound = nb;
            if (p.tid == 0) s.cnt[0] = 0;
            p.sync();
            top = biased ? prodFilterBiased(p, c, s, (const float4*)grow, b4, bound, sv)
                         : prodFilter(p, c, s, (const float4*)grow, bound, sv);
            p.sync();
            ns = s.cnt[0];
            top = ctaTop(top);
            generic = !okCount(ns) || (ns < wantN && (!biased || bound > lowest));
          }
        }
        if (pg.missRate > 0.2f && canExact) { // stop guessing for a while; the spell doubles each time (<= 4096 rows)
          pg.exactRows = pg.exactSpell;
          pg.exactSpell = pg.exactSpell < 4096 ? pg.exactSpell * 2 : 4096;
          pg.missRate = 0.0f;
        }
      }
    }
    if (!generic && okCount(ns)) {
      // ---- rank the survivors: linear histogram over [bound, row maximum], exact order inside bins
      {
        const unsigned tk = __reduce_max_sync(0xffffffffu, orderedKey32(top));
        if (lane == 0) bnd[32 + warp] = orderedKey32Inv(tk);
      }
      p.sync();
      top = bnd[32];
      for (int i = 1; i < nw; ++i) top = fmaxf(top, bnd[32 + i]);
      const float range = top - bound;
      const float scale = (range > 0.0f && range < 3.0e38f) ? (float)kProdBins / range : 0.0f;
      unsigned long long mine[kStreamSPT];
      int myBin[kStreamSPT], mySlot[kStreamSPT];
#pragma unroll
      for (int z = 0; z < kStreamSPT; ++z) {
        const int a = p.tid + z * p.nthr;
        myBin[z] = -1;
        mySlot[z] = 0;
        mine[z] = 0ull;
        if (a < ns) {
          mine[z] = sv[a];
          int bin = (int)((topmKeyVal(mine[z]) - bound) * scale);
          bin = bin > kProdBins - 1 ? kProdBins - 1 : (bin < 0 ? 0 : bin);
          myBin[z] = bin;
          mySlot[z] = atomAdd(&hist[bin], 1);
        }
      }
      p.sync();
      { // above[bin] = survivors in higher bins; every warp scans, warp 0 publishes
        const int4 h4 = *(const int4*)(hist + lane * 4);
        const int own = (h4.x + h4.y) + (h4.z + h4.w);
        int suf = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_down_sync(0xffffffffu, suf, o);
          if (lane + o < 32) suf += u;
        }
        int4 a4;
        int ab = suf - own;
        a4.w = ab;
        ab += h4.w;
        a4.z = ab;
        ab += h4.z;
        a4.y = ab;
        ab += h4.y;
        a4.x = ab;
        if (warp == 0) *(int4*)(above + lane * 4) = a4;
      }
      p.sync();
      // survivors grouped by bin (best bins first); only those that can rank < want are placed
#pragma unroll
      for (int z = 0; z < kStreamSPT; ++z)
        if (myBin[z] >= 0) {
          const int ab = above[myBin[z]];
          if (ab < want) s.sortBuf[ab + mySlot[z]] = mine[z];
          else myBin[z] = -1;
        }
      if (p.tid == 0) s.cnt[1] = (int)f32Bits(bound); // fewer than `want` survivors: the bound itself
      p.sync();
      const int nout = ns < want ? ns : want;
#pragma unroll
      for (int z = 0; z < kStreamSPT; ++z)
        if (myBin[z] >= 0) {
          const int ab = above[myBin[z]];
          const int cnt = hist[myBin[z]];
          int rr = ab;
          for (int k = 0; k < cnt; ++k) rr += s.sortBuf[ab + k] > mine[z] ? 1 : 0;
          if (rr < want) {
            const int tok = topmKeyTok(mine[z]);
            outTok[rr] = tok;
            outVal[rr] = biased ? __ldg(grow + tok) : topmKeyVal(mine[z]);
            if (rr == want - 1) {
              s.cnt[1] = (int)f32Bits(topmKeyVal(mine[z])); // the exact want-th value
              if (restricted) *outThr = topmKeyVal(mine[z]);
            }
          }
        }
      for (int j = nout + p.tid; j < c.M; j += p.nthr) { // short rows (biased: few expandable tokens)
        outTok[j] = -1;
        outVal[j] = 0.0f;
      }
      p.sync();
      for (int bn = p.tid; bn < kProdBins; bn += p.nthr) hist[bn] = 0;
      if (p.tid == 0) s.cnt[0] = 0;
      // next row's guess: below this row's want-th value by a margin that tracks the survivor count
      const float wth = bitsF32((uint32_t)s.cnt[1]);
      // (aim between 1.5 x and 2.5 x want survivors, and well inside the survivor capacity)
      const int hiT = 2 * want + want / 2 < (3 * c.capS) / 4 ? 2 * want + want / 2 : (3 * c.capS) / 4;
      const int loT = want + want / 2 < hiT - want / 4 ? want + want / 2 : hiT - want / 4;
      if (ns > hiT) pg.margin *= 0.85f;
      else if (ns < loT) pg.margin *= 1.25f;
      pg.margin = fminf(fmaxf(pg.margin, 0.02f), 4.0f);
      pg.g = wth - pg.margin * (top - wth);
      p.sync();
      return;
    }
    // adversarial row (ties en masse, -inf padding) or a long list in exact mode: generic path below
    p.sync(); // everyone has read cnt[0]
    for (int bn = p.tid; bn < kProdBins; bn += p.nthr) hist[bn] = 0;
    if (p.tid == 0) s.cnt[0] = 0;
    pg.g = bitsF32(0x7F800000u);
    p.sync();
  }
#else
  stageFree();
  (void)row;
  (void)pg;
#endif
  const float* bias = c.bias;
  if (biased) {
    topmSelect(p, c, s, N, want, [&](int i) {
      const float b = bias[i];
      return isNegInf(b) ? 0ull : topmKey(grow[i] + b, i);
    });
  } else {
    topmSelect(p, c, s, N, want, [&](int i) { return topmKey(grow[i], i); });
  }
  if (restricted && p.tid == 0) *outThr = topmKeyVal(s.sortBuf[c.bst - 1]);
  for (int j = p.tid; j < c.M; j += p.nthr) {
    const unsigned long long k = j < want ? s.sortBuf[j] : 0ull;
    const int tok = k ? topmKeyTok(k) : -1;
    outTok[j] = tok;
    outVal[j] = tok >= 0 ? (biased ? grow[tok] : topmKeyVal(k)) : 0.0f;
  }
  if (p.tid == 0) s.cnt[0] = 0; // the single-pass filter of the next row counts from zero
#if FLT_DEVICE_BUILD
  { // next row's guess from this row's exact result
    const unsigned long long kw = s.sortBuf[want - 1], k0 = s.sortBuf[0];
    if (kw && k0) {
      const float wth = topmKeyVal(kw), tp = topmKeyVal(k0);
      pg.g = wth - pg.margin * (tp - wth);
    }
  }
#endif
  p.sync();
}

struct StreamLay { // byte offsets from the CTA's shared-memory base
  int threads;     // 128 (lists up to ~340 entries) or 256
  int prod;        // scratch (TopMSmem)
  int row;         // staged emission row [N] fp32, 128-byte aligned
  int mbar;        // rowFull
  int total;
};

// One CTA streams rows bid, bid + nblk, ...
FLT_DEV void topmStreamCta(const Cta& p, const TopMCfg& tc, const StreamLay& sl, const TopMArgs& a, char* smem) {
  TopMSmem ps;
  carveTopM(smem + sl.prod, tc, ps);
  for (int i = p.tid; i < 2 * kProdBins + tc.capS; i += p.nthr) ps.rankCnt[i] = 0;
  if (p.tid < 4) ps.cnt[p.tid] = 0;
  float dummyThr = 0.0f;
  ProdGuess pg{bitsF32(0x7F800000u), 0.25f, 0.0f, 0, 64};
#if FLT_DEVICE_BUILD
  float* stage = (float*)(smem + sl.row);
  u64* full = (u64*)(smem + sl.mbar);
  const uint32_t rowBytes = (uint32_t)tc.N * 4u;
  if (p.tid == 0) {
    mbarInit(full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  p.sync();
  const u64 pol = l2EvictFirstPolicy();
  long long r = p.bid;
  if (r < a.rows && p.tid == 0) { // prime the stage
    mbarArriveExpectTx(full, rowBytes);
    bulkLoadHint(stage, a.emis + r * tc.N, rowBytes, full, pol);
  }
  uint32_t it = 0;
  for (; r < a.rows; r += p.nblk, ++it) {
    const long long next = r + p.nblk;
    mbarWait(full, it & 1);
    streamRow(p, tc, ps, pg, stage, a.emis + r * tc.N, a.outTok + r * tc.M, a.outVal + r * tc.M,
              a.outThr ? a.outThr + r : &dummyThr, [&]() {
                if (next < a.rows && p.tid == 0) { // the stage is free: stream the next row in behind the select
                  fenceProxyAsync();
                  mbarArriveExpectTx(full, rowBytes);
                  bulkLoadHint(stage, a.emis + next * tc.N, rowBytes, full, pol);
                }
              });
  }
#else
  p.sync();
  for (long long r = p.bid; r < a.rows; r += p.nblk)
    streamRow(p, tc, ps, pg, a.emis + r * tc.N, a.emis + r * tc.N, a.outTok + r * tc.M, a.outVal + r * tc.M,
              a.outThr ? a.outThr + r : &dummyThr, [] {});
  (void)sl;
#endif
}

} // namespace flt
