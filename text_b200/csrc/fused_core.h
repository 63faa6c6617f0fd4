// fused_core.h — token-beam select and beam step in ONE persistent kernel (lexicon-free decoder).
//
// One CTA per utterance: 8 consumer warps run the serial frame loop (beam_lf.h), 4 producer warps
// run ahead of them over the emission rows:
//
//   producer, row r:   wait(rowFull: the TMA bulk copy of row r into the shared-memory stage)
//                      -> pass 1 over the stage (per-thread maxima) -> the stage is free: one thread
//                      issues cp.async.bulk for row r+1, which streams in behind the rest of the
//                      select -> bound -> pass 2 (filter) re-reads row r from L2 -> survivors are
//                      ranked with a small histogram -> wait(listFree[r&1]) -> write list slot r&1
//                      -> arrive(listReady[r&1])
//   consumer, frame t: wait(listReady[t&1]) -> frame step with list t (its handful of per-hypothesis
//                      emission gathers hit L2: the row was just streamed) -> arrive(listFree[t&1])
//
// so that every emission row is read from HBM exactly once (4N bytes per frame, the §8d figure), no
// token list ever goes through HBM, and the bandwidth-bound select is hidden behind the
// latency-bound step: the producers run up to four rows ahead of the consumers.
//
// Replaces decoder/LexiconFreeDecoder.cpp:39-51 (partial_sort per frame) and :53-125 (decodeStep)
// together; same results as the two-kernel path (flt_k_topm + flt_k_decode), which remains for
// shapes the stage cannot hold.
#pragma once
#include "beam_core.h"
#include "beam_lf.h"
#include "beam_gx.h"
#include "topm_core.h"

namespace flt {

#ifndef FLT_FUSED_CONSUMERS
#define FLT_FUSED_CONSUMERS 256
#endif
constexpr int kFusedConsumers = FLT_FUSED_CONSUMERS; // threads (the first warps of the CTA)
constexpr int kFusedProducers = 128; // threads (the last 4 warps)
constexpr int kFusedRing = 4;        // token lists the producers may run ahead of the consumers (pow2)
constexpr int kFusedRingLog = 2;

struct FuseLay {       // byte offsets from the CTA's shared-memory base
  int ws;              // consumer workspace (DecCfg::lay)
  int prod;            // producer scratch (TopMSmem)
  int row;             // staged emission row [N] fp32, 16-byte aligned
  int list[kFusedRing]; // token list ring: int tok[M], float val[M]
  int thr[kFusedRing];  // cut value of the token set per ring slot
  int spec[kFusedRing]; // float2 per ring slot: e[blank], e[sil] of the row (beam_gx.h)
  int mbar;            // mbarriers: rowFull, listReady[ring], listFree[ring]
  int total;
};

enum { MB_ROW_FULL = 0, MB_LIST_READY0 = 1, MB_LIST_FREE0 = 1 + kFusedRing, MB_COUNT = 1 + 2 * kFusedRing };

/* ------------------------------------------------------------------ mbarrier / bulk copy ------ */
#if FLT_DEVICE_BUILD
FLT_DEV uint32_t smemU32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FLT_DEV void mbarInit(u64* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemU32(b)), "r"(count) : "memory");
}
FLT_DEV void mbarArrive(u64* b) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smemU32(b)) : "memory");
}
FLT_DEV void mbarArriveExpectTx(u64* b, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smemU32(b)),
               "r"(bytes)
               : "memory");
}
// producer-side wait: the producers run ahead and mostly wait for the consumers; back off so the
// polling does not take issue slots from the latency-critical consumer warps
FLT_DEV void mbarWaitRelaxed(u64* b, uint32_t parity) {
  const uint32_t addr = smemU32(b);
  for (;;) {
    uint32_t done = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(128);
  }
}
FLT_DEV void mbarWait(u64* b, uint32_t parity) {
  const uint32_t addr = smemU32(b);
  uint32_t done = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D bulk async copy global -> shared (TMA engine), completion counted in bytes on `b`
FLT_DEV void bulkLoad(void* dst, const void* src, uint32_t bytes, u64* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemU32(dst)),
               "l"(src), "r"(bytes), "r"(smemU32(b))
               : "memory");
}
// same with an L2 evict-first policy: emission rows are read once, the back-pointer history the
// step writes (and the backtrace re-reads) should be what stays in the 126 MB L2
FLT_DEV u64 l2EvictFirstPolicy() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
FLT_DEV void bulkLoadHint(void* dst, const void* src, uint32_t bytes, u64* b, u64 pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smemU32(dst)),
      "l"(src), "r"(bytes), "r"(smemU32(b)), "l"(pol)
      : "memory");
}
FLT_DEV void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

/* ------------------------------------------------------------------ producer: select ---------- */
constexpr int kProdBins = 128; // histogram bins of the survivor ranking (4 per lane)
constexpr int kProdCap = 512;  // survivor capacity (TopMCfg::capS of the fused producer)

// Running guess of the select bound, carried from row to row by the producer (uniform over its
// threads). A row is filtered in ONE pass over the staged copy against `g`; if at least `want`
// elements pass, the result is exact (every element >= the want-th largest was collected). The
// guess follows the previous row's exact want-th value minus a margin that adapts to how many
// elements passed. A miss (too few / too many survivors) costs one exact two-pass select from L2.
struct ProdGuess {
  float g;        // bound to try on the next row; +inf = none yet
  float margin;   // fraction of (row maximum - want-th value) the guess sits below the want-th value
  float missRate; // running rate of guess misses
  int exactRows;  // rows left in exact mode (two passes over the stage, no guess)
  int exactSpell; // length of the next exact-mode spell
};

#if FLT_DEVICE_BUILD
// collect every element >= bound of the row at r4 (shared or global) into sv[] (unordered keys);
// returns this thread's maximum. The scan itself is branch-free: a thread only notes WHICH of its
// 16-byte vectors hold a survivor (one bit each) and comes back for those few afterwards — a warp
// whose 32 lanes each test 4 elements would otherwise run the (rare per lane, common per warp)
// survivor path on most iterations.
FLT_DEV float prodFilter(const Cta& p, const TopMCfg& c, TopMSmem& s, const float4* r4, float bound,
                         unsigned long long* sv) {
  const int nvec = c.N >> 2;
  float top = bitsF32(0xFF800000u);
  for (int v0 = p.tid; v0 < nvec; v0 += 32 * p.nthr) {
    unsigned hits = 0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const int v = v0 + k * p.nthr;
      if (v < nvec) {
        const float4 x = r4[v];
        const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
        top = fmaxf(top, mx);
        hits |= (mx >= bound ? 1u : 0u) << k;
      }
    }
    while (hits) {
      const int k = __ffs(hits) - 1;
      hits &= hits - 1;
      const int v = v0 + k * p.nthr;
      const float4 x = r4[v];
      const unsigned m4 = (x.x >= bound ? 1u : 0u) | (x.y >= bound ? 2u : 0u) | (x.z >= bound ? 4u : 0u) |
                          (x.w >= bound ? 8u : 0u);
      int pos = atomAdd(&s.cnt[0], __popc(m4));
      if (m4 & 1u) {
        if (pos < c.capS) sv[pos] = topmKey(x.x, v * 4 + 0);
        ++pos;
      }
      if (m4 & 2u) {
        if (pos < c.capS) sv[pos] = topmKey(x.y, v * 4 + 1);
        ++pos;
      }
      if (m4 & 4u) {
        if (pos < c.capS) sv[pos] = topmKey(x.z, v * 4 + 2);
        ++pos;
      }
      if (m4 & 8u) {
        if (pos < c.capS) sv[pos] = topmKey(x.w, v * 4 + 3);
      }
    }
  }
  return top;
}
#endif

// Ranked list of one row: M entries (token, value) by value descending (ties: lower token), -1 / 0
// past the number of valid entries; *outThr = value of the beamSizeToken-th largest when the token
// set is restricted. No bias (lexicon-free decoder).
//   `row` is the shared-memory stage (TMA-filled), `grow` the same row in global memory (L2).
//   stageFree() is called as soon as every thread is done with the stage, so that the next row's
//   bulk copy overlaps the rest of the select; beforeWrite() is called before the outputs are
//   written (the caller waits there for the ring slot to be free).
template <class StageFree, class BeforeWrite>
FLT_DEV void topmRowStaged(const Cta& p, const TopMCfg& c, TopMSmem& s, ProdGuess& pg,
                           unsigned long long* stats, const float* row,
                           const float* grow, int* outTok, float* outVal, float* outThr,
                           StageFree stageFree, BeforeWrite beforeWrite) {
  const int N = c.N;
  const bool restricted = c.bst < N;
  const int want = restricted ? c.bst : c.M;
  bool done = false;
#if FLT_DEVICE_BUILD
  if (c.fast) {
    const int lane = p.tid & 31, warp = p.tid >> 5, nw = p.nthr >> 5;
    // scratch: sortBuf [0,capS) survivors grouped by bin | [capS,2capS) survivors as collected;
    // red: per-warp bounds / maxima; rankCnt: histogram [kProdBins], above [kProdBins]; a survivor's
    // (bin, slot) stays in its thread's registers
    float* bnd = (float*)s.red;          // [nw] per-warp bound, [32 + nw] per-warp maximum
    int* hist = s.rankCnt;               // [kProdBins] zero on entry, re-zeroed below
    int* above = s.rankCnt + kProdBins;  // [kProdBins]
    unsigned long long* sv = s.sortBuf + c.capS;
    const int minExpected = want < N ? want : N;
    const float ninf = bitsF32(0xFF800000u);
    auto okCount = [&](int n) { return n >= minExpected && n <= c.capS && n <= 2 * p.nthr; };
    float bound = pg.g, top;
    int ns;
    // exact bound from the per-thread maxima of the row at r4 (pass 1), then the filter (pass 2)
    auto exactSelect = [&](const float4* r4, bool global) {
      const int nvec = N >> 2;
      float m = ninf;
#pragma unroll 4
      for (int v = p.tid; v < nvec; v += p.nthr) {
        const float4 x = global ? __ldg(r4 + v) : r4[v];
        m = fmaxf(m, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
      }
      const float sorted = warpSortDesc(m, lane);
      const int r = (want + nw - 1) / nw; // every warp certifies r elements >= its r-th largest maximum
      const float tw = __shfl_sync(0xffffffffu, sorted, r - 1);
      if (lane == 0) bnd[warp] = tw;
      if (p.tid == 0) s.cnt[0] = 0;
      p.sync();
      float bd = bnd[0];
      for (int i = 1; i < nw; ++i) bd = fminf(bd, bnd[i]);
      bd = fmaxf(bd, bitsF32(0xFF7FFFFFu)); // -inf never passes
      p.sync(); // bnd[] is reused below
      const float tp = prodFilter(p, c, s, r4, bd, sv);
      p.sync();
      bound = bd;
      return tp;
    };
    if (pg.exactRows > 0) {
      // ---- exact mode (the guesses kept missing: peaky rows whose level moves from frame to frame):
      // both passes read the stage, which is released after the second
      --pg.exactRows;
      top = exactSelect((const float4*)row, false);
      stageFree();
      ns = s.cnt[0];
    } else {
      // ---- guess mode: one pass over the stage against the running guess, stage released at once
      top = prodFilter(p, c, s, (const float4*)row, bound, sv);
      p.sync(); // every thread is done reading the stage
      stageFree();
      ns = s.cnt[0];
      const bool miss = !okCount(ns);
      pg.missRate = 0.9f * pg.missRate + (miss ? 0.1f : 0.0f);
      if (miss) {
        // exact two-pass select of this row from L2
        if (stats && p.tid == 0) atomicAdd(stats + 11, 1ull);
        p.sync(); // everyone has read cnt[0]
        top = exactSelect((const float4*)grow, true);
        ns = s.cnt[0];
        if (pg.missRate > 0.2f) { // stop guessing for a while; the spell doubles each time (<= 4096 rows)
          pg.exactRows = pg.exactSpell;
          pg.exactSpell = pg.exactSpell < 4096 ? pg.exactSpell * 2 : 4096;
          pg.missRate = 0.0f;
        }
      }
    }
    if (okCount(ns)) {
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
      unsigned long long mine[2];
      int myBin[2], mySlot[2];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        const int a = p.tid + z * p.nthr;
        myBin[z] = -1;
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
      for (int z = 0; z < 2; ++z)
        if (myBin[z] >= 0) {
          const int ab = above[myBin[z]];
          if (ab < want) s.sortBuf[ab + mySlot[z]] = mine[z];
          else myBin[z] = -1;
        }
      p.sync();
      beforeWrite();
#pragma unroll
      for (int z = 0; z < 2; ++z)
        if (myBin[z] >= 0) {
          const int ab = above[myBin[z]];
          const int cnt = hist[myBin[z]];
          int rr = ab;
          for (int k = 0; k < cnt; ++k) rr += s.sortBuf[ab + k] > mine[z] ? 1 : 0;
          if (rr < want) {
            outTok[rr] = topmKeyTok(mine[z]);
            outVal[rr] = topmKeyVal(mine[z]);
            if (rr == want - 1) {
              s.cnt[1] = (int)f32Bits(topmKeyVal(mine[z])); // the exact want-th value
              if (restricted) *outThr = topmKeyVal(mine[z]);
            }
          }
        }
      p.sync();
      for (int bn = p.tid; bn < kProdBins; bn += p.nthr) hist[bn] = 0;
      if (p.tid == 0) s.cnt[0] = 0;
      // next row's guess: below this row's want-th value by a margin that tracks the survivor count
      const float wth = bitsF32((uint32_t)s.cnt[1]);
      if (ns > 2 * want + want / 2) pg.margin *= 0.85f;
      else if (ns < want + want / 2) pg.margin *= 1.25f;
      pg.margin = fminf(fmaxf(pg.margin, 0.02f), 4.0f);
      pg.g = wth - pg.margin * (top - wth);
      p.sync();
      return;
    }
    // adversarial row (ties en masse, -inf padding): generic path below
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
  (void)stats;
#endif
  if (!done) topmSelect(p, c, s, N, want, [&](int i) { return topmKey(grow[i], i); });
  beforeWrite();
  if (restricted && p.tid == 0) *outThr = topmKeyVal(s.sortBuf[c.bst - 1]);
  for (int j = p.tid; j < c.M; j += p.nthr) {
    const unsigned long long k = j < want ? s.sortBuf[j] : 0ull;
    outTok[j] = k ? topmKeyTok(k) : -1;
    outVal[j] = k ? topmKeyVal(k) : 0.0f;
  }
  if (p.tid == 0) s.cnt[0] = 0; // the single-pass filter of the next row counts from zero
  p.sync();
}

/* ------------------------------------------------------------------ the fused CTA ------------- */
struct FusedView {
  char* smem;
  const FuseLay* fl;
  FLT_DEV float* row() const { return (float*)(smem + fl->row); }
  FLT_DEV int* listTok(int k) const { return (int*)(smem + fl->list[k]); }
  FLT_DEV float* listVal(int k, int M) const { return (float*)(smem + fl->list[k]) + M; }
  FLT_DEV float* thr(int k) const { return (float*)(smem + fl->thr[k]); }
  FLT_DEV float* spec(int k) const { return (float*)(smem + fl->spec[k]); }
  FLT_DEV u64* mbar(int k) const { return (u64*)(smem + fl->mbar) + k; }
};

FLT_DEV FrameIn fusedFrameIn(const DecCfg& c, const BatchArgs& a, const FusedView& v, int b, int t,
                             int slot) {
  FrameIn f;
  f.e = a.emis + ((long long)b * a.T + t) * c.N;
  f.topTok = v.listTok(slot);
  f.topVal = v.listVal(slot, c.M);
  f.listLen = c.M;
  f.thrVal = c.setAll ? 0.0f : *v.thr(slot);
  f.first = t == 0;
  f.listIsSet = 1;
  f.specReady = 0; // the per-hypothesis emissions come from L2 (the row was just streamed)
  f.eBlank = 0.0f;
  f.eSil = 0.0f;
  f.listInfo = nullptr;
#if FLT_DEVICE_BUILD
  if (c.gx) { // the producer read e[blank], e[sil] from the staged row
    f.specReady = 1;
    f.eBlank = v.spec(slot)[0];
    f.eSil = v.spec(slot)[1];
  }
#endif
  f.eNext = nullptr; // set by the caller
  const long long h = ((long long)b * (a.T + 2) + (t + 1)) * c.K;
  f.hParent = a.hParent + h;
  f.hTok = a.hTok + h;
  f.hWord = nullptr;
  f.hScore = nullptr;
  f.hCount = nullptr;
  f.hRow = t + 1;
  f.hSkip = ((t + 1) & (kCpRows - 1)) == 0 ? a.hSkip + ((long long)b * a.nCp + ((t + 1) >> kCpShift)) * c.K : nullptr;
  return f;
}

// `whole` describes the launched CTA (kFusedConsumers + kFusedProducers threads on the device, one
// thread in the host model, which runs the producer's work inline where the consumer waits for it).
FLT_DEV void fusedCta(const Cta& whole, const DecCfg& c, const TopMCfg& tc, const FuseLay& fl,
                      const BatchArgs& a, char* smem) {
  const FusedView v{smem, &fl};
  const Ws w = wsOf(smem + fl.ws, c, nullptr); // one region (lexicon-free step)
  TopMSmem ps;
  carveTopM(smem + fl.prod, tc, ps);
#if FLT_DEVICE_BUILD
  const uint32_t rowBytes = (uint32_t)c.N * 4u;
  if (whole.tid == 0) {
    for (int k = 0; k < MB_COUNT; ++k) mbarInit(v.mbar(k), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool isConsumer = whole.tid < kFusedConsumers;
  if (!isConsumer) {
    /* ---------------- producer warps ---------------- */
    const Cta p{whole.tid - kFusedConsumers, kFusedProducers, whole.bid, whole.nblk, 2};
    for (int i = p.tid; i < 2 * kProdBins + tc.capS; i += p.nthr) ps.rankCnt[i] = 0;
    if (p.tid < 4) ps.cnt[p.tid] = 0;
    p.sync();
    ProdGuess pg{bitsF32(0x7F800000u), 0.25f, 0.0f, 0, 64};
    const u64 pol = l2EvictFirstPolicy();
    // rows of this CTA in order: (b, t) for b = bid, bid + nblk, ...; `r` counts them
    uint32_t r = 0;
    int nb = whole.bid, nt = 0; // the next row to stage
    auto advance = [&](int& b, int& t) { // first row with t < len at or after (b, t); b >= B when none
      while (b < a.B) {
        const int len = a.lengths ? a.lengths[b] : a.T;
        if (t < len) return;
        b += whole.nblk;
        t = 0;
      }
    };
    advance(nb, nt);
    if (nb < a.B && p.tid == 0) { // prime the stage
      mbarArriveExpectTx(v.mbar(MB_ROW_FULL), rowBytes);
      bulkLoadHint(v.row(), a.emis + ((long long)nb * a.T + nt) * c.N, rowBytes, v.mbar(MB_ROW_FULL), pol);
    }
    while (nb < a.B) {
      const int b = nb, t = nt;
      nt = t + 1;
      advance(nb, nt); // (nb, nt) = the row after (b, t)
      const float* grow = a.emis + ((long long)b * a.T + t) * c.N;
      mbarWait(v.mbar(MB_ROW_FULL), r & 1);
      const int slot = (int)(r & (kFusedRing - 1));
      const bool hasNext = nb < a.B;
      const float* gnext = hasNext ? a.emis + ((long long)nb * a.T + nt) * c.N : nullptr;
      float rowBlank = 0.0f, rowSil = 0.0f;
      if (c.gx && p.tid == 0) { // from the stage, before it is released
        if (c.ctc) rowBlank = v.row()[c.blank];
        rowSil = v.row()[c.sil];
      }
      topmRowStaged(
          p, tc, ps, pg, a.stats, v.row(), grow, v.listTok(slot), v.listVal(slot, c.M), v.thr(slot),
          [&]() {
            if (hasNext && p.tid == 0) { // the stage is free: stream the next row in behind the select
              fenceProxyAsync();
              mbarArriveExpectTx(v.mbar(MB_ROW_FULL), rowBytes);
              bulkLoadHint(v.row(), gnext, rowBytes, v.mbar(MB_ROW_FULL), pol);
            }
          },
          [&]() {
            mbarWaitRelaxed(v.mbar(MB_LIST_FREE0 + slot), ((r >> kFusedRingLog) + 1) & 1);
            if (c.gx && p.tid == 0) {
              v.spec(slot)[0] = rowBlank;
              v.spec(slot)[1] = rowSil;
            }
          });
      if (p.tid == 0) mbarArrive(v.mbar(MB_LIST_READY0 + slot));
      ++r;
    }
    return;
  }
  /* ---------------- consumer warps ---------------- */
  const Cta cta{whole.tid, kFusedConsumers, whole.bid, whole.nblk, 1};
#else
  const Cta cta = whole;
  for (int i = 0; i < 2 * kProdBins + tc.capS; ++i) ps.rankCnt[i] = 0;
#endif
  if (c.gx) gxInitWorkspace(cta, c, w);
  else ctaInitWorkspace(cta, c, w, smem + fl.ws);
  uint32_t g = 0; // frames consumed so far by this CTA (same sequence as the producer's r)
  for (int b = whole.bid; b < a.B; b += whole.nblk) {
    const int len = a.lengths ? a.lengths[b] : a.T;
    int curIdx = 0;
    cta.sync(); // previous utterance fully retired
    if (cta.tid == 0) seedUtterance(c, w, a, b);
    LfCarry carry{0.0f, 0.0f, 0.0f, 0};
    GxCarry gcarry = gxCarryInit();
    cta.sync();
    if (c.gx) gxBeginUtterance(cta, c, w, 1);
    else {
      lfTabBuild(cta, w, 0, w.sc()[SC_NH]); // fingerprint table of the seed beam
      cta.sync();
    }
    int gxAlive = 1; // hypotheses in the current beam (beam_gx.h); 0 = the beam died, lists are still consumed
    for (int t = 0; t < len; ++t, ++g) {
      const int slot = (int)(g & (kFusedRing - 1)); // ring slot = running row count mod ring, on both sides
      LfPhaseClock oc; // time spent waiting for the producers
      oc.start(cta, a.stats);
#if FLT_DEVICE_BUILD
      mbarWait(v.mbar(MB_LIST_READY0 + slot), (g >> kFusedRingLog) & 1);
#else
      { // host model: the producer's work for row (b, t), inline
        const float* grow = a.emis + ((long long)b * a.T + t) * c.N;
        ProdGuess pg{0.0f, 0.0f, 0.0f, 0, 64};
        topmRowStaged(cta, tc, ps, pg, nullptr, grow, grow, v.listTok(slot), v.listVal(slot, c.M), v.thr(slot), [] {}, [] {});
      }
#endif
      oc.mark(5);
      FrameIn f = fusedFrameIn(c, a, v, b, t, slot);
      f.eNext = t + 1 < len ? f.e + c.N : nullptr;
      if (c.gx) {
        if (gxAlive) {
          gxAlive = gxFrameStep<false>(cta, c, w, curIdx, f, a.status + b, a.stats, gcarry);
          if (gxAlive) curIdx ^= 1;
        }
      } else {
        lfFrameStep(cta, c, w, curIdx, f, a.stats, carry); // ends with a barrier
        curIdx ^= 1;
      }
#if FLT_DEVICE_BUILD
      if (cta.tid == 0) mbarArrive(v.mbar(MB_LIST_FREE0 + slot));
#endif
    }
    int nFin = 0;
    if (c.gx) {
      if (gxAlive) {
        const FrameIn f = finishFrameIn(c, a, b, len);
        nFin = gxFinish<false>(cta, c, w, curIdx, f);
        curIdx ^= 1;
      }
    } else if (w.sc()[SC_NH] != 0) {
      const FrameIn f = finishFrameIn(c, a, b, len);
      lfFinish(cta, c, w, curIdx, f);
      curIdx ^= 1;
      nFin = w.sc()[SC_NH];
    }
    cta.sync();
    writeFinals(cta, c, w, a, b, curIdx, nFin);
  }
}

} // namespace flt
