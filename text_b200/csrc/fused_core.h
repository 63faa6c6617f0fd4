// fused_core.h — token-beam select and beam step in ONE persistent kernel (lexicon-free decoder).
//
// One CTA per utterance: 8 consumer warps run the serial frame loop (beam_lf.h), 4 producer warps
// run ahead of them over the emission rows:
//
//   producer, row r:   wait(rowFree)  -> one thread: cp.async.bulk (TMA, 1-D) row r HBM -> smem stage,
//                      completion on the rowFull mbarrier -> all producer warps: select the ranked
//                      token list of row r from the staged row (two passes over shared memory)
//                      -> wait(listFree[r&1]) -> write list slot r&1 -> arrive(listReady[r&1])
//   consumer, frame t: wait(listReady[t&1]) -> frame step with list t and the emissions gathered
//                      for its own tokens -> arrive(listFree[t&1]) -> wait(rowFull, row t+1)
//                      -> gather e[t+1][own tokens / blank / sil] of the NEW beam from the staged
//                      row -> arrive(rowFree)
//
// so that every emission row is read from HBM exactly once (4N bytes per frame, the §8d figure),
// the scattered per-hypothesis emission gathers become shared-memory reads, no token list ever
// goes through HBM, and the bandwidth-bound select is hidden behind the latency-bound step.
// While the consumers step frame t the producers load and select row t+1.
//
// Replaces decoder/LexiconFreeDecoder.cpp:39-51 (partial_sort per frame) and :53-125 (decodeStep)
// together; same results as the two-kernel path (flt_k_topm + flt_k_decode), which remains for
// shapes the stage cannot hold.
#pragma once
#include "beam_core.h"
#include "beam_lf.h"
#include "topm_core.h"

namespace flt {

constexpr int kFusedConsumers = 256; // threads (warps 0..7)
constexpr int kFusedProducers = 128; // threads (warps 8..11)

struct FuseLay {       // byte offsets from the CTA's shared-memory base
  int ws;              // consumer workspace (DecCfg::lay)
  int prod;            // producer scratch (TopMSmem)
  int row;             // staged emission row [N] fp32, 16-byte aligned
  int list[2];         // token list ring: int tok[M], float val[M]
  int thr[2];          // cut value of the token set per ring slot
  int mbar;            // 6 mbarriers: rowFull, rowFree, listReady[2], listFree[2]
  int total;
};

enum { MB_ROW_FULL = 0, MB_ROW_FREE, MB_LIST_READY0, MB_LIST_READY1, MB_LIST_FREE0, MB_LIST_FREE1, MB_COUNT };

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
FLT_DEV void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

/* ------------------------------------------------------------------ producer: select from smem - */
// Ranked list of one staged row: M entries (token, value) by value descending (ties: lower token),
// -1 / 0 past the number of valid entries; *outThr = value of the beamSizeToken-th largest when the
// token set is restricted. No bias (lexicon-free decoder).
FLT_DEV void topmRowStaged(const Cta& p, const TopMCfg& c, TopMSmem& s, const float* row, int* outTok,
                           float* outVal, float* outThr) {
  const int N = c.N;
  const bool restricted = c.bst < N;
  const int want = restricted ? c.bst : c.M;
  bool done = false;
#if FLT_DEVICE_BUILD
  if (c.fast) {
    // two passes over the staged row: per-thread maxima -> bound, then filter against the bound
    const int lane = p.tid & 31, warp = p.tid >> 5, nw = p.nthr >> 5;
    const int nvec = N >> 2;
    const float4* r4 = (const float4*)row;
    const float ninf = bitsF32(0xFF800000u);
    float m = ninf;
    for (int v = p.tid; v < nvec; v += p.nthr) {
      const float4 x = r4[v];
      m = fmaxf(m, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
    }
    const float sorted = warpSortDesc(m, lane);
    const int r = (want + nw - 1) / nw; // every warp certifies r elements >= its r-th largest maximum
    const float tw = __shfl_sync(0xffffffffu, sorted, r - 1);
    float* tauS = (float*)s.red;
    if (lane == 0) tauS[warp] = tw;
    if (p.tid == 0) s.cnt[0] = 0;
    p.sync();
    float tau = tauS[0];
    for (int i = 1; i < nw; ++i) tau = fminf(tau, tauS[i]);
    tau = fmaxf(tau, bitsF32(0xFF7FFFFFu)); // -inf never passes
    int cntMine = 0;
    for (int v = p.tid; v < nvec; v += p.nthr) {
      const float4 x = r4[v];
      cntMine += (x.x >= tau ? 1 : 0) + (x.y >= tau ? 1 : 0) + (x.z >= tau ? 1 : 0) + (x.w >= tau ? 1 : 0);
    }
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
      for (int v = p.tid; v < nvec; v += p.nthr) {
        const float4 x = r4[v];
        const float e4[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          if (e4[z] >= tau) {
            if (pos < c.capS) src[pos] = topmKey(e4[z], v * 4 + z);
            ++pos;
          }
        }
      }
    }
    p.sync();
    const int ns = s.cnt[0];
    const int minExpected = want < N ? want : N;
    if (ns <= c.capS && ns >= minExpected) {
      int* rankCnt = s.rankCnt; // zero on entry, re-zeroed below
      const int parts = ns >= p.nthr ? 1 : p.nthr / ns;
      const int slice = (ns + parts - 1) / parts;
      for (int t = p.tid; t < ns * parts; t += p.nthr) {
        const int a = t % ns, part = t / ns;
        const int lo = part * slice, hi = lo + slice < ns ? lo + slice : ns;
        const unsigned long long ka = src[a];
        int cnt = 0;
        for (int b = lo; b < hi; ++b) cnt += src[b] > ka ? 1 : 0;
        if (cnt) atomAdd(&rankCnt[a], cnt);
      }
      p.sync();
      for (int a = p.tid; a < ns; a += p.nthr) {
        s.sortBuf[rankCnt[a]] = src[a];
        rankCnt[a] = 0;
      }
      for (int a = ns + p.tid; a < want; a += p.nthr) s.sortBuf[a] = 0ull;
      p.sync();
      done = true;
    } else {
      p.sync(); // everyone has read cnt[0] before the generic path reuses it
    }
  }
#endif
  if (!done) topmSelect(p, c, s, N, want, [&](int i) { return topmKey(row[i], i); });
  if (restricted && p.tid == 0) *outThr = topmKeyVal(s.sortBuf[c.bst - 1]);
  for (int j = p.tid; j < c.M; j += p.nthr) {
    const unsigned long long k = j < want ? s.sortBuf[j] : 0ull;
    outTok[j] = k ? topmKeyTok(k) : -1;
    outVal[j] = k ? topmKeyVal(k) : 0.0f;
  }
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
  f.specReady = 1;
  const long long h = ((long long)b * (a.T + 2) + (t + 1)) * c.K;
  f.hParent = a.hParent + h;
  f.hTok = a.hTok + h;
  f.hWord = nullptr;
  return f;
}

// `whole` describes the launched CTA (kFusedConsumers + kFusedProducers threads on the device, one
// thread in the host model, which runs the producer's work inline where the consumer waits for it).
FLT_DEV void fusedCta(const Cta& whole, const DecCfg& c, const TopMCfg& tc, const FuseLay& fl,
                      const BatchArgs& a, char* smem) {
  const FusedView v{smem, &fl};
  const Ws w{smem + fl.ws, &c};
  TopMSmem ps;
  carveTopM(smem + fl.prod, tc, ps);
  const uint32_t rowBytes = (uint32_t)c.N * 4u;
#if FLT_DEVICE_BUILD
  if (whole.tid == 0) {
    for (int k = 0; k < MB_COUNT; ++k) mbarInit(v.mbar(k), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool isConsumer = whole.tid < kFusedConsumers;
  if (!isConsumer) {
    /* ---------------- producer warps ---------------- */
    const Cta p{whole.tid - kFusedConsumers, kFusedProducers, whole.bid, whole.nblk, 2};
    for (int i = p.tid; i < tc.capS; i += p.nthr) ps.rankCnt[i] = 0;
    p.sync();
    uint32_t r = 0; // rows staged so far by this CTA
    for (int b = whole.bid; b < a.B; b += whole.nblk) {
      const int len = a.lengths ? a.lengths[b] : a.T;
      const float* g = a.emis + (long long)b * a.T * c.N;
      for (int t = 0; t < len; ++t, ++r) {
        mbarWait(v.mbar(MB_ROW_FREE), (r + 1) & 1); // the consumers gathered from the previous row
        if (p.tid == 0) {
          fenceProxyAsync();
          mbarArriveExpectTx(v.mbar(MB_ROW_FULL), rowBytes);
          bulkLoad(v.row(), g + (long long)t * c.N, rowBytes, v.mbar(MB_ROW_FULL));
        }
        mbarWait(v.mbar(MB_ROW_FULL), r & 1);
        const int slot = (int)(r & 1);
        mbarWait(v.mbar(MB_LIST_FREE0 + slot), ((r >> 1) + 1) & 1);
        topmRowStaged(p, tc, ps, v.row(), v.listTok(slot), v.listVal(slot, c.M), v.thr(slot));
        if (p.tid == 0) mbarArrive(v.mbar(MB_LIST_READY0 + slot));
      }
    }
    return;
  }
  /* ---------------- consumer warps ---------------- */
  const Cta cta{whole.tid, kFusedConsumers, whole.bid, whole.nblk, 1};
#else
  const Cta cta = whole;
  for (int i = 0; i < tc.capS; ++i) ps.rankCnt[i] = 0;
  auto produce = [&](int b, int t, int slot) { // host model: the producer's work for row (b, t), inline
    const float* gp = a.emis + ((long long)b * a.T + t) * c.N;
    float* row = v.row();
    for (int i = 0; i < c.N; ++i) row[i] = gp[i];
    topmRowStaged(cta, tc, ps, row, v.listTok(slot), v.listVal(slot, c.M), v.thr(slot));
  };
#endif
  ctaInitWorkspace(cta, c, w, smem + fl.ws);
  uint32_t g = 0; // frames consumed so far by this CTA (same sequence as the producer's r)
  for (int b = whole.bid; b < a.B; b += whole.nblk) {
    const int len = a.lengths ? a.lengths[b] : a.T;
    int curIdx = 0;
    cta.sync(); // previous utterance fully retired
    if (cta.tid == 0) seedUtterance(c, w, a, b);
    cta.sync();
    for (int t = 0; t < len; ++t, ++g) {
      if (t == 0) { // emissions of the seed hypothesis from row 0
#if FLT_DEVICE_BUILD
        mbarWait(v.mbar(MB_ROW_FULL), g & 1);
#else
        produce(b, 0, (int)(g & 1));
#endif
        lfGatherSpec(cta, c, w, w.beam(curIdx), w.sc()[SC_NH], v.row());
        cta.sync();
#if FLT_DEVICE_BUILD
        if (cta.tid == 0) mbarArrive(v.mbar(MB_ROW_FREE));
#endif
      }
#if FLT_DEVICE_BUILD
      mbarWait(v.mbar(MB_LIST_READY0 + (g & 1)), (g >> 1) & 1);
#endif
      // ring slot = running row count & 1, on both sides
      const FrameIn f = fusedFrameIn(c, a, v, b, t, (int)(g & 1));
      lfFrameStep(cta, c, w, w.beam(curIdx), w.beam(curIdx ^ 1), f, a.stats); // ends with a barrier
      curIdx ^= 1;
#if FLT_DEVICE_BUILD
      if (cta.tid == 0) mbarArrive(v.mbar(MB_LIST_FREE0 + (g & 1)));
#endif
      if (t + 1 < len) { // emissions of the new beam from row t+1
#if FLT_DEVICE_BUILD
        mbarWait(v.mbar(MB_ROW_FULL), (g + 1) & 1);
#else
        produce(b, t + 1, (int)((g + 1) & 1));
#endif
        lfGatherSpec(cta, c, w, w.beam(curIdx), w.sc()[SC_NH], v.row());
        cta.sync();
#if FLT_DEVICE_BUILD
        if (cta.tid == 0) mbarArrive(v.mbar(MB_ROW_FREE));
#endif
      }
    }
    int nFin = 0;
    if (w.sc()[SC_NH] != 0) {
      const FrameIn f = finishFrameIn(c, a, b, len);
      lfFinish(cta, c, w, w.beam(curIdx), w.beam(curIdx ^ 1), f);
      curIdx ^= 1;
      nFin = w.sc()[SC_NH];
    }
    cta.sync();
    writeFinals(cta, c, w, a, b, curIdx, nFin);
  }
}

} // namespace flt
