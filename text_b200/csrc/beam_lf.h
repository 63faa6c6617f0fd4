// beam_lf.h — the lexicon-free frame step (LexiconFreeDecoder + ZeroLM, max-merge), written for
// latency: five CTA barriers per frame, no allocation atomics, no merge table, no radix passes.
//
// Replaces LexiconFreeDecoder::decodeStep's expansion (decoder/LexiconFreeDecoder.cpp:53-112) and
// candidatesStore (decoder/Utils.h:146-225) for one frame; decodeEnd (LexiconFreeDecoder.cpp:127-158)
// is lfFinish below. Same exact-pruning argument as beam_core.h, restated in hypothesis-index
// space so that no row ranking is needed:
//
//   * The beam is sorted by score (index = rank). Hypotheses sharing an LM state form a row; a row
//     of this decoder has at most two members, (S, x, prevBlank=0) and (S, blank, prevBlank=1).
//     Hypotheses 0..i therefore span >= i/2+1 rows whose best member scores >= score(i), and the
//     new-token candidate of hypothesis i with the j-th ranked token is dominated by
//     (i/2+1)*(j-1) distinct merge groups: it is generated only if (i/2+1)*(j-2) <= K.
//   * Merge groups (decoder/Utils.h:176-198, max-merge) are resolved where candidates are
//     generated, so every materialised candidate is a distinct group:
//       - same row, same new token / blank: only the better (lower-index) eligible member emits;
//       - new token n from state S vs the repeat of a hypothesis already in child(S, n): both
//         sides look the other up in a table of the beam's LM-state fingerprints and only the
//         winner (higher score, then lower parent index) emits.
//   * Top-K: candidate scores lie in [tau, U] (tau = corner bound, U = best hypothesis + row
//     maximum); a monotone linear map sends them to NB bins and a per-warp suffix scan of the
//     histogram gives every candidate the number of candidates in higher bins. Candidates with
//     fewer than K above them (the top-K plus the rest of the cut bin) get their exact rank by
//     comparing with the other members of their own bin; rank < K = survivor, placed at
//     ranked[rank]. Crowded bins only lengthen those comparisons, they never change the result.
#pragma once
#include "beam_core.h"

namespace flt {

struct LfTab { // fingerprint -> row members
  int* a;      // [capRH] first member (claims the slot), -1 = empty
  int* b;      // [capRH] second member or -1
  int* slotOf; // [K] slot of hypothesis i
  uint32_t mask;
};

// One table per beam (index = beam index): the table of the NEXT beam is filled by the threads that create its
// hypotheses (end of lfFrameStep) while the current beam's is emptied, so a frame starts with its table ready.
FLT_DEV LfTab lfTab(const Ws& w, int beamIdx) {
  const Lay& L = w.c->lay;
  const int cap = w.c->capRH;
  return LfTab{(int*)(w.base + L.rowHash) + beamIdx * cap, (int*)(w.base + L.lfSlotB) + beamIdx * cap,
               (int*)(w.base + L.lfSlotOf) + beamIdx * w.c->K, (uint32_t)cap - 1};
}
// hypothesis i of `beam` (fingerprint fa, fb already stored in the beam) into its table
FLT_DEV void lfTabInsert(const LfTab& t, const Beam& beam, int i, u64 fa, u64 fb) {
  uint32_t s = (uint32_t)fa & t.mask;
  for (;;) {
    const int old = atomCAS(&t.a[s], -1, i);
    if (old == -1) break;
    if (beam.fpA(old) == fa && beam.fpB(old) == fb) {
      t.b[s] = i; // a row has at most two members
      break;
    }
    s = (s + 1) & t.mask;
  }
  t.slotOf[i] = (int)s;
}
// table of a beam that was not produced by lfFrameStep (seed, restored stream beam); the caller synchronises
FLT_DEV void lfTabBuild(const Cta& cta, const Ws& w, int beamIdx, int nH) {
  const LfTab t = lfTab(w, beamIdx);
  const Beam beam = w.beam(beamIdx);
  for (int i = cta.tid; i < nH; i += cta.nthr) lfTabInsert(t, beam, i, beam.fpA(i), beam.fpB(i));
}
FLT_DEV void lfTabClear(const Cta& cta, const Ws& w, int beamIdx, int nH) {
  const LfTab t = lfTab(w, beamIdx);
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    const int s = t.slotOf[i];
    t.a[s] = -1;
    t.b[s] = -1;
  }
}
FLT_DEV unsigned short* lfCbin(const Ws& w) { return (unsigned short*)(w.base + w.c->lay.lfCbin); }

// new-token test of hypothesis p for token n (LexiconFreeDecoder.cpp:69-71)
FLT_DEV bool lfEligible(const DecCfg& c, const Beam& cur, int p, int n) {
  if (c.ctc) return n != c.blank && (n != cur.tok(p) || cur.pb(p));
  return n != cur.tok(p);
}

// row members of the LM state with fingerprint (xa, xb); false if the state is not in the beam
FLT_DEV bool lfProbe(const LfTab& t, const Beam& cur, u64 xa, u64 xb, int& ma, int& mb) {
  uint32_t s = (uint32_t)xa & t.mask;
  for (;;) {
    const int a = t.a[s];
    if (a < 0) return false;
    if (cur.fpA(a) == xa && cur.fpB(a) == xb) {
      ma = a;
      mb = t.b[s];
      return true;
    }
    s = (s + 1) & t.mask;
  }
}

// deterministic order of two candidates: higher score, lower parent, lower token, prevBlank
FLT_DEV bool lfBetter(const Cand& cd, int a, int b) {
  const double sa = cd.score(a), sb = cd.score(b);
  if (sa != sb) return sa > sb;
  if (cd.par(a) != cd.par(b)) return cd.par(a) < cd.par(b);
  if (cd.tok(a) != cd.tok(b)) return cd.tok(a) < cd.tok(b);
  return (cd.flags(a) & CF_PB) < (cd.flags(b) & CF_PB);
}

// candidate score of a hypothesis with score ps taking token n with emission ev
// (LexiconFreeDecoder.cpp:64-67, :75 with ZeroLM's 0.0f; transitions only enter
// emittingModelScore, :59-63)
FLT_DEV double lfScoreOf(const DecCfg& c, double ps, int n, float ev, bool isNew) {
  double score = ps + (double)ev;
  if (n == c.sil) score += c.silScore;
  if (isNew) score = score + c.lmWeight * (double)0.0f;
  return score;
}
FLT_DEV double lfScore(const DecCfg& c, const Beam& cur, int p, int n, float ev, bool isNew) {
  return lfScoreOf(c, cur.score(p), n, ev, isNew);
}

// everything an item needs about its hypothesis, loaded up front so that the shared-memory
// latencies overlap instead of chaining behind the item's branches
struct LfHyp {
  int i, tok, pb;
  int sa, sb;   // the (at most two) members of its row
  double score;
  u64 fa, fb;   // fingerprint of its LM state (cells) or of the parent state (repeat items)
};
FLT_DEV LfHyp lfLoadHyp(const Beam& cur, const LfTab& t, int i, bool parentFp) {
  LfHyp h;
  h.i = i;
  h.tok = cur.tok(i);
  h.pb = cur.pb(i);
  h.score = cur.score(i);
  h.fa = parentFp ? cur.pfpA(i) : cur.fpA(i);
  h.fb = parentFp ? cur.pfpB(i) : cur.fpB(i);
  const int s = t.slotOf[i];
  h.sa = t.a[s];
  h.sb = t.b[s];
  return h;
}
FLT_DEV bool lfEligibleH(const DecCfg& c, const LfHyp& h, int n) {
  if (c.ctc) return n != c.blank && (n != h.tok || h.pb);
  return n != h.tok;
}

// new-token candidate of hypothesis h with token n: true if it is to be materialised
FLT_DEV bool lfCell(const DecCfg& c, const Beam& cur, const LfTab& t, const LfHyp& h, int n, float ev,
                    double tau, double& score) {
  if (!lfEligibleH(c, h, n)) return false;
  score = lfScoreOf(c, h.score, n, ev, true);
  if (score < tau) return false; // about half of the cells end here: test the bound first
  const int i = h.i;
  const int partner = h.sa == i ? h.sb : h.sa;
  if (partner >= 0 && partner < i && lfEligible(c, cur, partner, n)) return false; // the better member emits
  // a hypothesis already in the child state whose repeat has the same key (child(S,n), n, 0)
  u64 ca, cb;
  fpChild(h.fa, h.fb, n, ca, cb);
  int ma = -1, mb = -1;
  if (lfProbe(t, cur, ca, cb, ma, mb)) {
    int m = -1;
    if (cur.tok(ma) == n && !cur.pb(ma)) m = ma;
    else if (mb >= 0 && cur.tok(mb) == n && !cur.pb(mb)) m = mb;
    if (m >= 0) {
      const double sm = lfScore(c, cur, m, n, ev, false);
      if (sm > score || (sm == score && m < i)) return false; // the repeat wins the merge
    }
  }
  return true;
}

// repeat candidate of hypothesis h (LexiconFreeDecoder.cpp:98-110); h carries the PARENT state's fingerprint
FLT_DEV bool lfRepeat(const DecCfg& c, const Beam& cur, const LfTab& t, const FrameIn& f, const LfHyp& h,
                      float eOwn, double tau, double& score) {
  const int n = h.tok, i = h.i;
  const bool isRepeat = c.ctc ? (!h.pb && n != c.blank) : true;
  if (!(isRepeat && n >= 0 && n < c.N && inTokenSetV(c, f, n, eOwn))) return false;
  score = lfScoreOf(c, h.score, n, eOwn, false);
  if (score < tau) return false;
  // new token n from the parent state merges into the same key
  int ma = -1, mb = -1;
  if (lfProbe(t, cur, h.fa, h.fb, ma, mb)) {
    int p = -1;
    if (lfEligible(c, cur, ma, n)) p = ma;
    if (mb >= 0 && lfEligible(c, cur, mb, n) && (p < 0 || mb < p)) p = mb;
    if (p >= 0) {
      const double sp = lfScore(c, cur, p, n, eOwn, true);
      if (sp > score || (sp == score && p < i)) return false;
    }
  }
  return true;
}

// blank candidate of hypothesis h (LexiconFreeDecoder.cpp:86-97): the better member of a row emits
FLT_DEV bool lfBlank(const DecCfg& c, const FrameIn& f, const LfHyp& h, float eBlank, double tau,
                     double& score) {
  if (!c.ctc) return false;
  const int partner = h.sa == h.i ? h.sb : h.sa;
  if (partner >= 0 && partner < h.i) return false;
  if (!inTokenSetV(c, f, c.blank, eBlank)) return false;
  score = lfScoreOf(c, h.score, c.blank, eBlank, false);
  return !(score < tau);
}

// e[own token] of every hypothesis, e[blank], e[sil] of one emission row -> spec[]
FLT_DEV void lfGatherSpec(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& beam, int nH,
                          const float* row) {
  float* spec = w.spec();
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    const int n = beam.tok(i);
    spec[i] = (n >= 0 && n < c.N) ? row[n] : 0.0f;
  }
  if (cta.tid == cta.nthr - 1) {
    spec[c.K] = c.ctc ? row[c.blank] : 0.0f;
    spec[c.K + 1] = row[c.sil];
  }
}

// phase timing for the benchmark (thread 0, only while the decoder's timing is on): cycles spent
// up to each barrier exit, summed over frames into stats[4 + phase]
struct LfPhaseClock {
  unsigned long long* stats;
  long long t0;
  FLT_DEV void start(const Cta& cta, unsigned long long* st) {
    stats = cta.tid == 0 ? st : nullptr;
#if FLT_DEVICE_BUILD
    if (stats) t0 = clock64();
#endif
  }
  FLT_DEV void mark(int phase) {
#if FLT_DEVICE_BUILD
    if (stats) {
      const long long t1 = clock64();
      atomicAdd(stats + 4 + phase, (unsigned long long)(t1 - t0));
      t0 = t1;
    }
#else
    (void)phase;
#endif
  }
};

FLT_DEV int* lfItemDesc(const Ws& w) { return (int*)(w.base + w.c->lay.lfDesc); }
// Work items of a frame: K repeat items, K blank items and (silScore > 0 only) K sil cells first — the repeat
// item of hypothesis p is item p, i.e. it belongs to thread p, which created the hypothesis and already holds
// its emission of this frame in a register — then the cells (hypothesis, ranked column) of all K hypotheses,
// column-major. Item x owns candidate slot x; items of hypotheses >= nH are dead.
// desc = hypothesis | column << 12 | kind << 24.
FLT_DEV void lfBuildItemDesc(const Cta& cta, const DecCfg& c, const Ws& w) {
  // the host orders the cells column-major (column 0 of every hypothesis, then column 1, ...): the
  // cells most likely to survive the bound sit together in the first warps' first sweep, and the
  // late sweeps are mostly cells that end at the bound test
  int* desc = lfItemDesc(w);
  for (int x = cta.tid; x < c.capC; x += cta.nthr) desc[x] = c.lfDesc[x];
}

FLT_DEV void lfFrameStep(const Cta& cta, const DecCfg& c, const Ws& w, int curIdx, const FrameIn& f,
                         unsigned long long* stats, LfCarry& carry) {
  int* sc = w.sc();
  const int nH = sc[SC_NH];
  if (nH == 0) return; // the beam died (Utils.h:155-158)
  LfPhaseClock pc;
  pc.start(cta, stats);
  const int K = c.K;
  const Beam cur = w.beam(curIdx), nxt = w.beam(curIdx ^ 1);
  const LfTab t = lfTab(w, curIdx);     // filled when this beam was created
  const LfTab tn = lfTab(w, curIdx ^ 1); // empty; filled below with the new beam
  const Cand cd = w.cand();
  float* spec = w.spec();
  int* hist = w.hist();
  const int NB = c.lfBins;

  // The scattered emission reads of the special items (e[own token] per hypothesis, e[blank], e[sil]). With
  // K <= threads every special item is processed by a thread that holds its value in a register: thread p
  // created hypothesis p at the end of the previous frame and loaded e_next[token] then (carry); every thread
  // loaded e_next[blank] and e_next[sil]. No shared-memory hand-over, no barrier before the expansion.
  const bool direct = K <= cta.nthr;
  float eOwnR = 0.0f, eBlankR = 0.0f, eSilR = 0.0f;
  if (direct) {
    if (carry.valid) {
      eOwnR = carry.eOwn;
      eBlankR = carry.eBlank;
      eSilR = carry.eSil;
    } else { // first frame of the utterance (or of a stream chunk)
      const int n = cta.tid < nH ? cur.tok(cta.tid) : -1;
      eOwnR = (n >= 0 && n < c.N) ? f.e[n] : 0.0f;
      eBlankR = c.ctc ? f.e[c.blank] : 0.0f;
      eSilR = f.e[c.sil];
    }
  } else { // beams wider than the CTA: published through shared memory
    if (carry.valid) {
      if (cta.tid < nH) spec[cta.tid] = carry.eOwn;
      for (int i = cta.tid + cta.nthr; i < nH; i += cta.nthr) {
        const int n = cur.tok(i);
        spec[i] = (n >= 0 && n < c.N) ? f.e[n] : 0.0f;
      }
      if (cta.tid == cta.nthr - 1) {
        spec[K] = carry.eBlank;
        spec[K + 1] = carry.eSil;
      }
    } else {
      lfGatherSpec(cta, c, w, cur, nH, f.e);
    }
    cta.sync();
  }

  // corner bound (beam_core.h frameStep): lanes 0..nTau-1 of every warp take one rectangle each
  auto cornerBound = [&]() {
  double tau = negInf();
  {
#if FLT_DEVICE_BUILD
    const int lane = cta.tid & 31;
    if (lane < c.nTau) {
      const int i = 2 * c.tauA[lane] - 2, col = c.tauCol[lane];
      if (i < nH && col < f.listLen && f.topTok[col] >= 0)
        tau = cur.score(i) + (double)f.topVal[col] + c.lmWeight * (double)0.0f;
    }
    for (int o = 8; o > 0; o >>= 1) { // nTau <= 16
      const double u = __shfl_xor_sync(0xffffffffu, tau, o);
      tau = u > tau ? u : tau;
    }
    tau = __shfl_sync(0xffffffffu, tau, 0); // lanes 0..15 hold the maximum
#else
    for (int k = 0; k < c.nTau; ++k) {
      const int i = 2 * c.tauA[k] - 2, col = c.tauCol[k];
      if (i < nH && col < f.listLen && f.topTok[col] >= 0) {
        const double corner = cur.score(i) + (double)f.topVal[col] + c.lmWeight * (double)0.0f;
        tau = corner > tau ? corner : tau;
      }
    }
#endif
    if (c.silScore < 0) tau += c.silScore; // keeps the bound valid if a counted cell is the sil one
  }
  return tau;
  };
  double tau = cornerBound();
  // candidate scores lie in [tau, upper]: monotone linear map to NB bins
  const float eTop = (f.listLen > 0 && f.topTok[0] >= 0) ? f.topVal[0] : 0.0f;
  double upper = cur.score(0) + (double)eTop;
  if (c.silScore > 0) upper += c.silScore;
  // Guessed bound. The corner bound is safe but loose (about six candidates pass it for every survivor). In
  // steady state the K-th best candidate sits about as far below `upper` as it did one frame earlier, so the
  // frame is first expanded against upper - gfac * (that distance): if at least K candidates pass, they contain
  // the K best (every materialised candidate is a distinct merge group) and the result is exact; if fewer
  // pass, the frame is expanded again against the corner bound and guessing pauses for a few frames.
  // (the three scalars of the guess live in the workspace, not in registers: the kernel is at its register cap)
  bool guessed = false;
  {
    const float gap = bitsF32((uint32_t)sc[SC_LFGAP]);
    if (gap >= 0.0f && sc[SC_LFHOLD] == 0 && nH == K && !(c.dbg & 16)) {
      const double g = upper - (double)(gap * bitsF32((uint32_t)sc[SC_LFFAC]));
      if (g > tau) {
        tau = g;
        guessed = true;
      }
    }
  }
  pc.mark(0); // (no barrier: the beam, its table and the zeroed histogram were published by the previous frame's last one)

  // (2) candidates, each in the slot of its work item, and the histogram of their scores
  // item x -> (hypothesis, column | kind) is fixed for the CTA's lifetime (lfItemDesc, built once):
  // K repeat, K blank and K sil-cell items first, then the cells of all K hypotheses
  const int* desc = lfItemDesc(w);
  const int items = c.wideTotal + (c.silScore > 0 ? 3 : 2) * K;
  unsigned short* cbin = lfCbin(w);
  unsigned short* cslot = cbin + c.capC;
  int* list = w.rep();  // [capC] relevant candidates
  u64* lkey = w.rkey(); // [capC] their ordered score keys
  int total = 0;        // live candidates of the frame
  int nRel = 0;         // relevant ones (list length)
  unsigned short* above;
  for (;;) { // at most twice: against the guessed bound, then (on a miss) against the corner bound
  // (double -> float conversion, the float subtraction of a constant and the multiplication by a
  // positive constant are all monotone, so bins never invert the score order)
  const float rangeF = (float)(upper - tau);
  const bool binned = rangeF > 0.0f && rangeF < 3.0e38f; // finite, non-degenerate
  const float scaleF = binned ? (float)NB / rangeF : 0.0f;
  for (int x = cta.tid; x < items; x += cta.nthr) {
    bool alive = false;
    double score = 0.0;
    int tok = 0, flags = 0;
    float ev = 0.0f;
    const int dsc = desc[x];
    const int par = dsc & 0xFFF, kind = dsc >> 24;
    if (par >= nH) {
      cd.parflag(x) = 0;
      continue;
    }
    const LfHyp h = lfLoadHyp(cur, t, par, kind == 1);
    if (kind == 0) {
      const int j = (dsc >> 12) & 0xFFF;
      if (j < f.listLen) {
        tok = f.topTok[j];
        ev = f.topVal[j];
        flags = CF_NEW;
        // a boosted sil is not rank-dominated: every hypothesis proposes it as a special item
        if (tok >= 0 && !(tok == c.sil && c.silScore > 0)) alive = lfCell(c, cur, t, h, tok, ev, tau, score);
      }
    } else {
      if (kind == 1) {
        tok = h.tok;
        ev = direct ? eOwnR : spec[par]; // direct: item par belongs to thread par
        alive = lfRepeat(c, cur, t, f, h, ev, tau, score);
      } else if (kind == 2) {
        tok = c.blank;
        ev = direct ? eBlankR : spec[K];
        flags = CF_PB;
        alive = lfBlank(c, f, h, ev, tau, score);
      } else {
        tok = c.sil;
        ev = direct ? eSilR : spec[K + 1];
        flags = CF_NEW;
        if (inTokenSetV(c, f, tok, ev)) alive = lfCell(c, cur, t, h, tok, ev, tau, score);
      }
    }
    if (alive) {
      int bin = 0;
      if (binned) {
        const float pos = (float)(score - tau) * scaleF;
        bin = pos >= (float)(NB - 1) ? NB - 1 : (int)pos;
        bin = bin < 0 ? 0 : bin;
      }
      cbin[x] = (unsigned short)bin;
      cslot[x] = (unsigned short)atomAdd(&hist[bin], 1); // arrival order inside the bin
      cd.score(x) = score;
      cd.parflag(x) = (par << 4) | flags | CF_ALIVE;
      cd.tok(x) = tok;
      cd.ce(x) = ev;
    } else {
      cd.parflag(x) = 0;
    }
  }
  cta.sync(); // ---- B2
  pc.mark(1);

  // (3) above[bin] = candidates in higher bins, from a suffix scan of the histogram that every
  // warp does for itself into a private copy (lane L owns NB/32 consecutive bins). Candidates with
  // above < K (the K best and the rest of the cut bin) are "relevant" and go to list position
  // above[bin] + arrival order: the list is grouped by bin, best bins first, with no further atomics.
  total = 0;
  nRel = 0;
#if FLT_DEVICE_BUILD
  above = (unsigned short*)(w.base + c.lay.lfAbove) + (cta.tid >> 5) * NB;
  {
    const int lane = cta.tid & 31;
    if (NB == 256) {
      const int4 lo = *(const int4*)(hist + lane * 8), hi = *(const int4*)(hist + lane * 8 + 4);
      const int h[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      const int own = ((h[0] + h[1]) + (h[2] + h[3])) + ((h[4] + h[5]) + (h[6] + h[7]));
      int suf = own; // inclusive suffix over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += u;
      }
      total = __shfl_sync(0xffffffffu, suf, 0);
      int ab = suf - own; // candidates in bins of higher lanes
      unsigned short av[8];
      int rel = 0;
#pragma unroll
      for (int k = 7; k >= 0; --k) {
        av[k] = (unsigned short)(ab > 65535 ? 65535 : ab);
        rel += ab < K ? h[k] : 0;
        ab += h[k];
      }
      nRel = __reduce_add_sync(0xffffffffu, rel);
      uint4 pk;
      pk.x = av[0] | ((unsigned)av[1] << 16);
      pk.y = av[2] | ((unsigned)av[3] << 16);
      pk.z = av[4] | ((unsigned)av[5] << 16);
      pk.w = av[6] | ((unsigned)av[7] << 16);
      *(uint4*)(above + lane * 8) = pk;
    } else {
      const int per = NB >> 5;
      int own = 0;
      for (int k = 0; k < per; ++k) own += hist[lane * per + k];
      int suf = own;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += u;
      }
      total = __shfl_sync(0xffffffffu, suf, 0);
      int ab = suf - own;
      int rel = 0;
      for (int k = per - 1; k >= 0; --k) {
        const int hk = hist[lane * per + k];
        above[lane * per + k] = (unsigned short)(ab > 65535 ? 65535 : ab);
        rel += ab < K ? hk : 0;
        ab += hk;
      }
      nRel = __reduce_add_sync(0xffffffffu, rel);
    }
    __syncwarp();
  }
#else
  above = (unsigned short*)(w.base + c.lay.lfAbove);
  for (int bn = NB - 1; bn >= 0; --bn) {
    above[bn] = (unsigned short)(total > 65535 ? 65535 : total);
    if (total < K) nRel += hist[bn];
    total += hist[bn];
  }
#endif
  if (!(guessed && total < K)) break;
  // the guess cut too deep (fewer than K candidates passed): the same frame against the corner bound
  cta.sync(); // every warp has read the histogram
  for (int bn = cta.tid; bn < NB; bn += cta.nthr) hist[bn] = 0;
  tau = cornerBound();
  guessed = false;
  if (cta.tid == 0) {
    const float fac = bitsF32((uint32_t)sc[SC_LFFAC]) * 1.25f;
    sc[SC_LFFAC] = (int)f32Bits(fac < 3.0f ? fac : 3.0f);
    sc[SC_LFHOLD] = 9; // decremented below: 8 frames without guessing
#if FLT_DEVICE_BUILD
    if (stats) atomicAdd(stats + 13, 1ull);
#endif
  }
  cta.sync();
  } // for (;;)
#if FLT_DEVICE_BUILD
  if (stats && cta.tid == 0 && guessed) atomicAdd(stats + 12, 1ull);
#endif
  const int nSel = total < K ? total : K;
  for (int x = cta.tid; x < items; x += cta.nthr) {
    if (!(cd.parflag(x) & CF_ALIVE)) continue;
    const int ab = above[cbin[x]];
    if (ab >= K) continue; // K candidates score strictly higher
    const int q = ab + cslot[x];
    list[q] = x;
    lkey[q] = orderedKey64(cd.score(x));
  }
  cta.sync(); // ---- B3
  pc.mark(2);
  // (4) exact ranks among the nRel relevant candidates by counting: `parts` adjacent lanes share
  // one candidate and split the list between them (equal scores: lower parent, then lower token)
  int* ranked = w.surv() + c.capP;
  {
    int lg = 0;
    while (lg < 5 && (nRel << (lg + 1)) <= cta.nthr) ++lg;
    const int parts = 1 << lg;
    const int per = cta.nthr >> lg; // candidates per sweep
    const int part = cta.tid & (parts - 1);
    for (int a0 = 0; a0 < nRel; a0 += per) {
      const int qa = a0 + (cta.tid >> lg);
      int cnt = 0;
      int xa = -1;
      if (qa < nRel) {
        xa = list[qa];
        const u64 ka = lkey[qa];
        bool tie = false;
#pragma unroll 4
        for (int qb = part; qb < nRel; qb += parts) {
          const u64 kb = lkey[qb];
          cnt += kb > ka ? 1 : 0;
          tie |= (kb == ka) & (qb != qa);
        }
        if (tie) // another candidate with exactly this score: settle by work item (rare)
          for (int qb = part; qb < nRel; qb += parts)
            if (lkey[qb] == ka && qb != qa && lfBetter(cd, list[qb], xa)) ++cnt;
      }
#if FLT_DEVICE_BUILD
      for (int o = 1; o < parts; o <<= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#endif
      if (xa >= 0 && part == 0 && cnt < K) ranked[cnt] = xa;
    }
  }
  cta.sync(); // ---- B4
  pc.mark(3);
  for (int bn = cta.tid; bn < NB; bn += cta.nthr) hist[bn] = 0;
#if FLT_DEVICE_BUILD
  if (stats && cta.tid == 0) {
    atomicAdd(stats + 0, 1ull);
    atomicAdd(stats + 1, (unsigned long long)items);
    atomicAdd(stats + 2, (unsigned long long)total); // live candidates
    atomicAdd(stats + 3, (unsigned long long)nSel);
  }
#else
  (void)stats;
#endif

  // (5) the new beam (Utils.h:161-165 threshold against the best, then the K best in rank order) and its
  // fingerprint table; the current beam's table is emptied for the beam after next
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    const int s = t.slotOf[i];
    t.a[s] = -1;
    t.b[s] = -1;
  }
  if (nSel == 0) {
    if (cta.tid == 0) sc[SC_NH] = 0;
  } else {
    const double thrScore = cd.score(ranked[0]) - c.beamThreshold;
    // next frame's guess: how far below `upper` the K-th best candidate was; the factor follows how many
    // candidates passed (aim: 1.5 K .. 3 K)
    if (cta.tid == 0) {
      if (sc[SC_LFHOLD] > 0) --sc[SC_LFHOLD];
      if (nSel == K) {
        sc[SC_LFGAP] = (int)f32Bits((float)(upper - cd.score(ranked[K - 1])));
        if (guessed) {
          float fac = bitsF32((uint32_t)sc[SC_LFFAC]);
          if (total > 3 * K) fac = fac * 0.97f > 1.05f ? fac * 0.97f : 1.05f;
          else if (total < K + K / 2) fac *= 1.04f;
          sc[SC_LFFAC] = (int)f32Bits(fac);
        }
      } else {
        sc[SC_LFGAP] = (int)f32Bits(-1.0f);
      }
    }
    for (int q = cta.tid; q < nSel; q += cta.nthr) {
      const int x = ranked[q];
      const double score = cd.score(x);
      if (!(score >= thrScore)) continue;
      if (q + 1 == nSel || !(cd.score(ranked[q + 1]) >= thrScore)) sc[SC_NH] = q + 1;
      const int p = cd.par(x);
      const int fl = cd.flags(x);
      const int n = cd.tok(x);
      nxt.score(q) = score;
      nxt.am(q) = cur.am(p) + amOf(c, f, cd.ce(x), n, cur.tok(p));
      nxt.lm(q) = cur.lm(p) + (double)0.0f;
      nxt.lex(q) = 0;
      nxt.tok(q) = n;
      nxt.pb(q) = (fl & CF_PB) ? 1 : 0;
      u64 fa, fb;
      if (fl & CF_NEW) {
        fpChild(cur.fpA(p), cur.fpB(p), n, fa, fb);
        nxt.pfpA(q) = cur.fpA(p);
        nxt.pfpB(q) = cur.fpB(p);
      } else {
        fa = cur.fpA(p);
        fb = cur.fpB(p);
        nxt.pfpA(q) = cur.pfpA(p);
        nxt.pfpB(q) = cur.pfpB(p);
      }
      nxt.fpA(q) = fa;
      nxt.fpB(q) = fb;
      { // into the new beam's table. A slot already claimed by another new hypothesis q2 is compared through
        // what q2 was made FROM (its candidate and parent, all written before the last barrier), not through
        // the fingerprint q2's thread is storing in this very phase: no ordering between the two threads needed
        uint32_t s = (uint32_t)fa & tn.mask;
        for (;;) {
          const int q2 = atomCAS(&tn.a[s], -1, q);
          if (q2 == -1) break;
          const int x2 = ranked[q2];
          const int p2 = cd.par(x2);
          u64 oa = cur.fpA(p2), ob = cur.fpB(p2);
          if (cd.flags(x2) & CF_NEW) fpChild(oa, ob, cd.tok(x2), oa, ob);
          if (oa == fa && ob == fb) {
            tn.b[s] = q; // a row has at most two members
            break;
          }
          s = (s + 1) & tn.mask;
        }
        tn.slotOf[q] = (int)s;
      }
      f.hParent[q] = p;
      f.hTok[q] = n;
      const int an = skipCarry(cur, f.hRow, p);
      nxt.anc(q) = an;
      if (f.hSkip) f.hSkip[q] = an;
      if (f.hScore) {
        f.hScore[3 * q] = score;
        f.hScore[3 * q + 1] = nxt.am(q);
        f.hScore[3 * q + 2] = nxt.lm(q);
      }
      // the emission this hypothesis needs in the next frame: issue the (L2 / HBM) load now, it
      // lands while the barrier and the next frame's first phase run
      if (q == cta.tid && f.eNext && n >= 0 && n < c.N) carry.eOwn = f.eNext[n];
    }
  }
  carry.valid = f.eNext != nullptr;
  if (f.eNext && (direct || cta.tid == cta.nthr - 1)) { // direct: every thread keeps its own copy
    carry.eBlank = c.ctc ? f.eNext[c.blank] : 0.0f;
    carry.eSil = f.eNext[c.sil];
  }
  cta.sync(); // ---- B5
  pc.mark(4);
  if (f.hCount && cta.tid == 0) *f.hCount = sc[SC_NH];
}

// decodeEnd (LexiconFreeDecoder.cpp:127-158) with ZeroLM: finish() returns the same state and 0,
// every hypothesis proposes (state, sil, prevBlank=0); the two members of a row merge (max).
// The beam is already sorted, so the survivors keep their order.
FLT_DEV void lfFinish(const Cta& cta, const DecCfg& c, const Ws& w, int curIdx, const FrameIn& f) {
  int* sc = w.sc();
  const int nH = sc[SC_NH];
  if (nH == 0) return;
  const Beam cur = w.beam(curIdx), nxt = w.beam(curIdx ^ 1);
  const LfTab t = lfTab(w, curIdx); // filled when the beam was created (the caller has synchronised since)
  int* keep = w.surv(); // [capP] flags
  const double best = cur.score(0) + c.lmWeight * (double)0.0f;
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    const int s = t.slotOf[i];
    const int a = t.a[s], b = t.b[s];
    const int partner = a == i ? b : a;
    const double score = cur.score(i) + c.lmWeight * (double)0.0f;
    keep[i] = !(partner >= 0 && partner < i) && score >= best - c.beamThreshold;
  }
  cta.sync();
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    const int s = t.slotOf[i];
    t.a[s] = -1;
    t.b[s] = -1;
    int q = 0;
    for (int j = 0; j < i; ++j) q += keep[j];
    if (keep[i]) {
      nxt.score(q) = cur.score(i) + c.lmWeight * (double)0.0f;
      nxt.am(q) = cur.am(i);
      nxt.lm(q) = cur.lm(i) + (double)0.0f;
      f.hParent[q] = i;
      f.hTok[q] = c.sil;
      f.hSkip[q] = skipCarry(cur, f.hRow, i);
      if (f.hScore) {
        f.hScore[3 * q] = nxt.score(q);
        f.hScore[3 * q + 1] = nxt.am(q);
        f.hScore[3 * q + 2] = nxt.lm(q);
      }
    }
    if (i == nH - 1) sc[SC_NH] = q + keep[i];
  }
  cta.sync();
  if (f.hCount && cta.tid == 0) *f.hCount = sc[SC_NH];
}

} // namespace flt
