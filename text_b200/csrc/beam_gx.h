// beam_gx.h — the single-pass frame step with a guessed cut ("gx"), for both decoders.
//
// Replaces, for one frame, LexiconDecoder::decodeStep / LexiconFreeDecoder::decodeStep's expansion
// (decoder/LexiconDecoder.cpp:54-215, decoder/LexiconFreeDecoder.cpp:53-112) and candidatesStore
// (decoder/Utils.h:146-225) in the max-merge, word-level-LM, CTC (lexicon) / CTC+ASG (lexicon-free)
// configurations — the ones BASELINE.json's configs use. Everything else stays on beam_core.h.
//
// Idea. The reference materialises every candidate of a frame, then keeps the K best merge groups.
// A group's score is the maximum of its members (Utils.h:176-198), so for ANY threshold G the K best
// groups are found exactly from the candidates scoring >= G alone, provided those candidates form
// at least K groups (a group with a member >= G has its best member >= G; a group without scores
// < G <= every kept group). The step therefore
//   E   proposes candidates and materialises only those >= G, where G is GUESSED from the previous
//       frame (best hypothesis + best proposal of this frame - last frame's exact spread x margin);
//       each materialised candidate is merged at once into a CTA-private table on its 128-bit
//       (LM state, lex node, token, prevBlank) key — compare-and-swap on the slot, the loser is
//       marked dead — so there is no separate merge pass;
//   C   compacts the surviving group representatives (one per key) and empties the table;
//   RF  ranks the representatives by counting and the thread that finds rank q < K writes
//       hypothesis q of the new beam directly (no ranked[] round trip): threshold against the best
//       (Utils.h:161-165), LM-state fingerprint / n-gram context, back-pointer record, skip pointer.
// Three CTA barriers per frame. If the guess was too high (< K groups although something was cut) or
// too low (candidate capacity exceeded), the frame is redone EXACTLY: one histogram pass over all
// proposals (256 monotone bins) picks G, then E again; the kept set is verified the same way. The
// result never depends on the guess — only the time does.
//
// Work distribution in E (no row grouping, no prefix sums, no binary searches on the hot path):
//   * "walkers" = hypotheses that expand over the frame's RANKED token list (all hypotheses of the
//     lexicon-free decoder; hypotheses at the Trie root in the lexicon decoder, ranked by
//     e[n] + lmWeight * smeared score of root child n). Score is monotone along the list, so a group
//     of 4..32 lanes walks it in chunks and stops at the first chunk entirely below G. The best
//     walkers get a whole warp, the tail four lanes each (static schedule over walker ranks).
//   * "items" = everything enumerated directly: the Trie edges of non-root hypotheses, root children
//     that carry labels (single-token words), stay / blank / repeat. When a hypothesis is created
//     (RF of the previous frame) its thread reserves the hypothesis' items in a chunk table (8 items
//     per descriptor), so item x maps to (hypothesis, edge) with two shared-memory loads. Trie
//     edges are {token, child} pairs (one coalesced 8-byte load), nodes are 16-byte records
//     {smeared score, first edge, #edges | #labels, first label}: two dependent L2 round trips per
//     edge, issued for all items of a sweep before the walkers run, consumed after.
//   * n-gram word LM: a word end is probed only if its score with the LM's best possible score
//     still reaches G.
#pragma once
#include "beam_core.h"
#include "beam_lf.h"

namespace flt {

constexpr int kGxChunkLog = 3, kGxChunk = 1 << kGxChunkLog; // items per chunk descriptor
constexpr int kGxBins = 256;
constexpr int kGxSweep = 2; // items per thread and sweep (loads of a sweep are in flight together)

enum { // per-frame scalars, two sets (index = parity of the beam the frame reads)
  GX_NCAND = 0, GX_OVF, GX_CUT, GX_NREP, GX_NCHUNK, GX_SET
};

struct GxCarry { // per-utterance state of the guess; uniform over the CTA's threads (registers)
  double spread; // (reference level - cut) to use on the next frame; +inf = unknown: take everything
  float mu;      // safety factor applied to the last frame's exact spread
  float eBlank, eSil; // e[blank], e[sil] of the NEXT frame, loaded while the current one retires
  int eValid;
};
FLT_DEV GxCarry gxCarryInit() { return GxCarry{bitsF64(0x7FF0000000000000ull), 0.30f, 0.0f, 0.0f, 0}; }

struct GxFrame { // uniform per attempt
  double G;      // candidates scoring below are not materialised
  double floor;  // the reference's own filter: candidates below never survive (Utils.h:161-165)
  int mode;      // 0 = materialise, 1 = histogram only
  double hlo;
  float hscale;
};

/* ------------------------------------------------------------------ workspace views ---------- */
FLT_DEV int* gxSc(const Ws& w, int set) { return w.sc() + SC_GX + set * GX_SET; }
FLT_DEV int* gxCandX(const Ws& w) { return (int*)(w.base + w.c->lay.gxCandX); }      // [3][capC]
FLT_DEV int* gxChunks(const Ws& w, int set) {                                          // [2][capChunks] x int2
  return (int*)(w.base + w.c->lay.gxChunk) + (size_t)set * 2 * w.c->capChunks;
}
FLT_DEV uint32_t* gxBits(const Ws& w, int set) { // walker bitmap [2][(K+31)/32]
  return (uint32_t*)(w.base + w.c->lay.gxBits) + (size_t)set * ((w.c->K + 31) >> 5);
}
FLT_DEV int* gxListInfo(const Ws& w, int buf) { // [2][Mwide] x int4 {ms bits, child, eoff, degLab}
  return (int*)(w.base + w.c->lay.gxList) + (size_t)buf * 4 * w.c->Mwide;
}

FLT_DEV int gxSpecials(const DecCfg& c) { // directly enumerated specials per hypothesis
  if (c.lexicon) return (c.silScore > 0 ? 3 : 2);      // stay, blank, boosted-sil cell
  return (c.silScore > 0 ? 3 : 2);                     // repeat, blank, boosted-sil cell
}

// node record of the Trie: {smeared score bits, first edge, #edges | #labels << 24, first label or label offset}
FLT_DEV int gxNodeDeg(int degLab) { return degLab & 0xFFFFFF; }
FLT_DEV int gxNodeLabels(int degLab) { return (unsigned)degLab >> 24; }

/* ------------------------------------------------------------------ hypothesis registration -- */
// Called by the thread that creates hypothesis q of beam `set`: reserve its items in the chunk table
// (lexicon) and mark it as a walker if it sits at the Trie root.
FLT_DEV void gxRegister(const DecCfg& c, const Ws& w, int set, int q, int lex, int degLab) {
  if (!c.lexicon) return;
  int* sc = gxSc(w, set);
  const int edges = lex == 0 ? c.trie.nRootLab : gxNodeDeg(degLab);
  const int cnt = edges + gxSpecials(c);
  const int nch = (cnt + kGxChunk - 1) >> kGxChunkLog;
  const int base = atomAdd(&sc[GX_NCHUNK], nch);
  int* ch = gxChunks(w, set);
  for (int z = 0; z < nch; ++z) {
    const int k0 = z << kGxChunkLog;
    if (base + z < c.capChunks) {
      ch[2 * (base + z)] = q | (k0 << 12);
      ch[2 * (base + z) + 1] = cnt - k0 < kGxChunk ? cnt - k0 : kGxChunk;
    }
  }
  if (lex == 0) {
#if FLT_DEVICE_BUILD
    atomicOr(&gxBits(w, set)[q >> 5], 1u << (q & 31));
#else
    gxBits(w, set)[q >> 5] |= 1u << (q & 31);
#endif
  }
}

// per-hypothesis Trie cache of a beam entry (lexicon): first edge, #edges | #labels, smeared score
FLT_DEV void gxSetNode(const DecCfg& c, const Beam& b, int q, int lex, int eoff, int degLab, int msBits) {
  (void)c;
  (void)lex;
  b.xv[q] = eoff;
  b.xv[b.K + q] = degLab;
  b.xv[2 * b.K + q] = msBits;
}

// Rebuild the tables of beam `set` from its hypotheses (seed of an utterance, restored online beam).
FLT_DEV void gxRebuild(const Cta& cta, const DecCfg& c, const Ws& w, int set, int nH) {
  if (!c.lexicon) return;
  if (cta.tid == 0) gxSc(w, set)[GX_NCHUNK] = 0;
  for (int i = cta.tid; i < ((c.K + 31) >> 5); i += cta.nthr) gxBits(w, set)[i] = 0;
  cta.sync();
  const Beam b = w.beam(set);
  for (int q = cta.tid; q < nH; q += cta.nthr) {
    const int lex = b.lex(q);
    int4 nd;
    nd.x = 0, nd.y = 0, nd.z = 0, nd.w = 0;
    if (lex != 0) nd = c.trie.node[lex];
    gxSetNode(c, b, q, lex, nd.y, nd.z, lex == 0 ? 0 : nd.x);
    gxRegister(c, w, set, q, lex, nd.z);
  }
  cta.sync();
}

/* ------------------------------------------------------------------ candidates ---------------- */
// insert candidate x into the merge table; the better of two candidates with equal keys keeps the
// slot, the other is marked dead (Utils.h:176-198, max-merge). Callable while other threads insert.
FLT_DEV void gxInsert(const DecCfg& c, const Ws& w, int x) {
  const Cand cd = w.cand();
  int* mh = w.mh();
  const uint32_t mask = (uint32_t)c.capH - 1;
#if FLT_DEVICE_BUILD
  __threadfence_block(); // the record is visible before the slot names it
#endif
  const u64 ka = cd.keyA(x), kb = cd.keyB(x);
  uint32_t s = (uint32_t)ka & mask;
  for (;;) {
    int occ = atomCAS(&mh[s], -1, x);
    if (occ == -1) break;
    if (cd.keyA(occ) == ka && cd.keyB(occ) == kb) {
      for (;;) {
        if (!candBetter(cd, x, occ)) {
          cd.parflag(x) &= ~CF_ALIVE;
          break;
        }
        const int old = atomCAS(&mh[s], occ, x);
        if (old == occ) {
          cd.parflag(occ) &= ~CF_ALIVE;
          break;
        }
        occ = old;
      }
      break;
    }
    s = (s + 1) & mask;
  }
  w.cslot()[x] = (int)s;
}

FLT_DEV int gxBin(const GxFrame& fr, double score) {
  const float pos = (float)(score - fr.hlo) * fr.hscale;
  return pos >= (float)(kGxBins - 1) ? kGxBins - 1 : (pos > 0.0f ? (int)pos : 0);
}

// one proposal: histogram it (mode 1) or, if it reaches G, materialise and merge it (mode 0).
// x0..x2 = Trie cache of the candidate's lex node (first edge, #edges | #labels, smeared score bits).
FLT_DEV void gxOffer(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const GxFrame& fr,
                     int set, double score, int par, int tok, int word, int lex, int flags, float lmd,
                     float ev, int x0, int x1, int x2, int& cut) {
  if (fr.mode == 1) {
    if (score >= fr.floor) atomAdd(&w.hist()[gxBin(fr, score)], 1);
    return;
  }
  if (!(score >= fr.G)) {
    if (score >= fr.floor) cut = 1;
    return;
  }
  int* sc = gxSc(w, set);
  const int slot = aggInc(&sc[GX_NCAND], cta.tid);
  if (slot >= c.capC) {
    sc[GX_OVF] = 1;
    return;
  }
  putCand(c, w, cur, slot, score, par, tok, word, lex, flags, lmd, ev);
  if (c.lexicon) {
    int* cx = gxCandX(w);
    cx[slot] = x0;
    cx[c.capC + slot] = x1;
    cx[2 * c.capC + slot] = x2;
  }
  gxInsert(c, w, slot);
}

/* ------------------------------------------------------------------ walkers -------------------- */
// r-th set bit of the walker bitmap (hypothesis index of walker rank r), -1 if there is none
FLT_DEV int gxNthWalker(const uint32_t* bits, int words, int r) {
  for (int k = 0; k < words; ++k) {
    const uint32_t v = bits[k];
#if FLT_DEVICE_BUILD
    const int n = __popc(v);
    if (r < n) return (k << 5) + (int)__fns(v, 0, r + 1);
#else
    const int n = __builtin_popcount(v);
    if (r < n) {
      uint32_t u = v;
      for (int z = 0; z < r; ++z) u &= u - 1;
      return (k << 5) + __builtin_ctz(u);
    }
#endif
    r -= n;
  }
  return -1;
}

// static schedule of walker ranks over warp-sized slots: ranks 0,1 a whole warp each; 2..5 sixteen
// lanes; 6..21 eight; the tail four (the host model runs one walker per slot)
FLT_DEV int gxWalkSlots(int nWalk) {
#if FLT_DEVICE_BUILD
  if (nWalk <= 2) return nWalk;
  if (nWalk <= 6) return 2 + ((nWalk - 2 + 1) >> 1);
  if (nWalk <= 22) return 4 + ((nWalk - 6 + 3) >> 2);
  return 8 + ((nWalk - 22 + 7) >> 3);
#else
  return nWalk;
#endif
}
FLT_DEV void gxWalkLane(int slot, int lane, int& r, int& width, int& sub) {
#if FLT_DEVICE_BUILD
  if (slot < 2) {
    r = slot, width = 32;
  } else if (slot < 4) {
    r = 2 + ((slot - 2) << 1) + (lane >> 4), width = 16;
  } else if (slot < 8) {
    r = 6 + ((slot - 4) << 2) + (lane >> 3), width = 8;
  } else {
    r = 22 + ((slot - 8) << 3) + (lane >> 2), width = 4;
  }
  sub = lane & (width - 1);
#else
  (void)lane;
  r = slot, width = 1, sub = 0;
#endif
}

// One slot of walkers: each group of `width` lanes expands one hypothesis over the ranked list.
template <bool LEX>
FLT_DEV void gxWalkSlot(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const FrameIn& f,
                        const GxFrame& fr, int set, int slot, int nWalk, int& cut) {
  const int lane = cta.tid & 31;
  int r, width, sub;
  gxWalkLane(slot, lane, r, width, sub);
  int i = -1;
  if (r < nWalk) i = LEX ? gxNthWalker(gxBits(w, set), (c.K + 31) >> 5, r) : r;
  bool active = i >= 0;
  double si = 0.0;
  int ti = 0, pbi = 0;
  if (active) {
    si = cur.score(i);
    ti = cur.tok(i);
    pbi = cur.pb(i);
  }
  const int M = f.listLen;
  const int* info = LEX ? f.listInfo : nullptr;
  int j = sub;
  for (;;) {
    bool cont = false;
    if (active && j < M) {
      const int n = f.topTok[j];
      if (n >= 0) {
        const float ev = f.topVal[j];
        const double base = si + (double)ev;
        int child = 0, eoff = 0, degLab = 0, msBits = 0;
        float d = 0.0f;
        if (LEX) {
          msBits = info[4 * j];
          child = info[4 * j + 1];
          eoff = info[4 * j + 2];
          degLab = info[4 * j + 3];
          d = bitsF32((uint32_t)msBits) - 0.0f; // LexiconDecoder.cpp:94 with lexMaxScore = 0 at the root
        }
        // the list is ranked by the fp32 key e + bias; the fp64 score may swap near-equal neighbours
        const double lmPart = c.lmWeight * (double)d;
        const double approx = base + lmPart;
        double slack = 0.0;
        if (LEX) {
          const double a = (ev < 0 ? -(double)ev : (double)ev) + (lmPart < 0 ? -lmPart : lmPart);
          slack = 1e-3 + 1e-5 * a;
        }
        cont = fr.mode == 1 || !(approx + slack < fr.G);
        if (cont && (!LEX || child >= 0)) {
          bool ok;
          if (LEX) ok = pbi || n != ti; // LexiconDecoder.cpp:89-90 (CTC)
          else ok = c.ctc ? (n != c.blank && (n != ti || pbi)) : n != ti; // LexiconFreeDecoder.cpp:69-71
          if (n == c.sil && c.silScore > 0) ok = false; // proposed as a special item by every hypothesis
          if (ok) {
            double score = base;
            if (n == c.sil) score += c.silScore;
            score = score + c.lmWeight * (double)d; // ZeroLM / root child: lmWeight * (maxScore - 0)
            gxOffer(cta, c, w, cur, fr, set, score, i, n, -1, LEX ? child : 0, LEX ? 0 : CF_NEW, d, ev, eoff,
                    degLab, msBits, cut);
          }
        }
      }
    }
#if FLT_DEVICE_BUILD
    const unsigned b = __ballot_sync(0xffffffffu, cont);
    const unsigned gm = width == 32 ? 0xffffffffu : (((1u << width) - 1u) << (lane - sub));
    if (active && (b & gm) == 0) {
      if (j - sub + width < M && fr.G > fr.floor) cut = 1; // stopped before the end of the list: the rest is below G
      active = false;
    }
    if (b == 0) break;
#else
    if (!cont) {
      if (active && j + 1 < M && fr.G > fr.floor) cut = 1;
      break;
    }
#endif
    j += width;
  }
}

/* ------------------------------------------------------------------ items ---------------------- */
struct GxItem {
  int kind; // -1 none, 0 edge, 1 stay / repeat, 2 blank, 3 boosted-sil cell
  int q, n, child;
  float ev;
  int4 nd;
};

// word-level LM score of `label` after hypothesis p's state (ZeroLM: 0)
FLT_DEV float gxWordLm(const DecCfg& c, const Beam& cur, int p, int label) {
  if (c.lm.kind == 0) return 0.0f;
  return ngramScore(c.lm, cur.ctx(p), cur.nctx(p), c.lm.usr2lm[label]);
}

// stage 3 of one item: scores and proposals
template <bool LEX>
FLT_DEV void gxItemFinish(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const FrameIn& f,
                          const GxFrame& fr, int set, const GxItem& it, int& cut) {
  if (it.kind < 0) return;
  const int q = it.q;
  const double sq = cur.score(q);
  const int tq = cur.tok(q), pbq = cur.pb(q);
  if (!LEX) {
    if (it.kind == 1) { // repeat (LexiconFreeDecoder.cpp:98-110)
      const int n = tq;
      const bool isRepeat = c.ctc ? (!pbq && n != c.blank) : true;
      if (!(isRepeat && n >= 0 && n < c.N && inTokenSetV(c, f, n, it.ev))) return;
      double score = sq + (double)it.ev;
      if (n == c.sil) score += c.silScore;
      gxOffer(cta, c, w, cur, fr, set, score, q, n, -1, 0, 0, 0.0f, it.ev, 0, 0, 0, cut);
    } else if (it.kind == 2) { // blank (:86-97)
      if (!c.ctc || !inTokenSetV(c, f, c.blank, it.ev)) return;
      double score = sq + (double)it.ev;
      if (c.blank == c.sil) score += c.silScore;
      gxOffer(cta, c, w, cur, fr, set, score, q, c.blank, -1, 0, CF_PB, 0.0f, it.ev, 0, 0, 0, cut);
    } else { // boosted sil as a new token (:69-85)
      const int n = c.sil;
      const bool ok = c.ctc ? (n != c.blank && (n != tq || pbq)) : n != tq;
      if (!ok || !inTokenSetV(c, f, n, it.ev)) return;
      double score = sq + (double)it.ev;
      score += c.silScore;
      score = score + c.lmWeight * (double)0.0f;
      gxOffer(cta, c, w, cur, fr, set, score, q, n, -1, 0, CF_NEW, 0.0f, it.ev, 0, 0, 0, cut);
    }
    return;
  }
  const int lex = cur.lex(q);
  const int eoffQ = cur.xv[q], degLabQ = cur.xv[c.K + q], msQ = cur.xv[2 * c.K + q];
  if (it.kind == 1) { // (2) same node, LexiconDecoder.cpp:167-194
    if (!(!pbq || lex == 0)) return;
    const int n = it.n;
    double score = sq + (double)it.ev;
    if (n == c.sil) score += c.silScore;
    gxOffer(cta, c, w, cur, fr, set, score, q, n, -1, lex, 0, 0.0f, it.ev, eoffQ, degLabQ, msQ, cut);
    return;
  }
  if (it.kind == 2) { // (3) blank, :196-213
    const double score = sq + (double)it.ev;
    gxOffer(cta, c, w, cur, fr, set, score, q, c.blank, -1, lex, CF_PB, 0.0f, it.ev, eoffQ, degLabQ, msQ, cut);
    return;
  }
  if (it.kind == 3) { // boosted sil from the root through root child `sil` (emitSilCell)
    if (it.child < 0 || gxNodeDeg(it.nd.z) == 0) return;
    const int n = c.sil;
    if (!(pbq || n != tq) || n == c.blank) return;
    if (!inTokenSetV(c, f, n, it.ev)) return;
    double score = sq + (double)it.ev;
    score += c.silScore;
    const float d = bitsF32((uint32_t)it.nd.x) - 0.0f;
    score = score + c.lmWeight * (double)d;
    gxOffer(cta, c, w, cur, fr, set, score, q, n, -1, it.child, 0, d, it.ev, it.nd.y, it.nd.z, it.nd.x, cut);
    return;
  }
  // (1) one Trie edge: child node it.child reached by token it.n (LexiconDecoder.cpp:62-141)
  const int n = it.n;
  const float ev = it.ev;
  if (!inTokenSetV(c, f, n, ev)) return;
  const float lexMax = lex == 0 ? 0.0f : bitsF32((uint32_t)msQ);
  double score = sq + (double)ev;
  if (n == c.sil) score += c.silScore;
  const int deg = gxNodeDeg(it.nd.z), nLab = gxNodeLabels(it.nd.z);
  if (lex != 0 && deg > 0 && (pbq || n != tq)) { // (1a); root children with kids are the walkers' cells
    const float d = bitsF32((uint32_t)it.nd.x) - lexMax;
    const double s = score + c.lmWeight * (double)d;
    gxOffer(cta, c, w, cur, fr, set, s, q, n, -1, it.child, 0, d, ev, it.nd.y, it.nd.z, it.nd.x, cut);
  }
  if (nLab > 0 && !(lex == 0 && tq == n)) { // (1b) word ends, :114-141
    // with an n-gram LM: skip the probes when even the LM's best possible score cannot reach G
    if (c.lm.kind != 0 && fr.mode == 0) {
      const float dUp = c.lmUpper - lexMax;
      const double up = score + (c.lmWeight >= 0 ? c.lmWeight * (double)dUp : 0.0) + c.wordScore + 1e-6;
      if (c.lmWeight >= 0 && up < fr.G) {
        if (up >= fr.floor) cut = 1;
        return;
      }
    }
    for (int l = 0; l < nLab; ++l) {
      const int label = nLab == 1 ? it.nd.w : c.trie.labels[it.nd.w + l];
      const float d = gxWordLm(c, cur, q, label) - lexMax;
      const double s = score + c.lmWeight * (double)d + c.wordScore;
      gxOffer(cta, c, w, cur, fr, set, s, q, n, label, 0, CF_NEW, d, ev, 0, 0, 0, cut);
    }
  }
}

/* ------------------------------------------------------------------ phase E --------------------- */
template <bool LEX>
FLT_DEV void gxPhaseE(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const FrameIn& f,
                      const GxFrame& fr, int set, int nH, float eBlank, float eSil) {
  int cut = 0;
  const int warp = cta.tid >> 5, nw = (cta.nthr + 31) >> 5;
  const int nSpec = gxSpecials(c);
  // walkers
  int nWalk = nH;
  if (LEX) {
    nWalk = 0;
    const uint32_t* bits = gxBits(w, set);
    for (int k = 0; k < ((c.K + 31) >> 5); ++k) {
#if FLT_DEVICE_BUILD
      nWalk += __popc(bits[k]);
#else
      nWalk += __builtin_popcount(bits[k]);
#endif
    }
  }
  const int nSlots = gxWalkSlots(nWalk);
  // items
  int nItems;
  int kp2 = 1;
  if (LEX) {
    int nch = gxSc(w, set)[GX_NCHUNK];
    if (nch > c.capChunks) nch = c.capChunks; // overflow was flagged when the table was built
    nItems = nch << kGxChunkLog;
  } else {
    kp2 = c.capP; // pow2 >= K: item x = kind * kp2 + hypothesis
    nItems = nSpec * kp2;
  }
  const int* ch = LEX ? gxChunks(w, set) : nullptr;
  bool walked = false;
  for (int x0 = 0; x0 < nItems || !walked; x0 += kGxSweep * cta.nthr) {
    GxItem it[kGxSweep];
    int2 er[kGxSweep];
    // ---- stage 1: what each item is; edge records / own-token emissions requested
#pragma unroll
    for (int z = 0; z < kGxSweep; ++z) {
      GxItem& I = it[z];
      I.kind = -1;
      I.child = -1;
      I.ev = 0.0f;
      I.n = 0;
      I.q = 0;
      I.nd.x = 0, I.nd.y = 0, I.nd.z = 0, I.nd.w = 0;
      er[z].x = 0, er[z].y = -1;
      const int x = x0 + z * cta.nthr + cta.tid;
      if (x >= nItems) continue;
      if (!LEX) {
        const int kind = x / kp2, q = x & (kp2 - 1);
        if (q >= nH) continue;
        I.q = q;
        if (kind == 0) { // repeat
          I.kind = 1;
          const int n = cur.tok(q);
          I.n = n;
          if (n >= 0 && n < c.N) I.ev = f.e[n];
        } else if (kind == 1) {
          I.kind = 2;
          I.ev = eBlank;
        } else {
          I.kind = 3;
          I.ev = eSil;
        }
        continue;
      }
      const int cdesc = ch[2 * (x >> kGxChunkLog)], cval = ch[2 * (x >> kGxChunkLog) + 1];
      const int sub = x & (kGxChunk - 1);
      if (sub >= cval) continue;
      const int q = cdesc & 0xFFF, k = (int)((unsigned)cdesc >> 12) + sub;
      I.q = q;
      const int lex = cur.lex(q);
      const int edges = lex == 0 ? c.trie.nRootLab : gxNodeDeg(cur.xv[c.K + q]);
      if (k < edges) {
        I.kind = 0;
        er[z] = lex == 0 ? c.trie.rootLabEdge[k] : c.trie.edge[cur.xv[q] + k];
      } else if (k == edges) {
        I.kind = 1;
        const int n = lex == 0 ? c.sil : cur.tok(q);
        I.n = n;
        if (lex == 0) I.ev = eSil;
        else if (n >= 0 && n < c.N) I.ev = f.e[n];
      } else if (k == edges + 1) {
        I.kind = 2;
        I.ev = eBlank;
      } else if (lex == 0) { // boosted sil through the root child of `sil`
        I.kind = 3;
        I.ev = eSil;
        er[z].x = c.sil;
        er[z].y = c.trie.rootChild[c.sil];
      }
    }
    // ---- the walkers run while those loads are in flight (first sweep only)
    if (!walked) {
      for (int s = warp; s < nSlots; s += nw) gxWalkSlot<LEX>(cta, c, w, cur, f, fr, set, s, nWalk, cut);
      walked = true;
    }
    // ---- stage 2: emission and node record of each edge
    if (LEX) {
#pragma unroll
      for (int z = 0; z < kGxSweep; ++z) {
        GxItem& I = it[z];
        if (I.kind == 0 || (I.kind == 3 && er[z].y >= 0)) {
          I.n = er[z].x;
          I.child = er[z].y;
          if (I.kind == 0) I.ev = f.e[I.n];
          I.nd = c.trie.node[I.child];
        }
      }
    }
    // ---- stage 3
#pragma unroll
    for (int z = 0; z < kGxSweep; ++z) gxItemFinish<LEX>(cta, c, w, cur, f, fr, set, it[z], cut);
  }
  if (fr.mode == 0) {
#if FLT_DEVICE_BUILD
    if (__any_sync(0xffffffffu, cut) && (cta.tid & 31) == 0) gxSc(w, set)[GX_CUT] = 1;
#else
    if (cut) gxSc(w, set)[GX_CUT] = 1;
#endif
  }
}

/* ------------------------------------------------------------------ the frame step ------------- */
// highest bin b with at least `want` proposals in bins >= b (0 if there are fewer in total); *kept =
// proposals in bins >= b. One warp; the histogram is left intact.
FLT_DEV void gxFindCut(const Cta& cta, const Ws& w, int want, int* out) {
  const int* hist = w.hist();
#if FLT_DEVICE_BUILD
  if (cta.tid < 32) {
    const int lane = cta.tid;
    int h[8], part = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      h[k] = hist[255 - (lane * 8 + k)];
      part += h[k];
    }
    int incl = part;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0 && total < want) {
      out[0] = 0;
      out[1] = total;
    }
    const int excl = incl - part;
    if (excl < want && incl >= want) {
      int cum = excl;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (cum < want && cum + h[k] >= want) {
          out[0] = 255 - (lane * 8 + k);
          out[1] = cum + h[k];
        }
        cum += h[k];
      }
    }
  }
#else
  if (cta.tid == 0) {
    int cum = 0;
    out[0] = 0;
    for (int b = 255; b >= 0; --b) {
      cum += hist[b];
      if (cum >= want) {
        out[0] = b;
        break;
      }
    }
    out[1] = cum;
  }
#endif
}

// C: compact the group representatives, empty the merge table, find the best score
FLT_DEV void gxPhaseC(const Cta& cta, const DecCfg& c, const Ws& w, int set, int nCand) {
  const Cand cd = w.cand();
  int* sc = gxSc(w, set);
  int* rep = w.rep();
  u64* rkey = w.rkey();
  int* mh = w.mh();
  const int* cslot = w.cslot();
  u64 best = 0;
  for (int x = cta.tid; x < nCand; x += cta.nthr) {
    if (!(cd.parflag(x) & CF_ALIVE)) continue;
    const u64 k = orderedKey64(cd.score(x));
    mh[cslot[x]] = -1;
    const int r = aggInc(&sc[GX_NREP], cta.tid);
    rep[r] = x;
    rkey[r] = k; // keyA storage: every insert finished at the barrier before this phase
    best = k > best ? k : best;
  }
#if FLT_DEVICE_BUILD
  {
    unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(best >> 32));
    unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(best >> 32) == hi ? (unsigned)best : 0u);
    best = ((u64)hi << 32) | lo;
    if ((cta.tid & 31) == 0 && best) atomicMax((u64*)(w.base + c.lay.gxBest), best);
  }
#else
  {
    u64* gb = (u64*)(w.base + c.lay.gxBest);
    if (best > *gb) *gb = best;
  }
#endif
}

// RF: rank the representatives by counting; the thread that finds rank q < K writes hypothesis q of
// the new beam (phaseFinalize of beam_core.h, one hypothesis per thread) and registers it for the
// next frame's E.
FLT_DEV void gxPhaseRF(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const Beam& nxt,
                       const FrameIn& f, int set, int nRep) {
  const Cand cd = w.cand();
  const int K = c.K;
  int* sc = w.sc();
  const int* rep = w.rep();
  const u64* rkey = w.rkey();
  const u64 bestKey = *(const u64*)(w.base + c.lay.gxBest);
  // candidatesBestScore_ - beamThreshold (Utils.h:161-165; a max-merge keeps the same groups when the
  // filter runs after the merge)
  const double thrScore = keyToDouble(bestKey) - c.beamThreshold;
  int lg = 0;
  while (lg < 5 && ((long long)nRep << (lg + 1)) <= cta.nthr) ++lg;
  const int parts = 1 << lg;
  const int slice = (nRep + parts - 1) >> lg;
  for (int base = 0; base < (nRep << lg); base += cta.nthr) {
    const int t = base + cta.tid;
    const int a = t >> lg, part = t & (parts - 1);
    const bool valid = a < nRep;
    int cnt = 0, eq = 0;
    u64 ka = 0;
    if (valid) {
      ka = rkey[a];
      const int lo = part * slice, hi = lo + slice < nRep ? lo + slice : nRep;
#pragma unroll 4
      for (int b = lo; b < hi; ++b) {
        const u64 kb = rkey[b];
        cnt += kb > ka ? 1 : 0;
        eq |= kb == ka ? (b != a ? 1 : 0) : 0;
      }
    }
#if FLT_DEVICE_BUILD
    for (int o = 1; o < parts; o <<= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      eq |= __shfl_xor_sync(0xffffffffu, eq, o);
    }
#endif
    if (!(valid && part == 0)) continue;
    const int x = rep[a];
    if (eq) { // another group with exactly this score: settle by the deterministic order (rare)
      for (int b = 0; b < nRep; ++b)
        if (b != a && rkey[b] == ka && candBetter(cd, rep[b], x)) ++cnt;
    }
    const int q = cnt;
    if (q >= K) continue;
    const double score = cd.score(x);
    if (!(score >= thrScore)) continue;
    atomMax(&sc[SC_NH], q + 1);
    const int p = cd.par(x);
    const int fl = cd.flags(x);
    const int n = cd.tok(x);
    nxt.score(q) = score;
    nxt.am(q) = (fl & CF_FINISH) ? cur.am(p) : cur.am(p) + amOf(c, f, cd.ce(x), n, cur.tok(p));
    nxt.lm(q) = cur.lm(p) + (double)cd.lmd(x);
    const int lexNew = c.lexicon ? cd.lex(x) : 0;
    nxt.lex(q) = lexNew;
    nxt.tok(q) = n;
    nxt.pb(q) = (fl & CF_PB) ? 1 : 0;
    if (fl & CF_NEW) {
      const int lab = candLabel(c, cd, x);
      fpChild(cur.fpA(p), cur.fpB(p), lab, nxt.fpA(q), nxt.fpB(q));
      if (c.lm.kind) {
        const int wlm = lab < 0 ? c.lm.eos : c.lm.usr2lm[lab];
        nxt.nctx(q) = ngramAdvanceCtx(c.lm, cur.ctx(p), cur.nctx(p), wlm, nxt.ctx(q));
      }
    } else {
      nxt.fpA(q) = cur.fpA(p);
      nxt.fpB(q) = cur.fpB(p);
      if (c.lm.kind) {
        const int nc = cur.nctx(p);
        nxt.nctx(q) = nc;
        for (int k = 0; k < nc; ++k) nxt.ctx(q)[k] = cur.ctx(p)[k];
      }
    }
    f.hParent[q] = p;
    f.hTok[q] = n;
    if (f.hWord) f.hWord[q] = c.lexicon ? cd.word(x) : -1;
    const int an = skipCarry(cur, f.hRow, p);
    nxt.anc(q) = an;
    if (f.hSkip) f.hSkip[q] = an;
    if (f.hScore) {
      f.hScore[3 * q] = score;
      f.hScore[3 * q + 1] = nxt.am(q);
      f.hScore[3 * q + 2] = nxt.lm(q);
    }
    if (c.lexicon && !(fl & CF_FINISH)) {
      const int* cx = gxCandX(w);
      const int x0 = cx[x], x1 = cx[c.capC + x], x2 = cx[2 * c.capC + x];
      gxSetNode(c, nxt, q, lexNew, x0, x1, x2);
      gxRegister(c, w, set ^ 1, q, lexNew, x1);
    }
  }
}

// reference level of a frame: the best proposal of the best hypothesis (blank, or its best list cell)
template <bool LEX>
FLT_DEV double gxRefLevel(const DecCfg& c, const Beam& cur, const FrameIn& f, float eBlank, bool blankOk) {
  const double s0 = cur.score(0);
  double ref = negInf();
  if (blankOk) ref = s0 + (double)eBlank;
  if (f.listLen > 0 && f.topTok[0] >= 0) {
    double v = s0 + (double)f.topVal[0];
    if (LEX) v = v + c.lmWeight * (double)bitsF32((uint32_t)f.listInfo[0]);
    ref = v > ref ? v : ref;
  }
  return ref;
}

// One frame: cur (beam `set`) -> nxt (beam `set ^ 1`). All threads call this with identical arguments.
template <bool LEX>
FLT_DEV void gxFrameStep(const Cta& cta, const DecCfg& c, const Ws& w, int set, const FrameIn& f, int* status,
                         unsigned long long* stats, GxCarry& g) {
  int* sc = w.sc();
  const int nH = sc[SC_NH];
  if (nH == 0) return; // the beam died (Utils.h:155-158)
  const Beam cur = w.beam(set), nxt = w.beam(set ^ 1);
  int* gs = gxSc(w, set);
  LfPhaseClock pc;
  pc.start(cta, stats);
  const int K = c.K;
  // this frame's e[blank], e[sil]
  float eBlank = 0.0f, eSil = 0.0f;
  if (f.specReady) {
    eBlank = f.eBlank;
    eSil = f.eSil;
  } else if (g.eValid) {
    eBlank = g.eBlank;
    eSil = g.eSil;
  } else {
    if (c.ctc) eBlank = f.e[c.blank];
    eSil = f.e[c.sil];
  }
  if (f.eNext && !f.specReady) { // the next frame's, requested now
    g.eBlank = c.ctc ? f.eNext[c.blank] : 0.0f;
    g.eSil = f.eNext[c.sil];
  }
  g.eValid = f.eNext != nullptr && !f.specReady;
  const bool blankOk = c.ctc && (LEX || inTokenSetV(c, f, c.blank, eBlank));
  const double s0 = cur.score(0);
  GxFrame fr;
  fr.mode = 0;
  fr.hlo = 0.0;
  fr.hscale = 0.0f;
  // the best hypothesis' blank candidate always exists: the frame's best is at least that, and the
  // reference drops everything below it minus beamThreshold
  fr.floor = negInf();
  if (blankOk) {
    double sb = s0 + (double)eBlank;
    if (!LEX && c.blank == c.sil) sb += c.silScore;
    fr.floor = sb - c.beamThreshold;
  }
  const double ref = gxRefLevel<LEX>(c, cur, f, eBlank, blankOk);
  fr.G = ref - g.spread; // -inf while the spread is unknown
  if (!(fr.G == fr.G)) fr.G = negInf(); // inf - inf
  if (fr.G < fr.floor) fr.G = fr.floor;

  int nCand = 0, nRep = 0;
  int want = 2 * K + 32;
  bool histReady = false, last = false;
  int guessMiss = 0;
  for (int attempt = 0;; ++attempt) {
    gxPhaseE<LEX>(cta, c, w, cur, f, fr, set, nH, eBlank, eSil);
    cta.sync(); // ---- A
    if (attempt == 0) pc.mark(0);
    nCand = gs[GX_NCAND];
    const int ovf = gs[GX_OVF];
    if (nCand > c.capC) nCand = c.capC;
    gxPhaseC(cta, c, w, set, nCand);
    cta.sync(); // ---- B
    if (attempt == 0) pc.mark(1);
    nRep = gs[GX_NREP];
    const int cutAny = gs[GX_CUT];
    const bool miss = nRep < K && cutAny; // fewer than K groups although proposals were left out
    if ((!ovf && !miss) || last) break;
    // ---- exact redo: a histogram of every proposal picks G (the kept set is verified again)
    guessMiss = 1;
    cta.sync(); // everyone has read the scalars
    if (cta.tid == 0) {
      gs[GX_NCAND] = 0;
      gs[GX_OVF] = 0;
      gs[GX_CUT] = 0;
      gs[GX_NREP] = 0;
      *(u64*)(w.base + c.lay.gxBest) = 0;
    }
    if (!histReady) {
      double hi = s0;
      if (c.silScore > 0) hi += c.silScore;
      if (c.wordScore > 0) hi += c.wordScore;
      double span = 2.0 * (s0 - cur.score(nH - 1)) + 16.0;
      span = span > 128.0 ? 128.0 : span;
      if (c.beamThreshold + 4.0 < span) span = c.beamThreshold + 4.0;
      fr.hlo = hi - span;
      fr.hscale = (float)kGxBins / (float)span;
      fr.mode = 1;
      gxPhaseE<LEX>(cta, c, w, cur, f, fr, set, nH, eBlank, eSil);
      fr.mode = 0;
      histReady = true;
    } else if (ovf) {
      // the histogram's own choice overflowed (a crowded cut bin): aim lower; at K proposals nothing is
      // left to try — the host grows the capacity and redoes the batch
      if (want <= K + 16) last = true;
      want = want / 2 > K + 16 ? want / 2 : K + 16;
    } else {
      want = want * 2;
    }
    if (attempt >= 12) last = true;
    cta.sync();
    gxFindCut(cta, w, want, sc + SC_GXCUTBIN);
    cta.sync();
    const int cutBin = sc[SC_GXCUTBIN];
    if (cutBin <= 0) fr.G = fr.floor; // fewer proposals than wanted: take everything
    else fr.G = fr.hlo + ((double)cutBin - 0.02) / (double)fr.hscale;
    if (fr.G < fr.floor) fr.G = fr.floor;
    if (last && cta.tid == 0) *status |= 1; // decode on with what fits; the host discards the result
  }
  if (histReady) {
    for (int b = cta.tid; b < kGxBins; b += cta.nthr) w.hist()[b] = 0;
  }
  // the new beam's tables are built below: reset their counters, and this frame's scalars' twin set
  if (cta.tid == 0) {
    int* gn = gxSc(w, set ^ 1);
    gn[GX_NCAND] = 0;
    gn[GX_OVF] = 0;
    gn[GX_CUT] = 0;
    gn[GX_NREP] = 0;
    sc[SC_NH] = 0;
  }
  // (the twin set's chunk counter / walker bitmap were cleared in the previous frame's tail, below)
#if FLT_DEVICE_BUILD
  if (stats && cta.tid == 0) {
    atomicAdd(stats + 0, 1ull);
    atomicAdd(stats + 1, (unsigned long long)nCand);
    atomicAdd(stats + 2, (unsigned long long)nRep);
    atomicAdd(stats + 3, (unsigned long long)(nRep < K ? nRep : K));
    if (guessMiss) atomicAdd(stats + 12, 1ull);
    stats[30] = 1ull; // phase names of this step for the host
  }
#endif
  cta.sync(); // ---- (scalars reset before RF's atomics)
  gxPhaseRF(cta, c, w, cur, nxt, f, set, nRep);
  cta.sync(); // ---- C
  pc.mark(2);
  // tail: this frame's tables are dead — clear them for the frame after next; reset the best key
  if (LEX) {
    if (cta.tid == 0) gs[GX_NCHUNK] = 0;
    for (int i = cta.tid; i < ((K + 31) >> 5); i += cta.nthr) gxBits(w, set)[i] = 0;
  }
  if (cta.tid == 0) {
    *(u64*)(w.base + c.lay.gxBest) = 0;
    if (LEX && gxSc(w, set ^ 1)[GX_NCHUNK] > c.capChunks) *status |= 1;
  }
  // the guess for the next frame
  const int nHn = sc[SC_NH];
  if (nRep >= K && nHn == K) {
    const double cutScore = nxt.score(K - 1);
    const double exact = ref - cutScore;
    if (nRep < K + (K >> 2) + 8) g.mu = g.mu * 1.5f < 4.0f ? g.mu * 1.5f : 4.0f;
    else if (nRep > 2 * K + (K >> 1)) g.mu = g.mu * 0.8f > 0.04f ? g.mu * 0.8f : 0.04f;
    if (guessMiss) g.mu = g.mu * 1.5f < 4.0f ? g.mu * 1.5f : 4.0f;
    g.spread = exact > 0 ? exact * (1.0 + (double)g.mu) + 0.05 : 0.05;
  } else {
    g.spread = bitsF64(0x7FF0000000000000ull); // the beam is not full: nothing to cut against
  }
  if (f.hCount && cta.tid == 0) *f.hCount = nHn;
}

// decodeEnd (LexiconFreeDecoder.cpp:127-158, LexiconDecoder.cpp:231-274) through the same phases:
// one finish candidate per hypothesis (only those at the Trie root when any exists), merged, ranked.
template <bool LEX>
FLT_DEV void gxFinish(const Cta& cta, const DecCfg& c, const Ws& w, int set, const FrameIn& f) {
  int* sc = w.sc();
  const int nH = sc[SC_NH];
  if (nH == 0) return;
  const Beam cur = w.beam(set), nxt = w.beam(set ^ 1);
  int* gs = gxSc(w, set);
  if (cta.tid == 0) sc[SC_NICE] = 0;
  cta.sync();
  if (LEX) {
    int nice = 0;
    for (int i = cta.tid; i < nH; i += cta.nthr) nice |= cur.lex(i) == 0;
#if FLT_DEVICE_BUILD
    if (__any_sync(0xffffffffu, nice) && (cta.tid & 31) == 0) sc[SC_NICE] = 1;
#else
    if (nice) sc[SC_NICE] = 1;
#endif
    cta.sync();
  }
  const bool nice = LEX && sc[SC_NICE] != 0;
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    w.cand().parflag(i) = 0;
    if (nice && cur.lex(i) != 0) continue;
    float ls = 0.0f;
    int flags = CF_FINISH;
    if (c.lm.kind) { // KenLM::finish: score </s>, state = child(-1); ZeroLM: same state, 0
      ls = ngramScore(c.lm, cur.ctx(i), cur.nctx(i), c.lm.eos);
      flags |= CF_NEW;
    }
    const double score = cur.score(i) + c.lmWeight * (double)ls;
    putCand(c, w, cur, i, score, i, c.sil, -1, LEX ? cur.lex(i) : 0, flags, ls, 0.0f);
    gxInsert(c, w, i);
  }
  if (cta.tid == 0) {
    gs[GX_NREP] = 0;
    *(u64*)(w.base + c.lay.gxBest) = 0;
  }
  cta.sync();
  gxPhaseC(cta, c, w, set, nH);
  cta.sync();
  const int nRep = gs[GX_NREP];
  if (cta.tid == 0) sc[SC_NH] = 0;
  cta.sync();
  gxPhaseRF(cta, c, w, cur, nxt, f, set, nRep);
  cta.sync();
  if (cta.tid == 0) {
    gs[GX_NREP] = 0;
    *(u64*)(w.base + c.lay.gxBest) = 0;
  }
  if (f.hCount && cta.tid == 0) *f.hCount = sc[SC_NH];
}

// once per CTA: tables that persist over its utterances
FLT_DEV void gxInitWorkspace(const Cta& cta, const DecCfg& c, const Ws& w) {
  for (int i = cta.tid; i < c.capH; i += cta.nthr) w.mh()[i] = -1;
  for (int i = cta.tid; i < kGxBins; i += cta.nthr) w.hist()[i] = 0;
  for (int i = cta.tid; i < 2 * ((c.K + 31) >> 5); i += cta.nthr) gxBits(w, 0)[i] = 0;
  if (cta.tid == 0) {
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < GX_SET; ++k) gxSc(w, s)[k] = 0;
    *(u64*)(w.base + c.lay.gxBest) = 0;
  }
}

// per utterance, after the seed (or the restored beam) is in beam 0
FLT_DEV void gxBeginUtterance(const Cta& cta, const DecCfg& c, const Ws& w, int nH) {
  if (cta.tid == 0) {
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < GX_SET; ++k) gxSc(w, s)[k] = 0;
  }
  for (int i = cta.tid; i < 2 * ((c.K + 31) >> 5); i += cta.nthr) gxBits(w, 0)[i] = 0;
  cta.sync();
  gxRebuild(cta, c, w, 0, nH);
}

// Lexicon decoder: what the walkers need to know about each entry of a frame's ranked list — the root
// child reached by the token, its smeared score and Trie cache — gathered once per list entry.
// (The fused kernel's producer fills this; the two-kernel path gathers it a frame ahead.)
FLT_DEV void gxListEntry(const DecCfg& c, int n, int* out4) {
  int4 r;
  r.x = 0, r.y = -1, r.z = 0, r.w = 0;
  if (n >= 0) {
    const int child = c.trie.rootChild[n];
    if (child >= 0) {
      const int4 nd = c.trie.node[child];
      if (gxNodeDeg(nd.z) > 0) r.x = nd.x, r.y = child, r.z = nd.y, r.w = nd.z;
    }
  }
  out4[0] = r.x, out4[1] = r.y, out4[2] = r.z, out4[3] = r.w;
}

/* ------------------------------------------------------------------ whole-utterance driver ---- */
// Two-kernel path (token lists from flt_k_topm in HBM): one CTA decodes utterances bid, bid+nblk, ...
// The frame's list lives in the workspace (double-buffered); the NEXT frame's list — and, for the
// lexicon decoder, the root child of each of its entries — is pulled into registers at the start of
// a frame and stored after it, so those gathers never sit on the frame's critical path.
template <bool LEX>
FLT_DEV void gxLoadListDirect(const Cta& cta, const DecCfg& c, const Ws& w, const BatchArgs& a, long long row,
                              int buf) {
  for (int j = cta.tid; j < c.M; j += cta.nthr) {
    const int n = a.topTok[row * c.M + j];
    w.listTok(buf)[j] = n;
    w.listVal(buf)[j] = a.topVal[row * c.M + j];
    if (LEX && j < c.Mwide) gxListEntry(c, n, gxListInfo(w, buf) + 4 * j);
  }
}

template <bool LEX>
FLT_DEV void gxDecodeCta(const Cta& cta, const DecCfg& c, const BatchArgs& a, char* base) {
  const Ws w{base, &c};
  const int K = c.K;
  gxInitWorkspace(cta, c, w);
  for (int b = cta.bid; b < a.B; b += cta.nblk) {
    const int len = a.lengths ? a.lengths[b] : a.T;
    int set = 0;
    cta.sync(); // previous utterance fully retired
    if (a.streamBeam && a.streamRestore) streamRestoreBeam(cta, c, w, a);
    else if (cta.tid == 0) seedUtterance(c, w, a, b);
    cta.sync();
    gxBeginUtterance(cta, c, w, w.sc()[SC_NH]);
    GxCarry g = gxCarryInit();
    const long long row0 = (long long)b * a.T;
    if (len > 0) gxLoadListDirect<LEX>(cta, c, w, a, row0, 0);
    cta.sync();
    for (int t = 0; t < len; ++t) {
      const long long row = row0 + t;
      // next frame's list -> registers (two entries per thread; longer lists are copied after the step)
      int pfTok[2] = {-1, -1};
      float pfVal[2] = {0.0f, 0.0f};
      int pfInfo[2][4] = {{0, -1, 0, 0}, {0, -1, 0, 0}};
      const bool pf = t + 1 < len;
      if (pf) {
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          const int j = cta.tid + z * cta.nthr;
          if (j < c.M) {
            pfTok[z] = a.topTok[(row + 1) * c.M + j];
            pfVal[z] = a.topVal[(row + 1) * c.M + j];
          }
        }
        if (LEX) {
#pragma unroll
          for (int z = 0; z < 2; ++z) {
            const int j = cta.tid + z * cta.nthr;
            if (j < c.Mwide) gxListEntry(c, pfTok[z], pfInfo[z]);
          }
        }
      }
      FrameIn f;
      f.e = a.emis + row * c.N;
      f.topTok = w.listTok(t & 1);
      f.topVal = w.listVal(t & 1);
      f.listLen = c.M;
      f.thrVal = a.thrVal ? a.thrVal[row] : 0.0f;
      f.first = a.streamFrame0 + t == 0;
      f.listIsSet = !c.lexicon;
      f.specReady = 0;
      f.eBlank = 0.0f;
      f.eSil = 0.0f;
      f.listInfo = LEX ? gxListInfo(w, t & 1) : nullptr;
      f.hScore = a.hScore ? a.hScore + (long long)(t + 1) * K * 3 : nullptr;
      f.hCount = a.hCount ? a.hCount + (t + 1) : nullptr;
      f.eNext = t + 1 < len ? f.e + c.N : nullptr;
      const long long h = ((long long)b * (a.T + 2) + (t + 1)) * K;
      f.hParent = a.hParent + h;
      f.hTok = a.hTok + h;
      f.hWord = a.hWord ? a.hWord + h : nullptr;
      f.hRow = t + 1;
      f.hSkip = ((t + 1) & (kCpRows - 1)) == 0 ? a.hSkip + ((long long)b * a.nCp + ((t + 1) >> kCpShift)) * K : nullptr;
#if FLT_DEVICE_BUILD
      // the select kernel streamed this row long ago; pull the NEXT frame's row into L2 while this one
      // is processed (the lexicon step gathers one emission per Trie edge)
      if (LEX && f.eNext) {
        const char* nx = (const char*)f.eNext;
        for (int off = cta.tid * 128; off < c.N * 4; off += cta.nthr * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + off));
      }
#endif
      gxFrameStep<LEX>(cta, c, w, set, f, a.status + b, a.stats, g);
      if (pf) {
        const int nb = (t + 1) & 1;
#if FLT_DEVICE_BUILD
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          const int j = cta.tid + z * cta.nthr;
          if (j < c.M) {
            w.listTok(nb)[j] = pfTok[z];
            w.listVal(nb)[j] = pfVal[z];
            if (LEX && j < c.Mwide) {
              int* o = gxListInfo(w, nb) + 4 * j;
              o[0] = pfInfo[z][0], o[1] = pfInfo[z][1], o[2] = pfInfo[z][2], o[3] = pfInfo[z][3];
            }
          }
        }
        for (int j = cta.tid + 2 * cta.nthr; j < c.M; j += cta.nthr) {
          const int n = a.topTok[(row + 1) * c.M + j];
          w.listTok(nb)[j] = n;
          w.listVal(nb)[j] = a.topVal[(row + 1) * c.M + j];
          if (LEX && j < c.Mwide) gxListEntry(c, n, gxListInfo(w, nb) + 4 * j);
        }
#else
        (void)pfTok, (void)pfVal, (void)pfInfo;
        gxLoadListDirect<LEX>(cta, c, w, a, row + 1, nb);
#endif
      }
      if (w.sc()[SC_NH] == 0) break;
      set ^= 1;
      cta.sync();
    }
    if (a.streamBeam && a.streamNoFinish) { // decodeStep chunk: keep the beam for the next launch
      cta.sync();
      streamSaveBeam(cta, c, w, a, set);
      continue;
    }
    int nFin = 0;
    if (w.sc()[SC_NH] != 0) {
      const FrameIn f = finishFrameIn(c, a, b, len);
      gxFinish<LEX>(cta, c, w, set, f);
      set ^= 1;
      nFin = w.sc()[SC_NH];
    }
    cta.sync();
    writeFinals(cta, c, w, a, b, set, nFin);
  }
}

} // namespace flt
