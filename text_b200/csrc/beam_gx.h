// beam_gx.h — the two-pass frame step ("gx") of both decoders: histogram every proposal, cut, materialise
// only what can still reach the beam.
//
// Replaces, for one frame, LexiconDecoder::decodeStep / LexiconFreeDecoder::decodeStep's expansion
// (decoder/LexiconDecoder.cpp:54-215, decoder/LexiconFreeDecoder.cpp:53-112) and candidatesStore
// (decoder/Utils.h:146-225) in the max-merge, word-level-LM, CTC (lexicon) / CTC+ASG (lexicon-free)
// configurations — the ones BASELINE.json's configs use. Everything else stays on beam_core.h.
//
// Exactness. The reference materialises every candidate of a frame, then keeps the K best merge groups.
// A group's score is the maximum of its members (Utils.h:176-198), so for ANY threshold the K best
// groups are found exactly from the candidates at or above it alone, provided those candidates form at
// least K groups (a group with a member above the threshold has its best member there; a group
// without scores below every kept group). The threshold used here is a bin edge of a monotone map
// score -> 256 bins, chosen from an exact histogram of the frame's proposals; the kept set is verified
// to hold >= K groups after the merge (else the cut is lowered and the second pass repeated).
//
//   E1  every proposal is scored and histogrammed; nothing is stored except one byte per work item
//       (the best bin its proposals reached)
//   --  every warp finds the cut bin from the histogram for itself (no single-warp phase, no barrier)
//   E2  the work items whose byte reaches the cut are evaluated again and their proposals at or above
//       the cut are materialised (record + 128-bit merge key + ordered score key)
//   M   one candidate per thread: compare-and-swap into the CTA-private merge table on the
//       (LM state, lex node, token, prevBlank) key; the loser's score key is zeroed (max-merge)
//   RF  every live candidate ranks itself by counting larger score keys; the thread that finds rank
//       q < K writes hypothesis q of the new beam directly: threshold against the best (Utils.h:161-165),
//       LM-state fingerprint / n-gram context, back-pointer record, skip pointer — and registers the
//       hypothesis' work items for the next frame.
// Four CTA barriers per frame, no row grouping, no prefix sums, no binary searches, no compaction.
//
// Work distribution:
//   * "walkers" = hypotheses that expand over the frame's RANKED token list (all hypotheses of the
//     lexicon-free decoder; hypotheses at the Trie root in the lexicon decoder, ranked by
//     e[n] + lmWeight * smeared score of root child n). The score is monotone along the list, so eight
//     lanes per walker go down it in chunks and stop at the first chunk entirely below the bound.
//   * "items" = everything enumerated directly: the Trie edges of non-root hypotheses, root children
//     that carry labels (single-token words), stay / blank / repeat. When a hypothesis is created
//     (RF of the previous frame) its thread reserves the hypothesis' items in a chunk table (8 items
//     per descriptor), so item x maps to (hypothesis, edge) with two shared-memory loads. Trie
//     edges are {token, child} pairs (one coalesced 8-byte load), nodes are 16-byte records
//     {smeared score, first edge, #edges | #labels, first label}: two dependent L2 round trips per
//     edge, issued for all items of a sweep before the walkers run, consumed after.
//   * n-gram word LM: a word end is probed only if its score with the LM's best possible score still
//     reaches the bound.
#pragma once
#include "beam_core.h"
#include "beam_lf.h"

namespace flt {

constexpr int kGxChunkLog = 3, kGxChunk = 1 << kGxChunkLog; // items per chunk descriptor
constexpr int kGxBins = 256;
constexpr int kGxSweep = 2;     // items per thread and sweep (loads of a sweep are in flight together)
constexpr int kGxWalkLanes = 8; // lanes per walker (device)

enum { // scalars in two sets, index = beam parity p: NCAND / OVF / NDEAD belong to the frame that READS
       // beam p; NCHUNK / NWALK / NH describe beam p itself (item chunks, walkers, hypotheses)
  GX_NCAND = 0, GX_OVF, GX_NDEAD, GX_NCHUNK, GX_NWALK, GX_NH, GX_SET = 8
};

// Per-utterance state carried from frame to frame; uniform over the CTA's threads (registers).
struct GxCarry {
  double cutPrev, D;  // last cut and its offset to (previous cut + emission level): bound of the walkers'
  int have;           // histogram pass in the lexicon decoder (0 = nothing known, 2 = both)
  float span;         // score range the histogram covers below its top
  float eBlank, eSil; // e[blank], e[sil] of the NEXT frame, loaded while the current one retires
  int eValid;
};
FLT_DEV GxCarry gxCarryInit() { return GxCarry{0.0, 0.0, 0, 32.0f, 0.0f, 0.0f, 0}; }

struct GxFrame { // uniform per pass
  int mode;      // 1 = histogram pass, 0 = materialise
  double floor;  // the reference's own filter: candidates below never survive (Utils.h:161-165)
  double hlo;    // histogram map: bin = (score - hlo) * hscale, clamped to [0, 255]
  float hscale;
  int cutBin;    // pass 0: proposals in bins below are left out
  double stop;   // walkers / LM probes: nothing below this can matter in this pass
};

/* ------------------------------------------------------------------ workspace views ---------- */
FLT_DEV int* gxSc(const Ws& w, int set) { return w.sc() + SC_GX + set * GX_SET; }
FLT_DEV int* gxCandX(const Ws& w) { return (int*)(w.base + w.c->lay.gxCandX); }      // [3][capC]
FLT_DEV int* gxChunks(const Ws& w, int set) {                                          // [2][capChunks] x int2
  return (int*)(w.base + w.c->lay.gxChunk) + (size_t)set * 2 * w.c->capChunks;
}
FLT_DEV int* gxWalkList(const Ws& w, int set) { // [2][K] hypotheses at the Trie root of each beam
  return (int*)(w.base + w.c->lay.gxBits) + (size_t)set * w.c->K;
}
FLT_DEV int* gxListInfo(const Ws& w, int buf) { // [2][Mwide] x int4 {ms bits, child, eoff, degLab}
  return (int*)(w.base + w.c->lay.gxList) + (size_t)buf * 4 * w.c->Mwide;
}
FLT_DEV unsigned char* gxStash(const Ws& w) { return (unsigned char*)(w.base + w.c->lay.gxStash); }
FLT_DEV u64* gxSkey(const Ws& w) { return (u64*)(w.base + w.c->lay.skey); } // [capC] ordered score keys, 0 = dead

FLT_DEV int gxSpecials(const DecCfg& c) { // directly enumerated specials per hypothesis:
  return c.silScore > 0 ? 3 : 2;          // stay / repeat, blank, boosted-sil cell
}

// node record of the Trie: {smeared score bits, first edge, #edges | #labels << 24, first label or label offset}
FLT_DEV int gxNodeDeg(int degLab) { return degLab & 0xFFFFFF; }
FLT_DEV int gxNodeLabels(int degLab) { return (unsigned)degLab >> 24; }

/* ------------------------------------------------------------------ hypothesis registration -- */
// Called by the thread that creates hypothesis q of beam `set`: reserve its items in the chunk table
// (lexicon) and list it as a walker if it sits at the Trie root.
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
  if (lex == 0) gxWalkList(w, set)[atomAdd(&sc[GX_NWALK], 1)] = q;
}

// per-hypothesis Trie cache of a beam entry (lexicon): first edge, #edges | #labels, smeared score
FLT_DEV void gxSetNode(const Beam& b, int q, int eoff, int degLab, int msBits) {
  b.xv[q] = eoff;
  b.xv[b.K + q] = degLab;
  b.xv[2 * b.K + q] = msBits;
}

// Rebuild the tables of beam `set` from its hypotheses (seed of an utterance, restored online beam).
FLT_DEV void gxRebuild(const Cta& cta, const DecCfg& c, const Ws& w, int set, int nH) {
  if (!c.lexicon) return;
  if (cta.tid == 0) {
    gxSc(w, set)[GX_NCHUNK] = 0;
    gxSc(w, set)[GX_NWALK] = 0;
  }
  cta.sync();
  const Beam b = w.beam(set);
  for (int q = cta.tid; q < nH; q += cta.nthr) {
    const int lex = b.lex(q);
    int4 nd;
    nd.x = 0, nd.y = 0, nd.z = 0, nd.w = 0;
    if (lex != 0) nd = c.trie.node[lex];
    gxSetNode(b, q, nd.y, nd.z, lex == 0 ? 0 : nd.x);
    gxRegister(c, w, set, q, lex, nd.z);
  }
  cta.sync();
}

/* ------------------------------------------------------------------ proposals ------------------ */
FLT_DEV int gxBin(const GxFrame& fr, double score) {
  const float pos = (float)(score - fr.hlo) * fr.hscale;
  return pos >= (float)(kGxBins - 1) ? kGxBins - 1 : (pos > 0.0f ? (int)pos : 0);
}

// One proposal. Pass 1: histogram it; returns its bin (-1 = below the reference's own filter). Pass 0:
// materialise it if its bin reaches the cut — record, 128-bit merge key, ordered score key.
// x0..x2 = Trie cache of the candidate's lex node (first edge, #edges | #labels, smeared score bits).
FLT_DEV int gxOffer(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const GxFrame& fr, int set,
                    double score, int par, int tok, int word, int lex, int flags, float lmd, float ev, int x0,
                    int x1, int x2) {
  if (!(score >= fr.floor)) return -1;
  const int bin = gxBin(fr, score);
  if (fr.mode == 1) {
    atomAdd(&w.hist()[bin], 1);
    return bin;
  }
  if (bin < fr.cutBin) return bin;
  int* sc = gxSc(w, set);
  const int slot = aggInc(&sc[GX_NCAND], cta.tid);
  if (slot >= c.capC) {
    sc[GX_OVF] = 1;
    return bin;
  }
  const Cand cd = w.cand();
  cd.score(slot) = score;
  cd.parflag(slot) = (par << 4) | flags | CF_ALIVE;
  cd.tok(slot) = tok;
  cd.ce(slot) = ev;
  if (c.lexicon) {
    cd.word(slot) = word;
    cd.lex(slot) = lex;
    cd.lmd(slot) = lmd;
    int* cx = gxCandX(w);
    cx[slot] = x0;
    cx[c.capC + slot] = x1;
    cx[2 * c.capC + slot] = x2;
  }
  u64 sa = cur.fpA(par), sb = cur.fpB(par);
  if (flags & CF_NEW) {
    const int label = (flags & CF_FINISH) ? -1 : ((c.lexicon && !c.lmToken) ? word : tok);
    fpChild(sa, sb, label, sa, sb);
  }
  candKeyOf(sa, sb, lex, tok, flags & CF_PB, cd.keyA(slot), cd.keyB(slot));
  gxSkey(w)[slot] = orderedKey64(score);
  return bin;
}

// deterministic order of two candidates (candBetter of beam_core.h; the lexicon-free records carry no
// word / lex fields)
FLT_DEV bool gxBetter(const DecCfg& c, const Cand& cd, int a, int b) {
  if (c.lexicon) return candBetter(cd, a, b);
  const double sa = cd.score(a), sb = cd.score(b);
  if (sa != sb) return sa > sb;
  if (cd.par(a) != cd.par(b)) return cd.par(a) < cd.par(b);
  if (cd.tok(a) != cd.tok(b)) return cd.tok(a) < cd.tok(b);
  return (cd.flags(a) & CF_PB) < (cd.flags(b) & CF_PB);
}

// M: the merge, one candidate per thread (every record is visible: a barrier separates it from E2).
// Candidates with equal (LM state, lex node, token, prevBlank) keys meet in one slot of the CTA-private
// table; the better one keeps it, the other's score key is zeroed (Utils.h:176-198, max-merge). Also
// publishes the best score key of the frame and clears the twin set's scalars (the next frame's).
FLT_DEV void gxPhaseM(const Cta& cta, const DecCfg& c, const Ws& w, int set, int nCand) {
  const Cand cd = w.cand();
  int* mh = w.mh();
  int* cslot = w.cslot();
  u64* skey = gxSkey(w);
  int* sc = gxSc(w, set);
  const uint32_t mask = (uint32_t)c.capH - 1;
  u64 best = 0;
  int dead = 0;
  for (int x = cta.tid; x < nCand; x += cta.nthr) {
    const u64 ka = cd.keyA(x), kb = cd.keyB(x);
    const u64 sk = skey[x];
    best = sk > best ? sk : best;
    uint32_t s = (uint32_t)ka & mask;
    for (;;) {
      int occ = atomCAS(&mh[s], -1, x);
      if (occ == -1) break;
      if (cd.keyA(occ) == ka && cd.keyB(occ) == kb) {
        for (;;) {
          if (!gxBetter(c, cd, x, occ)) {
            skey[x] = 0;
            break;
          }
          const int old = atomCAS(&mh[s], occ, x);
          if (old == occ) {
            skey[occ] = 0;
            break;
          }
          occ = old;
        }
        ++dead;
        break;
      }
      s = (s + 1) & mask;
    }
    cslot[x] = (int)s;
  }
#if FLT_DEVICE_BUILD
  {
    const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(best >> 32));
    const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(best >> 32) == hi ? (unsigned)best : 0u);
    best = ((u64)hi << 32) | lo;
    if ((cta.tid & 31) == 0 && best) atomicMax((u64*)(w.base + c.lay.gxBest), best);
    dead = __reduce_add_sync(0xffffffffu, dead);
    if ((cta.tid & 31) == 0 && dead) atomicAdd(&sc[GX_NDEAD], dead);
  }
#else
  {
    u64* gb = (u64*)(w.base + c.lay.gxBest);
    if (best > *gb) *gb = best;
    sc[GX_NDEAD] += dead;
  }
#endif
  if (cta.tid == 0) {
    int* gn = gxSc(w, set ^ 1);
    gn[GX_NCAND] = 0;
    gn[GX_OVF] = 0;
    gn[GX_NDEAD] = 0;
    gn[GX_NH] = 0;
  }
}

/* ------------------------------------------------------------------ walkers -------------------- */
// One slot of walkers: each group of kGxWalkLanes lanes expands one hypothesis over the ranked list.
template <bool LEX>
FLT_DEV void gxWalkSlot(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const FrameIn& f,
                        const GxFrame& fr, int set, int slot, int nWalk) {
#if FLT_DEVICE_BUILD
  const int lane = cta.tid & 31;
  const int width = kGxWalkLanes, sub = lane & (width - 1);
  const int r = slot * (32 / width) + lane / width;
#else
  const int width = 1, sub = 0, r = slot;
#endif
  int i = -1;
  if (r < nWalk) i = LEX ? gxWalkList(w, set)[r] : r;
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
        cont = !(approx + slack < fr.stop);
        if (cont && (!LEX || child >= 0)) {
          bool ok;
          if (LEX) ok = pbi || n != ti; // LexiconDecoder.cpp:89-90 (CTC)
          else ok = c.ctc ? (n != c.blank && (n != ti || pbi)) : n != ti; // LexiconFreeDecoder.cpp:69-71
          if (n == c.sil && c.silScore > 0) ok = false; // proposed as a special item by every hypothesis
          if (ok) {
            double score = base;
            if (n == c.sil) score += c.silScore;
            score = score + c.lmWeight * (double)d; // ZeroLM / root child: lmWeight * (maxScore - 0)
            gxOffer(cta, c, w, cur, fr, set, score, i, n, -1, LEX ? child : 0, LEX ? 0 : CF_NEW, d, ev, eoff, degLab,
                    msBits);
          }
        }
      }
    }
#if FLT_DEVICE_BUILD
    const unsigned b = __ballot_sync(0xffffffffu, cont);
    const unsigned gm = ((1u << width) - 1u) << (lane - sub);
    if ((b & gm) == 0) active = false; // the whole chunk is below the bound: so is the rest of the list
    if (b == 0) break;
#else
    if (!cont) break;
#endif
    j += width;
  }
}
FLT_DEV int gxWalkSlots(int nWalk) {
#if FLT_DEVICE_BUILD
  return (nWalk + (32 / kGxWalkLanes) - 1) / (32 / kGxWalkLanes);
#else
  return nWalk;
#endif
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

// stage 3 of one item: scores and proposals. Returns 1 + the best bin a proposal reached, 0 if none,
// 255 if a word end was not probed (its bin is unknown: the second pass must look again).
template <bool LEX>
FLT_DEV int gxItemFinish(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const FrameIn& f,
                         const GxFrame& fr, int set, const GxItem& it) {
  if (it.kind < 0) return 0;
  const int q = it.q;
  const double sq = cur.score(q);
  const int tq = cur.tok(q), pbq = cur.pb(q);
  if (!LEX) {
    if (it.kind == 1) { // repeat (LexiconFreeDecoder.cpp:98-110)
      const int n = tq;
      const bool isRepeat = c.ctc ? (!pbq && n != c.blank) : true;
      if (!(isRepeat && n >= 0 && n < c.N && inTokenSetV(c, f, n, it.ev))) return 0;
      double score = sq + (double)it.ev;
      if (n == c.sil) score += c.silScore;
      return 1 + gxOffer(cta, c, w, cur, fr, set, score, q, n, -1, 0, 0, 0.0f, it.ev, 0, 0, 0);
    }
    if (it.kind == 2) { // blank (:86-97)
      if (!c.ctc || !inTokenSetV(c, f, c.blank, it.ev)) return 0;
      double score = sq + (double)it.ev;
      if (c.blank == c.sil) score += c.silScore;
      return 1 + gxOffer(cta, c, w, cur, fr, set, score, q, c.blank, -1, 0, CF_PB, 0.0f, it.ev, 0, 0, 0);
    }
    // boosted sil as a new token (:69-85)
    const int n = c.sil;
    const bool ok = c.ctc ? (n != c.blank && (n != tq || pbq)) : n != tq;
    if (!ok || !inTokenSetV(c, f, n, it.ev)) return 0;
    double score = sq + (double)it.ev;
    score += c.silScore;
    score = score + c.lmWeight * (double)0.0f;
    return 1 + gxOffer(cta, c, w, cur, fr, set, score, q, n, -1, 0, CF_NEW, 0.0f, it.ev, 0, 0, 0);
  }
  const int lex = cur.lex(q);
  const int eoffQ = cur.xv[q], degLabQ = cur.xv[c.K + q], msQ = cur.xv[2 * c.K + q];
  if (it.kind == 1) { // (2) same node, LexiconDecoder.cpp:167-194
    if (!(!pbq || lex == 0)) return 0;
    const int n = it.n;
    double score = sq + (double)it.ev;
    if (n == c.sil) score += c.silScore;
    return 1 + gxOffer(cta, c, w, cur, fr, set, score, q, n, -1, lex, 0, 0.0f, it.ev, eoffQ, degLabQ, msQ);
  }
  if (it.kind == 2) { // (3) blank, :196-213
    const double score = sq + (double)it.ev;
    return 1 + gxOffer(cta, c, w, cur, fr, set, score, q, c.blank, -1, lex, CF_PB, 0.0f, it.ev, eoffQ, degLabQ, msQ);
  }
  if (it.kind == 3) { // boosted sil from the root through root child `sil` (emitSilCell)
    if (it.child < 0 || gxNodeDeg(it.nd.z) == 0) return 0;
    const int n = c.sil;
    if (!(pbq || n != tq) || n == c.blank) return 0;
    if (!inTokenSetV(c, f, n, it.ev)) return 0;
    double score = sq + (double)it.ev;
    score += c.silScore;
    const float d = bitsF32((uint32_t)it.nd.x) - 0.0f;
    score = score + c.lmWeight * (double)d;
    return 1 + gxOffer(cta, c, w, cur, fr, set, score, q, n, -1, it.child, 0, d, it.ev, it.nd.y, it.nd.z, it.nd.x);
  }
  // (1) one Trie edge: child node it.child reached by token it.n (LexiconDecoder.cpp:62-141)
  const int n = it.n;
  const float ev = it.ev;
  if (!inTokenSetV(c, f, n, ev)) return 0;
  const float lexMax = lex == 0 ? 0.0f : bitsF32((uint32_t)msQ);
  double score = sq + (double)ev;
  if (n == c.sil) score += c.silScore;
  const int deg = gxNodeDeg(it.nd.z), nLab = gxNodeLabels(it.nd.z);
  int best = -1;
  if (lex != 0 && deg > 0 && (pbq || n != tq)) { // (1a); root children with kids are the walkers' cells
    const float d = bitsF32((uint32_t)it.nd.x) - lexMax;
    const double s = score + c.lmWeight * (double)d;
    const int b = gxOffer(cta, c, w, cur, fr, set, s, q, n, -1, it.child, 0, d, ev, it.nd.y, it.nd.z, it.nd.x);
    best = b > best ? b : best;
  }
  if (nLab > 0 && !(lex == 0 && tq == n)) { // (1b) word ends, :114-141
    // with an n-gram LM: no probes when even the LM's best possible score stays below the bound
    if (c.lm.kind != 0 && c.lmWeight >= 0) {
      const float dUp = c.lmUpper - lexMax;
      const double up = score + c.lmWeight * (double)dUp + c.wordScore + 1e-6;
      if (up < fr.stop) return fr.mode == 1 ? 255 : 1 + best;
    }
    for (int l = 0; l < nLab; ++l) {
      const int label = nLab == 1 ? it.nd.w : c.trie.labels[it.nd.w + l];
      const float d = gxWordLm(c, cur, q, label) - lexMax;
      const double s = score + c.lmWeight * (double)d + c.wordScore;
      const int b = gxOffer(cta, c, w, cur, fr, set, s, q, n, label, 0, CF_NEW, d, ev, 0, 0, 0);
      best = b > best ? b : best;
    }
  }
  return 1 + best;
}

/* ------------------------------------------------------------------ phase E (both passes) ------- */
template <bool LEX>
FLT_DEV void gxPhaseE(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const FrameIn& f,
                      const GxFrame& fr, int set, int nH, float eBlank, float eSil) {
  const int warp = cta.tid >> 5, nw = (cta.nthr + 31) >> 5;
  const int nSpec = gxSpecials(c);
  const int nWalk = LEX ? gxSc(w, set)[GX_NWALK] : nH;
  const int nSlots = gxWalkSlots(nWalk);
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
  unsigned char* stash = gxStash(w);
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
      if (fr.mode == 0) { // second pass: only the items whose best proposal reached the cut
        const int st = stash[x];
        if (st == 0 || (st != 255 && st - 1 < fr.cutBin)) continue;
      }
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
      for (int s = warp; s < nSlots; s += nw) gxWalkSlot<LEX>(cta, c, w, cur, f, fr, set, s, nWalk);
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
    for (int z = 0; z < kGxSweep; ++z) {
      const int x = x0 + z * cta.nthr + cta.tid;
      const int st = gxItemFinish<LEX>(cta, c, w, cur, f, fr, set, it[z]);
      if (fr.mode == 1 && x < nItems) stash[x] = (unsigned char)(st > 255 ? 255 : st);
    }
  }
}

/* ------------------------------------------------------------------ the cut -------------------- */
// From the histogram: the highest bin b with at least `want` proposals in bins >= b (0 if there are
// fewer in total). kept = proposals in bins >= b, above = proposals in bins > b. Every warp computes
// this for itself from the same data (identical results; no barrier, no single-warp phase).
FLT_DEV void gxFindCut(const Cta& cta, const Ws& w, int want, int& cutBin, int& kept, int& above) {
  const int* hist = w.hist();
#if FLT_DEVICE_BUILD
  const int lane = cta.tid & 31;
  int h[8], part = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    h[k] = hist[255 - (lane * 8 + k)];
    part += h[k];
  }
  int incl = part;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int excl = incl - part;
  int found = -1, fk = 0, fa = 0;
  if (excl < want && incl >= want) {
    int cum = excl;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (cum < want && cum + h[k] >= want) {
        found = 255 - (lane * 8 + k);
        fk = cum + h[k];
        fa = cum;
      }
      cum += h[k];
    }
  }
  const unsigned who = __ballot_sync(0xffffffffu, found >= 0);
  const int bin0 = __shfl_sync(0xffffffffu, h[7], 31);
  if (who) {
    const int src = __ffs(who) - 1;
    cutBin = __shfl_sync(0xffffffffu, found, src);
    kept = __shfl_sync(0xffffffffu, fk, src);
    above = __shfl_sync(0xffffffffu, fa, src);
  } else {
    cutBin = 0;
    kept = total;
    above = total - bin0;
  }
#else
  (void)cta;
  int cum = 0;
  cutBin = 0;
  for (int b = 255; b >= 1; --b) {
    if (cum + hist[b] >= want) {
      cutBin = b;
      above = cum;
      kept = cum + hist[b];
      return;
    }
    cum += hist[b];
  }
  above = cum;
  kept = cum + hist[0];
#endif
}

/* ------------------------------------------------------------------ phase RF -------------------- */
// Every live candidate ranks itself by counting larger score keys; the thread that finds rank q < K
// writes hypothesis q of the new beam (phaseFinalize of beam_core.h, one hypothesis per thread),
// registers it for the next frame's E, and frees its merge-table slot. The new beam's size
// accumulates in the twin set's GX_NH.
FLT_DEV void gxPhaseRF(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const Beam& nxt,
                       const FrameIn& f, int set, int nCand) {
  const Cand cd = w.cand();
  const int K = c.K;
  int* gn = gxSc(w, set ^ 1);
  const u64* skey = gxSkey(w);
  int* mh = w.mh();
  const int* cslot = w.cslot();
  const u64 bestKey = *(const u64*)(w.base + c.lay.gxBest);
  // candidatesBestScore_ - beamThreshold (Utils.h:161-165; a max-merge keeps the same groups when the
  // filter runs after the merge)
  const double thrScore = keyToDouble(bestKey) - c.beamThreshold;
  int lg = 0;
  while (lg < 5 && ((long long)nCand << (lg + 1)) <= cta.nthr) ++lg;
  const int parts = 1 << lg;
  const int slice = (nCand + parts - 1) >> lg;
  for (int base = 0; base < (nCand << lg); base += cta.nthr) {
    const int t = base + cta.tid;
    const int a = t >> lg, part = t & (parts - 1);
    u64 ka = 0;
    if (a < nCand) ka = skey[a];
    const bool valid = ka != 0; // live
    int cnt = 0, eq = 0;
    if (valid) {
      const int lo = part * slice, hi = lo + slice < nCand ? lo + slice : nCand;
#pragma unroll 4
      for (int b = lo; b < hi; ++b) {
        const u64 kb = skey[b];
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
    int keep = 0; // q + 1 if this thread writes hypothesis q
    const int x = a;
    if (valid && part == 0) {
      mh[cslot[x]] = -1; // leave the merge table empty
      if (eq) { // another group with exactly this score: settle by the deterministic order (rare)
        for (int b = 0; b < nCand; ++b)
          if (b != a && skey[b] == ka && gxBetter(c, cd, b, x)) ++cnt;
      }
      if (cnt < K && cd.score(x) >= thrScore) keep = cnt + 1;
    }
#if FLT_DEVICE_BUILD
    {
      const int mx = __reduce_max_sync(0xffffffffu, keep);
      if ((cta.tid & 31) == 0 && mx > 0) atomicMax(&gn[GX_NH], mx);
    }
#else
    if (keep > gn[GX_NH]) gn[GX_NH] = keep;
#endif
    if (!keep) continue;
    const int q = keep - 1;
    const double score = cd.score(x);
    const int p = cd.par(x);
    const int fl = cd.flags(x);
    const int n = cd.tok(x);
    nxt.score(q) = score;
    nxt.am(q) = (fl & CF_FINISH) ? cur.am(p) : cur.am(p) + amOf(c, f, cd.ce(x), n, cur.tok(p));
    nxt.lm(q) = c.lexicon ? cur.lm(p) + (double)cd.lmd(x) : cur.lm(p) + (double)0.0f;
    const int lexNew = c.lexicon ? cd.lex(x) : 0;
    nxt.lex(q) = lexNew;
    nxt.tok(q) = n;
    nxt.pb(q) = (fl & CF_PB) ? 1 : 0;
    if (fl & CF_NEW) {
      const int lab = (fl & CF_FINISH) ? -1 : ((c.lexicon && !c.lmToken) ? cd.word(x) : n);
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
      const int x1 = cx[c.capC + x];
      gxSetNode(nxt, q, cx[x], x1, cx[2 * c.capC + x]);
      gxRegister(c, w, set ^ 1, q, lexNew, x1);
    }
  }
}

/* ------------------------------------------------------------------ the frame step ------------- */
// emission level of a frame: the best raw proposal any hypothesis can make (blank, or the head of the
// ranked list with its LM part)
template <bool LEX>
FLT_DEV double gxLevel(const DecCfg& c, const FrameIn& f, float eBlank, bool blankOk, bool& known) {
  double lv = negInf();
  known = false;
  if (blankOk) {
    lv = (double)eBlank;
    known = true;
  }
  if (f.listLen > 0 && f.topTok[0] >= 0) {
    double v = (double)f.topVal[0];
    if (LEX) v = v + c.lmWeight * (double)bitsF32((uint32_t)f.listInfo[0]);
    lv = v > lv ? v : lv;
    known = true;
  }
  return lv;
}

// One frame: beam `set` -> beam `set ^ 1`. All threads call this with identical arguments; returns the
// number of hypotheses in the new beam (0 = the beam died, Utils.h:155-158).
template <bool LEX>
FLT_DEV int gxFrameStep(const Cta& cta, const DecCfg& c, const Ws& w, int set, const FrameIn& f, int* status,
                        unsigned long long* stats, GxCarry& g) {
  int* gs = gxSc(w, set);
  const int nH = gs[GX_NH];
  if (nH == 0) return 0;
  const Beam cur = w.beam(set), nxt = w.beam(set ^ 1);
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
  // the best hypothesis' blank candidate always exists: the frame's best is at least that, and the
  // reference drops everything below it minus beamThreshold
  fr.floor = negInf();
  if (blankOk) {
    double sb = s0 + (double)eBlank;
    if (!LEX && c.blank == c.sil) sb += c.silScore;
    fr.floor = sb - c.beamThreshold;
  }
  bool levelKnown;
  const double level = gxLevel<LEX>(c, f, eBlank, blankOk, levelKnown);
  // histogram range: proposals lie below best hypothesis + best emission (+ bonuses)
  double hi = s0 + (levelKnown ? level : 0.0);
  if (c.silScore > 0) hi += c.silScore;
  if (c.wordScore > 0) hi += c.wordScore;
  hi += 0.25;
  double span = (double)g.span;
  if (c.beamThreshold + 4.0 < span) span = c.beamThreshold + 4.0;
  fr.hlo = hi - span;
  fr.hscale = (float)((double)kGxBins / span);
  // Bound of the histogram pass for the walkers and the LM probes. Lexicon-free: the corner bound of
  // beam_core.h — rows 1..a x columns 0..col hold >= K distinct groups, so the cut is never below the
  // best corner (exact). Lexicon: the last cut moved by the emission level, minus a generous margin; a
  // proposal it hides is only missing from the histogram (the cut comes out lower, never wrong).
  double bound = negInf();
  if (!LEX) {
    for (int k = 0; k < c.nTau; ++k) {
      const int i = 2 * c.tauA[k] - 2, col = c.tauCol[k];
      if (i < nH && col < f.listLen && f.topTok[col] >= 0) {
        const double corner = cur.score(i) + (double)f.topVal[col] + c.lmWeight * (double)0.0f;
        bound = corner > bound ? corner : bound;
      }
    }
    if (c.silScore < 0) bound += c.silScore; // keeps the bound valid if a counted cell is the sil one
  } else if (g.have == 2 && levelKnown) {
    bound = g.cutPrev + level + g.D - 0.75;
  }
  if (!(bound == bound)) bound = negInf();
  int want = LEX ? K + (K >> 1) + 32 : 2 * K + 16;
  int nCand = 0, nRep = 0, cutBin = 0, kept = 0, above = 0, redo = 0;
  bool last = false;
  for (int hpass = 0;; ++hpass) {
    // ---- E1: histogram of every proposal
    fr.mode = 1;
    fr.cutBin = 0;
    fr.stop = bound > fr.floor ? bound : fr.floor;
    gxPhaseE<LEX>(cta, c, w, cur, f, fr, set, nH, eBlank, eSil);
    cta.sync(); // ---- B1
    if (hpass == 0) pc.mark(0);
    bool rehist = false;
    for (int attempt = 0;; ++attempt) {
      gxFindCut(cta, w, want, cutBin, kept, above);
      if (kept > c.capC - 16 && !last) {
        if (cutBin < kGxBins - 1 && above >= K + (K >> 2) && above <= c.capC - 16) {
          // a crowded cut bin: the bins above it may hold the K groups on their own (verified below)
          ++cutBin;
          kept = above;
        } else {
          // more proposals in (and above) the cut bin than candidate slots: histogram again at a finer
          // scale over the cut bin — or over a longer range if the cut fell below this one
          const double binW = 1.0 / (double)fr.hscale;
          if (cutBin == 0) {
            const double len = (double)kGxBins * binW;
            fr.hlo -= 7.0 * len;
            fr.hscale = (float)(1.0 / (8.0 * binW));
          } else {
            const double nw = binW / (double)(kGxBins - 2);
            fr.hlo = fr.hlo + (double)cutBin * binW - nw;
            fr.hscale = (float)(1.0 / nw);
          }
          rehist = true;
          break;
        }
      }
      // ---- E2: materialise what reaches the cut
      fr.mode = 0;
      fr.cutBin = cutBin;
      const double edge = cutBin > 0 ? fr.hlo + ((double)cutBin - 0.01) / (double)fr.hscale : negInf();
      fr.stop = edge > fr.floor ? edge : fr.floor;
      if (!LEX && bound > fr.stop) fr.stop = bound; // exact: nothing below the corner bound is ever needed
      gxPhaseE<LEX>(cta, c, w, cur, f, fr, set, nH, eBlank, eSil);
      cta.sync(); // ---- B2
      if (hpass == 0 && attempt == 0) pc.mark(1);
      nCand = gs[GX_NCAND];
      const int ovf = gs[GX_OVF];
      if (nCand > c.capC) nCand = c.capC;
      gxPhaseM(cta, c, w, set, nCand);
      cta.sync(); // ---- B3
      if (hpass == 0 && attempt == 0) pc.mark(2);
      nRep = nCand - gs[GX_NDEAD];
      // complete if nothing was left out (cut at bin 0; lexicon-free: and no corner bound in the way)
      const bool all = cutBin == 0 && (LEX || !(bound > fr.floor));
      const bool miss = nRep < K && !all;
      if ((!ovf && !miss) || last) break;
      // ---- rare: too few groups above the cut, or more candidates than the histogram promised. Undo.
      redo = 1;
      for (int x = cta.tid; x < nCand; x += cta.nthr)
        if (gxSkey(w)[x]) w.mh()[w.cslot()[x]] = -1;
      cta.sync();
      if (cta.tid == 0) {
        gs[GX_NCAND] = 0;
        gs[GX_OVF] = 0;
        gs[GX_NDEAD] = 0;
        *(u64*)(w.base + c.lay.gxBest) = 0;
      }
#if FLT_DEVICE_BUILD
      if (stats && cta.tid == 0) {
        atomicAdd(stats + 13, (unsigned long long)(ovf ? 1 : 0));
        atomicAdd(stats + 14, (unsigned long long)(miss ? 1 : 0));
      }
#endif
      cta.sync();
      if (ovf) { // proposals the bound hid from the histogram: count everything this time
        if (!(bound > negInf())) last = true; // nothing was hidden: the capacity itself is too small
        bound = negInf();
        rehist = true;
        break;
      }
      want = want * 2;
      if (attempt >= 8) last = true;
    }
    if (!rehist) break;
    redo = 1;
    if (hpass >= 4) last = true; // ties en masse: the host grows the capacity and redoes the batch
    for (int b = cta.tid; b < kGxBins; b += cta.nthr) w.hist()[b] = 0;
    cta.sync();
  }
  if (last && cta.tid == 0) *status |= 1;
  // leave the histogram empty for the next frame
  for (int b = cta.tid; b < kGxBins; b += cta.nthr) w.hist()[b] = 0;
#if FLT_DEVICE_BUILD
  if (stats && cta.tid == 0) {
    atomicAdd(stats + 0, 1ull);
    atomicAdd(stats + 1, (unsigned long long)nCand);
    atomicAdd(stats + 2, (unsigned long long)nRep);
    atomicAdd(stats + 3, (unsigned long long)(nRep < K ? nRep : K));
    if (redo) atomicAdd(stats + 12, 1ull);
    if (last) atomicAdd(stats + 15, 1ull);
    stats[30] = 1ull; // phase names of this step for the host
  }
#endif
  gxPhaseRF(cta, c, w, cur, nxt, f, set, nCand);
  cta.sync(); // ---- B4
  pc.mark(3);
  // tail: this frame's tables are dead — clear their counters for the frame after next
  const int nHn = gxSc(w, set ^ 1)[GX_NH];
  if (cta.tid == 0) {
    gs[GX_NCHUNK] = 0;
    gs[GX_NWALK] = 0;
    *(u64*)(w.base + c.lay.gxBest) = 0;
    if (LEX && gxSc(w, set ^ 1)[GX_NCHUNK] > c.capChunks) *status |= 1;
  }
  // carry: where the cut went (bound of the next histogram pass, range of the next histogram)
  if (nRep >= K && nHn == K && levelKnown) {
    const double cutNow = nxt.score(K - 1);
    if (g.have >= 1) {
      g.D = cutNow - g.cutPrev - level;
      g.have = 2;
    } else {
      g.have = 1;
    }
    g.cutPrev = cutNow;
    float sp = (float)((hi - cutNow) * 1.5 + 2.0);
    sp = sp < 4.0f ? 4.0f : (sp > 128.0f ? 128.0f : sp);
    g.span = 0.5f * g.span + 0.5f * sp;
  } else {
    g.have = 0;
    g.span = 32.0f;
  }
  if (f.hCount && cta.tid == 0) *f.hCount = nHn;
  return nHn;
}

// decodeEnd (LexiconFreeDecoder.cpp:127-158, LexiconDecoder.cpp:231-274) through the same phases:
// one finish candidate per hypothesis (only those at the Trie root when any exists), merged, ranked.
// Returns the number of final hypotheses (in beam set ^ 1, sorted).
template <bool LEX>
FLT_DEV int gxFinish(const Cta& cta, const DecCfg& c, const Ws& w, int set, const FrameIn& f) {
  int* sc = w.sc();
  int* gs = gxSc(w, set);
  const int nH = gs[GX_NH];
  if (nH == 0) return 0;
  const Beam cur = w.beam(set), nxt = w.beam(set ^ 1);
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
  GxFrame fr;
  fr.mode = 0;
  fr.floor = negInf();
  fr.hlo = 0.0;
  fr.hscale = 0.0f;
  fr.cutBin = 0;
  fr.stop = negInf();
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    if (nice && cur.lex(i) != 0) continue;
    float ls = 0.0f;
    int flags = CF_FINISH;
    if (c.lm.kind) { // KenLM::finish: score </s>, state = child(-1); ZeroLM: same state, 0
      ls = ngramScore(c.lm, cur.ctx(i), cur.nctx(i), c.lm.eos);
      flags |= CF_NEW;
    }
    const double score = cur.score(i) + c.lmWeight * (double)ls;
    gxOffer(cta, c, w, cur, fr, set, score, i, c.sil, -1, LEX ? cur.lex(i) : 0, flags, ls, 0.0f, 0, 0, 0);
  }
  cta.sync();
  const int nCand = gs[GX_NCAND] < c.capC ? gs[GX_NCAND] : c.capC; // <= K <= capC
  gxPhaseM(cta, c, w, set, nCand);
  cta.sync();
  gxPhaseRF(cta, c, w, cur, nxt, f, set, nCand);
  cta.sync();
  const int nFin = gxSc(w, set ^ 1)[GX_NH];
  if (cta.tid == 0) *(u64*)(w.base + c.lay.gxBest) = 0;
  if (f.hCount && cta.tid == 0) *f.hCount = nFin;
  return nFin;
}

// once per CTA: tables that persist over its utterances
FLT_DEV void gxInitWorkspace(const Cta& cta, const DecCfg& c, const Ws& w) {
  for (int i = cta.tid; i < c.capH; i += cta.nthr) w.mh()[i] = -1;
  for (int i = cta.tid; i < kGxBins; i += cta.nthr) w.hist()[i] = 0;
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
    gxSc(w, 0)[GX_NH] = nH;
    *(u64*)(w.base + c.lay.gxBest) = 0;
  }
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
  const Ws w = wsOf(base, c, nullptr); // one region
  const int K = c.K;
  gxInitWorkspace(cta, c, w);
  for (int b = cta.bid; b < a.B; b += cta.nblk) {
    const int len = a.lengths ? a.lengths[b] : a.T;
    int set = 0;
    cta.sync(); // previous utterance fully retired
    if (a.streamBeam && a.streamRestore) streamRestoreBeam(cta, c, w, a);
    else if (cta.tid == 0) seedUtterance(c, w, a, b);
    cta.sync();
    int nHcur = w.sc()[SC_NH];
    gxBeginUtterance(cta, c, w, nHcur);
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
      nHcur = gxFrameStep<LEX>(cta, c, w, set, f, a.status + b, a.stats, g);
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
      if (nHcur == 0) break; // the beam died; `set` still names the last beam that was read
      set ^= 1;
      cta.sync();
    }
    if (a.streamBeam && a.streamNoFinish) { // decodeStep chunk: keep the beam for the next launch
      cta.sync();
      if (cta.tid == 0) w.sc()[SC_NH] = nHcur;
      cta.sync();
      streamSaveBeam(cta, c, w, a, set);
      continue;
    }
    int nFin = 0;
    if (nHcur != 0) {
      const FrameIn f = finishFrameIn(c, a, b, len);
      nFin = gxFinish<LEX>(cta, c, w, set, f);
      set ^= 1;
    }
    cta.sync();
    writeFinals(cta, c, w, a, b, set, nFin);
  }
}

} // namespace flt
