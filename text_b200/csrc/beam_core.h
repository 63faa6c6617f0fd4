// beam_core.h — the per-frame beam expansion (K2/K3), finish (decodeEnd) and n-best backtrace (K4)
// of the B200 decode path, one CTA per utterance.
//
// What it replaces (reference, flashlight/lib/text/decoder/...):
//   LexiconFreeDecoder::decodeStep  LexiconFreeDecoder.cpp:53-122
//   LexiconDecoder::decodeStep      LexiconDecoder.cpp:54-225
//   candidatesAdd / candidatesStore Utils.h:131-225   (threshold filter, key merge, top-K)
//   decodeEnd                       LexiconFreeDecoder.cpp:127-158, LexiconDecoder.cpp:231-274
//   getAllHypothesis                Utils.h:229-266
//   LMState::child identity         lm/LM.h:24-49 (here: an interned id per (parent id, label))
//
// It is NOT a translation of those loops. The reference proposes beam x beamSizeToken candidates
// per frame and sorts them; here the candidate set is cut down *exactly* before anything is
// materialised:
//   * hypotheses that share an LM state (and lexicon node) form a "row"; for new-token expansions
//     only the best eligible member of a row can win a max-merge, because fl(x+e) is monotone in x;
//   * all rows of the "wide" kind (every row of the lexicon-free decoder; root-node rows of the
//     lexicon decoder) rank their expansions by the same per-frame token ordering, so candidate
//     (row rank r, column j) is dominated by (r')(j') for r' <= r, j' <= j: it can only reach the
//     top-K if r*(j-2) <= K. Only those cells are generated (K ln K instead of K*N);
//   * everything else (trie children of non-root rows, stay / blank, word ends) is enumerated
//     directly from the CSR trie.
// Candidates are then merged by exact key in a CTA-private hash table, the K best groups are
// found with a radix select on order-preserving 64-bit keys, and survivors are ranked.
// Arithmetic follows the reference's evaluation order (FP64 accumulators, FP32 sub-expressions,
// no FMA contraction: this file is compiled with -fmad=false).
#pragma once
#include "spmd.h"
#include "tables.h"

namespace flt {

/* ------------------------------------------------------------------ configuration ---------- */
struct DecCfg {
  int lexicon;   // 0 = LexiconFreeDecoder, 1 = LexiconDecoder
  int K;         // beamSize
  int N;         // tokens per frame
  int setAll;    // beamSizeToken >= N: every token is in the token set
  double beamThreshold, lmWeight, wordScore, unkScore, silScore;
  int logAdd, ctc, hasUnk;
  int sil, blank, unk;
  const float* trans; // [N*N] ASG transitions (device) or null
  int M;          // entries per frame in the token list produced by the select kernel
  int Mwide;      // columns the wide enumeration may use (<= M)
  int wideRanked; // 1 = wide rows use the ranked list; 0 = they enumerate their children directly
  int capC;       // candidate capacity
  int capH;       // merge table slots (pow2 >= 2*capC)
  int capRH;      // row table slots (pow2 >= 2*K)
  int capP;       // pow2 >= K (sort scratch)
  const int* wideOff; // [K+1]: wideOff[r] = sum_{q=1..r} J_q, J_q = min(Mwide, K/q + 3)
  TrieDev trie;
  LmDev lm;
};

/* Per-utterance arguments of one decode launch. */
struct BatchArgs {
  const float* emis; // [B,T,N]
  int B, T;
  const int* lengths;   // [B] or null
  const int* topTok;    // [B*T, M]
  const float* topVal;  // [B*T, M]
  const float* thrVal;  // [B*T] value of the beamSizeToken-th largest emission, or null (setAll)
  int* hParent;         // history [B, T+2, K]
  int* hTok;
  int* hWord;           // null for the lexicon-free decoder
  double* finScore;     // [B, K, 3]
  int* finCount;        // [B]
  int* status;          // [B] bit0 = candidate overflow, bit1 = state table overflow
  char* wsGlobal;       // per-CTA workspace slabs (used when the workspace does not fit smem)
  long long wsStride;
  unsigned long long* stateTab; // per-CTA LM-state intern tables
  long long stateCap;           // slots per table (pow2)
  int useSmem;
};

/* ------------------------------------------------------------------ workspace ---------- */
struct Beam {
  double* score;
  double* am;
  double* lm;
  int* sid;   // interned LM-state id (0 = the start state)
  int* spid;  // (parent id, label) that names this state: the exact, time-invariant identity
  int* slab;
  int* lex;   // trie node (0 = root); always 0 for the lexicon-free decoder
  int* tok;
  int* pb;    // prevBlank
  int* nctx;  // n-gram context length
  int* ctx;   // [K, kMaxCtx] LM vocabulary ids, most recent first
};

enum { // ws.sc[] scalars
  SC_NH = 0, SC_NCAND, SC_NREP, SC_NSEL, SC_NROWS, SC_NARROW_ITEMS, SC_WIDE_ITEMS, SC_OVF,
  SC_BIN, SC_NEED, SC_BINCOUNT, SC_TIES, SC_COUNT
};
enum { // ws.sq[] 64-bit scalars
  SQ_BEST = 0, SQ_MIN, SQ_PREFIX, SQ_COUNT
};
constexpr int CF_PB = 1, CF_NEW = 2, CF_ALIVE = 4, CF_FINISH = 8;
constexpr int kIntMax = 0x7FFFFFFF;

struct Ws {
  Beam beam[2];
  // rows
  int* rowHash;      // [capRH]
  int* rowOf;        // [K] leader of the row of hyp i (wide hyps)
  int* m2;           // [K] second-best member of the row led by i
  int* rank;         // [K] scan scratch / row rank
  int* rankTmp;      // [K]
  int* leaderOfRank; // [K]
  int* deg;          // [K+1] narrow items per hyp (scan)
  int* degTmp;       // [K+1]
  // candidates
  double* cscore;
  int* cpar;
  int* ctok;
  int* cword;
  int* clex;
  int* cflag;
  int* clab;
  float* clmd;
  float* ce;
  int* mh;   // [capH] merge table: candidate index of the group's best member, -1 empty
  int* rep;  // [capC] group representatives
  int* surv; // [capP] selected, then sorted
  int* survTmp;
  int* hist; // [256]
  int* sc;
  unsigned long long* sq;
  unsigned long long* red; // [64] CTA-reduction scratch
};

FLT_HD size_t alignUp(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carve the workspace out of `base` (nullptr => just compute the size).
FLT_HD size_t carveWs(char* base, const DecCfg& c, Ws& w) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off = alignUp(off + bytes, 16);
    return p;
  };
  const int K = c.K;
  for (int b = 0; b < 2; ++b) {
    Beam& B = w.beam[b];
    B.score = (double*)take(sizeof(double) * K);
    B.am = (double*)take(sizeof(double) * K);
    B.lm = (double*)take(sizeof(double) * K);
    B.sid = (int*)take(sizeof(int) * K);
    B.spid = (int*)take(sizeof(int) * K);
    B.slab = (int*)take(sizeof(int) * K);
    B.lex = (int*)take(sizeof(int) * K);
    B.tok = (int*)take(sizeof(int) * K);
    B.pb = (int*)take(sizeof(int) * K);
    B.nctx = (int*)take(sizeof(int) * K);
    B.ctx = (int*)take(sizeof(int) * K * (c.lm.kind ? kMaxCtx : 0));
  }
  w.rowHash = (int*)take(sizeof(int) * c.capRH);
  w.rowOf = (int*)take(sizeof(int) * K);
  w.m2 = (int*)take(sizeof(int) * K);
  w.rank = (int*)take(sizeof(int) * (K + 1));
  w.rankTmp = (int*)take(sizeof(int) * (K + 1));
  w.leaderOfRank = (int*)take(sizeof(int) * K);
  w.deg = (int*)take(sizeof(int) * (K + 1));
  w.degTmp = (int*)take(sizeof(int) * (K + 1));
  w.cscore = (double*)take(sizeof(double) * c.capC);
  w.cpar = (int*)take(sizeof(int) * c.capC);
  w.ctok = (int*)take(sizeof(int) * c.capC);
  w.cword = (int*)take(sizeof(int) * c.capC);
  w.clex = (int*)take(sizeof(int) * c.capC);
  w.cflag = (int*)take(sizeof(int) * c.capC);
  w.clab = (int*)take(sizeof(int) * c.capC);
  w.clmd = (float*)take(sizeof(float) * c.capC);
  w.ce = (float*)take(sizeof(float) * c.capC);
  w.mh = (int*)take(sizeof(int) * c.capH);
  w.rep = (int*)take(sizeof(int) * c.capC);
  w.surv = (int*)take(sizeof(int) * c.capP);
  w.survTmp = (int*)take(sizeof(int) * c.capP);
  w.hist = (int*)take(sizeof(int) * 256);
  w.sc = (int*)take(sizeof(int) * SC_COUNT);
  w.sq = (unsigned long long*)take(sizeof(unsigned long long) * SQ_COUNT);
  w.red = (unsigned long long*)take(sizeof(unsigned long long) * 64);
  return off;
}

FLT_HD double negInf() { return bitsF64(0xFFF0000000000000ull); }
FLT_HD double keyToDouble(unsigned long long key) { // inverse of orderedKey64
  return bitsF64((key >> 63) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key);
}

/* ------------------------------------------------------------------ CTA collectives ---------- */
// max over the CTA of a 64-bit key; every thread gets the result.
FLT_DEV unsigned long long ctaMax64(const Cta& cta, unsigned long long v, unsigned long long* red) {
#if FLT_DEVICE_BUILD
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u > v ? u : v;
  }
  const int warp = cta.tid >> 5, lane = cta.tid & 31, nw = (cta.nthr + 31) >> 5;
  cta.sync(); // red[] may still be read from a previous call
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

// Exclusive prefix sum of a[0..n) in place; a[n] receives the total. tmp has n+1 entries.
FLT_DEV void ctaExclusiveScan(const Cta& cta, int* a, int* tmp, int n) {
  // Hillis-Steele inclusive scan with ping-pong, then shift.
  int* src = a;
  int* dst = tmp;
  for (int d = 1; d < n; d <<= 1) {
    for (int i = cta.tid; i < n; i += cta.nthr) dst[i] = src[i] + (i >= d ? src[i - d] : 0);
    cta.sync();
    int* t = src;
    src = dst;
    dst = t;
  }
  // src holds the inclusive scan; write exclusive into dst then copy back if needed
  for (int i = cta.tid; i <= n; i += cta.nthr) dst[i] = i == 0 ? 0 : src[i - 1];
  cta.sync();
  if (dst != a) {
    for (int i = cta.tid; i <= n; i += cta.nthr) a[i] = dst[i];
    cta.sync();
  }
}

// largest r in [0, n) with off[r] <= x, for a non-decreasing off[0..n]
FLT_DEV int searchOffsets(const int* off, int n, int x) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= x) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

/* ------------------------------------------------------------------ LM-state interning ---------- */
// find-or-insert (parent id, label) in the CTA's open-addressing table; the id is slot+1
// (0 is the start state). Returns -1 when the table is full.
FLT_DEV int internState(unsigned long long* tab, long long cap, int pid, int label) {
  const unsigned long long key = ((unsigned long long)(unsigned)pid << 32) | (unsigned)label;
  const unsigned long long kEmpty = ~0ull;
  unsigned long long s = mix64(key) & (unsigned long long)(cap - 1);
  for (long long probes = 0; probes < cap; ++probes) {
    unsigned long long cur = tab[s];
    if (cur == key) return (int)s + 1;
    if (cur == kEmpty) {
      unsigned long long old = atomCAS64(&tab[s], kEmpty, key);
      if (old == kEmpty || old == key) return (int)s + 1;
    }
    s = (s + 1) & (unsigned long long)(cap - 1);
  }
  return -1;
}

/* ------------------------------------------------------------------ candidates ---------- */
FLT_DEV void putCand(Ws& w, int slot, double score, int par, int tok, int word, int lex, int flags,
                     float lmd, int lab, float ev) {
  w.cscore[slot] = score;
  w.cpar[slot] = par;
  w.ctok[slot] = tok;
  w.cword[slot] = word;
  w.clex[slot] = lex;
  w.cflag[slot] = flags | CF_ALIVE;
  w.clmd[slot] = lmd;
  w.clab[slot] = lab;
  w.ce[slot] = ev;
}

// exact identity of the candidate's LM state as a (parent id, label) pair
FLT_DEV void candState(const Ws& w, const Beam& cur, int c, int& pid, int& lab) {
  const int p = w.cpar[c];
  if (w.cflag[c] & CF_NEW) {
    pid = cur.sid[p];
    lab = w.clab[c];
  } else {
    pid = cur.spid[p];
    lab = cur.slab[p];
  }
}

FLT_DEV bool candKeyEq(const Ws& w, const Beam& cur, int a, int b) {
  if (w.ctok[a] != w.ctok[b] || w.clex[a] != w.clex[b]) return false;
  if ((w.cflag[a] ^ w.cflag[b]) & CF_PB) return false;
  int pa, la, pb_, lb;
  candState(w, cur, a, pa, la);
  candState(w, cur, b, pb_, lb);
  return pa == pb_ && la == lb;
}

// deterministic total order used wherever the reference leaves ties to libstdc++ internals:
// higher score first, then lower parent rank, token, word, lexicon node.
FLT_DEV bool candBetter(const Ws& w, int a, int b) {
  const double sa = w.cscore[a], sb = w.cscore[b];
  if (sa != sb) return sa > sb;
  if (w.cpar[a] != w.cpar[b]) return w.cpar[a] < w.cpar[b];
  if (w.ctok[a] != w.ctok[b]) return w.ctok[a] < w.ctok[b];
  if (w.cword[a] != w.cword[b]) return w.cword[a] < w.cword[b];
  if (w.clex[a] != w.clex[b]) return w.clex[a] < w.clex[b];
  return (w.cflag[a] & CF_PB) < (w.cflag[b] & CF_PB);
}

FLT_DEV uint32_t candKeyHash(const Ws& w, const Beam& cur, int c) {
  int pid, lab;
  candState(w, cur, c, pid, lab);
  uint64_t h = ((uint64_t)(uint32_t)pid << 32) | (uint32_t)lab;
  h = mix64(h) ^ (((uint64_t)(uint32_t)w.clex[c] << 32) | ((uint32_t)w.ctok[c] << 1) |
                  (uint32_t)(w.cflag[c] & CF_PB));
  return (uint32_t)mix64(h);
}

/* token-set membership for tokens that are not taken from the ranked list */
FLT_DEV bool inTokenSet(const DecCfg& c, const float* e, int n, float thrVal, const int* topTok,
                        int listLen) {
  if (c.setAll) return true;
  const float v = e[n];
  if (v > thrVal) return true;
  if (v < thrVal) return false;
  for (int j = 0; j < listLen; ++j) // equal to the cut value: membership = presence in the list
    if (topTok[j] == n) return true;
  return false;
}

/* ------------------------------------------------------------------ the frame step ---------- */
struct FrameIn {
  const float* e;      // emission row [N]
  const int* topTok;   // [M] ranked tokens (may be null when unused)
  const float* topVal; // [M]
  int listLen;         // valid entries in the list
  float thrVal;        // cut value of the token set (unused when setAll)
  int first;           // global frame 0 (ASG transitions are skipped, LexiconDecoder.cpp:70-73)
  int* hParent;        // history row to write (frame t+1), [K]
  int* hTok;
  int* hWord;
};

// Phase R: group wide hypotheses into rows (same LM state; lexicon: also lex == root).
FLT_DEV void phaseRows(const Cta& cta, const DecCfg& c, Ws& w, const Beam& cur, int nH) {
  for (int s = cta.tid; s < c.capRH; s += cta.nthr) w.rowHash[s] = -1;
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    w.m2[i] = kIntMax;
    w.rowOf[i] = -1;
  }
  cta.sync();
  const uint32_t mask = (uint32_t)c.capRH - 1;
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    if (c.lexicon && cur.lex[i] != 0) continue;
    uint32_t s = (uint32_t)mix64((uint64_t)(uint32_t)cur.sid[i]) & mask;
    for (;;) {
      int old = atomCAS(&w.rowHash[s], -1, i);
      if (old == -1) break;
      if (cur.sid[old] == cur.sid[i]) {
        atomMin(&w.rowHash[s], i);
        break;
      }
      s = (s + 1) & mask;
    }
  }
  cta.sync();
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    int isLeader = 0;
    if (!(c.lexicon && cur.lex[i] != 0)) {
      uint32_t s = (uint32_t)mix64((uint64_t)(uint32_t)cur.sid[i]) & mask;
      for (;;) {
        int occ = w.rowHash[s];
        if (cur.sid[occ] == cur.sid[i]) {
          w.rowOf[i] = occ;
          isLeader = occ == i;
          if (!isLeader) atomMin(&w.m2[occ], i);
          break;
        }
        s = (s + 1) & mask;
      }
    }
    w.rank[i] = isLeader;
  }
  cta.sync();
  ctaExclusiveScan(cta, w.rank, w.rankTmp, nH); // rank[i] = leaders before i; rank[nH] = rows
  for (int i = cta.tid; i < nH; i += cta.nthr)
    if (w.rowOf[i] == i) w.leaderOfRank[w.rank[i]] = i;
  if (cta.tid == 0) w.sc[SC_NROWS] = w.rank[nH];
  cta.sync();
}

FLT_DEV bool newTokenEligible(const DecCfg& c, const Beam& cur, int p, int n) {
  if (c.lexicon) return !c.ctc || cur.pb[p] || n != cur.tok[p]; // LexiconDecoder.cpp:89-90
  if (c.ctc) return n != cur.tok[p] || cur.pb[p];                // LexiconFreeDecoder.cpp:69-71
  return n != cur.tok[p];
}

FLT_DEV double transAdd(const DecCfg& c, const FrameIn& f, int n, int prevTok) {
  // returns the double `emittingModelScore` of the reference for token n after prevTok
  double am = (double)f.e[n];
  if (!c.ctc && !f.first && c.trans) am += (double)c.trans[(size_t)n * c.N + prevTok];
  return am;
}

// one wide cell: row led by hypothesis i, list column j
FLT_DEV void emitWide(const DecCfg& c, Ws& w, const Beam& cur, const FrameIn& f, int i, int j,
                      int slot, double& best) {
  w.cflag[slot] = 0;
  if (j >= f.listLen) return;
  const int n = f.topTok[j];
  if (n < 0) return;                 // short list (fewer eligible tokens than columns)
  if (c.ctc && n == c.blank) return; // blank is never a new token
  if (n == c.sil && c.silScore > 0) return; // boosted sil is not rank-dominated: emitSilCell
  int p = i;
  if (!newTokenEligible(c, cur, p, n)) {
    p = w.m2[i];
    if (p == kIntMax) return;
  }
  const float ev = f.topVal[j];
  if (!c.lexicon) {
    // LexiconFreeDecoder.cpp:64-85 with ZeroLM: score = prev + e (+sil) + lmWeight * 0
    double score = cur.score[p] + (double)ev;
    if (n == c.sil) score += c.silScore;
    score = score + c.lmWeight * (double)0.0f;
    putCand(w, slot, score, p, n, -1, 0, CF_NEW, 0.0f, n, ev);
    if (score > best) best = score;
  } else {
    // LexiconDecoder.cpp:62-110, prevLex == root, CTC (ranked mode excludes ASG)
    const int child = c.trie.rootChild[n];
    double score = cur.score[p] + (double)ev;
    if (n == c.sil) score += c.silScore;
    const float d = c.trie.maxScore[child] - 0.0f;
    score = score + c.lmWeight * (double)d;
    putCand(w, slot, score, p, n, -1, child, 0, d, -1, ev);
    if (score > best) best = score;
  }
}

// With silScore > 0 the sil expansion of a wide row is not dominated by the cells left of it in
// the ranked list, so every row proposes it explicitly (slot given by the caller).
FLT_DEV void emitSilCell(const DecCfg& c, Ws& w, const Beam& cur, const FrameIn& f, int i, int slot,
                         double& best) {
  w.cflag[slot] = 0;
  if (!(c.silScore > 0) || w.rowOf[i] != i) return;
  const int n = c.sil;
  if (n < 0 || n >= c.N || (c.ctc && n == c.blank)) return;
  if (!inTokenSet(c, f.e, n, f.thrVal, f.topTok, f.listLen)) return;
  int p = i;
  if (!newTokenEligible(c, cur, p, n)) {
    p = w.m2[i];
    if (p == kIntMax) return;
  }
  const float ev = f.e[n];
  if (!c.lexicon) {
    double score = cur.score[p] + (double)ev;
    score += c.silScore;
    score = score + c.lmWeight * (double)0.0f;
    putCand(w, slot, score, p, n, -1, 0, CF_NEW, 0.0f, n, ev);
    if (score > best) best = score;
  } else {
    const int child = c.trie.rootChild[n];
    if (child < 0 || c.trie.childOff[child + 1] == c.trie.childOff[child]) return;
    double score = cur.score[p] + (double)ev;
    score += c.silScore;
    const float d = c.trie.maxScore[child] - 0.0f;
    score = score + c.lmWeight * (double)d;
    putCand(w, slot, score, p, n, -1, child, 0, d, -1, ev);
    if (score > best) best = score;
  }
}

// stay / repeat and blank candidates of hypothesis i (slots base, base+1)
FLT_DEV void emitSpecials(const DecCfg& c, Ws& w, const Beam& cur, const FrameIn& f, int i, int base,
                          double& best) {
  w.cflag[base] = 0;
  w.cflag[base + 1] = 0;
  if (!c.lexicon) {
    // repeat (third branch, LexiconFreeDecoder.cpp:98-110): n == prevIdx and not a new token
    const int n = cur.tok[i];
    const bool isRepeat = c.ctc ? (!cur.pb[i] && n != c.blank) : true;
    if (isRepeat && n >= 0 && n < c.N && inTokenSet(c, f.e, n, f.thrVal, f.topTok, f.listLen)) {
      double score = cur.score[i] + (double)f.e[n];
      if (n == c.sil) score += c.silScore;
      putCand(w, base, score, i, n, -1, 0, 0, 0.0f, -1, f.e[n]);
      if (score > best) best = score;
    }
    if (c.ctc && c.blank >= 0 && c.blank < c.N &&
        inTokenSet(c, f.e, c.blank, f.thrVal, f.topTok, f.listLen)) {
      const int n = c.blank;
      double score = cur.score[i] + (double)f.e[n];
      if (n == c.sil) score += c.silScore;
      putCand(w, base + 1, score, i, n, -1, 0, CF_PB, 0.0f, -1, f.e[n]);
      if (score > best) best = score;
    }
  } else {
    const int lex = cur.lex[i];
    if (!c.ctc || !cur.pb[i] || lex == 0) { // (2) same node, LexiconDecoder.cpp:167-194
      const int n = lex == 0 ? c.sil : cur.tok[i];
      const double am = transAdd(c, f, n, cur.tok[i]);
      double score = cur.score[i] + am;
      if (n == c.sil) score += c.silScore;
      putCand(w, base, score, i, n, -1, lex, 0, 0.0f, -1, f.e[n]);
      if (score > best) best = score;
    }
    if (c.ctc) { // (3) blank, LexiconDecoder.cpp:196-213
      const int n = c.blank;
      double score = cur.score[i] + (double)f.e[n];
      putCand(w, base + 1, score, i, n, -1, lex, CF_PB, 0.0f, -1, f.e[n]);
      if (score > best) best = score;
    }
  }
}

FLT_DEV int allocCand(const DecCfg& c, Ws& w) {
  int s = atomAdd(&w.sc[SC_NCAND], 1);
  if (s >= c.capC) {
    w.sc[SC_OVF] = 1;
    return -1;
  }
  return s;
}

FLT_DEV float lmWordScore(const DecCfg& c, const Beam& cur, int p, int usrIdx) {
  if (c.lm.kind == 0) return 0.0f;
  const int wlm = c.lm.usr2lm[usrIdx];
  return ngramScore(c.lm, cur.ctx + (size_t)p * kMaxCtx, cur.nctx[p], wlm);
}

// one trie edge of hypothesis i: child node `child` reached by token n (LexiconDecoder.cpp:62-164)
FLT_DEV void emitEdge(const DecCfg& c, Ws& w, const Beam& cur, const FrameIn& f, int i, int n,
                      int child, bool labelsOnly, double& best) {
  if (!inTokenSet(c, f.e, n, f.thrVal, f.topTok, f.listLen)) return;
  const TrieDev& t = c.trie;
  const int lex = cur.lex[i];
  const float lexMax = lex == 0 ? 0.0f : t.maxScore[lex];
  const double am = transAdd(c, f, n, cur.tok[i]);
  double score = cur.score[i] + am;
  if (n == c.sil) score += c.silScore;
  const float ev = f.e[n];
  const bool hasKids = t.childOff[child + 1] > t.childOff[child];
  if (!labelsOnly && hasKids && newTokenEligible(c, cur, i, n)) {
    const float d = t.maxScore[child] - lexMax;
    const double s = score + c.lmWeight * (double)d;
    const int slot = allocCand(c, w);
    if (slot >= 0) putCand(w, slot, s, i, n, -1, child, 0, d, -1, ev);
    if (s > best) best = s;
  }
  const int l0 = t.labelOff[child], l1 = t.labelOff[child + 1];
  if (!(lex == 0 && cur.tok[i] == n)) { // LexiconDecoder.cpp:114-122
    for (int l = l0; l < l1; ++l) {
      const int label = t.labels[l];
      const float d = lmWordScore(c, cur, i, label) - lexMax;
      const double s = score + c.lmWeight * (double)d + c.wordScore;
      const int slot = allocCand(c, w);
      if (slot >= 0) putCand(w, slot, s, i, n, label, 0, CF_NEW, d, label, ev);
      if (s > best) best = s;
    }
  }
  if (l0 == l1 && c.hasUnk) { // LexiconDecoder.cpp:145-164
    const float d = lmWordScore(c, cur, i, c.unk) - lexMax;
    const double s = score + c.lmWeight * (double)d + c.unkScore;
    const int slot = allocCand(c, w);
    if (slot >= 0) putCand(w, slot, s, i, n, c.unk, 0, CF_NEW, d, c.unk, ev);
    if (s > best) best = s;
  }
}

// Phase M: merge candidates with equal (LM state, lex, token, prevBlank) keeping the best
// (Utils.h:168-198, max-merge). Representatives are collected into w.rep.
FLT_DEV void phaseMerge(const Cta& cta, const DecCfg& c, Ws& w, const Beam& cur, int nCand,
                        double thrScore) {
  for (int s = cta.tid; s < c.capH; s += cta.nthr) w.mh[s] = -1;
  if (cta.tid == 0) w.sc[SC_NREP] = 0;
  cta.sync();
  const uint32_t mask = (uint32_t)c.capH - 1;
  for (int x = cta.tid; x < nCand; x += cta.nthr) {
    if (!(w.cflag[x] & CF_ALIVE)) continue;
    if (!(w.cscore[x] >= thrScore)) { // Utils.h:161-165
      w.cflag[x] &= ~CF_ALIVE;
      continue;
    }
    uint32_t s = candKeyHash(w, cur, x) & mask;
    for (;;) {
      int occ = w.mh[s];
      if (occ == -1) {
        occ = atomCAS(&w.mh[s], -1, x);
        if (occ == -1) break;
      }
      if (candKeyEq(w, cur, occ, x)) {
        // same group: keep the better of the two in the slot
        while (candBetter(w, x, occ)) {
          int old = atomCAS(&w.mh[s], occ, x);
          if (old == occ) break;
          occ = old;
        }
        break;
      }
      s = (s + 1) & mask;
    }
  }
  cta.sync();
  for (int s = cta.tid; s < c.capH; s += cta.nthr) {
    const int x = w.mh[s];
    if (x >= 0) w.rep[atomAdd(&w.sc[SC_NREP], 1)] = x;
  }
  cta.sync();
}

// find, scanning bins from 255 down, the bin where the running count reaches `need`
FLT_DEV void findCutBin(const Cta& cta, Ws& w, int need) {
#if FLT_DEVICE_BUILD
  if (cta.tid < 32) {
    const int lane = cta.tid;
    int part = 0;
    for (int k = 0; k < 8; ++k) part += w.hist[255 - (lane * 8 + k)];
    int incl = part;
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    const int excl = incl - part;
    if (excl < need && incl >= need) {
      int cum = excl;
      for (int k = 0; k < 8; ++k) {
        const int b = 255 - (lane * 8 + k);
        const int h = w.hist[b];
        if (cum + h >= need) {
          w.sc[SC_BIN] = b;
          w.sc[SC_NEED] = need - cum;
          w.sc[SC_BINCOUNT] = h;
          break;
        }
        cum += h;
      }
    }
  }
#else
  if (cta.tid == 0) {
    int cum = 0;
    for (int b = 255; b >= 0; --b) {
      const int h = w.hist[b];
      if (cum + h >= need) {
        w.sc[SC_BIN] = b;
        w.sc[SC_NEED] = need - cum;
        w.sc[SC_BINCOUNT] = h;
        break;
      }
      cum += h;
    }
  }
#endif
}

// Phase Sel: choose the min(nRep, K) best representatives (Utils.h:200-220) into w.surv, then
// rank them (score descending, deterministic ties). Returns the number selected.
FLT_DEV int phaseSelect(const Cta& cta, const DecCfg& c, Ws& w, int nRep) {
  const int K = c.K;
  int nSel;
  if (nRep <= K) {
    for (int r = cta.tid; r < nRep; r += cta.nthr) w.surv[r] = w.rep[r];
    nSel = nRep;
    cta.sync();
  } else {
    // radix select on key - minKey, most significant differing byte first
    unsigned long long lmax = 0, lmin = ~0ull;
    for (int r = cta.tid; r < nRep; r += cta.nthr) {
      const unsigned long long k = orderedKey64(w.cscore[w.rep[r]]);
      lmax = k > lmax ? k : lmax;
      lmin = k < lmin ? k : lmin;
    }
    const unsigned long long kmax = ctaMax64(cta, lmax, w.red);
    const unsigned long long kmin = ~ctaMax64(cta, ~lmin, w.red);
    const unsigned long long range = kmax - kmin;
    int shift = 0;
    while (shift < 56 && (range >> shift) > 255ull) shift += 8;
    int need = K;
    unsigned long long prefix = 0; // bits above the current digit, already fixed
    bool wholeBin = false;
    if (cta.tid == 0) w.sc[SC_NSEL] = 0;
    for (;;) {
      for (int b = cta.tid; b < 256; b += cta.nthr) w.hist[b] = 0;
      cta.sync();
      for (int r = cta.tid; r < nRep; r += cta.nthr) {
        const unsigned long long k = orderedKey64(w.cscore[w.rep[r]]) - kmin;
        if (shift >= 56 || (k >> (shift + 8)) == (prefix >> (shift + 8)))
          atomAdd(&w.hist[(int)((k >> shift) & 255ull)], 1);
      }
      cta.sync();
      findCutBin(cta, w, need);
      cta.sync();
      const int bin = w.sc[SC_BIN];
      need = w.sc[SC_NEED];
      const int binCount = w.sc[SC_BINCOUNT];
      prefix |= (unsigned long long)bin << shift;
      if (binCount == need) {
        wholeBin = true;
        break;
      }
      if (shift == 0) break; // `need` of `binCount` identical keys: exact ties at the cut
      shift -= 8;
    }
    // keys strictly above the cut digit-prefix are selected; the cut bin is selected entirely
    // (wholeBin) or resolved among equals below.
    for (int r = cta.tid; r < nRep; r += cta.nthr) {
      const int x = w.rep[r];
      const unsigned long long k = (orderedKey64(w.cscore[x]) - kmin) >> shift;
      const unsigned long long p = prefix >> shift;
      if (k > p || (wholeBin && k == p)) w.surv[atomAdd(&w.sc[SC_NSEL], 1)] = x;
    }
    cta.sync();
    if (!wholeBin) {
      // rare: pick `need` of the equal-score groups by the deterministic order
      if (cta.tid == 0) {
        int n = w.sc[SC_NSEL];
        for (int q = 0; q < need; ++q) {
          int bestX = -1;
          for (int r = 0; r < nRep; ++r) {
            const int x = w.rep[r];
            if (((orderedKey64(w.cscore[x]) - kmin) >> shift) != (prefix >> shift)) continue;
            bool taken = false;
            for (int z = w.sc[SC_NSEL]; z < n; ++z) taken |= w.surv[z] == x;
            if (taken) continue;
            if (bestX < 0 || candBetter(w, x, bestX)) bestX = x;
          }
          w.surv[n++] = bestX;
        }
        w.sc[SC_NSEL] = n;
      }
      cta.sync();
    }
    nSel = w.sc[SC_NSEL];
  }
  // rank by counting (nSel <= K): position = number of strictly better survivors
  for (int a = cta.tid; a < nSel; a += cta.nthr) {
    const int x = w.surv[a];
    int pos = 0;
    for (int b = 0; b < nSel; ++b) pos += candBetter(w, w.surv[b], x) ? 1 : 0;
    w.survTmp[pos] = x;
  }
  cta.sync();
  return nSel;
}

// Phase F: materialise the new beam from the ranked survivors (in w.survTmp), intern new LM
// states, write the back-pointer records.
FLT_DEV void phaseFinalize(const Cta& cta, const DecCfg& c, Ws& w, const Beam& cur, Beam& nxt,
                           const FrameIn& f, int nSel, unsigned long long* stateTab,
                           long long stateCap, int* status) {
  for (int q = cta.tid; q < nSel; q += cta.nthr) {
    const int x = w.survTmp[q];
    const int p = w.cpar[x];
    const int fl = w.cflag[x];
    const int n = w.ctok[x];
    nxt.score[q] = w.cscore[x];
    if (fl & CF_FINISH) {
      nxt.am[q] = cur.am[p];
    } else {
      double am = (double)w.ce[x];
      if (!c.ctc && !f.first && c.trans) am += (double)c.trans[(size_t)n * c.N + cur.tok[p]];
      nxt.am[q] = cur.am[p] + am;
    }
    nxt.lm[q] = cur.lm[p] + (double)w.clmd[x];
    nxt.lex[q] = w.clex[x];
    nxt.tok[q] = n;
    nxt.pb[q] = (fl & CF_PB) ? 1 : 0;
    if (fl & CF_NEW) {
      const int lab = w.clab[x];
      // final states (decodeEnd) are never expanded again: no id needed
      int id = (fl & CF_FINISH) ? 0 : internState(stateTab, stateCap, cur.sid[p], lab);
      if (id < 0) {
        *status |= 2;
        id = 0;
      }
      nxt.sid[q] = id;
      nxt.spid[q] = cur.sid[p];
      nxt.slab[q] = lab;
      if (c.lm.kind) {
        const int wlm = lab < 0 ? c.lm.eos : c.lm.usr2lm[lab];
        nxt.nctx[q] = ngramAdvanceCtx(c.lm, cur.ctx + (size_t)p * kMaxCtx, cur.nctx[p], wlm,
                                      nxt.ctx + (size_t)q * kMaxCtx);
      }
    } else {
      nxt.sid[q] = cur.sid[p];
      nxt.spid[q] = cur.spid[p];
      nxt.slab[q] = cur.slab[p];
      if (c.lm.kind) {
        nxt.nctx[q] = cur.nctx[p];
        for (int k = 0; k < cur.nctx[p]; ++k)
          nxt.ctx[(size_t)q * kMaxCtx + k] = cur.ctx[(size_t)p * kMaxCtx + k];
      }
    }
    f.hParent[q] = p;
    f.hTok[q] = n;
    if (f.hWord) f.hWord[q] = w.cword[x];
  }
  if (cta.tid == 0) w.sc[SC_NH] = nSel;
  cta.sync();
}

// One frame: cur -> nxt. All threads of the CTA call this with identical arguments.
FLT_DEV void frameStep(const Cta& cta, const DecCfg& c, Ws& w, const Beam& cur, Beam& nxt,
                       const FrameIn& f, unsigned long long* stateTab, long long stateCap,
                       int* status) {
  const int nH = w.sc[SC_NH];
  if (nH == 0) return; // the beam died (Utils.h:155-158): every later frame is empty
  double best = negInf();
  const int K = c.K;

  int wideItems = 0;
  if (c.wideRanked) {
    phaseRows(cta, c, w, cur, nH);
    wideItems = c.wideOff[w.sc[SC_NROWS]];
  }
  const int specBase = wideItems;
  const int narrowBase = specBase + 3 * nH;
  if (cta.tid == 0) {
    w.sc[SC_NCAND] = narrowBase;
    w.sc[SC_OVF] = narrowBase > c.capC ? 1 : 0;
  }
  if (narrowBase > c.capC) { // cannot happen with a correctly sized capC; fail the utterance
    cta.sync();
    if (cta.tid == 0) {
      *status |= 1;
      w.sc[SC_NH] = 0;
    }
    cta.sync();
    return;
  }
  // wide cells
  if (c.wideRanked) {
    const int nRows = w.sc[SC_NROWS];
    for (int x = cta.tid; x < wideItems; x += cta.nthr) {
      const int r = searchOffsets(c.wideOff, nRows + 1, x); // row rank r (0-based)
      emitWide(c, w, cur, f, w.leaderOfRank[r], x - c.wideOff[r], x, best);
    }
  }
  // stay / repeat / blank
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    emitSpecials(c, w, cur, f, i, specBase + 3 * i, best);
    if (c.wideRanked) emitSilCell(c, w, cur, f, i, specBase + 3 * i + 2, best);
    else w.cflag[specBase + 3 * i + 2] = 0;
  }
  // trie edges
  if (c.lexicon) {
    const TrieDev& t = c.trie;
    for (int i = cta.tid; i < nH; i += cta.nthr) {
      const int lex = cur.lex[i];
      w.deg[i] = (c.wideRanked && lex == 0) ? t.nRootLab : t.childOff[lex + 1] - t.childOff[lex];
    }
    cta.sync();
    ctaExclusiveScan(cta, w.deg, w.degTmp, nH);
    const int items = w.deg[nH];
    for (int x = cta.tid; x < items; x += cta.nthr) {
      const int i = searchOffsets(w.deg, nH + 1, x);
      const int k = x - w.deg[i];
      const int lex = cur.lex[i];
      if (c.wideRanked && lex == 0) {
        const int n = t.rootLabTok[k];
        emitEdge(c, w, cur, f, i, n, t.rootChild[n], true, best);
      } else {
        const int e = t.childOff[lex] + k;
        emitEdge(c, w, cur, f, i, t.childTok[e], t.childNode[e], false, best);
      }
    }
  }
  const unsigned long long bestKey = ctaMax64(cta, orderedKey64(best), w.red);
  cta.sync();
  int nCand = w.sc[SC_NCAND];
  if (w.sc[SC_OVF]) {
    if (cta.tid == 0) *status |= 1;
    nCand = nCand < c.capC ? nCand : c.capC;
  }
  // candidatesBestScore_ - beamThreshold (LexiconDecoder.cpp:217-224)
  const double thrScore = keyToDouble(bestKey) - c.beamThreshold;
  phaseMerge(cta, c, w, cur, nCand, thrScore);
  const int nRep = w.sc[SC_NREP];
  const int nSel = phaseSelect(cta, c, w, nRep);
  phaseFinalize(cta, c, w, cur, nxt, f, nSel, stateTab, stateCap, status);
  (void)K;
}

// decodeEnd (LexiconFreeDecoder.cpp:127-158, LexiconDecoder.cpp:231-274) as one more "frame".
FLT_DEV void finishStep(const Cta& cta, const DecCfg& c, Ws& w, const Beam& cur, Beam& nxt,
                        const FrameIn& f, unsigned long long* stateTab, long long stateCap,
                        int* status) {
  const int nH = w.sc[SC_NH];
  if (nH == 0) return;
  if (cta.tid == 0) w.sc[SC_TIES] = 0;
  cta.sync();
  if (c.lexicon) {
    for (int i = cta.tid; i < nH; i += cta.nthr)
      if (cur.lex[i] == 0) w.sc[SC_TIES] = 1; // "nice ending" exists (benign same-value race)
    cta.sync();
  }
  const bool nice = c.lexicon && w.sc[SC_TIES] != 0;
  double best = negInf();
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    w.cflag[i] = 0;
    if (nice && cur.lex[i] != 0) continue;
    float ls = 0.0f;
    int flags = CF_FINISH;
    if (c.lm.kind) { // KenLM::finish: score </s>, state = child(-1)
      ls = ngramScore(c.lm, cur.ctx + (size_t)i * kMaxCtx, cur.nctx[i], c.lm.eos);
      flags |= CF_NEW;
    }
    const double score = cur.score[i] + c.lmWeight * (double)ls;
    putCand(w, i, score, i, c.sil, -1, cur.lex[i], flags, ls, -1, 0.0f);
    if (score > best) best = score;
  }
  const unsigned long long bestKey = ctaMax64(cta, orderedKey64(best), w.red);
  cta.sync();
  phaseMerge(cta, c, w, cur, nH, keyToDouble(bestKey) - c.beamThreshold);
  const int nSel = phaseSelect(cta, c, w, w.sc[SC_NREP]);
  phaseFinalize(cta, c, w, cur, nxt, f, nSel, stateTab, stateCap, status);
}

/* ------------------------------------------------------------------ whole-utterance driver ---- */
// One CTA decodes utterances bid, bid+nblk, ... start to finish.
FLT_DEV void decodeCta(const Cta& cta, const DecCfg& c, const BatchArgs& a, char* smem) {
  Ws w;
  char* base = a.useSmem ? smem : a.wsGlobal + (long long)cta.bid * a.wsStride;
  carveWs(base, c, w);
  unsigned long long* stateTab = a.stateTab + (long long)cta.bid * a.stateCap;
  const int K = c.K;
  for (int b = cta.bid; b < a.B; b += cta.nblk) {
    const int len = a.lengths ? a.lengths[b] : a.T;
    // reset the LM-state table and seed the beam (decodeBegin, LexiconDecoder.cpp:21-30)
    for (long long s = cta.tid; s < a.stateCap; s += cta.nthr) stateTab[s] = ~0ull;
    int curIdx = 0;
    if (cta.tid == 0) {
      Beam& B0 = w.beam[0];
      B0.score[0] = 0.0;
      B0.am[0] = 0.0;
      B0.lm[0] = 0.0;
      B0.sid[0] = 0;
      B0.spid[0] = -1;
      B0.slab[0] = -1;
      B0.lex[0] = 0;
      B0.tok[0] = c.sil;
      B0.pb[0] = 0;
      B0.nctx[0] = 0;
      if (c.lm.kind && c.lm.order > 1) {
        B0.ctx[0] = c.lm.bos;
        B0.nctx[0] = 1;
      }
      w.sc[SC_NH] = 1;
      a.status[b] = 0;
      int* hp = a.hParent + ((long long)b * (a.T + 2)) * K;
      hp[0] = -1;
      a.hTok[((long long)b * (a.T + 2)) * K] = c.sil;
      if (a.hWord) a.hWord[((long long)b * (a.T + 2)) * K] = -1;
    }
    cta.sync();
    for (int t = 0; t < len; ++t) {
      FrameIn f;
      const long long row = (long long)b * a.T + t;
      f.e = a.emis + row * c.N;
      f.topTok = a.topTok ? a.topTok + row * c.M : nullptr;
      f.topVal = a.topVal ? a.topVal + row * c.M : nullptr;
      f.listLen = c.M;
      f.thrVal = a.thrVal ? a.thrVal[row] : 0.0f;
      f.first = t == 0;
      const long long h = ((long long)b * (a.T + 2) + (t + 1)) * K;
      f.hParent = a.hParent + h;
      f.hTok = a.hTok + h;
      f.hWord = a.hWord ? a.hWord + h : nullptr;
      frameStep(cta, c, w, w.beam[curIdx], w.beam[curIdx ^ 1], f, stateTab, a.stateCap,
                a.status + b);
      if (w.sc[SC_NH] == 0) break;
      curIdx ^= 1;
      cta.sync();
    }
    int nFin = 0;
    if (w.sc[SC_NH] != 0) {
      FrameIn f;
      f.e = nullptr;
      f.topTok = nullptr;
      f.topVal = nullptr;
      f.listLen = 0;
      f.thrVal = 0.0f;
      f.first = 0;
      const long long h = ((long long)b * (a.T + 2) + (len + 1)) * K;
      f.hParent = a.hParent + h;
      f.hTok = a.hTok + h;
      f.hWord = a.hWord ? a.hWord + h : nullptr;
      finishStep(cta, c, w, w.beam[curIdx], w.beam[curIdx ^ 1], f, stateTab, a.stateCap,
                 a.status + b);
      curIdx ^= 1;
      nFin = w.sc[SC_NH];
    }
    cta.sync();
    const Beam& F = w.beam[curIdx];
    for (int q = cta.tid; q < nFin; q += cta.nthr) {
      double* o = a.finScore + ((long long)b * K + q) * 3;
      o[0] = F.score[q];
      o[1] = F.am[q];
      o[2] = F.lm[q];
    }
    if (cta.tid == 0) a.finCount[b] = nFin;
    cta.sync();
  }
}

/* ------------------------------------------------------------------ K4: n-best backtrace ------ */
struct BacktraceArgs {
  const int* hParent;
  const int* hTok;
  const int* hWord; // may be null
  const int* finCount;
  const int* lengths;
  int B, T, K, nbest;
  int* outTok;  // [B, nbest, T+2]
  int* outWord; // [B, nbest, T+2]
};

// item = (utterance b, rank r): walk the parent indices from the finish record to the seed
// (Utils.h:229-250); positions past len+1 are -1.
FLT_DEV void backtraceItem(const BacktraceArgs& a, long long item) {
  const int b = (int)(item / a.nbest), r = (int)(item % a.nbest);
  const int len = a.lengths ? a.lengths[b] : a.T;
  int* ot = a.outTok + item * (a.T + 2);
  int* ow = a.outWord + item * (a.T + 2);
  for (int i = len + 2; i < a.T + 2; ++i) {
    ot[i] = -1;
    ow[i] = -1;
  }
  if (r >= a.finCount[b]) {
    for (int i = 0; i < len + 2 && i < a.T + 2; ++i) {
      ot[i] = -1;
      ow[i] = -1;
    }
    return;
  }
  int k = r;
  for (int fidx = len + 1; fidx >= 0; --fidx) {
    const long long h = ((long long)b * (a.T + 2) + fidx) * a.K + k;
    ot[fidx] = a.hTok[h];
    ow[fidx] = a.hWord ? a.hWord[h] : -1;
    k = a.hParent[h];
  }
}

} // namespace flt
