// beam_core.h — the per-frame beam expansion (K2/K3), finish (decodeEnd) and n-best backtrace (K4)
// of the B200 decode path, one CTA per utterance.
//
// What it replaces (reference, flashlight/lib/text/decoder/...):
//   LexiconFreeDecoder::decodeStep  LexiconFreeDecoder.cpp:53-122
//   LexiconDecoder::decodeStep      LexiconDecoder.cpp:54-225
//   candidatesAdd / candidatesStore Utils.h:131-225   (threshold filter, key merge, top-K)
//   decodeEnd                       LexiconFreeDecoder.cpp:127-158, LexiconDecoder.cpp:231-274
//   getAllHypothesis                Utils.h:229-266
//   LMState::child identity         lm/LM.h:24-49
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
// Candidates are merged by key in a CTA-private hash table, the K best groups are found with a
// radix select on order-preserving 64-bit keys, and survivors are ranked.
//
// LM-state identity. The reference identifies an LM state by the address of a node in a child
// tree rooted at start() (lm/LM.h:24-49), i.e. by the sequence of labels that led to it. Here a
// state carries a 128-bit fingerprint of that label sequence (two independent 64-bit chains,
// extended with one multiply-mix per label): equal sequences give equal fingerprints at any time,
// with no table, no atomics and no per-utterance reset. Two different sequences collide with
// probability 2^-128 per pair.
//
// Arithmetic follows the reference's evaluation order (FP64 accumulators, FP32 sub-expressions,
// no FMA contraction: this file is compiled with -fmad=false).
#pragma once
#include "spmd.h"
#include "tables.h"

namespace flt {

typedef unsigned long long u64;

/* ------------------------------------------------------------------ configuration ---------- */
// Byte offsets of every workspace array from the CTA's workspace base. Computed once on the host
// (planFor) and passed in the kernel parameters, i.e. read from the constant bank.
struct Lay {
  int beamD[2], beamFp[2], beamI[2];
  int rowHash, rowI;
  int candScore, candKey, candI;
  int mh, rep, surv, skey, pos;
  int hist, sc, red;
  int list[2];
  int spec;
  int wideOff;
  int itemRow;   // int16 [wideOff[K]]: row rank of each wide work item
  int cslot;     // int [capC]: merge-table slot of each candidate
  int gath;      // int [64]: members of the cut bin (one-warp finish of the radix select)
  // lexicon-free fast step (beam_lf.h)
  int lfSlotB, lfSlotOf, lfCbin, lfAbove, lfDesc;
  int listNode, listMax; // int / float [Mwide] (lexicon, ranked rows): root child reached by each list
                         // token (-1: none or childless) and its smeared score, gathered once per frame
  int lnk, lhead; // int [capC] each (logAdd): members of a merge group chained behind its best member
  int pruneCache; // u8 per work item: 1 + best histogram bin its candidates reached in pass 1 (two-pass pruning)
  // single-pass step with a guessed cut (beam_gx.h)
  int beamX[2];   // int [3][K] per beam: Trie cache of each hypothesis (first edge, #edges | #labels, smeared score)
  int gxCandX;    // int [3][capC]: the same for each candidate's lex node
  int gxChunk;    // int2 [2][capChunks]: item chunk descriptors of each beam
  int gxBits;     // int [2][K]: hypotheses at the Trie root (walkers) of each beam
  int gxStash;    // u8 per work item: 1 + best histogram bin its proposals reached in the first pass
  int gxList;     // int4 [2][Mwide]: root child of each list entry (lexicon), double-buffered
  int gxBest;     // u64: best representative score key of the frame
  int total;      // bytes of both regions back to back (one region for the lexicon-free / single-pass steps)
  int small;      // bytes of the small region (Ws::base), 128-byte aligned
  int bigOff;     // offset of the capacity-sized region (Ws::big) from `base` when both sit together
  int bigBytes;   // its size: offsets of candScore / candKey / candI / mh / rep / cslot / lnk / lhead count from `big`
};

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
  int listInSmem; // the frame's token list is prefetched into the workspace (M <= 2 * threads)
  int capC;       // candidate capacity
  int capH;       // merge table slots (pow2 >= 2*capC)
  int capRH;      // row table slots (pow2 >= 2*K)
  int capP;       // pow2 >= K
  int wideTotal;  // wideOff[K]
  int prune2;     // 1 = two-pass histogram pruning of the candidates (lexicon decoder)
  int dbg;        // experiment switches (FLT_DBG): 1 = no direct-rank select (radix passes only)
  int pruneWant;  // candidates the kept bins must hold (>= K; the result is verified to hold K groups)
  int lfFast;     // 1 = lexicon-free fast step (beam_lf.h): cells indexed by hypothesis, no merge table
  int full;       // 1 = lexicon-free decoder expands every hypothesis x every token of the set (logAdd
                  //     merging or an n-gram token LM: no row / rank dominance to prune with)
  int rootList;   // 1 = lexicon decoder, beamSizeToken < N, no ranked rows: root hypotheses walk the
                  //     frame's token list (bst entries) instead of the root's ~N trie edges
  int wide;       // 1 = any of full / logAdd / lmToken / rootList: the kernels instantiated with W = true
                  //     (the max-merge kernels are compiled without those paths)
  int lmToken;    // 1 = LexiconDecoder with a token-level LM (isLmToken, LexiconDecoder.cpp:82-86)
  int lfBins;     // histogram bins of its select (pow2, multiple of 32)
  int gx;         // 1 = single-pass step with a guessed cut (beam_gx.h)
  int capChunks;  // its item chunk descriptors per beam
  int bigGlobal;  // 1 = the capacity-sized region (Ws::big) is the CTA's global slab, `base` is shared memory
  float lmUpper;  // n-gram LM: no word scores above this (log10): bound used to skip hopeless probes
  int nTau;       // pruning rectangles: rows 1..tauA[k] x columns 0..tauCol[k] hold >= K regular cells
  int tauA[16];
  int tauCol[16];
  const int* wideOff; // [K+1]: wideOff[r] = sum_{q=1..r} J_q, J_q = min(Mwide, K/q + 3 (+slack))
  const int* lfDesc;  // [wideTotal + 3K] work items of the lexicon-free step (beam_lf.h), column-major
  Lay lay;
  TrieDev trie;
  LmDev lm;
};

/* Per-launch arguments. */
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
  int* hSkip;           // [B, nCp, K] skip pointers: row 32j -> index of the ancestor in row 32(j-1)
  int* hSkipFin;        // [B, K]      finish row -> index of the ancestor in its checkpoint row
  int nCp;              // checkpoint rows per utterance: (T + 1) / 32 + 1
  double* finScore;     // [B, K, 3]
  int* finCount;        // [B]
  int* status;          // [B] bit0 = candidate overflow
  char* wsGlobal;       // per-CTA workspace slabs (used when the workspace does not fit smem)
  long long wsStride;
  unsigned long long* stats; // optional [4]: frames, candidates, merge groups, survivors (sums)
  // online decoding (Decoder.h:18-35): one utterance, the beam carried across launches
  char* streamBeam;     // saved beam (null = offline)
  int streamRestore;    // 1 = start from the saved beam instead of the seed (decodeBegin)
  int streamNoFinish;   // 1 = decodeStep chunk (no decodeEnd)
  int streamFrame0;     // frames decoded before this chunk (ASG transitions skip global frame 0)
  double streamShift;   // subtracted from the restored scores (prune()'s normalisation, Utils.h:333-341)
  double* hScore;       // [T+2, K, 3] score / emittingModelScore / lmScore per history record, or null
  int* hCount;          // [T+2] hypotheses per history row, or null
};

/* ------------------------------------------------------------------ workspace views ---------- */
struct Beam {
  double* d;  // [3][K] score, emittingModelScore, lmScore
  u64* fp;    // [2][K] LM-state fingerprint (+ [2][K] fingerprint of the parent state, beam_lf.h)
  int* iv;    // [5][K] lex, tok, prevBlank, nctx, anc; then (n-gram LM) ctx [K][kMaxCtx], the context's
              // back-off weights bo [K][kMaxCtx] (float bits) and their presence mask boMask [K]
  int K;
  int* xv = nullptr; // [3][K] Trie cache (beam_gx.h)
  FLT_DEV double& score(int i) const { return d[i]; }
  FLT_DEV double& am(int i) const { return d[K + i]; }
  FLT_DEV double& lm(int i) const { return d[2 * K + i]; }
  FLT_DEV u64& fpA(int i) const { return fp[i]; }
  FLT_DEV u64& fpB(int i) const { return fp[K + i]; }
  FLT_DEV u64& pfpA(int i) const { return fp[2 * K + i]; }
  FLT_DEV u64& pfpB(int i) const { return fp[3 * K + i]; }
  FLT_DEV int& lex(int i) const { return iv[i]; }
  FLT_DEV int& tok(int i) const { return iv[K + i]; }
  FLT_DEV int& pb(int i) const { return iv[2 * K + i]; }
  FLT_DEV int& nctx(int i) const { return iv[3 * K + i]; }
  FLT_DEV int& anc(int i) const { return iv[4 * K + i]; } // ancestor at the last checkpoint row (backtrace)
  FLT_DEV int* ctx(int i) const { return iv + 5 * K + i * kMaxCtx; }
  FLT_DEV float* bo(int i) const { return (float*)(iv + 5 * K + K * kMaxCtx) + i * kMaxCtx; } // tables.h
  FLT_DEV int& boMask(int i) const { return iv[5 * K + 2 * K * kMaxCtx + i]; }
};

constexpr int kPruneEdgeCap = 4096; // trie-edge work items whose pass-1 result is cached
constexpr int CF_PB = 1, CF_NEW = 2, CF_ALIVE = 4, CF_FINISH = 8;
constexpr int kIntMax = 0x7FFFFFFF;

struct Cand {
  double* sc; // [capC]
  u64* key;   // [2][capC] 128-bit merge key: (LM state, lex, token, prevBlank)
  int* iv;    // [6][capC] parent<<4|flags, token, e (float bits), word, lex, lm delta (float bits)
  int cap;
  FLT_DEV double& score(int x) const { return sc[x]; }
  FLT_DEV u64& keyA(int x) const { return key[x]; }
  FLT_DEV u64& keyB(int x) const { return key[cap + x]; }
  FLT_DEV int& parflag(int x) const { return iv[x]; }
  FLT_DEV int par(int x) const { return iv[x] >> 4; }
  FLT_DEV int flags(int x) const { return iv[x] & 15; }
  FLT_DEV int& tok(int x) const { return iv[cap + x]; }
  FLT_DEV float& ce(int x) const { return ((float*)iv)[2 * cap + x]; }
  FLT_DEV int& word(int x) const { return iv[3 * cap + x]; }
  FLT_DEV int& lex(int x) const { return iv[4 * cap + x]; }
  FLT_DEV float& lmd(int x) const { return ((float*)iv)[5 * cap + x]; }
};

struct Rows {
  int* hash; // [capRH]
  int* iv;
  int K;
  FLT_DEV int& rowOf(int i) const { return iv[i]; }
  FLT_DEV int& m2(int i) const { return iv[K + i]; }
  FLT_DEV int* rank() const { return iv + 2 * K; }             // [K+1]
  FLT_DEV int* rankTmp() const { return iv + 3 * K + 1; }      // [K+1]
  FLT_DEV int& leaderOfRank(int r) const { return iv[4 * K + 2 + r]; }
  FLT_DEV int* deg() const { return iv + 5 * K + 2; }          // [K+1]
  FLT_DEV int* degTmp() const { return iv + 6 * K + 3; }       // [K+1]
  FLT_DEV int& slotOf(int i) const { return iv[7 * K + 4 + i]; } // row-table slot of hyp i
  // [K] first Trie edge of hyp i's node. (64 ints further on: the scan's scratch degTmp() takes up to 64 ints
  // whatever K is and may run over slotOf(), which is dead by then — but must not reach this array.)
  FLT_DEV int* eoff() const { return iv + 8 * K + 4 + 64; }
};
constexpr int kRowsInts = 9; // K-sized int arrays (+ slack) behind Rows::iv

enum { // ws.sc[] scalars
  SC_NH = 0, SC_NCAND, SC_NREP, SC_NSEL, SC_NROWS, SC_OVF, SC_BIN, SC_NEED, SC_BINCOUNT, SC_NICE,
  SC_NGATH, SC_OR_LO, SC_OR_HI, SC_AND_LO, SC_AND_HI,
  SC_PMODE, SC_PCUT, SC_PLO_LO, SC_PLO_HI, SC_PSCALE, // two-pass candidate pruning (frameStep)
  SC_TLO, SC_THI, // previous phase stamp of thread 0 (counters on)
  SC_WANT, SC_WHOLD, // two-pass pruning: candidates to keep this frame; frames left at the wide setting
  SC_NACT, // active work items of the compacted second pruning pass
  SC_LFGAP, SC_LFFAC, SC_LFHOLD, // beam_lf.h guessed pruning bound: distance of the K-th best candidate below the
                                 // frame's upper bound in the previous frame (float bits, < 0: none yet), the
                                 // safety factor applied to it (float bits), frames left without guessing
  SC_GSPAN, SC_GHOLD, SC_GBIN, // guessed cut: kept score span below the top (float bits, 0 = none), frames left
                               // without guessing after a miss, this frame's guessed cut bin (0 = two passes)
  SC_GXCUTBIN, SC_GXKEPT, // beam_gx.h: cut bin of the exact redo and the proposals it keeps
  SC_GX,                  // beam_gx.h: two sets of scalars (2 x 8 ints)
  SC_WCNT = SC_GX + 16 /* 32 warp counters follow */,
  SC_COUNT = SC_WCNT + 32
};

// The workspace is addressed as base + constant-bank offset on every access (no pointer table in
// local memory; with the shared-memory base the compiler emits LDS/STS with immediate offsets).
// Two regions: `base` holds the small, hot arrays (beams, histogram, scalars, lists, row tables); `big` holds
// the arrays sized by the candidate capacity (candidate records, merge table, representatives). Both sit in
// shared memory when they fit (big = base + lay.bigOff); when the candidate capacity outgrows it, only `big`
// moves to the CTA's global slab (DecCfg::bigGlobal) and the per-item traffic of the expansion passes —
// histogram atomics, beam reads — stays on chip.
struct Ws {
  char* base;
  const DecCfg* c;
  int* itemBin = nullptr; // pass 1 of the two-pass pruning: best bin reached by the current work item
  char* big = nullptr;
  FLT_DEV Beam beam(int b) const {
    const Lay& L = c->lay;
    return Beam{(double*)(base + L.beamD[b]), (u64*)(base + L.beamFp[b]), (int*)(base + L.beamI[b]), c->K,
                (int*)(base + L.beamX[b])};
  }
  FLT_DEV Rows rows() const {
    return Rows{(int*)(base + c->lay.rowHash), (int*)(base + c->lay.rowI), c->K};
  }
  FLT_DEV Cand cand() const {
    const Lay& L = c->lay;
    return Cand{(double*)(big + L.candScore), (u64*)(big + L.candKey), (int*)(big + L.candI), c->capC};
  }
  FLT_DEV int* mh() const { return (int*)(big + c->lay.mh); }       // [capH] merge table
  FLT_DEV int* rep() const { return (int*)(big + c->lay.rep); }     // [capC] group representatives
  FLT_DEV int* surv() const { return (int*)(base + c->lay.surv); }   // [2][capP] selected / ranked
  FLT_DEV u64* skey() const { return (u64*)(base + c->lay.skey); }   // [capP] ordered score keys
  FLT_DEV int* pos() const { return (int*)(base + c->lay.pos); }     // [capP] rank counters
  FLT_DEV int* hist() const { return (int*)(base + c->lay.hist); }   // [256]
  FLT_DEV int* sc() const { return (int*)(base + c->lay.sc); }
  FLT_DEV u64* red() const { return (u64*)(base + c->lay.red); }     // [64]
  FLT_DEV int* listTok(int b) const { return (int*)(base + c->lay.list[b]); } // [M] token list
  FLT_DEV float* listVal(int b) const { return (float*)(base + c->lay.list[b] + 4 * c->M); }
  FLT_DEV float* spec() const { return (float*)(base + c->lay.spec); } // [K+2] e[own], e[blank], e[sil]
  FLT_DEV int* wideOff() const { return (int*)(base + c->lay.wideOff); } // [K+1]
  FLT_DEV short* itemRow() const { return (short*)(base + c->lay.itemRow); }
  FLT_DEV int* cslot() const { return (int*)(big + c->lay.cslot); }
  FLT_DEV int* gath() const { return (int*)(base + c->lay.gath); }
  FLT_DEV int* listNode() const { return (int*)(base + c->lay.listNode); }
  FLT_DEV float* listMax() const { return (float*)(base + c->lay.listMax); }
  FLT_DEV int* lnk() const { return (int*)(big + c->lay.lnk); }
  FLT_DEV int* lhead() const { return (int*)(big + c->lay.lhead); }
  FLT_DEV u64* rkey() const { return (u64*)(big + c->lay.candKey); } // reps' score keys (reuses keyA)
};

FLT_HD size_t alignUp(size_t x, size_t a) { return (x + a - 1) / a * a; }

FLT_HD void makeLayout(DecCfg& c) {
  size_t off = 0, offBig = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = alignUp(off + bytes, 16);
    return (int)o;
  };
  const int K = c.K;
  Lay& L = c.lay;
  const bool lf = c.lfFast != 0;
  const bool gx = c.gx != 0;
  const bool gxl = gx && c.lexicon;
  // the generic step keeps the capacity-sized arrays in a region of their own (Ws::big); the lexicon-free and
  // single-pass steps use one region (offsets from `base`, big == base)
  const bool split = !lf && !gx;
  auto takeBig = [&](size_t bytes) {
    if (!split) return take(bytes);
    const size_t o = offBig;
    offBig = alignUp(offBig + bytes, 16);
    return (int)o;
  };
  for (int b = 0; b < 2; ++b) {
    L.beamD[b] = take(sizeof(double) * 3 * K);
    L.beamFp[b] = take(sizeof(u64) * (lf ? 4 : 2) * K);
    L.beamI[b] = take(sizeof(int) * (5 * K + (c.lm.kind ? K * (2 * kMaxCtx + 1) : 0)));
    L.beamX[b] = take(gxl ? sizeof(int) * 3 * K : 0); // inside the beam block: saved / restored with it
  }
  L.rowHash = take(gx ? 0 : sizeof(int) * (lf ? 2 : 1) * c.capRH); // beam_lf.h: one fingerprint table per beam
  L.rowI = take(lf || gx ? 0 : sizeof(int) * (kRowsInts * K + 8 + 64));
  L.candScore = takeBig(sizeof(double) * c.capC);
  L.candKey = takeBig(sizeof(u64) * (lf ? 1 : 2) * c.capC);
  L.candI = takeBig(sizeof(int) * (lf ? 3 : 6) * c.capC);
  L.mh = takeBig(lf ? 0 : sizeof(int) * c.capH);
  L.rep = takeBig(gx ? 0 : sizeof(int) * c.capC);
  L.surv = take(gx ? 0 : sizeof(int) * 2 * c.capP);
  L.skey = take(lf ? 0 : sizeof(u64) * (gx ? c.capC : c.capP));
  L.pos = take(lf || gx ? 0 : sizeof(int) * c.capP);
  L.hist = take(sizeof(int) * (lf && c.lfBins > 256 ? c.lfBins : 256));
  L.sc = take(sizeof(int) * SC_COUNT);
  L.red = take(sizeof(u64) * 64);
  for (int b = 0; b < 2; ++b) L.list[b] = take(c.listInSmem ? 8 * (size_t)c.M : 0);
  L.spec = take(gx ? 0 : sizeof(float) * (K + 2));
  L.wideOff = take(gx ? 0 : sizeof(int) * (K + 1));
  L.itemRow = take(gx ? 0 : sizeof(short) * (c.wideTotal + 2));
  L.cslot = takeBig(lf ? 0 : sizeof(int) * c.capC);
  L.gath = take(gx ? 0 : sizeof(int) * 64);
  L.gxCandX = take(gxl ? sizeof(int) * 3 * c.capC : 0);
  L.gxChunk = take(gxl ? sizeof(int) * 2 * 2 * (size_t)c.capChunks : 0);
  L.gxBits = take(gxl ? sizeof(int) * 2 * K : 0);
  L.gxStash = take(gx ? (gxl ? (size_t)c.capChunks * 8 : 3 * (size_t)c.capP) : 0);
  L.gxList = take(gxl ? sizeof(int) * 4 * 2 * (size_t)c.Mwide : 0);
  L.gxBest = take(gx ? 16 : 0);
  L.lfSlotB = take(lf ? sizeof(int) * 2 * c.capRH : 0);
  L.lfSlotOf = take(lf ? sizeof(int) * 2 * K : 0);
  L.lfCbin = take(lf ? sizeof(unsigned short) * 2 * c.capC : 0); // bin, arrival order in the bin
  L.lfAbove = take(lf ? sizeof(unsigned short) * 16 * c.lfBins : 0); // one copy per warp (<= 16)
  L.lfDesc = take(lf ? sizeof(int) * c.capC : 0);                      // static work-item descriptors
  L.listNode = take(c.lexicon && c.wideRanked && !gx ? sizeof(int) * c.Mwide : 0);
  L.listMax = take(c.lexicon && c.wideRanked && !gx ? sizeof(float) * c.Mwide : 0);
  L.lnk = takeBig(c.logAdd ? sizeof(int) * c.capC : 0);
  L.lhead = takeBig(c.logAdd ? sizeof(int) * c.capC : 0);
  L.pruneCache = take(c.prune2 ? (size_t)c.wideTotal + c.K + kPruneEdgeCap : 0);
  L.small = (int)alignUp(off, 128);
  L.bigOff = split ? L.small : 0;
  L.bigBytes = (int)alignUp(offBig, 128);
  L.total = split ? L.small + L.bigBytes : (int)off; // both regions back to back
}

// the workspace views of a CTA: `base` = shared memory or the CTA's slab; with DecCfg::bigGlobal the
// capacity-sized region lives in the slab while `base` is shared memory
FLT_DEV Ws wsOf(char* base, const DecCfg& c, char* slab) {
  Ws w{base, &c};
  w.big = c.bigGlobal ? slab : base + c.lay.bigOff;
  return w;
}

FLT_HD double negInf() { return bitsF64(0xFFF0000000000000ull); }
FLT_HD double keyToDouble(u64 key) { // inverse of orderedKey64
  return bitsF64((key >> 63) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key);
}

/* ------------------------------------------------------------------ LM-state fingerprints ---- */
FLT_HD void fpRoot(u64& a, u64& b) {
  a = 0x243F6A8885A308D3ull;
  b = 0x13198A2E03707344ull;
}
// one 64-bit multiply + xor-shift per chain and label
FLT_HD void fpChild(u64 pa, u64 pb, int label, u64& a, u64& b) {
  const u64 l = (u64)(uint32_t)label;
  u64 x = (pa ^ (l + 0x632BE59BD9B4E019ull)) * 0x9E3779B97F4A7C15ull;
  u64 y = (pb + l * 0x165667B19E3779F9ull + 0x27D4EB2F165667C5ull) * 0xC2B2AE3D27D4EB4Full;
  a = x ^ (x >> 32);
  b = y ^ (y >> 29);
}
FLT_HD void candKeyOf(u64 sa, u64 sb, int lex, int tok, int pbFlag, u64& ka, u64& kb) {
  const u64 p = ((u64)(uint32_t)lex << 32) | ((u64)(uint32_t)tok << 1) | (u64)(pbFlag ? 1 : 0);
  u64 x = (sa ^ p) * 0xD6E8FEB86659FD93ull;
  u64 y = (sb + p) * 0xFF51AFD7ED558CCDull;
  ka = x ^ (x >> 31);
  kb = y ^ (y >> 30);
}

FLT_DEV int highBit64(u64 v) { // index of the most significant set bit (v != 0)
#if FLT_DEVICE_BUILD
  return 63 - __clzll((long long)v);
#else
  return 63 - __builtin_clzll(v);
#endif
}

// warp-aggregated counter increment: one atomic per converged group of callers
FLT_DEV int aggInc(int* ctr, int tid) {
#if FLT_DEVICE_BUILD
  const unsigned mask = __activemask();
  const int lane = tid & 31;
  const int leader = __ffs(mask) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(ctr, __popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
#else
  (void)tid;
  return (*ctr)++;
#endif
}

/* ------------------------------------------------------------------ CTA collectives ---------- */
// max and min over the CTA of 64-bit keys; every thread gets both.
FLT_DEV void ctaMaxMin64(const Cta& cta, u64& vmax, u64& vmin, u64* red) {
#if FLT_DEVICE_BUILD
  for (int o = 16; o > 0; o >>= 1) {
    const u64 u = __shfl_xor_sync(0xffffffffu, vmax, o);
    const u64 l = __shfl_xor_sync(0xffffffffu, vmin, o);
    vmax = u > vmax ? u : vmax;
    vmin = l < vmin ? l : vmin;
  }
  const int warp = cta.tid >> 5, lane = cta.tid & 31, nw = (cta.nthr + 31) >> 5;
  cta.sync(); // red[] may still be read from a previous call
  if (lane == 0) {
    red[warp] = vmax;
    red[32 + warp] = vmin;
  }
  cta.sync();
  u64 r = red[0], q = red[32];
  for (int i = 1; i < nw; ++i) {
    r = red[i] > r ? red[i] : r;
    q = red[32 + i] < q ? red[32 + i] : q;
  }
  vmax = r;
  vmin = q;
#else
  (void)cta;
  (void)red;
#endif
}
FLT_DEV u64 ctaMax64(const Cta& cta, u64 v, u64* red) {
  u64 mn = ~0ull;
  ctaMaxMin64(cta, v, mn, red);
  return v;
}

// Exclusive prefix sum of a[0..n) in place; a[n] receives the total. tmp has >= 64 ints.
// Each thread scans a contiguous slice, slices are combined with a shuffle scan (2 barriers).
FLT_DEV void ctaExclusiveScan(const Cta& cta, int* a, int* tmp, int n) {
#if FLT_DEVICE_BUILD
  const int per = (n + cta.nthr - 1) / cta.nthr;
  const int lo = cta.tid * per, hi = lo + per < n ? lo + per : n;
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  int incl = sum;
  const int lane = cta.tid & 31, warp = cta.tid >> 5, nw = (cta.nthr + 31) >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) tmp[warp] = incl;
  cta.sync();
  int base = 0;
  for (int wI = 0; wI < warp; ++wI) base += tmp[wI];
  int total = 0;
  for (int wI = 0; wI < nw; ++wI) total += tmp[wI];
  int run = base + incl - sum;
  for (int i = lo; i < hi; ++i) {
    const int v = a[i];
    a[i] = run;
    run += v;
  }
  if (cta.tid == 0) a[n] = total;
  cta.sync();
#else
  (void)tmp;
  int run = 0;
  for (int i = 0; i < n; ++i) {
    const int v = a[i];
    a[i] = run;
    run += v;
  }
  a[n] = run;
  (void)cta;
#endif
}

// largest r in [0, n) with off[r] <= x, for a non-decreasing off[0..n)
FLT_DEV int searchOffsets(const int* off, int n, int x) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= x) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

/* ------------------------------------------------------------------ candidates ---------- */
// label that names the candidate's new LM state: token (lexicon-free), word (lexicon), -1 (finish)
FLT_DEV int candLabel(const DecCfg& c, const Cand& cd, int x) {
  if (cd.flags(x) & CF_FINISH) return -1;
  return (c.lexicon && !c.lmToken) ? cd.word(x) : cd.tok(x);
}

FLT_DEV void putCand(const DecCfg& c, const Ws& w, const Beam& cur, int slot, double score, int par,
                     int tok, int word, int lex, int flags, float lmd, float ev) {
  const Cand cd = w.cand();
  cd.score(slot) = score;
  cd.parflag(slot) = (par << 4) | flags | CF_ALIVE;
  cd.tok(slot) = tok;
  cd.word(slot) = word;
  cd.lex(slot) = lex;
  cd.lmd(slot) = lmd;
  cd.ce(slot) = ev;
  u64 sa = cur.fpA(par), sb = cur.fpB(par);
  if (flags & CF_NEW) {
    const int label = (flags & CF_FINISH) ? -1 : ((c.lexicon && !c.lmToken) ? word : tok);
    fpChild(sa, sb, label, sa, sb);
  }
  candKeyOf(sa, sb, lex, tok, flags & CF_PB, cd.keyA(slot), cd.keyB(slot));
}

// deterministic total order used wherever the reference leaves ties to libstdc++ internals:
// higher score first, then lower parent rank, token, word, lexicon node, prevBlank.
FLT_DEV bool candBetter(const Cand& cd, int a, int b) {
  const double sa = cd.score(a), sb = cd.score(b);
  if (sa != sb) return sa > sb;
  if (cd.par(a) != cd.par(b)) return cd.par(a) < cd.par(b);
  if (cd.tok(a) != cd.tok(b)) return cd.tok(a) < cd.tok(b);
  if (cd.word(a) != cd.word(b)) return cd.word(a) < cd.word(b);
  if (cd.lex(a) != cd.lex(b)) return cd.lex(a) < cd.lex(b);
  return (cd.flags(a) & CF_PB) < (cd.flags(b) & CF_PB);
}

/* ------------------------------------------------------------------ backtrace skip pointers --- */
// A final hypothesis is traced back through T+2 history rows (Utils.h:229-250): a chain of T+2
// dependent loads. Every hypothesis therefore carries the index of its ancestor in the last
// checkpoint row (rows 0, 32, 64, ...); checkpoint rows and the finish row store it, and the
// backtrace first hops checkpoint to checkpoint (T/32 loads), then walks all 32-row segments in
// parallel.
constexpr int kCpShift = 5, kCpRows = 1 << kCpShift;
// ancestor index carried by the hypothesis written to history row hRow with parent p
FLT_DEV int skipCarry(const Beam& cur, int hRow, int p) {
  return ((hRow - 1) & (kCpRows - 1)) == 0 ? p : cur.anc(p);
}

/* ------------------------------------------------------------------ the frame step ---------- */
struct FrameIn {
  const float* e;      // emission row [N] (global)
  const int* topTok;   // [M] ranked tokens (workspace copy when listInSmem, else global)
  const float* topVal; // [M]
  int listLen;         // valid entries in the list
  float thrVal;        // cut value of the token set (unused when setAll)
  int first;           // global frame 0 (ASG transitions are skipped, LexiconDecoder.cpp:70-73)
  int listIsSet;       // the list holds the whole token set (lexicon-free, beamSizeToken < N)
  int specReady;       // spec[] (beam_lf.h) / eBlank, eSil (beam_gx.h) already hold this frame's emissions
  float eBlank, eSil;  // e[blank], e[sil] of this frame when specReady (beam_gx.h, fused kernel)
  const int* listInfo; // int4 per list entry: root child of the token (beam_gx.h, lexicon)
  const float* eNext;  // next frame's emission row or null (beam_lf.h prefetches its gathers)
  int hRow;            // index of the history row being written (frame t+1; len+1 for the finish)
  int* hSkip;          // where its skip pointers go: a checkpoint row (hRow % 32 == 0), the finish row, else null
  double* hScore;      // this row's [K,3] scores (online decoding) or null
  int* hCount;         // this row's hypothesis count (online decoding) or null
  int* hParent;        // history row to write (frame t+1), [K]
  int* hTok;
  int* hWord;
};

/* token-set membership for tokens that are not taken from the ranked list */
FLT_DEV bool inTokenSetV(const DecCfg& c, const FrameIn& f, int n, float v) {
  if (c.setAll) return true;
  if (v > f.thrVal) return true;
  if (v < f.thrVal) return false;
  if (!f.listIsSet) return true; // lexicon decoder: ties at the cut are taken (documented)
  for (int j = 0; j < f.listLen; ++j) // equal to the cut value: membership = presence in the list
    if (f.topTok[j] == n) return true;
  return false;
}

// Phase R: group wide hypotheses into rows (same LM state; lexicon: also lex == root), find each
// row's best and second-best member, and rank the rows by their best member. The row table is
// empty on entry and left empty (each leader clears its own slot). Ends with a barrier.
// `tail` runs between the last two barriers (used to publish per-frame scalars for free).
template <class Tail>
FLT_DEV void phaseRows(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, int nH, Tail tail) {
  const Rows R = w.rows();
  int* sc = w.sc();
  const uint32_t mask = (uint32_t)c.capRH - 1;
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    R.m2(i) = kIntMax;
    R.rowOf(i) = -1;
    if (c.lexicon && cur.lex(i) != 0) continue;
    const u64 fa = cur.fpA(i), fb = cur.fpB(i);
    uint32_t s = (uint32_t)fa & mask;
    for (;;) {
      const int old = atomCAS(&R.hash[s], -1, i);
      if (old == -1) break;
      if (cur.fpA(old) == fa && cur.fpB(old) == fb) {
        atomMin(&R.hash[s], i);
        break;
      }
      s = (s + 1) & mask;
    }
    R.slotOf(i) = (int)s;
  }
  cta.sync();
  // leaders, second members, row ranks (hypotheses are sorted by score: rank = leaders before me)
  const bool small = nH <= cta.nthr;
  int* rank = R.rank();
  int myRank = 0, isLeader = 0;
  for (int i = cta.tid; i < (small ? cta.nthr : nH); i += cta.nthr) {
    isLeader = 0;
    if (i < nH && !(c.lexicon && cur.lex(i) != 0)) {
      const int occ = R.hash[R.slotOf(i)];
      R.rowOf(i) = occ;
      isLeader = occ == i;
      if (!isLeader) atomMin(&R.m2(occ), i);
    }
#if FLT_DEVICE_BUILD
    if (small) {
      const unsigned bal = __ballot_sync(0xffffffffu, isLeader);
      myRank = __popc(bal & ((1u << (cta.tid & 31)) - 1u));
      if ((cta.tid & 31) == 0) sc[SC_WCNT + (cta.tid >> 5)] = __popc(bal);
    } else
#endif
    {
      if (i < nH) rank[i] = isLeader;
    }
  }
  cta.sync();
#if FLT_DEVICE_BUILD
  if (small) {
    const int warp = cta.tid >> 5, nw = (cta.nthr + 31) >> 5;
    int before = 0, total = 0;
    for (int k = 0; k < nw; ++k) {
      const int v = sc[SC_WCNT + k];
      before += k < warp ? v : 0;
      total += v;
    }
    if (isLeader) {
      R.leaderOfRank(before + myRank) = cta.tid;
      R.hash[R.slotOf(cta.tid)] = -1;
    }
    if (cta.tid == 0) sc[SC_NROWS] = total;
  } else
#endif
  {
    (void)myRank;
    ctaExclusiveScan(cta, rank, R.rankTmp(), nH); // rank[i] = leaders before i; rank[nH] = rows
    for (int i = cta.tid; i < nH; i += cta.nthr)
      if (R.rowOf(i) == i) {
        R.leaderOfRank(rank[i]) = i;
        R.hash[R.slotOf(i)] = -1;
      }
    if (cta.tid == 0) sc[SC_NROWS] = rank[nH];
  }
  tail();
  cta.sync();
}

FLT_DEV bool newTokenEligible(const DecCfg& c, const Beam& cur, int p, int n) {
  if (c.lexicon) return !c.ctc || cur.pb(p) || n != cur.tok(p); // LexiconDecoder.cpp:89-90
  if (c.ctc) return n != cur.tok(p) || cur.pb(p);                // LexiconFreeDecoder.cpp:69-71
  return n != cur.tok(p);
}

// the reference's double `emittingModelScore` for an emission value ev of token n after prevTok
FLT_DEV double amOf(const DecCfg& c, const FrameIn& f, float ev, int n, int prevTok) {
  double am = (double)ev;
  if (!c.ctc && !f.first && c.trans) am += (double)c.trans[(size_t)n * c.N + prevTok];
  return am;
}

// Two-pass pruning (lexicon decoder): pass 1 only histograms the candidate scores (mode 1), pass 2
// materialises the candidates whose bin is at or above the cut (mode 2). The map score -> bin is
// monotone, so everything left out scores no higher than anything kept; frameStep verifies that
// the kept set holds >= K merge groups (else it repeats the frame without the cut).
constexpr int kPruneBins = 256;
FLT_DEV int pruneBin(const int* sc, double score) {
  const double lo = bitsF64(((u64)(unsigned)sc[SC_PLO_HI] << 32) | (unsigned)sc[SC_PLO_LO]);
  const float pos = (float)(score - lo) * bitsF32((uint32_t)sc[SC_PSCALE]);
  // bins 0..254: 1 + bin fits the one-byte per-item cache of the two-pass pruning
  return pos >= (float)(kPruneBins - 2) ? kPruneBins - 2 : (pos > 0.0f ? (int)pos : 0);
}

// slot for a candidate that survived the pruning bound (compact: only live candidates are stored)
FLT_DEV int allocCand(const Cta& cta, const DecCfg& c, const Ws& w, double score) {
  const int mode = w.sc()[SC_PMODE];
  if (mode) {
    const int bin = pruneBin(w.sc(), score);
    if (mode == 1) {
      atomAdd(&w.hist()[bin], 1); // (warp-aggregating these with match.any was measured slower)
      if (w.itemBin && bin > *w.itemBin) *w.itemBin = bin;
      return -1;
    }
    if (bin < w.sc()[SC_PCUT]) return -1;
  }
  const int s = aggInc(&w.sc()[SC_NCAND], cta.tid);
  if (s >= c.capC) {
    w.sc()[SC_OVF] = 1;
    return -1;
  }
  return s;
}

// new-token expansion of the row led by hypothesis i with token n (value ev)
// (j = column of n in the frame's ranked list, whose root child is cached in the workspace; -1 = n is
// not taken from the list)
FLT_DEV void emitRowToken(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, int i, int n,
                          float ev, double tau, int j = -1) {
  int p = i;
  if (!newTokenEligible(c, cur, p, n)) {
    p = w.rows().m2(i);
    if (p == kIntMax) return;
  }
  if (!c.lexicon) {
    // LexiconFreeDecoder.cpp:64-85 with ZeroLM: score = prev + e (+sil) + lmWeight * 0
    double score = cur.score(p) + (double)ev;
    if (n == c.sil) score += c.silScore;
    score = score + c.lmWeight * (double)0.0f;
    if (score < tau) return;
    const int slot = allocCand(cta, c, w, score);
    if (slot >= 0) putCand(c, w, cur, slot, score, p, n, -1, 0, CF_NEW, 0.0f, ev);
  } else {
    // LexiconDecoder.cpp:62-110, prevLex == root, CTC (ranked mode excludes ASG)
    int child;
    float ms;
    if (j >= 0) { // gathered once per frame for the whole list (frameStep)
      child = w.listNode()[j];
      if (child < 0) return;
      ms = w.listMax()[j];
    } else {
      child = c.trie.rootChild[n];
      if (child < 0 || c.trie.childOff[child + 1] == c.trie.childOff[child]) return;
      ms = c.trie.maxScore[child];
    }
    double score = cur.score(p) + (double)ev;
    if (n == c.sil) score += c.silScore;
    const float d = ms - 0.0f;
    score = score + c.lmWeight * (double)d;
    if (score < tau) return;
    const int slot = allocCand(cta, c, w, score);
    if (slot >= 0) putCand(c, w, cur, slot, score, p, n, -1, child, 0, d, ev);
  }
}

// one wide cell: row led by hypothesis i, list column j
FLT_DEV void emitWide(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const FrameIn& f,
                      int i, int j, double tau) {
  if (j >= f.listLen) return;
  const int n = f.topTok[j];
  if (n < 0) return;                 // short list (fewer eligible tokens than columns)
  if (c.ctc && n == c.blank) return; // blank is never a new token
  if (n == c.sil && c.silScore > 0) return; // boosted sil is not rank-dominated: emitSilCell
  emitRowToken(cta, c, w, cur, i, n, f.topVal[j], tau, c.lexicon ? j : -1);
}

// With silScore > 0 the sil expansion of a wide row is not dominated by the cells left of it in
// the ranked list, so every row proposes it explicitly.
FLT_DEV void emitSilCell(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur,
                         const FrameIn& f, int i, double tau) {
  if (!(c.silScore > 0) || w.rows().rowOf(i) != i) return;
  const int n = c.sil;
  if (n < 0 || n >= c.N || (c.ctc && n == c.blank)) return;
  const float ev = w.spec()[c.K + 1];
  if (!inTokenSetV(c, f, n, ev)) return;
  emitRowToken(cta, c, w, cur, i, n, ev, tau);
}

// token whose emission hypothesis i needs for its stay / repeat candidate
FLT_DEV int ownToken(const DecCfg& c, const Beam& cur, int i) {
  return (c.lexicon && cur.lex(i) == 0) ? c.sil : cur.tok(i);
}

// stay / repeat and blank candidates of hypothesis i
FLT_DEV void emitSpecials(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur,
                          const FrameIn& f, int i, double tau) {
  const float eOwn = w.spec()[i], eBlank = w.spec()[c.K];
  if (!c.lexicon) {
    // repeat (third branch, LexiconFreeDecoder.cpp:98-110): n == prevIdx and not a new token
    const int n = cur.tok(i);
    const bool isRepeat = c.ctc ? (!cur.pb(i) && n != c.blank) : true;
    if (isRepeat && n >= 0 && n < c.N && inTokenSetV(c, f, n, eOwn)) {
      double score = cur.score(i) + (double)eOwn;
      if (n == c.sil) score += c.silScore;
      if (!(score < tau)) {
        const int slot = allocCand(cta, c, w, score);
        if (slot >= 0) putCand(c, w, cur, slot, score, i, n, -1, 0, 0, 0.0f, eOwn);
      }
    }
    if (c.ctc && inTokenSetV(c, f, c.blank, eBlank)) {
      const int n = c.blank;
      double score = cur.score(i) + (double)eBlank;
      if (n == c.sil) score += c.silScore;
      if (!(score < tau)) {
        const int slot = allocCand(cta, c, w, score);
        if (slot >= 0) putCand(c, w, cur, slot, score, i, n, -1, 0, CF_PB, 0.0f, eBlank);
      }
    }
  } else {
    const int lex = cur.lex(i);
    if (!c.ctc || !cur.pb(i) || lex == 0) { // (2) same node, LexiconDecoder.cpp:167-194
      const int n = lex == 0 ? c.sil : cur.tok(i);
      const double am = amOf(c, f, eOwn, n, cur.tok(i));
      double score = cur.score(i) + am;
      if (n == c.sil) score += c.silScore;
      if (!(score < tau)) {
        const int slot = allocCand(cta, c, w, score);
        if (slot >= 0) putCand(c, w, cur, slot, score, i, n, -1, lex, 0, 0.0f, eOwn);
      }
    }
    if (c.ctc) { // (3) blank, LexiconDecoder.cpp:196-213
      const double score = cur.score(i) + (double)eBlank;
      if (!(score < tau)) {
        const int slot = allocCand(cta, c, w, score);
        if (slot >= 0) putCand(c, w, cur, slot, score, i, c.blank, -1, lex, CF_PB, 0.0f, eBlank);
      }
    }
  }
}

FLT_DEV float lmWordScore(const DecCfg& c, const Beam& cur, int p, int usrIdx) {
  if (c.lm.kind == 0) return 0.0f;
  // the back-offs of hypothesis p's LM state were looked up when the state was created (phaseFinalize)
  return ngramScoreCached(c.lm, cur.ctx(p), cur.nctx(p), c.lm.usr2lm[usrIdx], cur.bo(p), cur.boMask(p));
}

// one trie edge of hypothesis i: child node `child` reached by token n (LexiconDecoder.cpp:62-164)
template <bool W>
FLT_DEV void emitEdge(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, const FrameIn& f,
                      int i, int n, int child, bool labelsOnly, double tau) {
  const float ev = f.e[n];
  if (!inTokenSetV(c, f, n, ev)) return;
  const TrieDev& t = c.trie;
  const int lex = cur.lex(i);
  const float lexMax = lex == 0 ? 0.0f : t.maxScore[lex];
  const double am = amOf(c, f, ev, n, cur.tok(i));
  double score = cur.score(i) + am;
  if (n == c.sil) score += c.silScore;
  const bool hasKids = t.childOff[child + 1] > t.childOff[child];
  const int l0 = t.labelOff[child], l1 = t.labelOff[child + 1];
  if (W && c.lmToken) {
    // token-level LM (LexiconDecoder.cpp:82-86): one LM step per trie edge, shared by the inner-node
    // candidate, the word ends and unk; the new LM state is child(state, n) for all of them
    const float ls = lmWordScore(c, cur, i, n);
    const double base = score + c.lmWeight * (double)ls;
    if (hasKids && newTokenEligible(c, cur, i, n) && !(base < tau)) {
      const int slot = allocCand(cta, c, w, base);
      if (slot >= 0) putCand(c, w, cur, slot, base, i, n, -1, child, CF_NEW, ls, ev);
    }
    if (!(lex == 0 && cur.tok(i) == n)) {
      for (int l = l0; l < l1; ++l) {
        const double s = base + c.wordScore;
        if (s < tau) continue;
        const int slot = allocCand(cta, c, w, s);
        if (slot >= 0) putCand(c, w, cur, slot, s, i, n, t.labels[l], 0, CF_NEW, ls, ev);
      }
    }
    if (l0 == l1 && c.hasUnk) {
      const double s = base + c.unkScore;
      if (!(s < tau)) {
        const int slot = allocCand(cta, c, w, s);
        if (slot >= 0) putCand(c, w, cur, slot, s, i, n, c.unk, 0, CF_NEW, ls, ev);
      }
    }
    return;
  }
  if (!labelsOnly && hasKids && newTokenEligible(c, cur, i, n)) {
    const float d = t.maxScore[child] - lexMax;
    const double s = score + c.lmWeight * (double)d;
    if (!(s < tau)) {
      const int slot = allocCand(cta, c, w, s);
      if (slot >= 0) putCand(c, w, cur, slot, s, i, n, -1, child, 0, d, ev);
    }
  }
  // n-gram word LM under the two-pass pruning: a word end's score is bounded from above with the LM's best
  // possible score (c.lmUpper) BEFORE the table probes (a chain of dependent DRAM reads each). Every step from
  // the LM score to the bin is monotone in floating point (float subtraction of lexMax, product with
  // lmWeight >= 0, the two additions, pruneBin), so bin(bound) >= bin(score):
  //   pass 1  the word end is not probed and not counted; only its bound's bin goes into the item's cache byte.
  //           The cut is then chosen from the other candidates alone and can only sit lower than with the word
  //           ends counted: pass 2 keeps a superset of what it kept before — still everything at or above the
  //           cut, so the frame stays exact (and is verified to hold >= K groups as before);
  //   pass 2  probed only if the bound reaches the cut.
  const int pmode = w.sc()[SC_PMODE];
  const bool lmBound = c.lm.kind != 0 && pmode != 0 && c.lmWeight >= 0;
  auto boundSkips = [&](double extra) __attribute__((always_inline)) { // true: this word end needs no probe now
    if (!lmBound) return false;
    const float dub = c.lmUpper - lexMax;
    const double sub = score + c.lmWeight * (double)dub + extra;
    if (sub < tau) return true;
    const int bub = pruneBin(w.sc(), sub);
    if (pmode == 1) {
      if (w.itemBin && bub > *w.itemBin) *w.itemBin = bub;
      return true;
    }
    return bub < w.sc()[SC_PCUT];
  };
  if (!(lex == 0 && cur.tok(i) == n)) { // LexiconDecoder.cpp:114-122
    for (int l = l0; l < l1; ++l) {
      if (boundSkips(c.wordScore)) break; // the bound is the same for every label of this node
      const int label = t.labels[l];
      const float d = lmWordScore(c, cur, i, label) - lexMax;
      const double s = score + c.lmWeight * (double)d + c.wordScore;
      if (s < tau) continue;
      const int slot = allocCand(cta, c, w, s);
      if (slot >= 0) putCand(c, w, cur, slot, s, i, n, label, 0, CF_NEW, d, ev);
    }
  }
  if (l0 == l1 && c.hasUnk && !boundSkips(c.unkScore)) { // LexiconDecoder.cpp:145-164
    const float d = lmWordScore(c, cur, i, c.unk) - lexMax;
    const double s = score + c.lmWeight * (double)d + c.unkScore;
    if (!(s < tau)) {
      const int slot = allocCand(cta, c, w, s);
      if (slot >= 0) putCand(c, w, cur, slot, s, i, n, c.unk, 0, CF_NEW, d, ev);
    }
  }
}

// Full expansion of the lexicon-free decoder (LexiconFreeDecoder.cpp:53-112): hypothesis i with
// token n of the token set (value ev). Used when nothing can be pruned by rank: logAdd merging (every
// member of a group counts) or an n-gram token LM (the LM score differs per hypothesis and token).
FLT_DEV void emitFullLf(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur, int i, int n,
                        float ev, double tau) {
  const int prevTok = cur.tok(i);
  double score = cur.score(i) + (double)ev; // transitions reach only emittingModelScore (:59-64)
  if (n == c.sil) score += c.silScore;
  const bool isNew = c.ctc ? (n != c.blank && (n != prevTok || cur.pb(i))) : n != prevTok;
  if (isNew) {
    const float ls = lmWordScore(c, cur, i, n);
    score = score + c.lmWeight * (double)ls;
    if (score < tau) return;
    const int slot = allocCand(cta, c, w, score);
    if (slot >= 0) putCand(c, w, cur, slot, score, i, n, -1, 0, CF_NEW, ls, ev);
  } else {
    if (score < tau) return;
    const int slot = allocCand(cta, c, w, score);
    if (slot >= 0) putCand(c, w, cur, slot, score, i, n, -1, 0, (c.ctc && n == c.blank) ? CF_PB : 0, 0.0f, ev);
  }
}

// Phase M: merge candidates with equal (LM state, lex, token, prevBlank) keeping the best
// (Utils.h:168-198, max-merge). The merge table is empty on entry; representatives (the group
// maxima) are collected into rep[] with their order-preserving score keys in rkey[], and the
// AND / OR of all keys is accumulated for the radix select. Two barriers.
template <bool W>
FLT_DEV void phaseMerge(const Cta& cta, const DecCfg& c, const Ws& w, int nCand) {
  const bool logAdd = W && c.logAdd;
  const Cand cd = w.cand();
  int* mh = w.mh();
  int* cslot = w.cslot();
  int* sc = w.sc();
  const uint32_t mask = (uint32_t)c.capH - 1;
  for (int x = cta.tid; x < nCand; x += cta.nthr) {
    if (!(cd.parflag(x) & CF_ALIVE)) continue;
    const u64 ka = cd.keyA(x), kb = cd.keyB(x);
    uint32_t s = (uint32_t)ka & mask;
    for (;;) {
      int occ = mh[s];
      if (occ == -1) {
        occ = atomCAS(&mh[s], -1, x);
        if (occ == -1) break;
      }
      if (cd.keyA(occ) == ka && cd.keyB(occ) == kb) {
        // same group: keep the better of the two in the slot
        while (candBetter(cd, x, occ)) {
          const int old = atomCAS(&mh[s], occ, x);
          if (old == occ) break;
          occ = old;
        }
        break;
      }
      s = (s + 1) & mask;
    }
    cslot[x] = (int)s;
    if (logAdd) w.lhead()[x] = -1;
  }
  cta.sync();
  // logAdd (Utils.h:161-165,185-195): candidates below best - beamThreshold are dropped BEFORE the
  // merge; the others of a group are chained behind its best member, which then adds them up in
  // descending score order exactly as the reference's sorted run does
  double thrPre = negInf();
  if (logAdd) {
    u64 mx = 0;
    for (int x = cta.tid; x < nCand; x += cta.nthr)
      if (cd.parflag(x) & CF_ALIVE) {
        const u64 k = orderedKey64(cd.score(x));
        mx = k > mx ? k : mx;
      }
#if FLT_DEVICE_BUILD
    mx = ctaMax64(cta, mx, w.red());
#endif
    thrPre = keyToDouble(mx) - c.beamThreshold;
    int* lnk = w.lnk();
    int* lhead = w.lhead();
    for (int x = cta.tid; x < nCand; x += cta.nthr) {
      if (!(cd.parflag(x) & CF_ALIVE)) continue;
      const int r = mh[cslot[x]];
      if (r == x || !(cd.score(x) >= thrPre)) continue;
#if FLT_DEVICE_BUILD
      lnk[x] = atomicExch(&lhead[r], x);
#else
      lnk[x] = lhead[r];
      lhead[r] = x;
#endif
    }
    cta.sync();
  }
  u64* rkey = w.rkey();
  int* rep = w.rep();
  u64 orK = 0, andK = ~0ull;
  for (int x = cta.tid; x < nCand; x += cta.nthr) {
    if (!(cd.parflag(x) & CF_ALIVE)) continue;
    if (mh[cslot[x]] != x) continue;
    if (logAdd) {
      if (!(cd.score(x) >= thrPre)) { // the group's best is below the threshold: so are all
        mh[cslot[x]] = -1;            // (the select clears the slots of the groups it is given)
        continue;
      }
      const int* lnk = w.lnk();
      const int head = w.lhead()[x];
      if (head >= 0) {
        double acc = cd.score(x);
        double lastS = 0.0;
        int lastI = -1; // members are visited by (score desc, index asc); -1 = none visited yet
        for (;;) {
          int best = -1;
          double bs = 0.0;
          for (int m = head; m >= 0; m = lnk[m]) {
            const double sm = cd.score(m);
            if (lastI >= 0 && !(sm < lastS || (sm == lastS && m > lastI))) continue;
            if (best < 0 || sm > bs || (sm == bs && m < best)) {
              best = m;
              bs = sm;
            }
          }
          if (best < 0) break;
          const double hi = acc > bs ? acc : bs, lo = acc > bs ? bs : acc;
          acc = hi + flt_log1p(flt_exp(lo - hi));
          lastS = bs;
          lastI = best;
        }
        cd.score(x) = acc;
      }
    }
    const int r = aggInc(&sc[SC_NREP], cta.tid);
    const u64 k = orderedKey64(cd.score(x));
    rep[r] = x;
    rkey[r] = k; // keyA storage is free again: every insert finished at the barrier above
    orK |= k;
    andK &= k;
  }
#if FLT_DEVICE_BUILD
  {
    const unsigned oLo = __reduce_or_sync(0xffffffffu, (unsigned)orK);
    const unsigned oHi = __reduce_or_sync(0xffffffffu, (unsigned)(orK >> 32));
    const unsigned aLo = __reduce_and_sync(0xffffffffu, (unsigned)andK);
    const unsigned aHi = __reduce_and_sync(0xffffffffu, (unsigned)(andK >> 32));
    if ((cta.tid & 31) == 0) {
      atomicOr((unsigned*)&sc[SC_OR_LO], oLo);
      atomicOr((unsigned*)&sc[SC_OR_HI], oHi);
      atomicAnd((unsigned*)&sc[SC_AND_LO], aLo);
      atomicAnd((unsigned*)&sc[SC_AND_HI], aHi);
    }
  }
#else
  sc[SC_OR_LO] |= (int)(unsigned)orK;
  sc[SC_OR_HI] |= (int)(unsigned)(orK >> 32);
  sc[SC_AND_LO] &= (int)(unsigned)andK;
  sc[SC_AND_HI] &= (int)(unsigned)(andK >> 32);
#endif
  cta.sync();
}

// find, scanning bins from 255 down, the bin where the running count reaches `need`; the bins are
// zeroed again as they are read (one warp)
FLT_DEV void findCutBin(const Cta& cta, const Ws& w, int need) {
  int* hist = w.hist();
  int* sc = w.sc();
#if FLT_DEVICE_BUILD
  if (cta.tid < 32) {
    const int lane = cta.tid;
    int h[8];
    int part = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      h[k] = hist[255 - (lane * 8 + k)];
      hist[255 - (lane * 8 + k)] = 0;
      part += h[k];
    }
    int incl = part;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    const int excl = incl - part;
    if (excl < need && incl >= need) {
      int cum = excl;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (cum < need && cum + h[k] >= need) {
          sc[SC_BIN] = 255 - (lane * 8 + k);
          sc[SC_NEED] = need - cum;
          sc[SC_BINCOUNT] = h[k];
        }
        cum += h[k];
      }
    }
  }
#else
  if (cta.tid == 0) {
    int cum = 0;
    bool found = false;
    for (int b = 255; b >= 0; --b) {
      const int h = hist[b];
      hist[b] = 0;
      if (!found && cum + h >= need) {
        sc[SC_BIN] = b;
        sc[SC_NEED] = need - cum;
        sc[SC_BINCOUNT] = h;
        found = true;
      }
      cum += h;
    }
  }
#endif
}

// Phase Sel: choose the min(nRep, K) best representatives (Utils.h:200-220), rank them (score
// descending, deterministic ties) into surv[capP..], and return how many there are.
// Radix select on the score keys, most significant *varying* bit first (from the AND/OR of the
// keys); a cut bin with <= 32 members is finished by one warp.
// `binned`: the groups come out of the two-pass pruning, i.e. their scores lie in the bins [SC_PCUT, 255) of
// the frame's monotone map (pruneBin) — the select then ranks by histogram instead of by counting.
FLT_DEV int phaseSelect(const Cta& cta, const DecCfg& c, const Ws& w, int nRep, bool binned = false) {
  const Cand cd = w.cand();
  const int K = c.K;
  int* sc = w.sc();
  int* surv = w.surv();
  int* ranked = w.surv() + c.capP;
  const int* rep = w.rep();
  const u64* rkey = w.rkey();
  u64* skey = w.skey();
  int* pos = w.pos();
  int* hist = w.hist();
  int* mh = w.mh();
  const int* cslot = w.cslot();
  int nSel;
  if (binned && nRep > K && nRep < 65536 && !(c.dbg & 4)) {
    // Histogram rank (as in beam_lf.h): the kept score range [lower edge of the cut bin, top] is spread over
    // 256 fine bins; a suffix scan gives every group the number of groups in higher bins, groups with fewer
    // than K above them are listed grouped by bin (position = groups above + arrival order in the bin) and
    // settle their exact rank against the members of their own bin only. O(groups x bin occupancy) instead
    // of O(groups^2); crowded bins only lengthen the comparisons, the order is the deterministic one.
    const double lo = bitsF64(((u64)(unsigned)sc[SC_PLO_HI] << 32) | (unsigned)sc[SC_PLO_LO]);
    const float pscale = bitsF32((uint32_t)sc[SC_PSCALE]);
    const float cutF = (float)sc[SC_PCUT];
    const float fine = 256.0f / (256.0f - cutF);
    int* bs = (int*)(w.rkey() + c.capC); // the merge keys' second half is dead by now: [capC] bin | slot, [capC] list
    int* list = bs + c.capC;
    for (int r = cta.tid; r < nRep; r += cta.nthr) {
      const int x = rep[r];
      mh[cslot[x]] = -1; // leave the merge table empty
      const float p = ((float)(cd.score(x) - lo) * pscale - cutF) * fine; // monotone in the score
      const int bin = p >= 255.0f ? 255 : (p > 0.0f ? (int)p : 0);
      bs[r] = (bin << 16) | atomAdd(&hist[bin], 1);
    }
    cta.sync();
#if FLT_DEVICE_BUILD
    if (cta.tid < 32) { // hist[b] <- groups in bins above b
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
      int cum = incl - part;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        hist[255 - (lane * 8 + k)] = cum;
        cum += h[k];
      }
    }
#else
    if (cta.tid == 0) {
      int cum = 0;
      for (int b = 255; b >= 0; --b) {
        const int h = hist[b];
        hist[b] = cum;
        cum += h;
      }
    }
#endif
    cta.sync();
    for (int r = cta.tid; r < nRep; r += cta.nthr) {
      const int ab = hist[bs[r] >> 16];
      if (ab < K) list[ab + (bs[r] & 0xFFFF)] = r;
    }
    cta.sync();
    for (int r = cta.tid; r < nRep; r += cta.nthr) {
      const int bin = bs[r] >> 16;
      const int ab = hist[bin];
      if (ab >= K) continue; // K groups score strictly higher
      const int end = bin == 0 ? nRep : hist[bin - 1]; // members of the bin sit at list[ab .. end)
      const u64 ka = rkey[r];
      const int xa = rep[r];
      int cnt = ab;
      for (int q = ab; q < end; ++q) {
        const int rb = list[q];
        const u64 kb = rkey[rb];
        cnt += (kb > ka || (kb == ka && rb != r && candBetter(cd, rep[rb], xa))) ? 1 : 0;
      }
      if (cnt < K) ranked[cnt] = xa;
    }
    cta.sync();
    for (int b = cta.tid; b < 256; b += cta.nthr) hist[b] = 0; // the next frame's first pass counts from zero
    cta.sync();
    return K;
  }
  if (nRep <= K) {
    for (int r = cta.tid; r < nRep; r += cta.nthr) {
      const int x = rep[r];
      surv[r] = x;
      skey[r] = rkey[r];
      pos[r] = 0;
      mh[cslot[x]] = -1; // leave the merge table empty
    }
    nSel = nRep;
    cta.sync();
  } else if (!(c.dbg & 1) && (long long)nRep * nRep <= 256LL * cta.nthr) { // <= 362 groups at 512 threads
    // A few hundred groups (the two-pass pruning keeps ~1.5 K candidates): rank every group by
    // counting the groups whose score key is larger — `parts` adjacent lanes share one group and add
    // their slices up with shuffles — and keep ranks < K. One barrier instead of the radix passes
    // below. Equal keys are rare; a group that has any is ranked again with the full deterministic
    // comparator, so the total order is the same as on the radix path.
    int lg = 0;
    while (lg < 5 && (nRep << (lg + 1)) <= cta.nthr) ++lg;
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
          eq += kb == ka ? 1 : 0;
        }
      }
#if FLT_DEVICE_BUILD
      for (int o = 1; o < parts; o <<= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        eq += __shfl_xor_sync(0xffffffffu, eq, o);
      }
#endif
      if (valid && part == 0) {
        const int xa = rep[a];
        if (eq > 1) { // ties (eq counts the group itself): exact order among the equal keys
          for (int b = 0; b < nRep; ++b)
            if (b != a && rkey[b] == ka && candBetter(cd, rep[b], xa)) ++cnt;
        }
        mh[cslot[xa]] = -1; // leave the merge table empty
        if (cnt < K) ranked[cnt] = xa;
      }
    }
    cta.sync();
    return K;
  } else {
    const u64 orK = ((u64)(unsigned)sc[SC_OR_HI] << 32) | (unsigned)sc[SC_OR_LO];
    const u64 andK = ((u64)(unsigned)sc[SC_AND_HI] << 32) | (unsigned)sc[SC_AND_LO];
    const u64 vary = orK ^ andK; // bit positions on which the keys differ
    const int top = vary ? highBit64(vary) : 0;
    int shift = top >= 7 ? top - 7 : 0; // first digit = the 8 bits ending at the top varying bit
    int width = top >= 7 ? 8 : top + 1;
    int need = K;
    u64 hiMask = 0, hiVal = 0; // digits already fixed (bits above the current digit)
    int mode = 0;              // 1 = whole cut bin selected, 2 = gathered bin resolved, 3 = ties
    int cutDigit = 0;
    for (;;) {
      const u64 dmask = (1ull << width) - 1ull;
      for (int r = cta.tid; r < nRep; r += cta.nthr) {
        const u64 k = rkey[r];
        if ((k & hiMask) == hiVal) atomAdd(&hist[(int)((k >> shift) & dmask)], 1);
      }
      cta.sync();
      findCutBin(cta, w, need);
      cta.sync();
      cutDigit = sc[SC_BIN];
      need = sc[SC_NEED];
      const int binCount = sc[SC_BINCOUNT];
      if (binCount == need) {
        mode = 1;
        break;
      }
      if (binCount <= 32) {
        mode = 2;
        break;
      }
      if (shift == 0) {
        mode = 3; // more than 32 identical keys straddle the cut
        break;
      }
      hiMask |= dmask << shift;
      hiVal |= (u64)cutDigit << shift;
      width = shift >= 8 ? 8 : shift;
      shift -= width;
    }
    const u64 dmask = (1ull << width) - 1ull;
    // selected outright: same fixed digits and a larger current digit, or (mode 1) the cut bin too
    for (int r = cta.tid; r < nRep; r += cta.nthr) {
      const u64 k = rkey[r];
      const int x = rep[r];
      mh[cslot[x]] = -1; // leave the merge table empty
      bool sel = false;
      if ((k & hiMask) != hiVal) {
        sel = (k & hiMask) > hiVal;
      } else {
        const int dg = (int)((k >> shift) & dmask);
        sel = dg > cutDigit || (mode == 1 && dg == cutDigit);
        if (mode >= 2 && dg == cutDigit) w.gath()[aggInc(&sc[SC_NGATH], cta.tid) & 63] = r;
      }
      if (sel) {
        const int a = aggInc(&sc[SC_NSEL], cta.tid);
        surv[a] = x;
        skey[a] = k;
        pos[a] = 0;
      }
    }
    cta.sync();
    if (mode == 2) {
      // one warp ranks the <= 32 members of the cut bin and keeps the best `need`
#if FLT_DEVICE_BUILD
      if (cta.tid < 32) {
        const int ng = sc[SC_NGATH];
        const int lane = cta.tid;
        const int r = lane < ng ? w.gath()[lane] : -1;
        const u64 k = r >= 0 ? rkey[r] : 0ull;
        const int x = r >= 0 ? rep[r] : -1;
        int better = 0;
        for (int o = 0; o < ng; ++o) {
          const u64 ko = __shfl_sync(0xffffffffu, k, o);
          const int xo = __shfl_sync(0xffffffffu, x, o);
          if (r >= 0 && o != lane) better += (ko > k || (ko == k && candBetter(cd, xo, x))) ? 1 : 0;
        }
        if (r >= 0 && better < need) {
          const int a = aggInc(&sc[SC_NSEL], cta.tid);
          surv[a] = x;
          skey[a] = k;
          pos[a] = 0;
        }
      }
#else
      if (cta.tid == 0) {
        const int ng = sc[SC_NGATH];
        for (int i = 0; i < ng; ++i) {
          const int r = w.gath()[i];
          int better = 0;
          for (int o = 0; o < ng; ++o) {
            const int ro = w.gath()[o];
            if (o != i) better += (rkey[ro] > rkey[r] || (rkey[ro] == rkey[r] && candBetter(cd, rep[ro], rep[r]))) ? 1 : 0;
          }
          if (better < need) {
            const int a = sc[SC_NSEL]++;
            surv[a] = rep[r];
            skey[a] = rkey[r];
            pos[a] = 0;
          }
        }
      }
#endif
      cta.sync();
    } else if (mode == 3) {
      // rare: `need` of many equal-score groups, by the deterministic order (one thread)
      if (cta.tid == 0) {
        const int n0 = sc[SC_NSEL];
        int n = n0;
        for (int q = 0; q < need; ++q) {
          int bestX = -1;
          u64 bk = 0;
          for (int r = 0; r < nRep; ++r) {
            const u64 k = rkey[r];
            if ((k & hiMask) != hiVal || (int)((k >> shift) & dmask) != cutDigit) continue;
            const int x = rep[r];
            bool taken = false;
            for (int z = n0; z < n; ++z) taken |= surv[z] == x;
            if (taken) continue;
            if (bestX < 0 || candBetter(cd, x, bestX)) {
              bestX = x;
              bk = k;
            }
          }
          surv[n] = bestX;
          skey[n] = bk;
          pos[n] = 0;
          ++n;
        }
        sc[SC_NSEL] = n;
      }
      cta.sync();
    }
    nSel = sc[SC_NSEL];
  }
  // rank by counting, all threads: thread (a, part) counts the survivors of its slice that beat a
  if (nSel > 0) {
    const int parts = nSel >= cta.nthr ? 1 : cta.nthr / nSel;
    const int slice = (nSel + parts - 1) / parts;
    for (int t = cta.tid; t < nSel * parts; t += cta.nthr) {
      const int a = t % nSel, part = t / nSel;
      const int lo = part * slice, hi = lo + slice < nSel ? lo + slice : nSel;
      const u64 ka = skey[a];
      const int xa = surv[a];
      int cnt = 0;
      for (int b = lo; b < hi; ++b) {
        const u64 kb = skey[b];
        cnt += (kb > ka || (kb == ka && b != a && candBetter(cd, surv[b], xa))) ? 1 : 0;
      }
      if (cnt) atomAdd(&pos[a], cnt);
    }
  }
  cta.sync();
  for (int a = cta.tid; a < nSel; a += cta.nthr) ranked[pos[a]] = surv[a];
  cta.sync();
  return nSel;
}

// Phase F: apply the beam threshold to the ranked survivors (Utils.h:161-165: keep score >=
// best - beamThreshold; for a max-merge filtering after the merge keeps the same groups),
// materialise the new beam, extend the LM-state fingerprints and n-gram contexts, write the
// back-pointer records, and reset the per-frame scalars.
FLT_DEV void phaseFinalize(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur,
                           const Beam& nxt, const FrameIn& f, int nSel) {
  const Cand cd = w.cand();
  int* sc = w.sc();
  const int* ranked = w.surv() + c.capP;
  if (nSel == 0) {
    if (cta.tid == 0) sc[SC_NH] = 0;
  } else {
    // candidatesBestScore_ - beamThreshold (LexiconDecoder.cpp:217-224)
    // (logAdd: the threshold was applied to the candidates before they were merged, phaseMerge)
    const double thrScore = c.logAdd ? negInf() : cd.score(ranked[0]) - c.beamThreshold;
    for (int q = cta.tid; q < nSel; q += cta.nthr) {
      const int x = ranked[q];
      const double score = cd.score(x);
      if (!(score >= thrScore)) continue;
      if (q + 1 == nSel || !(cd.score(ranked[q + 1]) >= thrScore)) sc[SC_NH] = q + 1;
      const int p = cd.par(x);
      const int fl = cd.flags(x);
      const int n = cd.tok(x);
      nxt.score(q) = score;
      if (fl & CF_FINISH) {
        nxt.am(q) = cur.am(p);
      } else {
        nxt.am(q) = cur.am(p) + amOf(c, f, cd.ce(x), n, cur.tok(p));
      }
      nxt.lm(q) = cur.lm(p) + (double)cd.lmd(x);
      nxt.lex(q) = cd.lex(x);
      nxt.tok(q) = n;
      nxt.pb(q) = (fl & CF_PB) ? 1 : 0;
      if (fl & CF_NEW) {
        const int lab = candLabel(c, cd, x);
        fpChild(cur.fpA(p), cur.fpB(p), lab, nxt.fpA(q), nxt.fpB(q));
        if (c.lm.kind) {
          const int wlm = lab < 0 ? c.lm.eos : c.lm.usr2lm[lab];
          nxt.nctx(q) = ngramAdvanceCtx(c.lm, cur.ctx(p), cur.nctx(p), wlm, nxt.ctx(q));
          nxt.boMask(q) = ngramContextBackoffs(c.lm, nxt.ctx(q), nxt.nctx(q), nxt.bo(q)); // once per new LM state
        }
      } else {
        nxt.fpA(q) = cur.fpA(p);
        nxt.fpB(q) = cur.fpB(p);
        if (c.lm.kind) {
          const int nc = cur.nctx(p);
          nxt.nctx(q) = nc;
          for (int k = 0; k < nc; ++k) {
            nxt.ctx(q)[k] = cur.ctx(p)[k];
            nxt.bo(q)[k] = cur.bo(p)[k];
          }
          nxt.boMask(q) = cur.boMask(p);
        }
      }
      f.hParent[q] = p;
      f.hTok[q] = n;
      if (f.hWord) f.hWord[q] = cd.word(x);
      const int an = skipCarry(cur, f.hRow, p);
      nxt.anc(q) = an;
      if (f.hSkip) f.hSkip[q] = an;
      if (f.hScore) {
        f.hScore[3 * q] = score;
        f.hScore[3 * q + 1] = nxt.am(q);
        f.hScore[3 * q + 2] = nxt.lm(q);
      }
    }
  }
  if (cta.tid == 0) { // scalars for the next frame's merge / select
    sc[SC_NREP] = 0;
    sc[SC_NSEL] = 0;
    sc[SC_NGATH] = 0;
    sc[SC_OR_LO] = 0;
    sc[SC_OR_HI] = 0;
    sc[SC_AND_LO] = -1;
    sc[SC_AND_HI] = -1;
  }
  cta.sync();
  if (f.hCount && cta.tid == 0) *f.hCount = sc[SC_NH];
}

// One frame: cur -> nxt. All threads of the CTA call this with identical arguments.
template <bool W>
FLT_DEV void frameStep(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur,
                       const Beam& nxt, const FrameIn& f, int* status, unsigned long long* stats) {
  const bool full = W && c.full, rootList = W && c.rootList;
  int* sc = w.sc();
  const int nH = sc[SC_NH];
  if (nH == 0) return; // the beam died (Utils.h:155-158): every later frame is empty

  // issue the scattered emission reads now; they are consumed after the row grouping
  float eOwn = 0.0f, eBlank = 0.0f, eSil = 0.0f;
  if (cta.tid < nH) {
    const int n = ownToken(c, cur, cta.tid);
    if (n >= 0 && n < c.N) eOwn = f.e[n];
  }
  if (cta.tid == cta.nthr - 1) {
    if (c.ctc) eBlank = f.e[c.blank];
    eSil = f.e[c.sil];
  }
  // ... and the Trie offsets of this thread's hypothesis (two L2 loads): degree and first edge, consumed by the
  // degree scan after the row grouping; the first edge is kept per hypothesis so that an edge item does not
  // chase childOff[lex] again (one L2 level less in every edge item of both passes)
  int myDeg = 0, myEoff = 0;
  if (c.lexicon && cta.tid < nH) {
    const int lex = cur.lex(cta.tid);
    if (!(lex == 0 && (c.wideRanked || rootList))) {
      myEoff = c.trie.childOff[lex];
      myDeg = c.trie.childOff[lex + 1] - myEoff;
    }
  }
  // lexicon, ranked rows: what the ~K ln K wide cells need to know about the root child of each list
  // token (node, has children, smeared score) is gathered ONCE per list entry, by the threads at the
  // far end of the CTA, instead of once per cell
  const bool listCache = c.lexicon && c.wideRanked;
  const int lj = cta.nthr - 2 - cta.tid; // list entry cached by this thread
  int ljNode = -1;
  float ljMax = 0.0f;
  auto listInfo = [&](int j, int& node, float& ms) __attribute__((always_inline)) {
    node = -1;
    ms = 0.0f;
    const int n = f.topTok[j];
    if (n < 0) return;
    const int child = c.trie.rootChild[n];
    if (child < 0 || c.trie.childOff[child + 1] == c.trie.childOff[child]) return;
    node = child;
    ms = c.trie.maxScore[child];
  };
  if (listCache && lj >= 0 && lj < f.listLen) listInfo(lj, ljNode, ljMax);
  float* spec = w.spec();
  auto publish = [&]() {
    if (listCache) {
      if (lj >= 0 && lj < f.listLen) {
        w.listNode()[lj] = ljNode;
        w.listMax()[lj] = ljMax;
      }
      for (int j = cta.nthr - 1 + cta.tid; j < f.listLen; j += cta.nthr) { // lists longer than the CTA
        int node;
        float ms;
        listInfo(j, node, ms);
        w.listNode()[j] = node;
        w.listMax()[j] = ms;
      }
    }
    if (cta.tid < nH) spec[cta.tid] = eOwn;
    for (int i = cta.tid + cta.nthr; i < nH; i += cta.nthr) { // beams wider than the CTA
      const int n = ownToken(c, cur, i);
      spec[i] = (n >= 0 && n < c.N) ? f.e[n] : 0.0f;
    }
    if (cta.tid == cta.nthr - 1) {
      spec[c.K] = eBlank;
      spec[c.K + 1] = eSil;
    }
    if (cta.tid == 0) {
      sc[SC_OVF] = 0;
      sc[SC_NCAND] = 0;
    }
  };
  // per-phase SM cycles of thread 0 (only when the host asked for counters: stats != null), summed
  // over frames into stats[4 + phase]; stats[31] marks the generic step's phase names for the host
#if FLT_DEVICE_BUILD
  auto mark = [&](int k) __attribute__((always_inline)) { // k < 0: start stamp only
    if (stats && cta.tid == 0) {
      const long long t = clock64();
      const long long t0 = (long long)(((u64)(unsigned)sc[SC_THI] << 32) | (unsigned)sc[SC_TLO]);
      if (k >= 0) atomicAdd(stats + 4 + k, (unsigned long long)(t - t0));
      sc[SC_TLO] = (int)(unsigned)t;
      sc[SC_THI] = (int)(unsigned)((u64)t >> 32);
    }
  };
  mark(-1);
#else
  auto mark = [&](int) {};
#endif
  int wideItems = 0;
  if (c.wideRanked) {
    phaseRows(cta, c, w, cur, nH, publish);
    wideItems = w.wideOff()[sc[SC_NROWS]];
  } else {
    publish();
    cta.sync();
  }
  mark(0); // rows
  // Pruning bound (exact): any rectangle rows 1..a x columns 0..col of the (row rank x ranked
  // token) grid holds >= a*(col-2) >= K regular cells, i.e. K distinct merge groups, each scoring
  // at least fl(s + e_col) where s bounds the a-th row's best member from below. A candidate below
  // the best such corner can never be among the K best groups. Lexicon-free rows have at most two
  // members, so the a-th row's leader is one of the first 2a-1 hypotheses.
  double tau = negInf();
  if (c.wideRanked && !c.lexicon) {
    for (int k = 0; k < c.nTau; ++k) {
      const int i = 2 * c.tauA[k] - 2, col = c.tauCol[k];
      if (i < nH && col < f.listLen && f.topTok[col] >= 0) {
        const double corner = cur.score(i) + (double)f.topVal[col] + c.lmWeight * (double)0.0f;
        tau = corner > tau ? corner : tau;
      }
    }
    if (c.silScore < 0) tau += c.silScore; // keeps the bound valid if a counted cell is the sil one
  }
  if (full && c.ctc) {
    // the best hypothesis' blank candidate exists whenever blank is in the token set, so the frame's
    // best candidate scores at least as much: anything below that minus beamThreshold is dropped by
    // the reference's own filter (Utils.h:131-144,161-165)
    const float eB = spec[c.K];
    if (inTokenSetV(c, f, c.blank, eB)) {
      double sb = cur.score(0) + (double)eB;
      if (c.blank == c.sil) sb += c.silScore;
      tau = sb - c.beamThreshold;
    }
  }
  // trie edge items: prefix sums of the hypotheses' degrees
  int edgeItems = 0;
  if (c.lexicon) {
    const TrieDev& t = c.trie;
    int* deg = w.rows().deg();
    int* eoff = w.rows().eoff();
    for (int i = cta.tid; i < nH; i += cta.nthr) {
      const int lex = cur.lex(i);
      if (lex == 0 && c.wideRanked) deg[i] = t.nRootLab;
      else if (lex == 0 && rootList) deg[i] = f.listLen;
      else if (i == cta.tid) { // loaded at the top of the frame
        deg[i] = myDeg;
        eoff[i] = myEoff;
      } else { // beams wider than the CTA
        eoff[i] = t.childOff[lex];
        deg[i] = t.childOff[lex + 1] - eoff[i];
      }
    }
    cta.sync();
    ctaExclusiveScan(cta, deg, w.rows().degTmp(), nH);
    edgeItems = deg[nH];
  }
  // pass 1 records, per work item, the best bin any of its candidates reached; pass 2 skips the
  // items that cannot reach the cut without recomputing them (their gathers are the expensive part)
  int localBin = -1;
  Ws wp = w;
  unsigned char* pcache = c.prune2 ? (unsigned char*)(w.base + c.lay.pruneCache) : nullptr;
  // pass: 0 = no pruning, 1 = histogram pass, 2 = materialise (items that cannot reach the cut are skipped from
  // the first pass' cache), 3 = materialise against a guessed cut (no first pass, every item is evaluated)
  auto emitAll = [&](int pass) __attribute__((always_inline)) {
    const int cut = pass == 2 ? sc[SC_PCUT] : 0;
    wp.itemBin = pass == 1 ? &localBin : nullptr;
    auto skip = [&](int slot) { return pass == 2 && pcache && (int)pcache[slot] <= cut; };
    auto note = [&](int slot) {
      if (pass == 1) pcache[slot] = (unsigned char)(localBin + 1);
      localBin = -1;
    };
    // the three kinds of work items (cache slots: [0, wideTotal) wide cells, then K specials, then the
    // first kPruneEdgeCap trie edges)
    auto doWide = [&](int x) __attribute__((always_inline)) {
      const int r = w.itemRow()[x]; // row rank (0-based)
      emitWide(cta, c, wp, cur, f, w.rows().leaderOfRank(r), x - w.wideOff()[r], tau);
    };
    auto doSpecial = [&](int i) __attribute__((always_inline)) {
      emitSpecials(cta, c, wp, cur, f, i, tau);
      if (c.wideRanked) emitSilCell(cta, c, wp, cur, f, i, tau);
    };
    auto doEdge = [&](int x) __attribute__((always_inline)) {
      const TrieDev& t = c.trie;
      const int* deg = w.rows().deg();
      const int i = searchOffsets(deg, nH + 1, x);
      const int k = x - deg[i];
      const int lex = cur.lex(i);
      if (c.wideRanked && lex == 0) {
        const int n = t.rootLabTok[k];
        emitEdge<W>(cta, c, wp, cur, f, i, n, t.rootChild[n], true, tau);
      } else if (rootList && lex == 0) {
        const int n = f.topTok[k];
        const int child = n >= 0 ? t.rootChild[n] : -1;
        if (child >= 0) emitEdge<W>(cta, c, wp, cur, f, i, n, child, false, tau);
      } else {
        const int e = w.rows().eoff()[i] + k;
        emitEdge<W>(cta, c, wp, cur, f, i, t.childTok[e], t.childNode[e], false, tau);
      }
    };
    if (pass == 2 && pcache && !full && !(c.dbg & 32)) {
      // Second pass, compacted: only about one item in ten reaches the cut, and a thread that owned two or
      // three of them ran their gather chains one after the other while most threads had none. The items
      // whose cached bin reaches the cut are listed first (one byte read per item) and then dealt out one per
      // thread. (The list reuses the representatives' array, idle until the merge; an item on it yields at
      // least one kept candidate, so the list is no longer than the candidate set.)
      int* act = w.rep();
      const int cachedEdges = edgeItems < kPruneEdgeCap ? edgeItems : kPruneEdgeCap;
      const int slots = c.wideTotal + c.K + (c.lexicon ? cachedEdges : 0);
      for (int sl = cta.tid; sl < slots; sl += cta.nthr) {
        const bool valid = sl < c.wideTotal ? (c.wideRanked && sl < wideItems)
                                            : (sl < c.wideTotal + c.K ? sl - c.wideTotal < nH : true);
        if (valid && (int)pcache[sl] > cut) {
          const int pos = aggInc(&sc[SC_NACT], cta.tid);
          if (pos < c.capC) act[pos] = sl;
          else sc[SC_OVF] = 1;
        }
      }
      cta.sync();
      const int nAct = sc[SC_NACT] < c.capC ? sc[SC_NACT] : c.capC;
      for (int a = cta.tid; a < nAct; a += cta.nthr) {
        const int sl = act[a];
        if (sl < c.wideTotal) doWide(sl);
        else if (sl < c.wideTotal + c.K) doSpecial(sl - c.wideTotal);
        else doEdge(sl - c.wideTotal - c.K);
      }
      for (int x = kPruneEdgeCap + cta.tid; x < edgeItems; x += cta.nthr) doEdge(x); // beyond the cache: all
      cta.sync();
      if (cta.tid == 0) sc[SC_NACT] = 0;
      return;
    }
    // wide cells
    if (c.wideRanked) {
      for (int x = cta.tid; x < wideItems; x += cta.nthr) {
        if (skip(x)) continue;
        doWide(x);
        note(x);
      }
    }
    // lexicon-free full expansion: every hypothesis x every token of the set (blank and repeat are
    // two of its cells)
    if (full) {
      const int S = c.setAll ? c.N : f.listLen;
      const long long items = (long long)nH * S;
      for (long long x = cta.tid; x < items; x += cta.nthr) {
        const int i = (int)(x / S), j = (int)(x - (long long)i * S);
        int n = j;
        float ev;
        if (c.setAll) {
          ev = f.e[j];
        } else {
          n = f.topTok[j];
          if (n < 0) continue;
          ev = f.topVal[j];
        }
        emitFullLf(cta, c, wp, cur, i, n, ev, tau);
      }
    }
    // stay / repeat / blank
    for (int i = cta.tid; i < (full ? 0 : nH); i += cta.nthr) {
      if (skip(c.wideTotal + i)) continue;
      doSpecial(i);
      note(c.wideTotal + i);
    }
    // trie edges
    if (c.lexicon) {
      for (int x = cta.tid; x < edgeItems; x += cta.nthr) {
        const bool cached = x < kPruneEdgeCap;
        if (cached && skip(c.wideTotal + c.K + x)) continue;
        doEdge(x);
        if (cached) note(c.wideTotal + c.K + x);
        else localBin = -1;
      }
    }
    cta.sync();
  };
  // Two-pass pruning for the lexicon decoder (no corner bound there): histogram the scores of
  // everything the frame would propose, keep the bins that hold the best ~3K candidates.
  const bool prune = c.prune2 != 0;
  int nCand = 0;
  bool done = false;
  if (prune) {
    if (cta.tid == 0) {
      // candidate scores lie below best hypothesis + bonuses for log-probability emissions; the
      // span follows the beam's own spread. Neither affects exactness (bins clamp, result verified).
      double hi = cur.score(0);
      if (c.silScore > 0) hi += c.silScore;
      if (c.wordScore > 0) hi += c.wordScore;
      double span = 2.0 * (cur.score(0) - cur.score(nH - 1)) + 12.0;
      span = span > 96.0 ? 96.0 : span;
      if (c.beamThreshold + 4.0 < span) span = c.beamThreshold + 4.0;
      const double lo = hi - span;
      const u64 lb = f64Bits(lo);
      const float scale = (float)kPruneBins / (float)span;
      sc[SC_PLO_LO] = (int)(unsigned)lb;
      sc[SC_PLO_HI] = (int)(unsigned)(lb >> 32);
      sc[SC_PSCALE] = (int)f32Bits(scale);
      sc[SC_BIN] = 0; // fewer candidates than wanted: keep every bin
      // keep ~1.5 K candidates; after a frame whose kept bins held fewer than K merge groups (it was
      // redone without the cut) fall back to 3K+64 for a while
      const int hold = sc[SC_WHOLD];
      sc[SC_WANT] = hold > 0 ? 3 * c.K + 64 : c.pruneWant;
      if (hold > 0) sc[SC_WHOLD] = hold - 1;
      // Guessed cut (experiment, FLT_DBG=8; OFF by default): in steady state the K-th best candidate sits about
      // as far below the top as it did in the previous frame, so the frame is first tried in ONE pass against
      // the bin of (top - that span): if the kept candidates fit the capacity and form >= K merge groups the
      // result is exact (everything at or above the cut was kept, as with a cut chosen from a histogram) and
      // the histogram pass is saved; otherwise the frame is done again with the two exact passes and guessing
      // pauses for a few frames. Measured on cfg 3 (B200): the kept count moves by a whole bin's worth of
      // candidates (~30) from frame to frame, 44 % of the guesses missed, 7.90 k utt/s against 8.27 k without.
      const float gspan = bitsF32((uint32_t)sc[SC_GSPAN]);
      const int ghold = sc[SC_GHOLD];
      int gbin = 0;
      if ((c.dbg & 8) && gspan > 0.0f && ghold == 0 && hold == 0 && nH == c.K) {
        const float pos = ((float)span - gspan) * scale;
        gbin = pos >= (float)(kPruneBins - 3) ? kPruneBins - 3 : (pos > 1.0f ? (int)pos : 0);
      }
      if (ghold > 0) sc[SC_GHOLD] = ghold - 1;
      sc[SC_GBIN] = gbin;
      sc[SC_PMODE] = gbin > 0 ? 2 : 1;
      sc[SC_PCUT] = gbin;
    }
    cta.sync();
    mark(1); // degrees + scan
    const int gbin = sc[SC_GBIN];
    if (gbin > 0) {
      emitAll(3);
      mark(4); // the single pass
      nCand = sc[SC_NCAND];
      const bool ovf = sc[SC_OVF] != 0;
      if (!ovf) {
        phaseMerge<W>(cta, c, w, nCand);
        done = sc[SC_NREP] >= c.K;
      }
      cta.sync();
      if (!done) { // undo: empty the merge table, reset the counters, exact passes below
        if (!ovf)
          for (int x = cta.tid; x < nCand; x += cta.nthr)
            if (w.cand().parflag(x) & CF_ALIVE) w.mh()[w.cslot()[x]] = -1;
        if (cta.tid == 0) {
          sc[SC_NREP] = 0;
          sc[SC_OR_LO] = 0;
          sc[SC_OR_HI] = 0;
          sc[SC_AND_LO] = -1;
          sc[SC_AND_HI] = -1;
          sc[SC_NCAND] = 0;
          sc[SC_OVF] = 0;
          sc[SC_PCUT] = 0;
          sc[SC_PMODE] = 1;
          sc[SC_GHOLD] = 8;
#if FLT_DEVICE_BUILD
          if (stats) atomicAdd(stats + 13, 1ull);
#endif
        }
      } else if (cta.tid == 0) {
        // steer the kept count towards 1.25 K .. 1.5 K + 32 candidates (merge groups ~ candidates)
        float g = bitsF32((uint32_t)sc[SC_GSPAN]);
        if (nCand > c.K + c.K / 2 + 32) g *= 0.95f;
        else if (nCand < c.K + c.K / 4) g *= 1.05f;
        sc[SC_GSPAN] = (int)f32Bits(g);
#if FLT_DEVICE_BUILD
        if (stats) atomicAdd(stats + 12, 1ull);
#endif
      }
      cta.sync();
    }
    if (!done) {
      emitAll(1); // pass 1: histogram only
      mark(2); // pass 1
      // cut = lowest bin with fewer than `want` candidates in higher bins (one warp; bins re-zeroed)
      findCutBin(cta, w, sc[SC_WANT]);
      cta.sync();
      if (cta.tid == 0) {
        sc[SC_PMODE] = 2;
        sc[SC_PCUT] = sc[SC_BIN];
        // next frame's guess: the span this exact cut keeps, widened to about twice the candidates
        const float scale = bitsF32((uint32_t)sc[SC_PSCALE]);
        const float kept = (float)(kPruneBins - sc[SC_BIN]) / scale;
        sc[SC_GSPAN] = sc[SC_BIN] > 0 ? (int)f32Bits(kept * 1.05f) : 0;
      }
      cta.sync();
    }
  }
  if (!done) {
    mark(3); // cut bin
    emitAll(prune ? 2 : 0);
    mark(4); // pass 2 (or the only pass)
    nCand = sc[SC_NCAND];
    if (sc[SC_OVF]) {
      if (cta.tid == 0) *status |= 1;
      nCand = nCand < c.capC ? nCand : c.capC;
    }
    phaseMerge<W>(cta, c, w, nCand);
    if (prune && sc[SC_PCUT] > 0 && sc[SC_NREP] < c.K) {
      // the kept bins hold fewer than K merge groups: take everything (rare)
      cta.sync();
      for (int x = cta.tid; x < nCand; x += cta.nthr)
        if (w.cand().parflag(x) & CF_ALIVE) w.mh()[w.cslot()[x]] = -1;
      if (cta.tid == 0) {
        sc[SC_NREP] = 0;
        sc[SC_OR_LO] = 0;
        sc[SC_OR_HI] = 0;
        sc[SC_AND_LO] = -1;
        sc[SC_AND_HI] = -1;
        sc[SC_NCAND] = 0;
        sc[SC_PCUT] = 0;
        sc[SC_WHOLD] = 64;
        sc[SC_GSPAN] = 0;
      }
      cta.sync();
      emitAll(2);
      nCand = sc[SC_NCAND];
      if (sc[SC_OVF]) {
        if (cta.tid == 0) *status |= 1;
        nCand = nCand < c.capC ? nCand : c.capC;
      }
      phaseMerge<W>(cta, c, w, nCand);
    }
  }
  if (prune && cta.tid == 0) sc[SC_PMODE] = 0; // decodeEnd's candidates are not pruned
  const int nRep = sc[SC_NREP];
  mark(5); // merge
  const int nSel = phaseSelect(cta, c, w, nRep, prune);
  mark(6); // select
#if FLT_DEVICE_BUILD
  if (stats && cta.tid == 0) {
    stats[31] = 1ull;
    atomicAdd(stats + 0, 1ull);
    atomicAdd(stats + 1, (unsigned long long)nCand);
    atomicAdd(stats + 2, (unsigned long long)nRep);
    atomicAdd(stats + 3, (unsigned long long)nSel);
  }
#else
  (void)stats;
#endif
  phaseFinalize(cta, c, w, cur, nxt, f, nSel);
  mark(7); // new beam + history
}

// decodeEnd (LexiconFreeDecoder.cpp:127-158, LexiconDecoder.cpp:231-274) as one more "frame".
template <bool W>
FLT_DEV void finishStep(const Cta& cta, const DecCfg& c, const Ws& w, const Beam& cur,
                        const Beam& nxt, const FrameIn& f) {
  int* sc = w.sc();
  const int nH = sc[SC_NH];
  if (nH == 0) return;
  if (cta.tid == 0) sc[SC_NICE] = 0;
  cta.sync();
  if (c.lexicon) {
    for (int i = cta.tid; i < nH; i += cta.nthr)
      if (cur.lex(i) == 0) sc[SC_NICE] = 1; // "nice ending" exists (benign same-value race)
    cta.sync();
  }
  const bool nice = c.lexicon && sc[SC_NICE] != 0;
  for (int i = cta.tid; i < nH; i += cta.nthr) {
    w.cand().parflag(i) = 0;
    if (nice && cur.lex(i) != 0) continue;
    float ls = 0.0f;
    int flags = CF_FINISH;
    if (c.lm.kind) { // KenLM::finish: score </s>, state = child(-1); ZeroLM: same state, 0
      ls = ngramScoreCached(c.lm, cur.ctx(i), cur.nctx(i), c.lm.eos, cur.bo(i), cur.boMask(i));
      flags |= CF_NEW;
    }
    const double score = cur.score(i) + c.lmWeight * (double)ls;
    putCand(c, w, cur, i, score, i, c.sil, -1, cur.lex(i), flags, ls, 0.0f);
  }
  cta.sync();
  phaseMerge<W>(cta, c, w, nH);
  const int nSel = phaseSelect(cta, c, w, sc[SC_NREP]);
  phaseFinalize(cta, c, w, cur, nxt, f, nSel);
}

/* ------------------------------------------------------------------ whole-utterance driver ---- */
// lexicon-free fast step, beam_lf.h
struct LfCarry { // emissions of the NEXT frame, loaded while the current one retires (registers)
  float eOwn, eBlank, eSil;
  int valid;
};
FLT_DEV void lfFrameStep(const Cta& cta, const DecCfg& c, const Ws& w, int curIdx, const FrameIn& f,
                         unsigned long long* stats, LfCarry& carry);
FLT_DEV void lfFinish(const Cta& cta, const DecCfg& c, const Ws& w, int curIdx, const FrameIn& f);
FLT_DEV void lfTabBuild(const Cta& cta, const Ws& w, int beamIdx, int nH);
FLT_DEV void lfTabClear(const Cta& cta, const Ws& w, int beamIdx, int nH);
FLT_DEV void lfBuildItemDesc(const Cta& cta, const DecCfg& c, const Ws& w);

// One CTA decodes utterances bid, bid+nblk, ... start to finish. `base` is the CTA's workspace:
// shared memory or a global slab.
// workspace tables that persist over the CTA's utterances (kept clean by their users)
FLT_DEV void ctaInitWorkspace(const Cta& cta, const DecCfg& c, const Ws& w, char* base) {
  const int K = c.K;
  for (int i = cta.tid; i <= K; i += cta.nthr) w.wideOff()[i] = c.wideOff[i];
  for (int i = cta.tid; i < (c.lfFast ? 2 : 1) * c.capRH; i += cta.nthr) w.rows().hash[i] = -1;
  if (c.lfFast) {
    int* slotB = (int*)(base + c.lay.lfSlotB);
    for (int i = cta.tid; i < 2 * c.capRH; i += cta.nthr) slotB[i] = -1;
    for (int i = cta.tid; i < c.lfBins; i += cta.nthr) w.hist()[i] = 0;
    lfBuildItemDesc(cta, c, w);
  } else {
    for (int i = cta.tid; i < c.capH; i += cta.nthr) w.mh()[i] = -1;
    for (int i = cta.tid; i < 256; i += cta.nthr) w.hist()[i] = 0;
  }
  if (cta.tid == 0) {
    int* sc = w.sc();
    sc[SC_NREP] = 0;
    sc[SC_NSEL] = 0;
    sc[SC_NGATH] = 0;
    sc[SC_OR_LO] = 0;
    sc[SC_OR_HI] = 0;
    sc[SC_AND_LO] = -1;
    sc[SC_AND_HI] = -1;
    sc[SC_WHOLD] = 0;
    sc[SC_GSPAN] = 0;
    sc[SC_GHOLD] = 0;
    sc[SC_LFGAP] = (int)f32Bits(-1.0f);
    sc[SC_LFFAC] = (int)f32Bits(1.3f);
    sc[SC_LFHOLD] = 0;
    sc[SC_NACT] = 0;
    sc[SC_PMODE] = 0; // allocCand reads these in every mode; only the two-pass pruning sets them
    sc[SC_PCUT] = 0;
    sc[SC_BIN] = 0;
  }
  for (int r = cta.tid; r < K; r += cta.nthr) // row rank of every wide work item
    for (int x = c.wideOff[r]; x < c.wideOff[r + 1]; ++x) w.itemRow()[x] = (short)r;
}

// decodeBegin (LexiconDecoder.cpp:21-30, LexiconFreeDecoder.cpp:20-28): one thread seeds beam 0
FLT_DEV void seedUtterance(const DecCfg& c, const Ws& w, const BatchArgs& a, int b) {
  const int K = c.K;
  const Beam B0 = w.beam(0);
  B0.score(0) = 0.0;
  B0.am(0) = 0.0;
  B0.lm(0) = 0.0;
  fpRoot(B0.fpA(0), B0.fpB(0));
  if (c.lfFast) { // the root state has no parent: a fingerprint no state carries
    B0.pfpA(0) = 0;
    B0.pfpB(0) = 0;
  }
  B0.lex(0) = 0;
  B0.tok(0) = c.sil;
  B0.pb(0) = 0;
  B0.nctx(0) = 0;
  if (c.lm.kind) {
    if (c.lm.order > 1) {
      B0.ctx(0)[0] = c.lm.bos;
      B0.nctx(0) = 1;
    }
    B0.boMask(0) = ngramContextBackoffs(c.lm, B0.ctx(0), B0.nctx(0), B0.bo(0));
  }
  w.sc()[SC_NH] = 1;
  w.sc()[SC_GSPAN] = 0; // the guessed cut of the two-pass pruning starts over with every utterance
  w.sc()[SC_GHOLD] = 0;
  w.sc()[SC_LFGAP] = (int)f32Bits(-1.0f); // and so does the lexicon-free step's guessed bound
  w.sc()[SC_LFFAC] = (int)f32Bits(1.3f);
  w.sc()[SC_LFHOLD] = 0;
  a.status[b] = 0;
  const long long h0 = ((long long)b * (a.T + 2)) * K;
  a.hParent[h0] = -1;
  a.hTok[h0] = c.sil;
  if (a.hWord) a.hWord[h0] = -1;
}

// the three scores of every final hypothesis and their count
FLT_DEV void writeFinals(const Cta& cta, const DecCfg& c, const Ws& w, const BatchArgs& a, int b,
                         int curIdx, int nFin) {
  const Beam F = w.beam(curIdx);
  for (int q = cta.tid; q < nFin; q += cta.nthr) {
    double* o = a.finScore + ((long long)b * c.K + q) * 3;
    o[0] = F.score(q);
    o[1] = F.am(q);
    o[2] = F.lm(q);
  }
  if (cta.tid == 0) a.finCount[b] = nFin;
}

FLT_DEV FrameIn finishFrameIn(const DecCfg& c, const BatchArgs& a, int b, int len) {
  FrameIn f;
  f.e = nullptr;
  f.topTok = nullptr;
  f.topVal = nullptr;
  f.listLen = 0;
  f.thrVal = 0.0f;
  f.first = 0;
  f.listIsSet = 0;
  f.specReady = 0;
  f.eNext = nullptr;
  f.hRow = len + 1;
  f.hSkip = a.hSkipFin + (long long)b * c.K;
  f.hScore = a.hScore ? a.hScore + (long long)(len + 1) * c.K * 3 : nullptr;
  f.hCount = a.hCount ? a.hCount + (len + 1) : nullptr;
  const long long h = ((long long)b * (a.T + 2) + (len + 1)) * c.K;
  f.hParent = a.hParent + h;
  f.hTok = a.hTok + h;
  f.hWord = a.hWord ? a.hWord + h : nullptr;
  return f;
}

// Online decoding: the live beam (its three workspace arrays and the hypothesis count) is kept in a
// global slot between decodeStep launches. Layout: [int nH, pad to 16][beamD][beamFp][beamI].
FLT_HD size_t streamBeamBytes(const DecCfg& c) {
  const Lay& L = c.lay;
  return 16 + (size_t)(L.beamFp[0] - L.beamD[0]) + (size_t)(L.beamI[0] - L.beamFp[0]) +
         (size_t)(L.beamD[1] - L.beamI[0]);
}
FLT_DEV void streamSaveBeam(const Cta& cta, const DecCfg& c, const Ws& w, const BatchArgs& a, int curIdx) {
  const Lay& L = c.lay;
  const int n = (int)(streamBeamBytes(c) - 16) / 4; // the three arrays are contiguous per beam
  const int* src = (const int*)(w.base + L.beamD[curIdx]);
  int* dst = (int*)(a.streamBeam + 16);
  for (int i = cta.tid; i < n; i += cta.nthr) dst[i] = src[i];
  if (cta.tid == 0) *(int*)a.streamBeam = w.sc()[SC_NH];
}
FLT_DEV void streamRestoreBeam(const Cta& cta, const DecCfg& c, const Ws& w, const BatchArgs& a) {
  const Lay& L = c.lay;
  const int n = (int)(streamBeamBytes(c) - 16) / 4;
  const int* src = (const int*)(a.streamBeam + 16);
  int* dst = (int*)(w.base + L.beamD[0]);
  for (int i = cta.tid; i < n; i += cta.nthr) dst[i] = src[i];
  const int nH = *(const int*)a.streamBeam;
  cta.sync();
  const Beam B0 = w.beam(0);
  for (int i = cta.tid; i < nH; i += cta.nthr) B0.score(i) -= a.streamShift; // Utils.h:339-341
  if (cta.tid == 0) {
    w.sc()[SC_NH] = nH;
    a.status[0] = 0;
  }
}

template <bool W>
FLT_DEV void decodeCta(const Cta& cta, const DecCfg& c, const BatchArgs& a, char* base) {
  const Ws w = wsOf(base, c, a.wsGlobal ? a.wsGlobal + (long long)cta.bid * a.wsStride : nullptr);
  const int K = c.K;
  ctaInitWorkspace(cta, c, w, base);
  for (int b = cta.bid; b < a.B; b += cta.nblk) {
    const int len = a.lengths ? a.lengths[b] : a.T;
    int curIdx = 0;
    cta.sync(); // previous utterance fully retired
    if (a.streamBeam && a.streamRestore) streamRestoreBeam(cta, c, w, a);
    else if (cta.tid == 0) seedUtterance(c, w, a, b);
    LfCarry carry{0.0f, 0.0f, 0.0f, 0};
    if (c.lfFast) { // fingerprint table of the first beam (beam_lf.h keeps one per beam from then on)
      cta.sync();
      lfTabBuild(cta, w, 0, w.sc()[SC_NH]);
    }
    // token list of frame 0 into the workspace
    const long long row0 = (long long)b * a.T;
    if (c.listInSmem && len > 0) {
      for (int j = cta.tid; j < c.M; j += cta.nthr) {
        w.listTok(0)[j] = a.topTok[row0 * c.M + j];
        w.listVal(0)[j] = a.topVal[row0 * c.M + j];
      }
    }
    cta.sync();
    for (int t = 0; t < len; ++t) {
      const long long row = row0 + t;
      // prefetch the next frame's list into registers (<= 2 entries per thread, listInSmem)
      int pfTok[2] = {-1, -1};
      float pfVal[2] = {0.0f, 0.0f};
      const bool pf = c.listInSmem && t + 1 < len;
      if (pf) {
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          const int j = cta.tid + z * cta.nthr;
          if (j < c.M) {
            pfTok[z] = a.topTok[(row + 1) * c.M + j];
            pfVal[z] = a.topVal[(row + 1) * c.M + j];
          }
        }
      }
      FrameIn f;
      f.e = a.emis + row * c.N;
      if (c.listInSmem) {
        f.topTok = w.listTok(t & 1);
        f.topVal = w.listVal(t & 1);
      } else {
        f.topTok = a.topTok ? a.topTok + row * c.M : nullptr;
        f.topVal = a.topVal ? a.topVal + row * c.M : nullptr;
      }
      f.listLen = c.M;
      f.thrVal = a.thrVal ? a.thrVal[row] : 0.0f;
      f.first = a.streamFrame0 + t == 0;
      f.listIsSet = !c.lexicon || c.rootList;
      f.specReady = 0;
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
      // The step gathers a few hundred scattered emissions per frame (trie edges, stay / blank); the
      // select kernel streamed this row long ago, so they would each pay a DRAM round trip. Pull the
      // NEXT frame's row into L2 now, one 128-byte line per thread, while this frame is processed.
      if (f.eNext && !c.lfFast) {
        const char* nx = (const char*)f.eNext;
        for (int off = cta.tid * 128; off < c.N * 4; off += cta.nthr * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + off));
      }
#endif
      if (c.lfFast) lfFrameStep(cta, c, w, curIdx, f, a.stats, carry);
      else frameStep<W>(cta, c, w, w.beam(curIdx), w.beam(curIdx ^ 1), f, a.status + b, a.stats);
      if (pf) {
#if FLT_DEVICE_BUILD
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          const int j = cta.tid + z * cta.nthr;
          if (j < c.M) {
            w.listTok((t + 1) & 1)[j] = pfTok[z];
            w.listVal((t + 1) & 1)[j] = pfVal[z];
          }
        }
#else
        (void)pfTok;
        (void)pfVal;
        for (int j = 0; j < c.M; ++j) { // model: the single thread copies the whole list
          w.listTok((t + 1) & 1)[j] = a.topTok[(row + 1) * c.M + j];
          w.listVal((t + 1) & 1)[j] = a.topVal[(row + 1) * c.M + j];
        }
#endif
      }
      if (w.sc()[SC_NH] == 0) break;
      curIdx ^= 1;
      cta.sync();
    }
    if (a.streamBeam && a.streamNoFinish) { // decodeStep chunk: keep the beam for the next launch
      cta.sync();
      streamSaveBeam(cta, c, w, a, curIdx);
      if (c.lfFast) lfTabClear(cta, w, curIdx, w.sc()[SC_NH]);
      continue;
    }
    int nFin = 0;
    if (w.sc()[SC_NH] != 0) {
      const FrameIn f = finishFrameIn(c, a, b, len);
      if (c.lfFast) lfFinish(cta, c, w, curIdx, f);
      else finishStep<W>(cta, c, w, w.beam(curIdx), w.beam(curIdx ^ 1), f);
      curIdx ^= 1;
      nFin = w.sc()[SC_NH];
    }
    cta.sync();
    writeFinals(cta, c, w, a, b, curIdx, nFin);
  }
}

/* ------------------------------------------------------------------ K4: n-best backtrace ------ */
struct BacktraceArgs {
  const int* hParent;
  const int* hTok;
  const int* hWord; // may be null
  const int* hSkip;    // [B, nCp, K]
  const int* hSkipFin; // [B, K]
  const int* finCount;
  const int* lengths;
  int B, T, K, nbest, nCp;
  int* outTok;  // [B, nbest, T+2]
  int* outWord; // [B, nbest, T+2]
};
constexpr int kBtMaxCp = 288; // checkpoint indices one item keeps in shared memory (T <= ~9200)

// rows (lo, hi] of item (b, r), starting from hypothesis k of row hi
FLT_DEV void backtraceSegment(const BacktraceArgs& a, int b, int* ot, int* ow, int hi, int lo, int k) {
  for (int row = hi; row > lo; --row) {
    const long long h = ((long long)b * (a.T + 2) + row) * a.K + k;
    ot[row] = a.hTok[h];
    ow[row] = a.hWord ? a.hWord[h] : -1;
    k = a.hParent[h];
  }
}

// item = (utterance b, rank r), handled by `nlane` cooperating lanes (a warp; 1 in the host model):
// getHypothesis (Utils.h:229-250); positions past len+1 are -1. cpIdx: >= kBtMaxCp ints of scratch.
FLT_DEV void backtraceItem(const BacktraceArgs& a, long long item, int lane, int nlane, int* cpIdx) {
  const int b = (int)(item / a.nbest), r = (int)(item % a.nbest);
  const int len = a.lengths ? a.lengths[b] : a.T;
  int* ot = a.outTok + item * (a.T + 2);
  int* ow = a.outWord + item * (a.T + 2);
  for (int i = len + 2 + lane; i < a.T + 2; i += nlane) {
    ot[i] = -1;
    ow[i] = -1;
  }
  if (r >= a.finCount[b]) {
    for (int i = lane; i < len + 2 && i < a.T + 2; i += nlane) {
      ot[i] = -1;
      ow[i] = -1;
    }
    return;
  }
  const int F = len + 1;               // finish row
  const int J = (F - 1) >> kCpShift;   // its checkpoint row is kCpRows * J
  if (J + 1 > kBtMaxCp) {              // very long utterance: plain walk
    if (lane == 0) backtraceSegment(a, b, ot, ow, F, -1, r);
    return;
  }
  if (lane == 0) { // hop checkpoint to checkpoint: cpIdx[j] = index of the ancestor in row kCpRows * j
    int k = a.hSkipFin[(long long)b * a.K + r];
    cpIdx[J] = k;
    for (int j = J; j >= 1; --j) {
      k = a.hSkip[((long long)b * a.nCp + j) * a.K + k];
      cpIdx[j - 1] = k;
    }
  }
#if FLT_DEVICE_BUILD
  __syncwarp();
#endif
  // segments: s = J + 1 is (kCpRows * J, F] from r; s = 1..J is (kCpRows (s-1), kCpRows s]; s = 0 is row 0
  for (int s = lane; s <= J + 1; s += nlane) {
    if (s == J + 1) backtraceSegment(a, b, ot, ow, F, J << kCpShift, r);
    else if (s == 0) backtraceSegment(a, b, ot, ow, 0, -1, cpIdx[0]);
    else backtraceSegment(a, b, ot, ow, s << kCpShift, (s - 1) << kCpShift, cpIdx[s]);
  }
}

} // namespace flt
