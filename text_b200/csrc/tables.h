// tables.h — flattened, read-only lookup tables the decode kernels gather from (HBM, L2-resident
// at the benchmark sizes): the lexicon Trie as CSR arrays and the n-gram LM as per-order
// open-addressing tables. Replaces, on the hot path, `TrieNode::children.find` / `labels` /
// `maxScore` (decoder/Trie.h:39-54, decoder/LexiconDecoder.cpp:58-67,113) and
// `KenLM::score` -> `lm::base::Model::BaseScore` (decoder/lm/KenLM.cpp:63-83).
#pragma once
#include "spmd.h"

namespace flt {

constexpr int kMaxOrder = 6;          // FL_TEXT_KENLM_MAX_ORDER, decoder/lm/CMakeLists.txt:3
constexpr int kMaxCtx = kMaxOrder - 1;

struct TrieDev {
  int nNodes;
  const int* childOff;    // [nNodes+1] CSR row offsets
  const int* childTok;    // [nEdges]   token of each edge, ascending within a node
  const int* childNode;   // [nEdges]   target node
  const float* maxScore;  // [nNodes]   smeared score (TrieNode::maxScore)
  const int* labelOff;    // [nNodes+1]
  const int* labels;      // [nLabels]  word ids (TrieNode::labels, <= 6 per node)
  const int* rootChild;   // [N]        node reached from the root by token n, or -1
  int nRootLab;           // root children that carry labels (single-token words)
  const int* rootLabTok;  // [nRootLab] their tokens
  // packed views of the same Trie for the single-pass step (beam_gx.h): one 8-byte record per edge,
  // one 16-byte record per node, so an edge costs two dependent loads instead of six
  const int2* edge;        // [nEdges]   {token, child node}
  const int4* node;        // [nNodes]   {smeared score bits, first edge, #edges | #labels << 24,
                           //             the label if there is exactly one, else the offset into labels[]}
  const int2* rootLabEdge; // [nRootLab] {token, child node} of the root children that carry labels
};

struct F2 {
  float x, y; // (log10 prob, back-off)
};

struct LmDev {
  int kind; // 0 = ZeroLM, 1 = n-gram
  int order, vocab, bos, eos, nUsr;
  const int* usr2lm;              // [nUsr] user word index -> LM vocabulary id (OOV -> 0 = <unk>)
  const F2* uni;                  // [vocab]
  const uint64_t* keys[kMaxOrder + 1]; // [n] -> open-addressing table of order n (n >= 2); 0 = empty
  const uint64_t* chk[kMaxOrder + 1];  // [n] -> second, independent 64-bit key of the same slot (verification)
  const F2* vals[kMaxOrder + 1];
  uint32_t mask[kMaxOrder + 1];
};

// Key of an n-gram = chained 64-bit mix over its words in REVERSED order (most recent first), so
// that growing the context by one older word extends the chain. KenLM's probing model likewise
// keys n-grams by a 64-bit hash of the word ids. A second chain with other constants is stored next to
// every entry and compared on a hit, so an n-gram is identified by 128 bits: two distinct n-grams (or
// a queried one that is not in the model and a stored one) are confused with probability 2^-128 per
// pair — against 2^-64 with one key, which at 5 M entries and ~1e9 probes per batch would have been a
// wrong LM score every few thousand batches.
FLT_HD uint64_t ngramChainStart(int w) { return mix64(0x9E3779B97F4A7C15ull + (uint64_t)(uint32_t)w); }
FLT_HD uint64_t ngramChainExtend(uint64_t h, int w) {
  return mix64(h * 0x100000001B3ull + (uint64_t)(uint32_t)w + 0x632BE59BD9B4E019ull);
}
FLT_HD uint64_t ngramFinalKey(uint64_t h) { return h == 0 ? 1 : h; }
FLT_HD uint64_t ngramChain2Start(int w) { return mix64(0xD6E8FEB86659FD93ull ^ ((uint64_t)(uint32_t)w << 1)); }
FLT_HD uint64_t ngramChain2Extend(uint64_t h, int w) {
  return mix64((h ^ 0xC2B2AE3D27D4EB4Full) * 0xFF51AFD7ED558CCDull + (uint64_t)(uint32_t)w);
}

FLT_HD bool ngramFind(const LmDev& lm, int n, uint64_t chain, uint64_t chain2, F2& out) {
  const uint64_t key = ngramFinalKey(chain);
  const uint32_t mask = lm.mask[n];
  const uint64_t* keys = lm.keys[n];
  if (!keys) return false;
  uint32_t s = (uint32_t)(key >> 17) & mask;
  for (;;) {
    uint64_t k = keys[s];
    if (k == key && lm.chk[n][s] == chain2) {
      out = lm.vals[n][s];
      return true;
    }
    if (k == 0) return false;
    s = (s + 1) & mask;
  }
}

// p(w | ctx) in log10 as a float, ctx most-recent-first. Same evaluation order as KenLM's
// FullScore: longest match grown one word at a time, then back-offs added from the shortest
// unused context to the longest, all in float.
FLT_HD float ngramScore(const LmDev& lm, const int* ctx, int nctx, int w) {
  const int L = nctx < lm.order - 1 ? nctx : lm.order - 1;
  float ret = lm.uni[w].x;
  int matchLen = 1;
  uint64_t h = ngramChainStart(w), h2 = ngramChain2Start(w);
  for (int k = 1; k <= L; ++k) {
    h = ngramChainExtend(h, ctx[k - 1]);
    h2 = ngramChain2Extend(h2, ctx[k - 1]);
    F2 v;
    if (!ngramFind(lm, k + 1, h, h2, v)) break;
    ret = v.x;
    matchLen = k + 1;
  }
  if (matchLen - 1 < L) {
    uint64_t g = 0, g2 = 0;
    for (int i = 0; i < L; ++i) {
      g = i == 0 ? ngramChainStart(ctx[0]) : ngramChainExtend(g, ctx[i]);
      g2 = i == 0 ? ngramChain2Start(ctx[0]) : ngramChain2Extend(g2, ctx[i]);
      if (i < matchLen - 1) continue;
      if (i == 0) {
        ret += lm.uni[ctx[0]].y;
      } else {
        F2 v;
        if (ngramFind(lm, i + 1, g, g2, v)) ret += v.y;
      }
    }
  }
  return ret;
}

// The back-off weights of a context, which ngramScore adds whenever the n-gram ending in the scored word is
// shorter than the context: they depend on the LM state only, so a decoder computes them once when the state
// is created and every word scored from that state reuses them (typically one table probe per word instead
// of three). bo[i] = back-off of the context's first i+1 words (bo[0] from the unigram table); bit i of the
// returned mask says whether that context is in the model (absent = nothing is added, as in ngramScore).
FLT_HD int ngramContextBackoffs(const LmDev& lm, const int* ctx, int nctx, float* bo) {
  const int L = nctx < lm.order - 1 ? nctx : lm.order - 1;
  int mask = 0;
  uint64_t g = 0, g2 = 0;
  for (int i = 0; i < L; ++i) {
    g = i == 0 ? ngramChainStart(ctx[0]) : ngramChainExtend(g, ctx[i]);
    g2 = i == 0 ? ngramChain2Start(ctx[0]) : ngramChain2Extend(g2, ctx[i]);
    bo[i] = 0.0f;
    if (i == 0) {
      bo[0] = lm.uni[ctx[0]].y;
      mask |= 1;
    } else {
      F2 v;
      if (ngramFind(lm, i + 1, g, g2, v)) {
        bo[i] = v.y;
        mask |= 1 << i;
      }
    }
  }
  return mask;
}
// ngramScore with the context's back-offs taken from ngramContextBackoffs (same float operations in the same
// order: bit-identical)
FLT_HD float ngramScoreCached(const LmDev& lm, const int* ctx, int nctx, int w, const float* bo, int mask) {
  const int L = nctx < lm.order - 1 ? nctx : lm.order - 1;
  float ret = lm.uni[w].x;
  int matchLen = 1;
  uint64_t h = ngramChainStart(w), h2 = ngramChain2Start(w);
  for (int k = 1; k <= L; ++k) {
    h = ngramChainExtend(h, ctx[k - 1]);
    h2 = ngramChain2Extend(h2, ctx[k - 1]);
    F2 v;
    if (!ngramFind(lm, k + 1, h, h2, v)) break;
    ret = v.x;
    matchLen = k + 1;
  }
  for (int i = matchLen - 1; i < L; ++i)
    if ((mask >> i) & 1) ret += bo[i];
  return ret;
}

// Context of the child state: (w, ctx...) truncated to order-1 words.
FLT_HD int ngramAdvanceCtx(const LmDev& lm, const int* ctx, int nctx, int w, int* out) {
  const int keep = lm.order - 1;
  int n = 0;
  if (keep > 0) {
    out[n++] = w;
    for (int i = 0; i < nctx && n < keep; ++i) out[n++] = ctx[i];
  }
  return n;
}

} // namespace flt
