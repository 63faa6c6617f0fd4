// ORACLE (test infrastructure, never shipped): CPU restatement of the reference's
// LexiconDecoder / LexiconFreeDecoder hot path. Exports the ora_* half of oracle_api.h.
//
// This is a from-scratch restatement (index arenas instead of shared_ptr graphs, interned
// LM-state ids instead of pointer identity) of the algorithm in
//   flashlight/lib/text/decoder/Utils.h:121-225,229-342      candidates*, backtrace, prune
//   flashlight/lib/text/decoder/LexiconFreeDecoder.cpp:20-227
//   flashlight/lib/text/decoder/LexiconDecoder.cpp:21-325
//   flashlight/lib/text/decoder/Trie.cpp:26-101
//   flashlight/lib/text/decoder/lm/LM.h:21-50, lm/ZeroLM.cpp:14-26, lm/KenLM.cpp:32-83
// PARITY PINNED: validated against (a) the reference's DecoderTest known answers
// (test/decoder/DecoderTest.cpp:107-120,148-155,184,190-194) and (b) the unmodified reference
// compiled in place (oracle/_ref/libflref.so) on seeded inputs — tests/test_oracle_*.py.
//
// Arithmetic follows SURVEY.md Appendix A.4: FP64 accumulators over FP32 emissions, the
// float-typed sub-expressions kept in float, no FMA contraction (-ffp-contract=off).
// Tie-breaking (equal scores at a cut, equal-score members of one merge group) is
// implementation-defined in the reference (libstdc++ sort internals / heap pointer order);
// here it is deterministic: earlier-created candidate first.

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "arpa_lm.hpp"
#include "oracle_api.h"

namespace {

thread_local std::string gErr;
const double kNegInf = -std::numeric_limits<double>::infinity();
const int kLookBackLimit = 100; // Utils.h:28
const int kTrieMaxLabel = 6;    // Trie.h:19

/* ------------------------------------------------------------------ Trie (Trie.cpp:26-101) */
struct TNode {
  std::unordered_map<int, int> kids; // token -> node index
  std::vector<int> kidOrder;         // insertion order (deterministic smear order)
  std::vector<int> labels;
  std::vector<float> scores;
  float maxScore = 0;
};
struct OTrie {
  int maxChildren;
  int rootIdx;
  std::vector<TNode> nodes;
  OTrie(int mc, int r) : maxChildren(mc), rootIdx(r), nodes(1) {}

  int insert(const int* idx, int n, int label, float score) {
    int cur = 0;
    for (int i = 0; i < n; ++i) {
      if (idx[i] < 0 || idx[i] >= maxChildren) {
        gErr = "[Trie] Invalid letter index: " + std::to_string(idx[i]);
        return -1;
      }
      auto it = nodes[cur].kids.find(idx[i]);
      if (it == nodes[cur].kids.end()) {
        int nn = (int)nodes.size();
        nodes[cur].kids.emplace(idx[i], nn);
        nodes[cur].kidOrder.push_back(idx[i]);
        nodes.emplace_back();
        cur = nn;
      } else {
        cur = it->second;
      }
    }
    if ((int)nodes[cur].labels.size() < kTrieMaxLabel) {
      nodes[cur].labels.push_back(label);
      nodes[cur].scores.push_back(score);
    }
    return cur;
  }
  int search(const int* idx, int n) const {
    int cur = 0;
    for (int i = 0; i < n; ++i) {
      if (idx[i] < 0 || idx[i] >= maxChildren) {
        gErr = "[Trie] Invalid letter index: " + std::to_string(idx[i]);
        return -2;
      }
      auto it = nodes[cur].kids.find(idx[i]);
      if (it == nodes[cur].kids.end()) return -1;
      cur = it->second;
    }
    return cur;
  }
  static double logAdd(double a, double b) { // Trie.cpp:66-77
    if (a < b) std::swap(a, b);
    double d = b - a;
    if (d < -39.14) return a;
    return a + std::log1p(std::exp(d));
  }
  void smearNode(int n, int mode) { // Trie.cpp:79-95 (maxScore is a float at every step)
    nodes[n].maxScore = -std::numeric_limits<float>::infinity();
    for (float s : nodes[n].scores) nodes[n].maxScore = (float)logAdd(nodes[n].maxScore, s);
    for (size_t k = 0; k < nodes[n].kidOrder.size(); ++k) {
      int c = nodes[n].kids[nodes[n].kidOrder[k]];
      smearNode(c, mode);
      if (mode == 2) {
        nodes[n].maxScore = (float)logAdd(nodes[n].maxScore, nodes[c].maxScore);
      } else if (mode == 1 && nodes[c].maxScore > nodes[n].maxScore) {
        nodes[n].maxScore = nodes[c].maxScore;
      }
    }
  }
  void smear(int mode) {
    if (mode != 0) smearNode(0, mode);
  }
};

/* ------------------------------------------------- LM states and LMs (LM.h, ZeroLM, KenLM) */
struct OLM {
  bool zero = true;
  std::unique_ptr<oracle::ArpaLM> arpa;
  std::vector<int> usr2lm;
};

// LMState identity = pointer identity of a node in the child tree (LM.h:24-49); restated as an
// interned id per (parent id, label). Lives for one decodeBegin..decodeBegin span.
struct StateArena {
  std::unordered_map<uint64_t, int> child;
  std::vector<std::array<int, oracle::kMaxOrder>> ctx;
  std::vector<int> nctx;
  bool needCtx = true; // ZeroLM states carry no context: skip the storage
  int count = 0;
  void clear() {
    child.clear();
    ctx.clear();
    nctx.clear();
    count = 0;
  }
  int fresh() {
    if (needCtx) {
      ctx.emplace_back();
      nctx.push_back(0);
    }
    return count++;
  }
  int childOf(int s, int label, bool& created) {
    uint64_t k = ((uint64_t)(uint32_t)s << 32) | (uint32_t)label;
    auto it = child.find(k);
    if (it != child.end()) {
      created = false;
      return it->second;
    }
    int id = fresh();
    child.emplace(k, id);
    created = true;
    return id;
  }
};

struct LMSession {
  const OLM* lm;
  StateArena A;
  int start(bool withNothing) {
    int s = A.fresh();
    if (!lm->zero && !withNothing) {
      A.ctx[s][0] = lm->arpa->bos();
      A.nctx[s] = 1;
    }
    return s;
  }
  void advance(int in, int w, int out, float& sc) {
    const auto& m = *lm->arpa;
    sc = m.score(A.ctx[in].data(), A.nctx[in], w);
    int keep = m.order() - 1, n = 0;
    std::array<int, oracle::kMaxOrder> c{};
    if (keep > 0) {
      c[n++] = w;
      for (int i = 0; i < A.nctx[in] && n < keep; ++i) c[n++] = A.ctx[in][i];
    }
    A.ctx[out] = c;
    A.nctx[out] = n;
  }
  // returns false on invalid index (KenLM.cpp:66-69 throws)
  bool score(int state, int usrIdx, int& out, float& sc) {
    bool created;
    if (lm->zero) { // ZeroLM.cpp:18-22
      out = A.childOf(state, usrIdx, created);
      sc = 0.0f;
      return true;
    }
    if (usrIdx < 0 || usrIdx >= (int)lm->usr2lm.size()) {
      gErr = "[KenLM] Invalid user token index: " + std::to_string(usrIdx);
      return false;
    }
    out = A.childOf(state, usrIdx, created);
    advance(state, lm->usr2lm[usrIdx], out, sc); // idempotent when the child already existed
    return true;
  }
  void finish(int state, int& out, float& sc) {
    if (lm->zero) { // ZeroLM.cpp:24-26: same state
      out = state;
      sc = 0.0f;
      return;
    }
    bool created;
    out = A.childOf(state, -1, created); // KenLM.cpp:77-83
    advance(state, lm->arpa->eos(), out, sc);
  }
};

/* ------------------------------------------------------------------------------- decoder */
struct Hyp {
  double score = 0;
  int lmState = -1;
  int lex = 0;      // trie node (0 = root); unused by the lexicon-free decoder
  int parent = -1;  // index into the previous frame's vector, -1 = none
  int token = -1;
  int word = -1;
  bool prevBlank = false;
  double am = 0;
  double lm = 0;
};

inline int cmpKey(const Hyp& a, const Hyp& b) { // LexiconDecoder.h:79-91 / LexiconFreeDecoder.h:68-78
  if (a.lmState != b.lmState) return a.lmState > b.lmState ? 1 : -1;
  if (a.lex != b.lex) return a.lex > b.lex ? 1 : -1;
  if (a.token != b.token) return a.token > b.token ? 1 : -1;
  if (a.prevBlank != b.prevBlank) return a.prevBlank > b.prevBlank ? 1 : -1;
  return 0;
}

struct ODecoder {
  bool lexicon = false;
  ora_options opt{};
  const OTrie* trie = nullptr;
  const OLM* lm = nullptr;
  int sil = 0, blank = -1, unk = -1;
  std::vector<float> trans;
  bool isLmToken = false;

  LMSession S;
  std::vector<std::vector<Hyp>> hyp;
  int nDecoded = 0, nPruned = 0;
  bool begun = false;

  std::vector<Hyp> cands;
  double best = kNegInf;
  std::vector<int> order;
  bool failed = false;
  // Tie detector (test infrastructure): events since begin() where the reference's result is left
  // to libstdc++ internals - equal-score members of one merge group (which parent survives),
  // equal scores straddling the beam cut (nth_element), equal emissions at the token-beam cut.
  long tieEvents = 0;

  void add(const Hyp& h) { // Utils.h:131-144
    if (h.score >= best) best = h.score;
    if (h.score >= best - opt.beamThreshold) cands.push_back(h);
  }

  void store(std::vector<Hyp>& out, bool /*sorted*/) { // Utils.h:146-225
    out.clear();
    if (cands.empty()) return;
    const double thr = best - opt.beamThreshold;
    order.clear();
    for (int i = 0; i < (int)cands.size(); ++i)
      if (cands[i].score >= thr) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      int c = cmpKey(cands[a], cands[b]);
      return c == 0 ? cands[a].score > cands[b].score : c > 0;
    });
    // logAdd: another implementation's exp/log1p may differ from this libm in the last bit, so
    // scores closer than `near` count as ties too (which member leads a group, who makes the cut)
    auto nearTie = [&](double a, double b) {
      return opt.logAdd && std::fabs(a - b) <= 1e-9 * std::max(1.0, std::fabs(a));
    };
    int n = 1;
    bool headFresh = true; // head.score is still the group's best member's own score
    for (int i = 1; i < (int)order.size(); ++i) {
      Hyp& head = cands[order[n - 1]];
      const Hyp& cur = cands[order[i]];
      if (cmpKey(cur, head) != 0) {
        order[n++] = order[i];
        headFresh = true;
      } else {
        if (cur.score == head.score && !opt.logAdd) ++tieEvents;
        if (headFresh && nearTie(cur.score, head.score)) ++tieEvents;
        headFresh = false;
        double mx = std::max(head.score, cur.score);
        if (opt.logAdd) {
          double mn = std::min(head.score, cur.score);
          head.score = mx + std::log1p(std::exp(mn - mx));
        } else {
          head.score = mx;
        }
      }
    }
    order.resize(std::min<size_t>(n, order.size()));
    n = (int)order.size();
    // select the beamSize best; deterministic: score desc, then creation order
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (cands[a].score != cands[b].score) return cands[a].score > cands[b].score;
      return a < b;
    });
    int fin = std::min(n, (int)opt.beamSize);
    if (fin < n && (cands[order[fin - 1]].score == cands[order[fin]].score ||
                    nearTie(cands[order[fin - 1]].score, cands[order[fin]].score)))
      ++tieEvents;
    for (int i = 0; i < fin; ++i) out.push_back(cands[order[i]]);
  }

  void begin() {
    S.lm = lm;
    S.A.clear();
    S.A.needCtx = !lm->zero;
    hyp.clear();
    hyp.emplace_back();
    Hyp h;
    h.score = 0.0;
    h.lmState = S.start(false);
    h.lex = 0;
    h.parent = -1;
    h.token = sil;
    h.word = -1;
    hyp[0].push_back(h);
    nDecoded = nPruned = 0;
    begun = true;
    failed = false;
    tieEvents = 0;
  }

  void selectTokens(const float* e, int N, std::vector<int>& idx) { // *Decoder.cpp "partial_sort"
    idx.resize(N);
    std::iota(idx.begin(), idx.end(), 0);
    if (N > opt.beamSizeToken) {
      std::partial_sort(idx.begin(), idx.begin() + opt.beamSizeToken, idx.end(),
                        [&](int l, int r) { return e[l] != e[r] ? e[l] > e[r] : l < r; });
      float cut = e[idx[opt.beamSizeToken - 1]];
      int atCut = 0;
      for (int i = 0; i < N; ++i) atCut += e[i] == cut;
      if (atCut > 1) ++tieEvents;
      idx.resize(opt.beamSizeToken);
    }
  }

  void stepFree(const float* e, int N, int frame, bool first, const std::vector<int>& idx) {
    const bool ctc = opt.criterion == 1, asg = opt.criterion == 0;
    const std::vector<Hyp>& prev = hyp[frame];
    for (int p = 0; p < (int)prev.size(); ++p) {
      const Hyp& ph = prev[p];
      const int prevIdx = ph.token;
      for (int n : idx) {
        double am = e[n];
        if (!first && asg) am += trans[(size_t)n * N + prevIdx];
        double score = ph.score + e[n]; // transitions reach only `am` (LexiconFreeDecoder.cpp:59-64)
        if (n == sil) score += opt.silScore;
        Hyp c;
        c.parent = p;
        c.token = n;
        c.am = ph.am + am;
        if ((asg && n != prevIdx) || (ctc && n != blank && (n != prevIdx || ph.prevBlank))) {
          int ns;
          float ls;
          if (!S.score(ph.lmState, n, ns, ls)) { failed = true; return; }
          c.score = score + opt.lmWeight * ls;
          c.lmState = ns;
          c.prevBlank = false;
          c.lm = ph.lm + ls;
        } else if (ctc && n == blank) {
          c.score = score;
          c.lmState = ph.lmState;
          c.prevBlank = true;
          c.lm = ph.lm;
        } else {
          c.score = score;
          c.lmState = ph.lmState;
          c.prevBlank = false;
          c.lm = ph.lm;
        }
        add(c);
      }
    }
  }

  void stepLex(const float* e, int N, int frame, bool first, const std::vector<int>& idx) {
    const bool ctc = opt.criterion == 1, asg = opt.criterion == 0;
    const std::vector<Hyp>& prev = hyp[frame];
    for (int p = 0; p < (int)prev.size(); ++p) {
      const Hyp& ph = prev[p];
      const TNode& pl = trie->nodes[ph.lex];
      const int prevIdx = ph.token;
      const float lexMax = ph.lex == 0 ? 0 : pl.maxScore;

      for (int n : idx) { // (1) children, LexiconDecoder.cpp:62-164
        auto it = pl.kids.find(n);
        if (it == pl.kids.end()) continue;
        const int ci = it->second;
        const TNode& cl = trie->nodes[ci];
        double am = e[n];
        if (!first && asg) am += trans[(size_t)n * N + prevIdx];
        double score = ph.score + am;
        if (n == sil) score += opt.silScore;

        int lmState = -1;
        double lmScore = 0.;
        if (isLmToken) {
          float ls;
          if (!S.score(ph.lmState, n, lmState, ls)) { failed = true; return; }
          lmScore = ls;
        }
        Hyp c;
        c.parent = p;
        c.token = n;
        c.prevBlank = false;
        c.am = ph.am + am;
        if (!ctc || ph.prevBlank || n != prevIdx) {
          if (!cl.kids.empty()) {
            if (!isLmToken) {
              lmState = ph.lmState;
              lmScore = cl.maxScore - lexMax; // float - float, widened after
            }
            c.score = score + opt.lmWeight * lmScore;
            c.lmState = lmState;
            c.lex = ci;
            c.word = -1;
            c.lm = ph.lm + lmScore;
            add(c);
          }
        }
        for (int label : cl.labels) {
          if (ph.lex == 0 && ph.token == n) continue;
          if (!isLmToken) {
            float ls;
            if (!S.score(ph.lmState, label, lmState, ls)) { failed = true; return; }
            lmScore = ls - lexMax;
          }
          c.score = score + opt.lmWeight * lmScore + opt.wordScore;
          c.lmState = lmState;
          c.lex = 0;
          c.word = label;
          c.lm = ph.lm + lmScore;
          add(c);
        }
        if (cl.labels.empty() && opt.unkScore > kNegInf) {
          if (!isLmToken) {
            float ls;
            if (!S.score(ph.lmState, unk, lmState, ls)) { failed = true; return; }
            lmScore = ls - lexMax;
          }
          c.score = score + opt.lmWeight * lmScore + opt.unkScore;
          c.lmState = lmState;
          c.lex = 0;
          c.word = unk;
          c.lm = ph.lm + lmScore;
          add(c);
        }
      }

      if (!ctc || !ph.prevBlank || ph.lex == 0) { // (2) same node, :167-194
        int n = ph.lex == 0 ? sil : prevIdx;
        double am = e[n];
        if (!first && asg) am += trans[(size_t)n * N + prevIdx];
        double score = ph.score + am;
        if (n == sil) score += opt.silScore;
        Hyp c;
        c.score = score;
        c.lmState = ph.lmState;
        c.lex = ph.lex;
        c.parent = p;
        c.token = n;
        c.word = -1;
        c.prevBlank = false;
        c.am = ph.am + am;
        c.lm = ph.lm;
        add(c);
      }
      if (ctc) { // (3) blank, :196-213
        double am = e[blank];
        Hyp c;
        c.score = ph.score + am;
        c.lmState = ph.lmState;
        c.lex = ph.lex;
        c.parent = p;
        c.token = blank;
        c.word = -1;
        c.prevBlank = true;
        c.am = ph.am + am;
        c.lm = ph.lm;
        add(c);
      }
    }
  }

  void step(const float* emis, int T, int N) {
    if (!begun) begin();
    int startFrame = nDecoded - nPruned;
    if ((int)hyp.size() < startFrame + T + 2) hyp.resize(startFrame + T + 2);
    std::vector<int> idx;
    for (int t = 0; t < T && !failed; ++t) {
      const float* e = emis + (size_t)t * N;
      selectTokens(e, N, idx);
      cands.clear();
      best = kNegInf;
      bool first = !(nDecoded + t > 0);
      if (lexicon) stepLex(e, N, startFrame + t, first, idx);
      else stepFree(e, N, startFrame + t, first, idx);
      store(hyp[startFrame + t + 1], false);
    }
    nDecoded += T;
  }

  void end() {
    int f = nDecoded - nPruned;
    if ((int)hyp.size() < f + 2) hyp.resize(f + 2);
    cands.clear();
    best = kNegInf;
    const std::vector<Hyp>& prev = hyp[f];
    bool nice = false;
    if (lexicon)
      for (const Hyp& h : prev)
        if (h.lex == 0) { nice = true; break; }
    for (int p = 0; p < (int)prev.size(); ++p) {
      const Hyp& ph = prev[p];
      if (lexicon && nice && ph.lex != 0) continue;
      int ns;
      float ls;
      S.finish(ph.lmState, ns, ls);
      Hyp c;
      c.score = ph.score + opt.lmWeight * ls;
      c.lmState = ns;
      c.lex = ph.lex;
      c.parent = p;
      c.token = sil;
      c.word = -1;
      c.prevBlank = false;
      c.am = ph.am;
      c.lm = ph.lm + ls;
      add(c);
    }
    store(hyp[f + 1], true);
    ++nDecoded;
  }

  // Utils.h:229-250: walk parents; length finalFrame+1
  void backtrace(int frame, int index, int finalFrame, std::vector<int>& tokens,
                 std::vector<int>& words) const {
    tokens.assign(finalFrame + 1, -1);
    words.assign(finalFrame + 1, -1);
    int i = 0, f = frame, k = index;
    while (k >= 0 && f >= 0) {
      const Hyp& h = hyp[f][k];
      if (finalFrame - i >= 0) {
        words[finalFrame - i] = lexicon ? h.word : -1;
        tokens[finalFrame - i] = h.token;
      }
      k = h.parent;
      --f;
      ++i;
    }
  }

  bool isComplete(int frame, int k) const { // LexiconDecoder.h:97-99 / LexiconFreeDecoder.h:84-86
    if (!lexicon) return true;
    const Hyp& h = hyp[frame][k];
    return h.parent < 0 || hyp[frame - 1][h.parent].word >= 0;
  }

  // Utils.h:268-310; returns (frame,index) of the ancestor or (-1,-1); updates lookBack
  std::pair<int, int> bestAncestor(int finalFrame, int& lookBack) const {
    const std::vector<Hyp>& fin = hyp[finalFrame];
    if (fin.empty()) return {-1, -1};
    int bk = 0;
    for (int r = 1; r < (int)fin.size(); ++r)
      if (fin[r].score > fin[bk].score) bk = r;
    int f = finalFrame, k = bk, n = 0;
    auto up = [&]() {
      k = hyp[f][k].parent;
      --f;
      if (k < 0) f = -1;
    };
    while (k >= 0 && n < lookBack) {
      ++n;
      up();
    }
    const int maxLB = lookBack + kLookBackLimit;
    while (k >= 0) {
      if (isComplete(f, k)) break;
      ++n;
      up();
      if (n == maxLB) break;
    }
    lookBack = n;
    return {k >= 0 ? f : -1, k};
  }

  void prune(int lookBack) { // *Decoder.cpp prune + Utils.h:312-342
    if (nDecoded - nPruned - lookBack < 1) return;
    int finalFrame = nDecoded - nPruned;
    auto anc = bestAncestor(finalFrame, lookBack);
    if (anc.second < 0) return;
    int startFrame = nDecoded - nPruned - lookBack;
    if (startFrame < 1) return;
    for (int i = 0; i < (int)hyp.size(); ++i) {
      if (i <= lookBack) hyp[i].swap(hyp[i + startFrame]);
      else hyp[i].clear();
    }
    for (Hyp& h : hyp[0]) h.parent = -1;
    double largest = hyp[lookBack].front().score;
    for (const Hyp& h : hyp[lookBack]) largest = std::max(largest, h.score);
    for (Hyp& h : hyp[lookBack]) h.score -= largest;
    nPruned = nDecoded - lookBack;
  }
};

int fillOne(const ODecoder& d, int frame, int k, int finalFrame, int stride, double* scores3,
            int* tokens, int* words) {
  std::vector<int> tk, wd;
  d.backtrace(frame, k, finalFrame, tk, wd);
  const Hyp& h = d.hyp[frame][k];
  scores3[0] = h.score;
  scores3[1] = h.am;
  scores3[2] = h.lm;
  for (int j = 0; j <= finalFrame && j < stride; ++j) {
    tokens[j] = tk[j];
    words[j] = wd[j];
  }
  return finalFrame + 1;
}

int fillAll(const ODecoder& d, int maxHyp, int stride, double* scores3, int* tokens, int* words,
            int* len) {
  int finalFrame = d.nDecoded - d.nPruned;
  if (!d.begun || finalFrame < 1 || finalFrame >= (int)d.hyp.size()) return 0;
  const auto& fin = d.hyp[finalFrame];
  for (int r = 0; r < (int)fin.size() && r < maxHyp; ++r) {
    int L = fillOne(d, finalFrame, r, finalFrame, stride, scores3 + 3 * r,
                    tokens + (size_t)r * stride, words + (size_t)r * stride);
    if (len) len[r] = L;
  }
  return (int)fin.size();
}

ODecoder* makeDecoder(int lexicon, const ora_options* opt, void* trie, void* lm, int sil,
                      int blank, int unk, const float* trans, int nTrans, int isLmToken) {
  auto* d = new ODecoder;
  d->lexicon = lexicon != 0;
  d->opt = *opt;
  d->trie = (const OTrie*)trie;
  d->lm = (const OLM*)lm;
  d->sil = sil;
  d->blank = blank;
  d->unk = unk;
  if (trans && nTrans > 0) d->trans.assign(trans, trans + nTrans);
  d->isLmToken = isLmToken != 0;
  return d;
}

} // namespace

extern "C" {

const char* ora_last_error(void) { return gErr.c_str(); }

void* ora_trie_create(int maxChildren, int rootIdx) { return new OTrie(maxChildren, rootIdx); }
int ora_trie_insert(void* trie, const int* idx, int n, int label, float score) {
  return ((OTrie*)trie)->insert(idx, n, label, score) < 0 ? -1 : 0;
}
void ora_trie_smear(void* trie, int mode) { ((OTrie*)trie)->smear(mode); }
int ora_trie_search(void* trie, const int* idx, int n, float* maxScore, int* nLabels,
                    int* labels6, float* scores6) {
  auto* t = (OTrie*)trie;
  int k = t->search(idx, n);
  if (k == -2) return -1;
  if (k < 0) return 0;
  const TNode& nd = t->nodes[k];
  if (maxScore) *maxScore = nd.maxScore;
  if (nLabels) *nLabels = (int)nd.labels.size();
  for (size_t i = 0; i < nd.labels.size() && i < 6; ++i) {
    if (labels6) labels6[i] = nd.labels[i];
    if (scores6) scores6[i] = nd.scores[i];
  }
  return 1;
}
void ora_trie_destroy(void* trie) { delete (OTrie*)trie; }

void* ora_lm_zero(void) { return new OLM; }
void* ora_lm_arpa(const char* path, const char* const* words, int nWords) {
  try {
    auto* m = new OLM;
    m->zero = false;
    m->arpa = std::make_unique<oracle::ArpaLM>(path);
    m->usr2lm.resize(nWords);
    for (int i = 0; i < nWords; ++i) m->usr2lm[i] = m->arpa->index(words[i]);
    return m;
  } catch (const std::exception& e) {
    gErr = e.what();
    return nullptr;
  }
}
int ora_lm_score_seq(void* lm, const int* usrIdx, int n, int withFinish, float* out) {
  LMSession S;
  S.lm = (const OLM*)lm;
  S.A.needCtx = !S.lm->zero;
  int st = S.start(false);
  for (int i = 0; i < n; ++i) {
    int ns;
    if (!S.score(st, usrIdx[i], ns, out[i])) return -1;
    st = ns;
  }
  if (withFinish) {
    int ns;
    S.finish(st, ns, out[n]);
  }
  return 0;
}
void ora_lm_destroy(void* lm) { delete (OLM*)lm; }

void* ora_decoder_lexfree(const ora_options* opt, void* lm, int sil, int blank,
                          const float* trans, int nTrans) {
  return makeDecoder(0, opt, nullptr, lm, sil, blank, -1, trans, nTrans, 0);
}
void* ora_decoder_lexicon(const ora_options* opt, void* trie, void* lm, int sil, int blank,
                          int unk, const float* trans, int nTrans, int isLmToken) {
  return makeDecoder(1, opt, trie, lm, sil, blank, unk, trans, nTrans, isLmToken);
}
void ora_decoder_destroy(void* dec) { delete (ODecoder*)dec; }

int ora_decode(void* dec, const float* emis, int T, int N, int maxHyp, double* scores3,
               int* tokens, int* words) {
  auto* d = (ODecoder*)dec;
  d->begin();
  d->step(emis, T, N);
  if (d->failed) return -1;
  d->end();
  return fillAll(*d, maxHyp, T + 2, scores3, tokens, words, nullptr);
}
void ora_decode_begin(void* dec) { ((ODecoder*)dec)->begin(); }
void ora_decode_step(void* dec, const float* emis, int T, int N) {
  ((ODecoder*)dec)->step(emis, T, N);
}
void ora_decode_end(void* dec) { ((ODecoder*)dec)->end(); }
void ora_prune(void* dec, int lookBack) { ((ODecoder*)dec)->prune(lookBack); }
long ora_tie_events(void* dec) { return ((ODecoder*)dec)->tieEvents; }
int ora_n_hypothesis(void* dec) {
  auto* d = (ODecoder*)dec;
  return (int)d->hyp[d->nDecoded - d->nPruned].size();
}
int ora_n_frames_in_buffer(void* dec) {
  auto* d = (ODecoder*)dec;
  return d->nDecoded - d->nPruned + 1;
}
int ora_best(void* dec, int lookBack, int maxLen, double* scores3, int* tokens, int* words) {
  auto* d = (ODecoder*)dec;
  int finalFrame = d->nDecoded - d->nPruned;
  if (d->lexicon && finalFrame - lookBack < 1) return 0; // LexiconDecoder.cpp:285-288
  int lb = lookBack;
  auto anc = d->bestAncestor(finalFrame, lb);
  if (anc.second < 0) return 0;
  // findBestAncestor updates lookBack through its int& parameter, and the result is sized
  // with the UPDATED value (LexiconDecoder.cpp:290-292, Utils.h:268-310)
  int ff = finalFrame - lb;
  return fillOne(*d, anc.first, anc.second, ff, maxLen, scores3, tokens, words);
}
int ora_all_final(void* dec, int maxHyp, int maxLen, double* scores3, int* tokens, int* words,
                  int* len) {
  return fillAll(*(ODecoder*)dec, maxHyp, maxLen, scores3, tokens, words, len);
}

double ora_bench_mt(int lexicon, const ora_options* opt, void* trie, void* lm, int sil,
                    int blank, int unk, int isLmToken, const float* emis, int B, int T, int N,
                    int nThreads, int warmup) {
  std::vector<std::unique_ptr<ODecoder>> decs;
  for (int i = 0; i < nThreads; ++i)
    decs.emplace_back(makeDecoder(lexicon, opt, trie, lm, sil, blank, unk, nullptr, 0, isLmToken));
  auto run = [&](int tid, int count) {
    for (int b = tid; b < count; b += nThreads) {
      ODecoder* d = decs[tid].get();
      d->begin();
      d->step(emis + (size_t)b * T * N, T, N);
      d->end();
    }
  };
  if (warmup > 0) {
    std::vector<std::thread> th;
    for (int i = 0; i < nThreads; ++i) th.emplace_back(run, i, std::min(B, nThreads * warmup));
    for (auto& t : th) t.join();
  }
  auto t0 = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> th;
    for (int i = 0; i < nThreads; ++i) th.emplace_back(run, i, B);
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
