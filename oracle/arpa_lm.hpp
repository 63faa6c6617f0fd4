// ORACLE (test infrastructure, never shipped): ARPA back-off n-gram LM.
//
// Restates the published query algorithm of KenLM, the third-party dependency the
// reference's KenLM adapter calls (flashlight/lib/text/decoder/lm/KenLM.cpp:32-83:
// LoadVirtual / BeginSentenceWrite / BaseScore / EndSentence / Vocabulary::Index).
// KenLM is NOT vendored under /root/reference and is absent from this image
// (pins: jacobkahn/kenlm @ 5bf7b46558e1c5595bf3b8c9b0b1f9d8d257040a,
// /root/reference/cmake/BuildKenlm.cmake:6-7; kpu/kenlm master in setup.py:79).
//
// Algorithm restated (KenLM lm/model.cc GenericModel::FullScore, published):
//   * vocabulary ids in ARPA unigram order, "<unk>" forced to id 0, OOV -> 0
//   * values are log10 probabilities stored as float, used as-is (no ln conversion)
//   * p(w | ctx): find the longest n-gram (ctx_k..ctx_1, w) present in the model,
//     growing the context one word at a time and stopping at the first miss;
//     ret = prob(longest match) as float; then for i = matchLen-1 .. ctxLen-1 in
//     ascending order ret += backoff(ctx_{i+1}..ctx_1)   (float adds; a context
//     n-gram that is absent contributes 0). KenLM's state minimisation only drops
//     context words whose back-off is exactly zero, so it does not change any value.
//   * start(false) = context {<s>}; start(true) = empty; finish = score(</s>).
// Pinned against the reference's own known answers for this boundary
// (flashlight/lib/text/test/decoder/DecoderTest.cpp:107-120,148-155,184,190-194)
// in tests/test_oracle_golden.py.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace oracle {

constexpr int kMaxOrder = 6; // FL_TEXT_KENLM_MAX_ORDER (decoder/lm/CMakeLists.txt:3)

struct NgramKey {
  uint64_t lo, hi; // up to 6 ids x 21 bits, exact (no hashing)
  bool operator==(const NgramKey& o) const { return lo == o.lo && hi == o.hi; }
};
struct NgramKeyHash {
  size_t operator()(const NgramKey& k) const {
    uint64_t x = k.lo * 0x9E3779B97F4A7C15ull ^ (k.hi + 0xD6E8FEB86659FD93ull);
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 29;
    return (size_t)x;
  }
};
struct ProbBackoff {
  float prob;
  float backoff;
};

class ArpaLM {
 public:
  explicit ArpaLM(const std::string& path) { load(path); }

  int order() const { return order_; }
  int vocabSize() const { return (int)unigrams_.size(); }
  int bos() const { return bos_; }
  int eos() const { return eos_; }

  // Vocabulary::Index: OOV -> 0 (<unk>)
  int index(const std::string& w) const {
    auto it = vocab_.find(w);
    return it == vocab_.end() ? 0 : it->second;
  }

  // ctx: most recent word first; only the first min(n, order-1) are used.
  float score(const int* ctx, int nctx, int w) const {
    int L = nctx < order_ - 1 ? nctx : order_ - 1;
    float ret = unigrams_[w].prob;
    int matchLen = 1;
    int words[kMaxOrder];
    for (int k = 1; k <= L; ++k) {
      // n-gram = ctx[k-1], ..., ctx[0], w  (natural order)
      for (int i = 0; i < k; ++i) words[i] = ctx[k - 1 - i];
      words[k] = w;
      const ProbBackoff* pb = find(words, k + 1);
      if (!pb) break;
      ret = pb->prob;
      matchLen = k + 1;
    }
    for (int i = matchLen - 1; i < L; ++i) {
      // back-off of the context n-gram made of the i+1 most recent words
      for (int j = 0; j <= i; ++j) words[j] = ctx[i - j];
      const ProbBackoff* pb = find(words, i + 1);
      if (pb) ret += pb->backoff;
    }
    return ret;
  }

  const ProbBackoff* find(const int* words, int n) const {
    if (n == 1) return &unigrams_[words[0]];
    auto it = ngrams_[n].find(pack(words, n));
    return it == ngrams_[n].end() ? nullptr : &it->second;
  }

  // Iteration support (used to flatten tables for the device in tests).
  const std::vector<ProbBackoff>& unigrams() const { return unigrams_; }
  const std::unordered_map<NgramKey, ProbBackoff, NgramKeyHash>& ngrams(int n) const {
    return ngrams_[n];
  }
  static void unpack(const NgramKey& k, int n, int* words) {
    for (int i = 0; i < n; ++i) {
      uint64_t v = i < 3 ? (k.lo >> (21 * i)) : (k.hi >> (21 * (i - 3)));
      words[i] = (int)(v & 0x1FFFFF);
    }
  }

 private:
  static NgramKey pack(const int* words, int n) {
    NgramKey k{0, 0};
    for (int i = 0; i < n; ++i) {
      uint64_t v = (uint64_t)(uint32_t)words[i] & 0x1FFFFF;
      if (i < 3) k.lo |= v << (21 * i);
      else k.hi |= v << (21 * (i - 3));
    }
    k.hi |= (uint64_t)n << 60; // length tag
    return k;
  }

  void load(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("[ArpaLM] cannot open " + path);
    std::string line;
    std::vector<long> counts(kMaxOrder + 2, 0);
    bool inData = false;
    int section = 0;
    ngrams_.resize(kMaxOrder + 2);
    std::vector<std::string> toks;
    while (std::getline(in, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty()) continue;
      if (line == "\\data\\") { inData = true; continue; }
      if (line == "\\end\\") break;
      if (line[0] == '\\') {
        // "\N-grams:"
        section = atoi(line.c_str() + 1);
        if (section < 1 || section > kMaxOrder)
          throw std::runtime_error("[ArpaLM] unsupported order in " + line);
        if (section == 1) {
          // <unk> is id 0 whether or not it is listed first
          vocab_["<unk>"] = 0;
          unigrams_.push_back(ProbBackoff{-100.0f, 0.0f});
        } else {
          ngrams_[section].reserve((size_t)(counts[section] * 1.3) + 16);
        }
        inData = false;
        continue;
      }
      if (inData) {
        if (line.compare(0, 6, "ngram ") == 0) {
          int n = atoi(line.c_str() + 6);
          size_t eq = line.find('=');
          if (n >= 1 && n <= kMaxOrder && eq != std::string::npos) {
            counts[n] = atol(line.c_str() + eq + 1);
            if (n > order_) order_ = n;
          }
        }
        continue;
      }
      if (section == 0) continue;
      // "<prob>\t<w1> ... <wn>[\t<backoff>]"
      toks.clear();
      size_t p = 0;
      while (p < line.size()) {
        size_t q = line.find_first_of(" \t", p);
        if (q == std::string::npos) q = line.size();
        if (q > p) toks.emplace_back(line.substr(p, q - p));
        p = q + 1;
      }
      if ((int)toks.size() < section + 1) continue;
      ProbBackoff pb;
      pb.prob = strtof(toks[0].c_str(), nullptr);
      pb.backoff = (int)toks.size() > section + 1 ? strtof(toks[section + 1].c_str(), nullptr) : 0.0f;
      if (section == 1) {
        const std::string& w = toks[1];
        if (w == "<unk>") {
          unigrams_[0] = pb;
        } else {
          int id = (int)unigrams_.size();
          vocab_[w] = id;
          unigrams_.push_back(pb);
        }
      } else {
        int words[kMaxOrder];
        for (int i = 0; i < section; ++i) words[i] = index(toks[1 + i]);
        ngrams_[section][pack(words, section)] = pb;
      }
    }
    if (unigrams_.size() >= (1u << 21)) throw std::runtime_error("[ArpaLM] vocabulary too large for exact keys");
    auto b = vocab_.find("<s>");
    auto e = vocab_.find("</s>");
    if (b == vocab_.end() || e == vocab_.end()) throw std::runtime_error("[ArpaLM] missing <s> or </s>");
    bos_ = b->second;
    eos_ = e->second;
  }

  int order_ = 0;
  int bos_ = -1, eos_ = -1;
  std::unordered_map<std::string, int> vocab_;
  std::vector<ProbBackoff> unigrams_;
  std::vector<std::unordered_map<NgramKey, ProbBackoff, NgramKeyHash>> ngrams_;
};

} // namespace oracle
