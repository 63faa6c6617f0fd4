// flashlight_text.h — C++17 mirror of flashlight/text's decoder-path classes over the C-ABI of the
// B200 decode path (include/flt_decoder.h). Header-only; link with text_b200/lib/libflt_decoder.so.
//
// Same namespace, class names, constructor signatures, method names and exception types as the
// reference, so that code written against
//   flashlight/lib/text/decoder/Decoder.h:16-74            CriterionType, Decoder
//   flashlight/lib/text/decoder/Utils.h:30-39              DecodeResult
//   flashlight/lib/text/decoder/LexiconDecoder.h:21-31,115-157     LexiconDecoderOptions, LexiconDecoder
//   flashlight/lib/text/decoder/LexiconFreeDecoder.h:20-28,100-139 LexiconFreeDecoderOptions, LexiconFreeDecoder
//   flashlight/lib/text/decoder/Trie.h:21-92               SmearingMode, TrieNode, Trie
//   flashlight/lib/text/decoder/lm/LM.h:21-85, lm/ZeroLM.h, lm/KenLM.h:52-67
//   flashlight/lib/text/dictionary/*                       Dictionary, loadWords, ... (flashlight_dictionary.h)
// compiles unchanged for the decode path. All decoding runs on the GPU; there is no CPU decoder
// behind these classes (constructing a decoder without a CUDA device throws std::runtime_error).
//
// Additive: Decoder::decodeBatch (B utterances in one call) and LexiconDecoder/LexiconFreeDecoder
// ::setNbest. Differences, all loud:
//   * LM objects other than ZeroLM / KenLM (ARPA file) cannot be used by the decoders
//     (std::invalid_argument): a user-defined LM::score cannot run on the device.
//   * The Trie keeps the reference's node tree on the host as well (TrieNode::children can be walked);
//     the smeared scores are taken over from the library after smear().
#pragma once
#include <cmath>
#include <cstdlib>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../../include/flt_decoder.h"
#include "flashlight_dictionary.h"

namespace fl {
namespace lib {
namespace text {

namespace detail {
inline void check(int code) {
  if (code == FLT_OK) return;
  const std::string msg = flt_last_error();
  switch (code) {
    case FLT_ERR_INVALID: throw std::invalid_argument(msg);
    case FLT_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
    default: throw std::runtime_error(msg);
  }
}
inline int deviceOrdinal() {
  const char* e = std::getenv("FLT_DEVICE");
  return e ? std::atoi(e) : 0;
}
} // namespace detail

/* ------------------------------------------------------------------ Decoder.h / Utils.h ------ */
enum class CriterionType { ASG = 0, CTC = 1, S2S = 2 };

struct DecodeResult {
  double score;
  double emittingModelScore;
  double lmScore;
  std::vector<int> words;
  std::vector<int> tokens;
  explicit DecodeResult(int length = 0)
      : score(0), emittingModelScore(0), lmScore(0), words(length, -1), tokens(length, -1) {}
};

/* ------------------------------------------------------------------ Trie.h ------------------ */
constexpr int kTrieMaxLabel = 6;
enum class SmearingMode { NONE = 0, MAX = 1, LOGADD = 2 };

struct TrieNode {
  explicit TrieNode(int idx) : idx(idx), maxScore(0) {
    labels.reserve(kTrieMaxLabel);
    scores.reserve(kTrieMaxLabel);
  }
  std::unordered_map<int, std::shared_ptr<TrieNode>> children; // letter index -> child, as in Trie.h:45
  int idx;
  std::vector<int> labels;
  std::vector<float> scores;
  float maxScore;
  int id_ = 0; // creation order (= node index inside the library): not part of the reference's struct
};
using TrieNodePtr = std::shared_ptr<TrieNode>;

// The node tree lives twice: here, with the reference's layout, for code that walks TrieNode::children
// (decoder/Trie.h:39-54), and flattened inside the library for the kernels. Both are built by the same
// sequence of inserts, so a node's creation index is its index in the library.
class Trie {
 public:
  Trie(int maxChildren, int rootIdx) : maxChildren_(maxChildren), root_(std::make_shared<TrieNode>(rootIdx)) {
    detail::check(flt_trie_create(maxChildren, rootIdx, &h_));
    byId_.push_back(root_.get());
  }
  ~Trie() { flt_trie_destroy(h_); }
  Trie(const Trie&) = delete;
  Trie& operator=(const Trie&) = delete;

  const TrieNode* getRoot() const { return root_.get(); }

  // Trie.cpp:26-48 (throws std::out_of_range on a token index outside [0, maxChildren); the nodes
  // created before the bad index stay, here as in the library)
  TrieNodePtr insert(const std::vector<int>& indices, int label, float score) {
    TrieNodePtr node = root_;
    bool bad = false;
    for (int idx : indices) {
      if (idx < 0 || idx >= maxChildren_) {
        bad = true;
        break;
      }
      auto it = node->children.find(idx);
      if (it == node->children.end()) {
        auto child = std::make_shared<TrieNode>(idx);
        child->id_ = (int)byId_.size();
        byId_.push_back(child.get());
        node->children[idx] = child;
        node = child;
      } else {
        node = it->second;
      }
    }
    detail::check(flt_trie_insert(h_, indices.data(), (int)indices.size(), label, score)); // throws if bad
    (void)bad;
    if (node->labels.size() < (size_t)kTrieMaxLabel) { // Trie.cpp:40-46 (the library prints the notice)
      node->labels.push_back(label);
      node->scores.push_back(score);
    }
    return node;
  }
  // Trie.cpp:50-64 (nullptr when the path does not exist)
  TrieNodePtr search(const std::vector<int>& indices) {
    TrieNodePtr node = root_;
    for (int idx : indices) {
      if (idx < 0 || idx >= maxChildren_)
        throw std::out_of_range("[Trie] Invalid letter index: " + std::to_string(idx));
      auto it = node->children.find(idx);
      if (it == node->children.end()) return nullptr;
      node = it->second;
    }
    return node;
  }
  // Trie.cpp:79-101: computed by the library (float arithmetic of the reference), mirrored here
  void smear(const SmearingMode smearMode) {
    detail::check(flt_trie_smear(h_, (int)smearMode));
    if (smearMode == SmearingMode::NONE) return;
    std::vector<float> ms(byId_.size());
    detail::check(flt_trie_max_scores(h_, ms.data(), (int64_t)ms.size()));
    for (size_t i = 0; i < byId_.size(); ++i) byId_[i]->maxScore = ms[i];
  }

  flt_trie* handle() const { return h_; }

  // Additive: the built (inserted + smeared) Trie as a table file (csrc/table_io.h) — written once, loaded by
  // every later process instead of repeating one insert per lexicon word.
  void save(const std::string& path) const { detail::check(flt_trie_save(h_, path.c_str())); }
  static std::shared_ptr<Trie> load(const std::string& path) {
    flt_trie* h = nullptr;
    detail::check(flt_trie_load(path.c_str(), &h));
    std::shared_ptr<Trie> t(new Trie(h));
    return t;
  }

 private:
  // takes over a library Trie and rebuilds the host node tree from its CSR export
  explicit Trie(flt_trie* h) : maxChildren_(0), h_(h) {
    int32_t meta[5];
    try {
      detail::check(flt_trie_export(h_, meta, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
      const size_t nn = (size_t)meta[2], ne = (size_t)meta[3], nl = (size_t)meta[4];
      std::vector<int32_t> childOff(nn + 1), childTok(ne + 1), childNode(ne + 1), labelOff(nn + 1), labels(nl + 1);
      std::vector<float> scores(nl + 1), maxScore(nn);
      detail::check(flt_trie_export(h_, meta, childOff.data(), childTok.data(), childNode.data(), labelOff.data(),
                                    labels.data(), scores.data(), maxScore.data()));
      maxChildren_ = meta[0];
      std::vector<TrieNodePtr> nodes(nn);
      nodes[0] = std::make_shared<TrieNode>(meta[1]);
      byId_.assign(nn, nullptr);
      // a child is always created after its parent, so parents come first in creation order
      for (size_t i = 0; i < nn; ++i) {
        TrieNode* nd = nodes[i].get();
        nd->id_ = (int)i;
        nd->maxScore = maxScore[i];
        nd->labels.assign(labels.begin() + labelOff[i], labels.begin() + labelOff[i + 1]);
        nd->scores.assign(scores.begin() + labelOff[i], scores.begin() + labelOff[i + 1]);
        byId_[i] = nd;
        for (int e = childOff[i]; e < childOff[i + 1]; ++e) {
          auto child = std::make_shared<TrieNode>(childTok[e]);
          nodes[(size_t)childNode[e]] = child;
          nd->children[childTok[e]] = child;
        }
      }
      root_ = nodes[0];
    } catch (...) {
      flt_trie_destroy(h_);
      throw;
    }
  }
  int maxChildren_;
  TrieNodePtr root_;
  std::vector<TrieNode*> byId_;
  flt_trie* h_ = nullptr;
};
using TriePtr = std::shared_ptr<Trie>;

/* ------------------------------------------------------------------ lm/LM.h ------------------ */
struct LMState {
  std::unordered_map<int, std::shared_ptr<LMState>> children;
  std::vector<int> history; // labels from the start state (this mirror's host-side scoring)

  template <typename T>
  std::shared_ptr<T> child(int usrIdx) {
    auto s = children.find(usrIdx);
    if (s == children.end()) {
      auto state = std::make_shared<T>();
      state->history = history;
      state->history.push_back(usrIdx);
      children[usrIdx] = state;
      return state;
    }
    return std::static_pointer_cast<T>(s->second);
  }
  // lm/LM.h:37-49: pointer order, throws on null
  int compare(const std::shared_ptr<LMState>& state) const {
    LMState* inState = state.get();
    if (!inState) throw std::runtime_error("a state is null");
    if (this == inState) return 0;
    return this < inState ? -1 : 1;
  }
};
using LMStatePtr = std::shared_ptr<LMState>;

class LM {
 public:
  virtual ~LM() = default;
  virtual LMStatePtr start(bool startWithNothing) = 0;
  virtual std::pair<LMStatePtr, float> score(const LMStatePtr& state, const int usrTokenIdx) = 0;
  virtual std::pair<LMStatePtr, float> finish(const LMStatePtr& state) = 0;
  virtual void updateCache(std::vector<LMStatePtr> /*stateIdices*/) {}
  // device-resident model behind this LM, or nullptr (user-defined LMs cannot decode on the device)
  virtual const flt_lm* handle() const { return nullptr; }

 protected:
  std::vector<int> usrToLmIdxMap_;
};
using LMPtr = std::shared_ptr<LM>;

// lm/ZeroLM.cpp:14-26
class ZeroLM : public LM {
 public:
  ZeroLM() { detail::check(flt_lm_zero_create(&h_)); }
  ~ZeroLM() override { flt_lm_destroy(h_); }
  LMStatePtr start(bool /*startWithNothing*/) override { return std::make_shared<LMState>(); }
  std::pair<LMStatePtr, float> score(const LMStatePtr& state, const int usrTokenIdx) override {
    return std::make_pair(state->child<LMState>(usrTokenIdx), 0.0f);
  }
  std::pair<LMStatePtr, float> finish(const LMStatePtr& state) override { return std::make_pair(state, 0.0f); }
  const flt_lm* handle() const override { return h_; }

 private:
  flt_lm* h_ = nullptr;
};
using ZeroLMPtr = std::shared_ptr<ZeroLM>;

// lm/KenLM.cpp:32-83 for ARPA files (or a table file written by save()): log10 scores, OOV -> <unk>, start = <s> context, finish = </s>
struct KenLMState : LMState {};
class KenLM : public LM {
 public:
  KenLM(const std::string& path, const Dictionary& usrTknDict) {
    const int n = (int)usrTknDict.indexSize();
    std::vector<std::string> words(n);
    std::vector<const char*> ptrs(n);
    for (int i = 0; i < n; ++i) {
      words[i] = usrTknDict.getEntry(i);
      ptrs[i] = words[i].c_str();
    }
    detail::check(flt_lm_ngram_load_arpa(path.c_str(), ptrs.data(), n, &h_));
  }
  ~KenLM() override { flt_lm_destroy(h_); }
  LMStatePtr start(bool startWithNothing) override {
    if (startWithNothing) throw std::runtime_error("[KenLM] start(true) (null context) is not supported by the device model");
    return std::make_shared<KenLMState>();
  }
  std::pair<LMStatePtr, float> score(const LMStatePtr& state, const int usrTokenIdx) override {
    auto out = state->child<KenLMState>(usrTokenIdx);
    std::vector<float> s(out->history.size());
    detail::check(flt_lm_score_seq(h_, out->history.data(), (int)out->history.size(), 0, s.data()));
    return std::make_pair(std::static_pointer_cast<LMState>(out), s.back());
  }
  std::pair<LMStatePtr, float> finish(const LMStatePtr& state) override {
    auto out = state->child<KenLMState>(-1);
    std::vector<float> s(state->history.size() + 1);
    detail::check(flt_lm_score_seq(h_, state->history.data(), (int)state->history.size(), 1, s.data()));
    return std::make_pair(std::static_pointer_cast<LMState>(out), s.back());
  }
  const flt_lm* handle() const override { return h_; }
  // Additive: write the hashed n-gram tables + vocabulary as a table file (csrc/table_io.h); the constructor
  // above loads such a file in place of ARPA text, the way the reference's takes a KenLM binary.
  void save(const std::string& path) const { detail::check(flt_lm_save(h_, path.c_str())); }

 private:
  flt_lm* h_ = nullptr;
};
using KenLMPtr = std::shared_ptr<KenLM>;

/* ------------------------------------------------------------------ options ------------------ */
struct LexiconDecoderOptions {
  int beamSize;
  int beamSizeToken;
  double beamThreshold;
  double lmWeight;
  double wordScore;
  double unkScore;
  double silScore;
  bool logAdd;
  CriterionType criterionType;
};
struct LexiconFreeDecoderOptions {
  int beamSize;
  int beamSizeToken;
  double beamThreshold;
  double lmWeight;
  double silScore;
  bool logAdd;
  CriterionType criterionType;
};

/* ------------------------------------------------------------------ Decoder ------------------ */
class Decoder {
 public:
  Decoder() = default;
  virtual ~Decoder() { flt_decoder_destroy(h_); }
  Decoder(const Decoder&) = delete;
  Decoder& operator=(const Decoder&) = delete;

  // Online decoding (Decoder.h:18-35): the beam stays on the device between chunks.
  virtual void decodeBegin() {
    online_ = true;
    begun_ = false; // flt_stream_begin needs N: issued by the first decodeStep
    final_.clear();
  }
  virtual void decodeStep(const float* emissions, int T, int N) {
    if (!online_) decodeBegin();
    if (!begun_) {
      detail::check(flt_stream_begin(h_, N));
      begun_ = true;
      onlineN_ = N;
    }
    detail::check(flt_stream_step(h_, emissions, T, N));
  }
  virtual void decodeEnd() {
    if (!online_) decodeBegin();
    if (!begun_) {
      detail::check(flt_stream_begin(h_, onlineN_ > 0 ? onlineN_ : 1));
      begun_ = true;
    }
    detail::check(flt_stream_end(h_));
  }
  // Decoder.h:51-57
  virtual std::vector<DecodeResult> decode(const float* emissions, int T, int N) {
    auto all = decodeBatch(emissions, 1, T, N);
    online_ = false;
    final_ = all[0];
    offlineFrames_ = T + 1;
    return std::move(all[0]);
  }
  virtual void prune(int lookBack = 0) {
    if (online_ && begun_) detail::check(flt_stream_prune(h_, lookBack));
  }
  virtual int nDecodedFramesInBuffer() const {
    if (!(online_ && begun_)) return online_ ? 1 : offlineFrames_ + 1;
    int n = 0;
    detail::check(flt_stream_frames_in_buffer(h_, &n));
    return n;
  }
  virtual DecodeResult getBestHypothesis(int lookBack = 0) const {
    if (!(online_ && begun_)) {
      if (online_ || lookBack != 0) return DecodeResult();
      return final_.empty() ? DecodeResult() : final_[0];
    }
    int frames = 0;
    detail::check(flt_stream_frames_in_buffer(h_, &frames));
    DecodeResult d(frames);
    double sc[3] = {0, 0, 0};
    int len = 0;
    detail::check(flt_stream_best(h_, lookBack, frames, d.tokens.data(), d.words.data(), sc, &len));
    if (len == 0) return DecodeResult();
    d.tokens.resize(len);
    d.words.resize(len);
    d.score = sc[0];
    d.emittingModelScore = sc[1];
    d.lmScore = sc[2];
    return d;
  }
  virtual std::vector<DecodeResult> getAllFinalHypothesis() const {
    if (!(online_ && begun_)) return final_;
    int frames = 0, count = 0;
    detail::check(flt_stream_frames_in_buffer(h_, &frames));
    const int K = beamSize();
    std::vector<int32_t> tok((size_t)K * frames), wrd((size_t)K * frames), lens(K);
    std::vector<double> sc((size_t)K * 3);
    detail::check(flt_stream_all_final(h_, K, frames, tok.data(), wrd.data(), sc.data(), lens.data(), &count));
    std::vector<DecodeResult> out;
    for (int r = 0; r < count && r < K; ++r) {
      DecodeResult d(lens[r]);
      d.score = sc[3 * r];
      d.emittingModelScore = sc[3 * r + 1];
      d.lmScore = sc[3 * r + 2];
      for (int i = 0; i < lens[r]; ++i) {
        d.tokens[i] = tok[(size_t)r * frames + i];
        d.words[i] = wrd[(size_t)r * frames + i];
      }
      out.push_back(std::move(d));
    }
    return out;
  }
  int nHypothesis() const {
    if (!(online_ && begun_)) return (int)final_.size();
    int n = 0;
    detail::check(flt_stream_n_hypothesis(h_, &n));
    return n;
  }

  // B utterances at once. emissions: row-major [B,T,N] fp32, host or device memory; lengths
  // (host, optional): valid frames per utterance; nbest <= beamSize hypotheses are materialised
  // per utterance (-1: all). result[b] is sorted by score, best first (Utils.h:252-266).
  std::vector<std::vector<DecodeResult>> decodeBatch(const float* emissions, int B, int T, int N,
                                                      const int* lengths = nullptr, int nbest = -1) {
    const int K = beamSize();
    const int nb = nbest < 0 ? K : (nbest < K ? nbest : K);
    detail::check(flt_decoder_set_nbest(h_, nb > 0 ? nb : 1));
    detail::check(flt_decode_batch(h_, emissions, B, T, N, lengths));
    const size_t L = (size_t)T + 2;
    std::vector<int32_t> tok((size_t)B * nb * L), wrd((size_t)B * nb * L), cnt(B);
    std::vector<double> sc((size_t)B * nb * 3);
    if (B > 0 && nb > 0) detail::check(flt_nbest_copy(h_, nb, tok.data(), wrd.data(), sc.data(), cnt.data()));
    std::vector<std::vector<DecodeResult>> out(B);
    for (int b = 0; b < B; ++b) {
      const int n = cnt[b] < nb ? cnt[b] : nb;
      const int len = (lengths ? lengths[b] : T) + 2;
      out[b].reserve(n);
      for (int r = 0; r < n; ++r) {
        DecodeResult d(len);
        const size_t o = ((size_t)b * nb + r);
        d.score = sc[o * 3 + 0];
        d.emittingModelScore = sc[o * 3 + 1];
        d.lmScore = sc[o * 3 + 2];
        for (int i = 0; i < len; ++i) {
          d.tokens[i] = tok[o * L + i];
          d.words[i] = wrd[o * L + i];
        }
        out[b].push_back(std::move(d));
      }
    }
    return out;
  }
  flt_decoder* handle() const { return h_; }

 protected:
  virtual int beamSize() const = 0;
  flt_decoder* h_ = nullptr;
  std::vector<DecodeResult> final_; // result of the last offline decode()
  bool online_ = false, begun_ = false;
  int onlineN_ = 0, offlineFrames_ = 0;
};

namespace detail {
inline const flt_lm* deviceLM(const LMPtr& lm) {
  if (!lm) throw std::invalid_argument("null LM");
  const flt_lm* h = lm->handle();
  if (!h)
    throw std::invalid_argument(
        "this LM has no device model: only ZeroLM and KenLM (ARPA) can decode on the GPU; "
        "user-defined LM::score cannot run there");
  return h;
}
} // namespace detail

// LexiconDecoder.h:115-157
class LexiconDecoder : public Decoder {
 public:
  LexiconDecoder(LexiconDecoderOptions opt, const TriePtr& lexicon, const LMPtr& lm, const int sil,
                 const int blank, const int unk, const std::vector<float>& transitions, const bool isLmToken)
      : opt_(std::move(opt)), lexicon_(lexicon), lm_(lm), sil_(sil), blank_(blank), unk_(unk),
        transitions_(transitions), isLmToken_(isLmToken) {
    if (!lexicon) throw std::invalid_argument("null lexicon");
    flt_options o{opt_.beamSize, opt_.beamSizeToken, opt_.beamThreshold, opt_.lmWeight, opt_.wordScore,
                  opt_.unkScore, opt_.silScore, opt_.logAdd ? 1 : 0, (int)opt_.criterionType};
    detail::check(flt_decoder_create_lexicon(&o, lexicon->handle(), detail::deviceLM(lm), sil, blank, unk,
                                             transitions_.data(), (int64_t)transitions_.size(),
                                             isLmToken ? 1 : 0, detail::deviceOrdinal(), &h_));
  }
  const LexiconDecoderOptions& getOptions() const { return opt_; }

 protected:
  int beamSize() const override { return opt_.beamSize; }
  LexiconDecoderOptions opt_;
  TriePtr lexicon_;
  LMPtr lm_;
  int sil_, blank_, unk_;
  std::vector<float> transitions_;
  bool isLmToken_;
};

// LexiconFreeDecoder.h:100-139
class LexiconFreeDecoder : public Decoder {
 public:
  LexiconFreeDecoder(LexiconFreeDecoderOptions opt, const LMPtr& lm, const int sil, const int blank,
                     const std::vector<float>& transitions)
      : opt_(std::move(opt)), lm_(lm), sil_(sil), blank_(blank), transitions_(transitions) {
    flt_options o{opt_.beamSize, opt_.beamSizeToken, opt_.beamThreshold, opt_.lmWeight, 0.0,
                  -std::numeric_limits<double>::infinity(), opt_.silScore, opt_.logAdd ? 1 : 0,
                  (int)opt_.criterionType};
    detail::check(flt_decoder_create_lexfree(&o, detail::deviceLM(lm), sil, blank, transitions_.data(),
                                             (int64_t)transitions_.size(), detail::deviceOrdinal(), &h_));
  }
  const LMPtr& getLMPtr() const { return lm_; }
  int getSilIdx() const { return sil_; }
  int getBlankIdx() const { return blank_; }
  const LexiconFreeDecoderOptions& getOptions() const { return opt_; }
  const std::vector<float>& getTransitions() const { return transitions_; }

 protected:
  int beamSize() const override { return opt_.beamSize; }
  LexiconFreeDecoderOptions opt_;
  LMPtr lm_;
  int sil_, blank_;
  std::vector<float> transitions_;
};

// One-shot setup (test/decoder/DecoderTest.cpp:126-146): every spelling of every lexicon word goes into
// a Trie with the word's LM score (lm->score from the start state), then the Trie is smeared.
inline TriePtr buildTrie(const LexiconMap& lexicon, const Dictionary& tokenDict, const Dictionary& wordDict,
                         const LMPtr& lm, int silIdx, int maxReps = 0, SmearingMode smear = SmearingMode::MAX) {
  auto trie = std::make_shared<Trie>((int)tokenDict.indexSize(), silIdx);
  auto start = lm->start(false);
  for (const auto& kv : lexicon) {
    const int usrIdx = wordDict.getIndex(kv.first);
    float score = 0.0f;
    if (!kv.second.empty()) score = lm->score(start, usrIdx).second;
    for (const auto& spelling : kv.second) trie->insert(tkn2Idx(spelling, tokenDict, maxReps), usrIdx, score);
  }
  trie->smear(smear);
  return trie;
}

} // namespace text
} // namespace lib
} // namespace fl
