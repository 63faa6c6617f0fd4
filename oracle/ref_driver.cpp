// ORACLE (test infrastructure, never shipped): C driver around the UNMODIFIED reference.
//
// Compiled by oracle/Makefile together with the reference's own sources where they lie
// under /root/reference (LexiconDecoder.cpp, LexiconFreeDecoder.cpp, Trie.cpp, lm/ZeroLM.cpp)
// into oracle/_ref/libflref.so. Nothing from the reference is copied into this repository;
// this file only calls its public C++ API (decoder/Decoder.h:36-74, LexiconDecoder.h:117-133,
// LexiconFreeDecoder.h:102-112, Trie.h:66-86, lm/LM.h:21-85).
//
// KenLM is not available (see oracle/arpa_lm.hpp), so the n-gram LM plugged into the
// reference decoder is `RefArpaLM`: an fl::lib::text::LM subclass with the same state
// discipline as the reference's KenLM adapter (child<State>(usrIdx) for identity,
// lm/KenLM.cpp:63-83) over oracle::ArpaLM for the arithmetic.

#include <atomic>
#include <chrono>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "flashlight/lib/text/decoder/LexiconDecoder.h"
#include "flashlight/lib/text/decoder/LexiconFreeDecoder.h"
#include "flashlight/lib/text/decoder/Trie.h"
#include "flashlight/lib/text/decoder/lm/ZeroLM.h"

#include "arpa_lm.hpp"
#include "oracle_api.h"

using namespace fl::lib::text;

namespace {

thread_local std::string gErr;

struct ArpaState : LMState {
  int ctx[oracle::kMaxOrder];
  int n = 0;
};

class RefArpaLM : public LM {
 public:
  RefArpaLM(const std::string& path, const char* const* words, int nWords) : model_(path) {
    usrToLmIdxMap_.resize(nWords);
    for (int i = 0; i < nWords; ++i) usrToLmIdxMap_[i] = model_.index(words[i]);
  }
  LMStatePtr start(bool startWithNothing) override {
    auto s = std::make_shared<ArpaState>();
    if (!startWithNothing) {
      s->ctx[0] = model_.bos();
      s->n = 1;
    }
    return s;
  }
  std::pair<LMStatePtr, float> score(const LMStatePtr& state, const int usrTokenIdx) override {
    if (usrTokenIdx < 0 || usrTokenIdx >= (int)usrToLmIdxMap_.size()) {
      throw std::runtime_error("[KenLM] Invalid user token index: " + std::to_string(usrTokenIdx));
    }
    auto in = std::static_pointer_cast<ArpaState>(state);
    auto out = in->child<ArpaState>(usrTokenIdx);
    float s = advance(*in, usrToLmIdxMap_[usrTokenIdx], *out);
    return {std::move(out), s};
  }
  std::pair<LMStatePtr, float> finish(const LMStatePtr& state) override {
    auto in = std::static_pointer_cast<ArpaState>(state);
    auto out = in->child<ArpaState>(-1);
    float s = advance(*in, model_.eos(), *out);
    return {std::move(out), s};
  }

 private:
  float advance(const ArpaState& in, int w, ArpaState& out) const {
    float s = model_.score(in.ctx, in.n, w);
    int keep = model_.order() - 1;
    out.ctx[0] = w;
    int n = 1;
    for (int i = 0; i < in.n && n < keep; ++i) out.ctx[n++] = in.ctx[i];
    out.n = keep > 0 ? n : 0;
    return s;
  }
  oracle::ArpaLM model_;
};

struct TrieBox { TriePtr t; };
struct LMBox { LMPtr lm; };
struct DecBox {
  std::unique_ptr<Decoder> d;
  bool lexicon = false;
};

LexiconDecoderOptions lexOpt(const ora_options* o) {
  return LexiconDecoderOptions{
      o->beamSize, o->beamSizeToken, o->beamThreshold, o->lmWeight, o->wordScore,
      o->unkScore, o->silScore, o->logAdd != 0, (CriterionType)o->criterion};
}
LexiconFreeDecoderOptions freeOpt(const ora_options* o) {
  return LexiconFreeDecoderOptions{
      o->beamSize, o->beamSizeToken, o->beamThreshold, o->lmWeight,
      o->silScore, o->logAdd != 0, (CriterionType)o->criterion};
}

int fill(const std::vector<DecodeResult>& res, int maxHyp, int stride, double* scores3,
         int* tokens, int* words, int* len) {
  int n = (int)res.size();
  for (int i = 0; i < n && i < maxHyp; ++i) {
    scores3[3 * i + 0] = res[i].score;
    scores3[3 * i + 1] = res[i].emittingModelScore;
    scores3[3 * i + 2] = res[i].lmScore;
    int L = (int)res[i].tokens.size();
    if (len) len[i] = L;
    for (int j = 0; j < L && j < stride; ++j) {
      tokens[(size_t)i * stride + j] = res[i].tokens[j];
      words[(size_t)i * stride + j] = res[i].words[j];
    }
  }
  return n;
}

std::unique_ptr<Decoder> makeDecoder(int lexicon, const ora_options* opt, void* trie, void* lm,
                                     int sil, int blank, int unk, const std::vector<float>& tr,
                                     int isLmToken) {
  if (lexicon) {
    return std::make_unique<LexiconDecoder>(
        lexOpt(opt), ((TrieBox*)trie)->t, ((LMBox*)lm)->lm, sil, blank, unk, tr, isLmToken != 0);
  }
  return std::make_unique<LexiconFreeDecoder>(freeOpt(opt), ((LMBox*)lm)->lm, sil, blank, tr);
}

} // namespace

extern "C" {

const char* ref_last_error(void) { return gErr.c_str(); }

void* ref_trie_create(int maxChildren, int rootIdx) {
  return new TrieBox{std::make_shared<Trie>(maxChildren, rootIdx)};
}
int ref_trie_insert(void* trie, const int* idx, int n, int label, float score) {
  try {
    ((TrieBox*)trie)->t->insert(std::vector<int>(idx, idx + n), label, score);
    return 0;
  } catch (const std::exception& e) {
    gErr = e.what();
    return -1;
  }
}
void ref_trie_smear(void* trie, int mode) { ((TrieBox*)trie)->t->smear((SmearingMode)mode); }
int ref_trie_search(void* trie, const int* idx, int n, float* maxScore, int* nLabels,
                    int* labels6, float* scores6) {
  try {
    auto node = ((TrieBox*)trie)->t->search(std::vector<int>(idx, idx + n));
    if (!node) return 0;
    if (maxScore) *maxScore = node->maxScore;
    if (nLabels) *nLabels = (int)node->labels.size();
    for (size_t i = 0; i < node->labels.size() && i < 6; ++i) {
      if (labels6) labels6[i] = node->labels[i];
      if (scores6) scores6[i] = node->scores[i];
    }
    return 1;
  } catch (const std::exception& e) {
    gErr = e.what();
    return -1;
  }
}
void ref_trie_destroy(void* trie) { delete (TrieBox*)trie; }

void* ref_lm_zero(void) { return new LMBox{std::make_shared<ZeroLM>()}; }
void* ref_lm_arpa(const char* path, const char* const* words, int nWords) {
  try {
    return new LMBox{std::make_shared<RefArpaLM>(path, words, nWords)};
  } catch (const std::exception& e) {
    gErr = e.what();
    return nullptr;
  }
}
int ref_lm_score_seq(void* lm, const int* usrIdx, int n, int withFinish, float* out) {
  try {
    auto& m = ((LMBox*)lm)->lm;
    auto st = m->start(false);
    for (int i = 0; i < n; ++i) {
      auto r = m->score(st, usrIdx[i]);
      st = r.first;
      out[i] = r.second;
    }
    if (withFinish) out[n] = m->finish(st).second;
    return 0;
  } catch (const std::exception& e) {
    gErr = e.what();
    return -1;
  }
}
void ref_lm_destroy(void* lm) { delete (LMBox*)lm; }

void* ref_decoder_lexfree(const ora_options* opt, void* lm, int sil, int blank,
                          const float* trans, int nTrans) {
  auto* b = new DecBox;
  std::vector<float> tr(trans, trans + (trans ? nTrans : 0));
  b->d = makeDecoder(0, opt, nullptr, lm, sil, blank, -1, tr, 0);
  return b;
}
void* ref_decoder_lexicon(const ora_options* opt, void* trie, void* lm, int sil, int blank,
                          int unk, const float* trans, int nTrans, int isLmToken) {
  auto* b = new DecBox;
  std::vector<float> tr(trans, trans + (trans ? nTrans : 0));
  b->d = makeDecoder(1, opt, trie, lm, sil, blank, unk, tr, isLmToken);
  b->lexicon = true;
  return b;
}
void ref_decoder_destroy(void* dec) { delete (DecBox*)dec; }

int ref_decode(void* dec, const float* emis, int T, int N, int maxHyp, double* scores3,
               int* tokens, int* words) {
  try {
    auto res = ((DecBox*)dec)->d->decode(emis, T, N);
    return fill(res, maxHyp, T + 2, scores3, tokens, words, nullptr);
  } catch (const std::exception& e) {
    gErr = e.what();
    return -1;
  }
}
void ref_decode_begin(void* dec) { ((DecBox*)dec)->d->decodeBegin(); }
void ref_decode_step(void* dec, const float* emis, int T, int N) {
  ((DecBox*)dec)->d->decodeStep(emis, T, N);
}
void ref_decode_end(void* dec) { ((DecBox*)dec)->d->decodeEnd(); }
void ref_prune(void* dec, int lookBack) { ((DecBox*)dec)->d->prune(lookBack); }
int ref_n_hypothesis(void* dec) {
  auto* b = (DecBox*)dec;
  if (b->lexicon) return static_cast<LexiconDecoder*>(b->d.get())->nHypothesis();
  return static_cast<LexiconFreeDecoder*>(b->d.get())->nHypothesis();
}
int ref_n_frames_in_buffer(void* dec) { return ((DecBox*)dec)->d->nDecodedFramesInBuffer(); }
int ref_best(void* dec, int lookBack, int maxLen, double* scores3, int* tokens, int* words) {
  std::vector<DecodeResult> r{((DecBox*)dec)->d->getBestHypothesis(lookBack)};
  int len = 0;
  fill(r, 1, maxLen, scores3, tokens, words, &len);
  return len;
}
int ref_all_final(void* dec, int maxHyp, int maxLen, double* scores3, int* tokens, int* words,
                  int* len) {
  auto res = ((DecBox*)dec)->d->getAllFinalHypothesis();
  return fill(res, maxHyp, maxLen, scores3, tokens, words, len);
}

// Steady-state throughput of the reference's decode() on nThreads host threads: one decoder
// object per thread (decoders are not thread-safe, decoder/Utils.h:60-63), utterances dealt
// round-robin, decode() per utterance so decodeBegin's teardown of the previous utterance's
// LMState trie is inside the timed region. Returns wall seconds for all B utterances.
double ref_bench_mt(int lexicon, const ora_options* opt, void* trie, void* lm, int sil,
                    int blank, int unk, int isLmToken, const float* emis, int B, int T, int N,
                    int nThreads, int warmup) {
  std::vector<float> tr;
  std::vector<std::unique_ptr<Decoder>> decs;
  for (int i = 0; i < nThreads; ++i)
    decs.push_back(makeDecoder(lexicon, opt, trie, lm, sil, blank, unk, tr, isLmToken));
  auto run = [&](int tid, int count) {
    for (int b = tid; b < count; b += nThreads) {
      auto res = decs[tid]->decode(emis + (size_t)b * T * N, T, N);
      (void)res;
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
