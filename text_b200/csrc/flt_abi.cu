// flt_abi.cu — implementation of include/flt_decoder.h: host-side Trie / ARPA builders with the
// reference's semantics, table upload, launch planning, and the three kernels (token-beam select,
// beam step, n-best backtrace). Compiled by nvcc for sm_100a into text_b200/lib/libflt_decoder.so.
// (tests/model builds the same file with g++ -DFLT_HOST_MODEL as a GPU-less logic harness; see
// spmd.h. That build is not shipped and nothing in the package loads it.)
#include "../../include/flt_decoder.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "runtime.h"
#include "beam_core.h"
#include "beam_lf.h"
#include "beam_gx.h"
#include "topm_core.h"
#include "fused_core.h"
#include "topm_stream.h"
#include "kernels.h"

using namespace flt;

/* =============================================================================== kernels ==== */
#if FLT_DEVICE_BUILD
__global__ void __launch_bounds__(256, 3) flt_k_topm(TopMCfg c, TopMArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  topmCta(cta, c, a, smem);
}
// streaming select (topm_stream.h): one TMA-staged row per 128-thread CTA, four CTAs per SM
// long lists (341..680 entries wanted): twice the threads rank twice the survivors, two CTAs per SM
__global__ void __launch_bounds__(2 * kStreamThreads, 2) flt_k_topm_stream256(TopMCfg c, StreamLay sl, TopMArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  topmStreamCta(cta, c, sl, a, smem);
}
__global__ void __launch_bounds__(kStreamThreads, 4) flt_k_topm_stream(TopMCfg c, StreamLay sl, TopMArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  topmStreamCta(cta, c, sl, a, smem);
}
// the beam-step kernels live in their own translation units (kern_step.cu, one object per kernel, and
// kern_gx.cu: declared in kernels.h) so that the library's device code compiles in parallel
// token-beam select + beam step fused: 8 consumer + 4 producer warps per utterance (fused_core.h)
__global__ void __launch_bounds__(kFusedConsumers + kFusedProducers, 2)
    flt_k_fused(DecCfg c, TopMCfg tc, FuseLay fl, BatchArgs a) {
  extern __shared__ __align__(128) char smem[];
  Cta cta{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  fusedCta(cta, c, tc, fl, a, smem);
}
// one warp per (utterance, rank): checkpoint hops by lane 0, then the 32-row segments in parallel
__global__ void __launch_bounds__(128) flt_k_backtrace(BacktraceArgs a) {
  __shared__ int cp[4][kBtMaxCp];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * 4 + warp;
  if (item < (long long)a.B * a.nbest) backtraceItem(a, item, lane, 32, cp[warp]);
}
#endif

namespace {

thread_local std::string gErr;
int fail(int code, const std::string& msg) {
  gErr = msg;
  return code;
}
struct FltError : std::runtime_error {
  int code;
  FltError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

constexpr int kThreads = 256; // token-select kernel
int decThreadsEnv() { // FLT_DEC_THREADS overrides the plan (experiments)
  static int t = [] {
    const char* e = getenv("FLT_DEC_THREADS");
    const int v = e ? atoi(e) : 0;
    return (v == 64 || v == 128 || v == 256 || v == 512) ? v : 0;
  }();
  return t;
}
constexpr int kTrieMaxLabel = 6; // decoder/Trie.h:19

/* ------------------------------------------------------------------ launches ---------- */
void launchTopM(const TopMCfg& c, const TopMArgs& a, int grid, size_t smem, rt::Stream s) {
#if FLT_DEVICE_BUILD
  flt_k_topm<<<grid, kThreads, smem, s>>>(c, a);
  FLT_RT_TRY(cudaGetLastError());
#else
  std::vector<char> sm(smem + 16, (char)0x5A); // shared memory is never zero for free
  for (int b = 0; b < grid; ++b) {
    Cta cta{0, 1, b, grid};
    topmCta(cta, c, a, sm.data());
  }
  (void)s;
#endif
}
void launchTopMStream(const TopMCfg& c, const StreamLay& sl, const TopMArgs& a, int grid, rt::Stream s) {
#if FLT_DEVICE_BUILD
  if (sl.threads == kStreamThreads) flt_k_topm_stream<<<grid, kStreamThreads, sl.total, s>>>(c, sl, a);
  else flt_k_topm_stream256<<<grid, 2 * kStreamThreads, sl.total, s>>>(c, sl, a);
  FLT_RT_TRY(cudaGetLastError());
#else
  std::vector<char> sm(sl.total + 128, (char)0x5A);
  for (int b = 0; b < grid; ++b) {
    Cta cta{0, 1, b, grid};
    topmStreamCta(cta, c, sl, a, sm.data());
  }
  (void)s;
#endif
}
// plan of the streaming select for a row length / list length / bias; false if the shape does not fit it
bool planStream(int N, int M, int bst, const float* dBias, float biasMax, TopMCfg& t, StreamLay& sl,
                float biasSpread = 1e30f) {
  const bool restricted = bst < N;
  const int want = restricted ? bst : M;
  // survivors aimed at: 1.5 .. 2.5 x want, inside 3/4 of the capacity (4 per thread)
  const int wantShort = getenv("FLT_STREAM_WANT") ? atoi(getenv("FLT_STREAM_WANT")) : 340;
  if (N % 4 != 0 || N < 64 || want > 2 * wantShort || M > 2048 || (dBias && restricted) || getenv("FLT_NO_STREAM")) return false;
  sl.threads = want > wantShort ? 2 * kStreamThreads : kStreamThreads;
  t = TopMCfg{};
  t.N = N;
  t.M = M;
  t.bst = restricted ? bst : N;
  t.bias = dBias;
  t.biasMax = biasMax;
  t.biasKeyed = biasSpread > 0.5f ? 1 : 0;
  t.P = std::max(sl.threads, nextPow2(want));
  t.capS = kStreamSPT * sl.threads;
  t.extra = 2 * kProdBins; // (512 ranking bins for long lists were measured slower: M = 205 39.7 % against 48.3 % of the HBM peak)
  t.fast = 1;
  t.stage = 0;
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) {
    off = (off + align - 1) / align * align;
    const size_t o = off;
    off += bytes;
    return (int)o;
  };
  TopMSmem ts;
  sl.prod = take(carveTopM(nullptr, t, ts), 16);
  sl.row = take((size_t)N * 4, 128);
  sl.mbar = take(8, 8);
  sl.total = (int)((off + 127) / 128 * 128);
  return sl.total <= 200 * 1024;
}
void launchDecode(const DecCfg& c, const BatchArgs& a, int grid, size_t smem, rt::Stream s, int threads) {
#if FLT_DEVICE_BUILD
  if (c.gx) {
    if (smem && threads == 512) (c.lexicon ? flt_k_gx512_lex : flt_k_gx512_lf)<<<grid, 512, smem, s>>>(c, a);
    else if (smem) (c.lexicon ? flt_k_gx_lex : flt_k_gx_lf)<<<grid, threads, smem, s>>>(c, a);
    else (c.lexicon ? flt_k_gx_gmem_lex : flt_k_gx_gmem_lf)<<<grid, threads > 256 ? 256 : threads, 0, s>>>(c, a);
  } else if (smem && threads == 1024 && !c.wide) flt_k_decode1024<<<grid, 1024, smem, s>>>(c, a);
  else if (smem && threads >= 512) (c.wide ? flt_k_decode512_wide : flt_k_decode512)<<<grid, 512, smem, s>>>(c, a);
  else if (smem) (c.wide ? flt_k_decode_wide : flt_k_decode)<<<grid, threads, smem, s>>>(c, a);
  else (c.wide ? flt_k_decode_gmem_wide : flt_k_decode_gmem)<<<grid, threads > 256 ? 256 : threads, 0, s>>>(c, a);
  FLT_RT_TRY(cudaGetLastError());
#else
  std::vector<char> sm(c.lay.total + 16, (char)0x5A); // shared memory is never zero for free
  for (int b = 0; b < grid; ++b) {
    Cta cta{0, 1, b, grid};
    if (c.gx && c.lexicon) gxDecodeCta<true>(cta, c, a, sm.data());
    else if (c.gx) gxDecodeCta<false>(cta, c, a, sm.data());
    else if (c.wide) decodeCta<true>(cta, c, a, sm.data());
    else decodeCta<false>(cta, c, a, sm.data());
  }
  (void)s;
  (void)smem;
  (void)threads;
#endif
}
void launchFused(const DecCfg& c, const TopMCfg& tc, const FuseLay& fl, const BatchArgs& a, int grid,
                 rt::Stream s) {
#if FLT_DEVICE_BUILD
  flt_k_fused<<<grid, kFusedConsumers + kFusedProducers, fl.total, s>>>(c, tc, fl, a);
  FLT_RT_TRY(cudaGetLastError());
#else
  std::vector<char> sm(fl.total + 128, (char)0x5A);
  for (int b = 0; b < grid; ++b) {
    Cta cta{0, 1, b, grid};
    fusedCta(cta, c, tc, fl, a, sm.data());
  }
  (void)s;
#endif
}
void launchBacktrace(const BacktraceArgs& a, rt::Stream s) {
  const long long items = (long long)a.B * a.nbest;
  if (items == 0) return;
#if FLT_DEVICE_BUILD
  flt_k_backtrace<<<(unsigned)((items + 3) / 4), 128, 0, s>>>(a);
  FLT_RT_TRY(cudaGetLastError());
#else
  std::vector<int> cp(kBtMaxCp);
  for (long long i = 0; i < items; ++i) backtraceItem(a, i, 0, 1, cp.data());
  (void)s;
#endif
}

template <class T>
T* upload(rt::DevBuf& buf, const std::vector<T>& v, rt::Stream s) {
  buf.reserve(sizeof(T) * std::max<size_t>(v.size(), 1));
  rt::h2d(buf.p, v.data(), sizeof(T) * v.size(), s);
  return buf.as<T>();
}

} // namespace

/* ================================================================================= Trie ===== */
struct HNode {
  std::map<int, int> kids; // token -> node, ascending (also the CSR edge order)
  std::vector<int> labels;
  std::vector<float> scores;
  float maxScore = 0;
};

struct flt_trie {
  int maxChildren, rootIdx;
  std::vector<HNode> nodes;
  // device images, one per CUDA device that a decoder uses the Trie on (built on first use there);
  // the Trie is frozen once the first exists
  struct Image {
    TrieDev dev{};
    rt::DevBuf childOff, childTok, childNode, maxScore, labelOff, labels, rootChild, rootLabTok, edge, node,
        rootLabEdge;
    ~Image() {
      for (rt::DevBuf* b : {&childOff, &childTok, &childNode, &maxScore, &labelOff, &labels, &rootChild,
                            &rootLabTok, &edge, &node, &rootLabEdge})
        b->release();
    }
  };
  mutable std::mutex mu;
  mutable std::map<int, std::unique_ptr<Image>> images;
  mutable bool frozen = false;

  static double logAdd(double a, double b) { // Trie.cpp:66-77
    if (a < b) std::swap(a, b);
    const double d = b - a;
    if (d < -39.14) return a;
    return a + std::log1p(std::exp(d));
  }
  void smearNode(int n, int mode) { // Trie.cpp:79-95; maxScore is a float after every step
    nodes[n].maxScore = -std::numeric_limits<float>::infinity();
    for (float s : nodes[n].scores) nodes[n].maxScore = (float)logAdd(nodes[n].maxScore, s);
    for (auto& kv : nodes[n].kids) {
      smearNode(kv.second, mode);
      const float cm = nodes[kv.second].maxScore;
      if (mode == FLT_SMEAR_LOGADD) nodes[n].maxScore = (float)logAdd(nodes[n].maxScore, cm);
      else if (mode == FLT_SMEAR_MAX && cm > nodes[n].maxScore) nodes[n].maxScore = cm;
    }
  }
  // the flattened Trie on `device` (the caller has made it current); thread-safe
  const TrieDev& deviceImage(int device, rt::Stream s) const {
    std::lock_guard<std::mutex> lock(mu);
    auto it = images.find(device);
    if (it != images.end()) return it->second->dev;
    std::unique_ptr<Image> im(new Image);
    const int nn = (int)nodes.size();
    std::vector<int> childOff(nn + 1, 0), childTok, childNode, labelOff(nn + 1, 0), labels;
    std::vector<float> maxScore(nn);
    for (int i = 0; i < nn; ++i) {
      childOff[i] = (int)childTok.size();
      for (auto& kv : nodes[i].kids) {
        childTok.push_back(kv.first);
        childNode.push_back(kv.second);
      }
      labelOff[i] = (int)labels.size();
      for (int l : nodes[i].labels) labels.push_back(l);
      maxScore[i] = nodes[i].maxScore;
    }
    childOff[nn] = (int)childTok.size();
    labelOff[nn] = (int)labels.size();
    std::vector<int> rootChildHost(std::max(maxChildren, 1), -1);
    std::vector<int> rootLabTok;
    for (auto& kv : nodes[0].kids) {
      if (kv.first >= 0 && kv.first < maxChildren) rootChildHost[kv.first] = kv.second;
      if (!nodes[kv.second].labels.empty()) rootLabTok.push_back(kv.first);
    }
    // packed records (tables.h)
    std::vector<int> edge(2 * std::max<size_t>(childTok.size(), 1)), node(4 * (size_t)nn),
        rootLabEdge(2 * std::max<size_t>(rootLabTok.size(), 1));
    for (size_t e = 0; e < childTok.size(); ++e) {
      edge[2 * e] = childTok[e];
      edge[2 * e + 1] = childNode[e];
    }
    for (int i = 0; i < nn; ++i) {
      const int deg = childOff[i + 1] - childOff[i], nl = labelOff[i + 1] - labelOff[i];
      if (deg >= (1 << 24)) throw std::runtime_error("trie node with more than 2^24 children");
      node[4 * (size_t)i] = (int)f32Bits(maxScore[i]);
      node[4 * (size_t)i + 1] = childOff[i];
      node[4 * (size_t)i + 2] = deg | (nl << 24);
      node[4 * (size_t)i + 3] = nl == 1 ? labels[labelOff[i]] : labelOff[i];
    }
    for (size_t k = 0; k < rootLabTok.size(); ++k) {
      rootLabEdge[2 * k] = rootLabTok[k];
      rootLabEdge[2 * k + 1] = rootChildHost[rootLabTok[k]];
    }
    TrieDev& dev = im->dev;
    dev.nNodes = nn;
    dev.childOff = upload(im->childOff, childOff, s);
    dev.childTok = upload(im->childTok, childTok, s);
    dev.childNode = upload(im->childNode, childNode, s);
    dev.maxScore = upload(im->maxScore, maxScore, s);
    dev.labelOff = upload(im->labelOff, labelOff, s);
    dev.labels = upload(im->labels, labels, s);
    dev.rootChild = upload(im->rootChild, rootChildHost, s);
    dev.nRootLab = (int)rootLabTok.size();
    dev.rootLabTok = upload(im->rootLabTok, rootLabTok, s);
    dev.edge = (const int2*)upload(im->edge, edge, s);
    dev.node = (const int4*)upload(im->node, node, s);
    dev.rootLabEdge = (const int2*)upload(im->rootLabEdge, rootLabEdge, s);
    rt::sync(s);
    frozen = true;
    return images.emplace(device, std::move(im)).first->second->dev;
  }
};

/* =================================================================================== LM ===== */
struct flt_lm {
  int kind = 0; // 0 zero, 1 ngram
  int order = 0, vocab = 0, bos = -1, eos = -1;
  std::vector<F2> uni;
  std::vector<uint64_t> keys[kMaxOrder + 1];
  std::vector<uint64_t> chk[kMaxOrder + 1];
  std::vector<F2> vals[kMaxOrder + 1];
  std::vector<int> usr2lm;
  std::vector<std::string> words; // vocabulary by LM id (kept for the table file, table_io.h)
  LmDev host{}; // view over the host vectors (host-side scoring)
  float upper = 0.0f; // no word scores above this: best probability + the positive back-offs
  struct Image {
    LmDev dev{};
    rt::DevBuf uni, usr, keys[kMaxOrder + 1], chk[kMaxOrder + 1], vals[kMaxOrder + 1];
    ~Image() {
      uni.release(), usr.release();
      for (int n = 0; n <= kMaxOrder; ++n) keys[n].release(), chk[n].release(), vals[n].release();
    }
  };
  mutable std::mutex mu;
  mutable std::map<int, std::unique_ptr<Image>> images;

  void makeHostView() {
    host = LmDev{};
    host.kind = kind;
    host.order = order;
    host.vocab = vocab;
    host.bos = bos;
    host.eos = eos;
    host.nUsr = (int)usr2lm.size();
    host.usr2lm = usr2lm.data();
    host.uni = uni.data();
    float best = -std::numeric_limits<float>::infinity(), bo = 0.0f;
    for (const F2& v : uni) best = std::max(best, v.x);
    for (int n = 2; n <= kMaxOrder; ++n) {
      host.keys[n] = keys[n].empty() ? nullptr : keys[n].data();
      host.chk[n] = chk[n].empty() ? nullptr : chk[n].data();
      host.vals[n] = vals[n].empty() ? nullptr : vals[n].data();
      host.mask[n] = keys[n].empty() ? 0 : (uint32_t)keys[n].size() - 1;
      for (size_t i = 0; i < keys[n].size(); ++i)
        if (keys[n][i]) best = std::max(best, vals[n][i].x);
    }
    // at most one back-off per context order is added to a score (tables.h ngramScore)
    for (int n = 1; n < std::max(order, 1); ++n) {
      float m = 0.0f;
      if (n == 1) {
        for (const F2& v : uni) m = std::max(m, v.y);
      } else {
        for (size_t i = 0; i < keys[n].size(); ++i)
          if (keys[n][i]) m = std::max(m, vals[n][i].y);
      }
      bo += m;
    }
    upper = kind == 0 ? 0.0f : best + bo + 1e-3f;
  }
  // the tables on `device` (the caller has made it current); thread-safe
  const LmDev& deviceImage(int device, rt::Stream s) const {
    std::lock_guard<std::mutex> lock(mu);
    auto it = images.find(device);
    if (it != images.end()) return it->second->dev;
    std::unique_ptr<Image> im(new Image);
    im->dev = host;
    if (kind == 1) {
      im->dev.usr2lm = upload(im->usr, usr2lm, s);
      im->dev.uni = upload(im->uni, uni, s);
      for (int n = 2; n <= kMaxOrder; ++n) {
        if (keys[n].empty()) continue;
        im->dev.keys[n] = upload(im->keys[n], keys[n], s);
        im->dev.chk[n] = upload(im->chk[n], chk[n], s);
        im->dev.vals[n] = upload(im->vals[n], vals[n], s);
      }
      rt::sync(s);
    }
    return images.emplace(device, std::move(im)).first->second->dev;
  }
};

namespace {

// ARPA reader: vocabulary ids in unigram order with <unk> forced to id 0 (KenLM convention), log10
// values kept as parsed floats, n-grams of order >= 2 hashed into per-order tables.
void loadArpa(const std::string& path, const char* const* usrWords, int nUsr, flt_lm& lm) {
  std::ifstream in(path);
  if (!in) throw FltError(FLT_ERR_RUNTIME, "[KenLM] LM loading failed: cannot open " + path);
  std::unordered_map<std::string, int> vocab;
  std::vector<long> counts(kMaxOrder + 2, 0);
  struct Entry {
    uint64_t key, key2;
    F2 v;
  };
  std::vector<Entry> pending[kMaxOrder + 1];
  std::string line;
  bool inData = false;
  int section = 0;
  std::vector<std::string> toks;
  lm.kind = 1;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    if (line == "\\data\\") {
      inData = true;
      continue;
    }
    if (line == "\\end\\") break;
    if (line[0] == '\\') {
      section = atoi(line.c_str() + 1);
      if (section < 1 || section > kMaxOrder)
        throw FltError(FLT_ERR_RUNTIME, "[KenLM] unsupported n-gram order in " + line);
      if (section == 1) {
        vocab["<unk>"] = 0;
        lm.uni.push_back(F2{-100.0f, 0.0f});
      } else {
        pending[section].reserve((size_t)counts[section]);
      }
      inData = false;
      continue;
    }
    if (inData) {
      if (line.compare(0, 6, "ngram ") == 0) {
        const int n = atoi(line.c_str() + 6);
        const size_t eq = line.find('=');
        if (n >= 1 && n <= kMaxOrder && eq != std::string::npos) {
          counts[n] = atol(line.c_str() + eq + 1);
          lm.order = std::max(lm.order, n);
        }
      }
      continue;
    }
    if (section == 0) continue;
    toks.clear();
    size_t p = 0;
    while (p < line.size()) {
      size_t q = line.find_first_of(" \t", p);
      if (q == std::string::npos) q = line.size();
      if (q > p) toks.emplace_back(line.substr(p, q - p));
      p = q + 1;
    }
    if ((int)toks.size() < section + 1) continue;
    F2 v;
    v.x = strtof(toks[0].c_str(), nullptr);
    v.y = (int)toks.size() > section + 1 ? strtof(toks[section + 1].c_str(), nullptr) : 0.0f;
    if (section == 1) {
      if (toks[1] == "<unk>") {
        lm.uni[0] = v;
      } else {
        vocab[toks[1]] = (int)lm.uni.size();
        lm.uni.push_back(v);
      }
    } else {
      // chain over the words in reversed order (tables.h)
      uint64_t h = 0, h2 = 0;
      for (int i = section; i >= 1; --i) {
        auto it = vocab.find(toks[i]);
        const int w = it == vocab.end() ? 0 : it->second;
        h = i == section ? ngramChainStart(w) : ngramChainExtend(h, w);
        h2 = i == section ? ngramChain2Start(w) : ngramChain2Extend(h2, w);
      }
      pending[section].push_back(Entry{ngramFinalKey(h), h2, v});
    }
  }
  if (lm.order < 1 || lm.uni.empty()) throw FltError(FLT_ERR_RUNTIME, "[KenLM] LM loading failed: empty model");
  auto b = vocab.find("<s>"), e = vocab.find("</s>");
  if (b == vocab.end() || e == vocab.end())
    throw FltError(FLT_ERR_RUNTIME, "[KenLM] LM vocabulary loading failed: missing <s> or </s>");
  lm.bos = b->second;
  lm.eos = e->second;
  lm.vocab = (int)lm.uni.size();
  lm.words.assign(lm.uni.size(), std::string());
  for (auto& kv : vocab) lm.words[(size_t)kv.second] = kv.first;
  for (int n = 2; n <= lm.order; ++n) {
    if (pending[n].empty()) continue;
    size_t cap = 16;
    while (cap < pending[n].size() * 2) cap <<= 1;
    lm.keys[n].assign(cap, 0);
    lm.chk[n].assign(cap, 0);
    lm.vals[n].assign(cap, F2{0, 0});
    const uint32_t mask = (uint32_t)cap - 1;
    for (const Entry& en : pending[n]) {
      uint32_t s = (uint32_t)(en.key >> 17) & mask;
      while (lm.keys[n][s] != 0 && !(lm.keys[n][s] == en.key && lm.chk[n][s] == en.key2)) s = (s + 1) & mask;
      lm.keys[n][s] = en.key; // a repeated n-gram keeps the last value, like a rebuilt table
      lm.chk[n][s] = en.key2;
      lm.vals[n][s] = en.v;
    }
    std::vector<Entry>().swap(pending[n]);
  }
  lm.usr2lm.resize(nUsr);
  for (int i = 0; i < nUsr; ++i) {
    auto it = vocab.find(usrWords[i]);
    lm.usr2lm[i] = it == vocab.end() ? 0 : it->second; // Vocabulary::Index: OOV -> <unk>
  }
  lm.makeHostView();
}

} // namespace

#include "table_io.h"

/* ============================================================================== decoder ===== */
struct flt_decoder {
  int lexicon = 0;
  flt_options opt{};
  const flt_trie* trie = nullptr;
  const flt_lm* lm = nullptr;
  int sil = 0, blank = -1, unk = -1;
  std::vector<float> trans;
  int isLmToken = 0;
  int device = 0;
  int nbest = 0;
  rt::Stream stream{};
  rt::Stream copyStream{};
  rt::Stream copyStream2{}; // second copy queue: two host->device copies in flight (flt_decode_batch, host input)
#if FLT_DEVICE_BUILD
  cudaEvent_t evCopied[3]{}, evFree[3]{};
  int numSMs = 148;
  std::vector<cudaEvent_t> evPool; // kernel timing: (start, stop) pairs, kind = index % 3
  size_t evUsed = 0;
#endif
  bool timing = false;   // CUDA events around each launch
  bool counters = false; // in-kernel work / phase counters (costs a few hundred cycles per frame)
  std::vector<int> evKinds;
  // plan
  int planN = -1;
  DecCfg cfg{};
  TopMCfg tcfg{};
  bool needTopM = false;
  // online decoding of one utterance (Decoder.h:18-35): the beam lives on the device between
  // decodeStep launches, the per-frame records are mirrored on the host for prune / best
  struct SHyp {
    double score, am, lm;
    int parent, token, word;
  };
  struct Online {
    bool begun = false;
    int N = 0, nDecoded = 0, nPruned = 0;
    bool haveBeam = false;
    double shift = 0.0;
    std::vector<std::vector<SHyp>> hyp; // rows in the buffer (row 0 = oldest kept frame)
  } on;
  rt::DevBuf sBeam, sBeamBackup, sScore, sCount, sEmis;
  int threads = 256;  // threads per utterance of the beam-step kernel
  bool fused = false; // select + step in one kernel (fused_core.h)
  TopMCfg ftcfg{};
  FuseLay flay{};
  bool streamSel = false; // token-beam select by the streaming kernel (topm_stream.h)
  TopMCfg stcfg{};
  StreamLay slay{};
  int streamGridMax = 1;
  int fusedGridMax = 1;
  size_t wsBytes = 0, topmSmem = 0; // wsBytes: whole workspace (both regions)
  float biasSpread = 0.0f; // max - min of the finite rank offsets of the root children (lexicon decoder)
  size_t smemBytes = 0, slabBytes = 0; // dynamic shared memory per CTA / global slab per CTA of the step kernel
  int gridMax = 1, topmGridMax = 1;
  std::vector<int> wideOffHost;
  rt::DevBuf dWideOff, dBias, dTrans, dLfDesc;
  // batch buffers
  rt::DevBuf hSkip, hSkipFin;
  rt::DevBuf topTok, topVal, thr, hPar, hTok, hWord, finScore, finCount, status, ws, outTok,
      outWord, dLengths, staging[3], dStats;
  int lastB = 0, lastT = 0, launches = 0;
  int lastNbest = 0; // nbest the last batch's outTok / outWord rows were sized and strided with
  int capBoost = 1; // candidate-capacity multiplier, grown after an overflow
  bool useSmemFlag = false;
  std::vector<int> lastLengths;
  bool haveLengths = false;

  ~flt_decoder() {
    for (rt::DevBuf* b : {&dWideOff, &dBias, &dTrans, &dLfDesc, &topTok, &topVal, &thr, &hPar, &hTok, &hWord,
                          &finScore, &finCount, &status, &ws, &outTok, &outWord, &dLengths,
                          &staging[0], &staging[1], &staging[2], &dStats, &hSkip, &hSkipFin, &sBeam, &sBeamBackup, &sScore, &sCount, &sEmis})
      b->release();
#if FLT_DEVICE_BUILD
    for (int i = 0; i < 3; ++i) {
      if (evCopied[i]) cudaEventDestroy(evCopied[i]);
      if (evFree[i]) cudaEventDestroy(evFree[i]);
    }
    for (cudaEvent_t e : evPool) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
    if (copyStream) cudaStreamDestroy(copyStream);
    if (copyStream2) cudaStreamDestroy(copyStream2);
#endif
  }
};

namespace {

void planFor(flt_decoder& d, int N) {
  if (d.planN == N) return;
  const flt_options& o = d.opt;
  DecCfg c{};
  c.lexicon = d.lexicon;
  c.K = o.beamSize;
  c.N = N;
  c.setAll = o.beamSizeToken >= N;
  c.beamThreshold = o.beamThreshold;
  c.lmWeight = o.lmWeight;
  c.wordScore = o.wordScore;
  c.unkScore = o.unkScore;
  c.silScore = o.silScore;
  c.logAdd = o.logAdd;
  c.ctc = o.criterionType == FLT_CRITERION_CTC;
  c.hasUnk = d.lexicon && o.unkScore > -std::numeric_limits<double>::infinity();
  c.sil = d.sil;
  c.blank = d.blank;
  c.unk = d.unk;
  if (o.beamSize < 1) throw FltError(FLT_ERR_INVALID, "beamSize must be >= 1");
  if (o.beamSizeToken < 1) throw FltError(FLT_ERR_INVALID, "beamSizeToken must be >= 1");
  if (!(o.beamThreshold >= 0)) throw FltError(FLT_ERR_INVALID, "beamThreshold must be >= 0");
  if (o.criterionType != FLT_CRITERION_CTC && o.criterionType != FLT_CRITERION_ASG)
    throw FltError(FLT_ERR_UNSUPPORTED, "criterion type must be ASG or CTC for these decoders");
  if (d.sil < 0 || d.sil >= N) throw FltError(FLT_ERR_INVALID, "sil index out of range for N");
  if (c.ctc && (d.blank < 0 || d.blank >= N))
    throw FltError(FLT_ERR_INVALID, "blank index out of range for N (CTC)");
  if (!c.ctc && !d.trans.empty() && (long long)d.trans.size() < (long long)N * N)
    throw FltError(FLT_ERR_INVALID, "transitions must hold N*N entries (ASG)");
  if (d.lexicon && d.trie->maxChildren < 1) throw FltError(FLT_ERR_INVALID, "empty trie");
  c.lmToken = d.lexicon && d.isLmToken;
  // lexicon-free decoder without rank dominance (logAdd merging, n-gram token LM): full expansion
  c.full = !d.lexicon && (o.logAdd || d.lm->kind != 0);
  if (d.lm->kind == 1) {
    // indices the LM will be asked about (KenLM::score throws on the first one out of range,
    // lm/KenLM.cpp:64-68; here the whole range is checked up front)
    int maxIdx = -1;
    if (!d.lexicon) {
      maxIdx = N - 1;
    } else if (c.lmToken) {
      for (auto& nd : d.trie->nodes)
        for (auto& kv : nd.kids) maxIdx = std::max(maxIdx, kv.first);
    } else {
      for (auto& nd : d.trie->nodes)
        for (int l : nd.labels) maxIdx = std::max(maxIdx, l);
      if (c.hasUnk) maxIdx = std::max(maxIdx, d.unk);
    }
    if (maxIdx >= (int)d.lm->usr2lm.size() || (d.lexicon && !c.lmToken && c.hasUnk && d.unk < 0))
      throw FltError(FLT_ERR_RUNTIME, "[KenLM] Invalid user token index: " + std::to_string(maxIdx));
  }

  // the lexicon step has ~4x the work items per frame of the lexicon-free one
  // ... and the full expansion with an n-gram token LM is bound by its table probes: twice the threads
  // keep twice the gathers in flight (measured: 87.8 -> 49.5 ms per step at beam 50, bst 50)
  d.threads = decThreadsEnv() ? decThreadsEnv() : ((d.lexicon || (c.full && d.lm->kind != 0)) ? 512 : 256);
  const int K = c.K;
  const int bstEff = std::min(o.beamSizeToken, N);
  // ranked wide rows need max-merge and scores that follow the per-frame token order
  c.wideRanked = d.lexicon ? (c.ctc && !c.hasUnk && !o.logAdd && !c.lmToken) : !c.full;
  d.needTopM = c.wideRanked || !c.setAll;
  TopMCfg t{};
  t.N = N;
  t.bst = c.setAll ? N : bstEff;
  t.capS = 2048;
  if (c.wideRanked) {
    const int slack = d.lexicon ? 2 : 0; // fp32 rank keys of e+bias may swap near-equal neighbours
    c.Mwide = std::min(K + 3 + slack, bstEff);
    c.M = (d.lexicon || c.setAll) ? c.Mwide : bstEff; // lexicon-free restricted: list = whole set
  } else {
    c.Mwide = 0;
    // full expansion / unranked lexicon rows with a token beam: the list is the whole token set
    c.rootList = d.lexicon && !c.setAll;
    c.M = ((c.full || c.rootList) && !c.setAll) ? bstEff : 1;
  }
  c.wide = c.full || o.logAdd || c.lmToken || c.rootList;
  c.dbg = getenv("FLT_DBG") ? atoi(getenv("FLT_DBG")) : 0;
  const int want = c.setAll ? c.M : bstEff;
  if (want > 2048)
    throw FltError(FLT_ERR_UNSUPPORTED,
                   "beamSize / beamSizeToken combination needs a token list longer than 2048 "
                   "(beamSizeToken < N and > 2048, or beamSize > 2040 in ranked mode)");
  t.M = c.M;
  t.P = std::max(kThreads, nextPow2(want));
  // register-resident fast path (alignment of the emission pointer is checked per launch)
  t.fast = (N % 4 == 0) && N <= 4 * kFastVec * kThreads && want <= 256 && c.M <= 256;
  t.stage = !t.fast && (size_t)N * 4 <= 100 * 1024;
  // single-pass step with a guessed cut (beam_gx.h): max-merge, word-level LM; lexicon: CTC without unk
  // (FLT_GX=1 / 0 forces it on / off where it applies)
  const bool gxCan = !o.logAdd && K <= 4095 && (d.lexicon ? c.wideRanked != 0 : !c.full);
  const bool gxAuto = false;
  c.gx = gxCan && (getenv("FLT_GX") ? atoi(getenv("FLT_GX")) != 0 : gxAuto);
  // lexicon-free fast step (beam_lf.h): ZeroLM max-merge, candidate indices fit 16 bits
  c.lfFast = !c.gx && !d.lexicon && !c.full && K <= 256 && !getenv("FLT_NO_LF");
  c.lfBins = std::min(1024, std::max(256, nextPow2(4 * K)));
  // wide offsets
  d.wideOffHost.assign(K + 1, 0);
  for (int r = 1; r <= K; ++r) {
    // fast step: item row = hypothesis index r-1, which spans >= (r-1)/2+1 rows
    const int rows = c.lfFast ? (r - 1) / 2 + 1 : r;
    d.wideOffHost[r] = d.wideOffHost[r - 1] + std::min(c.Mwide, K / rows + 3 + (d.lexicon ? 2 : 0));
  }
  // lexicon decoder, max-merge: two-pass histogram pruning keeps ~3K+64 candidates per frame (plus
  // the rest of the cut bin), so the workspace fits shared memory
  c.prune2 = !c.gx && (d.lexicon || (c.full && !getenv("FLT_NO_PRUNE2_FULL"))) && !o.logAdd && !getenv("FLT_NO_PRUNE2");
  // full expansion proposes up to K * |token set| candidates; a finite beamThreshold usually leaves
  // far fewer, so start from a budget and let the overflow retry (capBoost) grow it
  const long long fullCells = c.full ? (long long)K * (c.setAll ? N : bstEff) : 0;
  const long long narrowBudget = d.lexicon ? (c.prune2 ? 512 : std::max<long long>(4096, 24LL * K))
                                           : (c.full ? (c.prune2 ? 512 : std::max<long long>(8192, 64LL * K)) : 0);
  // candidates the two-pass pruning keeps per frame: 1.5K+32 (FLT_PRUNE_WANT=<percent of K> to change);
  // 3K+64 for 64 frames after a frame had to be redone (beam_core.h). Exactness does not depend on it: a
  // frame whose kept bins hold fewer than K merge groups is redone without the cut. Measured on cfg 3:
  // 300 % 31.3 ms, 200 % 30.0 ms, 150 % 29.3 ms, 125 % 49.5 ms (redo storms)
  c.pruneWant = std::max(K + 1, (int)((long long)K * (getenv("FLT_PRUNE_WANT") ? atoi(getenv("FLT_PRUNE_WANT")) : 150) / 100) + 32);
  // FLT_TEST_CAP=<n>: start from a tiny budget so that tests reach the overflow retry
  const long long budget0 = getenv("FLT_TEST_CAP") ? std::max(1, atoi(getenv("FLT_TEST_CAP"))) : narrowBudget;
  long long capC = c.prune2 ? 3LL * K + 64 + budget0 * d.capBoost
                            : (long long)(c.wideRanked ? d.wideOffHost[K] : 0) + 3LL * K + budget0 * d.capBoost;
  if (c.full && !c.prune2) capC = 3LL * K + std::min(fullCells, budget0 * d.capBoost);
  if (c.gx) {
    // the guess aims at 1.5 K .. 3 K merge groups; duplicates of a group (members of one row) come on top
    capC = getenv("FLT_TEST_CAP") ? std::max<long long>(K + 16, budget0) * d.capBoost : (4LL * K + 128) * d.capBoost;
    c.capChunks = (int)std::min<long long>((4LL * K + 64) * d.capBoost, 1 << 20);
  }
  capC = (capC + 63) / 64 * 64;
  if (capC > (1LL << 26)) throw FltError(FLT_ERR_RUNTIME, "candidate capacity exceeded");
  c.capC = (int)capC;
  c.capH = nextPow2((int)std::min<long long>(2 * capC, 1LL << 27));
  c.capRH = nextPow2((c.lfFast ? 8 : 2) * K); // fast step: sparse table, probes mostly end at once
  c.capP = nextPow2(K);
  c.wideTotal = (c.wideRanked && !c.gx) ? d.wideOffHost[K] : 0;
  // pruning rectangles (beam_core.h frameStep): a rows x (ceil(K/a)+3) columns
  c.nTau = 0;
  if (c.wideRanked && !d.lexicon) {
    const int as[16] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64};
    for (int k = 0; k < 16; ++k) {
      const int a = as[k];
      const int col = (K + a - 1) / a + 2;
      if (2 * a - 2 < K && col < c.Mwide) {
        c.tauA[c.nTau] = a;
        c.tauCol[c.nTau] = col;
        c.nTau++;
      }
    }
  }
  c.listInSmem = d.needTopM && (c.M <= 2 * d.threads || c.gx);

  rt::Stream s = d.stream;
  c.wideOff = upload(d.dWideOff, d.wideOffHost, s);
  c.lfDesc = nullptr;
  if (c.lfFast) {
    std::vector<int> desc; // beam_lf.h: repeat, blank (and boosted-sil) items first, then the cells column-major
    for (int kind = 1; kind <= (o.silScore > 0 ? 3 : 2); ++kind)
      for (int p = 0; p < K; ++p) desc.push_back(p | (kind << 24));
    for (int j = 0; j < c.Mwide; ++j)
      for (int p = 0; p < K; ++p)
        if (j < d.wideOffHost[p + 1] - d.wideOffHost[p]) desc.push_back(p | (j << 12));
    desc.resize(c.capC, 0xFFF); // padding: hypothesis 4095 >= nH, dead
    c.lfDesc = upload(d.dLfDesc, desc, s);
  }
  c.trans = nullptr;
  if (!c.ctc && !d.trans.empty()) c.trans = upload(d.dTrans, d.trans, s);
  if (d.lexicon) {
    c.trie = d.trie->deviceImage(d.device, s);
  }
  c.lm = d.lm->deviceImage(d.device, s);
  c.lmUpper = d.lm->upper;
  t.bias = nullptr;
  if (d.lexicon && c.wideRanked) {
    // rank key offset of a root child: lmWeight * smeared score; -inf = not expandable as (1a)
    std::vector<float> bias(N, -std::numeric_limits<float>::infinity());
    const flt_trie& tr = *d.trie;
    for (auto& kv : tr.nodes[0].kids) {
      if (kv.first < 0 || kv.first >= N) continue;
      if (tr.nodes[kv.second].kids.empty()) continue;
      bias[kv.first] = (float)(o.lmWeight * (double)tr.nodes[kv.second].maxScore);
    }
    t.bias = upload(d.dBias, bias, s);
    t.biasMax = 0.0f;
    float biasMin = 0.0f;
    bool any = false;
    for (float b : bias)
      if (!isNegInf(b)) {
        t.biasMax = any ? std::max(t.biasMax, b) : b;
        biasMin = any ? std::min(biasMin, b) : b;
        any = true;
      }
    d.biasSpread = any ? t.biasMax - biasMin : 0.0f;
  }
  rt::sync(s);

  makeLayout(c);
  d.wsBytes = ((size_t)c.lay.total + 255) / 256 * 256;
  TopMSmem ts;
  d.topmSmem = (carveTopM(nullptr, t, ts) + 255) / 256 * 256;
  d.cfg = c;
  d.tcfg = t;
  d.streamSel = d.needTopM && planStream(N, c.M, t.bst, t.bias, t.biasMax, d.stcfg, d.slay, d.biasSpread);
  // fused select + step: the row stage, the producer scratch and the consumer workspace share the
  // CTA's shared memory; two CTAs per SM need <= 113 KB each
  d.fused = false;
  if ((c.lfFast || (c.gx && !d.lexicon)) && !getenv("FLT_NO_FUSED") && N % 4 == 0 &&
      want <= 32 * (kFusedProducers / 32) && d.threads == 256) {
    TopMCfg ft = t;
    ft.P = std::max(kFusedProducers, nextPow2(want));
    ft.capS = kProdCap;
    ft.extra = 2 * kProdBins;
    ft.fast = 1;
    ft.stage = 0;
    ft.bias = nullptr;
    FuseLay fl{};
    size_t off = 0;
    auto take = [&](size_t bytes, size_t align) {
      off = (off + align - 1) / align * align;
      const size_t o = off;
      off += bytes;
      return (int)o;
    };
    fl.ws = take(c.lay.total, 16);
    TopMSmem ts2;
    fl.prod = take(carveTopM(nullptr, ft, ts2), 16);
    fl.row = take((size_t)N * 4, 128);
    for (int k = 0; k < kFusedRing; ++k) fl.list[k] = take(8 * (size_t)c.M, 16);
    for (int k = 0; k < kFusedRing; ++k) fl.thr[k] = take(4, 4);
    for (int k = 0; k < kFusedRing; ++k) fl.spec[k] = take(8, 8);
    fl.mbar = take(8 * MB_COUNT, 8);
    fl.total = (int)((off + 127) / 128 * 128);
    d.ftcfg = ft;
    d.flay = fl;
    d.fused = fl.total <= 113 * 1024;
  }
#if FLT_DEVICE_BUILD
  int dev = d.device;
  cudaDeviceProp prop;
  FLT_RT_TRY(cudaGetDeviceProperties(&prop, dev));
  d.numSMs = prop.multiProcessorCount;
  const size_t smemMax = prop.sharedMemPerBlockOptin;
  if (d.topmSmem > smemMax) throw FltError(FLT_ERR_UNSUPPORTED, "N too large for the select kernel's shared memory");
  FLT_RT_TRY(cudaFuncSetAttribute(flt_k_topm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)std::max<size_t>(d.topmSmem, 48 * 1024)));
  int occ = 1;
  FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, flt_k_topm, kThreads, d.topmSmem));
  d.topmGridMax = std::max(1, occ) * d.numSMs;
  if (d.streamSel) {
    if ((size_t)d.slay.total > smemMax) {
      d.streamSel = false;
    } else {
      auto* ks = d.slay.threads == kStreamThreads ? flt_k_topm_stream : flt_k_topm_stream256;
      FLT_RT_TRY(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(d.slay.total, 48 * 1024)));
      int occS = 1;
      FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occS, ks, d.slay.threads, d.slay.total));
      d.streamGridMax = std::max(1, occS) * d.numSMs;
    }
  }
  // <= 110 KB keeps two CTAs per SM; FLT_SMEM_KB raises the limit (one CTA per SM) for experiments
  // (the single-pass step keeps large beams on chip with one CTA per SM rather than spilling to a slab)
  const size_t smemLimit = getenv("FLT_SMEM_KB") ? (size_t)atoi(getenv("FLT_SMEM_KB")) * 1024
                                                 : (c.gx ? smemMax : 110 * 1024);
  // three placements of the workspace (beam_core.h Ws): everything in shared memory; the small region in
  // shared memory and the capacity-sized arrays in the CTA's global slab (the candidate capacity of a wide
  // beam, or one grown by the overflow retry, must not push the histogram and the beams off the chip: the
  // small region may then take a whole SM's shared memory); everything in the slab
  const size_t smallBytes = ((size_t)c.lay.small + 255) / 256 * 256;
  const bool smemOk = d.wsBytes <= std::min<size_t>(smemMax, smemLimit);
  const bool hybrid = !smemOk && c.lay.bigBytes > 0 && smallBytes <= smemMax && !getenv("FLT_NO_HYBRID");
  d.cfg.bigGlobal = c.bigGlobal = hybrid ? 1 : 0;
  d.smemBytes = smemOk ? d.wsBytes : (hybrid ? smallBytes : 0);
  d.slabBytes = smemOk ? 0 : (hybrid ? ((size_t)c.lay.bigBytes + 255) / 256 * 256 : d.wsBytes);
  int occ2 = 1;
  auto* k256 = c.gx ? (c.lexicon ? flt_k_gx_lex : flt_k_gx_lf) : (c.wide ? flt_k_decode_wide : flt_k_decode);
  auto* k512 = c.gx ? (c.lexicon ? flt_k_gx512_lex : flt_k_gx512_lf)
                    : (c.wide ? flt_k_decode512_wide : flt_k_decode512);
  auto* kGmem = c.gx ? (c.lexicon ? flt_k_gx_gmem_lex : flt_k_gx_gmem_lf)
                     : (c.wide ? flt_k_decode_gmem_wide : flt_k_decode_gmem);
  if (d.smemBytes) {
    FLT_RT_TRY(cudaFuncSetAttribute(k256, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)std::max<size_t>(d.smemBytes, 48 * 1024)));
    FLT_RT_TRY(cudaFuncSetAttribute(k512, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)std::max<size_t>(d.smemBytes, 48 * 1024)));
    if (d.threads == 512) {
      FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k512, 512, d.smemBytes));
      // one CTA per SM only (the small region of a very wide beam): give it all 32 warps of the SM
      if (occ2 == 1 && !c.gx && !c.wide && !decThreadsEnv() && !getenv("FLT_NO_1024")) {
        FLT_RT_TRY(cudaFuncSetAttribute(flt_k_decode1024, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)std::max<size_t>(d.smemBytes, 48 * 1024)));
        int occ1k = 0;
        FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1k, flt_k_decode1024, 1024, d.smemBytes));
        if (occ1k >= 1) d.threads = 1024;
      }
    } else
      FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k256, d.threads, d.smemBytes));
  } else {
    FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kGmem, d.threads > 256 ? 256 : d.threads, 0));
    occ2 = std::min(occ2, 4);
  }
  d.gridMax = std::max(1, occ2) * d.numSMs;
  d.useSmemFlag = d.smemBytes != 0;
  if (d.fused) {
    if ((size_t)d.flay.total > smemMax) {
      d.fused = false;
    } else {
      FLT_RT_TRY(cudaFuncSetAttribute(flt_k_fused, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      std::max(d.flay.total, 48 * 1024)));
      int occ3 = 1;
      FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &occ3, flt_k_fused, kFusedConsumers + kFusedProducers, d.flay.total));
      d.fusedGridMax = std::max(1, occ3) * d.numSMs;
    }
  }
#else
  d.gridMax = 4;
  d.topmGridMax = 4;
  d.streamGridMax = 3;
  d.fusedGridMax = 3;
  d.useSmemFlag = true;
  // host model: FLT_TEST_HYBRID=1 puts the capacity-sized region into a separate slab, as the device does
  // when the workspace outgrows shared memory
  d.cfg.bigGlobal = c.bigGlobal = (getenv("FLT_TEST_HYBRID") && c.lay.bigBytes > 0) ? 1 : 0;
  d.smemBytes = c.bigGlobal ? (size_t)c.lay.small : d.wsBytes;
  d.slabBytes = c.bigGlobal ? (size_t)c.lay.bigBytes : 0;
#endif
  d.planN = N;
  if (getenv("FLT_DBG_PLAN"))
    fprintf(stderr, "[flt plan] lexicon=%d K=%d N=%d M=%d gx=%d lfFast=%d fused=%d wide=%d prune2=%d threads=%d capC=%d "
                    "capChunks=%d ws=%zu B (%s) fusedSmem=%d grid<=%d streamSelect=%d (%d B, grid<=%d)\n",
            c.lexicon, c.K, N, c.M, c.gx, c.lfFast, (int)d.fused, c.wide, c.prune2, d.threads, c.capC, c.capChunks,
            d.wsBytes, d.slabBytes == 0 ? "shared" : (d.smemBytes ? "small region shared, capacity-sized arrays in a global slab" : "global slab"), d.fused ? d.flay.total : 0,
            d.fused ? d.fusedGridMax : d.gridMax, (int)d.streamSel, d.streamSel ? d.slay.total : 0, d.streamGridMax);
}

} // namespace

namespace {

struct KernelTimer { // records a CUDA-event pair around one launch when timing is on
  flt_decoder& d;
  int kind;
  KernelTimer(flt_decoder& dd, int k) : d(dd), kind(k) {
#if FLT_DEVICE_BUILD
    if (!d.timing) return;
    if (d.evUsed + 2 > d.evPool.size()) {
      for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        FLT_RT_TRY(cudaEventCreate(&e));
        d.evPool.push_back(e);
      }
    }
    FLT_RT_TRY(cudaEventRecord(d.evPool[d.evUsed], d.stream));
#endif
  }
  ~KernelTimer() {
#if FLT_DEVICE_BUILD
    if (!d.timing) return;
    cudaEventRecord(d.evPool[d.evUsed + 1], d.stream);
    d.evUsed += 2;
    d.evKinds.push_back(kind);
#endif
  }
};

// Run the three kernels over `Bc` utterances whose emissions are device-resident at dEmis, writing
// n-best rows [outBase, outBase+Bc) of the decoder's output buffers.
void runChunk(flt_decoder& d, const float* dEmis, int Bc, int T, int N, const int* dLen,
              long long outBase) {
  const DecCfg& c = d.cfg;
  const int K = c.K;
  rt::Stream s = d.stream;
  const long long rows = (long long)Bc * T;
  BatchArgs a{};
  a.emis = dEmis;
  a.B = Bc;
  a.T = T;
  a.lengths = dLen;
  const bool fused = d.fused && (reinterpret_cast<uintptr_t>(dEmis) & 15) == 0;
  if (d.needTopM && !fused) {
    d.topTok.reserve(sizeof(int) * rows * c.M);
    d.topVal.reserve(sizeof(float) * rows * c.M);
    if (!c.setAll) d.thr.reserve(sizeof(float) * rows);
    TopMArgs ta{};
    ta.emis = dEmis;
    ta.rows = rows;
    ta.outTok = d.topTok.as<int>();
    ta.outVal = d.topVal.as<float>();
    ta.outThr = c.setAll ? nullptr : d.thr.as<float>();
    TopMCfg tc = d.tcfg;
    const bool aligned = (reinterpret_cast<uintptr_t>(dEmis) & 15) == 0;
    if (!aligned) tc.fast = 0;
    const int grid = (int)std::min<long long>(rows, d.topmGridMax * 8LL);
    if (rows > 0) {
      KernelTimer kt(d, 0);
      if (d.streamSel && aligned) launchTopMStream(d.stcfg, d.slay, ta, (int)std::min<long long>(rows, d.streamGridMax), s);
      else launchTopM(tc, ta, grid, d.topmSmem, s);
      d.launches++;
    }
    a.topTok = ta.outTok;
    a.topVal = ta.outVal;
    a.thrVal = ta.outThr;
  }
  const long long hist = (long long)Bc * (T + 2) * K;
  d.hPar.reserve(sizeof(int) * hist);
  d.hTok.reserve(sizeof(int) * hist);
  if (d.lexicon) d.hWord.reserve(sizeof(int) * hist);
  a.nCp = (T + 1) / kCpRows + 1;
  d.hSkip.reserve(sizeof(int) * (size_t)Bc * a.nCp * K);
  d.hSkipFin.reserve(sizeof(int) * (size_t)Bc * K);
  a.hSkip = d.hSkip.as<int>();
  a.hSkipFin = d.hSkipFin.as<int>();
  a.hParent = d.hPar.as<int>();
  a.hTok = d.hTok.as<int>();
  a.hWord = d.lexicon ? d.hWord.as<int>() : nullptr;
  a.finScore = d.finScore.as<double>() + outBase * K * 3;
  a.finCount = d.finCount.as<int>() + outBase;
  a.status = d.status.as<int>() + outBase;
  const int grid = std::max(1, std::min(Bc, fused ? d.fusedGridMax : d.gridMax));
  a.stats = d.counters ? d.dStats.as<unsigned long long>() : nullptr;
  if (fused) {
    KernelTimer kt(d, 3);
    launchFused(c, d.ftcfg, d.flay, a, grid, s);
    d.launches++;
  } else {
    if (d.slabBytes) {
      d.ws.reserve(d.slabBytes * grid);
      a.wsGlobal = d.ws.as<char>();
      a.wsStride = (long long)d.slabBytes;
    }
    KernelTimer kt(d, 1);
    launchDecode(c, a, grid, d.smemBytes, s, d.threads);
    d.launches++;
  }
  BacktraceArgs b{};
  b.hParent = a.hParent;
  b.hTok = a.hTok;
  b.hWord = a.hWord;
  b.hSkip = a.hSkip;
  b.hSkipFin = a.hSkipFin;
  b.nCp = a.nCp;
  b.finCount = a.finCount;
  b.lengths = dLen;
  b.B = Bc;
  b.T = T;
  b.K = K;
  b.nbest = d.nbest;
  b.outTok = d.outTok.as<int>() + outBase * d.nbest * (T + 2);
  b.outWord = d.outWord.as<int>() + outBase * d.nbest * (T + 2);
  {
    KernelTimer kt(d, 2);
    launchBacktrace(b, s);
    d.launches++;
  }
}

void prepareBatch(flt_decoder& d, int B, int T, int N) {
  if (B < 0 || T < 0 || N < 1) throw FltError(FLT_ERR_INVALID, "bad batch shape");
  planFor(d, N);
  const int K = d.cfg.K;
  d.finScore.reserve(sizeof(double) * (size_t)std::max(B, 1) * K * 3);
  d.finCount.reserve(sizeof(int) * (size_t)std::max(B, 1));
  d.status.reserve(sizeof(int) * (size_t)std::max(B, 1));
  d.outTok.reserve(sizeof(int) * (size_t)std::max(B, 1) * d.nbest * (T + 2));
  d.outWord.reserve(sizeof(int) * (size_t)std::max(B, 1) * d.nbest * (T + 2));
  d.lastB = B;
  d.lastT = T;
  d.lastNbest = d.nbest;
  d.launches = 0;
  if (d.counters) {
    d.dStats.reserve(sizeof(unsigned long long) * 32);
    rt::devZero(d.dStats.p, sizeof(unsigned long long) * 32, d.stream);
  }
#if FLT_DEVICE_BUILD
  d.evUsed = 0;
#endif
  d.evKinds.clear();
}

// device-resident emissions: whole batch in slices that bound the history / list buffers
void decodeDevice(flt_decoder& d, const float* dEmis, int B, int T, int N, const int* dLen) {
  const long long perUtt = (long long)(T + 2) * d.cfg.K * 12 + (long long)T * d.cfg.M * 8 + 64;
  long long slice = std::max<long long>(1, (8LL << 30) / perUtt);
  slice = std::min<long long>(slice, B);
  const int gm = d.fused ? d.fusedGridMax : d.gridMax;
  if (slice >= gm) slice = slice / gm * gm; // whole waves
  for (long long b0 = 0; b0 < B; b0 += slice) {
    const int Bc = (int)std::min<long long>(slice, B - b0);
    runChunk(d, dEmis + b0 * T * N, Bc, T, N, dLen ? dLen + b0 : nullptr, b0);
  }
}

template <class F>
int guarded(F&& f) {
  try {
    f();
    return FLT_OK;
  } catch (const FltError& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(FLT_ERR_RUNTIME, "out of host memory");
  } catch (const std::exception& e) {
    return fail(FLT_DEVICE_BUILD ? FLT_ERR_CUDA : FLT_ERR_RUNTIME, e.what());
  }
}

// Makes `device` current for the duration of one C-ABI call and restores the caller's current device
// afterwards (a decoder is bound to the device it was created for; the caller's context is not ours to
// change).
struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int device) {
#if FLT_DEVICE_BUILD
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      cudaGetLastError();
      throw FltError(FLT_ERR_CUDA, "no CUDA device available: the decoder has no CPU path");
    }
    if (device < 0 || device >= n) throw FltError(FLT_ERR_INVALID, "bad CUDA device ordinal");
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    if (prev != device) FLT_RT_TRY(cudaSetDevice(device));
    else prev = -1; // nothing to restore
#else
    (void)device;
#endif
  }
  ~DeviceScope() {
#if FLT_DEVICE_BUILD
    if (prev >= 0) cudaSetDevice(prev);
#endif
  }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};

flt_decoder* makeDecoder(int lexicon, const flt_options* opt, const flt_trie* trie, const flt_lm* lm,
                         int sil, int blank, int unk, const float* trans, long long nTrans,
                         int isLmToken, int device) {
  if (!opt || !lm) throw FltError(FLT_ERR_INVALID, "null options or LM");
  if (lexicon && !trie) throw FltError(FLT_ERR_INVALID, "null trie");
  DeviceScope dev(device);
  std::unique_ptr<flt_decoder> d(new flt_decoder);
  d->lexicon = lexicon;
  d->opt = *opt;
  d->trie = trie;
  d->lm = lm;
  d->sil = sil;
  d->blank = blank;
  d->unk = unk;
  if (trans && nTrans > 0) d->trans.assign(trans, trans + nTrans);
  d->isLmToken = isLmToken;
  d->device = device;
  d->nbest = std::max(1, opt->beamSize);
#if FLT_DEVICE_BUILD
  FLT_RT_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  FLT_RT_TRY(cudaStreamCreateWithFlags(&d->copyStream, cudaStreamNonBlocking));
  FLT_RT_TRY(cudaStreamCreateWithFlags(&d->copyStream2, cudaStreamNonBlocking));
  for (int i = 0; i < 3; ++i) {
    FLT_RT_TRY(cudaEventCreateWithFlags(&d->evCopied[i], cudaEventDisableTiming));
    FLT_RT_TRY(cudaEventCreateWithFlags(&d->evFree[i], cudaEventDisableTiming));
  }
#endif
  return d.release();
}

void checkStatus(flt_decoder& d) {
  std::vector<int> st(std::max(d.lastB, 1));
  rt::d2h(st.data(), d.status.p, sizeof(int) * d.lastB, d.stream);
  rt::sync(d.stream);
  int bits = 0;
  for (int b = 0; b < d.lastB; ++b) bits |= st[b];
  if (bits & 1) throw FltError(FLT_ERR_RUNTIME, "candidate capacity exceeded");
}

} // namespace

namespace {

// One launch of online decoding: T frames of one utterance (T = 0 with finish = decodeEnd only).
// Appends the new history rows to the host mirror.
// Returns false (nothing mirrored, the saved beam untouched) if a frame overflowed the candidate
// capacity; runStream then grows the capacity and repeats the chunk.
bool runStreamOnce(flt_decoder& d, const float* emis, int T, int N, bool finish) {
  planFor(d, N);
  const DecCfg& c = d.cfg;
  const int K = c.K;
  rt::Stream s = d.stream;
  const float* dEmis = emis;
  if (T > 0 && !rt::isDevicePtr(emis)) {
    d.sEmis.reserve(sizeof(float) * (size_t)T * N);
    rt::h2d(d.sEmis.p, emis, sizeof(float) * (size_t)T * N, s);
    dEmis = d.sEmis.as<float>();
  }
  BatchArgs a{};
  a.emis = dEmis;
  a.B = 1;
  a.T = T;
  a.lengths = nullptr;
  if (d.needTopM && T > 0) {
    d.topTok.reserve(sizeof(int) * (size_t)T * c.M);
    d.topVal.reserve(sizeof(float) * (size_t)T * c.M);
    if (!c.setAll) d.thr.reserve(sizeof(float) * (size_t)T);
    TopMArgs ta{};
    ta.emis = dEmis;
    ta.rows = T;
    ta.outTok = d.topTok.as<int>();
    ta.outVal = d.topVal.as<float>();
    ta.outThr = c.setAll ? nullptr : d.thr.as<float>();
    TopMCfg tc = d.tcfg;
    const bool aligned = (reinterpret_cast<uintptr_t>(dEmis) & 15) == 0;
    if (!aligned) tc.fast = 0;
    if (d.streamSel && aligned) launchTopMStream(d.stcfg, d.slay, ta, std::min(T, d.streamGridMax), s);
    else launchTopM(tc, ta, std::min(T, d.topmGridMax), d.topmSmem, s);
    a.topTok = ta.outTok;
    a.topVal = ta.outVal;
    a.thrVal = ta.outThr;
  }
  const size_t rows = (size_t)T + 2;
  d.hPar.reserve(sizeof(int) * rows * K);
  d.hTok.reserve(sizeof(int) * rows * K);
  d.hWord.reserve(sizeof(int) * rows * K);
  d.sScore.reserve(sizeof(double) * rows * K * 3);
  d.sCount.reserve(sizeof(int) * rows);
  d.sBeam.reserve(streamBeamBytes(c));
  d.finScore.reserve(sizeof(double) * (size_t)K * 3);
  d.finCount.reserve(sizeof(int));
  d.status.reserve(sizeof(int));
  a.nCp = (T + 1) / kCpRows + 1;
  d.hSkip.reserve(sizeof(int) * (size_t)a.nCp * K);
  d.hSkipFin.reserve(sizeof(int) * (size_t)K);
  rt::devZero(d.sCount.p, sizeof(int) * rows, s);
  a.hSkip = d.hSkip.as<int>();
  a.hSkipFin = d.hSkipFin.as<int>();
  a.hParent = d.hPar.as<int>();
  a.hTok = d.hTok.as<int>();
  a.hWord = d.lexicon ? d.hWord.as<int>() : nullptr;
  a.finScore = d.finScore.as<double>();
  a.finCount = d.finCount.as<int>();
  a.status = d.status.as<int>();
  a.stats = nullptr;
  // the launch overwrites the saved beam: keep a copy so that an overflowed chunk can be repeated
  if (d.on.haveBeam) {
    d.sBeamBackup.reserve(streamBeamBytes(c));
    rt::d2d(d.sBeamBackup.p, d.sBeam.p, streamBeamBytes(c), s);
  }
  a.streamBeam = d.sBeam.as<char>();
  a.streamRestore = d.on.haveBeam ? 1 : 0;
  a.streamNoFinish = finish ? 0 : 1;
  a.streamFrame0 = d.on.nDecoded;
  a.streamShift = d.on.shift;
  a.hScore = d.sScore.as<double>();
  a.hCount = d.sCount.as<int>();
  if (d.slabBytes) {
    d.ws.reserve(d.slabBytes);
    a.wsGlobal = d.ws.as<char>();
    a.wsStride = (long long)d.slabBytes;
  }
  launchDecode(c, a, 1, d.smemBytes, s, d.threads);
  // mirror the new rows: frames 1..T (+ the finish row T+1)
  const int nNew = T + (finish ? 1 : 0);
  std::vector<int> par(rows * K), tok(rows * K), wrd(rows * K, -1), cnt(rows), st(1);
  std::vector<double> sc(rows * K * 3);
  rt::d2h(par.data(), d.hPar.p, sizeof(int) * rows * K, s);
  rt::d2h(tok.data(), d.hTok.p, sizeof(int) * rows * K, s);
  if (d.lexicon) rt::d2h(wrd.data(), d.hWord.p, sizeof(int) * rows * K, s);
  rt::d2h(sc.data(), d.sScore.p, sizeof(double) * rows * K * 3, s);
  rt::d2h(cnt.data(), d.sCount.p, sizeof(int) * rows, s);
  rt::d2h(st.data(), d.status.p, sizeof(int), s);
  rt::sync(s);
  if (st[0] & 1) {
    if (d.on.haveBeam) rt::d2d(d.sBeam.p, d.sBeamBackup.p, streamBeamBytes(c), s);
    rt::sync(s);
    return false;
  }
  for (int r = 1; r <= nNew; ++r) {
    std::vector<flt_decoder::SHyp> row(cnt[r]);
    for (int q = 0; q < cnt[r]; ++q) {
      const size_t o = (size_t)r * K + q;
      row[q] = flt_decoder::SHyp{sc[o * 3], sc[o * 3 + 1], sc[o * 3 + 2], par[o], tok[o], wrd[o]};
    }
    d.on.hyp.push_back(std::move(row));
  }
  d.on.haveBeam = true;
  d.on.shift = 0.0;
  d.on.nDecoded += nNew;
  return true;
}

void runStream(flt_decoder& d, const float* emis, int T, int N, bool finish) {
  for (int attempt = 0;; ++attempt) {
    if (runStreamOnce(d, emis, T, N, finish)) return;
    // data-dependent candidate counts (lexicon enumeration, full expansion): grow and redo the chunk
    if (attempt >= 6) throw FltError(FLT_ERR_RUNTIME, "candidate capacity exceeded");
    d.capBoost *= 4;
    d.planN = -1;
  }
}

using SHyp = flt_decoder::SHyp;
constexpr int kLookBackLimit = 100; // Utils.h:28

bool onlineComplete(const flt_decoder& d, int frame, int k) { // LexiconDecoder.h:97-99
  if (!d.lexicon) return true;
  const SHyp& h = d.on.hyp[frame][k];
  return h.parent < 0 || d.on.hyp[frame - 1][h.parent].word >= 0;
}
// Utils.h:268-310: (frame, index) of the ancestor or index -1; lookBack is updated
std::pair<int, int> onlineBestAncestor(const flt_decoder& d, int finalFrame, int& lookBack) {
  const auto& fin = d.on.hyp[finalFrame];
  if (fin.empty()) return {-1, -1};
  int bk = 0;
  for (int r = 1; r < (int)fin.size(); ++r)
    if (fin[r].score > fin[bk].score) bk = r;
  int f = finalFrame, k = bk, n = 0;
  auto up = [&]() {
    k = d.on.hyp[f][k].parent;
    --f;
    if (k < 0) f = -1;
  };
  while (k >= 0 && n < lookBack) {
    ++n;
    up();
  }
  const int maxLB = lookBack + kLookBackLimit;
  while (k >= 0) {
    if (onlineComplete(d, f, k)) break;
    ++n;
    up();
    if (n == maxLB) break;
  }
  lookBack = n;
  return {k >= 0 ? f : -1, k};
}
// Utils.h:229-250
int onlineFill(const flt_decoder& d, int frame, int k, int finalFrame, int stride, double* scores3,
               int32_t* tokens, int32_t* words) {
  const SHyp& h0 = d.on.hyp[frame][k];
  if (scores3) {
    scores3[0] = h0.score;
    scores3[1] = h0.am;
    scores3[2] = h0.lm;
  }
  for (int j = 0; j <= finalFrame && j < stride; ++j) {
    if (tokens) tokens[j] = -1;
    if (words) words[j] = -1;
  }
  int i = 0, f = frame;
  while (k >= 0 && f >= 0) {
    const SHyp& h = d.on.hyp[f][k];
    const int pos = finalFrame - i;
    if (pos >= 0 && pos < stride) {
      if (tokens) tokens[pos] = h.token;
      if (words) words[pos] = d.lexicon ? h.word : -1;
    }
    k = h.parent;
    --f;
    ++i;
  }
  return finalFrame + 1;
}
void requireOnline(const flt_decoder* d) {
  if (!d) throw FltError(FLT_ERR_INVALID, "null decoder");
  if (!d->on.begun) throw FltError(FLT_ERR_INVALID, "flt_stream_begin has not been called");
}

} // namespace

/* ================================================================================ C ABI ===== */
extern "C" {

const char* flt_last_error(void) { return gErr.c_str(); }

int flt_trie_create(int32_t maxChildren, int32_t rootIdx, flt_trie** out) {
  return guarded([&] {
    if (!out) throw FltError(FLT_ERR_INVALID, "null out");
    auto* t = new flt_trie;
    t->maxChildren = maxChildren;
    t->rootIdx = rootIdx;
    t->nodes.emplace_back();
    *out = t;
  });
}
int flt_trie_insert(flt_trie* trie, const int32_t* indices, int32_t n, int32_t label, float score) {
  return guarded([&] {
    if (!trie) throw FltError(FLT_ERR_INVALID, "null trie");
    if (trie->frozen) throw FltError(FLT_ERR_INVALID, "trie is frozen: a decoder already uses it");
    int cur = 0;
    for (int i = 0; i < n; ++i) {
      const int idx = indices[i];
      if (idx < 0 || idx >= trie->maxChildren) // Trie.cpp:31-34 (nodes created so far stay)
        throw FltError(FLT_ERR_OUT_OF_RANGE, "[Trie] Invalid letter index: " + std::to_string(idx));
      auto it = trie->nodes[cur].kids.find(idx);
      if (it == trie->nodes[cur].kids.end()) {
        const int nn = (int)trie->nodes.size();
        trie->nodes[cur].kids.emplace(idx, nn);
        trie->nodes.emplace_back();
        cur = nn;
      } else {
        cur = it->second;
      }
    }
    if ((int)trie->nodes[cur].labels.size() < kTrieMaxLabel) { // Trie.cpp:40-46
      trie->nodes[cur].labels.push_back(label);
      trie->nodes[cur].scores.push_back(score);
    } else {
      fprintf(stderr, "[Trie] Trie label number reached limit: %d\n", kTrieMaxLabel);
    }
  });
}
int flt_trie_smear(flt_trie* trie, int32_t mode) {
  return guarded([&] {
    if (!trie) throw FltError(FLT_ERR_INVALID, "null trie");
    if (trie->frozen) throw FltError(FLT_ERR_INVALID, "trie is frozen: a decoder already uses it");
    if (mode != FLT_SMEAR_NONE) trie->smearNode(0, mode);
  });
}
int flt_trie_search(const flt_trie* trie, const int32_t* indices, int32_t n, int32_t* found,
                    float* maxScore, int32_t* nLabels, int32_t* labels6, float* scores6) {
  return guarded([&] {
    if (!trie || !found) throw FltError(FLT_ERR_INVALID, "null argument");
    int cur = 0;
    *found = 0;
    for (int i = 0; i < n; ++i) {
      const int idx = indices[i];
      if (idx < 0 || idx >= trie->maxChildren)
        throw FltError(FLT_ERR_OUT_OF_RANGE, "[Trie] Invalid letter index: " + std::to_string(idx));
      auto it = trie->nodes[cur].kids.find(idx);
      if (it == trie->nodes[cur].kids.end()) return;
      cur = it->second;
    }
    const HNode& nd = trie->nodes[cur];
    *found = 1;
    if (maxScore) *maxScore = nd.maxScore;
    if (nLabels) *nLabels = (int)nd.labels.size();
    for (size_t i = 0; i < nd.labels.size() && i < 6; ++i) {
      if (labels6) labels6[i] = nd.labels[i];
      if (scores6) scores6[i] = nd.scores[i];
    }
  });
}
int flt_trie_num_nodes(const flt_trie* trie, int64_t* out) {
  return guarded([&] {
    if (!trie || !out) throw FltError(FLT_ERR_INVALID, "null argument");
    *out = (int64_t)trie->nodes.size();
  });
}
int flt_trie_max_scores(const flt_trie* trie, float* out, int64_t n) {
  return guarded([&] {
    if (!trie || (!out && n > 0)) throw FltError(FLT_ERR_INVALID, "null argument");
    if (n != (int64_t)trie->nodes.size()) throw FltError(FLT_ERR_INVALID, "n must equal flt_trie_num_nodes");
    for (int64_t i = 0; i < n; ++i) out[i] = trie->nodes[(size_t)i].maxScore;
  });
}
void flt_trie_destroy(flt_trie* trie) { delete trie; }
int flt_trie_save(const flt_trie* trie, const char* path) {
  return guarded([&] {
    if (!trie || !path) throw FltError(FLT_ERR_INVALID, "null argument");
    saveTrie(*trie, path);
  });
}
int flt_trie_export(const flt_trie* trie, int32_t* meta5, int32_t* childOff, int32_t* childTok, int32_t* childNode,
                    int32_t* labelOff, int32_t* labels, float* scores, float* maxScore) {
  return guarded([&] {
    if (!trie || !meta5) throw FltError(FLT_ERR_INVALID, "null argument");
    const size_t nn = trie->nodes.size();
    size_t ne = 0, nl = 0;
    for (const HNode& nd : trie->nodes) ne += nd.kids.size(), nl += nd.labels.size();
    meta5[0] = trie->maxChildren, meta5[1] = trie->rootIdx, meta5[2] = (int32_t)nn, meta5[3] = (int32_t)ne,
    meta5[4] = (int32_t)nl;
    if (!childOff) return; // sizes only
    if (!childTok || !childNode || !labelOff || !labels || !scores || !maxScore)
      throw FltError(FLT_ERR_INVALID, "null argument");
    size_t e = 0, l = 0;
    for (size_t i = 0; i < nn; ++i) {
      const HNode& nd = trie->nodes[i];
      childOff[i] = (int32_t)e, labelOff[i] = (int32_t)l, maxScore[i] = nd.maxScore;
      for (auto& kv : nd.kids) childTok[e] = kv.first, childNode[e] = kv.second, ++e;
      for (size_t k = 0; k < nd.labels.size(); ++k) labels[l] = nd.labels[k], scores[l] = nd.scores[k], ++l;
    }
    childOff[nn] = (int32_t)e, labelOff[nn] = (int32_t)l;
  });
}
int flt_trie_load(const char* path, flt_trie** out) {
  return guarded([&] {
    if (!out || !path) throw FltError(FLT_ERR_INVALID, "null argument");
    std::unique_ptr<flt_trie> t(new flt_trie);
    loadTrie(path, *t);
    *out = t.release();
  });
}

int flt_lm_zero_create(flt_lm** out) {
  return guarded([&] {
    if (!out) throw FltError(FLT_ERR_INVALID, "null out");
    auto* m = new flt_lm;
    m->kind = 0;
    m->makeHostView();
    *out = m;
  });
}
int flt_lm_ngram_load_arpa(const char* path, const char* const* usrWords, int32_t nUsrWords,
                           flt_lm** out) {
  return guarded([&] {
    if (!out || !path) throw FltError(FLT_ERR_INVALID, "null argument");
    std::unique_ptr<flt_lm> m(new flt_lm);
    // like KenLM's constructor, which takes ARPA text or its own binary format and tells them apart by magic
    if (isTableFile(path)) loadLmTables(path, usrWords, nUsrWords, *m);
    else loadArpa(path, usrWords, nUsrWords, *m);
    *out = m.release();
  });
}
int flt_lm_save(const flt_lm* lm, const char* path) {
  return guarded([&] {
    if (!lm || !path) throw FltError(FLT_ERR_INVALID, "null argument");
    saveLm(*lm, path);
  });
}
int flt_lm_score_seq(const flt_lm* lm, const int32_t* usrIdx, int32_t n, int32_t withFinish,
                     float* out) {
  return guarded([&] {
    if (!lm || !out) throw FltError(FLT_ERR_INVALID, "null argument");
    if (lm->kind == 0) {
      for (int i = 0; i < n + (withFinish ? 1 : 0); ++i) out[i] = 0.0f;
      return;
    }
    int ctx[kMaxCtx], nctx = 0, nxt[kMaxCtx];
    if (lm->order > 1) {
      ctx[0] = lm->bos;
      nctx = 1;
    }
    for (int i = 0; i < n + (withFinish ? 1 : 0); ++i) {
      int w;
      if (i < n) {
        if (usrIdx[i] < 0 || usrIdx[i] >= (int)lm->usr2lm.size()) // lm/KenLM.cpp:66-69
          throw FltError(FLT_ERR_RUNTIME, "[KenLM] Invalid user token index: " + std::to_string(usrIdx[i]));
        w = lm->usr2lm[usrIdx[i]];
      } else {
        w = lm->eos;
      }
      out[i] = ngramScore(lm->host, ctx, nctx, w);
      nctx = ngramAdvanceCtx(lm->host, ctx, nctx, w, nxt);
      for (int k = 0; k < nctx; ++k) ctx[k] = nxt[k];
    }
  });
}
void flt_lm_destroy(flt_lm* lm) { delete lm; }

int flt_decoder_create_lexfree(const flt_options* opt, const flt_lm* lm, int32_t sil, int32_t blank,
                               const float* transitions, int64_t nTransitions, int32_t device,
                               flt_decoder** out) {
  return guarded([&] {
    if (!out) throw FltError(FLT_ERR_INVALID, "null out");
    *out = makeDecoder(0, opt, nullptr, lm, sil, blank, -1, transitions, nTransitions, 0, device);
  });
}
int flt_decoder_create_lexicon(const flt_options* opt, const flt_trie* trie, const flt_lm* lm,
                               int32_t sil, int32_t blank, int32_t unk, const float* transitions,
                               int64_t nTransitions, int32_t isLmToken, int32_t device,
                               flt_decoder** out) {
  return guarded([&] {
    if (!out) throw FltError(FLT_ERR_INVALID, "null out");
    *out = makeDecoder(1, opt, trie, lm, sil, blank, unk, transitions, nTransitions, isLmToken, device);
  });
}
void flt_decoder_destroy(flt_decoder* dec) { delete dec; }

int flt_decoder_set_nbest(flt_decoder* dec, int32_t nbest) {
  return guarded([&] {
    if (!dec || nbest < 1) throw FltError(FLT_ERR_INVALID, "nbest must be >= 1");
    dec->nbest = std::min(nbest, dec->opt.beamSize);
  });
}

int flt_decode_batch_async(flt_decoder* dec, const float* dEmissions, int32_t B, int32_t T,
                           int32_t N, const int32_t* dLengths) {
  return guarded([&] {
    if (!dec || (!dEmissions && (long long)B * T > 0)) throw FltError(FLT_ERR_INVALID, "null argument");
    DeviceScope dev(dec->device);
    prepareBatch(*dec, B, T, N);
    dec->haveLengths = false;
    decodeDevice(*dec, dEmissions, B, T, N, dLengths);
  });
}

int flt_decode_batch(flt_decoder* dec, const float* emissions, int32_t B, int32_t T, int32_t N,
                     const int32_t* lengths) {
  return guarded([&] {
    if (!dec || (!emissions && (long long)B * T > 0)) throw FltError(FLT_ERR_INVALID, "null argument");
    DeviceScope dev(dec->device);
    flt_decoder& d = *dec;
    for (int attempt = 0;; ++attempt) {
      prepareBatch(d, B, T, N);
      const int* dLen = nullptr;
      if (lengths) {
        for (int b = 0; b < B; ++b)
          if (lengths[b] < 0 || lengths[b] > T) throw FltError(FLT_ERR_INVALID, "lengths[b] out of [0,T]");
        d.dLengths.reserve(sizeof(int) * (size_t)std::max(B, 1));
        rt::h2d(d.dLengths.p, lengths, sizeof(int) * B, d.stream);
        dLen = d.dLengths.as<int>();
      }
      const bool onDevice = (long long)B * T == 0 || rt::isDevicePtr(emissions);
      if (onDevice) {
        decodeDevice(d, emissions, B, T, N, dLen);
      } else {
#if FLT_DEVICE_BUILD
        // host emissions: stream them through staging buffers so the PCIe copy of slice i+1 overlaps the
        // kernels of slice i. Three buffers keep the copy queue from ever waiting for the kernels to free one
        // (two copy queues were measured slower: concurrent copies share the link and every slice arrives
        // later). The step is latency-bound — a slice of 13 utterances costs as much device
        // time as one of 52 — so FEW LARGE slices (2 GiB) keep the kernels hidden behind the copies, and the
        // part that cannot overlap is one step (the last slice's) either way. Measured at cfg 2 / cfg 3:
        // 512 MiB slices on two queues 1.10 k / 0.48 k utt/s, 2 GiB on two queues 1.15 k / 1.15 k,
        // 1 GiB on one queue (two buffers) 1.21 k / 0.83 k.
        const long long perUtt = (long long)T * N * sizeof(float);
        long long slice = std::max<long long>(1, (2LL << 30) / std::max<long long>(perUtt, 1));
        // ... and by the same history / list budget as decodeDevice (small N with a large beam)
        const long long perUttHist = (long long)(T + 2) * d.cfg.K * 12 + (long long)T * d.cfg.M * 8 + 64;
        slice = std::min<long long>(slice, std::max<long long>(1, (8LL << 30) / perUttHist));
        slice = std::min<long long>(slice, B);
        for (int i = 0; i < 3; ++i) d.staging[i].reserve((size_t)(slice * perUtt));
        int k = 0;
        for (long long b0 = 0; b0 < B; b0 += slice, k = (k + 1) % 3) {
          const int Bc = (int)std::min<long long>(slice, B - b0);
          rt::Stream cs = d.copyStream; // ONE queue: slices must arrive in order, each as early as possible
          FLT_RT_TRY(cudaStreamWaitEvent(cs, d.evFree[k], 0));
          FLT_RT_TRY(cudaMemcpyAsync(d.staging[k].p, emissions + b0 * T * N, (size_t)(Bc * perUtt),
                                     cudaMemcpyHostToDevice, cs));
          FLT_RT_TRY(cudaEventRecord(d.evCopied[k], cs));
          FLT_RT_TRY(cudaStreamWaitEvent(d.stream, d.evCopied[k], 0));
          runChunk(d, d.staging[k].as<float>(), Bc, T, N, dLen ? dLen + b0 : nullptr, b0);
          FLT_RT_TRY(cudaEventRecord(d.evFree[k], d.stream));
        }
#else
        decodeDevice(d, emissions, B, T, N, dLen);
#endif
      }
      try {
        checkStatus(d);
        break;
      } catch (const FltError& e) {
        // the lexicon decoder's direct enumeration is data dependent: grow and redo the batch
        if (std::string(e.what()) != "candidate capacity exceeded" || attempt >= 6) throw;
        d.capBoost *= 4;
        d.planN = -1;
      }
    }
    d.haveLengths = lengths != nullptr;
    if (lengths) d.lastLengths.assign(lengths, lengths + B);
  });
}

int flt_decoder_synchronize(flt_decoder* dec) {
  return guarded([&] {
    if (!dec) throw FltError(FLT_ERR_INVALID, "null decoder");
    rt::sync(dec->stream);
  });
}
void* flt_decoder_stream(flt_decoder* dec) { return dec ? (void*)dec->stream : nullptr; }

int flt_nbest_copy(flt_decoder* dec, int32_t nbest, int32_t* tokens, int32_t* words, double* scores,
                   int32_t* counts) {
  return guarded([&] {
    if (!dec) throw FltError(FLT_ERR_INVALID, "null decoder");
    flt_decoder& d = *dec;
    DeviceScope dev(d.device);
    const int B = d.lastB, T = d.lastT, K = d.cfg.K;
    // the rows of the last batch were sized and strided with the nbest setting of THAT decode, whatever
    // flt_decoder_set_nbest has been told since
    const int NB = d.lastNbest;
    if (B == 0) return;
    if (nbest < 1 || nbest > NB)
      throw FltError(FLT_ERR_INVALID, "nbest exceeds the nbest setting the last batch was decoded with");
    try {
      checkStatus(d);
    } catch (const FltError& e) {
      // flt_decode_batch_async cannot redo the batch itself: grow the capacity for the caller's retry
      if (std::string(e.what()) == "candidate capacity exceeded" && d.capBoost < 4096) {
        d.capBoost *= 4;
        d.planN = -1;
      }
      throw;
    }
    const size_t L = (size_t)T + 2;
    if (nbest == NB) {
      if (tokens) rt::d2h(tokens, d.outTok.p, sizeof(int) * B * nbest * L, d.stream);
      if (words) rt::d2h(words, d.outWord.p, sizeof(int) * B * nbest * L, d.stream);
    } else {
      for (int b = 0; b < B; ++b) {
        if (tokens) rt::d2h(tokens + (size_t)b * nbest * L, d.outTok.as<int>() + (size_t)b * NB * L,
                            sizeof(int) * nbest * L, d.stream);
        if (words) rt::d2h(words + (size_t)b * nbest * L, d.outWord.as<int>() + (size_t)b * NB * L,
                           sizeof(int) * nbest * L, d.stream);
      }
    }
    if (scores) {
      if (nbest == K) {
        rt::d2h(scores, d.finScore.p, sizeof(double) * B * K * 3, d.stream);
      } else {
        for (int b = 0; b < B; ++b)
          rt::d2h(scores + (size_t)b * nbest * 3, d.finScore.as<double>() + (size_t)b * K * 3,
                  sizeof(double) * nbest * 3, d.stream);
      }
    }
    if (counts) rt::d2h(counts, d.finCount.p, sizeof(int) * B, d.stream);
    rt::sync(d.stream);
  });
}

int flt_nbest_device_ptrs(flt_decoder* dec, int32_t** tokens, int32_t** words, double** scores,
                          int32_t** counts) {
  return guarded([&] {
    if (!dec) throw FltError(FLT_ERR_INVALID, "null decoder");
    if (tokens) *tokens = dec->outTok.as<int32_t>();
    if (words) *words = dec->outWord.as<int32_t>();
    if (scores) *scores = dec->finScore.as<double>();
    if (counts) *counts = dec->finCount.as<int32_t>();
  });
}

int flt_stream_begin(flt_decoder* dec, int32_t N) {
  return guarded([&] {
    if (!dec) throw FltError(FLT_ERR_INVALID, "null decoder");
    DeviceScope dev(dec->device);
    planFor(*dec, N);
    auto& on = dec->on;
    on = flt_decoder::Online{};
    on.begun = true;
    on.N = N;
    on.hyp.push_back({SHyp{0.0, 0.0, 0.0, -1, dec->sil, -1}}); // LexiconDecoder.cpp:26-27
  });
}
int flt_stream_step(flt_decoder* dec, const float* emissions, int32_t T, int32_t N) {
  return guarded([&] {
    requireOnline(dec);
    if (N != dec->on.N) throw FltError(FLT_ERR_INVALID, "N differs from flt_stream_begin");
    if (T < 0 || (!emissions && T > 0)) throw FltError(FLT_ERR_INVALID, "bad chunk");
    DeviceScope dev(dec->device);
    if (T > 0) runStream(*dec, emissions, T, N, false);
  });
}
int flt_stream_end(flt_decoder* dec) {
  return guarded([&] {
    requireOnline(dec);
    DeviceScope dev(dec->device);
    runStream(*dec, nullptr, 0, dec->on.N, true);
  });
}
int flt_stream_prune(flt_decoder* dec, int32_t lookBack) {
  return guarded([&] {
    requireOnline(dec);
    auto& on = dec->on;
    if (on.nDecoded - on.nPruned - lookBack < 1) return; // LexiconDecoder.cpp:304-309
    const int finalFrame = on.nDecoded - on.nPruned;
    int lb = lookBack;
    if (onlineBestAncestor(*dec, finalFrame, lb).second < 0) return;
    lookBack = lb;
    const int startFrame = on.nDecoded - on.nPruned - lookBack;
    if (startFrame < 1) return;
    // Utils.h:312-342: keep rows [startFrame, startFrame + lookBack], orphan row 0, normalise the
    // newest row (= the device beam: applied when the next launch restores it)
    std::vector<std::vector<SHyp>> kept(on.hyp.begin() + startFrame, on.hyp.begin() + startFrame + lookBack + 1);
    on.hyp.swap(kept);
    for (SHyp& h : on.hyp[0]) h.parent = -1;
    if (!on.hyp[lookBack].empty()) {
      double largest = on.hyp[lookBack].front().score;
      for (const SHyp& h : on.hyp[lookBack]) largest = std::max(largest, h.score);
      for (SHyp& h : on.hyp[lookBack]) h.score -= largest;
      if (lookBack == (int)on.hyp.size() - 1) on.shift += largest;
    }
    on.nPruned = on.nDecoded - lookBack;
  });
}
int flt_stream_frames_in_buffer(flt_decoder* dec, int32_t* out) {
  return guarded([&] {
    requireOnline(dec);
    if (!out) throw FltError(FLT_ERR_INVALID, "null out");
    *out = dec->on.nDecoded - dec->on.nPruned + 1; // LexiconDecoder.cpp:300-302
  });
}
int flt_stream_n_hypothesis(flt_decoder* dec, int32_t* out) {
  return guarded([&] {
    requireOnline(dec);
    if (!out) throw FltError(FLT_ERR_INVALID, "null out");
    const int finalFrame = dec->on.nDecoded - dec->on.nPruned;
    *out = (int)dec->on.hyp[finalFrame].size();
  });
}
int flt_stream_best(flt_decoder* dec, int32_t lookBack, int32_t maxLen, int32_t* tokens, int32_t* words,
                    double* scores3, int32_t* len) {
  return guarded([&] {
    requireOnline(dec);
    if (!len) throw FltError(FLT_ERR_INVALID, "null len");
    *len = 0;
    const int finalFrame = dec->on.nDecoded - dec->on.nPruned;
    if (dec->lexicon && finalFrame - lookBack < 1) return; // LexiconDecoder.cpp:285-288
    int lb = lookBack;
    const auto anc = onlineBestAncestor(*dec, finalFrame, lb);
    if (anc.second < 0) return;
    *len = onlineFill(*dec, anc.first, anc.second, finalFrame - lb, maxLen, scores3, tokens, words);
  });
}
int flt_stream_all_final(flt_decoder* dec, int32_t maxHyp, int32_t maxLen, int32_t* tokens, int32_t* words,
                         double* scores3, int32_t* lens, int32_t* count) {
  return guarded([&] {
    requireOnline(dec);
    if (!count) throw FltError(FLT_ERR_INVALID, "null count");
    const int finalFrame = dec->on.nDecoded - dec->on.nPruned;
    *count = 0;
    if (finalFrame < 1) return; // LexiconDecoder.cpp:276-283
    const auto& fin = dec->on.hyp[finalFrame];
    for (int r = 0; r < (int)fin.size() && r < maxHyp; ++r) {
      const int L = onlineFill(*dec, finalFrame, r, finalFrame, maxLen, scores3 ? scores3 + 3 * r : nullptr,
                               tokens ? tokens + (size_t)r * maxLen : nullptr,
                               words ? words + (size_t)r * maxLen : nullptr);
      if (lens) lens[r] = L;
    }
    *count = (int)fin.size();
  });
}

int flt_decoder_last_launches(const flt_decoder* dec, int32_t* out) {
  return guarded([&] {
    if (!dec || !out) throw FltError(FLT_ERR_INVALID, "null argument");
    *out = dec->launches;
  });
}
int flt_decoder_set_timing(flt_decoder* dec, int32_t on) {
  return guarded([&] {
    if (!dec) throw FltError(FLT_ERR_INVALID, "null decoder");
    dec->timing = on != 0;
    dec->counters = on >= 2;
  });
}
int flt_decoder_last_kernel_ms(flt_decoder* dec, float* ms3, int32_t* launches3) {
  return guarded([&] {
    if (!dec || !ms3) throw FltError(FLT_ERR_INVALID, "null argument");
    for (int k = 0; k < 4; ++k) {
      ms3[k] = 0;
      if (launches3) launches3[k] = 0;
    }
#if FLT_DEVICE_BUILD
    rt::sync(dec->stream);
    for (size_t i = 0; i < dec->evKinds.size(); ++i) {
      float ms = 0;
      FLT_RT_TRY(cudaEventElapsedTime(&ms, dec->evPool[2 * i], dec->evPool[2 * i + 1]));
      ms3[dec->evKinds[i]] += ms;
      if (launches3) launches3[dec->evKinds[i]]++;
    }
#endif
  });
}
int flt_decoder_last_stats(flt_decoder* dec, uint64_t* out4) {
  return guarded([&] {
    if (!dec || !out4) throw FltError(FLT_ERR_INVALID, "null argument");
    for (int k = 0; k < 32; ++k) out4[k] = 0;
    if (!dec->dStats.p) return;
    rt::d2h(out4, dec->dStats.p, sizeof(uint64_t) * 32, dec->stream);
    rt::sync(dec->stream);
  });
}
int flt_decoder_workspace_bytes(const flt_decoder* dec, int64_t* out) {
  return guarded([&] {
    if (!dec || !out) throw FltError(FLT_ERR_INVALID, "null argument");
    int64_t t = 0;
    for (const rt::DevBuf* b : {&dec->topTok, &dec->topVal, &dec->thr, &dec->hPar, &dec->hTok,
                                &dec->hWord, &dec->finScore, &dec->finCount, &dec->status, &dec->ws,
                                &dec->outTok, &dec->outWord, &dec->staging[0], &dec->staging[1], &dec->staging[2]})
      t += (int64_t)b->cap;
    *out = t;
  });
}

int flt_topm_rows(const float* dEmissions, int64_t rows, int32_t N, int32_t M, int32_t* dTok,
                  float* dVal, void* stream) {
  return flt_topm_rows_bias(dEmissions, rows, N, M, nullptr, dTok, dVal, stream);
}

int flt_topm_rows_bias(const float* dEmissions, int64_t rows, int32_t N, int32_t M, const float* dBias,
                       int32_t* dTok, float* dVal, void* stream) {
  return guarded([&] {
    if (M < 1 || M > 2048 || M > N) throw FltError(FLT_ERR_INVALID, "need 1 <= M <= min(N, 2048)");
    {
      TopMCfg st;
      StreamLay sl;
      if ((reinterpret_cast<uintptr_t>(dEmissions) & 15) == 0 && (reinterpret_cast<uintptr_t>(dBias) & 15) == 0 &&
          planStream(N, M, N, dBias, 0.0f, st, sl)) {
        TopMArgs sa{};
        sa.emis = dEmissions;
        sa.rows = rows;
        sa.outTok = dTok;
        sa.outVal = dVal;
        sa.outThr = nullptr;
        int gridS = 3;
#if FLT_DEVICE_BUILD
        int dev = 0, sms = 148, occ = 1;
        FLT_RT_TRY(cudaGetDevice(&dev));
        FLT_RT_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        auto* ks = sl.threads == kStreamThreads ? flt_k_topm_stream : flt_k_topm_stream256;
        FLT_RT_TRY(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(sl.total, 48 * 1024)));
        FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ks, sl.threads, sl.total));
        gridS = std::max(1, occ) * sms;
#endif
        if (rows > 0) launchTopMStream(st, sl, sa, (int)std::min<int64_t>(rows, gridS), (rt::Stream)stream);
        return;
      }
    }
    TopMCfg t{};
    t.N = N;
    t.M = M;
    t.bst = N;
    t.bias = dBias;
    t.capS = 2048;
    t.P = std::max(kThreads, nextPow2(M));
    t.fast = (N % 4 == 0) && N <= 4 * kFastVec * kThreads && M <= 256 &&
             (reinterpret_cast<uintptr_t>(dEmissions) & 15) == 0 && (reinterpret_cast<uintptr_t>(dBias) & 15) == 0;
    t.stage = !t.fast && (size_t)N * 4 <= 100 * 1024;
    TopMSmem ts;
    const size_t smem = (carveTopM(nullptr, t, ts) + 255) / 256 * 256;
    TopMArgs a{};
    a.emis = dEmissions;
    a.rows = rows;
    a.outTok = dTok;
    a.outVal = dVal;
    a.outThr = nullptr;
    int gridMax = 4;
#if FLT_DEVICE_BUILD
    int dev = 0, sms = 148, occ = 1;
    FLT_RT_TRY(cudaGetDevice(&dev));
    FLT_RT_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FLT_RT_TRY(cudaFuncSetAttribute(flt_k_topm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)std::max<size_t>(smem, 48 * 1024)));
    FLT_RT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, flt_k_topm, kThreads, smem));
    gridMax = std::max(1, occ) * sms * 8;
#endif
    if (rows > 0) launchTopM(t, a, (int)std::min<int64_t>(rows, gridMax), smem, (rt::Stream)stream);
  });
}

} // extern "C"
