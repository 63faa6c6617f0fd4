// table_io.h — on-disk format of the flattened decoder tables (Trie and n-gram LM).
//
// The reference rebuilds its tables in every process: 200 k Trie::insert calls plus smear
// (test/decoder/DecoderTest.cpp:126-146) and a KenLM load of the ARPA text (lm/KenLM.cpp:32-47; KenLM's own
// answer to that is its binary format, which it detects by magic in the same constructor). This file is the
// same idea for the device tables: one little-endian file per object, written once, read back with a
// handful of freads straight into the vectors the device image is uploaded from.
//
//   file   = header, section*
//   header = magic "FLTTBL\0\1" (8 bytes), kind u32 (1 = Trie, 2 = n-gram LM), nSections u32
//   section= tag u32, elemBytes u32, count u64, payload (count * elemBytes bytes), zero padding to 8 bytes
//
// Unknown tags are skipped on load, so the format can grow. Included by flt_abi.cu after flt_trie / flt_lm.
#pragma once
#include <cstdio>
#include <cstring>

namespace {

constexpr char kTblMagic[8] = {'F', 'L', 'T', 'T', 'B', 'L', 0, 1};
enum : uint32_t {
  TBL_KIND_TRIE = 1, TBL_KIND_LM = 2,
  // Trie sections
  TT_META = 0x100, TT_CHILD_OFF, TT_CHILD_TOK, TT_CHILD_NODE, TT_LABEL_OFF, TT_LABELS, TT_SCORES, TT_MAX_SCORE,
  // LM sections (TL_KEYS + n, TL_CHK + n, TL_VALS + n for order n)
  TL_META = 0x200, TL_UNI, TL_WORD_OFF, TL_WORD_BYTES, TL_KEYS = 0x210, TL_CHK = 0x220, TL_VALS = 0x230
};

struct TblWriter {
  FILE* f;
  uint32_t n = 0;
  explicit TblWriter(const std::string& path, uint32_t kind) : f(fopen(path.c_str(), "wb")) {
    if (!f) throw FltError(FLT_ERR_RUNTIME, "cannot open " + path + " for writing");
    const uint32_t hdr[2] = {kind, 0};
    put(kTblMagic, 8), put(hdr, 8);
  }
  ~TblWriter() {
    if (f) fclose(f);
  }
  void put(const void* p, size_t bytes) {
    if (bytes && fwrite(p, 1, bytes, f) != bytes) throw FltError(FLT_ERR_RUNTIME, "table file: short write");
  }
  void section(uint32_t tag, const void* p, uint32_t elem, uint64_t count) {
    const uint32_t h[2] = {tag, elem};
    put(h, 8), put(&count, 8), put(p, (size_t)(elem * count));
    const char zero[8] = {0};
    put(zero, (size_t)((8 - (elem * count) % 8) % 8));
    ++n;
  }
  template <class T>
  void vec(uint32_t tag, const std::vector<T>& v) { section(tag, v.data(), (uint32_t)sizeof(T), v.size()); }
  void finish() {
    if (fseek(f, 12, SEEK_SET) != 0) throw FltError(FLT_ERR_RUNTIME, "table file: seek failed");
    put(&n, 4);
    if (fclose(f) != 0) {
      f = nullptr;
      throw FltError(FLT_ERR_RUNTIME, "table file: close failed");
    }
    f = nullptr;
  }
};

struct TblReader {
  FILE* f;
  uint32_t kind = 0, nSections = 0;
  explicit TblReader(const std::string& path) : f(fopen(path.c_str(), "rb")) {
    if (!f) throw FltError(FLT_ERR_RUNTIME, "cannot open " + path);
    char magic[8];
    uint32_t hdr[2];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kTblMagic, 8) != 0 || fread(hdr, 4, 2, f) != 2) {
      fclose(f);
      f = nullptr;
      throw FltError(FLT_ERR_RUNTIME, path + " is not a flt table file");
    }
    kind = hdr[0], nSections = hdr[1];
  }
  ~TblReader() {
    if (f) fclose(f);
  }
  // next section header; payload must then be consumed with read<T>() or skip()
  uint32_t tag = 0, elem = 0;
  uint64_t count = 0;
  bool next() {
    uint32_t h[2];
    if (fread(h, 4, 2, f) != 2 || fread(&count, 8, 1, f) != 1) return false;
    tag = h[0], elem = h[1];
    return true;
  }
  void pad() {
    const long p = (long)((8 - (elem * count) % 8) % 8);
    if (p) fseek(f, p, SEEK_CUR);
  }
  void skip() {
    fseek(f, (long)(elem * count), SEEK_CUR);
    pad();
  }
  template <class T>
  void read(std::vector<T>& v) {
    if (elem != sizeof(T) || count > (1ull << 33)) throw FltError(FLT_ERR_RUNTIME, "table file: malformed section");
    v.resize((size_t)count);
    if (count && fread(v.data(), sizeof(T), (size_t)count, f) != count)
      throw FltError(FLT_ERR_RUNTIME, "table file: truncated section");
    pad();
  }
};

bool isTableFile(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char magic[8];
  const bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kTblMagic, 8) == 0;
  fclose(f);
  return ok;
}

/* ------------------------------------------------------------------------------- Trie ---------- */
void saveTrie(const flt_trie& t, const std::string& path) {
  const size_t nn = t.nodes.size();
  std::vector<int> childOff(nn + 1, 0), childTok, childNode, labelOff(nn + 1, 0), labels;
  std::vector<float> scores, maxScore(nn);
  for (size_t i = 0; i < nn; ++i) {
    const HNode& nd = t.nodes[i];
    childOff[i] = (int)childTok.size();
    for (auto& kv : nd.kids) childTok.push_back(kv.first), childNode.push_back(kv.second);
    labelOff[i] = (int)labels.size();
    labels.insert(labels.end(), nd.labels.begin(), nd.labels.end());
    scores.insert(scores.end(), nd.scores.begin(), nd.scores.end());
    maxScore[i] = nd.maxScore;
  }
  childOff[nn] = (int)childTok.size();
  labelOff[nn] = (int)labels.size();
  TblWriter w(path, TBL_KIND_TRIE);
  const std::vector<int> meta = {t.maxChildren, t.rootIdx, (int)nn};
  w.vec(TT_META, meta), w.vec(TT_CHILD_OFF, childOff), w.vec(TT_CHILD_TOK, childTok), w.vec(TT_CHILD_NODE, childNode);
  w.vec(TT_LABEL_OFF, labelOff), w.vec(TT_LABELS, labels), w.vec(TT_SCORES, scores), w.vec(TT_MAX_SCORE, maxScore);
  w.finish();
}

void loadTrie(const std::string& path, flt_trie& t) {
  TblReader r(path);
  if (r.kind != TBL_KIND_TRIE) throw FltError(FLT_ERR_RUNTIME, path + " does not hold a Trie");
  std::vector<int> meta, childOff, childTok, childNode, labelOff, labels;
  std::vector<float> scores, maxScore;
  while (r.next()) {
    switch (r.tag) {
      case TT_META: r.read(meta); break;
      case TT_CHILD_OFF: r.read(childOff); break;
      case TT_CHILD_TOK: r.read(childTok); break;
      case TT_CHILD_NODE: r.read(childNode); break;
      case TT_LABEL_OFF: r.read(labelOff); break;
      case TT_LABELS: r.read(labels); break;
      case TT_SCORES: r.read(scores); break;
      case TT_MAX_SCORE: r.read(maxScore); break;
      default: r.skip();
    }
  }
  const size_t nn = meta.size() == 3 ? (size_t)meta[2] : 0;
  if (nn == 0 || childOff.size() != nn + 1 || labelOff.size() != nn + 1 || maxScore.size() != nn ||
      childTok.size() != childNode.size() || labels.size() != scores.size() ||
      childOff[0] != 0 || labelOff[0] != 0 || (size_t)childOff[nn] != childTok.size() ||
      (size_t)labelOff[nn] != labels.size())
    throw FltError(FLT_ERR_RUNTIME, path + ": inconsistent Trie sections");
  t.maxChildren = meta[0];
  t.rootIdx = meta[1];
  t.nodes.assign(nn, HNode{});
  for (size_t i = 0; i < nn; ++i) {
    HNode& nd = t.nodes[i];
    if (childOff[i] > childOff[i + 1] || labelOff[i] > labelOff[i + 1])
      throw FltError(FLT_ERR_RUNTIME, path + ": inconsistent Trie offsets");
    auto hint = nd.kids.end();
    for (int e = childOff[i]; e < childOff[i + 1]; ++e) {
      // (an edge's token indexes the emission row on the device: same range as Trie::insert enforces)
      if (childNode[e] <= 0 || (size_t)childNode[e] >= nn || childTok[e] < 0 || childTok[e] >= meta[0])
        throw FltError(FLT_ERR_RUNTIME, path + ": bad Trie edge");
      hint = nd.kids.emplace_hint(hint, childTok[e], childNode[e]);
    }
    nd.labels.assign(labels.begin() + labelOff[i], labels.begin() + labelOff[i + 1]);
    nd.scores.assign(scores.begin() + labelOff[i], scores.begin() + labelOff[i + 1]);
    nd.maxScore = maxScore[i];
  }
}

/* --------------------------------------------------------------------------------- LM ---------- */
void saveLm(const flt_lm& m, const std::string& path) {
  if (m.kind != 1) throw FltError(FLT_ERR_INVALID, "only n-gram LMs have tables to save");
  TblWriter w(path, TBL_KIND_LM);
  const std::vector<int> meta = {m.order, m.vocab, m.bos, m.eos};
  w.vec(TL_META, meta), w.vec(TL_UNI, m.uni);
  std::vector<uint64_t> off(m.words.size() + 1, 0);
  std::string blob;
  for (size_t i = 0; i < m.words.size(); ++i) {
    off[i] = blob.size();
    blob += m.words[i];
  }
  off[m.words.size()] = blob.size();
  w.vec(TL_WORD_OFF, off);
  w.section(TL_WORD_BYTES, blob.data(), 1, blob.size());
  for (int n = 2; n <= kMaxOrder; ++n) {
    if (m.keys[n].empty()) continue;
    w.vec(TL_KEYS + n, m.keys[n]), w.vec(TL_CHK + n, m.chk[n]), w.vec(TL_VALS + n, m.vals[n]);
  }
  w.finish();
}

// usr index -> LM vocabulary id (Vocabulary::Index: OOV -> <unk> = 0), lm/KenLM.cpp:38-46
void mapUsrWords(flt_lm& m, const char* const* usrWords, int nUsr) {
  std::unordered_map<std::string, int> vocab;
  vocab.reserve(m.words.size() * 2);
  for (size_t i = 0; i < m.words.size(); ++i) vocab.emplace(m.words[i], (int)i);
  m.usr2lm.resize(nUsr);
  for (int i = 0; i < nUsr; ++i) {
    auto it = vocab.find(usrWords[i]);
    m.usr2lm[i] = it == vocab.end() ? 0 : it->second;
  }
}

void loadLmTables(const std::string& path, const char* const* usrWords, int nUsr, flt_lm& m) {
  TblReader r(path);
  if (r.kind != TBL_KIND_LM) throw FltError(FLT_ERR_RUNTIME, "[KenLM] LM loading failed: " + path + " does not hold an LM");
  std::vector<int> meta;
  std::vector<uint64_t> off;
  std::vector<char> blob;
  while (r.next()) {
    const uint32_t n = r.tag & 0xF;
    if (r.tag == TL_META) r.read(meta);
    else if (r.tag == TL_UNI) r.read(m.uni);
    else if (r.tag == TL_WORD_OFF) r.read(off);
    else if (r.tag == TL_WORD_BYTES) r.read(blob);
    else if ((r.tag & ~0xFu) == TL_KEYS && n >= 2 && n <= (uint32_t)kMaxOrder) r.read(m.keys[n]);
    else if ((r.tag & ~0xFu) == TL_CHK && n >= 2 && n <= (uint32_t)kMaxOrder) r.read(m.chk[n]);
    else if ((r.tag & ~0xFu) == TL_VALS && n >= 2 && n <= (uint32_t)kMaxOrder) r.read(m.vals[n]);
    else r.skip();
  }
  if (meta.size() != 4 || m.uni.empty() || (int)m.uni.size() != meta[1] || off.size() != m.uni.size() + 1 ||
      off.back() != blob.size())
    throw FltError(FLT_ERR_RUNTIME, "[KenLM] LM loading failed: inconsistent sections in " + path);
  for (int n = 2; n <= kMaxOrder; ++n) {
    const size_t cap = m.keys[n].size();
    if (cap != m.chk[n].size() || cap != m.vals[n].size() || (cap & (cap - 1)) != 0)
      throw FltError(FLT_ERR_RUNTIME, "[KenLM] LM loading failed: inconsistent n-gram table in " + path);
  }
  m.kind = 1;
  m.order = meta[0], m.vocab = meta[1], m.bos = meta[2], m.eos = meta[3];
  if (m.order < 1 || m.order > kMaxOrder || m.bos < 0 || m.bos >= m.vocab || m.eos < 0 || m.eos >= m.vocab)
    throw FltError(FLT_ERR_RUNTIME, "[KenLM] LM loading failed: bad header in " + path);
  m.words.resize(m.uni.size());
  for (size_t i = 0; i < m.words.size(); ++i) {
    if (off[i] > off[i + 1]) throw FltError(FLT_ERR_RUNTIME, "[KenLM] LM loading failed: bad vocabulary in " + path);
    m.words[i].assign(blob.data() + off[i], blob.data() + off[i + 1]);
  }
  mapUsrWords(m, usrWords, nUsr);
  m.makeHostView();
}

} // namespace
