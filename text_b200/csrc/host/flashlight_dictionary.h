// flashlight_dictionary.h — the host setup path that feeds the decode path:
// token / word dictionaries, lexicon files, replabels. C++17 mirror of
//   flashlight/lib/text/dictionary/Dictionary.h:23-66     Dictionary
//   flashlight/lib/text/dictionary/Utils.h:21-60          LexiconMap, createWordDict, loadWords,
//                                                          splitWrd, packReplabels, unpackReplabels, tkn2Idx
//   flashlight/lib/text/dictionary/Defines.h:14-17        kUnkToken, kEosToken
// with the reference's names, signatures, file formats and exception types. These are small canonical
// algorithms behind a name-compatible API: packReplabels / unpackReplabels / tkn2Idx / Dictionary::addEntry
// RESTATE dictionary/Utils.cpp:94-162 and Dictionary.cpp:60-84 (same control flow and messages, so that
// DictionaryTest's vectors and error strings carry over); they are not an independent design.
// Plain host code: none of this is on the timed path.
#pragma once
#include <fstream>
#include <istream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace fl {
namespace lib {
namespace text {

constexpr const char* kUnkToken = "<unk>";
constexpr const char* kEosToken = "</s>";

namespace detail {
inline std::vector<std::string> fields(const std::string& line) { // split on whitespace, drop empties
  std::vector<std::string> out;
  std::istringstream is(line);
  std::string f;
  while (is >> f) out.push_back(f);
  return out;
}
} // namespace detail

// bidirectional map entry <-> index; several entries may share an index (Dictionary.h:23-66)
class Dictionary {
 public:
  Dictionary() = default;
  explicit Dictionary(std::istream& stream) { createFromStream(stream); }
  explicit Dictionary(const std::string& filename) {
    std::ifstream stream(filename);
    if (!stream) throw std::runtime_error("Dictionary - cannot open file  " + filename);
    createFromStream(stream);
  }
  explicit Dictionary(const std::vector<std::string>& tkns) {
    for (const auto& t : tkns) addEntry(t);
    if (!isContiguous()) throw std::runtime_error("Invalid dictionary format - not contiguous");
  }

  size_t entrySize() const { return entry2idx_.size(); }
  size_t indexSize() const { return idx2entry_.size(); }

  void addEntry(const std::string& entry, int idx) {
    if (entry2idx_.count(entry)) throw std::invalid_argument("Duplicate entry name in dictionary '" + entry + "'");
    entry2idx_[entry] = idx;
    idx2entry_.emplace(idx, entry); // the first entry of an index names it
  }
  void addEntry(const std::string& entry) {
    if (entry2idx_.count(entry)) throw std::invalid_argument("Duplicate entry in dictionary '" + entry + "'");
    int idx = (int)idx2entry_.size();
    while (idx2entry_.count(idx)) ++idx; // first free index
    addEntry(entry, idx);
  }
  std::string getEntry(int idx) const {
    auto it = idx2entry_.find(idx);
    if (it == idx2entry_.end()) throw std::invalid_argument("Unknown index in dictionary '" + std::to_string(idx) + "'");
    return it->second;
  }
  void setDefaultIndex(int idx) { defaultIndex_ = idx; }
  int getIndex(const std::string& entry) const {
    auto it = entry2idx_.find(entry);
    if (it != entry2idx_.end()) return it->second;
    if (defaultIndex_ < 0) throw std::invalid_argument("Unknown entry in dictionary: '" + entry + "'");
    return defaultIndex_;
  }
  bool contains(const std::string& entry) const { return entry2idx_.count(entry) > 0; }
  bool isContiguous() const {
    for (size_t i = 0; i < indexSize(); ++i)
      if (!idx2entry_.count((int)i)) return false;
    for (const auto& kv : entry2idx_)
      if (!idx2entry_.count(kv.second)) return false;
    return true;
  }
  std::vector<int> mapEntriesToIndices(const std::vector<std::string>& entries) const {
    std::vector<int> out;
    out.reserve(entries.size());
    for (const auto& e : entries) out.push_back(getIndex(e));
    return out;
  }
  std::vector<std::string> mapIndicesToEntries(const std::vector<int>& indices) const {
    std::vector<std::string> out;
    out.reserve(indices.size());
    for (int i : indices) out.push_back(getEntry(i));
    return out;
  }

 private:
  // one line per index; all entries of a line share it
  void createFromStream(std::istream& stream) {
    if (!stream) throw std::runtime_error("Unable to open dictionary input stream.");
    std::string line;
    while (std::getline(stream, line)) {
      if (line.empty()) continue;
      const int idx = (int)idx2entry_.size();
      for (const auto& t : detail::fields(line)) addEntry(t, idx);
    }
    if (!isContiguous()) throw std::runtime_error("Invalid dictionary format - not contiguous");
  }
  std::unordered_map<std::string, int> entry2idx_;
  std::unordered_map<int, std::string> idx2entry_;
  int defaultIndex_ = -1;
};
using DictionaryMap = std::unordered_map<int, Dictionary>;

// word -> its spellings (token strings), dictionary/Utils.h:21-22
using LexiconMap = std::unordered_map<std::string, std::vector<std::vector<std::string>>>;

// lexicon file: one spelling per line, "word tok tok ..."; <unk> is always added (Utils.cpp:28-62)
inline LexiconMap loadWords(const std::string& filename, int maxWords = -1) {
  std::ifstream in(filename);
  if (!in) throw std::invalid_argument("text::loadWords - can't open file " + filename);
  LexiconMap lexicon;
  std::string line;
  while ((maxWords < 0 || (size_t)maxWords != lexicon.size()) && std::getline(in, line)) {
    auto f = detail::fields(line);
    if (f.size() < 2) throw std::runtime_error("[loadWords] Invalid line: " + line);
    lexicon[f[0]].emplace_back(f.begin() + 1, f.end());
  }
  lexicon[kUnkToken] = {}; // present, never with spellings (dictionary/Utils.cpp:61)
  return lexicon;
}

// one entry per lexicon word (map iteration order, like the reference); unknown words -> <unk>
inline Dictionary createWordDict(const LexiconMap& lexicon) {
  Dictionary dict;
  for (const auto& kv : lexicon) dict.addEntry(kv.first);
  dict.setDefaultIndex(dict.getIndex(kUnkToken));
  return dict;
}

// UTF-8 aware split into single-character tokens
inline std::vector<std::string> splitWrd(const std::string& word) {
  std::vector<std::string> tokens;
  const size_t len = word.size();
  for (size_t i = 0; i < len;) {
    const unsigned char c = (unsigned char)word[i];
    const int n = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : -1;
    if (n < 0 || i + n > len) throw std::runtime_error("splitWrd: invalid UTF-8 : " + word);
    tokens.emplace_back(word, i, n);
    i += n;
  }
  return tokens;
}

// runs of a repeated token become token + "<k>" replabel (k extra copies, k <= maxReps)
inline std::vector<int> packReplabels(const std::vector<int>& tokens, const Dictionary& dict, int maxReps) {
  if (tokens.empty() || maxReps <= 0) return tokens;
  std::vector<int> rep(maxReps + 1);
  for (int i = 1; i <= maxReps; ++i) rep[i] = dict.getIndex("<" + std::to_string(i) + ">");
  std::vector<int> out;
  int prev = -1, reps = 0;
  for (int t : tokens) {
    if (t == prev && reps < maxReps) {
      ++reps;
      continue;
    }
    if (reps > 0) out.push_back(rep[reps]);
    reps = 0;
    out.push_back(t);
    prev = t;
  }
  if (reps > 0) out.push_back(rep[reps]);
  return out;
}

// inverse; a replabel with no token to repeat (at the start, or right after another replabel) is dropped
inline std::vector<int> unpackReplabels(const std::vector<int>& tokens, const Dictionary& dict, int maxReps) {
  if (tokens.empty() || maxReps <= 0) return tokens;
  std::unordered_map<int, int> value;
  for (int i = 1; i <= maxReps; ++i) value.emplace(dict.getIndex("<" + std::to_string(i) + ">"), i);
  std::vector<int> out;
  int prev = -1;
  for (int t : tokens) {
    auto it = value.find(t);
    if (it == value.end()) {
      out.push_back(t);
      prev = t;
    } else if (prev != -1) {
      out.insert(out.end(), it->second, prev);
      prev = -1;
    }
  }
  return out;
}

inline std::vector<int> tkn2Idx(const std::vector<std::string>& spelling, const Dictionary& tokenDict, int maxReps) {
  std::vector<int> idx;
  idx.reserve(spelling.size());
  for (const auto& t : spelling) idx.push_back(tokenDict.getIndex(t));
  return packReplabels(idx, tokenDict, maxReps);
}

} // namespace text
} // namespace lib
} // namespace fl
