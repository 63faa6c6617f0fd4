// Test program written against the C++ mirror exactly as one would write it against the reference
// (cf. flashlight/lib/text/test/decoder/DecoderTest.cpp:126-183: build the Trie from word
// spellings scored by the LM, smear, construct LexiconDecoder with positional options, decode,
// read DecodeResult). Input: N T W, W spellings (len, tokens...), T*N emissions. Output: JSON.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <limits>
#include <vector>

#include "../../text_b200/csrc/host/flashlight_text.h"

using namespace fl::lib::text;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  int hdr[3];
  in.read((char*)hdr, sizeof(hdr));
  const int N = hdr[0], T = hdr[1], W = hdr[2];
  auto lm = std::make_shared<ZeroLM>();
  auto trie = std::make_shared<Trie>(N, 0);
  auto startState = lm->start(false);
  for (int w = 0; w < W; ++w) {
    int len;
    in.read((char*)&len, 4);
    std::vector<int> sp(len);
    in.read((char*)sp.data(), 4 * len);
    float score;
    std::tie(std::ignore, score) = lm->score(startState, w);
    trie->insert(sp, w, score);
  }
  trie->smear(SmearingMode::MAX);
  std::vector<float> emissions((size_t)T * N);
  in.read((char*)emissions.data(), sizeof(float) * emissions.size());

  LexiconDecoderOptions opt{20, N, 1e9, 0.0, 0.3, -std::numeric_limits<double>::infinity(), 0.0, false, CriterionType::CTC};
  std::vector<float> transitions;
  LexiconDecoder decoder(opt, trie, lm, 0, N - 1, W, transitions, false);
  auto results = decoder.decode(emissions.data(), T, N);

  // error behaviour of the reference: out-of-range token -> std::out_of_range (Trie.cpp:31-34)
  bool threw = false;
  try {
    trie->search({N + 1});
  } catch (const std::out_of_range&) {
    threw = true;
  }
  if (!threw) return 3;

  std::cout.precision(17);
  std::cout << "[";
  for (size_t r = 0; r < results.size(); ++r) {
    const auto& d = results[r];
    std::cout << (r ? "," : "") << "{\"score\":" << d.score << ",\"am\":" << d.emittingModelScore << ",\"lm\":" << d.lmScore
              << ",\"tokens\":[";
    for (size_t i = 0; i < d.tokens.size(); ++i) std::cout << (i ? "," : "") << d.tokens[i];
    std::cout << "],\"words\":[";
    for (size_t i = 0; i < d.words.size(); ++i) std::cout << (i ? "," : "") << d.words[i];
    std::cout << "]}";
  }
  std::cout << "]\n";
  return 0;
}
