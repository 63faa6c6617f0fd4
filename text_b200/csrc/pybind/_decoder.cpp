// pybind11 module `flashlight_lib_text_decoder` over the C++ mirror (../host/flashlight_text.h):
// the Python names, constructor kwargs and method names of the reference's decoder bindings
// (bindings/python/flashlight/lib/text/_decoder.cpp:167-441, decoder/_kenlm.cpp:19-25,
// _dictionary.cpp:33-60) for the CTC/ASG decode path, plus `decode_batch`.
// `decode` / `decode_step` take a raw integer address of fp32 emissions like the reference
// (_decoder.cpp:96-126); the address may point to host or device memory.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>

#include "../host/flashlight_text.h"

namespace py = pybind11;
using namespace fl::lib::text;
using namespace py::literals;

namespace {

// user-defined LMs (the reference's PyLM trampoline, _decoder.cpp:39-56): constructible and
// callable from Python, but not usable by the device decoders (no handle)
class PyLM : public LM {
  using LM::LM;
  LMStatePtr start(bool startWithNothing) override {
    PYBIND11_OVERRIDE_PURE(LMStatePtr, LM, start, startWithNothing);
  }
  std::pair<LMStatePtr, float> score(const LMStatePtr& state, const int usrTokenIdx) override {
    PYBIND11_OVERRIDE_PURE(PYBIND11_TYPE(std::pair<LMStatePtr, float>), LM, score, state, usrTokenIdx);
  }
  std::pair<LMStatePtr, float> finish(const LMStatePtr& state) override {
    PYBIND11_OVERRIDE_PURE(PYBIND11_TYPE(std::pair<LMStatePtr, float>), LM, finish, state);
  }
};

template <class D>
std::vector<DecodeResult> decodePtr(D& d, uintptr_t emissions, int T, int N) {
  py::gil_scoped_release nogil;
  return d.decode(reinterpret_cast<const float*>(emissions), T, N);
}
template <class D>
void decodeStepPtr(D& d, uintptr_t emissions, int T, int N) {
  d.decodeStep(reinterpret_cast<const float*>(emissions), T, N);
}
template <class D>
std::vector<std::vector<DecodeResult>> decodeBatchPtr(D& d, uintptr_t emissions, int B, int T, int N,
                                                       py::object lengths, int nbest) {
  std::vector<int> len;
  if (!lengths.is_none()) len = lengths.cast<std::vector<int>>();
  if (!len.empty() && (int)len.size() != B) throw std::invalid_argument("lengths must hold B entries");
  py::gil_scoped_release nogil;
  return d.decodeBatch(reinterpret_cast<const float*>(emissions), B, T, N, len.empty() ? nullptr : len.data(), nbest);
}

template <class D, class C>
void defDecoderMethods(C& cls) {
  cls.def("decode_begin", &D::decodeBegin)
      .def("decode_step", &decodeStepPtr<D>, "emissions"_a, "T"_a, "N"_a)
      .def("decode_end", &D::decodeEnd)
      .def("decode", &decodePtr<D>, "emissions"_a, "T"_a, "N"_a)
      .def("decode_batch", &decodeBatchPtr<D>, "emissions"_a, "B"_a, "T"_a, "N"_a, "lengths"_a = py::none(),
           "nbest"_a = -1)
      .def("prune", &D::prune, "look_back"_a = 0)
      .def("get_best_hypothesis", &D::getBestHypothesis, "look_back"_a = 0)
      .def("get_all_final_hypothesis", &D::getAllFinalHypothesis)
      .def("n_hypothesis", &D::nHypothesis)
      .def("n_decoded_frames_in_buffer", &D::nDecodedFramesInBuffer);
}

} // namespace

PYBIND11_MODULE(flashlight_lib_text_decoder, m) {
  py::enum_<SmearingMode>(m, "SmearingMode")
      .value("NONE", SmearingMode::NONE)
      .value("MAX", SmearingMode::MAX)
      .value("LOGADD", SmearingMode::LOGADD);

  py::class_<TrieNode, TrieNodePtr>(m, "TrieNode")
      .def(py::init<int>(), "idx"_a)
      .def_readwrite("children", &TrieNode::children)
      .def_readwrite("idx", &TrieNode::idx)
      .def_readwrite("labels", &TrieNode::labels)
      .def_readwrite("scores", &TrieNode::scores)
      .def_readwrite("max_score", &TrieNode::maxScore);

  py::class_<Trie, TriePtr>(m, "Trie")
      .def(py::init<int, int>(), "max_children"_a, "root_idx"_a)
      // the reference returns the raw const TrieNode* of a shared_ptr-held type here, which makes
      // Python free it a second time; reference semantics are kept by policy
      .def("get_root", &Trie::getRoot, py::return_value_policy::reference_internal)
      .def("insert", &Trie::insert, "indices"_a, "label"_a, "score"_a)
      .def("search", &Trie::search, "indices"_a)
      .def("smear", &Trie::smear, "smear_mode"_a)
      // additive: table file of the built Trie (csrc/table_io.h)
      .def("save", &Trie::save, "path"_a)
      .def_static("load", &Trie::load, "path"_a);

  py::class_<LM, LMPtr, PyLM>(m, "LM")
      .def(py::init<>())
      .def("start", &LM::start, "start_with_nothing"_a)
      .def("score", &LM::score, "state"_a, "usr_token_idx"_a)
      .def("finish", &LM::finish, "state"_a);

  py::class_<LMState, LMStatePtr>(m, "LMState")
      .def(py::init<>())
      .def_readwrite("children", &LMState::children)
      .def("compare", &LMState::compare, "state"_a)
      .def("child", &LMState::child<LMState>, "usr_index"_a);

  py::class_<ZeroLM, ZeroLMPtr, LM>(m, "ZeroLM").def(py::init<>());

  // dictionary / lexicon setup path (bindings/python/flashlight/lib/text/_dictionary.cpp:33-60)
  py::class_<Dictionary>(m, "Dictionary")
      .def(py::init<>())
      .def(py::init<const std::string&>(), "filename"_a)
      .def(py::init<const std::vector<std::string>&>(), "tkns"_a)
      .def("entry_size", &Dictionary::entrySize)
      .def("index_size", &Dictionary::indexSize)
      .def("add_entry", py::overload_cast<const std::string&, int>(&Dictionary::addEntry), "entry"_a, "idx"_a)
      .def("add_entry", py::overload_cast<const std::string&>(&Dictionary::addEntry), "entry"_a)
      .def("get_entry", &Dictionary::getEntry, "idx"_a)
      .def("set_default_index", &Dictionary::setDefaultIndex, "idx"_a)
      .def("get_index", &Dictionary::getIndex, "entry"_a)
      .def("contains", &Dictionary::contains, "entry"_a)
      .def("is_contiguous", &Dictionary::isContiguous)
      .def("map_entries_to_indices", &Dictionary::mapEntriesToIndices, "entries"_a)
      .def("map_indices_to_entries", &Dictionary::mapIndicesToEntries, "indices"_a);
  m.def("create_word_dict", &createWordDict, "lexicon"_a);
  m.def("load_words", &loadWords, "filename"_a, "max_words"_a = -1);
  m.def("pack_replabels", &packReplabels, "tokens"_a, "dict"_a, "max_reps"_a);
  m.def("unpack_replabels", &unpackReplabels, "tokens"_a, "dict"_a, "max_reps"_a);
  m.def("tkn_to_idx", &tkn2Idx, "spelling"_a, "token_dict"_a, "max_reps"_a);
  m.def("split_wrd", &splitWrd, "word"_a);
  m.def("build_trie", &buildTrie, "lexicon"_a, "token_dict"_a, "word_dict"_a, "lm"_a, "sil_idx"_a,
        "max_reps"_a = 0, "smear_mode"_a = SmearingMode::MAX);

  py::class_<KenLM, KenLMPtr, LM>(m, "KenLM")
      .def(py::init<const std::string&, const Dictionary&>(), "path"_a, "usr_token_dict"_a)
      .def("save", &KenLM::save, "path"_a); // additive: table file; the constructor loads it in place of ARPA

  py::enum_<CriterionType>(m, "CriterionType")
      .value("ASG", CriterionType::ASG)
      .value("CTC", CriterionType::CTC)
      .value("S2S", CriterionType::S2S);

  py::class_<LexiconDecoderOptions>(m, "LexiconDecoderOptions")
      .def(py::init<const int, const int, const double, const double, const double, const double,
                    const double, const bool, const CriterionType>(),
           "beam_size"_a, "beam_size_token"_a, "beam_threshold"_a, "lm_weight"_a, "word_score"_a,
           "unk_score"_a, "sil_score"_a, "log_add"_a, "criterion_type"_a)
      .def_readwrite("beam_size", &LexiconDecoderOptions::beamSize)
      .def_readwrite("beam_size_token", &LexiconDecoderOptions::beamSizeToken)
      .def_readwrite("beam_threshold", &LexiconDecoderOptions::beamThreshold)
      .def_readwrite("lm_weight", &LexiconDecoderOptions::lmWeight)
      .def_readwrite("word_score", &LexiconDecoderOptions::wordScore)
      .def_readwrite("unk_score", &LexiconDecoderOptions::unkScore)
      .def_readwrite("sil_score", &LexiconDecoderOptions::silScore)
      .def_readwrite("log_add", &LexiconDecoderOptions::logAdd)
      .def_readwrite("criterion_type", &LexiconDecoderOptions::criterionType)
      .def(py::pickle(
          [](const LexiconDecoderOptions& p) {
            return py::make_tuple(p.beamSize, p.beamSizeToken, p.beamThreshold, p.lmWeight, p.wordScore,
                                  p.unkScore, p.silScore, p.logAdd, p.criterionType);
          },
          [](py::tuple t) {
            if (t.size() != 9)
              throw std::runtime_error("Cannot run __setstate__ on LexiconDecoderOptions - insufficient arguments provided.");
            return LexiconDecoderOptions{t[0].cast<int>(),    t[1].cast<int>(),    t[2].cast<double>(),
                                         t[3].cast<double>(), t[4].cast<double>(), t[5].cast<double>(),
                                         t[6].cast<double>(), t[7].cast<bool>(),   t[8].cast<CriterionType>()};
          }));

  py::class_<LexiconFreeDecoderOptions>(m, "LexiconFreeDecoderOptions")
      .def(py::init<const int, const int, const double, const double, const double, const bool,
                    const CriterionType>(),
           "beam_size"_a, "beam_size_token"_a, "beam_threshold"_a, "lm_weight"_a, "sil_score"_a,
           "log_add"_a, "criterion_type"_a)
      .def_readwrite("beam_size", &LexiconFreeDecoderOptions::beamSize)
      .def_readwrite("beam_size_token", &LexiconFreeDecoderOptions::beamSizeToken)
      .def_readwrite("beam_threshold", &LexiconFreeDecoderOptions::beamThreshold)
      .def_readwrite("lm_weight", &LexiconFreeDecoderOptions::lmWeight)
      .def_readwrite("sil_score", &LexiconFreeDecoderOptions::silScore)
      .def_readwrite("log_add", &LexiconFreeDecoderOptions::logAdd)
      .def_readwrite("criterion_type", &LexiconFreeDecoderOptions::criterionType)
      .def(py::pickle(
          [](const LexiconFreeDecoderOptions& p) {
            return py::make_tuple(p.beamSize, p.beamSizeToken, p.beamThreshold, p.lmWeight, p.silScore,
                                  p.logAdd, p.criterionType);
          },
          [](py::tuple t) {
            if (t.size() != 7)
              throw std::runtime_error("Cannot run __setstate__ on LexiconFreeDecoderOptions - insufficient arguments provided.");
            return LexiconFreeDecoderOptions{t[0].cast<int>(),    t[1].cast<int>(),  t[2].cast<double>(),
                                             t[3].cast<double>(), t[4].cast<double>(), t[5].cast<bool>(),
                                             t[6].cast<CriterionType>()};
          }));

  py::class_<DecodeResult>(m, "DecodeResult")
      .def(py::init<int>(), "length"_a)
      .def_readwrite("score", &DecodeResult::score)
      .def_readwrite("emittingModelScore", &DecodeResult::emittingModelScore)
      .def_readwrite("lmScore", &DecodeResult::lmScore)
      .def_readwrite("words", &DecodeResult::words)
      .def_readwrite("tokens", &DecodeResult::tokens);

  // NB: `decode`, `decode_step` and `decode_batch` expect raw emissions pointers (integers).
  py::class_<LexiconDecoder> lex(m, "LexiconDecoder");
  lex.def(py::init<LexiconDecoderOptions, const TriePtr, const LMPtr, const int, const int, const int,
                   const std::vector<float>&, const bool>(),
          "options"_a, "trie"_a, "lm"_a, "sil_token_idx"_a, "blank_token_idx"_a, "unk_token_idx"_a,
          "transitions"_a, "is_token_lm"_a);
  defDecoderMethods<LexiconDecoder>(lex);

  py::class_<LexiconFreeDecoder> lexfree(m, "LexiconFreeDecoder");
  lexfree.def(py::init<LexiconFreeDecoderOptions, const LMPtr, const int, const int, const std::vector<float>&>(),
              "options"_a, "lm"_a, "sil_token_idx"_a, "blank_token_idx"_a, "transitions"_a);
  defDecoderMethods<LexiconFreeDecoder>(lexfree);
  lexfree.def("get_options", &LexiconFreeDecoder::getOptions)
      .def("get_sil_idx", &LexiconFreeDecoder::getSilIdx)
      .def("get_blank_idx", &LexiconFreeDecoder::getBlankIdx)
      .def("get_transitions", &LexiconFreeDecoder::getTransitions)
      // same rules as the reference's pickling (bindings/python/flashlight/lib/text/_decoder.cpp:409-441): only a
      // decoder without state and with a ZeroLM can be pickled; the copy gets a ZeroLM of its own
      .def(py::pickle(
          [](const LexiconFreeDecoder& p) {
            if (p.getAllFinalHypothesis().size() != 0)
              throw std::runtime_error("LexiconFreeDecoder: cannot pickle decoder that has state");
            if (!std::dynamic_pointer_cast<ZeroLM>(p.getLMPtr()))
              throw std::runtime_error("LexiconFreeDecoder: cannot pickle a decoder with an "
                                       "integrated language model that is not ZeroLM");
            return py::make_tuple(p.getOptions(), p.getSilIdx(), p.getBlankIdx(), p.getTransitions());
          },
          [](py::tuple t) {
            if (t.size() != 4)
              throw std::runtime_error("Cannot run __setstate__ on LexiconFreeDecoder - insufficient arguments provided.");
            return std::make_unique<LexiconFreeDecoder>(t[0].cast<LexiconFreeDecoderOptions>(), std::make_shared<ZeroLM>(),
                                                        t[1].cast<int>(), t[2].cast<int>(),
                                                        t[3].cast<std::vector<float>>());
          }));
}
