/* flt_decoder.h — C-ABI of the B200-native beam-search decode hot path.
 *
 * Drop-in boundary for flashlight/text's LexiconDecoder / LexiconFreeDecoder decode path.
 * Plain pointers and sizes only (no torch / STL types). Every entry point names the reference
 * interface it replaces (paths relative to the reference root, flashlight/lib/text/...).
 * All functions return 0 on success, non-zero on error (message: flt_last_error(), thread-local);
 * nothing throws across this boundary. The C++ mirror classes (text_b200/csrc/host/) map codes
 * back to the exception types the reference throws.
 *
 * Threading: a decoder handle owns one CUDA stream and is single-owner, like the reference's
 * decoders (decoder/Utils.h:60-63). Tries and LMs are immutable once a decoder is created from
 * them and may be shared by several decoders, also across CUDA devices and threads: the flattened
 * tables are built once per device, under a lock. Every call makes the decoder's device current for
 * its own duration and restores the caller's current device before it returns.
 */
#ifndef FLT_DECODER_H
#define FLT_DECODER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLT_OK 0
#define FLT_ERR_INVALID 1      /* std::invalid_argument in the C++ mirror            */
#define FLT_ERR_OUT_OF_RANGE 2 /* std::out_of_range   (Trie.cpp:31-34)                */
#define FLT_ERR_RUNTIME 3      /* std::runtime_error  (lm/KenLM.cpp:36,40,67)         */
#define FLT_ERR_CUDA 4         /* CUDA runtime failure / no device                   */
#define FLT_ERR_UNSUPPORTED 5  /* configuration not implemented on the device path   */

typedef struct flt_trie flt_trie;
typedef struct flt_lm flt_lm;
typedef struct flt_decoder flt_decoder;

/* CriterionType, decoder/Decoder.h:16 */
#define FLT_CRITERION_ASG 0
#define FLT_CRITERION_CTC 1
/* SmearingMode, decoder/Trie.h:21-25 */
#define FLT_SMEAR_NONE 0
#define FLT_SMEAR_MAX 1
#define FLT_SMEAR_LOGADD 2

/* LexiconDecoderOptions (decoder/LexiconDecoder.h:21-31); LexiconFreeDecoderOptions
 * (decoder/LexiconFreeDecoder.h:20-28) is the same minus wordScore / unkScore, which the
 * lexicon-free decoder ignores. */
typedef struct {
  int32_t beamSize;
  int32_t beamSizeToken;
  double beamThreshold;
  double lmWeight;
  double wordScore;
  double unkScore;
  double silScore;
  int32_t logAdd;
  int32_t criterionType;
} flt_options;

const char* flt_last_error(void);

/* ---- Trie: decoder/Trie.h:66-86 (Trie::Trie / insert / search / smear). Built on the host with
 * the reference's semantics (<= 6 labels per node, Trie.cpp:40-46; out-of-range token ->
 * FLT_ERR_OUT_OF_RANGE, Trie.cpp:31-34), flattened to CSR tables in HBM when the first decoder
 * is created from it. */
int flt_trie_create(int32_t maxChildren, int32_t rootIdx, flt_trie** out);
int flt_trie_insert(flt_trie* trie, const int32_t* indices, int32_t n, int32_t label, float score);
int flt_trie_smear(flt_trie* trie, int32_t mode);
/* found = 0/1; maxScore / labels (<= 6) / scores of the node reached by `indices` */
int flt_trie_search(const flt_trie* trie, const int32_t* indices, int32_t n, int32_t* found,
                    float* maxScore, int32_t* nLabels, int32_t* labels6, float* scores6);
int flt_trie_num_nodes(const flt_trie* trie, int64_t* out);
/* TrieNode::maxScore (decoder/Trie.h:50) of every node, in the order the nodes were created by
 * flt_trie_insert (node 0 = the root): lets a host-side mirror of the node tree (TrieNode::children,
 * decoder/Trie.h:39-54) take the smeared scores over after flt_trie_smear. n = flt_trie_num_nodes. */
int flt_trie_max_scores(const flt_trie* trie, float* out, int64_t n);
void flt_trie_destroy(flt_trie* trie);
/* Table file of a built (inserted + smeared) Trie: what test/decoder/DecoderTest.cpp:126-146 rebuilds in every
 * process with one Trie::insert per word — saved once, loaded with a few reads (format: csrc/table_io.h). */
int flt_trie_save(const flt_trie* trie, const char* path);
int flt_trie_load(const char* path, flt_trie** out);
/* The node tree as CSR arrays, nodes in creation order (node 0 = the root), edges of a node ascending by
 * token: meta5 = {maxChildren, rootIdx, nNodes, nEdges, nLabels}; with childOff == NULL only meta5 is filled
 * (sizes for the second call). childOff / labelOff hold nNodes+1 entries. Lets a host mirror rebuild
 * TrieNode::children (decoder/Trie.h:39-54) for a Trie that came from flt_trie_load. */
int flt_trie_export(const flt_trie* trie, int32_t* meta5, int32_t* childOff, int32_t* childTok, int32_t* childNode,
                    int32_t* labelOff, int32_t* labels, float* scores, float* maxScore);

/* ---- LM: decoder/lm/LM.h:52-85. Two device-resident models:
 *   zero  = ZeroLM (lm/ZeroLM.cpp:14-26): score 0, a child state per (state, index)
 *   ngram = the KenLM adapter's role (lm/KenLM.cpp:32-83) for ARPA back-off models: log10 scores,
 *           usr index -> LM vocabulary map built from `usrWords` (OOV -> <unk>, id 0),
 *           start = <s> context, finish scores </s>. KenLM binary files are not supported. */
int flt_lm_zero_create(flt_lm** out);
int flt_lm_ngram_load_arpa(const char* path, const char* const* usrWords, int32_t nUsrWords,
                           flt_lm** out);
/* The hashed n-gram tables + vocabulary as a table file (the role of KenLM's binary format, which the
 * reference's KenLM constructor accepts in place of ARPA text, lm/KenLM.cpp:32-47): flt_lm_ngram_load_arpa
 * recognises such a file by its magic and loads it without parsing text. */
int flt_lm_save(const flt_lm* lm, const char* path);
/* LM::start(false) then LM::score per index (and LM::finish when withFinish): host-side query used
 * for Trie insertion scores (test/decoder/DecoderTest.cpp:126-141) and tests. out[n(+1)]. */
int flt_lm_score_seq(const flt_lm* lm, const int32_t* usrIdx, int32_t n, int32_t withFinish,
                     float* out);
void flt_lm_destroy(flt_lm* lm);

/* ---- Decoders: LexiconFreeDecoder(opt, lm, sil, blank, transitions)
 * (decoder/LexiconFreeDecoder.h:102-112) and LexiconDecoder(opt, trie, lm, sil, blank, unk,
 * transitions, isLmToken) (decoder/LexiconDecoder.h:117-133). `device` = CUDA ordinal. */
int flt_decoder_create_lexfree(const flt_options* opt, const flt_lm* lm, int32_t sil, int32_t blank,
                               const float* transitions, int64_t nTransitions, int32_t device,
                               flt_decoder** out);
int flt_decoder_create_lexicon(const flt_options* opt, const flt_trie* trie, const flt_lm* lm,
                               int32_t sil, int32_t blank, int32_t unk, const float* transitions,
                               int64_t nTransitions, int32_t isLmToken, int32_t device,
                               flt_decoder** out);
void flt_decoder_destroy(flt_decoder* dec);

/* How many of the final hypotheses are materialised (tokens / words backtrace) per utterance by
 * flt_decode_batch*. Default: beamSize, i.e. everything getAllFinalHypothesis returns. */
int flt_decoder_set_nbest(flt_decoder* dec, int32_t nbest);

/* Decoder::decode (decoder/Decoder.h:51-57: decodeBegin + decodeStep + decodeEnd +
 * getAllFinalHypothesis) for B utterances at once. `emissions` is row-major [B,T,N] fp32
 * (e[(b*T+t)*N+n], decoder/LexiconDecoder.cpp:50,69), HOST or DEVICE memory (detected with
 * cudaPointerGetAttributes), borrowed until the call returns. `lengths` (host, may be NULL =
 * all T) gives the number of valid frames per utterance. The n-best lists stay on the device
 * until flt_nbest_copy. B = 1 is the reference's single-utterance decode(). */
int flt_decode_batch(flt_decoder* dec, const float* emissions, int32_t B, int32_t T, int32_t N,
                     const int32_t* lengths);
/* Same, device emissions, enqueue only (no host synchronisation): for timing on a stream. A frame whose
 * (data-dependent) candidate count exceeds the planned capacity cannot be redone here: the error
 * ("candidate capacity exceeded") surfaces at flt_nbest_copy, which also grows the capacity, so calling
 * flt_decode_batch_async + flt_nbest_copy again succeeds. flt_decode_batch does that retry itself. */
int flt_decode_batch_async(flt_decoder* dec, const float* dEmissions, int32_t B, int32_t T,
                           int32_t N, const int32_t* dLengths);
int flt_decoder_synchronize(flt_decoder* dec);
/* cudaStream_t the decoder enqueues on (for CUDA-event timing by the caller). */
void* flt_decoder_stream(flt_decoder* dec);

/* getAllFinalHypothesis (decoder/Utils.h:229-266) for the last batch: for utterance b, hypothesis
 * r < counts[b] (sorted by score, best first), tokens/words hold T+2 entries
 * (seed, T frames, finish record; -1 padded past lengths[b]+2):
 *   tokens[(b*nbest + r)*(T+2) + i], words[...], scores[(b*nbest + r)*3 + {0: score,
 *   1: emittingModelScore, 2: lmScore}]. nbest <= the nbest setting the LAST BATCH was decoded with
 * (flt_decoder_set_nbest calls made after that decode do not apply to it); host buffers. counts[b] is
 * the total number of final hypotheses (may exceed nbest). */
int flt_nbest_copy(flt_decoder* dec, int32_t nbest, int32_t* tokens, int32_t* words,
                   double* scores, int32_t* counts);

/* ---- Online decoding of ONE utterance: decodeBegin / decodeStep(chunk) / decodeEnd / prune /
 * nDecodedFramesInBuffer / getBestHypothesis / getAllFinalHypothesis (decoder/Decoder.h:18-35,
 * decoder/LexiconDecoder.cpp:285-325, decoder/LexiconFreeDecoder.cpp:188-227, decoder/Utils.h:268-342).
 * The beam stays on the device between chunks; chunk emissions [T,N] may be host or device memory.
 * tokens / words rows hold maxLen entries; *len / lens[r] = frames in the buffer + 1 (0 = none). */
int flt_stream_begin(flt_decoder* dec, int32_t N);
int flt_stream_step(flt_decoder* dec, const float* emissions, int32_t T, int32_t N);
int flt_stream_end(flt_decoder* dec);
int flt_stream_prune(flt_decoder* dec, int32_t lookBack);
int flt_stream_frames_in_buffer(flt_decoder* dec, int32_t* out);
int flt_stream_n_hypothesis(flt_decoder* dec, int32_t* out);
int flt_stream_best(flt_decoder* dec, int32_t lookBack, int32_t maxLen, int32_t* tokens, int32_t* words,
                    double* scores3, int32_t* len);
int flt_stream_all_final(flt_decoder* dec, int32_t maxHyp, int32_t maxLen, int32_t* tokens,
                         int32_t* words, double* scores3, int32_t* lens, int32_t* count);

/* Device addresses of the last batch's n-best buffers (valid until the next flt_decode_batch* call
 * on this decoder; synchronise the decoder's stream before reading them from another stream):
 * tokens / words [B, nbestSetting, T+2] int32, scores [B, beamSize, 3] fp64, counts [B] int32.
 * For device-side consumers, e.g. the NCCL gather of n-best blocks across GPUs. */
int flt_nbest_device_ptrs(flt_decoder* dec, int32_t** tokens, int32_t** words, double** scores,
                          int32_t** counts);

/* Introspection for benchmarks: kernels launched by the last flt_decode_batch* call, and device
 * bytes currently held by the decoder workspace. */
int flt_decoder_last_launches(const flt_decoder* dec, int32_t* out);
/* Per-kernel device time of the last flt_decode_batch* call, from CUDA events recorded on the
 * decoder's stream around each launch (only while timing is on): ms4 / launches4 index
 * 0 = token-beam select, 1 = beam step, 2 = n-best backtrace, 3 = fused select + step.
 * Synchronises the stream. on = 1: events only; on = 2: also the in-kernel work / phase counters
 * read by flt_decoder_last_stats (they cost a few hundred cycles per frame). */
int flt_decoder_set_timing(flt_decoder* dec, int32_t on);
int flt_decoder_last_kernel_ms(flt_decoder* dec, float* ms4, int32_t* launches4);
/* Beam-step work counters of the last call (collected while timing is on), summed over frames:
 * out16[0..3] = {frames stepped, work items, live candidates (merge groups), survivors};
 * out16[4..10] = SM cycles thread 0 of the lexicon-free step spent in each phase (hash insert,
 * emit, scan, rank, new beam, wait for the producers, hand-over + emission gather); rest 0. */
int flt_decoder_last_stats(flt_decoder* dec, uint64_t* out16); /* 32 entries; [16..27] per-warp emit time */
int flt_decoder_workspace_bytes(const flt_decoder* dec, int64_t* out);

/* Stand-alone entry to the token-beam select kernel (decoder/LexiconFreeDecoder.cpp:39-51:
 * iota + partial_sort of one emission row), for kernel-level tests and roofline timing.
 * dEmissions [rows,N] device; outputs device: dTok/dVal [rows,M] sorted by value descending. */
int flt_topm_rows(const float* dEmissions, int64_t rows, int32_t N, int32_t M, int32_t* dTok,
                  float* dVal, void* stream);
/* The same with the lexicon decoder's ranking offset (decoder/LexiconDecoder.cpp:41-52 sorts raw emissions; the
 * device path ranks the root's children by e[n] + bias[n], bias[n] = lmWeight * smeared score of root child n,
 * -inf = token n does not start a word): tokens by e + bias descending (fp32 sum), dVal = the raw e[tok];
 * rows with fewer than M eligible tokens are padded with tok = -1, val = 0. dBias [N] device, may be NULL. */
int flt_topm_rows_bias(const float* dEmissions, int64_t rows, int32_t N, int32_t M, const float* dBias,
                       int32_t* dTok, float* dVal, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLT_DECODER_H */
