/* ORACLE (test infrastructure, never shipped).
 *
 * One C API, exported twice with different prefixes:
 *   ref_*  by oracle/_ref/libflref.so  = the UNMODIFIED reference sources compiled in place
 *          from /root/reference (oracle/Makefile), wrapped by oracle/ref_driver.cpp
 *   ora_*  by oracle/liboracle.so      = our CPU restatement of the same path
 *          (oracle/oracle_decoder.cpp), validated against ref_* and the golden vectors.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load either library.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t beamSize;
  int32_t beamSizeToken;
  double beamThreshold;
  double lmWeight;
  double wordScore; /* lexicon decoder only */
  double unkScore;  /* lexicon decoder only */
  double silScore;
  int32_t logAdd;
  int32_t criterion; /* 0 = ASG, 1 = CTC (Decoder.h:16) */
} ora_options;

#define ORA_DECLARE(PFX)                                                                          \
  void* PFX##_trie_create(int maxChildren, int rootIdx);                                         \
  int PFX##_trie_insert(void* trie, const int* idx, int n, int label, float score);              \
  void PFX##_trie_smear(void* trie, int mode);                                                   \
  int PFX##_trie_search(void* trie, const int* idx, int n, float* maxScore, int* nLabels,        \
                        int* labels6, float* scores6);                                           \
  void PFX##_trie_destroy(void* trie);                                                           \
  void* PFX##_lm_zero(void);                                                                     \
  void* PFX##_lm_arpa(const char* path, const char* const* words, int nWords);                   \
  int PFX##_lm_score_seq(void* lm, const int* usrIdx, int n, int withFinish, float* out);        \
  void PFX##_lm_destroy(void* lm);                                                               \
  void* PFX##_decoder_lexfree(const ora_options* opt, void* lm, int sil, int blank,              \
                              const float* trans, int nTrans);                                   \
  void* PFX##_decoder_lexicon(const ora_options* opt, void* trie, void* lm, int sil, int blank,  \
                              int unk, const float* trans, int nTrans, int isLmToken);           \
  void PFX##_decoder_destroy(void* dec);                                                         \
  int PFX##_decode(void* dec, const float* emis, int T, int N, int maxHyp, double* scores3,      \
                   int* tokens, int* words);                                                     \
  void PFX##_decode_begin(void* dec);                                                            \
  void PFX##_decode_step(void* dec, const float* emis, int T, int N);                            \
  void PFX##_decode_end(void* dec);                                                              \
  void PFX##_prune(void* dec, int lookBack);                                                     \
  int PFX##_n_hypothesis(void* dec);                                                             \
  int PFX##_n_frames_in_buffer(void* dec);                                                       \
  int PFX##_best(void* dec, int lookBack, int maxLen, double* scores3, int* tokens, int* words); \
  int PFX##_all_final(void* dec, int maxHyp, int maxLen, double* scores3, int* tokens,           \
                      int* words, int* len);                                                     \
  double PFX##_bench_mt(int lexicon, const ora_options* opt, void* trie, void* lm, int sil,      \
                        int blank, int unk, int isLmToken, const float* emis, int B, int T,      \
                        int N, int nThreads, int warmup);                                        \
  const char* PFX##_last_error(void);

ORA_DECLARE(ref)
ORA_DECLARE(ora)
/* restatement only: events during the current utterance whose outcome the reference leaves to
 * libstdc++ internals (equal scores inside a merge group, at the beam cut, at the token-beam cut).
 * Tests exclude such utterances from bit-exact comparison. */
long ora_tie_events(void* dec);

#ifdef __cplusplus
}
#endif
