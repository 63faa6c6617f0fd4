#!/bin/bash
# compute-sanitizer over the full-expansion modes (logAdd chains, token LMs, list-walk rows) and the
# lexicon step's new select / list cache; then the whole GPU suite without the sanitizer.
OUT=gpurun_out/${1:-san2}; mkdir -p $OUT
SEL="lf_logadd_ctc or lf_logadd_bst_thr or lf_tokenlm_bst_thr or lf_tokenlm_logadd or lex_logadd_arpa or lex_logadd_bst_thr or lex_tokenlm_bst_unk or cfg3_scaled_bstN or arpa3_ctc_bst"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $OUT/memcheck.txt 2>&1; tail -3 $OUT/memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $OUT/racecheck.txt 2>&1; tail -3 $OUT/racecheck.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
