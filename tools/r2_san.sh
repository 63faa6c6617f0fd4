#!/bin/bash
# compute-sanitizer over the round-2 device code: the lexicon-free step with the table fold and the guessed
# bound (fused kernel), the streaming select (keyed filter, long lists), the lexicon step with the histogram
# select / LM bound / back-off cache, and the split workspace forced on small cases (FLT_SMEM_KB=16).
OUT=gpurun_out/${1:-san3}; mkdir -p $OUT
SEL="cfg2_scaled_bstN or cfg2_scaled_bstK or sil_positive_bst or asg_bst_thr or zero_ctc_bst_thr or arpa3_ctc or zero_unk or long_ragged or cfg3_scaled_bstN or arpa3_ctc_bst"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_streaming.py -x -q -m gpu -k "$SEL or streaming_cuda" > $OUT/memcheck.txt 2>&1; tail -3 $OUT/memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $OUT/racecheck.txt 2>&1; tail -3 $OUT/racecheck.txt
FLT_SMEM_KB=16 FLT_DBG_PLAN=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg3_scaled_bstN or arpa3_ctc or zero_unk or cfg4_scaled" > $OUT/memcheck_split.txt 2>&1; tail -3 $OUT/memcheck_split.txt; grep -a "flt plan" $OUT/memcheck_split.txt | head -2 | cut -c1-250
FLT_SMEM_KB=16 timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg3_scaled_bstN or arpa3_ctc" > $OUT/racecheck_split.txt 2>&1; tail -3 $OUT/racecheck_split.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_topm.py -x -q -m gpu -k "biased or many_rows" > $OUT/memcheck_topm.txt 2>&1; tail -3 $OUT/memcheck_topm.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_topm.py -x -q -m gpu -k "biased_and_long and (205 or 505)" > $OUT/racecheck_topm.txt 2>&1; tail -3 $OUT/racecheck_topm.txt
