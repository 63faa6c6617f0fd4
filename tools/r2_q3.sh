#!/bin/bash
set -u
TAG=${1:-q3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_random.py -q -m gpu --tb=short -k "lexfree_random or lexicon_random" 2>&1 | tail -120 ) > $OUT/pytest_random.txt
( timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg2_scaled_bstN or bst_small or cfg3_scaled_bstN or zero_ctc_silpos or arpa3_ctc_bst" 2>&1 | tail -40 ) > $OUT/racecheck.txt
# source-level profile of the lexicon gx kernel (short T) and the fused lexicon-free kernel
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_gx -s 1 -c 1 \
  -o $OUT/prof_lexicon python bench.py --steps 1 --warmup 1 --frames 60 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/prof_lexicon.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_fused -s 1 -c 1 \
  -o $OUT/prof_lexfree python bench.py --steps 1 --warmup 1 --frames 100 --no-e2e --no-cpu-baseline ) > $OUT/prof_lexfree.log 2>&1
tail -60 $OUT/pytest_random.txt; tail -12 $OUT/racecheck.txt; ls -la $OUT
