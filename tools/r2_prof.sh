#!/bin/bash
set -u
TAG=${1:-p1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_fused -s 1 -c 1 \
  -o $OUT/prof_lexfree python bench.py --steps 1 --warmup 1 --frames 100 --no-e2e --no-cpu-baseline ) > $OUT/prof_lexfree.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_gx -s 1 -c 1 \
  -o $OUT/prof_lexicon python bench.py --steps 1 --warmup 1 --frames 60 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/prof_lexicon.log 2>&1
ls -la $OUT
