#!/bin/bash
# A/B of library variants (tools/variants/*.so, FLT_LIB) on the lexicon workload
set -u
TAG=${1:-ab1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in s4 A B C; do
  ( FLT_LIB=$PWD/tools/variants/libflt_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/lexicon_$v.json 2> $OUT/lexicon_$v.err
done
ls -la $OUT
