#!/bin/bash
# A/B on the lexicon workload (cfg 3): current build, FLT_PRUNE_WANT variants (percent of K kept by the
# two-pass pruning), then the GPU suite
set -u
TAG=${1:-ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in default 200 150 125; do
  if [ $v = default ]; then unset FLT_PRUNE_WANT; else export FLT_PRUNE_WANT=$v; fi
  ( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/lexicon_want_$v.json 2> $OUT/lexicon_want_$v.err
done
unset FLT_PRUNE_WANT
( FLT_PRUNE_WANT=150 timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon_lm --frames 300 --batch 256 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/lexlm_want_150.json 2> $OUT/lexlm_want_150.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon_lm --frames 300 --batch 256 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/lexlm_want_default.json 2> $OUT/lexlm_want_default.err
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/pytest_gpu.txt
ls -la $OUT
