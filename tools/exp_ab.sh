#!/bin/bash
# cfg 3 with the experiment switches (FLT_DBG: 1 = radix select only, 2 = plain histogram adds), other
# lexicon workloads on the default build, then the GPU suite
set -u
TAG=${1:-ab5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in 0 1 2; do
  ( FLT_DBG=$v timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/lexicon_dbg$v.json 2> $OUT/lexicon_dbg$v.err
done
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon_lm --frames 300 --batch 256 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/lexlm.json 2> $OUT/lexlm.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --sigma 4 --no-e2e --no-cpu-baseline ) > $OUT/lexicon_sigma4.json 2> $OUT/lexicon_sigma4.err
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/pytest_gpu.txt
ls -la $OUT
