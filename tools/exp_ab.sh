#!/bin/bash
# lexicon workloads after the adaptive keep-count of the two-pass pruning + phase counters + GPU suite
set -u
TAG=${1:-ab3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/lexicon.json 2> $OUT/lexicon.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --bst 100 --no-e2e ) > $OUT/lexicon_bst100.json 2> $OUT/lexicon_bst100.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon_lm --frames 300 --batch 256 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/lexlm.json 2> $OUT/lexlm.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --sigma 4 --no-e2e --no-cpu-baseline ) > $OUT/lexicon_sigma4.json 2> $OUT/lexicon_sigma4.err
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/pytest_gpu.txt
ls -la $OUT
