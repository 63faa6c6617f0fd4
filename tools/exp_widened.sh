#!/bin/bash
# One gpurun call: GPU parity tests + bench lines of the full-expansion modes (logAdd, token LM).
# usage: tools/exp_widened.sh <tag>
set -u
TAG=${1:-w1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest_gpu.txt
( timeout 300 python bench.py --steps 3 --warmup 3 --log-add --bst 50 --threshold 25 --no-e2e ) > $OUT/bench_lexfree_logadd_bst50.json 2> $OUT/bench_lexfree_logadd_bst50.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexfree_tokenlm --bst 50 --threshold 25 --no-e2e ) > $OUT/bench_lexfree_tokenlm_bst50.json 2> $OUT/bench_lexfree_tokenlm_bst50.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --log-add --bst 100 --threshold 25 --no-e2e --batch 128 ) > $OUT/bench_lexicon_logadd_bst100.json 2> $OUT/bench_lexicon_logadd_bst100.err
( timeout 300 python bench.py --steps 3 --warmup 3 --log-add --threshold 25 --no-e2e --frames 200 --sigma 4 ) > $OUT/bench_lexfree_logadd_bstN_sigma4.json 2> $OUT/bench_lexfree_logadd_bstN_sigma4.err
ls -la $OUT
