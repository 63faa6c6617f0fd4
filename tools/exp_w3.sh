#!/bin/bash
# GPU parity suite + lexicon (cfg 3) bench after the packed edge records + full-expansion benches
set -u
TAG=${1:-w3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest_gpu.txt
( timeout 600 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e ) > $OUT/bench_lexicon.json 2> $OUT/bench_lexicon.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexfree_tokenlm --bst 50 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/bench_lexfree_tokenlm_bst50.json 2> $OUT/bench_lexfree_tokenlm_bst50.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --log-add --bst 100 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/bench_lexicon_logadd_bst100.json 2> $OUT/bench_lexicon_logadd_bst100.err
( timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline ) > $OUT/bench_lexfree.json 2> $OUT/bench_lexfree.err
ls -la $OUT
