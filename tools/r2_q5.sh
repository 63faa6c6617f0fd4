#!/bin/bash
set -u
TAG=${1:-q5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_topm.py -x -q -m gpu 2>&1 | tail -5 ) > $OUT/pytest_topm.txt; cat $OUT/pytest_topm.txt
( timeout 300 python tools/bench_topm.py 53 105 205 ) > $OUT/topm.jsonl 2> $OUT/topm.err; cat $OUT/topm.jsonl; tail -2 $OUT/topm.err
( FLT_NO_STREAM=1 timeout 300 python tools/bench_topm.py 53 205 ) > $OUT/topm_old.jsonl 2>/dev/null; cat $OUT/topm_old.jsonl
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
( FLT_DBG_PLAN=1 timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
