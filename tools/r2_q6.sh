#!/bin/bash
set -u
TAG=${1:-q6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 300 python tools/bench_topm.py 53 205 ) > $OUT/topm.jsonl 2> $OUT/topm.err; cat $OUT/topm.jsonl; tail -2 $OUT/topm.err
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:topm_stream -s 3 -c 1 -o $OUT/prof_topm python tools/bench_topm.py 53 ) > $OUT/prof_topm.log 2>&1; tail -3 $OUT/prof_topm.log
