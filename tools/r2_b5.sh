#!/bin/bash
# race-free table insert (racecheck of the fused kernel again), the 1024-thread step for beam 500, GPU suite.
set -u
TAG=${1:-b5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SEL="cfg2_scaled_bstN or cfg2_scaled_bstK or sil_positive_bst or asg_bst_thr or zero_ctc_bst_thr or long_ragged"
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $OUT/racecheck.txt 2>&1; tail -3 $OUT/racecheck.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "very_wide" > $OUT/memcheck_wide.txt 2>&1; tail -3 $OUT/memcheck_wide.txt
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
run() { # name env... -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( time env "${envs[@]}" FLT_DBG_PLAN=1 timeout 1200 python bench.py --no-cpu-baseline --no-secondary --no-e2e "$@" ) > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"], "mismatch", j["parity"]["mismatch"])
    w=j["beam_step_work"]; print("    ", w)
except Exception as ex:
    print("$name FAILED", ex)
PY
  grep -a "flt plan" $OUT/$name.err | tail -1 | cut -c1-330; grep real $OUT/$name.err
}
ARGS="--workload lexicon_lm --batch 512 --frames 1500 --threshold 25 --ngrams 2000000,2000000,1000000 --steps 2 --warmup 1"
run cfg2 -- --steps 5 --warmup 3
run cfg5shape_1024 -- $ARGS --beam 500 --batch 148 --frames 300
run cfg5shape_512 FLT_NO_1024=1 -- $ARGS --beam 500 --batch 148 --frames 300
