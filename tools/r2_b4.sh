#!/bin/bash
# Lexicon-free guessed pruning bound: parity tests, cfg 2 with / without it, peaky emissions.
set -u
TAG=${1:-b4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_fullsize.py tests/test_golden.py tests/test_gpu_random.py tests/test_gpu_fullsize.py tests/test_streaming.py tests/test_gpu_pybind.py -x -q -m gpu 2>&1 | tail -5 ) > $OUT/pytest_sel.txt; cat $OUT/pytest_sel.txt
run() { # name env... -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( time env "${envs[@]}" FLT_DBG_PLAN=1 timeout 1200 python bench.py --no-cpu-baseline --no-secondary --no-e2e "$@" ) > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "frac", round(j["roofline"]["frac"],3), "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"], "mismatch", j["parity"]["mismatch"])
    w=j["beam_step_work"]; print("    ", w)
except Exception as ex:
    print("$name FAILED", ex)
PY
  grep real $OUT/$name.err
}
run cfg2_guess -- --steps 5 --warmup 3
run cfg2_noguess FLT_DBG=16 -- --steps 5 --warmup 3
run cfg2_guess_sigma4 -- --steps 5 --warmup 3 --sigma 4
run cfg2_noguess_sigma4 FLT_DBG=16 -- --steps 5 --warmup 3 --sigma 4
run cfg2_guess_bst50 -- --steps 5 --warmup 3 --bst 50
run cfg3 -- --workload lexicon --steps 3 --warmup 2
