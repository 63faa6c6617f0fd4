#!/bin/bash
# LM bound for token-level LMs in the full expansion: widened parity tests, the token-LM bench with the bound
# (and, FLT_NO_PRUNE2_FULL=1, the old single-pass full expansion), fresh ncu capture of the fused kernel.
set -u
TAG=${1:-b9}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -3 ) > $OUT/pytest_sel.txt; cat $OUT/pytest_sel.txt
run() { # name env... -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --no-secondary --no-e2e "$@" ) > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"], "mismatch", j["parity"]["mismatch"])
    print("    ", j["beam_step_work"].get("candidates_per_frame"), j["beam_step_work"].get("phase_cycles_per_frame"))
except Exception as ex:
    print("$name FAILED", ex)
PY
}
run tokenlm_bst50 -- --workload lexfree_tokenlm --bst 50 --threshold 25 --steps 3 --warmup 2
run tokenlm_bst50_nop2 FLT_NO_PRUNE2_FULL=1 -- --workload lexfree_tokenlm --bst 50 --threshold 25 --steps 3 --warmup 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flt_k_fused -s 1 -c 1 \
  -o $OUT/prof python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $OUT/prof.log 2>&1
ls -la $OUT | tail -4
