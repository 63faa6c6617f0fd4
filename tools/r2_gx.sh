#!/bin/bash
# gx (two-pass step, FLT_GX=1) vs the default kernels on cfg 2 / cfg 3 (device-timed only)
set -u
TAG=${1:-gx}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { # name env... -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" FLT_DBG_PLAN=1 timeout 900 python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 3 --warmup 2 "$@" ) > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), "utt/s", round(j["ms_per_step"],2), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"])
    w=j["beam_step_work"]; print("    ", w)
except Exception as ex:
    print("$name FAILED", ex)
PY
  grep -a "flt plan" $OUT/$name.err | tail -1 | cut -c1-260
}
run cfg3_gx FLT_GX=1 -- --workload lexicon
run cfg2_gx FLT_GX=1 --
run cfg3_b512 -- --workload lexicon --batch 512
run cfg2_b512 -- --batch 512 --frames 500
