#!/bin/bash
# gx (two-pass step) vs the round-1 kernels on cfg 3 / cfg 4 / cfg 5 shapes
set -u
TAG=${1:-c1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { # name env... -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" FLT_DBG_PLAN=1 timeout 900 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 1 "$@" ) > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), "utt/s", round(j["ms_per_step"],2), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"])
    w=j["beam_step_work"]; print("    cand/frame", round(w["candidates_per_frame"]), w.get("phase_cycles_per_frame"), w.get("redo_events"))
except Exception as ex:
    print("$name FAILED", ex)
PY
  grep -a "flt plan" $OUT/$name.err | tail -1 | cut -c1-200
}
run cfg4_old -- --workload lexicon_lm --batch 256 --frames 300 --threshold 25
run cfg4_old_smem FLT_SMEM_KB=224 -- --workload lexicon_lm --batch 256 --frames 300 --threshold 25
run cfg4_gx FLT_GX=1 -- --workload lexicon_lm --batch 256 --frames 300 --threshold 25
run cfg3_gx FLT_GX=1 -- --workload lexicon
run cfg5_old_smem FLT_SMEM_KB=224 -- --workload lexicon_lm --batch 148 --frames 200 --threshold 25 --beam 500
run cfg5_gx FLT_GX=1 -- --workload lexicon_lm --batch 148 --frames 200 --threshold 25 --beam 500
run cfg2_gx FLT_GX=1 --
