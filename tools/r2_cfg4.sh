#!/bin/bash
# cfg 4 (BASELINE configs[3]: LexiconDecoder, 200k-word Trie + 4-gram, beam 200, beamThreshold 25, T=1500, B=512)
# at the SURVEY-sized LM (2M/2M/1M 2/3/4-grams), device-timed, with the variants of the step kernel.
set -u
TAG=${1:-cfg4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { # name env... -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( time env "${envs[@]}" FLT_DBG_PLAN=1 timeout 1200 python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 2 --warmup 1 "$@" ) > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), "utt/s", round(j["ms_per_step"],2), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"], "mismatch", j["parity"]["mismatch"])
    w=j["beam_step_work"]; print("    ", w); print("    setup", j["setup"])
except Exception as ex:
    print("$name FAILED", ex)
PY
  grep -a "flt plan" $OUT/$name.err | tail -1 | cut -c1-260; grep real $OUT/$name.err
}
ARGS="--workload lexicon_lm --batch 512 --frames 1500 --threshold 25 --ngrams 2000000,2000000,1000000"
run cfg4_default -- $ARGS
run cfg4_smem224 FLT_SMEM_KB=224 -- $ARGS
run cfg4_gx FLT_GX=1 -- $ARGS
run cfg5shape_smem224 FLT_SMEM_KB=224 -- $ARGS --beam 500 --batch 148 --frames 300
run cfg5shape_default -- $ARGS --beam 500 --batch 148 --frames 300
