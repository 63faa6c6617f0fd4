#!/bin/bash
# staged edge items + hoisted Trie offsets: lexicon parity tests, cfg 3 / cfg 4 / cfg 5 shape (compare with b6)
set -u
TAG=${1:-b7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_fullsize.py tests/test_golden.py tests/test_gpu_random.py tests/test_gpu_fullsize.py tests/test_streaming.py tests/test_decodertest_fixture.py tests/test_gpu_topm.py -x -q -m gpu 2>&1 | tail -4 ) > $OUT/pytest_sel.txt; cat $OUT/pytest_sel.txt
run() { # name env... -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( time env "${envs[@]}" timeout 1200 python bench.py --no-cpu-baseline --no-secondary --no-e2e "$@" ) > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"], "mismatch", j["parity"]["mismatch"])
    w=j["beam_step_work"]; print("    ", w.get("phase_cycles_per_frame"))
except Exception as ex:
    print("$name FAILED", ex)
PY
}
ARGS="--workload lexicon_lm --batch 512 --frames 1500 --threshold 25 --ngrams 2000000,2000000,1000000 --steps 2 --warmup 1"
run cfg3 -- --workload lexicon --steps 3 --warmup 2

run cfg4 -- $ARGS

run cfg5shape -- $ARGS --beam 500 --batch 148 --frames 300

