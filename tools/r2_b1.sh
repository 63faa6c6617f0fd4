#!/bin/bash
# Round-2 check after: keyed (biased) select + long-list variant, LM upper bound before n-gram probes,
# pinned result buffers in the e2e leg, NCCL-group side-effect experiment.
set -u
TAG=${1:-b1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_topm.py tests/test_gpu_parity.py tests/test_gpu_golden_fullsize.py tests/test_golden.py tests/test_gpu_random.py -x -q -m gpu 2>&1 | tail -8 ) > $OUT/pytest_sel.txt; cat $OUT/pytest_sel.txt
( timeout 300 python tools/bench_topm.py 53 105 205 505 ) > $OUT/topm.jsonl 2>/dev/null; cut -c1-200 $OUT/topm.jsonl
run() { # name env... -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( time env "${envs[@]}" FLT_DBG_PLAN=1 timeout 1200 python bench.py --no-cpu-baseline --no-secondary "$@" ) > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"], "mismatch", j["parity"]["mismatch"])
    e=j.get("e2e")
    if e: print("    e2e", round(e["value"]), round(e["h2d_gbs_per_gpu"],1), e["host_link"])
    w=j["beam_step_work"]; print("    ", w)
except Exception as ex:
    print("$name FAILED", ex)
PY
  grep -a "flt plan" $OUT/$name.err | tail -1 | cut -c1-260; grep real $OUT/$name.err
}
ARGS="--workload lexicon_lm --batch 512 --frames 1500 --threshold 25 --ngrams 2000000,2000000,1000000 --no-e2e --steps 2 --warmup 1"
run cfg4 -- $ARGS
run cfg5shape -- $ARGS --beam 500 --batch 148 --frames 300
run cfg2_e2e -- --steps 5 --warmup 3
run cfg2_pg BENCH_FORCE_PG=1 -- --steps 5 --warmup 3 --no-e2e
run cfg2_pg_noclk BENCH_FORCE_PG=1 BENCH_NO_CLOCKS=1 -- --steps 5 --warmup 3 --no-e2e
run cfg2_pg_20 BENCH_FORCE_PG=1 -- --steps 20 --warmup 3 --no-e2e
