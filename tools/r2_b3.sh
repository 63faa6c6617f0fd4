#!/bin/bash
# After the lexicon-free table fold (4 barriers per frame) and the plan-time choice of the biased filter:
# whole GPU suite, the default bench line with its cfg 3 / cfg 4 blocks (timed), cfg 4 at 256 threads per utterance.
set -u
TAG=${1:-b3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -8 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
( time timeout 1200 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
try:
    j=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("cfg2", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "frac", round(j["roofline"]["frac"],3), "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "e2e", round(j["e2e"]["value"]), round(j["e2e"]["h2d_gbs_per_gpu"],1), j["e2e"]["host_link"]["h2d_gbs_per_gpu_all_ranks_copying"])
    print("  cpu", j["cpu_baseline"]["value"], j["cpu_baseline_bst_beam"]["value"], "work", j["beam_step_work"])
    s=j["secondary"]; print("cfg3", round(s["value"]), "utt/s", round(s["ms_per_step"],3), {k:round(v["ms"],2) for k,v in s["kernels"].items()}, "parity", s["parity"]["exact_match"], "/", s["parity"]["utterances"], "e2e", round(s["e2e"]["value"]), "cpu", s["cpu_baseline"]["value"], s["cpu_baseline_bst_beam"]["value"])
    s=j["tertiary"]; print("cfg4", round(s["value"]), "utt/s", round(s["ms_per_step"],3), {k:round(v["ms"],2) for k,v in s["kernels"].items()}, "parity", s["parity"]["exact_match"], "/", s["parity"]["utterances"], s["setup"])
except Exception as ex:
    print("bench FAILED", ex)
PY
tail -5 $OUT/bench.err
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
run cfg4_t256 FLT_DEC_THREADS=256 -- $ARGS
run cfg3_t256 FLT_DEC_THREADS=256 -- --workload lexicon --steps 3 --warmup 2
run cfg3_guess -- --workload lexicon --steps 3 --warmup 2
run cfg3_noguess FLT_DBG=8 -- --workload lexicon --steps 3 --warmup 2
run cfg4_guess -- $ARGS
run cfg4_noguess FLT_DBG=8 -- $ARGS
