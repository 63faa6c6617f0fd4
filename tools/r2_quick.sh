#!/bin/bash
# Round-2 quick check on one B200: GPU suite, the two headline bench lines (cfg 2 default, cfg 3), plan print.
# usage: tools/r2_quick.sh <tag> [pytest-args]
set -u
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q ${2:-} 2>&1 | tail -25 ) > $OUT/pytest_gpu.txt
( FLT_DBG_PLAN=1 timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_lexfree.json 2> $OUT/bench_lexfree.err
( FLT_DBG_PLAN=1 timeout 600 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e ) > $OUT/bench_lexicon.json 2> $OUT/bench_lexicon.err
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.txt 2>&1
tail -3 $OUT/pytest_gpu.txt; cat $OUT/smoke.txt | tail -2
python - <<PY
import json
for n in ("lexfree","lexicon"):
    try:
        j=json.loads(open("$OUT/bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(j["value"]), "utt/s", j["ms_per_step"], "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "e2e", (j.get("e2e") or {}).get("value"))
        print("   work", j["beam_step_work"])
    except Exception as ex:
        print(n, "FAILED", ex)
PY
grep -a "flt plan" $OUT/*.err | sort | uniq -c | head
