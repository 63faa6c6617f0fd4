#!/bin/bash
set -u
TAG=${1:-q2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( FLT_DBG_PLAN=1 timeout 300 python tools/r2_dbg.py lexfree ) > $OUT/dbg_lexfree.json 2> $OUT/dbg_lexfree.err
( FLT_DBG_PLAN=1 timeout 300 python tools/r2_dbg.py lexicon ) > $OUT/dbg_lexicon.json 2> $OUT/dbg_lexicon.err
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $OUT/pytest_gpu.txt
( FLT_DBG_PLAN=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > $OUT/bench_lexfree.json 2> $OUT/bench_lexfree.err
( FLT_DBG_PLAN=1 timeout 600 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/bench_lexicon.json 2> $OUT/bench_lexicon.err
cat $OUT/dbg_lexfree.json $OUT/dbg_lexicon.json; tail -3 $OUT/dbg_lexfree.err
tail -3 $OUT/pytest_gpu.txt
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
