#!/bin/bash
set -u
TAG=${1:-q7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 300 python tools/bench_topm.py 53 105 ) > $OUT/topm.jsonl 2> $OUT/topm.err; cat $OUT/topm.jsonl; tail -2 $OUT/topm.err
( FLT_STREAM_WANT=340 timeout 300 python tools/bench_topm.py 205 ) > $OUT/topm205.jsonl 2>/dev/null; cat $OUT/topm205.jsonl
( timeout 900 python -m pytest tests/test_gpu_topm.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 ) > $OUT/pytest.txt; cat $OUT/pytest.txt
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
j=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("cfg2", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "frac", round(j["roofline"]["frac"],3), "parity", j["parity"], "e2e", round(j["e2e"]["value"]), j["e2e"]["host_link"])
s=j["secondary"]; print("cfg3", round(s["value"]), "utt/s", round(s["ms_per_step"],3), {k:round(v["ms"],2) for k,v in s["kernels"].items()}, "parity", s["parity"]["exact_match"], "e2e", round(s["e2e"]["value"]))
print(j["beam_step_work"])
PY
tail -3 $OUT/bench.err
