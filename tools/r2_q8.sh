#!/bin/bash
set -u
TAG=${1:-q8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( FLT_STREAM_WANT=340 timeout 300 python tools/bench_topm.py 205 ) > $OUT/topm205.jsonl 2>/dev/null; cat $OUT/topm205.jsonl
( timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -6 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
( timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
j=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("cfg2", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "frac", round(j["roofline"]["frac"],3), "parity", j["parity"]["exact_match"], "e2e", round(j["e2e"]["value"]), round(j["e2e"]["h2d_gbs_per_gpu"],1), j["e2e"]["host_link"]["h2d_gbs_per_gpu_all_ranks_copying"])
print("  cpu", j["cpu_baseline"]["value"], j["cpu_baseline_bst_beam"]["value"])
s=j["secondary"]; print("cfg3", round(s["value"]), "utt/s", round(s["ms_per_step"],3), {k:round(v["ms"],2) for k,v in s["kernels"].items()}, "parity", s["parity"]["exact_match"], "e2e", round(s["e2e"]["value"]), "cpu", s["cpu_baseline"]["value"], s["cpu_baseline_bst_beam"]["value"])
PY
tail -3 $OUT/bench.err
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json | cut -c1-600; tail -3 $OUT/bench_reference.err
