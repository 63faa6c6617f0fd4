#!/bin/bash
# 2 GPUs: the NCCL sharded-decode tests, then the default bench line at N = 2 (weak scaling) next to N = 1
set -u
TAG=${1:-mg2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_shard_nccl.py -x -q -m gpu -rs 2>&1 | tail -5 ) > $OUT/pytest_nccl.txt; cat $OUT/pytest_nccl.txt
show() {
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$1.json").read().strip().splitlines()[-1])
    print("$1", "n", j["n_gpus"], round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "e2e", j["e2e"] and round(j["e2e"]["value"]), j["e2e"] and j["e2e"].get("gathered_result_equals_local"), j["e2e"] and j["e2e"].get("host_link"))
    for k in ("secondary", "tertiary"):
        s=j.get(k)
        if s: print("  ", k, round(s["value"]), round(s["ms_per_step"],3), "parity", s["parity"]["exact_match"], "e2e", s.get("e2e") and s["e2e"].get("value") and round(s["e2e"]["value"]), s.get("e2e") and s["e2e"].get("gathered_result_equals_local"))
except Exception as ex:
    print("$1 FAILED", ex)
PY
  tail -2 $OUT/$1.err | cut -c1-300
}
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline ) > $OUT/bench_g2.json 2> $OUT/bench_g2.err; show bench_g2
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary ) > $OUT/bench_g1.json 2> $OUT/bench_g1.err; show bench_g1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e ) > $OUT/bench_g2_20.json 2> $OUT/bench_g2_20.err; show bench_g2_20
