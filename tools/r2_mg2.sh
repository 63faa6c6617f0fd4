#!/bin/bash
# 2 GPUs: the NCCL sharded-decode test, then bench at N = 2 with and without the pre-step
set -u
TAG=${1:-mg2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_shard_nccl.py -x -q -m gpu -rs 2>&1 | tail -5 ) > $OUT/pytest_nccl.txt; cat $OUT/pytest_nccl.txt
run() { # name, env
  ( env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline $3 ) > $OUT/$1.json 2> $OUT/$1.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$1.json").read().strip().splitlines()[-1])
    print("$1", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "e2e", j["e2e"] and round(j["e2e"]["value"]), j["e2e"] and j["e2e"].get("gathered_result_equals_local"), j["e2e"] and j["e2e"].get("host_link"))
    s=j.get("secondary")
    if s: print("   cfg3", round(s["value"]), round(s["ms_per_step"],3), "e2e", s["e2e"] and round(s["e2e"]["value"]), s["e2e"] and s["e2e"].get("gathered_result_equals_local"))
except Exception as ex:
    print("$1 FAILED", ex)
PY
  tail -2 $OUT/$1.err
}
run bench_g2 "A=1" ""
run bench_g2_noprestep "BENCH_NO_PRESTEP=1" "--no-e2e --no-secondary"
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary ) > $OUT/bench_g1.json 2> $OUT/bench_g1.err
python - <<PY
import json
j=json.loads(open("$OUT/bench_g1.json").read().strip().splitlines()[-1])
print("g1", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), round(j["e2e"]["h2d_gbs_per_gpu"],1), j["e2e"]["host_link"]["h2d_gbs_per_gpu_all_ranks_copying"])
PY
