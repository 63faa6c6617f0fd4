#!/bin/bash
# cfg 5 again on 8 GPUs with the final build (1024-thread step for beam 500)
set -u
TAG=${1:-mg8b}
G=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$OUT/nccl_debug.%p.txt FLT_DBG_PLAN=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 2 --warmup 1 --workload lexicon_lm --beam 500 --batch 512 --frames 1500 --threshold 25 --ngrams 2000000,2000000,1000000 --no-e2e --no-cpu-baseline ) > $OUT/bench_cfg5_g$G.json 2> $OUT/bench_cfg5_g$G.err
python - <<PY
import json
try:
    j=json.loads(open("$OUT/bench_cfg5_g$G.json").read().strip().splitlines()[-1])
    print("cfg5", "n", j["n_gpus"], round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"])
    print("   work", j["beam_step_work"])
except Exception as ex:
    print("cfg5 FAILED", ex)
PY
cat $OUT/nccl_debug.*.txt | grep -a "NVLS multicast\|Init COMPLETE" | head -10 | cut -c1-220 > $OUT/nccl_info.txt; rm -f $OUT/nccl_debug.*.txt; head -3 $OUT/nccl_info.txt; grep -a "flt plan" $OUT/bench_cfg5_g$G.err | tail -1 | cut -c1-330; grep real $OUT/bench_cfg5_g$G.err
