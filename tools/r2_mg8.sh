#!/bin/bash
# 8 GPUs of one box: multi-GPU tests, cfg 5 (BASELINE configs[4]: LexiconDecoder, 4-gram, beam 500, T = 1500,
# B = 512 per GPU = 4096 over the box) under torchrun, and the headline workload at N = 8 with its e2e leg.
set -u
TAG=${1:-mg8}
G=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_shard_nccl.py -x -q -m gpu -rs 2>&1 | tail -5 ) > $OUT/pytest_nccl.txt; cat $OUT/pytest_nccl.txt
show() {
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$1.json").read().strip().splitlines()[-1])
    print("$1", "n", j["n_gpus"], round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "ties", j["parity"]["excluded_for_ties"], "mismatch", j["parity"]["mismatch"])
    e=j.get("e2e")
    if e: print("   e2e", e.get("value") and round(e["value"]), e.get("h2d_gbs_per_gpu") and round(e["h2d_gbs_per_gpu"],1), e.get("gathered_result_equals_local"), e.get("host_link"))
    print("   work", j["beam_step_work"])
except Exception as ex:
    print("$1 FAILED", ex)
PY
  tail -2 $OUT/$1.err | cut -c1-300
}
( time NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$OUT/nccl_debug.%p.txt FLT_DBG_PLAN=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 2 --warmup 1 --workload lexicon_lm --beam 500 --batch 512 --frames 1500 --threshold 25 --ngrams 2000000,2000000,1000000 --no-e2e --no-cpu-baseline ) > $OUT/bench_cfg5_g$G.json 2> $OUT/bench_cfg5_g$G.err; show bench_cfg5_g$G
cat $OUT/nccl_debug.*.txt | grep -a "NCCL INFO.*\(NVLS\|nranks\|Connected all\)" | head -5 > $OUT/nccl_info.txt; cut -c1-200 $OUT/nccl_info.txt; grep -a "flt plan" $OUT/bench_cfg5_g$G.err | head -1 | cut -c1-330; grep real $OUT/bench_cfg5_g$G.err
grep -av "NCCL INFO" $OUT/bench_cfg5_g$G.err > $OUT/bench_cfg5_g$G.err.short; mv $OUT/bench_cfg5_g$G.err.short $OUT/bench_cfg5_g$G.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline --no-secondary ) > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err; show bench_g$G; grep real $OUT/bench_g$G.err
