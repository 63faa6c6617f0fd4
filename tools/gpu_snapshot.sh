#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list and one full capture of each kernel.
# usage: tools/gpu_snapshot.sh <tag>
set -u
TAG=${1:-snap}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.txt
( timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_lexfree.json 2> $OUT/bench_lexfree.err
( timeout 600 python bench.py --steps 5 --warmup 3 --bst 50 ) > $OUT/bench_lexfree_bst50.json 2> $OUT/bench_lexfree_bst50.err
( timeout 600 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e ) > $OUT/bench_lexicon.json 2> $OUT/bench_lexicon.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_ -s 3 -c 3 \
  -o $OUT/prof python bench.py --steps 1 --warmup 1 --frames 250 --no-e2e --no-cpu-baseline > $OUT/prof.log 2>&1
python - <<'PY' > $OUT/pcie.txt 2>&1
import torch, time
x = torch.empty(1 << 30, dtype=torch.float32, pin_memory=True)  # 4 GiB
d = torch.empty_like(x, device="cuda")
for chunk in (1 << 30, 1 << 28, 1 << 26):
    torch.cuda.synchronize(); t = time.perf_counter()
    for o in range(0, x.numel(), chunk):
        d[o:o + chunk].copy_(x[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"H2D pinned chunk={chunk * 4 >> 20} MiB: {x.numel() * 4 / dt / 1e9:.1f} GB/s")
torch.cuda.synchronize(); t = time.perf_counter(); x.copy_(d, non_blocking=True); torch.cuda.synchronize()
print(f"D2H pinned: {x.numel() * 4 / (time.perf_counter() - t) / 1e9:.1f} GB/s")
PY
ls -la $OUT
