#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch list and one full capture of each kernel.
# usage: tools/gpu_snapshot.sh <tag>
set -u
TAG=${1:-snap}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.txt
( timeout 600 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_lexfree.json 2> $OUT/bench_lexfree.err
( timeout 600 python bench.py --steps 5 --warmup 3 --sigma 4 --no-e2e --no-cpu-baseline ) > $OUT/bench_lexfree_sigma4.json 2> $OUT/bench_lexfree_sigma4.err
( timeout 600 python bench.py --steps 5 --warmup 3 --bst 50 ) > $OUT/bench_lexfree_bst50.json 2> $OUT/bench_lexfree_bst50.err
( timeout 600 python bench.py --steps 3 --warmup 3 --workload lexicon --no-e2e ) > $OUT/bench_lexicon.json 2> $OUT/bench_lexicon.err
( FLT_NO_FUSED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline ) > $OUT/bench_lexfree_twokernel.json 2> $OUT/bench_lexfree_twokernel.err
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err
# full-expansion modes on the same build
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexfree_tokenlm --bst 50 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/bench_lexfree_tokenlm_bst50.json 2> $OUT/bench_lexfree_tokenlm_bst50.err
( timeout 300 python bench.py --steps 3 --warmup 3 --log-add --bst 50 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/bench_lexfree_logadd_bst50.json 2> $OUT/bench_lexfree_logadd_bst50.err
( timeout 300 python bench.py --steps 3 --warmup 3 --workload lexicon --log-add --bst 100 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/bench_lexicon_logadd_bst100.json 2> $OUT/bench_lexicon_logadd_bst100.err
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.txt 2>&1
# launch list of the default bench command (per-launch durations, cold cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/launches.log 2>&1
# full capture of the dominant kernel at the benchmark's full size (traffic per launch) + source counters
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_ -s 2 -c 2 \
  -o $OUT/prof python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/prof.log 2>&1
# (gpurun copies back at most 64 MiB: one kernel per extra capture; the two-kernel capture of snapshot s4 stands)
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_decode -s 1 -c 1 \
  -o $OUT/prof_lexicon python bench.py --steps 1 --warmup 1 --frames 100 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/prof_lexicon.log 2>&1
du -sh $OUT; ls -la $OUT
