#!/bin/bash
# Round-2 evidence snapshot on one B200: GPU suite, default bench line (cfg 2 + cfg 3 secondary), reference arm,
# stand-alone select roofline, launch list, one full ncu capture of the dominant kernel, smoke.
# usage: tools/r2_snap.sh <tag>
set -u
TAG=${1:-snap}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -12 ) > $OUT/pytest_gpu.txt; tail -4 $OUT/pytest_gpu.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.txt 2>&1; tail -1 $OUT/smoke.txt
( timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
try:
    j=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("cfg2", round(j["value"]), "utt/s", round(j["ms_per_step"],3), "ms", {k:round(v["ms"],2) for k,v in j["kernels"].items()}, "frac", round(j["roofline"]["frac"],3), "parity", j["parity"]["exact_match"], "/", j["parity"]["utterances"], "e2e", round(j["e2e"]["value"]), round(j["e2e"]["h2d_gbs_per_gpu"],1), j["e2e"]["host_link"]["h2d_gbs_per_gpu_all_ranks_copying"])
    print("  cpu", j["cpu_baseline"]["value"], j["cpu_baseline_bst_beam"]["value"], "work", j["beam_step_work"])
    s=j["secondary"]; print("cfg3", round(s["value"]), "utt/s", round(s["ms_per_step"],3), {k:round(v["ms"],2) for k,v in s["kernels"].items()}, "parity", s["parity"]["exact_match"], "/", s["parity"]["utterances"], "e2e", round(s["e2e"]["value"]), "cpu", s["cpu_baseline"]["value"], s["cpu_baseline_bst_beam"]["value"])
    print("  work", s["beam_step_work"])
except Exception as ex:
    print("bench FAILED", ex)
PY
tail -3 $OUT/bench.err
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cut -c1-400 $OUT/bench_reference.json
( timeout 300 python tools/bench_topm.py 53 105 205 ) > $OUT/topm.jsonl 2>/dev/null; cat $OUT/topm.jsonl
# launch list of the default bench command (per-launch durations, cold cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $OUT/launches.log 2>&1
# full capture of the dominant kernel at the benchmark's full size (traffic per launch) + source counters
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_fused -s 1 -c 1 \
  -o $OUT/prof python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $OUT/prof.log 2>&1
du -sh $OUT; ls -la $OUT
# the lexicon step (cfg 3, T = 100): full capture for the source-level counters of the north star's target kernel
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:flt_k_decode512 -s 1 -c 1 \
  -o $OUT/prof_lexicon python bench.py --steps 1 --warmup 1 --frames 100 --workload lexicon --no-e2e --no-cpu-baseline ) > $OUT/prof_lexicon.log 2>&1
ls -la $OUT | tail -5
