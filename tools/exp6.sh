#!/bin/bash
OUT=gpurun_out/exp11; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree_fused.json 2> $OUT/lexfree_fused.err
FLT_NO_FUSED=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree_2k.json 2> $OUT/lexfree_2k.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --bst 50 > $OUT/lexfree_fused_bst50.json 2> $OUT/lexfree_fused_bst50.err
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --batch 1024 --frames 500 > $OUT/lexfree_fused_b1024.json 2> $OUT/lexfree_fused_b1024.err
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "cfg2_scaled_bstN or cfg2_scaled_bstK or asg" > $OUT/memcheck.txt 2>&1; tail -4 $OUT/memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "cfg2_scaled_bstN or sil_positive_bst" > $OUT/racecheck.txt 2>&1; tail -4 $OUT/racecheck.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flt_k_fused -s 1 -c 1 -o $OUT/prof_fused python bench.py --steps 1 --warmup 1 --frames 250 --no-e2e --no-cpu-baseline > $OUT/prof.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp11/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['roofline']['frac'], d['beam_step_work'], d['parity']['exact_match'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
