#!/bin/bash
OUT=gpurun_out/exp5; mkdir -p $OUT
for th in 256 512; do
  FLT_DEC_THREADS=$th timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree_$th.json 2> $OUT/lexfree_$th.err
done
FLT_DEC_THREADS=512 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --sigma 4 --threshold 25 > $OUT/lexfree_s4thr_512.json 2> $OUT/lexfree_s4thr_512.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flt_k_decode -s 1 -c 1 -o $OUT/prof_lf python bench.py --steps 1 --warmup 1 --frames 250 --no-e2e --no-cpu-baseline > $OUT/prof.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp5/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['beam_step_work'], d['parity']['exact_match'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-500:])
PY
