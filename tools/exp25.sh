#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree.json 2> $OUT/lexfree.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --sigma 4 > $OUT/lexfree_s4.json 2> $OUT/lexfree_s4.err
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg2_scaled_bstN or sil_positive_bst or masked" > $OUT/racecheck.txt 2>&1; tail -2 $OUT/racecheck.txt
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['beam_step_work'].get('phase_cycles_per_frame'), d['beam_step_work'].get('select_guess_misses'), d['parity']['exact_match'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
