#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
for b in 148 256 296; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --batch $b > $OUT/lexfree_b$b.json 2> $OUT/lexfree_b$b.err
done
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['beam_step_work'].get('phase_cycles_per_frame'), d['parity']['exact_match'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
