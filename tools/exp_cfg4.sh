#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 1200 python bench.py --workload lexicon_lm --batch 512 --frames 1500 --threshold 25 --steps 2 --warmup 2 --no-e2e --cpu-seconds 10 > $OUT/cfg4.json 2> $OUT/cfg4.err
tail -5 $OUT/cfg4.err
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['beam_step_work'], d['parity'], d['cpu_baseline'], d['workspace_bytes'])
    except Exception as e: print(f, 'ERR', e)
PY
