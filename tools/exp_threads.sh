#!/bin/bash
# experiment: beam-step CTA width
OUT=gpurun_out/exp1; mkdir -p $OUT
for th in 256 128 64; do
  FLT_DEC_THREADS=$th timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree_$th.json 2> $OUT/lexfree_$th.err
  FLT_DEC_THREADS=$th timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload lexicon > $OUT/lexicon_$th.json 2> $OUT/lexicon_$th.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp1/*.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], {k:v['ms'] for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
