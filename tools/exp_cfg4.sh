#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
for kb in 110 227; do
FLT_SMEM_KB=$kb timeout 1200 python bench.py --workload lexicon_lm --batch 512 --frames 1500 --threshold 25 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > $OUT/cfg4_$kb.json 2> $OUT/cfg4_$kb.err
done
FLT_SMEM_KB=227 FLT_DEC_THREADS=256 timeout 1200 python bench.py --workload lexicon_lm --batch 512 --frames 1500 --threshold 25 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > $OUT/cfg4_227_t256.json 2> $OUT/cfg4_227_t256.err
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['parity']['exact_match'], d['workspace_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
