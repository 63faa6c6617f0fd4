#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --workload lexicon --beam 500 --batch 64 --frames 200 > $OUT/lexicon_k500.json 2> $OUT/lexicon_k500.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --beam 500 --batch 64 --frames 200 > $OUT/lexfree_k500.json 2> $OUT/lexfree_k500.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --beam 100 > $OUT/lexfree_k100.json 2> $OUT/lexfree_k100.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --sigma 4 --threshold 25 > $OUT/lexfree_s4_thr25.json 2> $OUT/lexfree_s4_thr25.err
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['parity']['exact_match'], d['parity']['excluded_for_ties'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-500:])
PY
