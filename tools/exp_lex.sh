#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload lexicon > $OUT/lexicon.json 2> $OUT/lexicon.err
FLT_DEC_THREADS=512 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload lexicon > $OUT/lexicon_512.json 2> $OUT/lexicon_512.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload lexicon --bst 100 > $OUT/lexicon_bst100.json 2> $OUT/lexicon_bst100.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree.json 2> $OUT/lexfree.err
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), {k:round(v['ms'],2) for k,v in d['kernels'].items()}, {k:v for k,v in d['beam_step_work'].items() if k!='phase_cycles_per_frame'}, d['parity']['exact_match'], d['workspace_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
