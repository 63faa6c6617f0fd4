#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree_c256.json 2> $OUT/lexfree_c256.err
FLT_LIB=$PWD/text_b200/lib/libflt_decoder_c320.so timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree_c320.json 2> $OUT/lexfree_c320.err
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['beam_step_work'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
