#!/bin/bash
OUT=gpurun_out/exp2; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.txt
for th in 256 512; do
  FLT_DEC_THREADS=$th timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/lexfree_$th.json 2> $OUT/lexfree_$th.err
  FLT_DEC_THREADS=$th timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --sigma 4 > $OUT/lexfree_s4_$th.json 2> $OUT/lexfree_s4_$th.err
  FLT_DEC_THREADS=$th timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --bst 50 > $OUT/lexfree_bst50_$th.json 2> $OUT/lexfree_bst50_$th.err
done
FLT_DEC_THREADS=512 timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload lexicon > $OUT/lexicon_512.json 2> $OUT/lexicon_512.err
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "cfg1_shape or sil_positive_bst or asg_bst_thr" > $OUT/racecheck.txt 2>&1
tail -5 $OUT/racecheck.txt
cat $OUT/pytest_gpu.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp2/*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['beam_step_work'], d['parity']['exact_match'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-500:])
PY
