#!/bin/bash
# full-expansion modes of the lexicon-free decoder with 512 instead of 256 threads per utterance
OUT=gpurun_out/${1:-thr1}; mkdir -p $OUT
for t in 256 512; do
  ( FLT_DEC_THREADS=$t timeout 200 python bench.py --steps 3 --warmup 2 --workload lexfree_tokenlm --bst 50 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/tokenlm_$t.json 2> $OUT/tokenlm_$t.err
  ( FLT_DEC_THREADS=$t timeout 200 python bench.py --steps 3 --warmup 2 --log-add --bst 50 --threshold 25 --no-e2e --no-cpu-baseline ) > $OUT/logadd_$t.json 2> $OUT/logadd_$t.err
done
for f in $OUT/*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', round(d['value'],1), {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['parity']['exact_match'])"; done
