#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
G=${2:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $G --steps 1 --warmup 1 > $OUT/ref_g$G.json 2> $OUT/ref_g$G.err
tail -3 $OUT/bench_g$G.err; cat $OUT/bench_g$G.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'n',d['n_gpus'],'ms',d['ms_per_step'],'e2e',d['e2e'])"
tail -2 $OUT/ref_g$G.json | cut -c1-300
