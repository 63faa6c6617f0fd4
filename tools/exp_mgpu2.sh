#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
G=${2:-2}
nvidia-smi topo -m > $OUT/topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err
tail -1 $OUT/bench_g$G.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'n',d['n_gpus'],'ms',d['ms_per_step'],'e2e',d['e2e'])"
head -12 $OUT/topo.txt | cut -c1-200
