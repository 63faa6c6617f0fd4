#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
python -c "
import json; d=json.load(open('$OUT/bench_default.json')); print(round(d['value']), d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['parity'])"
tail -3 $OUT/bench_default.err
