#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flt_k_fused -s 1 -c 1 -o $OUT/prof python bench.py --steps 1 --warmup 1 --frames 250 --no-e2e --no-cpu-baseline > $OUT/prof.log 2>&1
ls -la $OUT
