#!/bin/bash
# compute-sanitizer over a representative subset of the GPU tests (memcheck, racecheck, synccheck).
OUT=gpurun_out/${1:-san}; mkdir -p $OUT
SEL="cfg2_scaled_bstN or cfg2_scaled_bstK or sil_positive_bst or asg_bst_thr or zero_ctc_bst_thr or arpa3_ctc or zero_unk or long_ragged or masked"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_streaming.py -x -q -m gpu -k "$SEL or streaming_cuda" > $OUT/memcheck.txt 2>&1; tail -3 $OUT/memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $OUT/racecheck.txt 2>&1; tail -3 $OUT/racecheck.txt
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg2_scaled_bstN or zero_ctc_bst_thr or long_ragged" > $OUT/synccheck.txt 2>&1; tail -3 $OUT/synccheck.txt
