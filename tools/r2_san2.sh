#!/bin/bash
# compute-sanitizer on the final build's lexicon step (compacted second pass, hoisted Trie offsets, 1024 threads)
OUT=gpurun_out/${1:-san4}; mkdir -p $OUT
SEL="cfg3_scaled_bstN or cfg3_mid or arpa3_ctc or arpa_unk or zero_unk or cfg4_scaled or zero_asg"
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $OUT/racecheck_lexicon.txt 2>&1; tail -3 $OUT/racecheck_lexicon.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL or very_wide or split" > $OUT/memcheck_lexicon.txt 2>&1; tail -3 $OUT/memcheck_lexicon.txt
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "very_wide" > $OUT/racecheck_wide.txt 2>&1; tail -3 $OUT/racecheck_wide.txt
