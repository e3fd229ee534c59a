#!/bin/bash
# new full-shape parity tests + both bench arms.  usage: gpurun --timeout 1500 -- 'bash scripts/gpu_tests_new.sh TAG'
TAG=${1:-t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nproc > $OUT/gpu.txt; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -k "config2_full or config3_full or config4_hea64 or config5_real or sampling" --durations=8 > $OUT/pytest_new.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_new.log
tail -25 $OUT/pytest_new.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 400 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
