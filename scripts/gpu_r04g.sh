#!/bin/bash
OUT=gpurun_out/r04g; mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1" > $OUT/pytest_a.log 2>&1; echo "pytest_a rc=$?"; tail -3 $OUT/pytest_a.log
MPS_B200_DBG_MODE=10 timeout 200 python bench.py --prep random --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peak --no-extras > $OUT/bench_dbg.json 2> $OUT/bench_dbg.err
grep "timing" $OUT/bench_dbg.err
MPS_B200_TRACE=1 timeout 200 python bench.py --prep random --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-peak --no-extras > $OUT/bench_trace.json 2> $OUT/bench_trace.err
grep "trace" $OUT/bench_trace.err | tail -20
