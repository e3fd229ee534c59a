#!/bin/bash
# parity tests + bench with A/B switches.  usage: gpurun --timeout 1200 -- 'bash scripts/gpu_ab.sh TAG'
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 $OUT/sanitizer.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 300 python bench.py --no-qr --no-e2e --no-cpu-baseline --no-peak > $OUT/bench_noqr.json 2>> $OUT/bench.err; cat $OUT/bench_noqr.json
