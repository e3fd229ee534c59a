#!/bin/bash
OUT=gpurun_out/r04j; mkdir -p $OUT
nproc > $OUT/gpu.txt; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $OUT/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=15 -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "GHZ-35|passed|failed|FAILED|Error" $OUT/pytest_gpu.log | tail -30
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
for k in ("value","ms_per_step","e2e","roofline","roofline_theta","roofline_theta_chi512","e2e_visitor","circuit","cpu_baseline","phases_ms_per_step"):
    print(k, json.dumps(d.get(k))[:600])
PY
timeout 400 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref.json
timeout 300 python scripts/site_kernels_bench.py > $OUT/site_kernels.json 2> $OUT/site_kernels.err; tail -5 $OUT/site_kernels.json | cut -c1-900
