#!/bin/bash
OUT=gpurun_out/r04f; mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "svd_engine_variants or exact_parity or config1" > $OUT/pytest_a.log 2>&1; echo "pytest_a rc=$?"; tail -5 $OUT/pytest_a.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" > $OUT/pytest_b.log 2>&1; echo "pytest_b rc=$?"; tail -5 $OUT/pytest_b.log
for V in "MPS_B200_JACOBI_CLUSTER=1" "MPS_B200_JACOBI_CLUSTER=0"; do
  env $V timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-peak --no-extras > $OUT/bench_$V.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$V.json"))
    print("[$V]", "value %.1f ms/step %.2f circuit %.0f ms" % (d["value"], d["ms_per_step"], d["circuit"]["wall_ms"]), d["phases_ms_per_step"])
except Exception as e:
    print("[$V]", "FAILED", e)
PY
done
tail -3 $OUT/bench.err
