#!/bin/bash
OUT=gpurun_out/r04h; mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "svd_engine_variants or exact_parity or config1" > $OUT/pytest_a.log 2>&1; echo "pytest_a rc=$?"; tail -3 $OUT/pytest_a.log
MPS_B200_DBG_MODE=10 timeout 200 python bench.py --prep random --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peak --no-extras > $OUT/bench_dbg.json 2> $OUT/bench_dbg.err
grep "timing" $OUT/bench_dbg.err
for Q in 50 14 8; do
for V in "MPS_B200_JACOBI_CLUSTER=1" "MPS_B200_JACOBI_CLUSTER=0"; do
  env $V timeout 200 python bench.py --qubits $Q --prep random --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-peak --no-extras > $OUT/bench_$Q_$V.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$Q_$V.json"))
    print("[q=$Q $V]", "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), d["phases_ms_per_step"])
except Exception as e:
    print("[$V]", "FAILED", e)
PY
done; done
tail -3 $OUT/bench.err
