#!/bin/bash
# quick validation after a kernel change: the SVD-facing parity tests, the lone-gate regime, the headline step
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "svd_engine_variants or exact_parity or config1 or full_size or config2_full_shape_parity or golden or sharded or config3_qaoa or config5_syc" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 300 python scripts/lone_gate_bench.py 256 512 1024 2> $OUT/lone.err | grep '"jacobi_cluster": 0' > $OUT/lone.jsonl; cut -c1-220 $OUT/lone.jsonl
timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-peak --no-extras > $OUT/bench.json 2>> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.1f ms/step %.2f circuit %.1f" % (d["value"], d["ms_per_step"], d["circuit"]["wall_ms"]), d["phases_ms_per_step"])
PY
