#!/bin/bash
OUT=gpurun_out/r04r; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" $OUT/pytest_gpu.log | tail -8
timeout 300 python scripts/sharded_abi_check.py --gpus 2 --qubits 50 --depth 20 --chi 256 --partition-by count > $OUT/shard_c2.json 2> $OUT/shard.err; cut -c1-600 $OUT/shard_c2.json
timeout 600 python bench.py --no-cpu-baseline --no-extras > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
for k in ("value","ms_per_step","e2e","circuit","phases_ms_per_step"):
    print(k, json.dumps(d.get(k))[:400])
PY
