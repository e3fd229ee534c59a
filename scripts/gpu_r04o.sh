#!/bin/bash
OUT=gpurun_out/r04o; mkdir -p $OUT
nproc > $OUT/gpu.txt; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $OUT/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=10 -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "GHZ-35|passed|failed|FAILED|Error" $OUT/pytest_gpu.log | tail -12
timeout 900 python scripts/configs_fullsize.py --which c3,c5 --fuse3 0 --chi5 256,512 --fuse-both-upto 0 --budget 300 --out $OUT/configs.jsonl > $OUT/configs.log 2>&1; cut -c1-800 $OUT/configs.jsonl
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
for k in ("value","ms_per_step","e2e","e2e_visitor","circuit","phases_ms_per_step"):
    print(k, json.dumps(d.get(k))[:500])
PY
