#!/bin/bash
# short bench runs only, one per argument (environment assignment list).  usage: gpurun -- 'bash scripts/gpu_variants.sh TAG "A=1 B=2" ...'
TAG=${1:-var}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for V in "$@"; do
  i=$((i+1))
  env $V timeout 300 python bench.py --prep random --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-peak > $OUT/bench_v$i.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_v$i.json"))
    print("$V", "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), d["phases_ms_per_step"])
except Exception as e:
    print("$V", "FAILED", e)
PY
done
tail -3 $OUT/bench.err
