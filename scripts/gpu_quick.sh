#!/bin/bash
# default bench (no e2e/cpu) + variants; circuit prep.  usage: gpurun -- 'bash scripts/gpu_quick.sh TAG "" "A=1" ...'
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
i=0
for V in "$@"; do
  i=$((i+1))
  env $V timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-peak > $OUT/bench_v$i.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_v$i.json"))
    print("[$V]", "value %.1f ms/step %.2f circuit %.0f ms" % (d["value"], d["ms_per_step"], d["circuit"]["wall_ms"]), d["phases_ms_per_step"])
except Exception as e:
    print("[$V]", "FAILED", e)
PY
done
grep "phase timing" $OUT/bench.err | tail -2; tail -2 $OUT/bench.err
