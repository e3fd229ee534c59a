#!/bin/bash
# parity tests + bench with A/B switches.  usage: gpurun --timeout 1200 -- 'bash scripts/gpu_ab.sh TAG "ENV1=.. ENV2=.." ...'
# every extra argument is an environment assignment list for one more short bench run (no e2e / cpu baseline)
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
i=0
for V in "$@"; do
  i=$((i+1))
  echo "== variant $i: $V"
  env $V timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-peak > $OUT/bench_v$i.json 2>> $OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_v$i.json"))
print("$V", "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), d["phases_ms_per_step"])
PY
done
