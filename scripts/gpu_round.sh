#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the reference arm, the chi=512 theta micro-measurement.
# usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh rNN'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 300 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
# north_star: theta contraction at chi >= 512 against the FP64 tensor peak (24-qubit chain, random saturated sites)
timeout 300 python bench.py --qubits 24 --chi 512 --prep random --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_chi512.json 2>> $OUT/bench.err
python - <<PY
import json
d = json.load(open("$OUT/bench_chi512.json"))
print("chi512:", "value %.1f gates/s" % d["value"], "theta", d["roofline_theta"], "phases", d["phases_ms_per_step"])
PY
ls -la $OUT
