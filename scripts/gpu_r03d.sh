#!/bin/bash
TAG=${1:-r03d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "variants or fuse_2q or exact_parity or truncated_gate_against_lapack or config" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_sel.log
for v in "MPS_B200_WIDE_TASKS=0" "MPS_B200_WIDE_TASKS=1"; do
  echo "== $v"
  env $v timeout 120 python scripts/configs_fullsize.py --which c5 --chi5 256 --fuse-both-upto 0 --budget 60 2>> $OUT/err.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('max_bond_dim','run_s','nn_gates_2q_per_s','jacobi_sweeps','launches','norm','amp0_re')})"
done | tee $OUT/c5_ab.txt
echo "== c3 auto"; timeout 120 python scripts/configs_fullsize.py --which c3 --fuse3 0 2>> $OUT/err.log | cut -c1-700 | tee -a $OUT/c5_ab.txt
timeout 300 python bench.py --no-cpu-baseline > $OUT/bench.json 2>> $OUT/err.log; python -c "
import json; d=json.load(open('$OUT/bench.json')); print('bench', d['value'], d['e2e']['value'], d['phases_ms_per_step'], d['circuit']['wall_ms'])"
tail -3 $OUT/err.log
