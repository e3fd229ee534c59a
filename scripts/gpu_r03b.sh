#!/bin/bash
# r03b: A/B of the resident-CTA rule of the persistent sweep kernel on the routed (few gates per layer) configs
TAG=${1:-r03b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "MPS_B200_CTAS_PER_SM=4" "MPS_B200_CTAS_PER_SM=0" "MPS_B200_CTAS_PER_SM=1" "MPS_B200_BLOCK16=1"; do
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
