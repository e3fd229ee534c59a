#!/bin/bash
TAG=${1:-r03e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 python scripts/configs_fullsize.py --which c5 --chi5 256 --fuse-both-upto 0 --budget 60 2>> $OUT/err.log | cut -c1-420
timeout 300 python bench.py --no-cpu-baseline > $OUT/bench.json 2>> $OUT/err.log; python -c "
import json; d=json.load(open('$OUT/bench.json')); print('bench', d['value'], d['e2e']['value'], d['phases_ms_per_step'], d['circuit']['wall_ms'])"
MPS_B200_WIDE_TASKS=0 MPS_B200_CTAS_PER_SM=4 timeout 300 python bench.py --no-cpu-baseline --no-e2e > $OUT/bench_old.json 2>> $OUT/err.log; python -c "
import json; d=json.load(open('$OUT/bench_old.json')); print('bench old rule', d['value'], d['phases_ms_per_step'], d['circuit']['wall_ms'])"
tail -3 $OUT/err.log
