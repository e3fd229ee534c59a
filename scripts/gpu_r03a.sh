#!/bin/bash
# r03a: fuse_2q parity test + full-shape configs 3 and 5 on one B200
TAG=${1:-r03a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q -k "fuse_2q or compiled or config3 or config5" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_sel.log
timeout 200 python scripts/configs_fullsize.py --which c3 --out $OUT/configs.jsonl 2> $OUT/c3.err | cut -c1-900; tail -3 $OUT/c3.err
timeout 330 python scripts/configs_fullsize.py --which c5 --chi5 256,512 --budget 90 --out $OUT/configs.jsonl 2> $OUT/c5.err | cut -c1-900; tail -3 $OUT/c5.err
