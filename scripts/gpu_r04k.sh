#!/bin/bash
OUT=gpurun_out/r04k; mkdir -p $OUT
timeout 600 python -m pytest tests/test_visitor.py tests/test_gpu_parity.py -m gpu -q -k "visitor or teacher or sampling" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "GHZ-35|passed|failed|FAILED" $OUT/pytest.log | tail
timeout 900 python scripts/configs_fullsize.py --which c3,c5 --fuse3 0 --chi5 256,512 --fuse-both-upto 0 --budget 300 --out $OUT/configs.jsonl > $OUT/configs.log 2>&1; echo "configs rc=$?"; cut -c1-900 $OUT/configs.jsonl; tail -3 $OUT/configs.log
