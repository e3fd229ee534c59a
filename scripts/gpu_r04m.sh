#!/bin/bash
OUT=gpurun_out/r04m; mkdir -p $OUT
nvidia-smi -L > $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -k "sharded or visitor_site or sampling_large_register_many" -s > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "GHZ-35|passed|failed|FAILED|Error" $OUT/pytest.log | tail
timeout 900 python scripts/configs_fullsize.py --which c5 --chi5 512 --fuse-both-upto 0 --budget 300 --out $OUT/configs_1gpu.jsonl > $OUT/configs.log 2>&1; cut -c1-700 $OUT/configs_1gpu.jsonl
timeout 900 python scripts/configs_fullsize.py --gpus 2 --which c3,c5 --fuse3 0 --chi5 256,512 --fuse-both-upto 0 --budget 300 --out $OUT/configs_2gpu.jsonl >> $OUT/configs.log 2>&1; cut -c1-700 $OUT/configs_2gpu.jsonl; tail -3 $OUT/configs.log
