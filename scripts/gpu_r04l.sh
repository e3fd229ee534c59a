#!/bin/bash
OUT=gpurun_out/r04l; mkdir -p $OUT
MPS_B200_TRACE=1 timeout 600 python scripts/configs_fullsize.py --which c5 --chi5 512 --fuse-both-upto 0 --budget 300 --out $OUT/configs.jsonl > $OUT/configs.log 2> $OUT/trace.err
grep -c "sweep" $OUT/trace.err; grep "NOT converged" $OUT/trace.err | head -10
grep "NOT converged" $OUT/trace.err | head -3 | while read l; do L=$(echo "$l" | sed 's/.*layer \([0-9]*\) chunk.*/\1/'); grep "layer $L chunk" $OUT/trace.err | tail -45 | cut -c1-120; done > $OUT/nonconv_detail.txt
head -60 $OUT/nonconv_detail.txt
gzip -9 $OUT/trace.err
