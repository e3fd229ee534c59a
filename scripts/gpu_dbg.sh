#!/bin/bash
OUT=gpurun_out/r01i; mkdir -p $OUT
MPS_B200_DBG_MODE=10 timeout 90 python bench.py --prep random --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peak > $OUT/bench.json 2> $OUT/err.txt
grep "phase timing" $OUT/err.txt; tail -2 $OUT/err.txt
