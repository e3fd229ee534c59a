#!/bin/bash
# ncu evidence only: launch list of one bench step + full-set captures of the Jacobi and GEMM kernels.
# usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_prof.sh rNN'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
BARGS="--prep random --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-peak"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches.csv \
  python bench.py $BARGS > $OUT/ncu_bench.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; cat $OUT/launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi -s 300 -c 3 -o $OUT/prof_jacobi \
  python bench.py $BARGS > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zgemm_dmma -s 2 -c 3 -o $OUT/prof_gemm \
  python bench.py $BARGS >> $OUT/ncu_full.log 2>&1
rm -f $OUT/launches.csv.gz; gzip -9 $OUT/launches.csv
ls -la $OUT
