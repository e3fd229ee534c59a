#!/bin/bash
# ncu evidence for round 2: launch list of one bench step + full-set captures of the dominant kernels (digested on the box)
TAG=${1:-prof}; OUT=gpurun_out/$TAG; mkdir -p $OUT
BARGS="--prep random --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-peak --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv python bench.py $BARGS > $OUT/ncu_bench.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; cat $OUT/launches_summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:jacobi_sweep -s 12 -c 2 -o $OUT/prof_jacobi_sweep python bench.py $BARGS > $OUT/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"qr_panel|qr_update" -s 70 -c 4 -o $OUT/prof_qr python bench.py $BARGS >> $OUT/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:zgemm_dmma -s 2 -c 3 -o $OUT/prof_gemm python bench.py $BARGS >> $OUT/ncu_full.log 2>&1
for r in jacobi_sweep qr gemm; do python scripts/ncu_digest.py $OUT/prof_$r.ncu-rep > $OUT/ncu_$r.txt 2>&1; done
head -40 $OUT/ncu_jacobi_sweep.txt
rm -f $OUT/launches.csv.gz; gzip -9 $OUT/launches.csv
ls -la $OUT
