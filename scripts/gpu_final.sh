#!/bin/bash
# End-of-round validation on one B200: full gpu test suite, bench (+ reference arm, chi=512 line), ncu launch list of one
# bench step, full-shape configs 3 and 5.  usage (here): gpurun --timeout 900 -- 'bash scripts/gpu_final.sh rNN'
TAG=${1:-r03z}
OUT=gpurun_out/$TAG
bash scripts/gpu_round.sh $TAG 2>&1 | tail -12
BARGS="--prep random --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-peak"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
  python bench.py $BARGS > $OUT/ncu_bench.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; cat $OUT/launches_summary.txt
rm -f $OUT/launches.csv.gz; gzip -9 $OUT/launches.csv
timeout 150 python scripts/configs_fullsize.py --which c3 --out $OUT/configs.jsonl 2> $OUT/c3.err | cut -c1-500; tail -2 $OUT/c3.err
timeout 60 python scripts/configs_fullsize.py --which c5 --chi5 256 --fuse-both-upto 0 --budget 40 --out $OUT/configs.jsonl 2> $OUT/c5.err | cut -c1-500; tail -2 $OUT/c5.err
