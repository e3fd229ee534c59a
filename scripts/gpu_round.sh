#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the reference arm, the ncu launch list and one full capture.
# usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh rNN'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 300 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
if [ "$2" != "noprof" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv \
  python bench.py --prep random --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-peak > $OUT/ncu_bench.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; cat $OUT/launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_step -s 300 -c 3 -o $OUT/prof_jacobi \
  python bench.py --prep random --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-peak > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zgemm_dmma -s 2 -c 3 -o $OUT/prof_gemm \
  python bench.py --prep random --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-peak >> $OUT/ncu_full.log 2>&1
rm -f $OUT/launches.csv.gz; gzip -9 $OUT/launches.csv
fi
ls -la $OUT
