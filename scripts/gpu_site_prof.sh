#!/bin/bash
# HBM-side kernels: CUDA-event numbers of the 1q layer and the transfer sweeps + ncu full-set digests of gate1q_kernel and of
# the mid-chain (saturated) transfer-sweep GEMMs.  usage (here): gpurun --timeout 600 -- "bash scripts/gpu_site_prof.sh rNN"
TAG=${1:-r03h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 python scripts/site_kernels_bench.py --qubits 100 --chi 256 > $OUT/site_kernels.json 2> $OUT/site.err; cat $OUT/site_kernels.json; tail -3 $OUT/site.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gate1q -s 1 -c 2 -o $OUT/prof_1q \
  python scripts/site_kernels_bench.py --qubits 100 --chi 256 --reps 2 > $OUT/ncu_site.log 2>&1; tail -1 $OUT/ncu_site.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"site_dot|zgemm_small|trace_pair|zgemm_dmma_kernel1" -s 160 -c 8 -o $OUT/prof_sweep \
  python scripts/site_kernels_bench.py --qubits 100 --chi 256 --reps 2 >> $OUT/ncu_site.log 2>&1; tail -1 $OUT/ncu_site.log
python scripts/ncu_digest.py $OUT/prof_1q.ncu-rep > $OUT/ncu_gate1q.txt 2>&1
python scripts/ncu_digest.py $OUT/prof_sweep.ncu-rep > $OUT/ncu_transfer_sweep.txt 2>&1
grep -E "kernel:|gpu__time_duration|dram__bytes|dram__throughput|tensor_cycles" $OUT/ncu_gate1q.txt $OUT/ncu_transfer_sweep.txt | cut -c1-160
rm -f $OUT/*.ncu-rep
