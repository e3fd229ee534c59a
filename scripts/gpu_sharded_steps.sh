OUT=gpurun_out/r05d; mkdir -p $OUT
for G in 4 8; do
  timeout 200 python scripts/sharded_step_bench.py --gpus $G > $OUT/step_n$G.json 2> $OUT/step.err; cut -c1-400 $OUT/step_n$G.json
  timeout 200 python scripts/sharded_step_bench.py --gpus $G --chi 512 --qubits 50 --steps 2 > $OUT/step512_n$G.json 2>> $OUT/step.err; cut -c1-400 $OUT/step512_n$G.json
done
