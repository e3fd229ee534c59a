#!/bin/bash
OUT=gpurun_out/r04i; mkdir -p $OUT
for G in 2 4; do
  timeout 200 python scripts/sharded_step_bench.py --gpus $G > $OUT/step_n$G.json 2> $OUT/step_n$G.err; cat $OUT/step_n$G.json; tail -2 $OUT/step_n$G.err
  timeout 200 python scripts/sharded_step_bench.py --gpus $G --chi 512 --qubits 32 --steps 2 > $OUT/step512_n$G.json 2> $OUT/step512_n$G.err; cat $OUT/step512_n$G.json; tail -2 $OUT/step512_n$G.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_n4.json 2> $OUT/bench_n4.err; echo "bench4 rc=$?"; tail -3 $OUT/bench_n4.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n4.json").read().strip().splitlines()[-1])
print("N=4 value", d["value"], "e2e", d["e2e"]["value"]); print(json.dumps(d["circuit_sharded"], indent=1))
PY
