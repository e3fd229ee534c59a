#!/bin/bash
# 8-GPU scaling evidence (run with gpurun --gpus 8)
TAG=${1:-scale8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpu.txt
timeout 300 python -m pytest tests -m gpu -q -k "sharded or nccl" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/config4_sweep.py > $OUT/config4.json 2> $OUT/config4.err; grep config $OUT/config4.json | cut -c1-600; tail -2 $OUT/config4.err
for G in 2 4 8; do
  timeout 200 python scripts/sharded_step_bench.py --gpus $G > $OUT/step_n$G.json 2> $OUT/step.err; cut -c1-500 $OUT/step_n$G.json
  timeout 200 python scripts/sharded_step_bench.py --gpus $G --chi 512 --qubits 50 --steps 2 > $OUT/step512_n$G.json 2>> $OUT/step.err; cut -c1-500 $OUT/step512_n$G.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench8 rc=$?"; tail -2 $OUT/bench_n8.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value", d["value"], "e2e", d["e2e"]["value"]); print(json.dumps(d["circuit_sharded"])[:1800])
PY
timeout 900 python scripts/configs_fullsize.py --gpus 8 --which c3,c5 --fuse3 0 --chi5 512,1024 --fuse-both-upto 0 --budget 400 --out $OUT/configs_8gpu.jsonl > $OUT/configs.log 2>&1; cut -c1-900 $OUT/configs_8gpu.jsonl; tail -2 $OUT/configs.log
