#!/bin/bash
# what the driver does at round end, on a 2-GPU box: smoke(), bench at N=1 is covered by gpu_round.sh; here N=2 under torchrun (both arms)
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench2 rc=$?"; tail -2 $OUT/bench_n2.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"], "launches", d["gpu_launches"]); print(json.dumps(d["circuit_sharded"])[:1500])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; echo "ref2 rc=$?"; cut -c1-200 $OUT/bench_ref_n2.json
