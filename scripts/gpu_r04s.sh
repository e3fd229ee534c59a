#!/bin/bash
OUT=gpurun_out/r04s; mkdir -p $OUT
timeout 300 python scripts/sharded_blocks_check.py > $OUT/blocks.jsonl 2> $OUT/blocks.err; python - <<PY
import json
w = {}
for l in open("$OUT/blocks.jsonl"):
    d = json.loads(l)
    for k in ("dnorm_first", "damp", "dz", "dzz", "dnorm_after", "bond_mismatch"):
        w[k] = max(w.get(k, 0), d[k])
print("blocks check on real devices, worst deviations:", w)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench8 rc=$?"; tail -2 $OUT/bench_n8.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value", d["value"], "e2e", d["e2e"]["value"]); print(json.dumps(d["circuit_sharded"])[:1800])
PY
timeout 600 python scripts/configs_fullsize.py --gpus 8 --which c3 --fuse3 0 --out $OUT/configs_8gpu.jsonl > $OUT/configs.log 2>&1; cut -c1-700 $OUT/configs_8gpu.jsonl
