#!/bin/bash
OUT=gpurun_out/r04e; mkdir -p $OUT
nvidia-smi -L > $OUT/gpu.txt
timeout 300 python scripts/sharded_abi_check.py --gpus 2 --qubits 16 --depth 8 --chi 16 > $OUT/shard_small.json 2> $OUT/shard_small.err; echo "small rc=$?"; cat $OUT/shard_small.json; tail -3 $OUT/shard_small.err
timeout 300 python scripts/sharded_abi_check.py --gpus 2 --qubits 50 --depth 20 --chi 256 > $OUT/shard_c2.json 2> $OUT/shard_c2.err; echo "c2 rc=$?"; cat $OUT/shard_c2.json; tail -3 $OUT/shard_c2.err
timeout 300 python scripts/sharded_abi_check.py --gpus 2 --qubits 50 --depth 20 --chi 256 --partition-by count > $OUT/shard_c2_count.json 2>> $OUT/shard_c2.err; cat $OUT/shard_c2_count.json
timeout 1200 python -m pytest tests -m gpu -q --durations=12 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/pytest_gpu.log
