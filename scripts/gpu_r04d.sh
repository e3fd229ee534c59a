#!/bin/bash
OUT=gpurun_out/r04d; mkdir -p $OUT
nvidia-smi -L > $OUT/gpu.txt
timeout 300 python scripts/sharded_abi_check.py --gpus 2 --qubits 16 --depth 8 --chi 16 > $OUT/shard_small.json 2> $OUT/shard_small.err; echo "small rc=$?"; cat $OUT/shard_small.json; tail -3 $OUT/shard_small.err
timeout 300 python scripts/sharded_abi_check.py --gpus 2 --qubits 50 --depth 20 --chi 256 > $OUT/shard_c2.json 2> $OUT/shard_c2.err; echo "c2 rc=$?"; cat $OUT/shard_c2.json; tail -3 $OUT/shard_c2.err
timeout 300 python scripts/sharded_abi_check.py --gpus 2 --qubits 50 --depth 20 --chi 256 --partition-by count > $OUT/shard_c2_count.json 2>> $OUT/shard_c2.err; cat $OUT/shard_c2_count.json
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_config2_full_shape_parity --deselect tests/test_gpu_parity.py::test_config3_full_qubit_count_parity -k "not config5_real" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
timeout 600 python scripts/parity_diag.py c2s c2 > $OUT/diag.jsonl 2> $OUT/diag.err; tail -3 $OUT/diag.err; cat $OUT/diag.jsonl | cut -c1-1500
