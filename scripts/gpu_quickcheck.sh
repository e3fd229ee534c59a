OUT=gpurun_out/r04w; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "svd_engine_variants or exact_parity or config1 or full_size or config2_full_shape_parity or golden or sharded" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-peak --no-extras > $OUT/bench.json 2>> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.1f ms/step %.2f circuit %.1f" % (d["value"], d["ms_per_step"], d["circuit"]["wall_ms"]), d["phases_ms_per_step"])
PY
