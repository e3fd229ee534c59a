OUT=gpurun_out/r04z; mkdir -p $OUT
BARGS="--qubits 24 --chi 512 --prep random --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-peak --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_chi512.csv python bench.py $BARGS > $OUT/ncu_bench.log 2>&1
python scripts/summarize_launches.py $OUT/launches_chi512.csv > $OUT/launches_chi512_summary.txt 2>&1; cat $OUT/launches_chi512_summary.txt
# a lone chi=512 gate per layer (the routed-circuit regime): 4-qubit chain, gates (1,2) only
python - <<PY > $OUT/lone_gate.json
import json, time, math, numpy as np, sys
sys.path.insert(0, ".")
import tnqvm_b200
chi, n = 512, 4
rng = np.random.default_rng(1)
e = tnqvm_b200.B200MPS(n, max_bond=chi)
dims = [1, chi, chi, chi, 1]
for k in range(n):
    t = (rng.standard_normal((dims[k], 2, dims[k + 1])) + 1j * rng.standard_normal((dims[k], 2, dims[k + 1]))) / math.sqrt(2 * dims[k] * dims[k + 1])
    e.set_site(k, t)
m = tnqvm_b200.gates.gate_matrix("fSim", (0.4, 1.1))
e.apply_2q(1, 2, m); e.sync()
e.set_option("profile", 1)
s0 = e.stats(); t0 = time.perf_counter()
for i in range(5):
    e.apply_2q(1, 2, m); e.flush()
e.sync(); dt = (time.perf_counter() - t0) / 5
s1 = e.stats()
print(json.dumps({"lone_gate_chi512_ms": dt * 1e3, "ms_theta": (s1["ms_theta"] - s0["ms_theta"]) / 5, "ms_svd": (s1["ms_svd"] - s0["ms_svd"]) / 5, "ms_qr_within_svd": (s1["ms_qr"] - s0["ms_qr"]) / 5,
                  "ms_writeback": (s1["ms_writeback"] - s0["ms_writeback"]) / 5, "sweeps": (s1["jacobi_sweeps"] - s0["jacobi_sweeps"]) / 5, "launches": (s1["launches"] - s0["launches"]) / 5}))
PY
cat $OUT/lone_gate.json
rm -f $OUT/launches_chi512.csv.gz; gzip -9 $OUT/launches_chi512.csv
