"""Saturated-state step (bench.py's unit of work) on a site-sharded handle vs one engine: wall ms per step.
usage: python scripts/sharded_step_bench.py --gpus 2 [--qubits 50 --chi 256 --steps 4]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tnqvm_b200
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--qubits", type=int, default=50)
ap.add_argument("--chi", type=int, default=256)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--partition-by", default="cost")
a = ap.parse_args()
n, chi = a.qubits, a.chi
sites = B.random_mps_sites(n, chi, 7)
steps = [tnqvm_b200.CompiledCircuit(B.step_circuit(n, i, 7)) for i in range(a.steps + 2)]
out = {"qubits": n, "chi": chi, "steps": a.steps}
for tag, devs in (("1gpu", [0]), ("sharded", list(range(a.gpus)))):
    e = tnqvm_b200.B200MPS(n, max_bond=chi, devices=devs, partition_by=a.partition_by)
    for k, t in enumerate(sites):
        e.set_site(k, t)
    for i in range(2):
        e.run(steps[i]); e.sync()
    t0 = time.perf_counter()
    for i in range(2, 2 + a.steps):
        e.run(steps[i]); e.flush()
    e.sync()
    dt = (time.perf_counter() - t0) / a.steps
    st = e.stats()
    out[tag] = {"ms_per_step": dt * 1e3, "gates_per_s": (n - 1) / dt, "layout": e.shard_layout(), "norm": e.norm(), "sweeps": st["jacobi_sweeps"], "layers": st["layers"]}
    e.close()
out["speedup"] = out["1gpu"]["ms_per_step"] / out["sharded"]["ms_per_step"]
print(json.dumps(out))
