"""Lone saturated 2q gates (the routed-circuit regime of configs 3 and 5: one to a few gates per dependency layer): ms per gate with
the streaming and with the cluster-resident Jacobi sweep.  usage: python scripts/lone_gate_bench.py [chi ...]"""
import json, math, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tnqvm_b200
for chi in [int(x) for x in sys.argv[1:]] or [256, 512, 1024]:
    for ngate in (1, 3):
        n = 2 + 2 * ngate
        rng = np.random.default_rng(1)
        dims = [1] + [chi] * (n - 1) + [1]
        sites = [(rng.standard_normal((dims[k], 2, dims[k + 1])) + 1j * rng.standard_normal((dims[k], 2, dims[k + 1]))) / math.sqrt(2 * dims[k] * dims[k + 1]) for k in range(n)]
        m = tnqvm_b200.gates.gate_matrix("fSim", (0.4, 1.1))
        for cluster in (0, 1):
            e = tnqvm_b200.B200MPS(n, max_bond=chi, jacobi_cluster=cluster)
            for k in range(n):
                e.set_site(k, sites[k])
            def layer():
                for g in range(ngate):
                    e.apply_2q(1 + 2 * g, 2 + 2 * g, m)
                e.flush()
            layer(); e.sync()
            for k in range(n):
                e.set_site(k, sites[k])
            e.set_option("profile", 1)
            s0 = e.stats(); t0 = time.perf_counter()
            reps = 3
            for i in range(reps):
                layer()
            e.sync(); dt = (time.perf_counter() - t0) / reps
            s1 = e.stats()
            print(json.dumps({"chi": chi, "gates_per_layer": ngate, "jacobi_cluster": cluster, "layer_ms": dt * 1e3, "ms_svd": (s1["ms_svd"] - s0["ms_svd"]) / reps,
                              "ms_qr_within_svd": (s1["ms_qr"] - s0["ms_qr"]) / reps, "sweeps": (s1["jacobi_sweeps"] - s0["jacobi_sweeps"]) / reps,
                              "norm": e.norm()}), flush=True)
            e.close()
