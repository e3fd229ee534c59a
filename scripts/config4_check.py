"""Config 4 (64-qubit hardware-efficient ansatz, 4 layers, max-bond-dim 64, batched parameter sets) on one GPU:
R parameter sets as R registers of one handle (their gates share launches).  Prints one JSON line."""
import json, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tnqvm_b200
from tnqvm_b200 import circuits as Cc

n, L, chi = 64, 4, 64
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
circs = [Cc.hea(n, L, seed=s) for s in range(R)]
comp = [tnqvm_b200.CompiledCircuit(c, offset=r * n) for r, c in enumerate(circs)]
for rep in range(2):
    e = tnqvm_b200.B200MPS(n, max_bond=chi, n_registers=R)
    e.sync()
    t0 = time.perf_counter()
    for c in comp:
        e.run(c)
    e.sync()
    t_run = time.perf_counter() - t0
    t0 = time.perf_counter()
    z = np.array([e.expval_z_all(reg=r) for r in range(R)])
    t_obs = time.perf_counter() - t0
    st = e.stats()
    e.close()
n2 = sum(1 for g in circs[0] if len(g[1]) == 2) * R
print(json.dumps({"check": "config4_hea", "registers": R, "qubits": n, "layers": L, "max_bond_dim": chi, "gates_2q": n2,
                  "run_ms": t_run * 1e3, "gates_2q_per_s": n2 / t_run, "observables_ms": t_obs * 1e3, "engine_layers": st["layers"],
                  "launches": st["launches"], "jacobi_sweeps": st["jacobi_sweeps"], "z_checksum": float(np.abs(z).sum())}))
