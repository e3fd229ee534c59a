"""Sharded handle with many blocks (all on GPU 0, or spread over the GPUs present) vs one engine: gates, observables, wrap-edge ZZ."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tnqvm_b200
from tnqvm_b200 import circuits as Cc
ng = torch.cuda.device_count()
for name, n, circ, chi in (("qaoa", 40, Cc.nearest_neighbor(Cc.qaoa_ring(40, 2, seed=7)), 16), ("brick", 40, Cc.brickwork(40, 12, seed=3), 16),
                           ("qaoa_exact", 16, Cc.nearest_neighbor(Cc.qaoa_ring(16, 2, seed=5)), 0)):
    a = tnqvm_b200.B200MPS(n, max_bond=chi); a.run(circ)
    za, na = a.expval_z_all(), a.norm()
    edges = [(i, (i + 1) % n) for i in range(n)]
    zza = a.expval_zz_pairs(edges)
    amp_a = a.amplitude([k % 2 for k in range(n)])
    for P in (2, 3, 5, 8):
        for part in ("cost", "count"):
            devs = [d % ng for d in range(P)]
            b = tnqvm_b200.B200MPS(n, max_bond=chi, devices=devs, partition_by=part); b.run(circ)
            nb1 = b.norm()                      # before any other observable: the left chain alone
            amp_b = b.amplitude([k % 2 for k in range(n)])   # gather path
            zb = b.expval_z_all(); zzb = b.expval_zz_pairs(edges); nb2 = b.norm()
            print(json.dumps({"circ": name, "blocks": P, "part": part, "layout": b.shard_layout(), "dnorm_first": abs(nb1 - na) / abs(na), "damp": abs(amp_a - amp_b) / max(1e-300, abs(amp_a)),
                              "dz": float(np.abs(za - zb).max()), "dzz": float(np.abs(zza - zzb).max()), "dnorm_after": abs(nb2 - na) / abs(na),
                              "bond_mismatch": int((np.asarray(a.bond_dims()) != np.asarray(b.bond_dims())).sum())}), flush=True)
            b.close()
    a.close()
