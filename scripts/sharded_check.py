"""Site-sharded MPS over NCCL on the GPUs of one box, checked against the single-GPU engine and timed.
   torchrun --nproc-per-node N scripts/sharded_check.py [--qubits 32 --depth 16 --chi 128]
Rank 0 prints one JSON line: parity of <Z_k>/norm with the unsharded run of the same circuit, circuit wall times."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tnqvm_b200                      # noqa: E402
from tnqvm_b200 import circuits as Cc  # noqa: E402
from tnqvm_b200 import sharded         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=32)
    ap.add_argument("--depth", type=int, default=16)
    ap.add_argument("--chi", type=int, default=128)
    ap.add_argument("--seed", type=int, default=9)
    ap.add_argument("--partition-by", default="count", choices=["count", "cost"])
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    circ = Cc.nearest_neighbor(Cc.brickwork(a.qubits, a.depth, seed=a.seed, prefix_ghz=True))
    n2 = sum(1 for g in circ if len(g[1]) == 2)

    out = {}
    for rep in range(2):   # first pass warms NCCL connections and allocations
        sm = sharded.ShardedMPS(a.qubits, max_bond=a.chi, device=local, partition_by=a.partition_by)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        sm.run(circ)
        sm.flush()
        sm.loc.eng.sync()
        dist.barrier(); torch.cuda.synchronize()
        t_sh = time.perf_counter() - t0
        full = sm.gather_to_root()
        if rank == 0:
            z_sh, nrm_sh = full.eng.expval_z_all(), full.eng.norm()
            full.close()
        ex, by = sm.exchanges, sm.bytes_sent
        sm.close()
    tot = torch.tensor([ex, by], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        for rep in range(2):
            e = tnqvm_b200.B200MPS(a.qubits, max_bond=a.chi, device=local)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e.run(circ); e.sync()
            t_1 = time.perf_counter() - t0
            z_1, nrm_1 = e.expval_z_all(), e.norm()
            e.close()
        out = {"check": "site_sharded_vs_single_gpu", "n_gpus": world, "partition_by": a.partition_by, "qubits": a.qubits, "depth": a.depth, "max_bond_dim": a.chi,
               "gates_2q": n2, "max_abs_dz": float(np.abs(z_sh - z_1).max()), "abs_dnorm": float(abs(nrm_sh - nrm_1)),
               "wall_ms_sharded": t_sh * 1e3, "wall_ms_single_gpu": t_1 * 1e3, "speedup": t_1 / t_sh,
               "boundary_exchanges": int(tot[0].item()), "bytes_over_nvlink": int(tot[1].item())}
        print(json.dumps(out))
        assert out["max_abs_dz"] < 5e-6 and out["abs_dnorm"] < 5e-6, out
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
