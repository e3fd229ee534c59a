"""BASELINE config 4 as specified: 64-qubit hardware-efficient VQE ansatz (4 layers), 64 parameter sets sharded across the GPUs
of the box (torchrun, one rank per GPU, tnqvm_b200.sharded.run_parameter_sweep: no data-path collective, the <Z_k> rows are
gathered at the end), every set checked against the oracle on rank 0.  Prints one JSON line on rank 0.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/config4_sweep.py"""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnqvm_b200 import circuits as Cc
from tnqvm_b200.sharded import run_parameter_sweep

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, L, chi, R = 64, 4, 64, 64
circs = [Cc.hea(n, L, seed=s) for s in range(R)]
run_parameter_sweep(n, circs[:world], max_bond=chi, device=local)   # warm-up (lazy module loading, allocators)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
z = run_parameter_sweep(n, circs, max_bond=chi, device=local)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    from oracle import oracle as O   # checker only
    worst = 0.0
    for s, c in enumerate(circs):
        o = O.OracleMPS(n, max_bond=chi).run(c)
        worst = max(worst, float(np.abs(z[s] - np.array([o.expval_z([k]) for k in range(n)])).max()))
    n2 = sum(1 for g in circs[0] if len(g[1]) == 2) * R
    print(json.dumps({"config": "c4_hea64_64_parameter_sets", "gpus": world, "sets_per_gpu": R // world, "qubits": n, "layers": L, "max_bond_dim": chi,
                      "gates_2q": n2, "wall_ms_including_observables_and_gather": dt * 1e3, "gates_2q_per_s": n2 / dt,
                      "max_abs_dz_vs_oracle_over_all_sets": worst}))
dist.destroy_process_group()
