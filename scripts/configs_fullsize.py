"""BASELINE configs 3 and 5 at their full shape on ONE B200 (the sharded runs are scripts/sharded_check.py): wall time,
outputs, and the size-independent checks that exist at this size.  One JSON line per run, appended to --out.

  C3  100-qubit ring MaxCut QAOA p=4, max-bond-dim 512, <Z_i Z_j> on the 100 ring edges and the cut energy; run gate for
      gate (reference-faithful) and with option fuse_2q (CX.Rz.CX -> one ZZ gate).  Routing the wrap edge drags a qubit
      through the whole chain, the bonds reach 512 and truncation is active, so the two runs differ at truncation level;
      --oracle also runs the CPU restatement (more than 20 minutes on 8 cores at this size).
  C5  the reference's examples/sycamore/resources/sycamore_53_14_0.xasm (committed fixture, 53 qubits, 14 cycles; 1897 NN 2q gates
      after the nearest-neighbour pass), amplitude of |0...0>, norm, fidelity estimate prod(1 - w) at the bond dimensions given
      by --chi5; the circuit is fed in chunks so that a run exceeding --budget seconds stops and says how far it got.
  --gpus N shards the sites over N GPUs inside the library (mps_create_sharded).

usage (GPU box): python scripts/configs_fullsize.py --which c3,c5 --chi5 256,512,1024 --budget 240 --out gpurun_out/rNN/configs.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tnqvm_b200
from tnqvm_b200 import circuits as Cc


def emit(out, rec):
    line = json.dumps(rec)
    print(line, flush=True)
    if out:
        os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
        with open(out, "a") as f:
            f.write(line + "\n")


def run_c3(args):
    n, p, chi = args.n3, 4, args.chi3
    circ = Cc.nearest_neighbor(Cc.qaoa_ring(n, p, seed=7))
    n1, n2 = Cc.count_gates(circ)
    edges = [(i, (i + 1) % n) for i in range(n)]
    cc = tnqvm_b200.CompiledCircuit(circ)
    zz_by_mode = {}
    modes = [int(x) for x in args.fuse3.split(",")]
    for fuse in modes:
        e = tnqvm_b200.B200MPS(n, max_bond=chi, fuse_2q=fuse, devices=list(range(args.gpus)))
        e.run(cc); e.sync(); e.reset()          # warm-up: workspace, site buffers, pinned staging at their final sizes
        s0 = e.stats()
        t0 = time.perf_counter()
        e.run(cc); e.sync()
        t_run = time.perf_counter() - t0
        t0 = time.perf_counter()
        zz = e.expval_zz_pairs(edges)
        t_obs = time.perf_counter() - t0
        s1 = e.stats()
        zz_by_mode[fuse] = zz
        emit(args.out, {"config": "c3_qaoa_ring", "gpus": args.gpus, "site_blocks": e.shard_layout(), "boundary_exchanges": s1["boundary_exchanges"] - s0["boundary_exchanges"], "qubits": n, "p": p, "max_bond_dim": chi, "fuse_2q": fuse, "gates_1q": n1, "gates_2q_nn": n2,
                        "gates_2q_executed": s1["gates_2q"] - s0["gates_2q"], "gates_2q_fused": s1["gates_2q_fused"] - s0["gates_2q_fused"],
                        "run_ms": 1e3 * t_run, "nn_gates_2q_per_s": n2 / t_run, "zz_100_edges_ms": 1e3 * t_obs,
                        "energy": float(((1 - zz) / 2).sum()), "norm": e.norm(), "max_bond_reached": int(e.bond_dims().max()),
                        "discarded_weight": e.discarded_weight(), "jacobi_sweeps": s1["jacobi_sweeps"] - s0["jacobi_sweeps"],
                        "launches": s1["launches"] - s0["launches"]})
        e.close()
    if len(modes) == 2:
        emit(args.out, {"config": "c3_qaoa_ring", "check": "fused_vs_gate_by_gate", "max_abs_dzz": float(np.abs(zz_by_mode[0] - zz_by_mode[1]).max())})
    if args.oracle:
        from oracle import oracle as O   # checker only
        t0 = time.perf_counter()
        o = O.OracleMPS(n, max_bond=chi).run(circ)
        t_or = time.perf_counter() - t0
        ref = np.array([o.expval_z([i, j]) for i, j in edges])
        emit(args.out, {"config": "c3_qaoa_ring", "check": "oracle_full_size", "oracle_run_s": t_or, "cores": os.cpu_count(),
                        "max_abs_dzz": {str(k): float(np.abs(v - ref).max()) for k, v in zz_by_mode.items()},
                        "energy_oracle": float(((1 - ref) / 2).sum()), "abs_dnorm": abs(o.norm() - 1.0)})


def run_c5(args):
    n, raw = Cc.sycamore_53(args.depth5)
    circ = Cc.nearest_neighbor(raw)
    n1, n2 = Cc.count_gates(circ)
    chunk = args.chunk   # instructions per ABI call; large, so that the dependency layers keep their natural width
    for chi in [int(x) for x in args.chi5.split(",") if x]:
        for fuse in ([0, 1] if chi <= args.fuse_both_upto else [args.fuse5]):
            e = tnqvm_b200.B200MPS(n, max_bond=chi, fuse_2q=fuse, devices=list(range(args.gpus)))
            t0 = time.perf_counter()
            done2 = 0
            complete = True
            for i in range(0, len(circ), chunk):
                part = circ[i:i + chunk]
                e.run(tnqvm_b200.CompiledCircuit(part)); e.sync()
                done2 += Cc.count_gates(part)[1]
                if time.perf_counter() - t0 > args.budget:
                    complete = i + chunk >= len(circ)
                    break
            t_run = time.perf_counter() - t0
            s = e.stats()
            rec = {"config": "c5_sycamore_53_14_0", "gpus": args.gpus, "site_blocks": e.shard_layout(), "boundary_exchanges": s["boundary_exchanges"],
                   "svd_nonconverged": s["svd_nonconverged"], "norm_guard_violations": s["norm_guard_violations"], "qubits": n, "depth": args.depth5, "max_bond_dim": chi, "fuse_2q": fuse, "gates_1q": n1, "gates_2q_nn": n2,
                   "complete": complete, "gates_2q_nn_done": done2, "gates_2q_executed": s["gates_2q"], "gates_2q_fused": s["gates_2q_fused"],
                   "run_s": t_run, "nn_gates_2q_per_s": done2 / t_run, "max_bond_reached": int(e.bond_dims().max()),
                   "bonds_at_max": int((e.bond_dims() >= chi).sum()), "jacobi_sweeps": s["jacobi_sweeps"], "launches": s["launches"]}
            if complete:
                t0 = time.perf_counter()
                amp = e.amplitude([0] * n)
                rec.update({"amp0_re": amp.real, "amp0_im": amp.imag, "amp0_abs2_times_2^53": abs(amp) ** 2 * 2.0 ** n, "amp_ms": 1e3 * (time.perf_counter() - t0)})
                nrm = e.norm()
                dw = e.discarded_weight()
                # fidelity estimate of the truncated run: prod over truncations of (1 - discarded / total weight); the norm is a
                # different quantity (the reference gauge is never renormalised nor canonical)
                rec.update({"norm": nrm, "discarded_weight_sum": dw, "fidelity_estimate": e.fidelity_estimate()})
            emit(args.out, rec)
            e.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c3,c5")
    ap.add_argument("--n3", type=int, default=100)
    ap.add_argument("--chi3", type=int, default=512)
    ap.add_argument("--fuse3", default="0,1")
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--chi5", default="256,512")
    ap.add_argument("--depth5", type=int, default=14)
    ap.add_argument("--fuse5", type=int, default=0)
    ap.add_argument("--fuse-both-upto", type=int, default=256)
    ap.add_argument("--budget", type=float, default=240.0)
    ap.add_argument("--chunk", type=int, default=512)
    ap.add_argument("--out", default="")
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    if "c3" in a.which:
        run_c3(a)
    if "c5" in a.which:
        run_c5(a)
