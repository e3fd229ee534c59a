#!/usr/bin/env python
"""bench.py -- MPS 2q gates/sec at chi=256 (complex128) on B200, BASELINE.json config[1].

Workload ("c2_brickwork_n50_chi256"): a 50-qubit random brickwork circuit (rcs gate set + CNOT, seeded) is run
from |0...0> with max-bond-dim 256 for `--depth` (20) layers -- that run is timed once and reported as
`circuit` (the "circuit wall time" half of the metric) and leaves every interior bond saturated at chi=256.
A STEP is then two further brickwork layers (one even, one odd: 100 random 1q gates + 49 CNOTs, fresh seeded
gates every step) on that saturated state with truncation 512 -> 256 active: the unit of work of SURVEY.md
section 8(d).  `value` = 2q gates/s over K steps, device-timed with CUDA events on the engine's own stream,
state resident in HBM.  `e2e` = the same step driven with HOST-resident state: every step uploads all site
tensors from pinned host memory (mps_set_site), applies the gates, reads <Z_k>, the norm and all site tensors
back (mps_get_site).

N > 1 (torchrun): the path shards by independent circuits (config 4 style): every rank runs its own
seeded 50-qubit circuit, no data-path collective, weak scaling; time = max over ranks.

--impl reference: the CPU restatement of the reference algorithm (oracle/, host OpenBLAS zgemm + zgesdd, all host
threads) on a bounded sample of the same step (a few saturated-bond gates per step on a random chi=256 MPS).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mps_2q_gates_per_sec_chi256_c128"
UNIT = "gates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=50)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--gauge", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the chi=512 theta line, the visitor-path run and the sharded circuit")
    ap.add_argument("--no-peak", action="store_true", help="skip the live cuBLAS DGEMM peak measurement (profiling runs)")
    ap.add_argument("--no-qr", action="store_true", help="A/B: disable the QR pre-reduction of the SVD")
    ap.add_argument("--jacobi-tol", type=float, default=0.0, help="A/B: override the Jacobi convergence tolerance")
    ap.add_argument("--prep", default="circuit", choices=["circuit", "random"],
                    help="how the saturated state is made: run the depth-D circuit (default) or load random chi-saturated sites (profiling runs)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ workload
def step_circuit(n, step_idx, seed):
    """Two brickwork layers (even pairs then odd pairs) with fresh seeded 1q gates: tnqvm_b200.circuits.brickwork
    restricted to layers (2*step_idx, 2*step_idx+1) of an endless circuit."""
    from tnqvm_b200 import circuits as Cc
    rng = np.random.Generator(np.random.PCG64([seed, step_idx]))
    c = []
    for layer in range(2):
        for q in range(n):
            g = Cc.RCS_GATES[int(rng.integers(0, len(Cc.RCS_GATES)))]
            c.append((g, (q,), (float(rng.uniform(-math.pi, math.pi)),) if g in ("Rx", "Ry", "Rz") else ()))
        for j in range(layer, n - 1, 2):
            c.append(("CNOT", (j, j + 1), ()))
    return c


def gate_flops(bonds, n, circ):
    """Algorithmic flops of the 2q gates of `circ` at the given bond dimensions:
    theta: 8*(2 chiL)(2 chiR) chi + 128 chiL chiR ; SVD (LAPACK-equivalent): 88 * M * N * min(M, N), M = 2 chiL, N = 2 chiR."""
    dims = [1] + [int(b) for b in bonds] + [1]
    th = sv = 0.0
    sat = 0
    for g in circ:
        if len(g[1]) == 2:
            lo = min(g[1])
            cl, ch, cr = dims[lo], dims[lo + 1], dims[lo + 2]
            th += 8.0 * (2 * cl) * (2 * cr) * ch + 128.0 * cl * cr
            sv += 88.0 * (2 * cl) * (2 * cr) * min(2 * cl, 2 * cr)
            sat += int(cl == ch == cr)
    return th, sv, sat


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        hi = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(torch, dev):
    """FP64 tensor (DMMA) denominator: MEASURED_PEAKS.json carries no FP64 figure, so cuBLAS DGEMM 8192^3 is timed
    live (burst: best of 5, CUDA events) -- 'of measured cuBLAS DGEMM'."""
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    b = torch.randn(n, n, device=dev, dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def bond_profile(n, chi):
    return [min(chi, 2 ** min(k + 1, n - 1 - k)) for k in range(n - 1)]


def random_mps_sites(n, chi, seed):
    """Random complex Gaussian sites with the bond profile of a saturated 50-qubit chain, scaled ~left-isometric."""
    rng = np.random.default_rng(seed)
    dims = [1] + bond_profile(n, chi) + [1]
    out = []
    for k in range(n):
        dl, dr = dims[k], dims[k + 1]
        t = (rng.standard_normal((dl, 2, dr)) + 1j * rng.standard_normal((dl, 2, dr))) / math.sqrt(2.0 * 2 * dl)
        out.append(np.asfortranarray(t))
    return out


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O   # checker / CPU baseline (the one place bench.py executes oracle/)
    cores = os.cpu_count() or 1
    O.lib().oracle_set_threads(cores)
    n, chi = args.qubits, args.chi
    o = O.OracleMPS(n, max_bond=chi, gesdd=True, gauge=args.gauge)
    for k, t in enumerate(random_mps_sites(n, chi, args.seed)):
        o.set_site(k, t)
    def one_step(i):
        # the SAME step the B200 arm times: two brickwork layers, all 2n 1q gates and all n-1 CNOTs
        n2 = 0
        for g in step_circuit(n, i, args.seed):
            o.apply(g[0], g[1], g[2])
            n2 += len(g[1]) == 2
        return n2

    for i in range(args.warmup):
        one_step(i)
    t0 = time.perf_counter()
    n2 = 0
    for i in range(args.steps):
        n2 += one_step(args.warmup + i)
    dt = time.perf_counter() - t0
    val = n2 / dt
    sample = ("the full step of the B200 arm (2 brickwork layers: %d 1q + %d 2q gates, chain ends included) on a random chi=%d-saturated "
              "%d-qubit MPS; zgemm+zgesdd (scipy OpenBLAS)" % (2 * n, n - 1, chi, n))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": "c2_brickwork_n%d_chi%d" % (n, chi), "qubits": n, "max_bond_dim": chi, "gauge": "reference" if args.gauge == 0 else "canonical",
                       "step": "2 brickwork layers = %d 1q + %d 2q gates on the saturated state" % (2 * n, n - 1), "same_step_as_b200_arm": True,
                       "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import tnqvm_b200
    from tnqvm_b200 import circuits as Cc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")   # host-side barrier: an NCCL barrier would park a spinning kernel on the idle GPUs

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, chi = args.qubits, args.chi
    seed = args.seed + 1000 * rank
    fp64_peak = measure_fp64_peak(torch, dev) if (rank == 0 and not args.no_peak) else None

    eng = tnqvm_b200.B200MPS(n, max_bond=chi, gauge=args.gauge, device=local)
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    if args.no_qr:
        eng.set_option("qr_prereduce", 0)
    if args.jacobi_tol > 0:
        eng.set_option("jacobi_tol", args.jacobi_tol)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- the depth-`depth` circuit from |0...0> (also brings every interior bond to chi)
    circ0 = Cc.brickwork(n, args.depth, seed=seed)
    n1_0, n2_0 = Cc.count_gates(circ0)
    st0 = eng.stats()
    if args.prep == "circuit":
        # first pass untimed: CUDA lazy module loading, arena / site / pinned-buffer growth; then |0...0> again and the timed pass
        eng.run(circ0); eng.flush(); eng.sync()
        eng.reset()
        st0 = eng.stats()
        cc0 = tnqvm_b200.CompiledCircuit(circ0)
        circuit_ms = timed(lambda: (eng.run(cc0), eng.flush()))
    else:
        for k, t in enumerate(random_mps_sites(n, chi, seed)):
            eng.set_site(k, t)
        circuit_ms = float("nan")
    st1 = eng.stats()
    bonds = list(eng.bond_dims())
    norm0 = eng.norm()

    steps = [step_circuit(n, i, seed) for i in range(args.warmup + 2 * args.steps + 1)]
    n2_step = sum(1 for g in steps[0] if len(g[1]) == 2)
    # gate names -> matrices once, outside the timed regions (the C++ visitor does this per visit(); from Python it costs
    # ~25 us per gate); a step is then ONE call of the C ABI's mps_apply_gates plus the flush
    compiled = [tnqvm_b200.CompiledCircuit(c) for c in steps]

    def do_step(i):
        eng.run(compiled[i])
        eng.flush()

    for i in range(args.warmup):
        do_step(i)
    eng.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sa = eng.stats()
    ms = timed(lambda: [do_step(args.warmup + i) for i in range(args.steps)])
    sb = eng.stats()
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.steps * n2_step / (ms * 1e-3)
    launches = int(sb["launches"] - sa["launches"])
    sweeps_per_layer = (sb["jacobi_sweeps"] - sa["jacobi_sweeps"]) / max(1.0, sb["layers"] - sa["layers"])

    # ---- per-phase split (separate profiled pass: events around theta / SVD / write-back of every layer)
    bonds = list(eng.bond_dims())
    eng.set_option("profile", 1)
    pa = eng.stats()
    th_fl = sv_fl = 0.0
    sat = 0
    for i in range(args.steps):
        k = args.warmup + args.steps + i
        a, b, c = gate_flops(bonds, n, steps[k])
        th_fl += a; sv_fl += b; sat += c
        do_step(k)
    eng.sync()
    pb = eng.stats()
    eng.set_option("profile", 0)
    ms_theta, ms_svd, ms_wb, ms_qr = (pb[k] - pa[k] for k in ("ms_theta", "ms_svd", "ms_writeback", "ms_qr"))
    jac_flops = pb["jacobi_dmma_flops"] - pa["jacobi_dmma_flops"]   # real DMMA flops the Jacobi pair tasks executed
    jac_tf = jac_flops / max(1e-9, (ms_svd - ms_qr) * 1e-3) / 1e12
    jac_launches = None

    # ---- e2e: host-resident state, pinned host buffers both ways, observables read back
    e2e = None
    if not args.no_e2e:
        shapes = [eng.get_site(k).shape for k in range(n)]
        hin = [torch.empty(int(np.prod(s)) * 2, dtype=torch.float64).pin_memory() for s in shapes]
        hout = [torch.empty(int(np.prod(s)) * 2, dtype=torch.float64).pin_memory() for s in shapes]
        shp = np.zeros(3, dtype=np.int32)
        for k in range(n):
            eng._ck(eng.L.mps_get_site(eng.h, k, hin[k].data_ptr(), shp.ctypes.data))
        zbuf = np.zeros(n)

        import ctypes as C
        ks = (C.c_int * n)(*range(n))
        dls = (C.c_int * n)(*[s_[0] for s_ in shapes])
        drs = (C.c_int * n)(*[s_[2] for s_ in shapes])
        pin = (C.c_void_p * n)(*[t.data_ptr() for t in hin])
        pout = (C.c_void_p * n)(*[t.data_ptr() for t in hout])

        def e2e_step(i):
            # the whole host-resident state up (one call), the gates, the observables, the whole state down (one call)
            eng._ck(eng.L.mps_set_sites(eng.h, n, ks, pin, dls, drs))
            eng.run(compiled[i])
            z = eng.expval_z_all()
            nr = eng.norm()
            eng._ck(eng.L.mps_get_sites(eng.h, n, ks, pout))
            return z, nr

        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(args.warmup + i)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        nbytes = int(sum(int(np.prod(s)) * 16 for s in shapes))
        e2e = {"value": world * args.steps * n2_step / dt, "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes + 8 * (n + 1), "ms_per_step": dt / args.steps * 1e3}

    # ---- north_star target line: theta contraction at chi >= 512 against the FP64 peak (24-qubit chain, random saturated sites)
    theta512 = None
    if rank == 0 and not args.no_extras and chi < 512:
        n5, chi5 = 24, 512
        e5 = tnqvm_b200.B200MPS(n5, max_bond=chi5, device=local)
        for k, t in enumerate(random_mps_sites(n5, chi5, seed)):
            e5.set_site(k, t)
        st5 = [tnqvm_b200.CompiledCircuit(step_circuit(n5, i, seed)) for i in range(4)]
        for i in range(2):
            e5.run(st5[i]); e5.flush()
        e5.sync()
        b5 = list(e5.bond_dims())
        e5.set_option("profile", 1)
        q0 = e5.stats()
        fl5 = 0.0
        for i in range(2, 4):
            fl5 += gate_flops(b5, n5, step_circuit(n5, i, seed))[0]
            e5.run(st5[i]); e5.flush()
        e5.sync()
        q1 = e5.stats()
        e5.close()
        ms5 = q1["ms_theta"] - q0["ms_theta"]
        tf5 = fl5 / (ms5 * 1e-3) / 1e12 if ms5 > 0 else None
        theta512 = {"bound": "tensor", "kernel": "zgemm_dmma_kernel<theta>", "workload": "24-qubit chain, chi=512, 2 steps", "achieved": tf5, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": (tf5 / fp64_peak) if (tf5 and fp64_peak) else None,
                    "ms_per_step": {"theta": ms5 / 2, "svd": (q1["ms_svd"] - q0["ms_svd"]) / 2, "writeback": (q1["ms_writeback"] - q0["ms_writeback"]) / 2}}

    # ---- the same depth-D circuit through the reference-facing C++ surface: XASM text -> TNQVM::execute restated
    # (b200_tnqvm_run: setOptions, initialize, nearest-neighbour pass, visit() per instruction, finalize with the norm)
    e2e_visitor = None
    drv = os.path.join(ROOT, "tnqvm_b200", "lib", "b200_tnqvm_run")
    if rank == 0 and not args.no_extras and os.path.exists(drv):
        xasm = Cc.to_xasm(circ0)
        try:
            r = subprocess.run([drv, "--xasm", "-", "--qubits", str(n), "--max-bond-dim", str(chi), "--device", str(local), "--repeat", "3"],
                               input=xasm, capture_output=True, text=True, timeout=300)
            doc = json.loads(r.stdout.strip().splitlines()[-1])
            wall = min(doc["execute_ms"][1:])
            e2e_visitor = {"circuit_wall_ms": wall, "gates_2q_per_s": n2_0 / (wall * 1e-3), "norm": doc.get("norm"),
                           "path": "XASM -> B200MpsVisitor::initialize/visit()/finalize via b200_tnqvm_run (host wall clock, second and third execute() on one visitor instance)",
                           "device_timed_same_circuit_ms": circuit_ms}
        except Exception as ex:   # reported, never silently dropped
            e2e_visitor = {"error": repr(ex)[:300]}

    # ---- strong scaling of ONE circuit: sites sharded over the N GPUs of the box inside the library (mps_create_sharded, SURVEY 8e /
    # configs 3 and 5).  The launch contract gives one process per GPU; a site-sharded handle is one process driving N devices,
    # so rank 0 runs it over all N GPUs while the other ranks wait at a barrier with their GPUs idle.
    sharded = None
    if world > 1 and not args.no_extras:
        barrier()
        dist.barrier(group=cpu_group)
        if rank == 0:
            sharded = {"n_devices": world, "timing": "host wall clock around run + sync, best of 2 after one untimed pass", "circuits": []}
            for label, nq_s, depth_s, chi_s in (("c2_brickwork_n%d_depth%d_chi%d" % (n, args.depth, chi), n, args.depth, chi),
                                                ("brickwork_n%d_depth%d_chi512" % (n, args.depth), n, args.depth, 512)):
                circ_s = Cc.brickwork(nq_s, depth_s, seed=seed)
                cc_s = tnqvm_b200.CompiledCircuit(circ_s)
                res = {}
                for tag, devs in (("1gpu", [local]), ("sharded", list(range(world)))):
                    es = tnqvm_b200.B200MPS(nq_s, max_bond=chi_s, devices=devs)
                    best = 1e30
                    for r in range(3):
                        es.reset(); es.sync()
                        t0 = time.perf_counter()
                        es.run(cc_s); es.sync()
                        if r:
                            best = min(best, time.perf_counter() - t0)
                    st_s = es.stats()
                    res[tag] = dict(ms=best * 1e3, z=es.expval_z_all(), norm=es.norm(), layout=es.shard_layout(), exch=st_s["boundary_exchanges"] / 3,
                                    mb=st_s["peer_bytes"] / 3e6)
                    # the metric's own unit of work on this handle: saturated-state steps (two brickwork layers each) after the circuit
                    sat_steps = [tnqvm_b200.CompiledCircuit(step_circuit(nq_s, 100 + i, seed)) for i in range(4)]
                    es.run(sat_steps[0]); es.sync()
                    t0 = time.perf_counter()
                    for i in range(1, 4):
                        es.run(sat_steps[i]); es.flush()
                    es.sync()
                    res[tag]["step_ms"] = (time.perf_counter() - t0) / 3 * 1e3
                    es.close()
                sharded["circuits"].append({
                    "name": label, "gates_2q": Cc.count_gates(circ_s)[1], "wall_ms_1gpu": res["1gpu"]["ms"], "wall_ms_sharded": res["sharded"]["ms"],
                    "speedup_vs_1gpu": res["1gpu"]["ms"] / res["sharded"]["ms"], "site_blocks": res["sharded"]["layout"],
                    "saturated_step_ms_1gpu": res["1gpu"]["step_ms"], "saturated_step_ms_sharded": res["sharded"]["step_ms"],
                    "saturated_step_gates_per_s_sharded": (nq_s - 1) / (res["sharded"]["step_ms"] * 1e-3),
                    "saturated_step_speedup_vs_1gpu": res["1gpu"]["step_ms"] / res["sharded"]["step_ms"],
                    "boundary_exchanges": res["sharded"]["exch"], "peer_mbytes": res["sharded"]["mb"],
                    "parity_max_abs_dz_vs_1gpu": float(np.abs(res["1gpu"]["z"] - res["sharded"]["z"]).max()),
                    "parity_rel_dnorm_vs_1gpu": abs(res["1gpu"]["norm"] - res["sharded"]["norm"]) / abs(res["1gpu"]["norm"])})
        dist.barrier(group=cpu_group)

    # ---- CPU baseline beside it (rank 0, N = 1): the oracle on a bounded sample of the same step, same state
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O   # checker / CPU baseline only
        cores = os.cpu_count() or 1
        O.lib().oracle_set_threads(cores)
        o = O.OracleMPS(n, max_bond=chi, gesdd=True, gauge=args.gauge)
        for k in range(n):
            o.set_site(k, eng.get_site(k))
        circ = steps[-1]
        o.apply("CNOT", (n // 2, n // 2 + 1), ())   # warm the BLAS threads
        t0 = time.perf_counter()
        for g in circ:
            o.apply(g[0], g[1], g[2])
        dt = time.perf_counter() - t0
        cpu = {"value": n2_step / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "one full step (%d 1q + %d 2q gates, the step the GPU arm times) on the GPU run's own chi=%d state; oracle zgemm+zgesdd, scipy OpenBLAS" % (2 * n, n2_step, chi)}

    if rank == 0:
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("jacobi_sweep_dram_bytes_per_launch")   # ncu capture, see the file
            except Exception:
                traffic = None
        svd_tf = sv_fl / (ms_svd * 1e-3) / 1e12 if ms_svd > 0 else None
        th_tf = th_fl / (ms_theta * 1e-3) / 1e12 if ms_theta > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
            "data": "synthetic",
            "config": {"workload": "c2_brickwork_n%d_chi%d" % (n, chi), "qubits": n, "max_bond_dim": chi,
                       "step": "2 brickwork layers = %d 1q + %d 2q gates on the saturated state (%d of them with all three bonds at chi)" % (2 * n, n2_step, sat // max(1, args.steps)),
                       "gauge": "reference" if args.gauge == 0 else "canonical", "sharding": "independent circuits per GPU" if world > 1 else "none",
                       "l2": "working set (state %.0f MB + per-layer workspace) exceeds the 126 MB L2; no explicit flush" % (sum(2 * a * b * 16 for a, b in zip([1] + bonds, bonds + [1])) / 1e6),
                       "norm_after_depth%d" % args.depth: norm0},
            "clocks": clocks,
            "gpu_launches": launches,
            "e2e": e2e,
            "roofline": {"bound": "tensor", "kernel": "SVD phase: qr_panel/qr_update (pre-reduction) + jacobi_sweep_kernel", "achieved": svd_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": (svd_tf / fp64_peak) if (svd_tf and fp64_peak) else None, "traffic": traffic,
                         "executed_dmma_tflops_jacobi": jac_tf, "executed_frac_jacobi": (jac_tf / fp64_peak) if fp64_peak else None,
                         "note": "achieved = LAPACK-equivalent SVD flops (88 M N min(M,N) per gate) / CUDA-event time of the SVD phase; executed_dmma_tflops_jacobi = flops the Jacobi pair tasks really issued on the DMMA pipe (Gram 10 + apply 32 DMMA per row chunk, counted on the device) / time of the Jacobi sweeps; peak = cuBLAS DGEMM 8192^3 measured live in this run (MEASURED_PEAKS.json has no FP64 figure)",
                         "share_of_step": ms_svd / max(1e-9, ms_theta + ms_svd + ms_wb)},
            "roofline_theta": {"bound": "tensor", "kernel": "zgemm_dmma_kernel<theta>", "achieved": th_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                               "frac": (th_tf / fp64_peak) if (th_tf and fp64_peak) else None},
            "phases_ms_per_step": {"theta": ms_theta / args.steps, "svd": ms_svd / args.steps, "writeback": ms_wb / args.steps, "qr_prereduce_within_svd": ms_qr / args.steps,
                                   "jacobi_sweeps_per_layer": sweeps_per_layer},
            "circuit": {"name": "brickwork n=%d depth=%d chi<=%d from |0>" % (n, args.depth, chi), "wall_ms": circuit_ms, "gates_1q": n1_0, "gates_2q": n2_0,
                        "gates_2q_per_s": n2_0 / (circuit_ms * 1e-3), "launches": int(st1["launches"] - st0["launches"])},
            "roofline_theta_chi512": theta512,
            "e2e_visitor": e2e_visitor,
            "circuit_sharded": sharded,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
