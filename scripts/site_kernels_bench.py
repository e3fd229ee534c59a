"""HBM-side kernels of the path on a saturated chi-bond chain ("1q gate, partial norms: HBM-bound; <Z> sweeps:
DMMA per site, latency along the chain"): achieved GB/s of gate1q_kernel against MEASURED_PEAKS.json, and the time of the
<Z_k>-for-all-k / <Z_i Z_j> / norm transfer sweeps with their algorithmic flops.  Timed with CUDA events on the handle's own
stream; run the same command under `ncu --set full -k regex:gate1q|site_dot|zgemm_small|trace_pair` for the counters.

usage (GPU box): python scripts/site_kernels_bench.py --qubits 50 --chi 256 [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tnqvm_b200
from tnqvm_b200.gates import gate_matrix
from bench import random_mps_sites, bond_profile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=100)   # 176 MB of sites: larger than the 126 MB L2
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    n, chi = a.qubits, a.chi
    e = tnqvm_b200.B200MPS(n, max_bond=chi, fuse_1q=0)   # fuse_1q=0: every 1q gate reaches gate1q_kernel
    for k, t in enumerate(random_mps_sites(n, chi, 7)):
        e.set_site(k, t)
    e.sync()
    stream = torch.cuda.ExternalStream(e.stream())
    dims = [1] + bond_profile(n, chi) + [1]
    site_elems = sum(2 * dims[k] * dims[k + 1] for k in range(n))
    rng = np.random.default_rng(1)

    def timed(fn, reps, before=None):
        if before:
            before()
        fn()
        e.sync()
        t = []
        for _ in range(reps):
            if before:
                before()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            fn()
            a1.record(stream)
            a1.synchronize()
            t.append(a0.elapsed_time(a1))
        return float(np.median(t)), float(np.min(t))

    # ---- one Rx on every qubit = ONE gate1q_kernel launch over all sites (read + write each site once)
    def queue_1q():   # host side only: the gates wait in the handle's queue until the flush
        for q in range(n):
            e.apply_1q(q, gate_matrix("Rx", (float(rng.uniform(-3, 3)),)))

    ms_med, ms_min = timed(e.flush, a.reps, before=queue_1q)
    bytes_1q = 2 * 16 * site_elems   # 64 chi_L chi_R per site
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = None
    for key in ("hbm_gbs", "hbm_gbps"):
        if isinstance(peaks.get(key), (int, float)):
            hbm = float(peaks[key]); break
    out = {"qubits": n, "chi": chi, "state_bytes": 16 * site_elems,
           "gate1q_layer": {"ms_median": ms_med, "ms_min": ms_min, "algorithmic_bytes": bytes_1q, "gbps_median": bytes_1q / ms_med / 1e6,
                            "gbps_best": bytes_1q / ms_min / 1e6, "hbm_peak_gbps": hbm,
                            "frac_of_hbm_peak": (bytes_1q / ms_med / 1e6 / hbm) if hbm else None,
                            "note": "one launch for the whole layer; the timed region is the flush: descriptor upload + gate1q_kernel; "
                                    "state_bytes above says whether the state exceeds the 126 MB L2 (n=100: 176 MB, it does)"}}

    # ---- transfer sweeps: <Z_k> for all k + norm (one left + one right sweep), <Z_i Z_{i+1}> for all bonds, norm alone
    flops_site = lambda k: 2 * 8.0 * dims[k] * dims[k] * 2 * dims[k + 1] + 2 * 8.0 * dims[k] * 2 * dims[k + 1] * dims[k + 1]

    def bump():   # observables are cached per state version: touch the state so that every repetition recomputes
        e.apply_1q(0, gate_matrix("Rz", (0.1,)))
        e.flush()

    ms_z, _ = timed(lambda: (bump(), e.expval_z_all()), max(3, a.reps // 4))
    pairs = [(i, i + 1) for i in range(n - 1)]
    ms_zz, _ = timed(lambda: (bump(), e.expval_zz_pairs(pairs)), max(3, a.reps // 4))
    ms_n, _ = timed(lambda: (bump(), e.norm()), max(3, a.reps // 4))
    fl = sum(flops_site(k) for k in range(n))
    out["expval_z_all"] = {"ms": ms_z, "algorithmic_flops_two_sweeps": 2 * fl, "tflops": 2 * fl / ms_z / 1e9}
    out["expval_zz_all_bonds"] = {"ms": ms_zz, "pairs": len(pairs)}
    out["norm"] = {"ms": ms_n, "algorithmic_flops_one_sweep": fl, "tflops": fl / ms_n / 1e9}
    print(json.dumps(out))
    e.close()


if __name__ == "__main__":
    main()
