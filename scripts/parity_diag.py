"""Diagnostic (GPU box): free-running GPU vs oracle deviations on the full-shape configs, oracle with LAPACK zgesvd / zgesdd and
with the engine's numerically-null rule mirrored.  Prints one JSON line per config."""
import json, os, sys, math, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tnqvm_b200
from tnqvm_b200 import circuits as Cc
from oracle import oracle as O

O.lib().oracle_set_threads(os.cpu_count() or 1)
which = sys.argv[1:] or ["c3", "c5", "c2"]
for w in which:
    if w == "c2":
        n, chi = 50, 256; circ = Cc.brickwork(n, 20, seed=12345)
    elif w == "c2s":
        n, chi = 30, 32; circ = Cc.brickwork(n, 20, seed=12345)
    elif w == "c3":
        n, chi = 100, 32; circ = Cc.nearest_neighbor(Cc.qaoa_ring(n, 4, seed=7))
    elif w == "c5":
        n, raw = Cc.sycamore_53(14); chi = 32; circ = Cc.nearest_neighbor(raw)
    out = {"config": w, "n": n, "chi": chi}
    for nt in (None, 0.0, 1e-300):
        e = tnqvm_b200.B200MPS(n, max_bond=chi)
        if nt is not None:
            e.set_option("null_tol", nt)
        t0 = time.time(); e.run(circ); zg = e.expval_z_all(); ng = e.norm(); tg = time.time() - t0
        bg = np.asarray(e.bond_dims()); st = e.stats()
        out["gpu_null_%s" % nt] = {"norm": ng, "dw": e.discarded_weight(), "nonconv": st["svd_nonconverged"], "s": tg, "maxbond": int(bg.max()), "sumbond": int(bg.sum())}
        if nt is None:
            z0, n0, b0 = zg, ng, bg
        e.close()
    for name, kw in (("gesvd", dict(gesdd=False)), ("gesdd", dict(gesdd=True)), ("gesdd_null", dict(gesdd=True, null_tol=2e-14)), ("gesvd_null", dict(gesdd=False, null_tol=2e-14))):
        t0 = time.time()
        o = O.OracleMPS(n, max_bond=chi, **kw).run(circ)
        zo = np.array([o.expval_z([k]) for k in range(n)]); no = o.norm(); bo = np.asarray(o.bond_dims())
        out[name] = {"norm": no, "dw": o.discarded_weight(), "rel_dnorm": abs(n0 - no) / no, "max_dz_over_norm": float(np.abs(z0 - zo).max() / no),
                     "bond_mismatch": int((b0 != bo).sum()), "sumbond": int(bo.sum()), "s": time.time() - t0}
    print(json.dumps(out), flush=True)
