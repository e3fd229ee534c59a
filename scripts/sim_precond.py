"""Dev tool: sweep counts of the kernel's block Jacobi (8-column blocks, cross-pairs-only rule) under different
preconditioners of theta: plain QR (current), column sorting, pivoted QR, two QRs.  CPU only."""
import sys, os, math, time
import numpy as np, scipy.linalg as sl
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sim_jacobi import thetas, block_jacobi_order, order_round_robin, block_jacobi_fast

def variants(T):
    out = {}
    R = np.linalg.qr(T, mode='r'); out["qr"] = R.conj().T.copy()
    p = np.argsort(-np.linalg.norm(T, axis=0)); R = np.linalg.qr(T[:, p], mode='r'); out["sort+qr"] = R.conj().T.copy()
    R = sl.qr(T, mode='r', pivoting=True)[0]; out["qrcp"] = R[:T.shape[1]].conj().T.copy()
    R1 = np.linalg.qr(T, mode='r'); R2 = np.linalg.qr(R1.conj().T, mode='r'); out["qr+qr"] = R2.conj().T.copy()
    R1 = sl.qr(T, mode='r', pivoting=True)[0][:T.shape[1]]; R2 = np.linalg.qr(R1.conj().T, mode='r'); out["qrcp+qr"] = R2.conj().T.copy()
    R3 = np.linalg.qr(R2.conj().T, mode='r'); out["qrcp+qr+qr"] = R3.conj().T.copy()
    return out

if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    ths = thetas(n=n, depth=depth, chi=chi, seed=3)
    print("chi", chi, "thetas", len(ths), ths[0].shape, flush=True)
    for T in ths[-3:]:
        sref = np.linalg.svd(T, compute_uv=False)
        print("spectrum s0 %.2e s[chi] %.2e smin %.2e" % (sref[0], sref[chi], sref[-1]))
        for name, X in variants(T).items():
            steps = order_round_robin(X.shape[1] // 8)
            t0 = time.time()
            s, hist = block_jacobi_order(X, steps)
            s2, hist2 = block_jacobi_fast(X, b=8, tol=math.sqrt(X.shape[0]) * 2.2e-16)
            print("  %-12s sweeps %2d %s | full inner eig: %2d %s  relerr %.1e (%.0fs)" % (name, len(hist), hist, len(hist2), hist2, np.max(np.abs(s - sref)) / sref[0], time.time() - t0), flush=True)
