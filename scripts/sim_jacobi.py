"""CPU emulation of block one-sided Jacobi variants to study outer-sweep counts on realistic thetas (dev tool)."""
import sys, os, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from tnqvm_b200 import circuits as Cc
from tnqvm_b200.gates import gate_matrix

def thetas(n=14, depth=16, chi=32, seed=3, gauge=0):
    """Collect thetas (after gate) seen at saturated bonds."""
    o = O.OracleMPS(n, max_bond=chi, gauge=gauge)
    circ = Cc.brickwork(n, depth, seed=seed)
    out = []
    for g in circ:
        if len(g[1]) == 2:
            lo = min(g[1])
            A, B = o.get_site(lo), o.get_site(lo + 1)
            if A.shape[0] == chi and A.shape[2] == chi and B.shape[2] == chi:
                m = gate_matrix(g[0], g[2]).reshape(2, 2, 2, 2)
                if g[1][0] > g[1][1]:
                    m = m.transpose(1, 0, 3, 2)
                th = np.einsum('pqij,aijc->apqc', m, np.einsum('aik,kjc->aijc', A, B))
                out.append(th.reshape(2 * chi, 2 * chi, order='F'))
        o.apply(g[0], g[1], g[2])
    return out

def jacobi_eig_sweeps(W, tol, max_sweeps, dead):
    """two-sided cyclic Jacobi on Hermitian W; returns Q, nrot"""
    n = W.shape[0]
    W = W.copy(); Q = np.eye(n, dtype=complex); nrot = 0
    for s in range(max_sweeps):
        rot = 0
        for p in range(n - 1):
            for q in range(p + 1, n):
                a, b, g = W[p, p].real, W[q, q].real, W[p, q]
                g2 = abs(g) ** 2
                if a > dead and b > dead and g2 > tol * tol * a * b:
                    d = b - a
                    h = math.sqrt(d * d + 4 * g2)
                    u = (2.0 if d >= 0 else -2.0) / (abs(d) + h)
                    c = 1 / math.sqrt(1 + u * u * g2)
                    sg = c * u * g
                    J = np.array([[c, sg], [-np.conj(sg), c]])
                    W[:, [p, q]] = W[:, [p, q]] @ J
                    W[[p, q], :] = J.conj().T @ W[[p, q], :]
                    Q[:, [p, q]] = Q[:, [p, q]] @ J
                    rot += 1
        nrot += rot
        if rot == 0:
            break
    return Q, nrot

def block_jacobi(T, b=8, inner=1, tol=None, max_outer=60, sort=False, verbose=False):
    M, N = T.shape
    X = T.copy()
    if sort:
        X = X[:, np.argsort(-np.linalg.norm(X, axis=0))]
    tol = tol or math.sqrt(M) * 2.2e-16
    nb = (N + b - 1) // b
    nbe = nb + (nb & 1)
    fro2 = np.linalg.norm(X) ** 2
    dead = (10 * tol) ** 2 * fro2 / N
    hist = []
    for sweep in range(max_outer):
        dirty = 0
        for step in range(nbe - 1):
            for pi in range(nbe // 2):
                if pi == 0: A_, B_ = nbe - 1, step
                else: A_, B_ = (step + pi) % (nbe - 1), (step + nbe - 1 - pi) % (nbe - 1)
                cols = [c for blk in (A_, B_) if blk < nb for c in range(blk * b, min(N, blk * b + b))]
                if not cols: continue
                Xs = X[:, cols]
                W = Xs.conj().T @ Xs
                d = np.real(np.diag(W))
                off = np.abs(W) ** 2 > tol * tol * np.outer(d, d)
                np.fill_diagonal(off, False)
                off &= np.outer(d > dead, d > dead)
                if not off.any(): continue
                dirty += 1
                Q, _ = jacobi_eig_sweeps(W, tol, inner, dead)
                X[:, cols] = Xs @ Q
        hist.append(dirty)
        if dirty == 0: break
    s = np.sort(np.linalg.norm(X, axis=0))[::-1]
    return s, hist

if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    gauge = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ths = thetas(chi=chi, gauge=gauge)
    print("collected", len(ths), "thetas", ths[0].shape)
    for T in ths[-3:]:
        sref = np.linalg.svd(T, compute_uv=False)
        print("cond: s0 %.3e s[chi] %.3e smin %.3e" % (sref[0], sref[chi], sref[-1]))
        for (b, inner, tol, sort) in [(8, 1, None, False), (8, 1, 1e-13, False), (8, 30, None, False), (8, 30, 1e-13, False), (16, 30, 1e-13, False), (32, 30, 1e-13, False), (8, 30, 1e-13, True)]:
            s, hist = block_jacobi(T, b=b, inner=inner, tol=tol, sort=sort)
            print("b=%2d inner=%2d tol=%s sort=%d: outer sweeps %2d  hist %s  relerr %.2e" % (b, inner, tol, sort, len(hist), hist, np.max(np.abs(s - sref) / sref[0])))

def eigh_Q(W):
    w, Q = np.linalg.eigh(W)
    # order columns so Q is as close to identity as possible (Jacobi-like, no gratuitous permutation)
    n = W.shape[0]
    A = np.abs(Q) ** 2
    perm = -np.ones(n, dtype=int); used = np.zeros(n, bool)
    for _ in range(n):
        i, j = np.unravel_index(np.argmax(np.where(used[None, :] | (perm[:, None] >= 0), -1, A)), A.shape)
        perm[i] = j; used[j] = True
    return Q[:, perm]

def block_jacobi_fast(T, b=8, tol=1e-13, max_outer=60, order="rr", qr=False):
    M, N = T.shape
    X = T.copy()
    if qr:
        # QR preconditioning: T = Q R ; work on L = R^H (lower triangular) columns
        R = np.linalg.qr(X, mode='r')
        X = R.conj().T.copy()
    nb = (N + b - 1) // b
    nbe = nb + (nb & 1)
    fro2 = np.linalg.norm(X) ** 2
    dead = (10 * tol) ** 2 * fro2 / N
    hist = []
    for sweep in range(max_outer):
        dirty = 0
        for step in range(nbe - 1):
            for pi in range(nbe // 2):
                if pi == 0: A_, B_ = nbe - 1, step
                else: A_, B_ = (step + pi) % (nbe - 1), (step + nbe - 1 - pi) % (nbe - 1)
                cols = [c for blk in (A_, B_) if blk < nb for c in range(blk * b, min(N, blk * b + b))]
                if not cols: continue
                Xs = X[:, cols]
                W = Xs.conj().T @ Xs
                d = np.real(np.diag(W))
                off = np.abs(W) ** 2 > tol * tol * np.outer(d, d)
                np.fill_diagonal(off, False)
                off &= np.outer(d > dead, d > dead)
                if not off.any(): continue
                dirty += 1
                X[:, cols] = Xs @ eigh_Q(W)
        hist.append(dirty)
        if dirty == 0: break
    s = np.sort(np.linalg.norm(X, axis=0))[::-1]
    return s, hist

def study(chi, gauge=0, n=None, depth=None):
    n = n or (2 * int(math.log2(chi)) + 6)
    depth = depth or (2 * int(math.log2(chi)) + 8)
    ths = thetas(n=n, depth=depth, chi=chi, gauge=gauge)
    print("chi", chi, "collected", len(ths), "thetas", ths[0].shape)
    for T in ths[-2:]:
        sref = np.linalg.svd(T, compute_uv=False)
        print("cond: s0 %.3e s[chi] %.3e smin %.3e" % (sref[0], sref[chi], sref[-1]))
        for qr in (False, True):
            for b in (8, 16, 32, 64):
                if 2 * b > T.shape[1]: continue
                s, hist = block_jacobi_fast(T, b=b, qr=qr)
                print("  qr=%d b=%2d nb=%2d: outer sweeps %2d hist %s relerr %.1e" % (qr, b, T.shape[1] // b, len(hist), hist, np.max(np.abs(s - sref)) / sref[0]))


# ---------------------------------------------------------------------------------------------------------------------
# Orderings for the block tournament.  Each returns the list of steps of ONE sweep; a step
# is a perfect matching [(A, B), ...] of the nbe blocks.
def order_round_robin(nbe):
    steps = []
    for step in range(nbe - 1):
        st = []
        for pi in range(nbe // 2):
            if pi == 0: st.append((nbe - 1, step))
            else: st.append(((step + pi) % (nbe - 1), (step + nbe - 1 - pi) % (nbe - 1)))
        steps.append(st)
    return steps


def order_resident(nbe):
    """Recursive bipartite tournament (nbe a power of two): split the blocks into halves R and T; for m = |R| steps pair
    R_i with T_(i+s) -- block R_i never leaves its task (it can stay in shared memory), the T blocks stream past; then both
    halves recurse concurrently.  nbe - 1 steps per sweep, every step a perfect matching, every pair once."""
    assert nbe & (nbe - 1) == 0
    def rec(groups):
        m = len(groups[0]) // 2
        if m == 0:
            return []
        steps = []
        for s in range(m):
            st = []
            for g in groups:
                R, T = g[:m], g[m:]
                st += [(R[i], T[(i + s) % m]) for i in range(m)]
            steps.append(st)
        return steps + rec([h for g in groups for h in (g[:m], g[m:])])
    return rec([list(range(nbe))])


def block_jacobi_order(T, steps_of_sweep, b=8, tol=None, max_outer=60):
    """One-sided block Jacobi with the cross-pairs-only rule of the kernel: the pairs inside a block are rotated at the first
    step of a sweep only (`within`), every later step rotates the 64 cross pairs of its two blocks (one cyclic pass)."""
    M, N = T.shape
    X = T.copy()
    tol = tol or math.sqrt(M) * 2.2e-16
    fro2 = np.linalg.norm(X) ** 2
    dead = (10 * tol) ** 2 * fro2 / N
    hist = []
    for sweep in range(max_outer):
        dirty = 0
        for si, st in enumerate(steps_of_sweep):
            for (A_, B_) in st:
                cols = list(range(A_ * b, A_ * b + b)) + list(range(B_ * b, B_ * b + b))
                Xs = X[:, cols]
                W = Xs.conj().T @ Xs
                Q = np.eye(2 * b, dtype=complex)
                rot = 0
                pairs = [(p, q) for p in range(b) for q in range(b, 2 * b)]
                if si == 0:
                    pairs = [(p, q) for p in range(2 * b) for q in range(p + 1, 2 * b)]
                for (p, q) in pairs:
                    a, c, g = W[p, p].real, W[q, q].real, W[p, q]
                    g2 = abs(g) ** 2
                    if a > dead and c > dead and g2 > tol * tol * a * c:
                        d = c - a
                        h = math.sqrt(d * d + 4 * g2)
                        u = (2.0 if d >= 0 else -2.0) / (abs(d) + h)
                        cs = 1 / math.sqrt(1 + u * u * g2)
                        sg = cs * u * g
                        J = np.array([[cs, sg], [-np.conj(sg), cs]])
                        W[:, [p, q]] = W[:, [p, q]] @ J
                        W[[p, q], :] = J.conj().T @ W[[p, q], :]
                        Q[:, [p, q]] = Q[:, [p, q]] @ J
                        rot += 1
                if rot:
                    dirty += 1
                    X[:, cols] = Xs @ Q
        hist.append(dirty)
        if dirty == 0:
            break
    s = np.sort(np.linalg.norm(X, axis=0))[::-1]
    return s, hist


def study_orderings(chi=64, n=16, depth=20, count=4):
    ths = thetas(n=n, depth=depth, chi=chi, seed=3)
    print("chi", chi, "thetas", len(ths), ths[0].shape)
    for T in ths[-count:]:
        sref = np.linalg.svd(T, compute_uv=False)
        X = np.linalg.qr(T, mode='r').conj().T.copy()
        nbe = X.shape[1] // 8
        for name, order in (("round_robin", order_round_robin), ("resident", order_resident)):
            steps = order(nbe)
            assert len(steps) == nbe - 1 and len({tuple(sorted(p)) for st in steps for p in st}) == nbe * (nbe - 1) // 2
            s, hist = block_jacobi_order(X, steps)
            print("  %-12s sweeps %2d  dirty tasks per sweep %s  relerr %.1e" % (name, len(hist), hist, np.max(np.abs(s - sref)) / sref[0]))


# ---------------------------------------------------------------------------------------------------------------------
# Mixed precision: complex64 sweeps that accumulate V, V re-unitarised in complex128 by
# Newton-Schulz, G V formed in complex128, complex128 polishing sweeps.  How many polishing sweeps remain?
def fp32_presweeps(G, b=8, max_outer=12):
    M, N = G.shape
    X = G.astype(np.complex64)
    V = np.eye(N, dtype=np.complex64)
    tol = np.float32(math.sqrt(M) * 6e-8)
    nbe = N // b
    steps = order_round_robin(nbe)
    hist = []
    for sweep in range(max_outer):
        dirty = 0
        for si, st in enumerate(steps):
            for (A_, B_) in st:
                cols = list(range(A_ * b, A_ * b + b)) + list(range(B_ * b, B_ * b + b))
                Xs = X[:, cols]
                W = (Xs.conj().T @ Xs).astype(np.complex64)
                Q = np.eye(2 * b, dtype=np.complex64)
                rot = 0
                pairs = [(p, q) for p in range(b) for q in range(b, 2 * b)] if si else [(p, q) for p in range(2 * b) for q in range(p + 1, 2 * b)]
                for (p, q) in pairs:
                    a, c, g = np.float32(W[p, p].real), np.float32(W[q, q].real), W[p, q]
                    g2 = np.float32(abs(g) ** 2)
                    if g2 > tol * tol * a * c:
                        d = np.float32(c - a)
                        h = np.float32(math.sqrt(d * d + 4 * g2))
                        u = np.float32((2.0 if d >= 0 else -2.0) / (abs(d) + h))
                        cs = np.float32(1 / math.sqrt(1 + u * u * g2))
                        sg = np.complex64(cs * u * g)
                        J = np.array([[cs, sg], [-np.conj(sg), cs]], dtype=np.complex64)
                        W[:, [p, q]] = W[:, [p, q]] @ J
                        W[[p, q], :] = J.conj().T @ W[[p, q], :]
                        Q[:, [p, q]] = Q[:, [p, q]] @ J
                        rot += 1
                if rot:
                    dirty += 1
                    X[:, cols] = Xs @ Q
                    V[:, cols] = V[:, cols] @ Q
        hist.append(dirty)
        if dirty == 0:
            break
    return V, hist


def study_mixed(chi=64, n=16, depth=20, count=3):
    ths = thetas(n=n, depth=depth, chi=chi, seed=3)
    for T in ths[-count:]:
        sref = np.linalg.svd(T, compute_uv=False)
        G = np.linalg.qr(T, mode='r').conj().T.copy()
        N = G.shape[1]
        steps = order_round_robin(N // 8)
        _, h64 = block_jacobi_order(G, steps)
        for pre in (6, 8, 12):
            V32, h32 = fp32_presweeps(G, max_outer=pre)
            V = V32.astype(np.complex128)
            u0 = np.linalg.norm(V.conj().T @ V - np.eye(N))
            for _ in range(2):
                V = V @ (1.5 * np.eye(N) - 0.5 * (V.conj().T @ V))
            u2 = np.linalg.norm(V.conj().T @ V - np.eye(N))
            s, hp = block_jacobi_order(G @ V, steps)
            print("fp64 only: %d sweeps | fp32 pre-sweeps %2d %s  unitarity %.1e -> %.1e | fp64 polish sweeps %d %s  relerr %.1e"
                  % (len(h64), len(h32), h32[-3:], u0, u2, len(hp), hp, np.max(np.abs(s - sref)) / sref[0]))
