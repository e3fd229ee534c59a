"""Developer check: one 2q gate on random sites of chosen bond dims; compares the new site pair with numpy."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqvm_b200
from tnqvm_b200.gates import gate_matrix

rng = np.random.default_rng(0)
def rnd(*s): return (rng.standard_normal(s) + 1j * rng.standard_normal(s)) / np.sqrt(2 * s[0] * s[-1])

def one(cl, ch, cr, gate="fSim", params=(0.7, 0.3), rev=False, max_bond=0):
    e = tnqvm_b200.B200MPS(4, max_bond=max_bond)
    S = [rnd(1, 2, cl), rnd(cl, 2, ch), rnd(ch, 2, cr), rnd(cr, 2, 1)]
    for k in range(4): e.set_site(k, S[k])
    m = gate_matrix(gate, params)
    q = (2, 1) if rev else (1, 2)
    e.apply_2q(q[0], q[1], m); e.sync()
    A, B = e.get_site(1), e.get_site(2)
    D = np.einsum('apk,kqc->apqc', S[1], S[2])
    g = m.reshape(2, 2, 2, 2)
    if rev: th = np.einsum('qpji,aijc->apqc', g, D)   # index = 2*bit(q0=hi)+bit(q1=lo)
    else: th = np.einsum('pqij,aijc->apqc', g, D)
    got = np.einsum('apk,kqc->apqc', A, B)
    if max_bond:
        U, sv_, Vh = np.linalg.svd(th.reshape(2 * cl, 2 * cr), full_matrices=False)
        kk = min(max_bond, len(sv_))
        th = ((U[:, :kk] * sv_[:kk]) @ Vh[:kk]).reshape(cl, 2, 2, cr)
    s = e.singular_values(1)
    sref = np.linalg.svd(th.reshape(2 * cl, 2 * cr), compute_uv=False)
    r = len(s)
    st = e.stats()
    print("dims", (cl, ch, cr), "rev", rev, "bond", A.shape[2], "recon err %.3e" % np.abs(got - th).max(),
          "sv err %.3e" % np.abs(s - sref[:r]).max(), "sweeps", st["jacobi_sweeps"], flush=True)
    e.close()

for d in [(4, 4, 4), (16, 16, 16), (32, 32, 32), (33, 33, 33), (32, 33, 32), (33, 32, 32), (32, 32, 33), (40, 20, 40), (64, 64, 64), (64, 32, 16), (16, 32, 64),
          (100, 100, 100), (128, 128, 128)]:
    one(*d)
one(64, 64, 64, rev=True)
one(48, 64, 80, gate="CNOT", params=())

one(64, 64, 64, max_bond=64)
one(64, 64, 64, max_bond=100)
one(128, 128, 128, max_bond=128)
one(32, 64, 32, max_bond=16)
print("rank-deficient cases")
one(32, 4, 32)
one(32, 1, 32)
one(17, 3, 33)
one(33, 3, 17)
def ortho(cl, ch, cr):
    e = tnqvm_b200.B200MPS(4)
    S = [rnd(1, 2, cl), rnd(cl, 2, ch), rnd(ch, 2, cr), rnd(cr, 2, 1)]
    for k in range(4): e.set_site(k, S[k])
    e.apply_2q(1, 2, gate_matrix("fSim", (0.7, 0.3))); e.sync()
    A = e.get_site(1); L = A.reshape(2 * cl, -1, order="F")
    nrm = np.linalg.norm(L, axis=0)
    Gm = np.abs(L.conj().T @ L) / np.maximum(np.outer(nrm, nrm), 1e-300)
    np.fill_diagonal(Gm, 0)
    print("ortho", (cl, ch, cr), "col norms", nrm[:3], nrm[-3:], "max rel offdiag %.2e" % Gm.max(), "argmax", np.unravel_index(Gm.argmax(), Gm.shape), e.stats()["jacobi_sweeps"])
ortho(32, 4, 32); ortho(32, 1, 32)
