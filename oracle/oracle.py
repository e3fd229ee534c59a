"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

ctypes front-end for
  * oracle/libmps_oracle.so   CPU restatement of the exatn-mps visitor algorithm
                              (see mps_oracle.cpp for the reference file:line map), and
  * oracle/_ref/libdense_ref.so  the reference's own header-only dense simulator / sampler
                              (tnqvm/base/Gates.hpp, tnqvm/utils/GateMatrixAlgebra.hpp,
                              tnqvm/utils/RandomEngine.hpp), compiled where the sources lie.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
"""
import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "libmps_oracle.so")
    src = os.path.join(_HERE, "mps_oracle.cpp")
    need = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    ref_so = os.path.join(_HERE, "_ref", "libdense_ref.so")
    if os.path.isdir("/root/reference/tnqvm/base") and not os.path.exists(ref_so):
        need = True
    if need:
        subprocess.check_call(["make", "-C", _HERE, "CXX=g++"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)


def _blas_path():
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so"))
    if not cands:
        raise RuntimeError("scipy OpenBLAS not found")
    return os.path.abspath(cands[0])


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "libmps_oracle.so"))
        L.oracle_init_blas.argtypes = [C.c_char_p]
        rc = L.oracle_init_blas(_blas_path().encode())
        if rc != 0:
            raise RuntimeError("oracle_init_blas failed: %d" % rc)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_set_renorm.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_null_tol.argtypes = [C.c_void_p, C.c_double]
        L.oracle_apply_1q.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_apply_2q.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_bond_dims.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_singular_values.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_discarded_weight.restype = C.c_double
        L.oracle_discarded_weight.argtypes = [C.c_void_p]
        L.oracle_fidelity_estimate.restype = C.c_double
        L.oracle_fidelity_estimate.argtypes = [C.c_void_p]
        L.oracle_times.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_site.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_set_site.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.oracle_statevector.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_norm.restype = C.c_double
        L.oracle_norm.argtypes = [C.c_void_p]
        L.oracle_expval_z.restype = C.c_double
        L.oracle_expval_z.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_amplitude.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_measure.argtypes = [C.c_void_p, C.c_int]
        L.oracle_clear_measure.argtypes = [C.c_void_p]
        L.oracle_sample_statevector.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_sample_rdm_shot.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_gate_matrix.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def gate_matrix(name, params=()):
    """Restated tnqvm/base/Gates.hpp matrices (row = output)."""
    p = np.asarray(list(params) + [0.0] * 3, dtype=np.float64)
    out = np.zeros(16, dtype=np.complex128)
    dim = lib().oracle_gate_matrix(name.encode(), p.ctypes.data, len(params), out.ctypes.data)
    return out[: dim * dim].reshape(dim, dim).copy()


class OracleMPS:
    """CPU restatement of ExatnMpsVisitor (reference gauge by default)."""

    def __init__(self, n, max_bond=0, svd_cutoff=-1.0, cutoff_on_sqrt=False, gesdd=False, gauge=0, seed=None, null_tol=0.0):
        self.L = lib()
        self.n = n
        self.h = self.L.oracle_create(n, int(max_bond), float(svd_cutoff), int(cutoff_on_sqrt), int(gesdd), int(gauge))
        if null_tol > 0:   # NOT the reference: mirrors the engine's numerically-null rule
            self.L.oracle_set_null_tol(self.h, float(null_tol))
        if seed is not None:
            self.seed(seed)

    def __del__(self):
        try:
            if self.h:
                self.L.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def seed(self, s):
        self.L.oracle_seed(self.h, int(s))

    def apply_1q(self, q, m):
        m = np.ascontiguousarray(m, dtype=np.complex128)
        self.L.oracle_apply_1q(self.h, q, m.ctypes.data)

    def apply_2q(self, q0, q1, m):
        m = np.ascontiguousarray(m, dtype=np.complex128)
        rc = self.L.oracle_apply_2q(self.h, q0, q1, m.ctypes.data)
        if rc:
            raise RuntimeError("oracle_apply_2q rc=%d" % rc)

    def apply(self, name, qubits, params=()):
        if name == "Measure":
            self.L.oracle_measure(self.h, qubits[0])
            return
        if name == "I":
            return
        m = gate_matrix(name, params)
        if name == "Swap" and len(qubits) == 2 and qubits[0] < qubits[1]:
            qubits = (qubits[1], qubits[0])   # visit(Swap), ExaTnMpsVisitor.cpp:1030-1033
        if m.shape[0] == 2:
            self.apply_1q(qubits[0], m)
        else:
            self.apply_2q(qubits[0], qubits[1], m)

    def run(self, circuit):
        for g in circuit:
            self.apply(g[0], g[1], g[2] if len(g) > 2 else ())
        return self

    def bond_dims(self):
        out = np.zeros(max(self.n - 1, 1), dtype=np.int32)
        self.L.oracle_bond_dims(self.h, out.ctypes.data)
        return out[: self.n - 1]

    def singular_values(self, bond):
        out = np.zeros(1 << 14, dtype=np.float64)
        c = self.L.oracle_singular_values(self.h, bond, out.ctypes.data, out.size)
        return out[:c].copy()

    def discarded_weight(self):
        return self.L.oracle_discarded_weight(self.h)

    def fidelity_estimate(self):
        return self.L.oracle_fidelity_estimate(self.h)

    def times(self):
        out = np.zeros(3)
        self.L.oracle_times(self.h, out.ctypes.data)
        return dict(gemm=out[0], svd=out[1], trunc=out[2])

    def get_site(self, k):
        shp = np.zeros(3, dtype=np.int32)
        self.L.oracle_get_site(self.h, k, None, shp.ctypes.data)
        out = np.zeros(int(shp.prod()), dtype=np.complex128)
        self.L.oracle_get_site(self.h, k, out.ctypes.data, shp.ctypes.data)
        return out.reshape(tuple(int(x) for x in shp), order="F")

    def set_site(self, k, t):
        t = np.asfortranarray(t, dtype=np.complex128)
        self.L.oracle_set_site(self.h, k, t.ctypes.data, t.shape[0], t.shape[2])

    def statevector(self):
        out = np.zeros(1 << self.n, dtype=np.complex128)
        rc = self.L.oracle_statevector(self.h, out.ctypes.data)
        assert rc == 0
        return out

    def norm(self):
        return self.L.oracle_norm(self.h)

    def expval_z(self, qubits):
        q = np.asarray(qubits, dtype=np.int32)
        return self.L.oracle_expval_z(self.h, q.ctypes.data, len(q))

    def amplitude(self, bits):
        b = np.asarray(bits, dtype=np.int8)
        out = np.zeros(1, dtype=np.complex128)
        self.L.oracle_amplitude(self.h, b.ctypes.data, out.ctypes.data)
        return complex(out[0])

    def measure(self, q):
        self.L.oracle_measure(self.h, q)

    def sample(self, shots, nq):
        """finalize() sampling: n<20 -> GenerateSamples on the state vector, else per-shot RDM."""
        if self.n < 20:
            buf = C.create_string_buffer(shots * nq + 1)
            m = self.L.oracle_sample_statevector(self.h, shots, buf)
            raw = buf.raw
            return [raw[i * nq:(i + 1) * nq].decode() for i in range(m)]
        res = []
        for _ in range(shots):
            buf = C.create_string_buffer(nq + 1)
            self.L.oracle_sample_rdm_shot(self.h, buf)
            res.append(buf.raw[:nq].decode())
        return res


# ------------------------------------------------------------------ reference's own dense simulator
def ref():
    global _REF
    if _REF is None:
        build()
        p = os.path.join(_HERE, "_ref", "libdense_ref.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_gate_matrix.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
        R.ref_apply_1q.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p]
        R.ref_apply_cnot.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        R.ref_set_seed.argtypes = [C.c_uint64]
        R.ref_rand_prob.restype = C.c_double
        R.ref_generate_samples.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _REF = R
    return _REF


def ref_gate_matrix(name, params=()):
    p = np.asarray(list(params) + [0.0] * 3, dtype=np.float64)
    out = np.zeros(16, dtype=np.complex128)
    dim = ref().ref_gate_matrix(name.encode(), p.ctypes.data, out.ctypes.data)
    return out[: dim * dim].reshape(dim, dim).copy()


def dense_apply(state, n, name, qubits, params=(), use_ref=True):
    """Dense state-vector gate application, qubit 0 = LSB.  1q gates and CNOT go through the
    reference's own ApplySingleQubitGate / ApplyCNOTGate when oracle/_ref is built; other 2q
    gates use the restated matrix with index = 2*bit(q0)+bit(q1) (ExaTnMpsVisitor.cpp:1492-1499)."""
    R = ref() if use_ref else None
    p = np.asarray(list(params) + [0.0] * 3, dtype=np.float64)
    if name in ("I", "Measure"):
        return state
    nm = "CNOT" if name == "CX" else name
    if len(qubits) == 1 and R is not None:
        R.ref_apply_1q(state.ctypes.data, n, qubits[0], nm.encode(), p.ctypes.data)
        return state
    if nm == "CNOT" and R is not None:
        R.ref_apply_cnot(state.ctypes.data, n, qubits[0], qubits[1])
        return state
    m = gate_matrix(nm, params)
    psi = state.reshape([2] * n)            # axis k <-> qubit n-1-k (C order, qubit 0 fastest)
    if len(qubits) == 1:
        ax = n - 1 - qubits[0]
        psi = np.moveaxis(np.tensordot(m, psi, axes=([1], [ax])), 0, ax)
    else:
        a0, a1 = n - 1 - qubits[0], n - 1 - qubits[1]
        g = m.reshape(2, 2, 2, 2)           # [b0', b1', b0, b1], b0 = bit(q0) is the MSB
        psi = np.moveaxis(np.tensordot(g, psi, axes=([2, 3], [a0, a1])), [0, 1], [a0, a1])
    state[:] = np.ascontiguousarray(psi).reshape(-1)
    return state


def dense_run(n, circuit, use_ref=True):
    state = np.zeros(1 << n, dtype=np.complex128)
    state[0] = 1.0
    for g in circuit:
        dense_apply(state, n, g[0], g[1], g[2] if len(g) > 2 else (), use_ref)
    return state


def dense_expval_z(state, n, qubits):
    idx = np.arange(1 << n, dtype=np.uint64)
    par = np.zeros(1 << n, dtype=np.int64)
    for q in qubits:
        par ^= ((idx >> np.uint64(q)) & np.uint64(1)).astype(np.int64)
    return float(np.sum((1 - 2 * par) * np.abs(state) ** 2))


if __name__ == "__main__":
    build(force=True)
    print("oracle built;", "ref present" if ref() is not None else "ref absent", file=sys.stderr)
