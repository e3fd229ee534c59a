"""Python host-side handle over the C ABI (tests, bench, multi-GPU plumbing).  Mirrors the calls the
C++ visitor makes; there is deliberately no CPU path here."""
import ctypes as C

import numpy as np

from . import abi
from .gates import gate_matrix

GAUGE_REFERENCE, GAUGE_LEFT, GAUGE_RIGHT = 0, 1, 2


class CompiledCircuit:
    """A circuit lowered once to the arrays mps_apply_gates takes: the gate-name -> matrix step of the visitor
    (ExatnUtils.cpp:36-133) done ahead of time, visit(Swap)'s bit sort (:1030-1033) included."""

    def __init__(self, circuit, offset=0):
        q0, q1, mats, self.measure = [], [], [], []
        for g in circuit:
            name, qs = g[0], g[1]
            if name == "Measure":
                self.measure.append(qs[0])
                continue
            if name == "I":
                continue
            m = gate_matrix(name, g[2] if len(g) > 2 else ())
            buf = np.zeros(16, dtype=np.complex128)
            if m.shape[0] == 2:
                buf[:4] = m.reshape(-1)
                q0.append(qs[0] + offset); q1.append(-1)
            else:
                if name == "Swap" and qs[0] < qs[1]:
                    qs = (qs[1], qs[0])
                buf[:] = m.reshape(-1)
                q0.append(qs[0] + offset); q1.append(qs[1] + offset)
            mats.append(buf)
        self.count = len(q0)
        self.q0 = np.ascontiguousarray(q0, dtype=np.int32)
        self.q1 = np.ascontiguousarray(q1, dtype=np.int32)
        self.mats = np.ascontiguousarray(mats, dtype=np.complex128).reshape(-1) if mats else np.zeros(0, dtype=np.complex128)


class B200MPS:
    def __init__(self, n_qubits, max_bond=0, svd_cutoff=-1.0, gauge=GAUGE_REFERENCE, device=0, seed=0, n_registers=1,
                 devices=None, partition_by="cost", **options):
        """devices = [d0, d1, ...]: the sites are sharded over these GPUs inside the library (mps_create_sharded: contiguous
        site blocks, boundary sites exchanged by peer copies over NVLink); otherwise one engine on `device`."""
        self.L = abi.load_library()
        self.n = n_qubits
        self.nreg = n_registers
        self.h = C.c_void_p()
        if devices is not None and len(devices) > 1:
            if n_registers != 1:
                raise ValueError("a site-sharded handle holds one register")
            dv = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.L.mps_create_sharded(n_qubits, int(max_bond), float(svd_cutoff), int(gauge), len(devices), dv,
                                           int(partition_by == "cost"), int(seed), C.byref(self.h))
        else:
            if devices:
                device = devices[0]
            rc = self.L.mps_create(n_qubits, n_registers, int(max_bond), float(svd_cutoff), int(gauge), int(device), int(seed),
                                   C.byref(self.h))
        if rc != 0:
            msg = self.L.mps_last_error(None)
            self.h = None
            raise abi.MpsError("mps_create failed: %s" % (msg.decode() if msg else rc))
        for k, v in options.items():
            self.set_option(k, v)

    def _ck(self, rc):
        if rc != 0:
            msg = self.L.mps_last_error(self.h)
            raise abi.MpsError(msg.decode() if msg else "error %d" % rc)

    def close(self):
        if getattr(self, "h", None):
            self.L.mps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self._ck(self.L.mps_set_option(self.h, key.encode(), float(value)))

    def reset(self):
        self._ck(self.L.mps_reset(self.h))

    def snapshot(self):
        """Remember the current state on the device (VQE mode: the ansatz state, TNQVM.cpp:52-92)."""
        self._ck(self.L.mps_snapshot(self.h))

    def restore(self):
        self._ck(self.L.mps_restore(self.h))

    def clear_measure(self):
        self._ck(self.L.mps_clear_measure(self.h))

    # ---- gates
    def apply_1q(self, q, m):
        m = np.ascontiguousarray(m, dtype=np.complex128)
        self._ck(self.L.mps_apply_1q(self.h, q, m.ctypes.data))

    def apply_2q(self, q0, q1, m):
        m = np.ascontiguousarray(m, dtype=np.complex128)
        self._ck(self.L.mps_apply_2q(self.h, q0, q1, m.ctypes.data))

    def apply_layer(self, q0, q1, mats):
        q0 = np.ascontiguousarray(q0, dtype=np.int32)
        q1 = np.ascontiguousarray(q1, dtype=np.int32)
        mats = np.ascontiguousarray(mats, dtype=np.complex128)
        self._ck(self.L.mps_apply_layer(self.h, len(q0), q0.ctypes.data, q1.ctypes.data, mats.ctypes.data))

    def apply(self, name, qubits, params=()):
        """One XACC-named instruction, with the visitor's semantics (ExaTnMpsVisitor.cpp:870-1085)."""
        if name == "Measure":
            self._ck(self.L.mps_measure(self.h, qubits[0]))
            return
        if name == "I":
            return
        m = gate_matrix(name, params)
        if name == "Swap" and qubits[0] < qubits[1]:
            qubits = (qubits[1], qubits[0])   # visit(Swap) sorts the bits, :1030-1033
        if m.shape[0] == 2:
            self.apply_1q(qubits[0], m)
        else:
            self.apply_2q(qubits[0], qubits[1], m)

    def run(self, circuit, offset=0):
        if isinstance(circuit, CompiledCircuit):
            return self.run_compiled(circuit)
        for g in circuit:
            qs = tuple(q + offset for q in g[1]) if g[0] != "Measure" else g[1]
            self.apply(g[0], qs, g[2] if len(g) > 2 else ())
        return self

    def run_compiled(self, cc):
        """One ABI call for a whole instruction list (mps_apply_gates)."""
        if cc.count:
            self._ck(self.L.mps_apply_gates(self.h, cc.count, cc.q0.ctypes.data, cc.q1.ctypes.data, cc.mats.ctypes.data))
        for q in cc.measure:
            self._ck(self.L.mps_measure(self.h, q))
        return self

    def flush(self):
        self._ck(self.L.mps_flush(self.h))

    def sync(self):
        self._ck(self.L.mps_sync(self.h))

    # ---- observables
    def norm(self, reg=0):
        out = C.c_double()
        self._ck(self.L.mps_norm(self.h, reg, C.byref(out)))
        return out.value

    def expval_z(self, qubits, reg=0):
        q = np.ascontiguousarray(qubits, dtype=np.int32)
        out = C.c_double()
        self._ck(self.L.mps_expval_z(self.h, reg, len(q), q.ctypes.data, C.byref(out)))
        return out.value

    def expval_z_all(self, reg=0):
        out = np.zeros(self.n, dtype=np.float64)
        self._ck(self.L.mps_expval_z_all(self.h, reg, out.ctypes.data))
        return out

    def expval_zz_pairs(self, pairs, reg=0):
        qi = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int32)
        qj = np.ascontiguousarray([p[1] for p in pairs], dtype=np.int32)
        out = np.zeros(len(pairs), dtype=np.float64)
        self._ck(self.L.mps_expval_zz_pairs(self.h, reg, len(pairs), qi.ctypes.data, qj.ctypes.data, out.ctypes.data))
        return out

    def amplitude(self, bits, reg=0):
        b = np.ascontiguousarray(bits, dtype=np.int8)
        nopen = int((b < 0).sum())
        out = np.zeros(1 << nopen, dtype=np.complex128)
        ln = C.c_size_t()
        self._ck(self.L.mps_amplitude(self.h, reg, b.ctypes.data, out.ctypes.data, C.byref(ln)))
        return out if nopen else complex(out[0])

    def statevector(self, reg=0):
        out = np.zeros(1 << self.n, dtype=np.complex128)
        self._ck(self.L.mps_statevector(self.h, reg, out.ctypes.data))
        return out

    def measure(self, q):
        self._ck(self.L.mps_measure(self.h, q))

    def seed(self, s):
        self._ck(self.L.mps_seed(self.h, int(s)))

    def n_measured(self):
        out = C.c_int()
        self._ck(self.L.mps_n_measured(self.h, C.byref(out)))
        return out.value

    def sample(self, shots, reg=0):
        """(raw chars, strings produced); the stride of a string is n_measured() (the handle's own measure list, which
        grows with every Measure until reset() / clear_measure())."""
        cap = max(1, shots) * max(1, self.n_measured())
        buf = C.create_string_buffer(cap + 1)
        n_out = C.c_int()
        self._ck(self.L.mps_sample(self.h, reg, shots, buf, cap, C.byref(n_out)))
        return buf.raw, n_out.value

    def sample_strings(self, shots, n_measured, reg=0):
        raw, cnt = self.sample(shots, reg)
        return [raw[i * n_measured:(i + 1) * n_measured].decode() for i in range(cnt)]

    # ---- introspection
    def bond_dims(self):
        nt = self.n * self.nreg
        out = np.zeros(max(nt - 1, 1), dtype=np.int32)
        self._ck(self.L.mps_bond_dims(self.h, out.ctypes.data))
        return out[: nt - 1]

    def singular_values(self, bond):
        out = np.zeros(1 << 14, dtype=np.float64)
        cnt = C.c_int()
        self._ck(self.L.mps_singular_values(self.h, bond, out.ctypes.data, out.size, C.byref(cnt)))
        return out[: cnt.value].copy()

    def discarded_weight(self):
        out = C.c_double()
        self._ck(self.L.mps_discarded_weight(self.h, C.byref(out)))
        return out.value

    def fidelity_estimate(self):
        """prod over truncations of (1 - discarded/total weight)"""
        out = C.c_double()
        self._ck(self.L.mps_fidelity_estimate(self.h, C.byref(out)))
        return out.value

    def get_site(self, k):
        shp = np.zeros(3, dtype=np.int32)
        self._ck(self.L.mps_get_site(self.h, k, None, shp.ctypes.data))
        out = np.zeros(int(shp.prod()), dtype=np.complex128)
        self._ck(self.L.mps_get_site(self.h, k, out.ctypes.data, shp.ctypes.data))
        return out.reshape(tuple(int(x) for x in shp), order="F")

    def set_site(self, k, t):
        t = np.asfortranarray(t, dtype=np.complex128)
        assert t.ndim == 3 and t.shape[1] == 2
        self._ck(self.L.mps_set_site(self.h, k, t.ctypes.data, t.shape[0], t.shape[2]))

    def set_sites(self, tensors):
        """{site index: (dl, 2, dr) array}: all uploads in one call, one wait (mps_set_sites)."""
        ks = sorted(tensors)
        arrs = [np.asfortranarray(tensors[k], dtype=np.complex128) for k in ks]
        n = len(ks)
        kk = (C.c_int * n)(*ks)
        dl = (C.c_int * n)(*[a.shape[0] for a in arrs])
        dr = (C.c_int * n)(*[a.shape[2] for a in arrs])
        pp = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        self._ck(self.L.mps_set_sites(self.h, n, kk, pp, dl, dr))

    def get_sites(self, ks):
        """[site indices] -> list of (dl, 2, dr) arrays: all downloads in one call, one wait (mps_get_sites)."""
        ks = list(ks)
        shp = np.zeros(3, dtype=np.int32)
        outs = []
        for k in ks:
            self._ck(self.L.mps_get_site(self.h, k, None, shp.ctypes.data))
            outs.append(np.zeros(tuple(int(x) for x in shp), dtype=np.complex128, order="F"))
        n = len(ks)
        kk = (C.c_int * n)(*ks)
        pp = (C.c_void_p * n)(*[a.ctypes.data for a in outs])
        self._ck(self.L.mps_get_sites(self.h, n, kk, pp))
        return outs

    def site_device_ptr(self, k):
        shp = np.zeros(3, dtype=np.int32)
        p = C.c_void_p()
        self._ck(self.L.mps_site_device_ptr(self.h, k, C.byref(p), shp.ctypes.data))
        return p.value, tuple(int(x) for x in shp)

    def resize_site(self, k, dl, dr):
        p = C.c_void_p()
        self._ck(self.L.mps_resize_site(self.h, k, dl, dr, C.byref(p)))
        return p.value

    def stream(self):
        """cudaStream_t (as an int) the handle issues its work on."""
        p = C.c_void_p()
        self._ck(self.L.mps_get_stream(self.h, C.byref(p)))
        return p.value or 0

    def shard_layout(self):
        """First site of every device block (+ n at the end); [0, n] for a single-device handle."""
        nd = C.c_int()
        self._ck(self.L.mps_shard_layout(self.h, C.byref(nd), None))
        first = (C.c_int * (nd.value + 1))()
        self._ck(self.L.mps_shard_layout(self.h, C.byref(nd), first))
        return list(first)

    def stats(self):
        out = np.zeros(15, dtype=np.float64)
        self._ck(self.L.mps_stats(self.h, out.ctypes.data, 15))
        keys = ["gates_2q", "gates_1q_kernel", "layers", "jacobi_sweeps", "launches", "ms_theta", "ms_svd", "ms_writeback", "ms_qr",
                "jacobi_dmma_flops", "gates_2q_fused", "svd_nonconverged", "norm_guard_violations", "boundary_exchanges",
                "peer_bytes"]
        return dict(zip(keys, out.tolist()))
