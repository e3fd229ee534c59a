"""Host logic of the multi-GPU drivers (tnqvm_b200/sharded.py) on CPU: world_size-2/3 gloo process groups, with the
local engine replaced by the CPU oracle (tests may use oracle/; the product never does).  What is checked is the
partition, ownership, boundary-exchange schedule and the gather -- the same code that drives NCCL on the GPUs."""
import contextlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tnqvm_b200 import circuits as Cc          # noqa: E402
from tnqvm_b200 import sharded as S            # noqa: E402


class OracleLocal:
    """Stand-in for sharded.B200Local with the same protocol, backed by the CPU oracle."""

    def __init__(self, n_sites, max_bond=0, svd_cutoff=-1.0, gauge=0, **_):
        from oracle import oracle as O
        self.o = O.OracleMPS(n_sites, max_bond=max_bond, svd_cutoff=svd_cutoff, gauge=gauge)
        self.comm_device = torch.device("cpu")
        self.pending = {}

    def apply(self, name, qubits, params=()):
        self.o.apply(name, qubits, params)

    def flush(self):
        pass

    def export_site(self, k):
        t = self.o.get_site(k)
        flat = np.ascontiguousarray(t.ravel(order="F")).view(np.float64)
        return torch.from_numpy(flat.copy()), t.shape[0], t.shape[2]

    def import_site(self, k, dl, dr):
        buf = torch.empty(4 * dl * dr, dtype=torch.float64)
        self.pending[k] = (buf, dl, dr)
        return buf

    def commit_site(self, k):
        buf, dl, dr = self.pending.pop(k)
        self.o.set_site(k, buf.numpy().view(np.complex128).reshape((dl, 2, dr), order="F"))

    def comm_context(self):
        return contextlib.nullcontext()

    def close(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, circ, max_bond, q, partition_by="count"):
    try:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        sh = S.ShardedMPS(n, max_bond=max_bond, local_factory=lambda ns, **kw: OracleLocal(ns, **kw), partition_by=partition_by)
        sh.run(circ)
        full = sh.gather_to_root()
        if rank == 0:
            o = full.o
            q.put(("ok", [o.expval_z([k]) for k in range(n)], o.norm(), o.bond_dims().tolist(), sh.exchanges))
        else:
            q.put(("peer", sh.exchanges, sh.bytes_sent))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # pragma: no cover
        import traceback
        q.put(("err", "rank %d: %s\n%s" % (rank, e, traceback.format_exc())))


def _run_sharded(world, n, circ, max_bond, partition_by="count"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, circ, max_bond, q, partition_by)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    errs = [r for r in res if r[0] == "err"]
    assert not errs, errs[0][1]
    return [r for r in res if r[0] == "ok"][0], [r for r in res if r[0] == "peer"]


def test_partition_and_ownership():
    assert S.partition(10, 2) == [(0, 5), (5, 10)]
    assert S.partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert S.partition(53, 8)[-1][1] == 53 and all(e - s in (6, 7) for s, e in S.partition(53, 8))
    b = S.partition(100, 8)
    assert [S.owner_of(q, b) for q in (0, 12, 13, 99)] == [0, 0, 1, 7]
    with pytest.raises(ValueError):
        S.partition(3, 4)
    assert S.shard_items(64, 3, 8) == list(range(24, 32))
    assert sorted(sum((S.shard_items(10, r, 4) for r in range(4)), [])) == list(range(10))


@pytest.mark.parametrize("world,n,depth", [(2, 10, 6), (3, 11, 5), (2, 2, 3)])
def test_site_sharded_equals_single_process(world, n, depth):
    from oracle import oracle as O
    circ = Cc.brickwork(n, depth, seed=21 + n, prefix_ghz=True)
    ok, peers = _run_sharded(world, n, circ, 0)
    ref = O.OracleMPS(n).run(circ)
    z = np.array(ok[1])
    zr = np.array([ref.expval_z([k]) for k in range(n)])
    assert np.max(np.abs(z - zr)) < 1e-12
    assert abs(ok[2] - ref.norm()) < 1e-12
    assert ok[3] == ref.bond_dims().tolist()
    # every boundary gate is exactly one site tensor each way
    bounds = S.partition(n, world)
    edges = {e for (_, e) in bounds[:-1]}
    nb = sum(1 for g in circ if len(g[1]) == 2 and max(g[1]) in edges and min(g[1]) == max(g[1]) - 1)
    total_sends = ok[4] + sum(p[1] for p in peers)
    gather_sends = n - (bounds[0][1] - bounds[0][0])
    assert total_sends == 2 * nb + gather_sends


def test_site_sharded_with_truncation_and_routed_gates():
    from oracle import oracle as O
    n = 9
    circ = Cc.nearest_neighbor(Cc.qaoa_ring(n, 2, seed=5))
    ok, _ = _run_sharded(2, n, circ, 8)
    ref = O.OracleMPS(n, max_bond=8).run(circ)
    z = np.array(ok[1])
    zr = np.array([ref.expval_z([k]) for k in range(n)])
    assert np.max(np.abs(z - zr)) < 1e-10
    assert ok[3] == ref.bond_dims().tolist()


def _sweep_worker(rank, world, port, n, circs, q):
    try:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        from oracle import oracle as O

        class MultiReg:
            """n_registers independent oracle chains behind the B200MPS multi-register surface."""

            def __init__(self, nq, nreg, max_bond=0, **_):
                self.nq, self.regs = nq, [O.OracleMPS(nq, max_bond=max_bond) for _ in range(nreg)]

            def run(self, circ, offset=0):
                r = offset // self.nq
                for g in circ:
                    self.regs[r].apply(g[0], g[1], g[2] if len(g) > 2 else ())

            def expval_z_all(self, reg=0):
                return np.array([self.regs[reg].expval_z([k]) for k in range(self.nq)])

            def close(self):
                pass

        out = S.run_parameter_sweep(n, circs, engine_factory=lambda nq, nreg, **kw: MultiReg(nq, nreg, **kw))
        q.put(("ok" if rank == 0 else "peer", out))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # pragma: no cover
        import traceback
        q.put(("err", "rank %d: %s\n%s" % (rank, e, traceback.format_exc())))


def test_parameter_sweep_sharded_across_ranks():
    from oracle import oracle as O
    n, world = 6, 2
    circs = [Cc.hea(n, 2, seed=s) for s in range(5)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_sweep_worker, args=(r, world, port, n, circs, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    errs = [r for r in res if r[0] == "err"]
    assert not errs, errs[0][1]
    for _, out in res:
        for i, c in enumerate(circs):
            ref = O.OracleMPS(n).run(c)
            assert np.allclose(out[i], [ref.expval_z([k]) for k in range(n)], atol=1e-12)


def test_cost_balanced_partition():
    """balance the blocks by the SVD work of the saturated bond profile, not by site count."""
    def cost(n, chi, bounds):
        dims = [1] + [min(chi, 2 ** min(k + 1, n - 1 - k)) for k in range(n - 1)] + [1]
        w = [0.0 if k == n - 1 else (2.0 * dims[k]) * (2.0 * dims[k + 2]) * min(2.0 * dims[k], 2.0 * dims[k + 2]) for k in range(n)]
        return [sum(w[s:e]) for s, e in bounds]
    for n, world, chi in ((50, 4, 256), (50, 8, 256), (100, 8, 512), (53, 8, 1024), (10, 4, 4), (5, 4, 2), (4, 4, 64)):
        b = S.partition(n, world, chi)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(e > s for s, e in b)
        if n >= 50:
            even = S.partition(n, world)
            assert max(cost(n, chi, b)) < 0.9 * max(cost(n, chi, even))      # the heaviest block got lighter
            assert b[0][1] - b[0][0] > even[0][1] - even[0][0]               # because the cheap chain ends got more sites
    assert S.partition(50, 4, 0) == S.partition(50, 4) and S.partition(50, 1, 256) == [(0, 50)]


def test_site_sharded_cost_partition_equals_single_process():
    from oracle import oracle as O
    n, world, chi = 11, 3, 4
    assert S.partition(n, world, chi) != S.partition(n, world)
    circ = Cc.brickwork(n, 6, seed=3, prefix_ghz=True)
    ok, _ = _run_sharded(world, n, circ, chi, partition_by="cost")
    ref = O.OracleMPS(n, max_bond=chi).run(circ)
    assert np.max(np.abs(np.array(ok[1]) - np.array([ref.expval_z([k]) for k in range(n)]))) < 1e-9
    assert abs(ok[2] - ref.norm()) < 1e-9 and ok[3] == ref.bond_dims().tolist()
