"""Host logic of the site-sharded handle (mps_create_sharded), checked WITHOUT a GPU: the schedule a flush builds for a gate list
(mps_shard_plan_debug) is interpreted by a small simulator with one in-order executor per device.  Replaces, for the C-ABI path,
what tests/test_sharded_host.py does for the torch.distributed path with gloo.  Reference scheme being replaced:
ExaTnMpsVisitor.cpp:347-531 (site blocks), :2059-2170 (boundary gate dispatch)."""
import ctypes

import numpy as np
import pytest

from tnqvm_b200 import abi, circuits as Cc

LAYER, SEND, RECV = 0, 1, 2


def plan(n, world, circ, chi=0, by_cost=0):
    L = abi.load_library()
    first = (ctypes.c_int * (world + 1))()
    assert L.mps_shard_partition(n, world, chi, by_cost, first) == 0
    gates = [g for g in circ if g[0] not in ("Measure", "I")]
    q0 = np.array([g[1][0] for g in gates], dtype=np.int32)
    q1 = np.array([g[1][1] if len(g[1]) == 2 else -1 for g in gates], dtype=np.int32)
    cap = 16 * len(gates) + 64
    out = np.zeros(cap, dtype=np.int32)
    used = ctypes.c_int()
    assert L.mps_shard_plan_debug(n, world, first, len(gates), q0.ctypes.data, q1.ctypes.data, out.ctypes.data, cap, ctypes.byref(used)) == 0
    ops = [[] for _ in range(world)]
    i = 0
    while i < used.value:
        d, kind, site, slot, ng = (int(x) for x in out[i:i + 5])
        ops[d].append((kind, site, slot, [int(x) for x in out[i + 5:i + 5 + ng]]))
        i += 5 + ng
    owner = np.zeros(n, dtype=int)
    for d in range(world):
        owner[first[d]:first[d + 1]] = d
    return gates, ops, owner


def simulate(n, world, gates, ops, owner):
    """Runs the op lists with one in-order executor per device; a RECV blocks until its slot was published.  Checks: no deadlock;
    every gate finds its sites on the device that runs it; per site the gates run in program order; a site is never touched
    by its owner while it is lent out; every boundary exchange moves exactly one site each way; all sites are home at the end."""
    where = list(owner)            # device currently holding site k
    published = {}                 # slot -> (site, from device)
    pc = [0] * world
    done_gates = []
    next_on_site = [0] * n         # per site: position in its own program-order gate list
    per_site = [[] for _ in range(n)]
    for gi, g in enumerate(gates):
        for q in g[1]:
            per_site[q].append(gi)
    moves = 0
    progress = True
    while progress:
        progress = False
        for d in range(world):
            while pc[d] < len(ops[d]):
                kind, site, slot, gl = ops[d][pc[d]]
                if kind == RECV:
                    if slot not in published:
                        break                                   # blocked: try the other devices
                    s_site, src = published.pop(slot)
                    assert s_site == site and where[site] == -1
                    where[site] = d
                    moves += 1
                elif kind == SEND:
                    assert where[site] == d, ("sending a site the device does not hold", d, site)
                    where[site] = -1                              # in flight
                    published[slot] = (site, d)
                else:
                    sites_in_layer = set()
                    for gi in gl:
                        qs = gates[gi][1]
                        assert owner[min(qs)] == d                # a gate runs on the owner of its left site
                        for q in qs:
                            assert where[q] == d, ("gate on a site that is not on this device", d, gi, q, where[q])
                            assert q not in sites_in_layer, "two gates of one layer share a site"
                            sites_in_layer.add(q)
                            assert per_site[q][next_on_site[q]] == gi, ("program order violated on site", q)
                            next_on_site[q] += 1
                        done_gates.append(gi)
                pc[d] += 1
                progress = True
    assert all(pc[d] == len(ops[d]) for d in range(world)), "deadlock: " + str([(pc[d], len(ops[d])) for d in range(world)])
    assert sorted(done_gates) == list(range(len(gates))) and not published
    assert where == list(owner), "a site did not come home"
    return moves


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("kind", ["brickwork", "qaoa_ring", "sycamore", "random_1q_2q"])
def test_shard_schedule_is_deadlock_free_and_keeps_program_order(world, kind):
    rng = np.random.default_rng(world)
    if kind == "brickwork":
        n, circ, chi = 50, Cc.brickwork(50, 6, seed=3), 256
    elif kind == "qaoa_ring":
        n, circ, chi = 40, Cc.nearest_neighbor(Cc.qaoa_ring(40, 2, seed=7)), 64
    elif kind == "sycamore":
        n, raw = Cc.sycamore_53(14)
        circ, chi = Cc.nearest_neighbor(raw)[:1500], 1024
    else:
        n, chi, circ = 17, 0, []
        for _ in range(400):
            if rng.integers(0, 3):
                a = int(rng.integers(0, n - 1))
                circ.append(("CNOT", (a, a + 1) if rng.integers(0, 2) else (a + 1, a), ()))
            else:
                circ.append(("H", (int(rng.integers(0, n)),), ()))
    for by_cost in (0, 1):
        gates, ops, owner = plan(n, world, circ, chi, by_cost)
        moves = simulate(n, world, gates, ops, owner)
        crossing = sum(1 for g in gates if len(g[1]) == 2 and owner[g[1][0]] != owner[g[1][1]])
        assert moves == 2 * crossing                               # one site each way per boundary gate, nothing else moves
        # a device's layers contain only its own gates and every gate appears exactly once
        assert sum(len(gl) for d in range(world) for k, _, _, gl in ops[d] if k == LAYER) == len(gates)


def test_shard_schedule_single_device_has_no_transfers():
    gates, ops, owner = plan(12, 1, Cc.brickwork(12, 4, seed=1))
    assert all(k == LAYER for k, _, _, _ in ops[0]) and simulate(12, 1, gates, ops, owner) == 0
