"""Pins the CPU oracle (oracle/mps_oracle.cpp) BEFORE it is trusted as the checker:
  * every known-answer test the reference holds for the exatn-mps path (tests/reference_cases.py),
  * the reference's own header-only dense simulator + sampler compiled into oracle/_ref (when present),
  * the committed golden fixtures tests/golden/*.json (generated from oracle/_ref by tests/golden/make_golden.py)."""
import json
import math
import os

import numpy as np
import pytest

import reference_cases as RC
from oracle import oracle as O
from tnqvm_b200 import circuits as Cc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_oracle(n, circ, **kw):
    circ = Cc.nearest_neighbor(circ)   # TNQVM.cpp:119-124
    return O.OracleMPS(n, **kw).run(circ)


@pytest.mark.parametrize("case", RC.PROB_CASES, ids=[c["name"] for c in RC.PROB_CASES])
def test_reference_probability_cases(case):
    o = run_oracle(case["n"], case["circuit"])
    measured = [g[1][0] for g in case["circuit"] if g[0] == "Measure"]
    probs = RC.probs_from_state(o.statevector(), case["n"], measured)
    for s, p in case["expect"].items():
        assert abs(probs.get(s, 0.0) - p) < 1e-12, (case["cite"], s, probs)
    assert abs(o.norm() - 1.0) < 1e-12


def test_deuteron_table():
    # MpsGateTester.cpp:359-407, table to 6 digits
    for t, ref in zip(RC.deuteron_angles(), RC.DEUTERON_TABLE):
        o = run_oracle(2, RC.deuteron_circuit(t))
        assert abs(o.expval_z([0, 1]) - ref) < 2e-6


def test_grover():
    o = run_oracle(3, RC.grover_circuit())
    probs = RC.probs_from_state(o.statevector(), 3, [2, 1, 0])
    assert probs.get("110", 0.0) > 0.5   # MpsGateTester.cpp:409-499


def test_rx_expectation_law():
    for th in np.linspace(-math.pi, math.pi, 7):
        o = run_oracle(3, [("Rx", (1,), (float(th),))])
        assert abs(o.expval_z([1]) - RC.rx_expz(th)) < 1e-12   # ITensorMPSVisitorTester.cpp:322-332


def test_ghz35_rdm_sampling_and_seed():
    # MpsMeasurementTester.cpp:7-35 (n >= 20 branch) and :37-66 (seed determinism)
    o = run_oracle(35, RC.ghz35(), seed=5)
    s = o.sample(6, 4)
    assert set(s) <= {"0000", "1111"}
    runs = []
    for _ in range(3):
        o4 = run_oracle(4, RC.ghz4_measured(), seed=123)
        s4 = o4.sample(8192, 4)
        runs.append({k: s4.count(k) for k in set(s4)})
    assert runs[0] == runs[1] == runs[2] and set(runs[0]) == {"0000", "1111"}


def test_norm_of_rcs_circuit():
    # NumericalTesterCheckNorm.cpp:15-62: 10-qubit rcs, 15 layers, norm = 1 +- 1e-6 with and without svd-cutoff 1e-16
    c = Cc.rcs(10, 15, seed=4)
    assert abs(run_oracle(10, c).norm() - 1.0) < 1e-6
    assert abs(run_oracle(10, c, svd_cutoff=1e-16).norm() - 1.0) < 1e-6


needs_ref = pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (reference tree absent)")


@needs_ref
def test_gate_matrices_equal_reference_headers():
    for nm, pr in [("H", ()), ("X", ()), ("Y", ()), ("Z", ()), ("T", ()), ("Tdg", ()), ("Rx", (0.3,)), ("Ry", (0.7,)), ("Rz", (-1.1,)),
                   ("U", (0.3, 0.4, 0.5)), ("CNOT", ()), ("CZ", ()), ("CY", ()), ("CH", ()), ("CRZ", (0.9,)), ("CPhase", (0.2,)),
                   ("Swap", ()), ("iSwap", ()), ("fSim", (0.4, 0.6)), ("I", ())]:
        assert np.array_equal(O.gate_matrix(nm, pr), O.ref_gate_matrix(nm, pr)), nm


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_vs_reference_dense_simulator(seed):
    n = 9
    c = Cc.brickwork(n, 7, seed=seed, prefix_ghz=(seed == 1))
    o = run_oracle(n, c)
    d = O.dense_run(n, c)   # reference ApplySingleQubitGate / ApplyCNOTGate
    assert np.abs(o.statevector() - d).max() < 1e-13
    for q in range(n):
        assert abs(o.expval_z([q]) - O.dense_expval_z(d, n, [q])) < 1e-13


@needs_ref
def test_sampler_equals_reference_generate_samples():
    import ctypes
    n = 6
    c = Cc.brickwork(n, 5, seed=8)
    o = run_oracle(n, c, seed=77)
    for q in (3, 0, 5):
        o.measure(q)
    mine = o.sample(500, 3)
    R = O.ref()
    R.ref_set_seed(77)
    sv = o.statevector()
    bits = np.array([3, 0, 5], dtype=np.int32)
    buf = ctypes.create_string_buffer(500 * 3 + 1)
    m = R.ref_generate_samples(sv.ctypes.data, n, 500, bits.ctypes.data, 3, buf)
    theirs = [buf.raw[i * 3:(i + 1) * 3].decode() for i in range(m)]
    assert mine == theirs


def test_golden_fixtures():
    files = sorted(f for f in os.listdir(GOLD) if f.endswith(".json"))
    assert files, "golden fixtures missing"
    for f in files:
        gold = json.load(open(os.path.join(GOLD, f)))
        circ = [(g[0], tuple(g[1]), tuple(g[2])) for g in gold["circuit"]]
        o = run_oracle(gold["n"], circ)
        z = np.array([o.expval_z([q]) for q in range(gold["n"])])
        assert np.abs(z - np.array(gold["expz"])).max() < 1e-12, f
        for bits, (re, im) in gold["amplitudes"]:
            assert abs(o.amplitude(bits) - complex(re, im)) < 1e-12, f
        assert abs(o.norm() - gold["norm"]) < 1e-12


def test_oracle_null_rule_option_is_harmless_on_exact_runs_and_only_removes_noise_slices():
    """OracleMPS(null_tol=...) mirrors the engine's documented deviation (numerically-null singular values count as exact zeros,
    DESIGN.md section 1).  It is NOT part of the reference; this pins what it does: on exact (untruncated) runs the state is the
    reference's to 1e-12 and only weightless bond slices disappear; on a truncated run of a rank-deficient circuit (GHZ prefix)
    it never keeps more than the reference's rule does."""
    n = 10
    c = Cc.brickwork(n, 6, seed=4, prefix_ghz=True)
    a = run_oracle(n, c)
    b = O.OracleMPS(n, null_tol=1e-13).run(c)
    assert np.abs(a.statevector() - b.statevector()).max() < 1e-12
    assert (np.asarray(b.bond_dims()) <= np.asarray(a.bond_dims())).all() and (np.asarray(b.bond_dims()) < np.asarray(a.bond_dims())).any()
    for k in range(n - 1):
        sa, sb = a.singular_values(k), b.singular_values(k)
        m = min(len(sa), len(sb))
        # what was dropped carried no weight; the 1e-9-level differences are the reference rule's own amplified noise (sqrt of a
        # 1e-17 singular value lands on the neighbouring sites)
        assert np.abs(sa[:m] - sb[:m]).max() < 2e-8 * sa[0] and (sa[m:] < 1e-8 * sa[0]).all()
    at = O.OracleMPS(n, max_bond=8).run(c)
    bt = O.OracleMPS(n, max_bond=8, null_tol=1e-13).run(c)
    assert (np.asarray(bt.bond_dims()) <= np.asarray(at.bond_dims())).all()
    assert 0.0 <= bt.fidelity_estimate() <= 1.0 and 0.0 <= at.fidelity_estimate() <= 1.0
