"""The C++ host side above the C ABI: tnqvm::B200MpsVisitor driven like TNQVM::execute (tnqvm/TNQVM.cpp:106-139) by
b200_tnqvm_run.  CPU part: XASM reader + nearest-neighbour pass.  GPU part: the reference gtests' known answers
(tnqvm/visitors/exatn-mps/tests/MpsGateTester.cpp, MpsMeasurementTester.cpp, NumericalTesterCheckNorm.cpp) through
the visitor's own AcceleratorBuffer outputs ("norm", "exp-val-z", bit-string counts)."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

import reference_cases as RC
from oracle import oracle as O   # checker only
from tnqvm_b200 import circuits as Cc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "tnqvm_b200", "lib", "b200_tnqvm_run")


def run(circ, n, *args):
    p = subprocess.run([RUN, "--xasm", "-", "--qubits", str(n)] + [str(a) for a in args], input=Cc.to_xasm(circ), capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return p.stdout


def test_driver_is_built_and_links_the_cuda_library():
    assert os.path.exists(RUN)
    out = subprocess.run(["ldd", RUN], capture_output=True, text=True).stdout
    assert "libmps_b200.so" in out and "libtnqvm_b200_visitor.so" in out


def test_cpp_nearest_neighbor_pass_matches_host_restatement():
    # NearestNeighborTransform.hpp:43-135 restated twice (C++ pre-pass and tnqvm_b200.circuits): they must agree
    rng = np.random.default_rng(5)
    circ = []
    for _ in range(40):
        a, b = rng.choice(12, 2, replace=False)
        circ.append([("CNOT", (int(a), int(b)), ()), ("fSim", (int(a), int(b)), (0.3, -0.2)), ("Rz", (int(a),), (1.25,))][int(rng.integers(0, 3))])
    out = run(circ, 12, "--dump-nn").strip().splitlines()
    ref = Cc.nearest_neighbor(circ)
    assert len(out) == len(ref)
    for line, g in zip(out, ref):
        tok = line.split()
        assert tok[0] == g[0] and tuple(int(t) for t in tok[1:1 + len(g[1])]) == tuple(g[1])
        assert np.allclose([float(t) for t in tok[1 + len(g[1]):]], list(g[2]))
    syc = "/root/reference/examples/sycamore/resources/sycamore_53_14_0.xasm"
    if os.path.exists(syc):   # build container only
        p = subprocess.run([RUN, "--xasm", syc, "--dump-nn"], capture_output=True, text=True)
        lines = p.stdout.strip().splitlines()
        assert sum(1 for l in lines if len(l.split()) >= 3 and l.split()[0] in ("Swap", "fSim")) == 1897


def test_no_gpu_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([RUN, "--xasm", "-", "--qubits", "2"], input="H(q[0]);\n", capture_output=True, text=True)
    assert p.returncode != 0 and "no CUDA device" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", RC.PROB_CASES, ids=[c["name"] for c in RC.PROB_CASES])
def test_visitor_reference_gtests_sampled(case):
    # shots >= 1 -> appendMeasurement strings; the gtests assert probabilities within 0.05 / 0.01 / 0.1 on 10 000 shots
    res = json.loads(run(case["circuit"], case["n"], "--shots", 10000, "--seed", 123))
    tot = sum(res["counts"].values())
    assert tot == 10000
    for s, p in case["expect"].items():
        assert abs(res["counts"].get(s, 0) / tot - p) < 0.03, (case["cite"], res["counts"])
    assert abs(res["norm"] - 1.0) < 1e-12


@pytest.mark.gpu
def test_visitor_expval_z_deuteron_and_seed_determinism():
    for t, ref in zip(RC.deuteron_angles(), RC.DEUTERON_TABLE):
        res = json.loads(run(RC.deuteron_circuit(t), 2))     # shots < 1 -> "exp-val-z"  (MpsGateTester.cpp:359-407)
        assert abs(res["exp-val-z"] - ref) < 2e-6
    a = json.loads(run(RC.ghz4_measured(), 4, "--shots", 8192, "--seed", 123))["counts"]
    for _ in range(3):                                        # MpsMeasurementTester.cpp:37-66
        assert json.loads(run(RC.ghz4_measured(), 4, "--shots", 8192, "--seed", 123))["counts"] == a
    assert set(a) == {"0000", "1111"}


@pytest.mark.gpu
def test_visitor_large_register_sampling_and_norm():
    res = json.loads(run(RC.ghz35(), 35, "--shots", 20, "--seed", 9))   # MpsMeasurementTester.cpp:7-35
    assert set(res["counts"]) <= {"0000", "1111"} and sum(res["counts"].values()) == 20
    circ = [g for g in Cc.rcs(10, 15, seed=4)]
    for extra in ([], ["--svd-cutoff", 1e-16]):                          # NumericalTesterCheckNorm.cpp:15-62
        res = json.loads(run(circ, 10, "--shots", 8, *extra))
        assert abs(res["norm"] - 1.0) < 1e-6
    # amplitude output (computeWaveFuncSlice with a closed bit string, ExaTnMpsVisitor.cpp:2588-2675): GHZ 1/sqrt(2)
    res = json.loads(run([g for g in RC.ghz4_measured() if g[0] != "Measure"], 4, "--bitstring", "1111"))
    assert abs(res["amplitude"][0] - math.sqrt(0.5)) < 1e-12 and abs(res["amplitude"][1]) < 1e-12


@pytest.mark.gpu
def test_visitor_bitstring_option_amplitude_and_slice():
    """{"bitstring", vector<int>} as in ExaTnMpsVisitor.cpp:776-822: every leg fixed -> "amplitude-real"/"amplitude-imag";
    legs marked -1 -> the normalised slice in "amplitude-real-vec"/"amplitude-imag-vec" (ExatnVisitorTester.cpp:533-651 asserts
    the same GHZ values on the exatn visitor).  Also the b200-fuse-2q switch through the visitor options."""
    ghz = [g for g in RC.ghz4_measured() if g[0] != "Measure"]
    res = json.loads(run(ghz, 4, "--bitstring", "0000"))
    assert abs(res["amplitude"][0] - math.sqrt(0.5)) < 1e-12 and "amplitude_slice" not in res
    res = json.loads(run(ghz, 4, "--bitstring", "0110"))
    assert abs(res["amplitude"][0]) < 1e-12 and abs(res["amplitude"][1]) < 1e-12
    # q0 = q3 = 1 fixed, q1 and q2 open: only |11> of the open pair survives, normalised to 1 (open qubit q1 fastest)
    res = json.loads(run(ghz, 4, "--bitstring", "1xx1"))
    sl = np.array([complex(a, b) for a, b in res["amplitude_slice"]])
    assert sl.shape == (4,) and abs(abs(sl[3]) - 1.0) < 1e-12 and np.abs(sl[:3]).max() < 1e-12
    # all legs open on a product-free circuit: the slice is the normalised state vector
    circ = Cc.brickwork(5, 4, seed=2)
    res = json.loads(run(circ, 5, "--bitstring", "xxxxx", "--state", "--fuse-2q"))
    sl = np.array([complex(a, b) for a, b in res["amplitude_slice"]])
    sv = np.array([complex(a, b) for a, b in res["state"]])
    assert np.abs(sl - sv / np.linalg.norm(sv)).max() < 1e-12
    # wrong length -> the reference's error
    p = subprocess.run([RUN, "--xasm", "-", "--qubits", "4", "--bitstring", "01"], input=Cc.to_xasm(ghz), capture_output=True, text=True)
    assert p.returncode != 0 and "Bitstring size must match" in p.stderr


@pytest.mark.gpu
def test_visitor_vqe_mode_one_ansatz_many_terms():
    """TNQVM.cpp:52-92 with supportVqeMode(): the ansatz runs once, every observable term goes through
    getExpectationValueZ (change of basis + Measure) and must leave the ansatz state behind it
    (ITensorMPSVisitor.cpp:440-464 is the reference's implementation of that contract).  Each term must equal the
    "exp-val-z" of a full separate run of ansatz + term (VQEModeTester.cpp compares the two modes the same way)."""
    n = 6
    ansatz = Cc.hea(n, 2, seed=5) + [("CNOT", (0, 3), ()), ("Ry", (2,), (0.7,))]    # one long-range gate: routed by the pass
    terms = ["Z0", "X0X1", "Y2Y3", "Z1Z4", "X0Y2Z5", "Z0"]
    res = json.loads(run(ansatz, n, "--observe", ";".join(terms), "--state"))
    assert len(res["vqe_terms"]) == len(terms)
    assert abs(res["vqe_terms"][0] - res["vqe_terms"][-1]) < 1e-13                # the ansatz state came back
    basis = {"X": lambda q: [("H", (q,), ())], "Y": lambda q: [("Rx", (q,), (math.pi / 2,))], "Z": lambda q: []}
    import re
    for t, val in zip(terms, res["vqe_terms"]):
        ops = [(m.group(1), int(m.group(2))) for m in re.finditer(r"([XYZ])(\d+)", t)]
        circ = list(ansatz)
        for p, q in ops:
            circ += basis[p](q)
        circ += [("Measure", (q,), ()) for _, q in ops]
        one = json.loads(run(circ, n))
        assert abs(one["exp-val-z"] - val) < 1e-10, t
        # and the reference's own dense simulator (oracle/_ref: Gates.hpp + GateMatrixAlgebra.hpp) on the same term
        dense = O.dense_run(n, [g for g in circ if g[0] != "Measure"])
        assert abs(O.dense_expval_z(dense, n, [q for _, q in ops]) - val) < 1e-10, t
    base = json.loads(run(ansatz, n, "--state"))
    sv = np.array([complex(a, b) for a, b in res["state"]])
    sv0 = np.array([complex(a, b) for a, b in base["state"]])
    assert np.abs(sv - sv0).max() < 1e-12


@pytest.mark.gpu
def test_visitor_execution_info_has_the_reference_stat_buckets():
    """getExecutionInfo() carries the reference's per-phase statistics under the reference's names (FunctionCallStat buckets,
    ExatnUtils.hpp:57-126; ExaTnMpsVisitor.cpp:345, 680, 1292, 1523, 1626, 1718, 1729) plus the engine counters; with
    "b200-profile" the three GPU phases of the two-qubit step are timed with CUDA events."""
    n = 12
    circ = Cc.brickwork(n, 6, seed=2)
    doc = json.loads(run(circ, n, "--max-bond-dim", 16, "--profile").strip().splitlines()[-1])
    info = doc["execution_info"]
    n1, n2 = Cc.count_gates(circ)
    assert info["Initialize [calls]"] == 1 and info["Finalize [calls]"] == 1
    assert info["One-qubit Gate Total [calls]"] == n1 and info["Two-qubit Gate Total [calls]"] == n2
    for key in ("Contract Two-Qubit Gate Tensor [secs]", "Decompose Tensor SVD [secs]", "Truncate SVD Tensor [secs]"):
        assert info[key] > 0.0, key
    assert info["Two-qubit Gate Total [gpu gates]"] == n2 and info["b200-kernel-launches"] > 0
    assert info["b200-svd-nonconverged"] == 0


@pytest.mark.gpu
def test_visitor_site_sharded_option_matches_single_device():
    """{"b200-devices", [..]} through the C++ visitor: the sharded run (two site blocks; both on GPU 0 when the box has one GPU)
    gives the buffer outputs of the single-device run."""
    import torch
    n = 14
    circ = Cc.brickwork(n, 7, seed=4, prefix_ghz=True) + [("Measure", (q,), ()) for q in (0, 5, 13)]
    devs = "0,1" if torch.cuda.device_count() > 1 else "0,0"
    a = json.loads(run(circ, n, "--max-bond-dim", 16).strip().splitlines()[-1])
    b = json.loads(run(circ, n, "--max-bond-dim", 16, "--devices", devs).strip().splitlines()[-1])
    assert abs(a["norm"] - b["norm"]) < 1e-12 and abs(a["exp-val-z"] - b["exp-val-z"]) < 1e-12
    assert a["bond_dims"] == b["bond_dims"]
