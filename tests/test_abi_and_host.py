"""CPU-side checks: the C-ABI library loads and exports every symbol include/mps_b200.h declares (no compute
without a GPU), the product refuses to run without CUDA (no fallback), and the host-side circuit logic."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import tnqvm_b200
from tnqvm_b200 import abi, circuits as Cc, gates

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mps_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mps_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 25
    L = ctypes.CDLL(abi.lib_path())
    for s in syms:
        assert hasattr(L, s), "libmps_b200.so does not export %s" % s
    assert set(syms) == set(abi.SYMBOLS), set(syms) ^ set(abi.SYMBOLS)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(abi.MpsError, match="no CUDA device|CUDA"):
        tnqvm_b200.B200MPS(4)


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = "import sys; sys.path.insert(0, %r); import tnqvm_b200; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'" % ROOT
    subprocess.check_call([sys.executable, "-c", code])
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tnqvm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, os.path.join(dirpath, f)


def test_gate_matrices_unitary_and_conventions():
    for nm, pr in [("H", ()), ("Rx", (0.3,)), ("U", (0.1, 0.2, 0.3)), ("CNOT", ()), ("fSim", (0.4, 0.6)), ("iSwap", ()), ("CPhase", (0.7,))]:
        m = gates.gate_matrix(nm, pr)
        assert np.allclose(m @ m.conj().T, np.eye(m.shape[0]), atol=1e-14)
    assert gates.gate_matrix("CNOT")[3, 2] == 1 and gates.gate_matrix("fSim", (0.0, 0.5))[3, 3] == np.exp(-0.5j)
    assert np.array_equal(gates.gate_matrix("Bogus"), np.eye(2))   # ExatnUtils.cpp:112


def test_xasm_roundtrip_and_counts():
    c = Cc.brickwork(6, 3, seed=1, prefix_ghz=True) + [("fSim", (1, 2), (0.25, -0.5)), ("Measure", (3,), ())]
    n, c2 = Cc.load_xasm(Cc.to_xasm(c))
    assert n == 6 and len(c2) == len(c)
    for a, b in zip(c, c2):
        assert a[0] == b[0] and tuple(a[1]) == tuple(b[1]) and np.allclose(a[2], b[2])


def test_nearest_neighbor_matches_reference_swap_bookkeeping():
    # NearestNeighborTransformTester.cpp: all 2q distances become 1; CNOT(0,3) -> swaps (0,1) then (3,2)... back again
    out = Cc.nearest_neighbor([("CNOT", (0, 3), ())])
    assert all(abs(g[1][0] - g[1][1]) == 1 for g in out if len(g[1]) == 2)
    assert [g[0] for g in out] == ["Swap", "Swap", "CNOT", "Swap", "Swap"]
    assert out[0][1] == (0, 1) and out[1][1] == (3, 2) and out[2][1] == (1, 2)
    out = Cc.nearest_neighbor([("CNOT", (3, 1), ())], max_distance=2)
    assert out == [("CNOT", (3, 1), ())]
    # swap count: distance d needs 2*(d-1) swaps
    for d in range(2, 9):
        o = Cc.nearest_neighbor([("CZ", (0, d), ())])
        assert sum(1 for g in o if g[0] == "Swap") == 2 * (d - 1)


def test_config_generators_shapes():
    c2 = Cc.brickwork(50, 20, seed=12345)
    assert Cc.count_gates(c2) == (1000, 490)
    q = Cc.nearest_neighbor(Cc.qaoa_ring(100, 4))
    assert all(abs(g[1][0] - g[1][1]) == 1 for g in q if len(g[1]) == 2)
    assert Cc.count_gates(Cc.hea(64, 4, seed=0)) == (512, 252)
    syc = "/root/reference/examples/sycamore/resources/sycamore_53_14_0.xasm"
    if os.path.exists(syc):
        n, c = Cc.load_xasm(open(syc).read())
        assert n == 53 and Cc.count_gates(c) == (2527, 301)
        assert Cc.count_gates(Cc.nearest_neighbor(c))[1] == 1897


def test_sycamore_grid_stand_in_has_the_shape_of_the_resource_file():
    """C5 on the GPU box (no /root/reference there): 53 qubits, 14 cycles, ~20-27 fSim per cycle in the ABCDCDAB pattern
    order, per-coupler angles near (pi/2, pi/6), long-range couplers that the nearest-neighbour pass routes."""
    import math
    c = Cc.sycamore_grid()
    assert max(max(g[1]) for g in c) == 52
    fs = [g for g in c if g[0] == "fSim"]
    assert len(fs) == 320 and {abs(g[1][0] - g[1][1]) for g in fs} == {1, 6}
    assert all(abs(g[2][0] - math.pi / 2) < 0.07 and abs(g[2][1] - math.pi / 6) < 0.07 for g in fs)
    ang = {}
    for g in fs:
        assert ang.setdefault(g[1], g[2]) == g[2]          # one angle pair per coupler, every time it fires
    nn = Cc.nearest_neighbor(c)
    assert Cc.count_gates(nn) == (1240, 2200)
    assert all(abs(g[1][0] - g[1][1]) == 1 for g in nn if len(g[1]) == 2)
    assert Cc.sycamore_grid(seed=0) == c and Cc.sycamore_grid(seed=1) != c


def test_plugin_activator_registers_the_visitor_service():
    """tnqvm_b200/csrc/visitor/plugin/: the CppMicroServices activator (pattern of ExaTnMpsActivator.cpp:14-19), compiled against the
    in-tree shim by build(), registers exactly one tnqvm::TNQVMVisitor named "exatn-mps"; manifest.json names the bundle; the
    plugin CMakeLists.txt configures (XACC's CMake functions and targets stubbed)."""
    import json
    import shutil
    import subprocess
    import tempfile
    plug = os.path.join(ROOT, "tnqvm_b200", "csrc", "visitor", "plugin")
    exe = os.path.join(ROOT, "tnqvm_b200", "lib", "b200_activator_check")
    assert os.path.exists(exe), "run build() first"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "name=exatn-mps" in r.stdout, r.stdout + r.stderr
    man = json.load(open(os.path.join(plug, "manifest.json")))
    assert man["bundle.symbolic_name"] == "tnqvm_b200_mps" and man["bundle.activator"] is True
    if shutil.which("cmake"):
        with tempfile.TemporaryDirectory() as tmp:
            c = subprocess.run(["cmake", "-S", os.path.join(plug, "cmake_check"), "-B", tmp, "-DMPS_B200_ROOT=" + ROOT], capture_output=True, text=True, timeout=300)
            assert c.returncode == 0, c.stdout[-1500:] + c.stderr[-1500:]


def test_shard_partition_formula_matches_the_host_restatement():
    """mps_shard_partition (the site-block formula of mps_create_sharded, callable without a device) against
    tnqvm_b200.sharded.partition: equal counts and equal estimated SVD cost; contiguous, non-empty blocks covering all sites."""
    from tnqvm_b200 import abi
    L = abi.load_library()
    try:
        from tnqvm_b200.sharded import partition
    except Exception:
        partition = None
    for n in (2, 5, 16, 50, 53, 100):
        for world in (1, 2, 3, 4, 8):
            if world > n:
                continue
            for chi, by_cost in ((0, 0), (64, 1), (256, 1), (1024, 1)):
                out = (ctypes.c_int * (world + 1))()
                assert L.mps_shard_partition(n, world, chi, by_cost, out) == 0
                first = list(out)
                assert first[0] == 0 and first[-1] == n and all(a < b for a, b in zip(first, first[1:])), (n, world, chi, first)
                if partition is not None:
                    ref = partition(n, world, chi if by_cost else 0)
                    assert first == [s for s, _ in ref] + [n], (n, world, chi, first, ref)
    out = (ctypes.c_int * 4)()
    assert L.mps_shard_partition(2, 3, 0, 0, out) != 0   # fewer sites than devices


def test_python_gate_table_equals_the_reference_table_and_knows_s_sdg_u3():
    """tnqvm_b200.gates.gate_matrix (the Python host path) against the oracle's table, which is pinned bit for bit to the
    reference's Gates.hpp (tests/test_oracle_golden.py); S, Sdg and the U3 alias -- gates the reference header names but gives no
    matrix for, and which the C++ B200MpsVisitor applies natively -- have their textbook matrices on both host surfaces."""
    import cmath
    from oracle import oracle as O   # checker only
    from tnqvm_b200.gates import gate_matrix
    for nm, pr in [("H", ()), ("X", ()), ("Y", ()), ("Z", ()), ("T", ()), ("Tdg", ()), ("Rx", (0.3,)), ("Ry", (0.7,)), ("Rz", (-1.1,)),
                   ("U", (0.3, 0.4, 0.5)), ("CNOT", ()), ("CZ", ()), ("CY", ()), ("CH", ()), ("CRZ", (0.9,)), ("CPhase", (0.2,)),
                   ("Swap", ()), ("iSwap", ()), ("fSim", (0.4, 0.6)), ("I", ())]:
        assert np.array_equal(np.asarray(gate_matrix(nm, pr)), np.asarray(O.gate_matrix(nm, pr))), nm
    assert np.allclose(gate_matrix("S", ()), [[1, 0], [0, 1j]]) and np.allclose(gate_matrix("Sdg", ()), [[1, 0], [0, -1j]])
    assert np.array_equal(np.asarray(gate_matrix("U3", (0.3, 0.4, 0.5))), np.asarray(gate_matrix("U", (0.3, 0.4, 0.5))))
    assert np.allclose(np.asarray(gate_matrix("S", ())) @ np.asarray(gate_matrix("S", ())), np.asarray(gate_matrix("Z", ())))
    assert abs(np.asarray(gate_matrix("T", ()))[1, 1] - cmath.exp(0.25j * cmath.pi)) < 1e-16


def test_xasm_parameters_are_parsed_not_evaluated():
    """load_xasm must not eval() gate parameters (a .xasm file is untrusted input).  Arithmetic with pi works,
    anything else is rejected."""
    n, circ = Cc.load_xasm("Rx(q[0], pi/2);\nRz(q[1], -0.25*pi + 1e-3);\nfSim(q[0], q[3], (1.5707963267948966), 2**-1);\nCX(q[1], q[2]);\n")
    assert n == 4 and circ[0][0] == "Rx" and abs(circ[0][2][0] - math.pi / 2) < 1e-15
    assert abs(circ[1][2][0] - (-0.25 * math.pi + 1e-3)) < 1e-15 and circ[2][2] == (1.5707963267948966, 0.5) and circ[3][0] == "CNOT"
    for bad in ("Rx(q[0], __import__('os').system('true'));",
                "Rx(q[0], ().__class__.__base__.__subclasses__());",
                "Rx(q[0], open('/etc/passwd'));",
                "Rx(q[0], pi if 1 else 2);",
                "Rx(q[0], [1][0]);",
                "Rx(q[0], 'a');"):
        with pytest.raises(ValueError):
            Cc.load_xasm(bad)


def test_real_sycamore_fixture_is_the_reference_circuit():
    """tests/golden/circuits/sycamore_53_14_0.json.gz (written by tests/golden/make_sycamore_fixture.py from the reference's
    examples/sycamore/resources/sycamore_53_14_0.xasm): 53 qubits, 2527 1q + 301 fSim gates, coupler distances 1..10, and
    1897 nearest-neighbour 2q gates after the routing pass.  In the build container the fixture is also
    compared gate for gate with the resource file itself."""
    n, circ = Cc.sycamore_53(14)
    assert n == 53 and Cc.count_gates(circ) == (2527, 301)
    assert {g[0] for g in circ} == {"Rx", "Ry", "Rz", "fSim"}
    dist = {abs(g[1][0] - g[1][1]) for g in circ if len(g[1]) == 2}
    assert min(dist) == 1 and max(dist) == 10
    nn = Cc.nearest_neighbor(circ)
    assert Cc.count_gates(nn) == (2527, 1897) and all(abs(g[1][0] - g[1][1]) == 1 for g in nn if len(g[1]) == 2)
    src = "/root/reference/examples/sycamore/resources/sycamore_53_14_0.xasm"
    if os.path.exists(src):
        n2, ref = Cc.load_xasm(open(src).read())
        assert n2 == n and len(ref) == len(circ)
        for a, b in zip(ref, circ):
            assert a[0] == b[0] and tuple(a[1]) == tuple(b[1]) and np.allclose(a[2], b[2], rtol=0, atol=0)
