"""GPU parity of the Feynman path sum (csrc/qsv_pathsum.cu through Circuit.run_method1 and the raw qsv.pathsum
call) against the reference's own run_method1 weights (tests/golden) and the oracle.  pytest -m gpu."""
import numpy as np
import pytest

import statevec_oracle as so
from helpers import build_circuit

pytestmark = pytest.mark.gpu
TOL = 1e-12

M1_CASES = ["paper5", "builtin_mix", "shor15_a7"] + ["random_seed%d" % s for s in (1, 2, 3, 4, 5, 6, 7, 8, 10)]


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"
    from qsv import _cabi, pathsum
    lib = _cabi.load()
    return lib, pathsum


@pytest.mark.parametrize("name", M1_CASES)
def test_path_sum_kernel_matches_reference_method1(golden, gpu, name, monkeypatch):
    lib, pathsum = gpu
    meta, cases, arrays = golden
    case = cases[name]
    c = build_circuit(case, arrays)
    for i, in_v in enumerate(case["inputs"]):
        ref = arrays[name + "/weights_m1"][i]
        before = lib.qsv_launch_count()
        w = pathsum.weights(c, list(in_v))
        assert lib.qsv_launch_count() > before                      # the kernel ran
        assert w.dtype == np.float64 and w.shape == ref.shape
        assert np.max(np.abs(w - ref)) <= TOL
        monkeypatch.setenv("QSV_METHOD1", "pathsum")
        w_api = c.run_method1(list(in_v))
        assert isinstance(w_api, list) and np.array_equal(np.array(w_api), w)
        monkeypatch.setenv("QSV_METHOD1", "engine")                 # transposed state-vector evolution: same weights
        assert np.max(np.abs(np.array(c.run_method1(list(in_v))) - w)) <= TOL


def test_path_sum_shor21_ten_internal_wires(golden, gpu, monkeypatch):
    """main/Shor.py circuit for N = 21, a = 17: 10 qubits, 10 internal wires, one 1024 x 1024 gate table."""
    lib, pathsum = gpu
    meta, cases, arrays = golden
    from qsv import workloads
    c, q = workloads.shor_circuit(21, 17, dense=True)
    assert len(c.get_internal_wires()) == 10 and pathsum.wanted(c)
    w = pathsum.weights(c, list(meta["shor21"]["input"]))
    assert np.max(np.abs(w - arrays["shor21_a17/weights_m1"][0])) <= TOL


def test_path_sum_direct_wires_and_layers(gpu):
    """Inputs wired straight to outputs (main/circuit.py:240-252) and a 7-qubit layered circuit against the oracle's
    loop-for-loop path sum / transposed evolution."""
    lib, pathsum = gpu
    import main.circuit as mc
    c = mc.Circuit(3)
    h = c.add(mc.H(1, "h"))
    c.add_wire(c, 0, c, 2)
    c.add_wire(c, 1, h, 0)
    c.add_wire(h, 0, c, 0)
    c.add_wire(c, 2, c, 1)
    for in_v in ([0, 0, 0], [1, 0, 1], [0, 1, 1], [1, 1, 0]):
        w = pathsum.weights(c, in_v)
        assert np.max(np.abs(w - so.run_method1_path_sum(c, in_v))) <= TOL
        assert np.count_nonzero(w) == 2
    import random
    random.seed(5)
    big = mc.Circuit.random_circuit(7, 6)
    n_paths = len(big) + len(big.get_internal_wires())
    assert n_paths <= pathsum.MAX_LOG2_PATHS
    for in_v in ([0] * 7, [1, 0, 1, 1, 0, 0, 1]):
        w = pathsum.weights(big, in_v)
        assert np.max(np.abs(w - so.run_method1(big, in_v))) <= TOL
        assert abs(w.sum() - 1.0) <= 1e-10


def test_path_sum_rejects_bad_sizes(gpu):
    import ctypes as C
    lib, pathsum = gpu
    assert lib.qsv_path_sum(None, 1, None, None, None, 0, 1, 0, None, 0, None, None) != 0
    assert b"qsv_path_sum" in lib.qsv_last_error()
