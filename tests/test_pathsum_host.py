"""Feynman path sum (Circuit.run_method1, reference main/circuit.py:221-276): the oracle's loop-for-loop restatement
and the kernel's packed description (qsv.pathsum.describe, evaluated by the numpy emulation of csrc/qsv_pathsum.cu)
against the weights the reference's own run_method1 produced (tests/golden).  CPU only."""
import numpy as np
import pytest

import statevec_oracle as so
from helpers import build_circuit

TOL = 1e-12
SMALL = ["paper5", "builtin_mix"] + ["random_seed%d" % s for s in (1, 2, 3, 4, 5, 6, 7, 8, 10)]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_path_sum_matches_reference_method1(golden, name):
    meta, cases, arrays = golden
    case = cases[name]
    c = build_circuit(case, arrays)
    for i, in_v in enumerate(case["inputs"][:4]):
        w = so.run_method1_path_sum(c, list(in_v))
        assert np.max(np.abs(w - arrays[name + "/weights_m1"][i])) <= TOL
        assert np.max(np.abs(w - so.run_method1(c, list(in_v)))) <= TOL      # == transposed evolution


@pytest.mark.parametrize("name", SMALL)
def test_packed_description_evaluates_to_reference_method1(golden, name):
    from qsv import pathsum
    meta, cases, arrays = golden
    case = cases[name]
    c = build_circuit(case, arrays)
    assert pathsum.feasible(c)
    for i, in_v in enumerate(case["inputs"][:4]):
        desc = pathsum.describe(c, list(in_v))
        wires, hdr, ports, mats, direct, n_int = desc
        assert n_int == len(c.get_internal_wires()) and hdr.shape == (len(c.gates), 3)
        assert wires.dtype == np.uint8 and hdr.dtype == np.uint32 and ports.dtype == np.uint32
        w = pathsum.emulate(desc, len(c))
        assert np.max(np.abs(w - arrays[name + "/weights_m1"][i])) <= TOL


def test_direct_wire_lets_only_the_input_value_through():
    """main/circuit.py:240-252: an input wired straight to an output zeroes every `out` that disagrees."""
    import main.circuit as mc
    from qsv import pathsum
    c = mc.Circuit(3)
    h = c.add(mc.H(1, "h"))
    c.add_wire(c, 0, c, 2)                  # direct, crossing positions
    c.add_wire(c, 1, h, 0)
    c.add_wire(h, 0, c, 0)
    c.add_wire(c, 2, c, 1)                  # direct
    for in_v in ([0, 0, 0], [1, 0, 1], [0, 1, 1], [1, 1, 0]):
        desc = pathsum.describe(c, in_v)
        assert desc[4].shape == (2, 2)
        w = pathsum.emulate(desc, 3)
        assert np.max(np.abs(w - so.run_method1_path_sum(c, in_v))) <= TOL
        assert np.max(np.abs(w - so.run_method1(c, in_v))) <= TOL
        assert abs(w.sum() - 1.0) <= TOL and np.count_nonzero(w) == 2


def test_selector_limits(monkeypatch):
    import main.circuit as mc
    from qsv import pathsum
    c = mc.Circuit.random_circuit(4, 3)
    monkeypatch.delenv("QSV_METHOD1", raising=False)
    assert pathsum.wanted(c)
    monkeypatch.setenv("QSV_METHOD1", "engine")
    assert not pathsum.wanted(c)
    big = mc.Circuit(14)                    # a 14-qubit gate: no dense table
    g = big.add(mc.H(14, "h"))
    for i in range(14):
        big.add_wire(big, i, g, i)
        big.add_wire(g, i, big, i)
    monkeypatch.setenv("QSV_METHOD1", "pathsum")
    assert not pathsum.wanted(big)


def _direct_cases():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pathsum_direct.json")
    with open(path) as f:
        return json.load(f)["cases"]


def _build_direct(case):
    import main.circuit as mc
    c = mc.Circuit(case["n"])
    objs = [c.add(getattr(mc, g["cls"])(*g["args"])) for g in case["gates"]]
    for s, sp, d, dp in case["wires"]:
        c.add_wire(c if s < 0 else objs[s], sp, c if d < 0 else objs[d], dp)
    return c


@pytest.mark.parametrize("idx", range(4))
def test_direct_wire_circuits_match_the_reference(idx):
    """tests/golden/pathsum_direct.json (oracle/gen_pathsum_golden.py: the unmodified reference's run_method1 and
    run_method2 on circuits whose inputs are wired straight to outputs, every basis input): the oracle's loop-for-loop
    path sum, its transposed evolution, its run_method2 and the kernel's packed description must all reproduce them."""
    from qsv import pathsum
    case = _direct_cases()[idx]
    c = _build_direct(case)
    assert any(w.right is c for w in c.input), "fixture without a direct wire"
    for in_v, w1, w2 in zip(case["inputs"], case["weights_m1"], case["weights_m2"]):
        w1, w2 = np.array(w1), np.array(w2)
        assert np.max(np.abs(so.run_method1_path_sum(c, list(in_v)) - w1)) <= TOL
        assert np.max(np.abs(so.run_method1(c, list(in_v)) - w1)) <= TOL
        assert np.max(np.abs(so.run_method2(c, list(in_v)) - w2)) <= TOL
        desc = pathsum.describe(c, list(in_v))
        assert desc[4].shape[0] == sum(1 for w in c.input if w.right is c)
        assert np.max(np.abs(pathsum.emulate(desc, len(c)) - w1)) <= TOL
