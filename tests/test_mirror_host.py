"""Host logic of the drop-in mirror (quantum-circuits_b200/main) against reference fixtures.  CPU only."""
import inspect
import pickle
import random
import sys

import numpy as np
import pytest

import main.circuit as mc
import main.matrix as mm
from main.matrix import Matrix, MatrixError, tensor, matrix_product
from helpers import build_circuit


def _rows(m):
    return [[complex(re, im) for re, im in row] for row in m]


def test_matrix_factories_bit_identical_to_reference(golden):
    meta = golden[0]
    built = {
        "H1": Matrix.H(), "H3": Matrix.H(3), "X2": Matrix.X(2), "Y2": Matrix.Y(2), "Z2": Matrix.Z(2),
        "SqrtNot2": Matrix.SqrtNot(2), "Phase0.3_2": Matrix.PhaseShift(0.3, 2), "QFT1": Matrix.QFT(1),
        "QFT3": Matrix.QFT(3), "QFT4": Matrix.QFT(4), "Cnot": Matrix.Cnot(), "T": Matrix.T(),
        "Perm201": Matrix.Permutation([2, 0, 1]), "Id4": Matrix.Id(4), "Zero2x3": Matrix.Zero(2, 3),
    }
    A, B = Matrix.QFT(2), Matrix.SqrtNot(2)
    built["QFT2*SqrtNot2"] = A * B
    built["QFT2.conj"] = A.getConjugate()
    built["QFT2.round3"] = A.Round(3)
    built["tensor3"] = tensor([Matrix.H(), Matrix.Y(), Matrix.PhaseShift(1.1)])
    built["matrix_product"] = matrix_product([Matrix.H(), Matrix.Y(), Matrix.SqrtNot()])
    built["apply"] = Matrix.vector([1, 0, 0, 0]).apply([Matrix.H(2), Matrix.Cnot()])
    for name, m in built.items():
        ref = _rows(meta["matrices"][name])
        got = [[complex(v) for v in row] for row in m.rows]
        assert got == ref, name
        ints = all(isinstance(v, int) for row in m.rows for v in row)
        assert ints == meta["matrix_int_entries"][name], name
    assert Matrix.QFT(3).isUnitary() is meta["isUnitary"]["QFT3"]
    assert Matrix.T().isUnitary() is meta["isUnitary"]["T"]
    assert Matrix([[1, 1], [0, 1]]).isUnitary() is meta["isUnitary"]["nonunitary"]
    assert str(Matrix.Cnot()) == meta["str_Cnot"]
    assert repr(Matrix.X()) == meta["repr_X"]


def test_matrix_errors_and_symbolic_powers():
    with pytest.raises(MatrixError, match="Please specify the rows of the matrix."):
        Matrix([])
    with pytest.raises(MatrixError, match="Row dimensions are not consistent."):
        Matrix([[1, 2], [3]])
    with pytest.raises(MatrixError, match="Matrix dimensions must agree."):
        Matrix.Id(2) * Matrix.Id(4)
    with pytest.raises(MatrixError, match="This method is reserved for vectors."):
        Matrix.Id(2).apply([Matrix.Id(2)])
    big = Matrix.H(30)                       # symbolic: no 2^30 x 2^30 list is ever built
    assert len(big) == 2 ** 30 and big.structure()[0] == "power"
    with pytest.raises(MatrixError):
        big.rows
    h2 = Matrix.H(2)
    assert h2.structure() is not None
    assert h2[3][3] == (2 ** -0.5) * (2 ** -0.5)     # materialises on demand
    assert h2.structure() is None
    m = Matrix.X(2)
    m.rows[0] = [1, 0, 0, 0]                            # callers write into .rows (main/Grover.py:18)
    assert m.structure() is None
    t = Matrix([[1, 2], [3, 4]])
    assert t.transpose() is None and t.rows == [[1, 3], [2, 4]]
    assert pickle.loads(pickle.dumps(Matrix.H(2))).rows == Matrix.H(2).rows


def test_star_import_surface():
    ns = {}
    exec("from main.circuit import *", ns)
    for name in ("Matrix", "tensor", "Counter", "product", "choices", "choice", "randint", "log", "modf", "uuid4",
                 "Wire", "Circuit", "Gate", "X", "Y", "Z", "H", "SqrtNot", "QFT", "CNot", "T"):
        assert name in ns, name
    check = lambda t: inspect.isclass(t) and issubclass(t, mc.Gate) and t is not mc.Gate   # gui/main_window.py:269
    palette = sorted(n for n, _ in inspect.getmembers(sys.modules["main.circuit"], check))
    assert palette == ["CNot", "H", "QFT", "SqrtNot", "T", "X", "Y", "Z"]
    assert (mc.Gate.SIZE, mc.CNot.SIZE, mc.T.SIZE) == (1, 2, 3)
    for cls in (mc.X, mc.Y, mc.Z, mc.H, mc.SqrtNot, mc.QFT, mc.CNot, mc.T):
        g = cls()                                   # gui/scene.py:197 instantiates with defaults
        assert pickle.loads(pickle.dumps(cls)) is cls
        assert len(g) == cls.SIZE


def test_graph_logic_matches_reference(golden):
    meta, cases, arrays = golden
    for name, case in cases.items():
        c = build_circuit(case, arrays)
        assert c.check() and not c.contains_cycle()
        before = {g: i for i, g in enumerate(c.gates)}
        c.sort()
        assert [before[g] for g in c.gates] == case["sorted_order"], name     # same DFS order as the reference
        assert [str(g) for g in c.gates] == case["sorted_gate_names"], name
        assert [[c.startqbit(i, p) for p in range(len(g))] for i, g in enumerate(c.gates)] == case["targets"], name
        assert c.startqbits() == case["startqbits"], name
        from qsv.lowering import gate_targets, topological_gates
        table, outq = gate_targets(c, c.gates)
        assert table == case["targets"] and outq == case["startqbits"]
        order = topological_gates(c)
        seen = set()
        for g in order:
            for w in g.in_wires:
                assert w.left is c or w.left in seen
            seen.add(g)


def test_random_circuit_consumes_rng_like_reference(golden):
    meta, cases, arrays = golden
    for seed in range(1, 13):
        case = cases["random_seed%d" % seed]
        random.seed(seed)
        c = mc.Circuit.random_circuit()
        assert len(c) == case["n"]
        c.sort()
        assert [str(g) for g in c.gates] == case["sorted_gate_names"]
        assert [type(g).__name__ for g in c.gates] == [case["gates"][i]["cls"] for i in case["sorted_order"]]
        assert c.startqbits() == case["startqbits"]


def test_error_messages():
    c = mc.Circuit(2)
    h = c.add_gate(mc.H)
    with pytest.raises(RuntimeError, match="Component is already added to a circuit."):
        c.add(h)
    other = mc.H()
    with pytest.raises(RuntimeError, match="Cannot add wire: Start point is not a valid component."):
        c.add_wire(other, 0, h, 0)
    with pytest.raises(RuntimeError, match="Cannot add wire: End point is not a valid component."):
        c.add_wire(h, 0, other, 0)
    with pytest.raises(RuntimeError, match="Cannot add wire: H does not have 2 output ports."):
        c.add_wire(h, 1, c, 0)
    with pytest.raises(RuntimeError, match="Cannot add wire: H does not have 3 input ports."):
        c.add_wire(c, 0, h, 2)
    with pytest.raises(RuntimeError, match="Number of output ports does not match number of input ports."):
        c.add_wires(c, (0, 1), h, ())
    with pytest.raises(RuntimeError, match="Cannot sort - please finish building the circuit first."):
        c.sort()
    with pytest.raises(RuntimeError, match="Cannot sort - please finish building the circuit first."):
        c.run_method2([0, 0])
    with pytest.raises(RuntimeError, match="Gates should be represented by square matrices."):
        mc.Gate(Matrix([[1, 0, 0], [0, 1, 0]]), "bad")
    with pytest.raises(RuntimeError, match="Gates should be represented by unitary matrices."):
        mc.Gate(Matrix([[1, 1], [0, 1]]), "bad")
    with pytest.raises(RuntimeError, match="Invalid gate dimension."):
        mc.Gate(Matrix.Id(3), "bad")
    c2 = mc.Circuit(1)
    x = c2.add_gate(mc.X)
    c2.add_wire(c2, 0, x, 0)
    c2.add_wire(x, 0, c2, 0)
    with pytest.raises(RuntimeError, match="Input length does not match the size of the circuit."):
        c2.run_method2([0, 1])
    with pytest.raises(RuntimeError, match="Invalid input - values must be 0 or 1."):
        c2.run_method2([2])
    with pytest.raises(RuntimeError, match="Please choose one of the avalible methods: 1 or 2"):
        c2.run([0], 3)


def test_cycle_detection_and_removal():
    c = mc.Circuit(1)
    a, b = c.add_gate(mc.CNot), c.add_gate(mc.CNot)
    c.add_wire(c, 0, a, 0)
    c.add_wire(a, 0, b, 0)
    c.add_wire(b, 0, a, 1)
    c.add_wire(a, 1, b, 1)
    c.add_wire(b, 1, c, 0)
    assert c.contains_cycle() and not c.check()
    w = b.out_wires[0]
    c.remove_wire(w)
    assert b.out_wires[0] is None and a.in_wires[1] is None and w not in c.wires
    assert not c.contains_cycle()
    c.remove_gate(b)
    assert b not in c.gates and c.output[0] is None and a.out_wires[0] is None
    assert str(a) == "CNot1"          # two gates named CNot were registered: ids are shown


def test_circuit_pickle_roundtrip(golden):
    meta, cases, arrays = golden
    c = build_circuit(cases["paper5"], arrays)
    d = pickle.loads(pickle.dumps(c))
    d.sort()
    assert [str(g) for g in d.gates] == cases["paper5"]["sorted_gate_names"]
    assert d.startqbits() == cases["paper5"]["startqbits"]
    assert {g: 1 for g in d}                      # gates hashable (gui/scene.py:81)


def test_reference_saved_circuits_unpickle_with_the_mirror(golden):
    """Resources/Saved/*.cir start with a pickled Circuit (gui/scene.py:73-81).  With the mirror on the
    path the reference's own files load into mirror objects (same module paths and attribute layout) and
    lower to the same state as the fixture circuit.  Only where the reference checkout is mounted."""
    import os
    path = "/root/reference/Resources/Saved/5.cir"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    meta, cases, arrays = golden
    with open(path, "rb") as f:
        c = pickle.load(f)
    assert type(c) is mc.Circuit and len(c) == 6 and c.check()
    assert sorted(type(g).__name__ for g in c.gates) == ["CNot", "CNot", "CNot", "T", "X", "X", "X"]
    import statevec_oracle as so
    for i, in_v in enumerate(cases["paper5"]["inputs"]):
        w = so.run_method2(c, in_v)
        assert np.array_equal(w, arrays["paper5/weights_m2"][i])
    with open("/root/reference/Resources/Saved/Grover.cir", "rb") as f:
        g = pickle.load(f)
    assert len(g) == 7 and len(g.gates) == 25 and isinstance(g.gates[1].matrix, Matrix)
    from qsv.lowering import lower
    g.sort()
    ops, src = lower(g)
    assert {o.kind for o in ops} <= {"dense", "mcx", "diag", "phase"}


def test_lazy_kronecker_products_and_identities():
    """(f-2) Matrix.bintensor with a symbolic factor / a large result and Matrix.Id(2^12+) stay symbolic: building
    `tensor([Matrix.H(n), Matrix.Id(2)])` (main/Grover.py:42-44) at n = 24 costs nothing, the lowering takes the factors
    one by one, and `.rows` - when somebody does ask - holds exactly the reference's entries."""
    import time
    from main.matrix import Matrix, tensor
    import main.circuit as mc
    from qsv import lowering
    t0 = time.perf_counter()
    big = tensor([Matrix.H(24), Matrix.Id(2)])
    gate = mc.Gate(big, "h", False)                      # error checks on: unitarity is decided on the structure
    assert time.perf_counter() - t0 < 0.5 and len(gate) == 25 and big.structure()[0] == "kron"
    c = mc.Circuit(25)
    g = c.add(gate)
    c.add_wires(c, (), g, ())
    c.add_wires(g, (), c, ())
    c.sort()
    ops, src = lowering.lower(c)
    assert len(ops) == 24 and all(o.kind == "dense" and len(o.targets) == 1 for o in ops)
    assert sorted(o.targets[0] for o in ops) == list(range(1, 25))        # wire 24 (position 0) carries the identity
    # small products with a symbolic factor: the expansion equals the eager product entry by entry
    lazy = tensor([Matrix.H(3), Matrix.Id(2), Matrix.PhaseShift(0.3)])
    assert lazy.structure()[0] == "kron"
    eager = Matrix(Matrix.H(3).rows).bintensor(Matrix.Id(2)).bintensor(Matrix.PhaseShift(0.3))
    assert eager.structure() is None and lazy.rows == eager.rows and lazy.structure() is None
    assert Matrix.Id(1 << 12).structure() == ("id", 1 << 12) and Matrix.Id(1 << 12).isUnitary()
    m = Matrix.Id(1 << 12)
    m[5][5] = 0                                          # a write expands it, like the reference's list of lists
    m[5][6] = 1
    assert m.structure() is None and m.rows[5][6] == 1 and m.rows[7][7] == 1
    assert not tensor([Matrix([[1, 0], [0, 2]]), Matrix.H(2)]).isUnitary()


def test_sort_is_remembered_per_graph_and_forgets_on_any_change():
    """Circuit.sort() (reference main/circuit.py:173-184) walks the whole graph; run() calls it every time.  The mirror
    remembers the result per graph structure: the remembered order must be the order an uncached sort produces, any
    change of the graph (new gate, re-wired port, re-ordered gate list) must be seen, an unfinished circuit must raise
    as before, and nothing is stored on the instance (pickles stay what the reference writes)."""
    import pickle
    import random
    import main.circuit as mc
    random.seed(11)
    c = mc.Circuit.random_circuit(6, 9)
    names0 = [id(g) for g in c.gates]
    orders = []
    for _ in range(4):
        c.sort()
        orders.append([id(g) for g in c.gates])
    # the same sequence with the memory wiped before every call
    c.gates = [g for i in names0 for g in c.gates if id(g) == i]
    plain = []
    for _ in range(4):
        mc._SORT_CACHE.clear()
        c.sort()
        plain.append([id(g) for g in c.gates])
    assert orders == plain
    assert c in mc._SORT_CACHE and "_SORT_CACHE" not in c.__dict__ and not any(k.startswith("_qsv") for k in c.__dict__)
    blob = pickle.dumps(c)
    assert b"SORT_CACHE" not in blob
    # a structure tuple is returned for the lowering; it changes with the graph
    s1 = c._sorted()
    assert s1 is not None and s1 == c._structure()         # (a re-sort may order commuting gates differently again)
    h = c.gates[-1]
    c.gates.reverse()                                   # same gates, another order
    assert c._structure() != s1
    c.sort()
    # re-wire: cut an output wire of the last gate -> unfinished circuit, same error as the reference
    w = h.out_wires[0]
    c.remove_wire(w)
    with pytest.raises(RuntimeError, match="finish building"):
        c.sort()
    c.add_wire(h, w.lind, w.right, w.rind)
    c.sort()
    before = len(mc._SORT_CACHE[c])
    g = c.add(mc.H(1, "late"))                          # a new gate on output 0's wire
    wo = c.output[0]
    src, sp = wo.left, wo.lind
    c.remove_wire(wo)
    c.add_wire(src, sp, g, 0)
    c.add_wire(g, 0, c, 0)
    c.sort()
    assert g in c.gates and len(mc._SORT_CACHE[c]) >= min(before + 1, 4) - 1
    pos = {id(x): i for i, x in enumerate(c.gates)}
    for x in c.gates:                                   # a topological order: every gate after its producers
        for wv in x.in_wires:
            assert wv.left is c or pos[id(wv.left)] < pos[id(x)]


def test_remembered_sort_does_not_keep_circuits_alive():
    import gc
    import random
    import weakref
    import main.circuit as mc
    from qsv import lowering
    random.seed(3)
    c = mc.Circuit.random_circuit(5, 6)
    st = c._sorted()
    c._sorted()
    lowering.lower(c, structure=c._sorted())
    assert c in mc._SORT_CACHE
    r = weakref.ref(c)
    del c
    gc.collect()
    assert r() is None, "a cache entry holds the circuit (or one of its gates, which point back at it)"
    assert st is not None
