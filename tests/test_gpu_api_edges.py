"""Edge cases of the mirror API on the GPU: tiny registers, wires without gates, every built-in gate
class with a size argument, wide and symbolic gates, RNG consumption, large registers (lazy weights)."""
import io
import contextlib
import random

import numpy as np
import pytest

import statevec_oracle as so

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _oracle_weights(c, in_v):
    return so.run_method2(c, in_v)


def test_one_and_two_qubit_circuits():
    import main.circuit as mc
    for cls in (mc.X, mc.Y, mc.Z, mc.H, mc.SqrtNot, mc.QFT):
        c = mc.Circuit(1)
        g = c.add_gate(cls)
        c.add_wire(c, 0, g, 0)
        c.add_wire(g, 0, c, 0)
        for v in ([0], [1]):
            assert np.max(np.abs(np.array(c.run_method2(v)) - _oracle_weights(c, v))) <= TOL
    c = mc.Circuit(2)
    cn = c.add_gate(mc.CNot)
    h = c.add_gate(mc.H)
    c.add_wire(c, 1, h, 0)
    c.add_wire(h, 0, cn, 0)          # control = wire 1 after H
    c.add_wire(c, 0, cn, 1)
    c.add_wire(cn, 0, c, 0)          # outputs crossed
    c.add_wire(cn, 1, c, 1)
    for v in ([0, 0], [0, 1], [1, 0], [1, 1]):
        assert np.max(np.abs(np.array(c.run_method2(v)) - _oracle_weights(c, v))) <= TOL
        assert np.max(np.abs(np.array(c.run_method1(v)) - so.run_method1(c, v))) <= TOL


def test_circuit_without_gates_and_with_bare_wires():
    import main.circuit as mc
    c = mc.Circuit(3)
    c.add_wire(c, 0, c, 2)           # pure wiring: output 2 shows input 0, ...
    c.add_wire(c, 1, c, 0)
    c.add_wire(c, 2, c, 1)
    for v in ([1, 0, 0], [0, 1, 1]):
        w = c.run_method2(v)
        assert np.array_equal(np.array(w), _oracle_weights(c, v))
        with contextlib.redirect_stdout(io.StringIO()):
            assert c.run(v, 2) == [v[1], v[2], v[0]]
    x = c.add_gate(mc.X)             # incomplete circuit -> the reference's error
    with pytest.raises(RuntimeError, match="Cannot sort - please finish building the circuit first."):
        c.run_method2([0, 0, 0])


def test_sized_gate_classes_and_wide_gates():
    import main.circuit as mc
    from main.matrix import Matrix
    n = 8
    c = mc.Circuit(n)
    layers = [mc.H(3), mc.Y(2), mc.SqrtNot(2), mc.Z(2), mc.X(3), mc.QFT(6), mc.QFT(2)]
    spans = [(0, 1, 2), (3, 4), (5, 6), (7, 0), (1, 3, 5), (7, 6, 4, 2, 0, 1), (3, 5)]
    last = [(c, i) for i in range(n)]
    for g, wires in zip(layers, spans):
        c.add(g)
        for p, w in enumerate(wires):
            c.add_wire(last[w][0], last[w][1], g, p)
            last[w] = (g, p)
    dense = mc.Gate(Matrix([[0.6, 0.8], [0.8j, -0.6j]]), "U", False)     # unitary check runs (host)
    c.add(dense)
    c.add_wire(last[4][0], last[4][1], dense, 0)
    last[4] = (dense, 0)
    for w in range(n):
        c.add_wire(last[w][0], last[w][1], c, (w + 3) % n)
    for v in ([0] * n, [1, 0, 1, 1, 0, 0, 1, 0]):
        w = np.array(c.run_method2(v))
        assert np.max(np.abs(w - _oracle_weights(c, v))) <= TOL
        assert abs(w.sum() - 1) < 1e-12


def test_symbolic_qft_ladder_matches_dense_reference_matrix():
    import main.circuit as mc
    n = 10
    c = mc.Circuit(n)
    h = c.add_gate(mc.H, n)
    q = c.add_gate(mc.QFT, 9)                       # above QFT_DENSE_MAX: H + controlled-phase ladder
    c.add_wires(c, (), h, ())
    c.add_wires(h, tuple(range(1, 10)), q, ())
    c.add_wire(h, 0, c, 9)
    c.add_wires(q, (), c, tuple(range(9)))
    from qsv import lowering
    v = [1, 0, 1, 1, 0, 1, 0, 0, 1, 1]
    amps = lowering.simulate_amplitudes(c, v)
    steps = [(so.kron_power(so.mat_H(), n), list(range(n))), (so.mat_QFT(9), list(range(1, 10)))]
    ref = so.evolve(steps, [1, 2, 3, 4, 5, 6, 7, 8, 9, 0], v)
    assert np.max(np.abs(amps - ref)) <= TOL


def test_run_consumes_exactly_one_random_number():
    import main.circuit as mc
    c = mc.Circuit(4)
    h = c.add_gate(mc.H, 4)
    c.add_wires(c, (), h, ())
    c.add_wires(h, (), c, ())
    for method in (None, 1, 2):
        random.seed(99)
        with contextlib.redirect_stdout(io.StringIO()):
            out = c.run([0, 1, 0, 1], method)
        after = random.random()
        random.seed(99)
        u = random.random()
        assert after == random.random()
        w = so.run_method2(c, [0, 1, 0, 1])
        assert out == so.index_bits(so.sample_index(w, u), 4)


def test_large_register_lazy_weights_and_sample():
    import main.circuit as mc
    from qsv import lowering
    n = 26
    c = mc.Circuit(n)
    h = c.add_gate(mc.H, n - 1)
    x = c.add_gate(mc.X)
    c.add_wires(c, tuple(range(n - 1)), h, ())
    c.add_wire(c, n - 1, x, 0)
    c.add_wires(h, (), c, tuple(range(n - 1)))
    c.add_wire(x, 0, c, n - 1)
    w = c.run_method2([0] * n)
    assert isinstance(w, lowering.LazyWeights) and len(w) == 1 << n
    assert w[0] == 0.0 and abs(w[1] - 2.0 ** -(n - 1)) < 1e-20        # last wire flipped to 1
    assert abs(float(w.tensor().sum()) - 1.0) < 1e-10
    random.seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        out = c.run([0] * n, 2)
    assert len(out) == n and out[-1] == 1
    random.seed(5)
    u = random.random()
    # uniform distribution over the 2^(n-1) odd indices: the outcome index is floor-determined by u
    idx = int("".join(map(str, out)), 2)
    assert abs(idx / 2.0 ** n - u) < 2.0 ** -(n - 2)


def test_get_subcircuit_matches_circuit_action():
    import main.circuit as mc
    c = mc.Circuit(3)
    h = c.add_gate(mc.H)
    cn = c.add_gate(mc.CNot)
    t = c.add_gate(mc.T)
    c.add_wire(c, 0, h, 0)
    c.add_wire(h, 0, cn, 0)
    c.add_wire(c, 1, cn, 1)
    c.add_wire(cn, 0, t, 0)
    c.add_wire(cn, 1, t, 1)
    c.add_wire(c, 2, t, 2)
    c.add_wires(t, (), c, ())
    g = c.get_subcircuit("sub")
    assert len(g) == 3 and g.name == "sub"
    U = np.array(g.matrix.rows, dtype=complex)
    steps, sq = so.trace(c)
    for col in range(8):
        v = so.index_bits(col, 3)
        assert np.max(np.abs(U[:, col] - so.evolve(steps, sq, v))) <= TOL


def test_get_subcircuit_against_reference_matrices(golden):
    """Circuit.get_subcircuit (main/circuit.py:122-157) pinned on the reference's own output for the small
    random circuits.  For random_seed3 (Z, CNot on crossed wires) the reference's swap bookkeeping
    (main/circuit.py:130-138, different from run_method2's) yields a non-unitary matrix and its own Gate
    check raises; the engine returns the true circuit unitary there."""
    meta, cases, arrays = golden
    from helpers import build_circuit
    seen = 0
    for name, info in meta["subcircuit_cases"].items():
        c = build_circuit(cases[name], arrays)
        g = c.get_subcircuit("sub")
        U = np.array(g.matrix.rows, dtype=complex)
        if info["error"] is None:
            assert np.max(np.abs(U - arrays["subcircuit/" + name])) <= TOL, name
            seen += 1
        else:
            steps, sq = so.trace(c)
            for col in range(1 << info["n"]):
                assert np.max(np.abs(U[:, col] - so.evolve(steps, sq, so.index_bits(col, info["n"])))) <= TOL
    assert seen >= 5


def test_get_subcircuit_batched_10_qubits():
    """All 1024 columns of a 10-qubit layered circuit in ONE 20-qubit batch register (lowering.circuit_unitary):
    sampled columns against the oracle, unitarity of the whole matrix, and a single D2H of the matrix."""
    from qsv import lowering, workloads
    n = 10
    c = workloads.random_layered_circuit(n, 5, seed=21)
    U = lowering.circuit_unitary(c)
    st = lowering.last_run_stats()
    assert U.shape == (1 << n, 1 << n) and st["d2h_bytes"] == 16 << (2 * n)
    steps = so.random_layered_steps(n, 5, seed=21)
    for col in (0, 1, 513, 1023, 700):
        ref = np.zeros(1 << n, dtype=np.complex128)
        ref[col] = 1
        for M, wires, _ in steps:
            ref = so.apply_gate(ref, n, wires, M)
        assert np.max(np.abs(U[:, col] - ref)) <= TOL
    assert np.max(np.abs(U.conj().T @ U - np.eye(1 << n))) <= 1e-12
    g = c.get_subcircuit("layered")
    # (a second sort() of the already sorted circuit may order commuting gates differently: last-bit differences)
    assert len(g) == n and abs(g.matrix.rows[3][5] - complex(U[3, 5])) <= 1e-14
