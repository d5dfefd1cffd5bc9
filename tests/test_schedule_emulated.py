"""Lowering + pass planning + program packing, executed by a numpy emulation of the tile kernel that
decodes the packed device image (tests/helpers.py).  Checks all index arithmetic on the CPU."""
import os

import numpy as np
import pytest

import statevec_oracle as so
from helpers import build_circuit, circuit_from_steps, emulate_ops
from qsv import schedule
from qsv.lowering import lower, basis_index

TOL = 1e-12


def _final_output_order(psi, src_pos):
    n = len(src_pos)
    z = np.arange(1 << n)
    y = np.zeros_like(z)
    for b in range(n):
        y |= ((z >> b) & 1) << src_pos[b]
    return psi[y]


def _run_emulated(c, in_v, transposed=False, tile_bits=None, low_bits=None):
    c.sort()
    ops, src_pos, flip = lower(c, transposed, True, in_flip=True)      # the engine's path (lowering.evolve)
    n = len(c)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[basis_index(in_v) ^ flip] = 1
    psi = emulate_ops(ops, psi, n, tile_bits, low_bits)
    return _final_output_order(psi, src_pos), ops


def test_x_propagation_leaves_no_plain_x_and_keeps_results(golden):
    """Uncontrolled X gates are pushed back to the input (a flip of the basis index); without in_flip the
    leftover flips stay as explicit ops at the head of the list.  Both forms must equal the unpropagated
    circuit on every basis input."""
    import os
    from qsv.schedule import propagate_x
    meta, cases, arrays = golden
    for name in ("paper5", "builtin_mix", "random_seed3", "random_seed7", "layered_n6_d2"):
        c = build_circuit(cases[name], arrays)
        c.sort()
        n = len(c)
        os.environ["QSV_PROPAGATE_X"] = "0"
        try:
            plain, src0 = lower(c, False, True)
        finally:
            del os.environ["QSV_PROPAGATE_X"]
        ops, src_pos, flip = lower(c, False, True, in_flip=True)
        compat, src1 = lower(c, False, True)
        assert src0 == src_pos == src1
        assert not any(op.kind == "mcx" and not op.controls for op in ops)
        again, flip2 = propagate_x(ops)
        assert flip2 == 0 and len(again) == len(ops)
        for idx in range(1 << n):
            e = np.zeros(1 << n, dtype=np.complex128)
            e[idx] = 1
            ref = emulate_ops(plain, e, n)
            f = np.zeros(1 << n, dtype=np.complex128)
            f[idx ^ flip] = 1
            assert np.max(np.abs(emulate_ops(ops, f, n) - ref)) <= 1e-14
            assert np.max(np.abs(emulate_ops(compat, e, n) - ref)) <= 1e-14
            if n > 5 and idx > 8:
                break


@pytest.mark.parametrize("name", [
    "paper5", "grover7", "shor15_a7", "layered_n5_d3", "layered_n6_d2", "layered_n7_d2", "builtin_mix"]
    + ["random_seed%d" % s for s in range(1, 13)])
@pytest.mark.parametrize("geom", [(None, None), (4, 2), (3, 0)])
def test_emulated_engine_matches_reference(golden, name, geom):
    meta, cases, arrays = golden
    case = cases[name]
    if name in ("grover7", "shor15_a7") and geom != (None, None):
        pytest.skip("large fixtures only with the default tile")
    c = build_circuit(case, arrays)
    for i, in_v in enumerate(case["inputs"]):
        psi, ops = _run_emulated(c, in_v, tile_bits=geom[0], low_bits=geom[1])
        assert np.max(np.abs(psi - arrays[name + "/amps"][i])) <= TOL
        if name + "/weights_m1" in arrays:
            psi1, _ = _run_emulated(c, in_v, transposed=True, tile_bits=geom[0], low_bits=geom[1])
            assert np.max(np.abs(np.abs(psi1) ** 2 - arrays[name + "/weights_m1"][i])) <= TOL


def test_permutation_circuit_is_bit_exact(golden):
    meta, cases, arrays = golden
    c = build_circuit(cases["paper5"], arrays)
    for i, in_v in enumerate(cases["paper5"]["inputs"]):
        psi, ops = _run_emulated(c, in_v, tile_bits=4, low_bits=1)
        assert all(op.kind == "mcx" for op in ops)
        assert np.array_equal(psi, arrays["paper5/amps"][i])


def test_classification_of_reference_operators():
    # Grover oracle F (main/Grover.py:14-22): one multi-controlled X with 0/1 control values
    F = so.grover_oracle_matrix(6, 13)
    ops = schedule.classify_matrix(F, tuple(range(6, -1, -1)))
    assert len(ops) == 1 and ops[0].kind == "mcx" and ops[0].targets == (0,)
    assert sorted(zip(ops[0].controls, ops[0].cvals)) == [(1 + b, (13 >> b) & 1) for b in range(6)]
    # U0 (x) I: phase -1 on "search register != 0" is not a single controlled phase -> diag table on 6 bits
    U = so.grover_u0_matrix(6)
    ops = schedule.classify_matrix(U, tuple(range(6, -1, -1)))
    assert [o.kind for o in ops] == ["diag"] and ops[0].targets == (6, 5, 4, 3, 2, 1)
    # H^6 (x) I: six 1-qubit dense ops, no op on the ancilla
    HI = np.kron(so.kron_power(so.mat_H(), 6), np.eye(2))
    ops = schedule.classify_matrix(HI, tuple(range(6, -1, -1)))
    assert [o.kind for o in ops] == ["dense"] * 6 and sorted(o.targets[0] for o in ops) == [1, 2, 3, 4, 5, 6]
    prod = np.eye(1)
    for o in sorted(ops, key=lambda o: -o.targets[0]):
        prod = np.kron(prod, o.matrix)
    assert np.max(np.abs(np.kron(prod, np.eye(2)) - HI)) < 1e-15
    # CNot / Toffoli / CZ / Z / phase
    assert [(o.kind, o.targets, o.controls) for o in schedule.classify_matrix(so.mat_Cnot(), (5, 2))] == [("mcx", (2,), (5,))]
    t = schedule.classify_matrix(so.mat_Toffoli(), (7, 3, 1))[0]
    assert t.kind == "mcx" and t.targets == (1,) and sorted(t.controls) == [3, 7] and set(t.cvals) == {1}
    z = schedule.classify_matrix(so.mat_Z(), (4,))[0]
    assert z.kind == "phase" and z.controls == (4,) and z.scalar == -1
    cz = schedule.classify_matrix(so.mat_CZ(), (4, 9))[0]
    assert cz.kind == "phase" and sorted(cz.controls) == [4, 9]
    # Shor U (1024 x 1024 permutation) -> exact wide permutation; QFT(5) -> dense 5-qubit tile op
    perm = so.modexp_permutation(10, 17, 21)
    Umat = np.zeros((1024, 1024))
    Umat[perm, np.arange(1024)] = 1
    ops = schedule.classify_matrix(Umat, tuple(range(9, -1, -1)))
    assert [o.kind for o in ops] == ["wide_perm"] and np.array_equal(ops[0].perm, perm)
    ops = schedule.classify_matrix(so.mat_QFT(5), (9, 8, 7, 6, 5))
    assert [o.kind for o in ops] == ["dense"] and ops[0].matrix.shape == (32, 32)
    assert schedule.classify_matrix(np.eye(8), (2, 1, 0)) == []


def test_pass_planner_respects_dependencies_and_tile_capacity():
    n = 16
    steps = so.random_layered_steps(n, 4, seed=3)
    c = circuit_from_steps(steps, n)
    c.sort()
    ops, src_pos = lower(c)
    for T, L in ((12, 6), (8, 3), (13, 5)):
        passes = schedule.plan_passes(ops, n, T, L)
        assert sum(len(p.ops) for p in passes) == len(ops)
        for p in passes:
            assert len(p.hi_pos) == p.tile_bits - p.low_bits and list(p.hi_pos) == sorted(set(p.hi_pos))
            for op in p.ops:
                assert all(p.loc(b) >= 0 for b in op.nondiag_bits())
        # order preserved for every pair of ops that do not commute structurally
        flat = [id(o) for p in passes for o in p.ops]
        rank = {o: i for i, o in enumerate(flat)}
        for i, a in enumerate(ops):
            for b in ops[i + 1:]:
                a_nd, b_nd = set(a.nondiag_bits()), set(b.nondiag_bits())
                a_all, b_all = a_nd | set(a.diag_bits()), b_nd | set(b.diag_bits())
                if (a_nd & b_all) or (b_nd & a_all):
                    assert rank[id(a)] < rank[id(b)]
    # and the result is still right
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    got = emulate_ops(ops, psi, n, 8, 3)
    ref = so.initial_state([0] * n)
    for M, wires, _ in steps:
        ref = so.apply_gate(ref, n, wires, M)
    assert np.max(np.abs(_final_output_order(got, src_pos) - ref)) <= TOL


def test_single_gate_pass_keeps_controls_external():
    """A lone CNot / CZ on high qubits must leave its control outside the tile so that the kernel skips
    the tiles the control switches off (half / quarter traffic)."""
    n = 20
    cn = schedule.classify_matrix(so.mat_Cnot(), (19, 17))
    p = schedule.plan_passes(cn, n)[0]
    assert p.loc(17) >= 0 and p.loc(19) < 0
    cz = schedule.classify_matrix(so.mat_CZ(), (19, 18))
    p = schedule.plan_passes(cz, n)[0]
    assert p.loc(19) < 0 and p.loc(18) < 0
    packed = schedule.pack_passes([p], register_kernel=False)
    hdr = np.frombuffer(packed.blob.tobytes(), dtype=schedule.OP_DTYPE, count=1, offset=16)[0]
    assert int(hdr["ext_ctrl_mask"]) == (1 << 19) | (1 << 18) and int(hdr["n_ins"]) == 0
    import os
    os.environ["QSV_REG_MIN_OPS"] = "0"          # force the register kernel even for a one-op pass
    try:
        packed = schedule.pack_passes([p], register_kernel=True)
    finally:
        del os.environ["QSV_REG_MIN_OPS"]
    ps, off, marker, (prog, n_rounds) = packed.passes[0]
    assert marker < 0 and n_rounds == 1
    rop = np.frombuffer(prog.tobytes(), dtype=schedule.ROP_DTYPE, count=1, offset=32)[0]
    assert int(rop["ext_mask"]) == (1 << 19) | (1 << 18) and int(rop["kind"]) == schedule.ROP_PHASE_T
    assert int(rop["flags"]) & schedule.ROPF_NEG


def test_round_planner_groups_ops_by_register_bits():
    n = 16
    steps = so.random_layered_steps(n, 6, seed=9)
    c = circuit_from_steps(steps, n)
    c.sort()
    ops, src_pos = lower(c)
    passes = schedule.plan_passes(ops, n, 12, 6)
    total_rounds = 0
    for p in passes:
        rounds = schedule.plan_rounds(p)
        assert rounds is not None
        assert sum(len(R.pre) + len(R.reg) + len(R.post) for R in rounds) == len(p.ops)
        for R in rounds:
            rb = R.rb
            assert len(rb) == schedule.reg_bits() and rb == sorted(set(rb)) and all(0 <= b < 12 for b in rb)
            assert all(op.kind == "mcx" for op in R.pre + R.post)
            for op in R.reg:
                assert op.kind != "mcx"
                for b in op.nondiag_bits():
                    assert p.loc(b) in rb
        total_rounds += len(rounds)
    assert total_rounds < len(ops) / 3          # the point of rounds: several ops per shared-memory trip
    both = []
    for reg in (False, True):
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[5] = 1
        from helpers import emulate_program
        both.append(emulate_program(schedule.pack_passes(passes, register_kernel=reg), psi))
    assert np.max(np.abs(both[0] - both[1])) < 1e-14
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[5] = 1
    for M, wires, _ in steps:
        ref = so.apply_gate(ref, n, wires, M)
    assert np.max(np.abs(both[1] - ref)) <= TOL


def test_symbolic_gates_lower_without_materialising():
    import main.circuit as mc
    n = 30
    c = mc.Circuit(n)
    h = c.add_gate(mc.H, n)
    x = c.add_gate(mc.X, n)
    c.add_wires(c, (), h, ())
    c.add_wires(h, (), x, ())
    c.add_wires(x, (), c, ())
    c.sort()
    ops, src_pos = lower(c)
    assert len(ops) == n and src_pos == list(range(n))      # X.H fused into one 2x2 matrix per qubit
    assert all(op.kind == "dense" and np.allclose(op.matrix, so.mat_X() @ so.mat_H()) for op in ops)
    assert h.matrix.structure() is not None          # still symbolic: nothing built a 2^30 x 2^30 list
    passes = schedule.plan_passes(ops, n)
    assert len(passes) <= 4 + 1


def test_large_qft_gate_uses_ladder_and_relabel():
    import main.circuit as mc
    n = 10
    c = mc.Circuit(n)
    q = c.add_gate(mc.QFT, 9)
    c.add_wires(c, tuple(range(1, 10)), q, ())
    c.add_wire(c, 0, c, 0)
    c.add_wires(q, (), c, tuple(range(1, 10)))
    c.sort()
    ops, src_pos = lower(c)
    assert all(o.kind in ("dense", "phase") for o in ops) and len(ops) == 9 + 36
    rng = np.random.default_rng(1)
    psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    got = _final_output_order(emulate_ops(ops, psi0, n), src_pos)
    ref = so.apply_gate(psi0, n, list(range(1, 10)), so.mat_QFT(9))
    assert np.max(np.abs(got - ref)) < 1e-12


def test_lowering_cache_is_keyed_by_matrix_entries_not_object_identity():
    """lower() caches its result per circuit CONTENT: wires, output wiring and the entries of every gate matrix.
    Callers write into Matrix.rows (main/Grover.py:18, main/Shor.py:78-80), so a gate whose rows were changed in
    place must be lowered afresh; an untouched circuit must come back as a new list with the same ops."""
    from main.circuit import Circuit, H, CNot
    from qsv import lowering
    from qsv.schedule import ops_digest
    c = Circuit(3)
    h, cn = c.add(H()), c.add(CNot())
    c.add_wire(c, 0, h, 0)
    c.add_wire(h, 0, cn, 0)
    c.add_wire(c, 1, cn, 1)
    c.add_wire(cn, 0, c, 0)
    c.add_wire(cn, 1, c, 1)
    c.add_wire(c, 2, c, 2)
    c.sort()
    ops1, src1 = lowering.lower(c)
    ops2, src2 = lowering.lower(c)
    assert ops1 is not ops2 and ops_digest(ops1) == ops_digest(ops2) and src1 == src2
    assert getattr(ops2, "digest", None) == ops_digest(ops2)
    ops2.append("scribble")                       # the caller's list is its own
    assert len(lowering.lower(c)[0]) == len(ops1)
    rows = h.matrix.rows                          # turn the H into a Z in place
    rows[0][0], rows[0][1], rows[1][0], rows[1][1] = 1, 0, 0, -1
    ops3, _ = lowering.lower(c)
    assert ops_digest(ops3) != ops_digest(ops1)
    assert not any(op.kind == "dense" for op in ops3)


# ------------------------------------------------------------------------------------------------
# native tile search (csrc/qsv_plan.cu) and the randomised schedule search
# ------------------------------------------------------------------------------------------------

def _layered_ops(n, depth, seed, g=0):
    from qsv import workloads
    c = workloads.random_layered_circuit(n, depth, seed=seed)
    c.sort()
    ops, src_pos, flip = lower(c, False, True, n_global=g, in_flip=True, free_swap=bool(g), fuse_swaps=bool(g))
    return ops, src_pos, flip


def _schedule_signature(items):
    out = []
    for kind, it in items:
        if kind == "pass":
            out.append(("pass", tuple(it.hi_pos), tuple(id(o) for o in it.ops)))
        elif kind == "swap_pass":
            out.append(("swap_pass", id(it[0]), tuple(it[1].hi_pos), tuple(id(o) for o in it[1].ops), it[2]))
        else:
            out.append(("op", id(it)))
    return out


def test_native_tile_search_reproduces_the_python_planner(monkeypatch):
    """qsv_plan_improve_tile with seed 0 visits the candidates in the same order as schedule._improve_tile: the
    schedules are identical, unsharded and sharded (joint placement + fused swaps)."""
    monkeypatch.setenv("QSV_DISK_CACHE", "0")
    monkeypatch.setenv("QSV_JIT", "force")
    monkeypatch.setenv("QSV_PLACE_SEEDS", "1")      # the shuffled placement candidates exist in the native search only
    for n, g, seed in ((18, 0, 5), (22, 0, 6), (24, 2, 7)):
        sigs = []
        for native in ("1", "0"):
            monkeypatch.setenv("QSV_PLAN_NATIVE", native)
            schedule._PLACE_CACHE.clear()
            from qsv import lowering
            lowering._LOWER_CACHE.clear()
            ops, _, _ = _layered_ops(n, 10, seed, g)
            items = schedule.plan_schedule(ops, n - g, fuse_swaps=bool(g))
            sigs.append([(k, tuple(it.hi_pos), len(it.ops)) if k == "pass" else (k,) for k, it in items])
        assert sigs[0] == sigs[1]


def test_native_absorbed_counts_match_python():
    import ctypes as C
    from qsv import _cabi
    rng = np.random.default_rng(3)
    lib = _cabi.load()
    for _ in range(50):
        n_ops = int(rng.integers(1, 80))
        masks = []
        for _ in range(n_ops):
            nd = int(rng.integers(0, 1 << 16)) & int(rng.integers(0, 1 << 16)) & int(rng.integers(0, 1 << 16))
            dg = int(rng.integers(0, 1 << 16)) & int(rng.integers(0, 1 << 16)) & ~nd
            if rng.random() < 0.05:
                nd |= schedule._NOT_TILED
            masks.append((nd, dg))
        allowed = int(rng.integers(0, 1 << 16))
        nd = np.array([m[0] & schedule._M64 for m in masks], dtype=np.uint64)
        dg = np.array([m[1] & schedule._M64 for m in masks], dtype=np.uint64)
        nv = np.array([1 if m[0] >> 200 else 0 for m in masks], dtype=np.uint8)
        got = lib.qsv_plan_absorbed(C.c_void_p(nd.ctypes.data), C.c_void_p(dg.ctypes.data), C.c_void_p(nv.ctypes.data),
                                    n_ops, allowed)
        assert got == schedule._absorbed(masks, allowed, n_ops)


def test_schedule_search_is_valid_deterministic_and_never_worse_than_greedy(monkeypatch, tmp_path):
    """The randomised search returns a schedule that (a) holds every op exactly once in an order the emulation
    accepts (same final state as the greedy schedule), (b) is the same on every call (fixed seeds: the ranks of a
    sharded register must agree), (c) costs no more than the greedy schedule under the model, (d) survives the
    on-disk plan cache unchanged."""
    monkeypatch.setenv("QSV_JIT", "force")
    monkeypatch.setenv("QSV_CACHE_DIR", str(tmp_path))
    n = 14
    ops, src_pos, flip = _layered_ops(n, 8, 11)
    monkeypatch.setattr(schedule, "PLAN_SEARCH_MIN_QUBITS", 0)
    greedy = schedule.plan_schedule(ops, n)
    monkeypatch.setenv("QSV_DISK_CACHE", "0")
    a = schedule.plan_schedule_search(ops, n, restarts=6)
    b = schedule.plan_schedule_search(ops, n, restarts=6)
    assert _schedule_signature(a) == _schedule_signature(b)
    assert sorted(id(o) for k, it in a for o in it.ops) == sorted(id(o) for o in ops)
    assert schedule.schedule_ms(a, n, 0) <= schedule.schedule_ms(greedy, n, 0) + 1e-9
    monkeypatch.setenv("QSV_DISK_CACHE", "1")
    c1 = schedule.plan_schedule_search(ops, n, restarts=6)          # stores
    c2 = schedule.plan_schedule_search(ops, n, restarts=6)          # loads
    assert _schedule_signature(c1) == _schedule_signature(a) == _schedule_signature(c2)
    assert any(f.startswith("plan-") for f in os.listdir(tmp_path))
    # both schedules give the same state on the emulated kernels
    psi0 = np.zeros(1 << n, dtype=np.complex128)
    psi0[flip] = 1
    ref = emulate_ops(ops, psi0.copy(), n)
    reordered = [o for k, it in a for o in it.ops]
    got = emulate_ops(reordered, psi0.copy(), n)
    assert np.max(np.abs(got - ref)) <= TOL


@pytest.mark.parametrize("idx", range(4))
def test_emulated_engine_on_direct_wire_circuits(idx):
    """Circuits whose inputs are wired straight to outputs (fixtures of oracle/gen_pathsum_golden.py, the reference's own
    run_method1 / run_method2 weights): lowering + pass schedule, emulated on the CPU, for both methods."""
    from test_pathsum_host import _direct_cases, _build_direct
    case = _direct_cases()[idx]
    c = _build_direct(case)
    for in_v, w1, w2 in zip(case["inputs"], case["weights_m1"], case["weights_m2"]):
        psi, _ = _run_emulated(c, in_v)
        assert np.max(np.abs(np.abs(psi) ** 2 - np.array(w2))) <= TOL
        psi1, _ = _run_emulated(c, in_v, transposed=True, tile_bits=2, low_bits=1)
        assert np.max(np.abs(np.abs(psi1) ** 2 - np.array(w1))) <= TOL
