"""GPU parity: the CUDA engine (through the mirror API and through the raw C-ABI) against the
reference's golden vectors and the CPU oracle.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import io
import contextlib
import math
import random

import numpy as np
import pytest

import statevec_oracle as so
from helpers import build_circuit, circuit_from_steps

pytestmark = pytest.mark.gpu
TOL = 1e-12   # north_star: max-abs per amplitude for floating-point gates; index work is bit-exact


@pytest.fixture(scope="module")
def qsv_gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"
    from qsv import _cabi, engine, lowering
    return _cabi.load(), engine, lowering, torch


ALL_CASES = ["paper5", "grover7", "shor15_a2", "shor15_a7", "shor15_a13", "layered_n5_d3", "layered_n6_d2",
             "layered_n7_d2", "builtin_mix"] + ["random_seed%d" % s for s in range(1, 13)]


@pytest.mark.parametrize("name", ALL_CASES)
def test_golden_circuits_weights_and_amplitudes(golden, qsv_gpu, name):
    lib, engine, lowering, torch = qsv_gpu
    meta, cases, arrays = golden
    case = cases[name]
    c = build_circuit(case, arrays)
    for i, in_v in enumerate(case["inputs"]):
        w = c.run_method2(list(in_v))
        assert isinstance(w, list) and len(w) == 1 << case["n"]
        assert np.max(np.abs(np.array(w) - arrays[name + "/weights_m2"][i])) <= TOL
        amps = lowering.simulate_amplitudes(c, in_v)
        assert np.max(np.abs(amps - arrays[name + "/amps"][i])) <= TOL
        if name + "/weights_m1" in arrays:
            w1 = c.run_method1(list(in_v))
            assert np.max(np.abs(np.array(w1) - arrays[name + "/weights_m1"][i])) <= TOL


def test_paper_circuit_truth_table_bit_exact(golden, qsv_gpu):
    """main/computational-tests.py:9-71: out = [a, b, !a, !b, !a&!b, a|b] through run() (auto -> method 1)
    and run(in_v, 2); permutation-only circuit -> amplitudes bit-identical to the reference's."""
    lib, engine, lowering, torch = qsv_gpu
    meta, cases, arrays = golden
    case = cases["paper5"]
    c = build_circuit(case, arrays)
    for i, in_v in enumerate(case["inputs"]):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            o1 = c.run(list(in_v))
            o2 = c.run(list(in_v), 2)
        assert buf.getvalue() == "Running method 1...\n"
        a, b = in_v[0], in_v[1]
        assert o1 == o2 == [a, b, 1 - a, 1 - b, (1 - a) & (1 - b), a | b] == case["run_outputs"][i][0]
        assert np.array_equal(lowering.simulate_amplitudes(c, in_v), arrays["paper5/amps"][i])
        assert np.array_equal(np.array(c.run_method2(list(in_v))), arrays["paper5/weights_m2"][i])


def test_grover_reference_script_circuit(golden, qsv_gpu):
    lib, engine, lowering, torch = qsv_gpu
    meta, cases, arrays = golden
    from qsv import workloads
    c = workloads.grover_circuit_dense(6, 13)
    w = np.array(c.run_method2([0] * 6 + [1]))
    assert np.max(np.abs(w - arrays["grover7/weights_m2"][0])) <= TOL
    assert abs(w[26] - 0.4982928403934023) <= TOL and abs(w[27] - 0.4982928403934023) <= TOL
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        random.seed(0)
        out = c.run([0] * 6 + [1])
    assert buf.getvalue() == "Running method 2...\n"
    random.seed(0)
    expect = so.index_bits(so.sample_index(arrays["grover7/weights_m2"][0], random.random()), 7)
    assert out == expect
    # structured gates (oracle + diffusion kernels) give the same state
    cs = workloads.grover_circuit(6, 13, 6)
    amps = lowering.simulate_amplitudes(cs, [0] * 6 + [1])
    assert np.max(np.abs(amps - arrays["grover7/amps"][0])) <= TOL


def test_shor15_outcome_sequences_bit_exact(golden, qsv_gpu):
    """random.seed(s); Shor(15): every run() outcome and the factor list equal the reference's
    (one random() per run, sampled index bit-exact)."""
    meta, cases, arrays = golden
    from qsv import workloads
    for dense in (True, False):
        for seed, rec in meta["shor15_sequences"].items():
            random.seed(int(seed))
            runs = []
            with contextlib.redirect_stdout(io.StringIO()):
                factors = workloads.shor_factor(15, dense=dense, on_run=lambda a, out: runs.append(out))
            assert runs == rec["runs"], (seed, dense)
            assert factors == rec["factors"]


def test_shor21_example_weights_and_modexp_kernel(golden, qsv_gpu):
    lib, engine, lowering, torch = qsv_gpu
    meta, cases, arrays = golden
    from qsv import workloads
    in_v = meta["shor21"]["input"]
    for dense in (True, False):
        c, q = workloads.shor_circuit(21, 17, dense=dense)
        w = np.array(c.run_method1(list(in_v)))
        assert np.max(np.abs(w - arrays["shor21_a17/weights_m1"][0])) <= TOL
    # the modexp kernel alone is an exact index permutation equal to the reference's matrix
    for key in [k for k in arrays.files if k.startswith("modexp_perm/")]:
        N, a = (int(x) for x in key.split("/")[1].split("_"))
        q = (len(bin(N)) - 2) * 2
        sv = engine.StateVector(q)
        rng = np.random.default_rng(q)
        psi = rng.normal(size=1 << q) + 1j * rng.normal(size=1 << q)
        sv.state.copy_(torch.from_numpy(psi))
        h = q // 2
        from qsv.schedule import Op
        sv.run_ops([Op("modexp", tuple(range(h - 1, -1, -1)), tuple(range(q - 1, h - 1, -1)),
                       params={"a": a, "N": N})])
        out = sv.amplitudes()
        expect = np.empty_like(psi)
        expect[arrays[key]] = psi
        assert np.array_equal(out, expect)


def test_factor_lists_2_to_31(golden, qsv_gpu):
    """main/tests.txt (output of Shor_test(32)): same prime factor multisets."""
    meta = golden[0]
    from qsv import workloads
    random.seed(12345)
    for n in range(2, 32):
        with contextlib.redirect_stdout(io.StringIO()):
            f = workloads.shor_factor(n)
        assert sorted(f) == meta["tests_txt_factors"][str(n)], n


# ---------------------------------------------------------------------------------------------
# raw C-ABI entry points on random states, every target position class (low / mid / high bits)
# ---------------------------------------------------------------------------------------------

def _rand_state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return psi / np.linalg.norm(psi)


def _wires(n, positions):
    return [n - 1 - p for p in positions]


def _rand_unitary(k, seed):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    Q, _ = np.linalg.qr(A)
    return Q


@pytest.mark.parametrize("n", [1, 3, 9, 14])
def test_cabi_single_gate_entry_points(qsv_gpu, n):
    lib, engine, lowering, torch = qsv_gpu
    from qsv import _cabi
    rng = random.Random(n)
    sv = engine.StateVector(n)
    st = sv._stream()
    for trial in range(24):
        psi = _rand_state(n, 100 * n + trial)
        sv.state.copy_(torch.from_numpy(psi))
        k = rng.randint(1, min(n, 5))
        pos = rng.sample(range(n), k)
        rest = [p for p in range(n) if p not in pos]
        kind = trial % 4
        if kind == 0:
            nc = rng.randint(0, min(2, len(rest)))
            ctr = rng.sample(rest, nc)
            U = _rand_unitary(k, trial)
            flat = np.ascontiguousarray(U).view(np.float64).reshape(-1)
            _cabi.check(lib.qsv_apply_dense(sv._ptr(sv.state), n, k, _cabi.int_array(pos), nc, _cabi.int_array(ctr),
                                            flat.ctypes.data_as(C.POINTER(C.c_double)), 0, st), "dense")
            M = np.eye(1 << (k + nc), dtype=complex)
            M[-(1 << k):, -(1 << k):] = U
            expect = so.apply_gate(psi, n, _wires(n, ctr + pos), M)
        elif kind == 1:
            d = np.exp(1j * np.random.default_rng(trial).uniform(0, 6.28, size=1 << k))
            flat = np.ascontiguousarray(d).view(np.float64)
            _cabi.check(lib.qsv_apply_diag(sv._ptr(sv.state), n, k, _cabi.int_array(pos),
                                           flat.ctypes.data_as(C.POINTER(C.c_double)), 0, st), "diag")
            expect = so.apply_gate(psi, n, _wires(n, pos), np.diag(d))
        elif kind == 2:
            ph = cmath_phase = complex(math.cos(0.1 * trial), math.sin(0.1 * trial))
            _cabi.check(lib.qsv_apply_phase(sv._ptr(sv.state), n, k, _cabi.int_array(pos), ph.real, ph.imag, 0, st),
                        "phase")
            d = np.ones(1 << k, dtype=complex)
            d[-1] = ph
            expect = so.apply_gate(psi, n, _wires(n, pos), np.diag(d))
        else:
            tgt, ctr = pos[0], pos[1:]
            _cabi.check(lib.qsv_apply_cx(sv._ptr(sv.state), n, len(ctr), _cabi.int_array(ctr), tgt, 0, st), "cx")
            M = np.eye(1 << k, dtype=complex)
            M[-2:, -2:] = [[0, 1], [1, 0]]
            expect = so.apply_gate(psi, n, _wires(n, ctr + [tgt]), M)
        got = sv.amplitudes()
        if kind == 3:
            assert np.array_equal(got, expect)          # pure data movement: bit exact
        else:
            assert np.max(np.abs(got - expect)) <= TOL


def test_cabi_wide_gates(qsv_gpu):
    lib, engine, lowering, torch = qsv_gpu
    from qsv import _cabi
    n, k = 9, 7
    sv = engine.StateVector(n)
    psi = _rand_state(n, 5)
    pos = [8, 0, 3, 5, 1, 7, 2]
    U = _rand_unitary(k, 3)
    sv.state.copy_(torch.from_numpy(psi))
    out = torch.empty_like(sv.state)
    Md = torch.from_numpy(np.ascontiguousarray(U)).cuda()
    _cabi.check(lib.qsv_apply_dense_wide(sv._ptr(sv.state), sv._ptr(out), n, k, _cabi.int_array(pos), sv._ptr(Md),
                                         sv._stream()), "dense_wide")
    assert np.max(np.abs(out.cpu().numpy() - so.apply_gate(psi, n, _wires(n, pos), U))) <= TOL
    perm = np.random.default_rng(1).permutation(1 << k).astype(np.uint32)
    Pd = torch.from_numpy(perm.view(np.int32)).cuda()
    _cabi.check(lib.qsv_apply_perm_wide(sv._ptr(sv.state), sv._ptr(out), n, k, _cabi.int_array(pos), sv._ptr(Pd),
                                        sv._stream()), "perm_wide")
    Pm = np.zeros((1 << k, 1 << k))
    Pm[perm, np.arange(1 << k)] = 1
    assert np.array_equal(out.cpu().numpy(), so.apply_gate(psi, n, _wires(n, pos), Pm))
    d = np.exp(1j * np.random.default_rng(2).uniform(0, 6.28, size=1 << k))
    Dd = torch.from_numpy(d).cuda()
    _cabi.check(lib.qsv_apply_diag_wide(sv._ptr(sv.state), n, k, _cabi.int_array(pos), sv._ptr(Dd), sv._stream()),
                "diag_wide")
    assert np.max(np.abs(sv.amplitudes() - so.apply_gate(psi, n, _wires(n, pos), np.diag(d)))) <= TOL


@pytest.mark.parametrize("geom", [(12, 6), (13, 5), (8, 3), (5, 0)])
def test_layered_circuit_16q_all_tile_geometries(qsv_gpu, geom):
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    n = 16
    c = workloads.random_layered_circuit(n, 6, seed=11)
    c.sort()
    ops, src_pos = lowering.lower(c)
    sv = engine.StateVector(n)
    sv.tile_bits, sv.low_bits = geom
    sv.init_basis(0)
    sv.run_ops(ops)
    ref = so.initial_state([0] * n)
    for M, wires, _ in so.random_layered_steps(n, 6, seed=11):
        ref = so.apply_gate(ref, n, wires, M)
    assert src_pos == list(range(n))
    assert np.max(np.abs(sv.amplitudes() - ref)) <= TOL


def test_measurement_kernels_and_sampling(qsv_gpu):
    lib, engine, lowering, torch = qsv_gpu
    n = 13
    sv = engine.StateVector(n)
    psi = _rand_state(n, 77)
    sv.state.copy_(torch.from_numpy(psi))
    src = list(np.random.default_rng(3).permutation(n))
    z = np.arange(1 << n)
    y = np.zeros_like(z)
    for b in range(n):
        y |= ((z >> b) & 1) << int(src[b])
    w_ref = (psi[y].real ** 2 + psi[y].imag ** 2)
    w = sv.probs(src).cpu().numpy()
    assert np.array_equal(w, w_ref) or np.max(np.abs(w - w_ref)) < 1e-18
    sums = sv.block_sums(src, 5).cpu().numpy()
    assert np.max(np.abs(sums - w_ref.reshape(-1, 32).sum(axis=1))) < 1e-15
    rng = random.Random(5)
    for _ in range(40):
        u = rng.random()
        assert sv.sample(u, src) == so.sample_index(w, u)
    assert sv.sample(0.0, src) == so.sample_index(w, 0.0)
    assert sv.sample(0.9999999999999999, src) == so.sample_index(w, 0.9999999999999999)
    # the exact fixed-point descent (used above 2^20 outcomes and on sharded registers) forced at this size: identical
    # to the brute-force bisect over exact cumulative sums of the DEVICE's weights, and to the float rule away from ties
    import helpers
    w_phys = sv.probs(None).cpu().numpy()
    engine_mod = engine
    old = engine_mod.SCAN_BITS_EXACT
    engine_mod.SCAN_BITS_EXACT = 4
    try:
        for _ in range(40):
            u = rng.random()
            got = sv.sample(u, src)
            assert got == helpers.exact_draw_bruteforce(w_phys, n, src, u)
            assert got == so.sample_index(w, u)
        for u in (0.0, 0.9999999999999999):
            assert sv.sample(u, src) == helpers.exact_draw_bruteforce(w_phys, n, src, u)
    finally:
        engine_mod.SCAN_BITS_EXACT = old


@pytest.mark.parametrize("n,kind", [(14, "identity"), (15, "reversed"), (16, "shuffled"), (16, "sparse")])
def test_exact_sampler_descent_bit_exact(qsv_gpu, n, kind):
    """Every level shape of the radix descent (high-position levels, levels with one or two bucket bits inside a run,
    leaf) against the integer brute force: the outcome index is bit-exact by construction."""
    lib, engine, lowering, torch = qsv_gpu
    import helpers
    sv = engine.StateVector(n)
    psi = _rand_state(n, 31 + n)
    if kind == "sparse":
        keep = np.random.default_rng(9).random(1 << n) < 0.002
        psi = np.where(keep, psi, 0)
        psi /= np.linalg.norm(psi)
    sv.state.copy_(torch.from_numpy(psi))
    src = {"identity": list(range(n)), "reversed": list(range(n - 1, -1, -1))}.get(kind)
    if src is None:
        src = [int(v) for v in np.random.default_rng(n).permutation(n)]
    w_phys = sv.probs(None).cpu().numpy()
    old = engine.SCAN_BITS_EXACT
    engine.SCAN_BITS_EXACT = 4
    rng = random.Random(n)
    try:
        for u in [0.0, 1.0 - 2.0 ** -53] + [rng.random() for _ in range(12)]:
            assert sv.sample(u, src) == helpers.exact_draw_bruteforce(w_phys, n, src, u)
    finally:
        engine.SCAN_BITS_EXACT = old
    sv.state.zero_()
    engine.SCAN_BITS_EXACT = 4
    try:
        with pytest.raises(ValueError):
            sv.sample(0.5, src)
    finally:
        engine.SCAN_BITS_EXACT = old


def test_exact_sampler_at_22_qubits_matches_vectorised_integer_rule(qsv_gpu):
    """Above 2^20 outcomes the sampler is the default path: 22 qubits, shuffled output wiring, against exact integer
    cumulative sums built with numpy object arrays."""
    lib, engine, lowering, torch = qsv_gpu
    from qsv.sampling import u_fixed
    n = 22
    sv = engine.StateVector(n)
    sv.state.copy_(torch.from_numpy(_rand_state(n, 5)))
    src = [int(v) for v in np.random.default_rng(22).permutation(n)]
    w_out = sv.probs(src).cpu().numpy()
    mant, ex = np.frexp(w_out)
    m = (mant * float(1 << 53)).astype(np.uint64).astype(object)
    sh = ex.astype(np.int64) - 53 + 96
    q = np.array([int(a) << int(b) if b >= 0 else int(a) >> int(-b) for a, b in zip(m, sh)], dtype=object)
    cum = np.cumsum(q)
    total = int(cum[-1])
    rng = random.Random(1)
    for u in [0.0] + [rng.random() for _ in range(6)]:
        mu, su = u_fixed(u)
        target = (total * mu) >> su
        lo, hi = 0, (1 << n) - 1
        while lo < hi:                       # first index with cum > target
            mid = (lo + hi) // 2
            if cum[mid] > target:
                hi = mid
            else:
                lo = mid + 1
        assert sv.sample(u, src) == lo


def test_grover_kernels_closed_form_25q(qsv_gpu):
    """Oracle + fused diffusion at a size no dense operator could exist: amplitude of the marked pair
    follows sin((2K+1) theta)/sqrt(2), theta = asin(2^(-n/2))."""
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    n, K, x = 24, 12, 0x5A5A5A
    c = workloads.grover_circuit(n, x, K)
    c.sort()
    ops, src_pos = lowering.lower(c)
    sv = engine.StateVector(n + 1)
    sv.init_basis(1)
    sv.run_ops(ops)
    pair = sv.state[2 * x: 2 * x + 2].cpu().numpy()
    theta = math.asin(2.0 ** (-n / 2))
    expect = math.sin((2 * K + 1) * theta) / math.sqrt(2)
    assert abs(pair[0] - expect) < 1e-12 and abs(pair[1] + expect) < 1e-12
    total = float(sv.block_sums(None, 12).sum())
    assert abs(total - 1.0) < 1e-10


def test_round_trip_and_norm_at_full_size(qsv_gpu):
    """BASELINE config-3 size (30 qubits, 17 GB): circuit followed by its inverse returns |0...0>
    exactly enough, and the norm is preserved — size-independent properties where the oracle cannot go."""
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    from qsv.schedule import Op
    n = 30
    c = workloads.random_layered_circuit(n, 2)
    c.sort()
    ops, _ = lowering.lower(c)
    inv = []
    for op in reversed(ops):
        if op.kind == "dense":
            inv.append(Op("dense", op.targets, op.controls, cvals=op.cvals, matrix=op.matrix.conj().T))
        elif op.kind == "phase":
            inv.append(Op("phase", op.targets, op.controls, cvals=op.cvals, scalar=op.scalar.conjugate()))
        else:
            assert op.kind == "mcx"
            inv.append(op)
    sv = engine.StateVector(n)
    sv.init_basis(0)
    sv.run_ops(ops)
    norm = float(sv.block_sums(None, 12).sum())
    assert abs(norm - 1.0) < 1e-10
    sv.run_ops(inv)
    a0 = complex(sv.state[0].item())
    assert abs(a0 - 1.0) < 1e-10
    assert abs(float(sv.block_sums(None, 12).sum()) - 1.0) < 1e-10
    assert sv.sample(0.37) == 0


def test_22q_layered_circuit_against_oracle(qsv_gpu):
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    n = 22
    c = workloads.random_layered_circuit(n, 2, seed=5)
    c.sort()
    ops, _ = lowering.lower(c)
    sv = engine.StateVector(n)
    sv.init_basis(3)
    sv.run_ops(ops)
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[3] = 1
    for M, wires, _ in so.random_layered_steps(n, 2, seed=5):
        ref = so.apply_gate(ref, n, wires, M)
    assert np.max(np.abs(sv.amplitudes() - ref)) <= TOL


# ---- pass-specialised kernels (csrc/qsv_jit.cu): forced on at sizes the oracle can follow ---------

def _jit_counts(lib):
    n_c, n_h, secs = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
    lib.qsv_jit_stats(C.byref(n_c), C.byref(n_h), C.byref(secs))
    return int(n_c.value), int(n_h.value)


@pytest.mark.parametrize("workload", ["layered16", "qft15", "layered22"])
def test_specialised_kernels_match_oracle_and_interpreter(qsv_gpu, monkeypatch, workload):
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    if workload == "qft15":
        n = 15
        c = workloads.qft_ladder_circuit(n)
        in_v = [(i + 1) % 2 for i in range(n)]
        steps = None
    else:
        n = int(workload[7:])
        depth = 6 if n == 16 else 2
        c = workloads.random_layered_circuit(n, depth, seed=21)
        in_v = [0] * n
        steps = so.random_layered_steps(n, depth, seed=21)
    c.sort()
    ops, src_pos = lowering.lower(c)
    idx = lowering.basis_index(in_v)
    out = {}
    for mode in ("0", "force"):
        monkeypatch.setenv("QSV_JIT", mode)
        before = _jit_counts(lib)
        sv = engine.StateVector(n)
        sv.init_basis(idx)
        sv.run_ops(ops)
        out[mode] = sv.amplitudes()
        after = _jit_counts(lib)
        if mode == "force":
            assert after[0] + after[1] > before[0] + before[1], "the specialised path did not run"
        else:
            assert after == before
    assert np.max(np.abs(out["force"] - out["0"])) <= 1e-13
    if steps is not None:
        ref = so.initial_state(in_v)
        for M, wires, _ in steps:
            ref = so.apply_gate(ref, n, wires, M)
        assert np.max(np.abs(out["force"] - ref)) <= TOL
    else:
        # QFT of a basis state: every outcome has probability 2^-n (the interpreter comparison above
        # pins the phases)
        w = np.abs(out["force"]) ** 2
        assert np.max(np.abs(w - 2.0 ** -n)) <= TOL


def test_specialised_kernels_external_conditions(qsv_gpu, monkeypatch):
    """A shard of a larger register: controls and diagonal factors above n_local come from rank_bits."""
    lib, engine, lowering, torch = qsv_gpu
    from qsv.schedule import Op
    from helpers import emulate_ops
    n_local = 14
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
    ph = complex(math.cos(0.3), math.sin(0.3))
    ops = [Op("dense", (3,), (), matrix=H), Op("mcx", (2,), (15,)), Op("phase", (), (14, 3), scalar=ph),
           Op("phase", (), (15, 12), scalar=-1.0 + 0j), Op("dense", (11,), (), matrix=H), Op("mcx", (7,), (11, 14)),
           Op("phase", (), (14,), scalar=ph), Op("dense", (0,), (14,), matrix=H), Op("mcx", (13,), (1,)),
           Op("diag", (2, 14, 9), (), matrix=np.exp(1j * np.arange(8)))]
    monkeypatch.setenv("QSV_JIT", "force")
    rng = np.random.default_rng(4)
    psi = rng.standard_normal(1 << n_local) + 1j * rng.standard_normal(1 << n_local)
    psi /= np.linalg.norm(psi)
    for rank in range(4):
        sv = engine.StateVector(n_local)
        sv.rank = rank              # pretend to be shard `rank` of a 16-qubit register
        sv.state.copy_(torch.from_numpy(psi).to(sv.device))
        sv.run_ops(ops)
        ref = emulate_ops(ops, psi, n_local, rank_bits=rank << n_local)
        assert np.max(np.abs(sv.amplitudes() - ref)) <= 1e-13


def test_specialised_kernels_full_size_against_interpreter(qsv_gpu, monkeypatch):
    """30 qubits (tile bits up to position 29, 2^18 tiles): the specialised kernels and the interpreting
    kernel must produce the same 17 GB state (the oracle cannot follow at this size)."""
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    n = 30
    c = workloads.random_layered_circuit(n, 3, seed=77)
    c.sort()
    ops, _ = lowering.lower(c)
    monkeypatch.setenv("QSV_JIT", "0")
    sv = engine.StateVector(n)
    sv.init_basis(5)
    sv.run_ops(ops)
    sv.synchronize()
    ref = sv.state.clone()
    monkeypatch.setenv("QSV_JIT", "force")
    before = _jit_counts(lib)
    sv.init_basis(5)
    sv.run_ops(ops)
    sv.synchronize()
    after = _jit_counts(lib)
    assert after[0] + after[1] > before[0] + before[1]
    d = torch.view_as_real(sv.state) - torch.view_as_real(ref)
    assert float(d.abs().max()) <= 1e-13
    del d, ref
    assert abs(float(sv.block_sums(None, 12).sum()) - 1.0) < 1e-10


def test_support_hint_gives_the_same_state_as_full_sweeps(qsv_gpu, monkeypatch):
    """A register that starts in a basis state skips the tiles that are still zero (schedule.Support ->
    qsv_run_pass_rounds_sparse).  Bit-for-bit the same state as with the hint switched off, at a size where the
    first passes touch a small fraction of the tiles; the hinted run must launch the same kernels."""
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    monkeypatch.setenv("QSV_JIT", "force")
    n = 24
    c = workloads.random_layered_circuit(n, 8, seed=31)
    c.sort()
    ops, _, flip = lowering.lower(c, in_flip=True)
    idx = 0b1011 ^ flip
    sv = engine.StateVector(n)
    plan = sv.prepare(ops)
    hints = [sp for kind, seg, sp, zf in sv._walk(plan, idx) if kind == "pass"]
    assert sum(1 for sp in hints if sp[0]) >= 2, "expected the first passes to carry a support hint"
    monkeypatch.setenv("QSV_SPARSE_START", "0")
    sv.init_basis(idx)
    sv.execute(plan)
    sv.synchronize()
    ref = sv.state.clone()
    monkeypatch.delenv("QSV_SPARSE_START")
    sv.init_basis(idx)
    sv.execute(plan)
    sv.synchronize()
    assert torch.equal(torch.view_as_real(sv.state), torch.view_as_real(ref))
    assert abs(float(sv.block_sums(None, 12).sum()) - 1.0) < 1e-10
    # a state of unknown support (no init_basis before execute) is swept completely
    sv.execute(plan)
    assert sv._basis is None


def test_run_from_basis_without_the_clear_is_bit_identical(qsv_gpu, monkeypatch):
    """StateVector.run_from_basis skips the 16 * 2^n byte clear when the plan's leading passes can zero-fill
    (qsv_init_basis_sparse + qsv_run_pass_rounds_fresh).  On a register deliberately filled with NaN the final state
    must equal init_basis + execute bit for bit; with QSV_LAZY_INIT=0, and for plans that do not qualify, it IS
    init_basis + execute."""
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    monkeypatch.setenv("QSV_JIT", "force")
    for n, depth, seed in ((20, 10, 41), (24, 8, 31)):
        c = workloads.random_layered_circuit(n, depth, seed=seed)
        c.sort()
        ops, _, flip = lowering.lower(c, in_flip=True)
        idx = 0b1101 ^ flip
        sv = engine.StateVector(n)
        plan = sv.prepare(ops)
        sv.init_basis(idx)
        sv.execute(plan)
        sv.synchronize()
        ref = sv.state.clone()
        n_fresh = sv._fresh_prefix(plan)
        assert n_fresh >= 2, "expected the plan to start with zero-fill passes"
        launches = lib.qsv_launch_count()
        torch.view_as_real(sv.state).fill_(float("nan"))
        sv.run_from_basis(idx, plan)
        sv.synchronize()
        assert torch.equal(torch.view_as_real(sv.state), torch.view_as_real(ref))
        assert not torch.isnan(torch.view_as_real(sv.state)).any()
        monkeypatch.setenv("QSV_LAZY_INIT", "0")
        assert sv._fresh_prefix(plan) == 0
        torch.view_as_real(sv.state).fill_(float("nan"))
        sv.run_from_basis(idx, plan)
        sv.synchronize()
        assert torch.equal(torch.view_as_real(sv.state), torch.view_as_real(ref))
        monkeypatch.delenv("QSV_LAZY_INIT")
    # a shallow circuit never sweeps the whole register: the plan does not qualify and the register is cleared
    c = workloads.random_layered_circuit(20, 1, seed=3)
    c.sort()
    ops, _, flip = lowering.lower(c, in_flip=True)
    sv = engine.StateVector(20)
    plan = sv.prepare(ops)
    torch.view_as_real(sv.state).fill_(float("nan"))
    sv.run_from_basis(flip, plan)
    sv.synchronize()
    assert not torch.isnan(torch.view_as_real(sv.state)).any()
    assert abs(float(sv.block_sums(None, 12).sum()) - 1.0) < 1e-10


@pytest.mark.parametrize("static_form,by_cost", [("1", "1"), ("1", "0"), ("0", "0")])
def test_last_pass_weighs_the_state_for_the_sampler(qsv_gpu, monkeypatch, static_form, by_cost):
    """execute(plan, sample_src_pos=...) lets the last pass add up the sampler's first-level bucket sums
    (qsv_run_pass_rounds_sums).  The draws must equal the ones of the plain path (qsv_sample_level over the whole state)
    for every u, with identity and permuted output wiring; the state itself is bit-identical; a sample with another
    wiring, or after the state changed, does not use the stale sums."""
    lib, engine, lowering, torch = qsv_gpu
    from qsv import workloads
    monkeypatch.setenv("QSV_JIT", "force")
    # "1": buckets from generation-time constants where the last round's stores allow it (qsv_tile_sum_s, up to three
    # bucket bits inside the tile); "0": every amplitude's bucket from its final index (at most two inside)
    monkeypatch.setenv("QSV_SUM_STATIC", static_form)
    # by_cost "1": the engine weighs the cost of bucket bits inside the tile against the sweep they save (default);
    # "0": as many top output bits as the kernel can take (up to three inside the tile)
    monkeypatch.setenv("QSV_PRESUM_COST", by_cost)
    us = [0.0, 0.123, 0.5, 0.87654321, 1.0 - 2.0 ** -53]
    inside_seen, cases_used = set(), 0
    for n, seed in ((23, 17), (23, 18), (23, 19), (24, 31), (22, 5)) + tuple((22, s_) for s_ in range(40, 52)):
        c = workloads.random_layered_circuit(n, 8, seed=seed)
        c.sort()
        ops, src_pos, flip = lowering.lower(c, in_flip=True)
        sv = engine.StateVector(n)
        plan = sv.prepare(ops)
        sv.run_from_basis(flip, plan)
        sv.synchronize()
        ref_state = sv.state.clone()
        plain, totals = [], []
        for u in us:
            plain.append(sv.sample(u, src_pos))
            totals.append(sv._sampler.sel.cpu().tolist()[4:6])
        used = 0
        for u, want, tot in zip(us, plain, totals):
            sv.run_from_basis(flip, plan, sample_src_pos=src_pos)
            used += sv._presum is not None
            k = sv._presum[0] if sv._presum is not None else 0
            assert sv.sample(u, src_pos) == want
            assert sv._sampler.sel.cpu().tolist()[4:6] == tot          # the exact 128-bit total weight
            assert sv._presum is None
        assert used in (0, len(us))             # (0: the plan ends in a 1-2 op pass of the shared-memory kernel)
        cases_used += used > 0
        assert torch.equal(torch.view_as_real(sv.state), torch.view_as_real(ref_state))
        last = [seg for kind, seg, sp, zf in sv._walk(plan, flip) if kind == "pass"][-1][0][0]
        tile = set(range(last.low_bits)) | set(last.hi_pos)
        if used:
            inside_seen.add(sum(1 for b in range(n - k, n) if src_pos[b] in tile))
    assert cases_used >= 6, "the last pass was expected to carry the sums in at least six of the circuits"
    assert max(inside_seen) >= (2 if by_cost == "0" else 1), "no case with bucket bits inside the last pass's tile"
    n, seed = 23, 17
    c = workloads.random_layered_circuit(n, 8, seed=seed)
    c.sort()
    ops, src_pos, flip = lowering.lower(c, in_flip=True)
    sv = engine.StateVector(n)
    plan = sv.prepare(ops)
    # a different output wiring: the sums on the device do not apply, the sampler sweeps the state itself
    other = list(src_pos)
    other[0], other[n - 1] = other[n - 1], other[0]
    sv.run_from_basis(flip, plan, sample_src_pos=src_pos)
    z_other = sv.sample(0.5, other)
    sv.run_from_basis(flip, plan)
    assert sv.sample(0.5, other) == z_other
    # state changed after the sums were taken: they are dropped
    sv.run_from_basis(flip, plan, sample_src_pos=src_pos)
    sv.execute(plan)
    assert sv._presum is None


def test_support_hint_is_dropped_when_the_state_changes_outside_execute(qsv_gpu, monkeypatch):
    """init_basis -> run_program(H) -> run_ops(ops): the register is no longer the basis state when execute() runs,
    so the passes must sweep everything (a stale hint would skip non-zero tiles silently)."""
    lib, engine, lowering, torch = qsv_gpu
    import numpy as np
    from qsv import workloads
    from qsv.schedule import Op
    monkeypatch.setenv("QSV_JIT", "force")
    n = 20
    c = workloads.random_layered_circuit(n, 6, seed=9)
    c.sort()
    ops, _, flip = lowering.lower(c, in_flip=True)
    s_ = 2 ** -0.5
    pre = [Op("dense", (n - 1,), (), matrix=np.array([[s_, s_], [s_, -s_]], dtype=np.complex128)),
           Op("dense", (n - 3,), (), matrix=np.array([[s_, s_], [s_, -s_]], dtype=np.complex128))]
    sv = engine.StateVector(n)
    h = sv.upload_program(pre)
    sv.init_basis(flip)
    sv.run_program(h)
    assert sv._basis is None
    sv.run_ops(ops)
    sv.synchronize()
    got = sv.state.clone()
    monkeypatch.setenv("QSV_SPARSE_START", "0")
    sv.init_basis(flip)
    sv.run_program(h)
    sv.run_ops(ops)
    sv.synchronize()
    assert torch.equal(torch.view_as_real(got), torch.view_as_real(sv.state))
    assert abs(float(sv.block_sums(None, 12).sum()) - 1.0) < 1e-10
