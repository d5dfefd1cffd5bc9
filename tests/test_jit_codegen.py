"""The pass-specialised kernel generator (csrc/qsv_jit.cu) without a GPU: its output is compiled as host
C++ (-DQSV_JIT_HOST) and must reproduce the numpy emulation of the interpreting kernel on the same packed
round programs; NVRTC must accept the same source for sm_100a."""
import ctypes as C
import math
import random

import numpy as np
import pytest

from helpers import emulate_round_program, jit_cpass, jit_host_run, jit_source

TOL = 1e-14      # same arithmetic up to FMA contraction and the deferred 2^-1/2 factors


@pytest.fixture(autouse=True)
def _programs_for_the_specialised_kernel(monkeypatch):
    """Pack the programs the way large states get them (in-register permutations, big programs): the planner
    asks the library which kernel will run a pass, and QSV_JIT=force answers 'the specialised one'."""
    monkeypatch.setenv("QSV_JIT", "force")


def _rand_state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return psi / np.linalg.norm(psi)


def _round_programs(ops, n_local):
    from qsv import schedule
    packed = schedule.pack_passes(schedule.plan_passes(ops, n_local, 12, 6))
    blob = packed.blob.tobytes()
    return [(ps, extra[0], blob[off:]) for ps, off, n_ops, extra in packed.passes if n_ops < 0]


def _check(ops, n_local, tmp_path, rank_bits=0, seed=1):
    progs = _round_programs(ops, n_local)
    assert progs, "expected at least one register-resident pass"
    psi = _rand_state(n_local, seed)
    stats = {"general": 0, "rounds": 0}
    for ps, prog, tables in progs:
        ref = emulate_round_program(prog, tables, ps, psi, rank_bits)
        got = jit_host_run(ps, prog, tables, psi, rank_bits, str(tmp_path))
        assert np.max(np.abs(got - ref)) <= TOL
        from qsv.schedule import ROUND_DTYPE
        nr = int(np.frombuffer(bytes(prog), dtype="<u4", count=1)[0])
        rounds = np.frombuffer(bytes(prog), dtype=ROUND_DTYPE, count=nr, offset=16)
        stats["rounds"] += nr
        stats["general"] += sum(1 for R in rounds if (int(R["n_pre"]) | int(R["n_post"])) & 0x8000)
        psi = ref
    return stats


@pytest.mark.parametrize("pad", ["default", "0", "all"])
def test_layered_circuit_matches_interpreter_emulation(tmp_path, monkeypatch, pad):
    """pad: shared-memory layout of the tile - padded runs + per-round thread-bit mapping where a round keeps
    tile bits 0..2 in registers (default), never ("0"), or in every pass ("all")."""
    from qsv import workloads, lowering
    if pad != "default":
        monkeypatch.setenv("QSV_JIT_PAD", pad)
    n = 14
    c = workloads.random_layered_circuit(n, 6, seed=3)
    c.sort()
    ops, _ = lowering.lower(c)
    monkeypatch.setenv("QSV_REG_MCX", "all")
    st = _check(ops, n, tmp_path)
    assert st["rounds"] >= 4
    from qsv.schedule import ROP_DTYPE, ROP_MCX_REG
    kinds = set()
    for ps, prog, _ in _round_programs(ops, n):
        hdr = np.frombuffer(bytes(prog), dtype="<u4", count=4)
        rops = np.frombuffer(bytes(prog), dtype=ROP_DTYPE, count=int(hdr[2]), offset=16 + 16 * int(hdr[0]))
        kinds |= set(int(k) for k in rops["kind"])
    assert ROP_MCX_REG in kinds, "the layered circuit should produce in-register CNots"


def test_qft_ladder_controlled_phases(tmp_path):
    from qsv import workloads, lowering
    n = 13
    c = workloads.qft_ladder_circuit(n)
    c.sort()
    ops, _ = lowering.lower(c)
    _check(ops, n, tmp_path, seed=2)


@pytest.mark.parametrize("workload", ["layered", "toffoli"])
def test_bank_search_mapping_keeps_results_identical(tmp_path, monkeypatch, workload):
    """QSV_JIT_BANKSEARCH (default on, 0 = static rule): the low thread-index bits of every round are chosen by scoring the actual slot
    addresses, permutations included.  Any choice is a relabelling of the threads: the state must come out exactly
    as with the default mapping, and the generated source must differ (the search did pick other bits)."""
    from qsv import workloads, lowering
    from qsv.schedule import Op
    if workload == "layered":
        n = 14
        c = workloads.random_layered_circuit(n, 6, seed=3)
        c.sort()
        ops, _ = lowering.lower(c)
    else:
        monkeypatch.setenv("QSV_REG_MCX", "0")
        n = 13
        rng = random.Random(7)
        H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
        ops = []
        for _ in range(50):
            a, b, c_ = rng.sample(range(n), 3)
            ops.append(rng.choice([Op("mcx", (a,), (b, c_)), Op("mcx", (a,), (b,)), Op("dense", (a,), (), matrix=H),
                                   Op("phase", (), (a, b), scalar=1j)]))
    progs = _round_programs(ops, n)
    psi = _rand_state(n, 9)
    differs = False
    for ps, prog, tables in progs:
        monkeypatch.setenv("QSV_JIT_BANKSEARCH", "0")
        src0 = jit_source(ps, prog)
        ref = jit_host_run(ps, prog, tables, psi, 0, str(tmp_path))
        monkeypatch.setenv("QSV_JIT_BANKSEARCH", "1")
        src1 = jit_source(ps, prog)
        got = jit_host_run(ps, prog, tables, psi, 0, str(tmp_path))
        differs = differs or src0 != src1
        assert np.array_equal(got, ref)
        psi = ref
    assert differs or workload == "layered", "the search never changed a mapping"


@pytest.mark.parametrize("reg_mcx", ["1", "all", "0"])
def test_toffoli_heavy_circuit_general_permutations(tmp_path, monkeypatch, reg_mcx):
    """Toffoli / CNot chains: as in-register permutations (slot renaming, predicated exchanges), and - with
    those switched off - as address permutations whose conditions read several register-dependent bits, which
    leave the XOR-linear form: the generator must emit the per-slot general form for those lists."""
    from qsv.schedule import Op
    monkeypatch.setenv("QSV_REG_MCX", reg_mcx)
    n = 13
    rng = random.Random(7)
    ops = []
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
    Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    S = np.array([[0.5 + 0.5j, 0.5 - 0.5j], [0.5 - 0.5j, 0.5 + 0.5j]], dtype=np.complex128)
    for _ in range(60):
        r = rng.randrange(6)
        a, b, c = rng.sample(range(n), 3)
        if r == 0:
            ops.append(Op("mcx", (a,), (b, c)))
        elif r == 1:
            ops.append(Op("mcx", (a,), (b,), cvals=(rng.randrange(2),)))
        elif r == 2:
            ops.append(Op("dense", (a,), (), matrix=rng.choice([H, Y, S])))
        elif r == 3:
            ops.append(Op("phase", (), (a, b), cvals=(1, rng.randrange(2)), scalar=complex(math.cos(1.1), math.sin(1.1))))
        elif r == 4:
            ops.append(Op("phase", (), (a,), scalar=-1.0 + 0j))
        else:
            ops.append(Op("dense", (a,), (b,), matrix=H))
    st = _check(ops, n, tmp_path, seed=5)
    if reg_mcx == "0":
        assert st["general"] >= 1, "the circuit was meant to exercise the general permutation form"


def test_external_conditions_read_rank_bits(tmp_path):
    """Controls / diagonal factors on positions outside the shard (multi-GPU) come from rank_bits."""
    from qsv.schedule import Op
    n_local = 13
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
    ph = complex(math.cos(0.3), math.sin(0.3))
    ops = [Op("dense", (3,), (), matrix=H), Op("mcx", (2,), (14,)), Op("phase", (), (13, 3), scalar=ph),
           Op("phase", (), (14, 12), scalar=-1.0 + 0j), Op("dense", (11,), (), matrix=H), Op("mcx", (7,), (11, 13)),
           Op("phase", (), (13,), scalar=ph), Op("dense", (0,), (13,), matrix=H),
           Op("diag", (2, 13, 9), (), matrix=np.exp(1j * np.arange(8)))]
    for rank in range(4):
        _check(ops, n_local, tmp_path, rank_bits=rank << n_local, seed=rank)


def test_pass_whose_ops_are_all_conditional_skips_tiles(tmp_path):
    from qsv.schedule import Op
    n_local = 14
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
    ops = [Op("dense", (1,), (13,), matrix=H), Op("mcx", (2,), (13,)), Op("phase", (), (13, 4), scalar=1j),
           Op("dense", (5,), (13,), matrix=H)]
    _check(ops, n_local, tmp_path)


def test_nvrtc_accepts_generated_source_for_sm100a():
    from qsv import _cabi, workloads, lowering
    lib = _cabi.load()
    n = 14
    c = workloads.random_layered_circuit(n, 3, seed=9)
    c.sort()
    ops, _ = lowering.lower(c)
    ps, prog, _ = _round_programs(ops, n)[0]
    src = jit_source(ps, prog)
    assert "qsv_jit_pass" in src and "cp.async.bulk" in src
    cp = jit_cpass(ps)
    need = C.c_size_t(0)
    rc = lib.qsv_jit_compile(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), None, 0, C.byref(need))
    if rc != 0 and b"cannot load libnvrtc" in lib.qsv_last_error():
        pytest.skip("NVRTC is not installed here")
    assert rc == 0, lib.qsv_last_error()
    cubin = C.create_string_buffer(need.value)
    assert lib.qsv_jit_compile(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), cubin, need.value,
                               C.byref(need)) == 0
    assert cubin.raw[:4] == b"\x7fELF"
    n_c, n_h, secs = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
    lib.qsv_jit_stats(C.byref(n_c), C.byref(n_h), C.byref(secs))
    assert n_c.value >= 1 and n_h.value >= 1 and secs.value > 0


def test_policy_small_states_stay_on_the_interpreter(monkeypatch):
    from qsv import _cabi, workloads, lowering
    lib = _cabi.load()
    c = workloads.random_layered_circuit(14, 2, seed=9)
    c.sort()
    ops, _ = lowering.lower(c)
    ps, prog, _ = _round_programs(ops, 14)[0]
    cp = jit_cpass(ps)
    args = (C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size))
    monkeypatch.delenv("QSV_JIT", raising=False)
    assert lib.qsv_jit_wanted(*args) == 0
    cp.n_local = 30
    assert lib.qsv_jit_wanted(*args) == 1
    monkeypatch.setenv("QSV_JIT", "0")
    assert lib.qsv_jit_wanted(*args) == 0
    cp.n_local = 14
    monkeypatch.setenv("QSV_JIT", "force")
    assert lib.qsv_jit_wanted(*args) == 1


def test_long_programs_capacity_and_splitting(tmp_path, monkeypatch):
    """A pass with more ops than the interpreting kernel's parameter block holds (120) is one program for
    the specialised kernel and is split into two sweeps for the interpreter; both give the same state."""
    from qsv import workloads, lowering, schedule
    from helpers import emulate_program
    n = 15
    c = workloads.qft_ladder_circuit(n)
    c.sort()
    ops, _ = lowering.lower(c)
    passes = schedule.plan_passes(ops, n, 12, 6)
    psi = _rand_state(n, 11)
    monkeypatch.setattr(schedule, "RPROG_MAX_OPS", 40)      # stand-in for 120 at a size the emulation follows

    monkeypatch.setenv("QSV_JIT", "force")
    packed = schedule.pack_passes(passes)
    sizes = [int(np.frombuffer(bytes(m[3][0]), dtype="<u4", count=3)[2]) for m in packed.passes if m[2] < 0]
    assert max(sizes) > schedule.RPROG_MAX_OPS, sizes
    assert len(packed.passes) == len(passes)
    ref = emulate_program(packed, psi)
    blob = packed.blob.tobytes()
    got = psi
    for ps, off, n_ops, extra in packed.passes:
        assert n_ops < 0
        got = jit_host_run(ps, extra[0], blob[off:], got, 0, str(tmp_path))
    assert np.max(np.abs(got - ref)) <= 1e-13

    monkeypatch.setenv("QSV_JIT", "0")
    split = schedule.pack_passes(passes)
    assert len(split.passes) > len(passes)
    for m in split.passes:
        assert m[2] < 0 and int(np.frombuffer(bytes(m[3][0]), dtype="<u4", count=3)[2]) <= schedule.RPROG_MAX_OPS
    assert np.max(np.abs(emulate_program(split, psi) - ref)) <= 1e-13


@pytest.mark.parametrize("g,positions", [(1, [13]), (2, [9, 14]), (3, [12, 13, 14])])
def test_fused_swap_pass_matches_swap_then_pass(tmp_path, g, positions):
    """The fused variant (qsv_run_pass_rounds_swap) on the host: every rank pulls its tiles from the shards before
    the swap and applies the pass on the way.  Reference: the swap as an index exchange in numpy, then the unfused
    generated pass on every shard.  Ops read the swapped and the global positions as controls / phases."""
    from helpers import jit_host_run_swap
    from qsv.schedule import Op, plan_schedule, pack_passes
    n_local = 15
    P = 1 << g
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
    S = np.array([[0.5 + 0.5j, 0.5 - 0.5j], [0.5 - 0.5j, 0.5 + 0.5j]], dtype=np.complex128)
    ph = complex(math.cos(0.7), math.sin(0.7))
    free = [p for p in range(6, n_local) if p not in positions]
    swap = Op("global_swap", (), (), params={"n_swap": g, "positions": positions})
    ops = [swap,
           Op("dense", (0,), (), matrix=H), Op("dense", (free[0],), (), matrix=S), Op("mcx", (free[1],), (positions[0],)),
           Op("phase", (), (positions[-1], 3), scalar=ph), Op("phase", (), (n_local, free[2]), scalar=-1.0 + 0j),
           Op("dense", (free[2],), (n_local + g - 1,), matrix=H), Op("mcx", (2,), (free[0], positions[0])),
           Op("dense", (free[3],), (), matrix=H), Op("phase", (), (free[3],), scalar=ph), Op("dense", (1,), (), matrix=S),
           Op("diag", (2, positions[0], n_local), (), matrix=np.exp(1j * np.arange(8))), Op("dense", (4,), (), matrix=H)]
    sched = plan_schedule(ops, n_local, 12, 6, fuse_swaps=True)
    assert sched[0][0] == "swap_pass" and sched[0][1][2] == 0, [k for k, _ in sched]
    ps = sched[0][1][1]
    assert not set(ps.hi_pos) & set(positions)
    assert len(ps.ops) >= 8
    packed = pack_passes([ps])
    ps_, off, n_ops, extra = packed.passes[0]
    assert n_ops < 0
    prog, tables = extra[0], packed.blob.tobytes()[off:]
    shards = [_rand_state(n_local, 40 + r) for r in range(P)]
    got = jit_host_run_swap(ps_, prog, tables, shards, positions, n_local, str(tmp_path))
    # reference: exchange (field = j, rest) of rank r with (field = r, rest) of rank j, then the plain pass
    idx = np.arange(1 << n_local)
    field = np.zeros_like(idx)
    for k, q in enumerate(positions):
        field |= ((idx >> q) & 1) << k
    fmask = sum(1 << q for q in positions)
    for r in range(P):
        swapped = np.empty(1 << n_local, dtype=np.complex128)
        for j in range(P):
            sel = field == j
            src_idx = (idx[sel] & ~fmask) | sum(((r >> k) & 1) << q for k, q in enumerate(positions))
            swapped[sel] = shards[j][src_idx]
        ref = jit_host_run(ps_, prog, tables, swapped, r << n_local, str(tmp_path))
        assert np.max(np.abs(got[r] - ref)) == 0.0, "rank %d" % r


@pytest.mark.parametrize("g,positions", [(1, [12]), (3, [10, 13, 14])])
def test_fused_pass_then_swap_matches_pass_then_swap(tmp_path, g, positions):
    """'Pass first' form: the last pass before a swap runs inside the swap kernel, on the tiles as they are pulled
    from the peers; its ops see the coordinates (rank, swapped field) the tile had on the source rank."""
    from helpers import jit_host_run_swap
    from qsv.schedule import Op, plan_schedule, pack_passes
    n_local = 15
    P = 1 << g
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
    S = np.array([[0.5 + 0.5j, 0.5 - 0.5j], [0.5 - 0.5j, 0.5 + 0.5j]], dtype=np.complex128)
    ph = complex(math.cos(0.7), math.sin(0.7))
    free = [p for p in range(6, n_local) if p not in positions]
    swap = Op("global_swap", (), (), params={"n_swap": g, "positions": positions})
    ops = [Op("dense", (0,), (), matrix=H), Op("dense", (free[0],), (), matrix=S), Op("mcx", (free[1],), (positions[0],)),
           Op("phase", (), (positions[-1], 3), scalar=ph), Op("phase", (), (n_local, free[2]), scalar=-1.0 + 0j),
           Op("dense", (free[2],), (n_local + g - 1,), matrix=H), Op("mcx", (2,), (free[0], positions[0])),
           Op("dense", (free[3],), (), matrix=H), Op("phase", (), (free[3],), scalar=ph), Op("dense", (1,), (), matrix=S),
           Op("diag", (2, positions[0], n_local), (), matrix=np.exp(1j * np.arange(8))), Op("dense", (4,), (), matrix=H),
           swap]
    sched = plan_schedule(ops, n_local, 12, 6, fuse_swaps=True)
    assert [k for k, _ in sched] == ["swap_pass"] and sched[0][1][2] == 1, [k for k, _ in sched]
    ps = sched[0][1][1]
    packed = pack_passes([ps])
    ps_, off, n_ops, extra = packed.passes[0]
    prog, tables = extra[0], packed.blob.tobytes()[off:]
    shards = [_rand_state(n_local, 60 + r) for r in range(P)]
    got = jit_host_run_swap(ps_, prog, tables, shards, positions, n_local, str(tmp_path), pass_first=1)
    after_pass = [jit_host_run(ps_, prog, tables, shards[r], r << n_local, str(tmp_path)) for r in range(P)]
    idx = np.arange(1 << n_local)
    field = np.zeros_like(idx)
    for k, q in enumerate(positions):
        field |= ((idx >> q) & 1) << k
    fmask = sum(1 << q for q in positions)
    for r in range(P):
        ref = np.empty(1 << n_local, dtype=np.complex128)
        for j in range(P):
            sel = field == j
            src_idx = (idx[sel] & ~fmask) | sum(((r >> k) & 1) << q for k, q in enumerate(positions))
            ref[sel] = after_pass[j][src_idx]
        assert np.max(np.abs(got[r] - ref)) == 0.0, "rank %d" % r


def test_support_hint_skips_only_zero_tiles(tmp_path):
    """qsv_run_pass_rounds_sparse on the host: with the state confined to index bits (fixed_mask, fixed_val) outside
    the tile, the generated kernel gives the same result with and without the hint, and leaves tiles outside the
    support untouched (they are poisoned with NaN here: a kernel that read them would spread the NaN)."""
    from qsv import workloads, lowering, schedule
    n = 15
    c = workloads.random_layered_circuit(n, 3, seed=5)
    c.sort()
    ops, _ = lowering.lower(c)
    ps, prog, tables = _round_programs(ops, n)[0]
    outside = [p for p in range(ps.low_bits, n) if p not in ps.hi_pos]
    assert len(outside) == 3
    fixed_mask = (1 << outside[0]) | (1 << outside[2])
    fixed_val = 1 << outside[2]
    idx = np.arange(1 << n)
    inside = (idx & fixed_mask) == fixed_val
    psi = _rand_state(n, 3)
    psi[~inside] = 0
    dense = jit_host_run(ps, prog, tables, psi, 0, str(tmp_path))
    poisoned = psi.copy()
    poisoned[~inside] = np.nan
    hinted = jit_host_run(ps, prog, tables, poisoned, 0, str(tmp_path), fixed_mask=fixed_mask, fixed_val=fixed_val)
    assert np.all(np.isnan(hinted[~inside])) and not np.any(np.isnan(hinted[inside]))
    assert np.array_equal(hinted[inside], dense[inside])
    assert np.all(dense[~inside] == 0)


def test_zero_fill_form_never_reads_outside_the_support(tmp_path):
    """qsv_run_pass_rounds_fresh on the host: the register holds NaN everywhere except on the support of the state
    (index bits fixed outside AND inside the tile).  With (zf_mask, zf_val) = the fixed bits inside the tile the pass
    gives, on every tile it touches, exactly the result of the dense run on the zero-padded state - whatever the
    uninitialised memory held - and writes those tiles in full; tiles outside the support keep their poison.
    Plain loads and loads that carry address permutations (X / CNot riding on round 0) are both covered."""
    from qsv import workloads, lowering, schedule
    n = 15
    seen_permuted = seen_plain = 0
    for seed, which in ((5, 0), (5, 1), (6, 1), (7, 0), (8, 1)):
        c = workloads.random_layered_circuit(n, 3, seed=seed)
        c.sort()
        ops, _ = lowering.lower(c)
        ps, prog, tables = _round_programs(ops, n)[which]
        rounds0_pre = int(np.frombuffer(prog.tobytes()[16 + 12:16 + 14], dtype=np.uint16)[0]) & 0x7fff
        seen_permuted += rounds0_pre > 0
        seen_plain += rounds0_pre == 0
        outside = [p for p in range(ps.low_bits, n) if p not in ps.hi_pos]
        tile_pos = list(range(ps.low_bits)) + list(ps.hi_pos)
        rng = np.random.default_rng(seed)
        fixed_out = [outside[0], outside[2]]
        fixed_in = [int(t) for t in rng.choice(12, size=5, replace=False)]           # tile-local bits still fixed
        vals = {p: int(rng.integers(0, 2)) for p in fixed_out + [tile_pos[t] for t in fixed_in]}
        fixed_mask = sum(1 << p for p in fixed_out)
        fixed_val = sum(vals[p] << p for p in fixed_out)
        zf_mask = sum(1 << t for t in fixed_in)
        zf_val = sum(vals[tile_pos[t]] << t for t in fixed_in)
        idx = np.arange(1 << n)
        in_tiles = (idx & fixed_mask) == fixed_val
        in_support = in_tiles.copy()
        for t in fixed_in:
            in_support &= ((idx >> tile_pos[t]) & 1) == vals[tile_pos[t]]
        psi = _rand_state(n, seed)
        psi[~in_support] = 0
        dense = jit_host_run(ps, prog, tables, psi, 0, str(tmp_path))
        poisoned = psi.copy()
        poisoned[~in_support] = np.nan
        fresh = jit_host_run(ps, prog, tables, poisoned, 0, str(tmp_path), fixed_mask=fixed_mask, fixed_val=fixed_val,
                             zf_mask=zf_mask, zf_val=zf_val)
        assert np.all(np.isnan(fresh[~in_tiles])) and not np.any(np.isnan(fresh[in_tiles]))
        assert np.array_equal(fresh[in_tiles], dense[in_tiles])
    assert seen_plain and seen_permuted


def test_zero_fill_form_writes_tiles_no_op_acts_on(tmp_path):
    """A pass whose ops all hang on a control OUTSIDE the tile skips the tiles where that control is 0 - unless it runs
    in the zero-fill form with fixed bits inside the tile: such a tile holds uninitialised memory next to its
    amplitudes, so it must be loaded, masked and written in full even though no op acts on it."""
    from qsv.schedule import Op
    n = 15
    s = 2 ** -0.5
    H = np.array([[s, s], [s, -s]], dtype=np.complex128)
    ops = [Op("dense", (3,), (14,), matrix=H), Op("mcx", (2,), (14,)), Op("dense", (7,), (14,), matrix=H),
           Op("phase", (), (14, 9), scalar=1j), Op("dense", (1,), (14,), matrix=H)]
    progs = _round_programs(ops, n)
    assert len(progs) == 1
    ps, prog, tables = progs[0]
    assert ps.loc(14) < 0
    tile_pos = list(range(ps.low_bits)) + list(ps.hi_pos)
    fixed_in = [t for t in range(12) if tile_pos[t] not in (1, 2, 3, 7)][:6]
    zf_mask = sum(1 << t for t in fixed_in)
    zf_val = sum(1 << t for t in fixed_in[::2])
    idx = np.arange(1 << n)
    in_support = np.ones(1 << n, dtype=bool)
    for k, t in enumerate(fixed_in):
        in_support &= ((idx >> tile_pos[t]) & 1) == (1 if k % 2 == 0 else 0)
    psi = _rand_state(n, 4)
    psi[~in_support] = 0
    dense = jit_host_run(ps, prog, tables, psi, 0, str(tmp_path))
    poisoned = psi.copy()
    poisoned[~in_support] = np.nan
    fresh = jit_host_run(ps, prog, tables, poisoned, 0, str(tmp_path), zf_mask=zf_mask, zf_val=zf_val)
    assert not np.any(np.isnan(fresh)) and np.array_equal(fresh, dense)
    # without the zero-fill form the tiles with control 14 = 0 are skipped (and keep the poison)
    plain = jit_host_run(ps, prog, tables, poisoned, 0, str(tmp_path))
    assert np.any(np.isnan(plain[(idx >> 14) & 1 == 0]))


def test_sums_variant_adds_exact_tile_weights_to_the_sampler_buckets(tmp_path):
    """qsv_run_pass_rounds_sums on the host: the state comes out exactly as from the plain kernel, and the bucket array
    holds, per bucket (index bits at the given positions outside the tile, rank bits included), the EXACT sum of
    floor(w * 2^96) over its amplitudes with w = fma(re, re, round(im * im)) - big-integer arithmetic here, four 32-bit
    limbs there.  Tiles no op acts on are weighed too; the sums kernel also compiles for sm_100a."""
    import ctypes as C
    from fractions import Fraction
    from qsv import workloads, lowering, _cabi
    from helpers import jit_cpass
    n = 15
    c = workloads.random_layered_circuit(n, 3, seed=8)
    c.sort()
    ops, _ = lowering.lower(c)
    ps, prog, tables = _round_programs(ops, n)[-1]
    outside = [p for p in range(ps.low_bits, n) if p not in ps.hi_pos]
    # two local positions outside the tile, one global (from rank_bits), and three INSIDE the tile (a low bit and two
    # of the tile's high positions: register- or thread-held in the last round, whatever the round planner chose)
    positions = [outside[2], ps.hi_pos[1], outside[0], 3, n + 1, ps.hi_pos[4]]
    rank_bits = 0b10 << n
    psi = _rand_state(n, 12)
    psi[::7] = 0
    psi[5] = 3e-20 + 1e-25j                              # below the 2^-96 resolution after squaring: contributes 0
    ref = jit_host_run(ps, prog, tables, psi, rank_bits, str(tmp_path))
    bk = np.zeros(4 << len(positions), dtype=np.uint64)
    got = jit_host_run(ps, prog, tables, psi, rank_bits, str(tmp_path), sums=(positions, bk))
    assert np.array_equal(got, ref)
    expect = [0] * (1 << len(positions))
    for i, a in enumerate(got):
        im2 = float(a.imag) * float(a.imag)
        w = float(Fraction(float(a.real)) * Fraction(float(a.real)) + Fraction(im2))      # correctly rounded = fma
        g = i | rank_bits
        b = sum(((g >> p) & 1) << k for k, p in enumerate(positions))
        expect[b] += int(Fraction(w) * (1 << 96)) if w >= 2.0 ** -1022 else 0
    have = [int(bk[4 * b]) + (int(bk[4 * b + 1]) << 32) + (int(bk[4 * b + 2]) << 64) + (int(bk[4 * b + 3]) << 96)
            for b in range(1 << len(positions))]
    assert have == expect
    assert sum(1 for v in expect if v) == 32            # the global position pins one half of the bucket index
    lib = _cabi.load()
    cp = jit_cpass(ps)
    need = C.c_size_t(0)
    rc = lib.qsv_jit_compile_sums(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), len(positions),
                                  _cabi.int_array(positions), None, 0, C.byref(need))
    assert rc == 0 and need.value > 1000, lib.qsv_last_error()
    # more than three bucket positions inside the tile are refused
    rc = lib.qsv_jit_compile_sums(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), 4,
                                  _cabi.int_array([0, 1, ps.hi_pos[0], ps.hi_pos[2]]), None, 0, C.byref(need))
    assert rc != 0 and b"inside the tile" in lib.qsv_last_error()


def test_every_kernel_of_the_8_gpu_bench_plan_compiles_for_sm100a():
    """The schedule bench.py runs at --gpus 8 (33 qubits, joint placement, fused swaps): every pass program - the
    plain ones and the fused swap+pass forms - must be accepted by NVRTC for sm_100a.  No GPU needed."""
    from qsv import _cabi, workloads, lowering, schedule
    lib = _cabi.load()
    n, g = 33, 3
    c = workloads.random_layered_circuit(n, 20, seed=20261019)
    c.sort()
    ops, _, _ = lowering.lower(c, n_global=g, in_flip=True, free_swap=True, fuse_swaps=True)
    sched = schedule.plan_schedule(ops, n - g, fuse_swaps=True)
    assert any(k == "swap_pass" for k, _ in sched) and sum(k == "pass" for k, _ in sched) <= 20
    compiled = fused = 0
    for kind, it in sched:
        if kind == "op":
            continue
        ps = it if kind == "pass" else it[1]
        packed = schedule.pack_passes([ps])
        for meta in packed.passes:
            if meta[2] >= 0:
                continue
            prog, cp, need = meta[3][0], jit_cpass(meta[0]), C.c_size_t(0)
            if kind == "swap_pass" and len(packed.passes) == 1:
                pos, gb = schedule.swap_fields(it[0], n - g)
                rc = lib.qsv_jit_compile_swap_ex(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), len(pos),
                                                 _cabi.int_array(pos), _cabi.int_array(gb), int(it[2]), None, 0,
                                                 C.byref(need))
                fused += 1
            else:
                rc = lib.qsv_jit_compile(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), None, 0,
                                         C.byref(need))
            if rc != 0 and b"cannot load libnvrtc" in lib.qsv_last_error():
                pytest.skip("NVRTC is not installed here")
            assert rc == 0, lib.qsv_last_error()
            assert need.value > 10000
            compiled += 1
    assert compiled >= 12 and fused >= 1


def test_bank_score_search_is_never_worse_than_the_default_mapping():
    """qsv_jit_bank_score: modelled bank-group collisions of a pass with the default mapping and with the mapping /
    layout the opt-in search picks.  The search starts from the default, so it can only improve on it; on the 30-qubit
    bench plan it clears the passes that ncu showed load conflicts in."""
    from qsv import _cabi, workloads, lowering, schedule
    lib = _cabi.load()
    n = 30
    c = workloads.random_layered_circuit(n, 20, seed=20261019)
    c.sort()
    ops, _, _ = lowering.lower(c, in_flip=True)
    tot_d = tot_s = 0
    for kind, ps in schedule.plan_schedule(ops, n):
        packed = schedule.pack_passes([ps])
        m = packed.passes[0]
        prog, cp = m[3][0], jit_cpass(m[0])
        d, s_ = C.c_uint64(0), C.c_uint64(0)
        assert lib.qsv_jit_bank_score(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), C.byref(d),
                                      C.byref(s_)) == 0, lib.qsv_last_error()
        assert s_.value <= d.value
        tot_d += d.value
        tot_s += s_.value
    assert tot_s < tot_d


@pytest.mark.parametrize("world_bits,positions,gbits,pass_first,hint", [
    (2, [13], [1], 0, False),            # partial swap: one of two global qubits, the HIGH one (peer = rank ^ 2)
    (2, [12], [0], 1, False),            # partial swap, pass first
    (3, [10, 14], [0, 2], 0, False),     # two of three global qubits, not adjacent
    (2, [9, 14], [0, 1], 0, True),       # full swap of a still-sparse state: pairs of zeros are skipped
    (2, [13], [1], 1, True),             # partial + hint + pass first
])
def test_fused_partial_swaps_and_support_hint(tmp_path, world_bits, positions, gbits, pass_first, hint):
    """General form of the fused kernel (qsv_run_pass_rounds_swap_ex) on the host: positions[k] trades places with
    global bit gbits[k], a subset of the register's global qubits; with the support hint, tile pairs that hold only
    zeros on both ranks are skipped (the output keeps NaN poison there, the expected state is zero).  Reference: the
    swap as an index exchange of the FULL state (helpers.swap_positions) and the plain generated pass."""
    from helpers import jit_host_run_swap, swap_positions
    from qsv.schedule import Op, plan_schedule, pack_passes
    n_local = 15
    n = n_local + world_bits
    P = 1 << world_bits
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
    S = np.array([[0.5 + 0.5j, 0.5 - 0.5j], [0.5 - 0.5j, 0.5 + 0.5j]], dtype=np.complex128)
    ph = complex(math.cos(0.7), math.sin(0.7))
    free = [p for p in range(6, n_local) if p not in positions]
    swap = Op("global_swap", (), (), params={"n_swap": len(positions), "positions": positions, "gbits": gbits})
    gates = [Op("dense", (0,), (), matrix=H), Op("dense", (free[0],), (), matrix=S), Op("mcx", (free[1],), (positions[0],)),
             Op("phase", (), (positions[-1], 3), scalar=ph), Op("phase", (), (n_local, free[2]), scalar=-1.0 + 0j),
             Op("dense", (free[2],), (n - 1,), matrix=H), Op("mcx", (2,), (free[0], positions[0])),
             Op("dense", (free[3],), (), matrix=H), Op("phase", (), (free[3],), scalar=ph), Op("dense", (1,), (), matrix=S),
             Op("diag", (2, positions[0], n_local + gbits[0]), (), matrix=np.exp(1j * np.arange(8))),
             Op("dense", (4,), (), matrix=H)]
    ops = gates + [swap] if pass_first else [swap] + gates
    sched = plan_schedule(ops, n_local, 12, 6, fuse_swaps=True)
    assert sched[0][0] == "swap_pass" and sched[0][1][2] == pass_first, [k for k, _ in sched]
    ps = sched[0][1][1]
    assert len(ps.ops) == len(gates)
    packed = pack_passes([ps])
    ps_, off, n_ops, extra = packed.passes[0]
    prog, tables = extra[0], packed.blob.tobytes()[off:]
    full = np.concatenate([_rand_state(n_local, 80 + r) for r in range(P)])
    support = (0, 0)
    if hint:
        # a state confined to: one swapped position, one global bit, one local position outside the tile
        outside = [p for p in range(6, n_local) if p not in ps.hi_pos and p not in positions]
        fixed_mask = (1 << positions[0]) | (1 << (n_local + gbits[-1])) | (1 << outside[0])
        fixed_val = (1 << (n_local + gbits[-1])) | (1 << outside[0])
        idx = np.arange(1 << n)
        full[(idx & fixed_mask) != fixed_val] = 0
        support = (fixed_mask, fixed_val)
    shards = [full[r << n_local: (r + 1) << n_local].copy() for r in range(P)]
    got = jit_host_run_swap(ps_, prog, tables, shards, positions, n_local, str(tmp_path), pass_first=pass_first,
                            gbits=gbits, support=support)

    def run_pass(state):
        return np.concatenate([jit_host_run(ps_, prog, tables, state[r << n_local: (r + 1) << n_local], r << n_local,
                                            str(tmp_path)) for r in range(P)])
    ref = swap_positions(run_pass(full), swap, n_local) if pass_first else run_pass(swap_positions(full, swap, n_local))
    got = np.concatenate(got)
    assert np.max(np.abs(got - ref)) == 0.0


@pytest.mark.parametrize("world_bits,positions,gbits,pass_first", [
    (1, [13], [0], 0), (1, [13], [0], 1), (2, [9, 14], [0, 1], 0), (2, [13], [1], 1), (3, [10, 14], [0, 2], 0)])
def test_fused_swap_zero_fill_form_on_poisoned_shards(tmp_path, world_bits, positions, gbits, pass_first):
    """qsv_run_pass_rounds_swap_fresh on the host: the register was not cleared, everything outside the support of the
    state is NaN.  Tiles pulled from outside the support count as zeros, tiles inside it keep only the amplitudes that
    agree with the fixed bits inside the tile; every tile of an active pair is written in full (no NaN survives there)
    and equals swap + pass of the clean state; pairs outside the support on both sides are left alone."""
    from helpers import jit_host_run_swap, swap_positions
    from qsv.schedule import Op, plan_schedule, pack_passes
    n_local = 15
    n = n_local + world_bits
    P = 1 << world_bits
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5
    S = np.array([[0.5 + 0.5j, 0.5 - 0.5j], [0.5 - 0.5j, 0.5 + 0.5j]], dtype=np.complex128)
    ph = complex(math.cos(0.7), math.sin(0.7))
    free = [p for p in range(6, n_local) if p not in positions]
    swap = Op("global_swap", (), (), params={"n_swap": len(positions), "positions": positions, "gbits": gbits})
    gates = [Op("dense", (0,), (), matrix=H), Op("dense", (free[0],), (), matrix=S), Op("mcx", (free[1],), (positions[0],)),
             Op("phase", (), (positions[-1], 3), scalar=ph), Op("phase", (), (n_local, free[2]), scalar=-1.0 + 0j),
             Op("dense", (free[2],), (n - 1,), matrix=H), Op("mcx", (2,), (free[0], positions[0])),
             Op("dense", (free[3],), (), matrix=H), Op("dense", (1,), (), matrix=S), Op("dense", (4,), (), matrix=H)]
    ops = gates + [swap] if pass_first else [swap] + gates
    sched = plan_schedule(ops, n_local, 12, 6, fuse_swaps=True)
    assert sched[0][0] == "swap_pass" and sched[0][1][2] == pass_first
    ps = sched[0][1][1]
    packed = pack_passes([ps])
    ps_, off, n_ops, extra = packed.passes[0]
    prog, tables = extra[0], packed.blob.tobytes()[off:]
    tile_phys = list(range(ps.low_bits)) + list(ps.hi_pos)
    outside = [p for p in range(6, n_local) if p not in ps.hi_pos and p not in positions]
    # support: the first swapped position = 0, one global bit = 1, a local position outside the tile = 1 (-> sp_mask),
    # and INSIDE the tile position 5 = 1, position 0 = 0 and the tile's highest position = 0 (-> zf_mask, tile-local)
    sp_mask = (1 << positions[0]) | (1 << (n_local + gbits[-1])) | (1 << outside[0])
    sp_val = (1 << (n_local + gbits[-1])) | (1 << outside[0])
    in_tile = {5: 1, 0: 0, ps.hi_pos[-1]: 0}
    zf_mask = sum(1 << tile_phys.index(p) for p in in_tile)
    zf_val = sum(v << tile_phys.index(p) for p, v in in_tile.items())
    full_mask = sp_mask | sum(1 << p for p in in_tile)
    full_val = sp_val | sum(v << p for p, v in in_tile.items())
    clean = np.concatenate([_rand_state(n_local, 90 + r) for r in range(P)])
    idx = np.arange(1 << n)
    inside = (idx & full_mask) == full_val
    clean[~inside] = 0
    poisoned = clean.copy()
    poisoned[~inside] = complex(np.nan, np.nan)
    shards = [poisoned[r << n_local: (r + 1) << n_local].copy() for r in range(P)]
    got = np.concatenate(jit_host_run_swap(ps_, prog, tables, shards, positions, n_local, str(tmp_path),
                                           pass_first=pass_first, gbits=gbits, support=(sp_mask, sp_val),
                                           fresh=(zf_mask, zf_val)))

    def run_pass(state):
        return np.concatenate([jit_host_run(ps_, prog, tables, state[r << n_local: (r + 1) << n_local], r << n_local,
                                            str(tmp_path)) for r in range(P)])
    ref = swap_positions(run_pass(clean), swap, n_local) if pass_first else run_pass(swap_positions(clean, swap, n_local))
    bad = np.isnan(got.real) | np.isnan(got.imag)
    assert np.count_nonzero(ref) > 0 and not np.any(bad & (ref != 0))
    assert np.array_equal(got[~bad], ref[~bad])
    # what was left alone is exactly the set of tile pairs outside the support on both sides: after the swap the support
    # has the swapped position's and the global bit's roles exchanged where they were part of it
    assert np.count_nonzero(~bad) >= 2 * np.count_nonzero(ref) // 8
    # the plain form of the same kernel would have used the poison as data
    plain = np.concatenate(jit_host_run_swap(ps_, prog, tables, shards, positions, n_local, str(tmp_path),
                                             pass_first=pass_first, gbits=gbits, support=(sp_mask, sp_val)))
    assert np.any(np.isnan(plain[ref != 0].real))


def test_sums_static_and_general_classification_agree(tmp_path, monkeypatch):
    """The sums variant classifies an amplitude into its bucket either from the final global index (general form) or,
    when nothing permutes the last round's stores, from (bits outside the tile | thread bits | register-slot bits) fixed
    at generation time (static form, qsv_tile_sum_s).  Both host forms must fill the bucket array identically, for
    bucket positions on register bits, thread bits and outside the tile, over several pass programs."""
    import ctypes as C
    from qsv import workloads, lowering, _cabi
    from helpers import jit_source_sums, jit_cpass
    lib = _cabi.load()
    n = 15
    static_seen = general_seen = 0
    kinds = set()
    for seed in (8, 9, 11, 12, 14):
        c = workloads.random_layered_circuit(n, 3, seed=seed)
        c.sort()
        ops, _ = lowering.lower(c)
        for ps, prog, tables in _round_programs(ops, n)[-2:]:
            outside = [p for p in range(ps.low_bits, n) if p not in ps.hi_pos]
            tile_phys = list(range(ps.low_bits)) + list(ps.hi_pos)
            hdr = prog[:16].view(np.uint32)
            rb = [int(b) for b in prog[16 + 16 * (int(hdr[0]) - 1):][:4]]          # register bits of the LAST round
            reg_pos = [tile_phys[b] for b in rb]
            thr_pos = [p for p in tile_phys if p not in reg_pos]
            for positions in ([outside[0], reg_pos[0], reg_pos[3], outside[1], reg_pos[1]],
                              [thr_pos[0], outside[1], thr_pos[5], thr_pos[7]],
                              [reg_pos[2], thr_pos[1], n + 1, thr_pos[6], outside[0]],
                              [outside[2], outside[0], n]):
                psi = _rand_state(n, seed)
                psi[::5] = 0
                res = {}
                for mode in ("1", "0"):
                    monkeypatch.setenv("QSV_SUM_STATIC", mode)
                    src = jit_source_sums(ps, prog, positions)
                    is_static = "qsv_tile_sum_s" in src
                    assert mode == "1" or not is_static
                    cp = jit_cpass(ps)
                    assert lib.qsv_jit_sums_static(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size)) == int(is_static)
                    bk = np.zeros(4 << len(positions), dtype=np.uint64)
                    out = jit_host_run(ps, prog, tables, psi, 0b01 << n, str(tmp_path), sums=(positions, bk))
                    res[mode] = (out, bk, is_static)
                assert np.array_equal(res["1"][0], res["0"][0])
                assert np.array_equal(res["1"][1], res["0"][1]), (seed, positions)
                assert res["0"][1].any()
                if res["1"][2]:
                    static_seen += 1
                    kinds.add((sum(p in reg_pos for p in positions), sum(p in thr_pos for p in positions)))
                else:
                    general_seen += 1
    assert static_seen >= 8, "the static form was expected for most last rounds (no permutation on the stores)"
    assert {(3, 0), (0, 3), (1, 2), (0, 0)} <= kinds
