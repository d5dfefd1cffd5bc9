"""Sharded execution without GPUs: placement (global<->local qubit swaps) checked on virtual ranks, and
the exchange plumbing checked with two real processes over gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import statevec_oracle as so
from helpers import circuit_from_steps, emulate_sharded
from qsv import schedule
from qsv.lowering import lower

TOL = 1e-12


def _permute_to_output(psi, src_pos):
    n = len(src_pos)
    z = np.arange(1 << n)
    y = np.zeros_like(z)
    for b in range(n):
        y |= ((z >> b) & 1) << src_pos[b]
    return psi[y]


@pytest.mark.parametrize("free_swap", [False, True, "joint", "joint-deep"])
@pytest.mark.parametrize("g", [1, 2, 3])
def test_placed_circuit_on_virtual_ranks_matches_oracle(g, free_swap, monkeypatch):
    """free_swap False: victims moved to the top positions first; True: list-order placement with swaps at any
    position; "joint": placement decided together with pass planning (schedule.place_joint, the default)."""
    n = 11
    monkeypatch.setattr(schedule, "SWAP_MIN_POS", 3)     # small register: let the free swaps use low positions too
    monkeypatch.setenv("QSV_PLACE", "joint" if str(free_swap).startswith("joint") else "list")
    monkeypatch.setenv("QSV_TILE_BITS", "6")
    monkeypatch.setenv("QSV_LOW_BITS", "2")
    depth = 14 if free_swap == "joint-deep" else 5
    steps = so.random_layered_steps(n, depth, seed=21 + g)
    c = circuit_from_steps(steps, n)
    c.sort()
    ops, src_pos = lower(c, n_global=g, free_swap=free_swap)
    n_swaps = sum(op.kind == "global_swap" for op in ops)
    assert n_swaps >= 1, "a layered circuit on every wire must touch the global qubits"
    if free_swap:
        assert not any(op.name == "relayout" and op.kind == "mcx" for op in ops), "free swaps need no re-layout CNots"
    if free_swap is True:
        assert any(op.params["positions"] != list(range(n - 2 * g, n - g)) for op in ops if op.kind == "global_swap")
    if free_swap == "joint-deep":
        assert n_swaps >= 2
    assert all(max(op.nondiag_bits(), default=0) < n - g for op in ops if op.kind != "global_swap")
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0b101] = 1
    got = emulate_sharded(ops, psi, n, g, tile_bits=6, low_bits=2)
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[0b101] = 1
    for M, wires, _ in steps:
        ref = so.apply_gate(ref, n, wires, M)
    assert np.max(np.abs(_permute_to_output(got, src_pos) - ref)) <= TOL


def test_diagonal_and_control_use_of_global_qubits_needs_no_swap():
    n, g = 10, 2
    steps = [(so.mat_H(), [9], "H"), (so.mat_Cnot(), [0, 9], "CNot"), (so.mat_CZ(), [0, 1], "CZ"),
             (so.mat_PhaseShift(0.3), [1], "P"), (so.mat_Toffoli(), [0, 1, 8], "T")]
    c = circuit_from_steps(steps, n)
    c.sort()
    ops, src_pos = lower(c, n_global=g)           # wires 0,1 are the global qubits: only controls / phases
    assert not any(op.kind == "global_swap" for op in ops)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[(1 << 9) | (1 << 8) | 1] = 1
    got = emulate_sharded(ops, psi, n, g, tile_bits=5, low_bits=2)
    ref = psi.copy()
    for M, wires, _ in steps:
        ref = so.apply_gate(ref, n, wires, M)
    assert np.max(np.abs(_permute_to_output(got, src_pos) - ref)) <= TOL


def test_exchange_plan():
    from qsv.dist import plan_exchange
    P, S, C, k = plan_exchange(33, 3, 1 << 30)
    assert (P, S) == (8, 1 << 30) and C * k == S and 16 * (P - 1) * C <= 1 << 30
    P, S, C, k = plan_exchange(10, 1, 1 << 30)
    assert (P, S, C, k) == (2, 512, 512, 1)


def _worker(rank, world, port, n, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qsv.dist import Comm
    comm = Comm(staging_bytes=16 * world * 8)        # tiny staging: forces many chunks
    g = comm.global_bits
    n_local = n - g
    full = np.random.default_rng(seed).normal(size=(1 << n, 2)).view(np.complex128).reshape(-1)
    state = torch.from_numpy(full.reshape(world, -1)[rank].copy())
    P, S = world, 1 << (n_local - g)

    peers = [j for j in range(P) if j != rank]

    def pack(staging, offset, C):
        staging.view(P - 1, C).copy_(state.view(P, S)[peers, offset:offset + C])

    def unpack(staging, offset, C):
        state.view(P, S)[peers, offset:offset + C] = staging.view(P - 1, C)

    comm.exchange(state, n_local, pack, unpack)
    expect = np.ascontiguousarray(full.reshape(world, world, -1).transpose(1, 0, 2)).reshape(world, -1)[rank]
    ok = np.array_equal(state.numpy(), expect)
    t = torch.tensor([1.0 + rank, 2.0], dtype=torch.complex128)
    comm.all_reduce_sum(t)
    ok = ok and complex(t[0]) == sum(1.0 + r for r in range(world)) and comm.swaps == 1
    ok = ok and comm.bytes_sent == 16 * (P - 1) * S
    with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_exchange_over_gloo(tmp_path, world):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, 9, 3, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / ("ok%d" % r)).read() == "1"


def _insert_zero(v, p):
    return ((v >> p) << (p + 1)) | (v & ((1 << p) - 1))


def _swap_expand(e, q):
    for p in q:
        e = _insert_zero(e, p)
    return e


@pytest.mark.parametrize("world_bits,gb,S,low,hint", [
    (1, [0], 1, 0, False), (1, [0], 2, 0, False), (1, [0], 8, 0, False), (2, [0, 1], 4, 0, False),
    (3, [0, 1, 2], 16, 0, False), (2, [0, 1], 8192, 0, False), (1, [0], 4096, 0, False),
    (1, [0], 8, 1, False), (2, [0, 1], 16, 2, False), (3, [0, 1, 2], 64, 3, False), (2, [0, 1], 8192, 5, False),
    # partial swaps: a subset of the global bits
    (2, [1], 16, 0, False), (2, [0], 2048, 3, False), (3, [0, 2], 64, 2, False), (3, [1], 4096, 4, False),
    # support hint: pairs that are zero on both sides are skipped, everything else still moves exactly once
    (2, [0, 1], 4096, 0, True), (3, [0, 2], 2048, 3, True), (1, [0], 8192, 5, True),
])
def test_peer_swap_kernel_index_logic(world_bits, gb, S, low, hint):
    """numpy restatement of k_swap_peer (csrc/qsv_peer.cu), chunk loop included: every element of every pair is
    swapped exactly once, by exactly one of the two ranks, and the result equals the exchange of the chosen local
    positions with the chosen global bits (low = 0: the top positions, i.e. the all-to-all transposition of blocks;
    low > 0: positions spread below the top, as the free-position placement produces them).  With the support hint
    only pairs of zeros are left alone."""
    P = 1 << world_bits
    j_sw = len(gb)
    n_local = j_sw + (S.bit_length() - 1)
    if low == 0:
        q = list(range(n_local - j_sw, n_local))
    else:
        q = sorted(np.random.default_rng(low).choice(np.arange(low, n_local), size=j_sw, replace=False).tolist())

    def field(x):
        return sum(((x >> gb[k]) & 1) << q[k] for k in range(j_sw))

    G = (S // 2 if S // 2 < 2048 else 2048) if S >= 2 else 1
    per_pair = S // 2 if S >= 2 else 1
    CH = min(per_pair, 512)
    n_peers = (1 << j_sw) - 1
    n_chunks = n_peers * (per_pair // CH)
    rng = np.random.default_rng(S + world_bits)
    N = 1 << n_local
    shards = rng.normal(size=(P, N))
    sp_mask = sp_val = 0
    if hint:
        cand = [p for p in range(n_local) if p not in q]
        sp_mask = (1 << q[0]) | (1 << cand[-1]) | (1 << cand[0]) | (1 << (n_local + gb[-1]))
        sp_val = (1 << cand[-1]) | (1 << (n_local + gb[-1]))
        gidx = np.arange(P * N)
        flat = shards.reshape(-1)
        flat[(gidx & sp_mask) != sp_val] = 0.0
    from helpers import swap_positions
    from qsv.schedule import Op
    swap_op = Op("global_swap", (), (), params={"n_swap": j_sw, "positions": q, "gbits": gb})
    expect = swap_positions(shards.reshape(-1).copy(), swap_op, n_local).reshape(P, N)
    if low == 0 and j_sw == world_bits and not hint:
        assert np.array_equal(expect, np.ascontiguousarray(shards.reshape(P, P, S).transpose(1, 0, 2)).reshape(P, N))
    work = shards.copy()
    touched = np.zeros((P, N), dtype=int)
    var_mask = _swap_expand(CH - 1, q)
    cm, cv = sp_mask & ~var_mask, sp_val & ~var_mask
    skipped_chunks = 0
    for r in range(P):
        fr = field(r)
        for c in range(n_chunks):
            s_ = 1 + c % n_peers
            dj = sum(((s_ >> k) & 1) << gb[k] for k in range(j_sw))
            j = r ^ dj
            fj = field(j)
            hb = (c // n_peers) * CH
            if S >= 2:
                e0 = ((hb // G) * 2 + (0 if r < j else 1)) * G + (hb % G)
            else:
                e0 = 0
                if r > j:
                    continue
            x0 = _swap_expand(e0, q)
            mine_g, theirs_g = (r << n_local) | x0 | fj, (j << n_local) | x0 | fr
            if (mine_g & cm) != cv and (theirs_g & cm) != cv:
                skipped_chunks += 1
                continue
            for l in range(CH):
                x = _swap_expand(e0 + l, q)
                if sp_mask:
                    mg, tg = (r << n_local) | x | fj, (j << n_local) | x | fr
                    if (mg & sp_mask) != sp_val and (tg & sp_mask) != sp_val:
                        continue
                mine, theirs = x | fj, x | fr
                a, b = work[r, mine], work[j, theirs]
                work[r, mine], work[j, theirs] = b, a
                touched[r, mine] += 1
                touched[j, theirs] += 1
    assert np.array_equal(work, expect)
    assert touched.max() <= 1
    if not hint:
        moves = np.ones((P, N), dtype=bool)
        loc = np.arange(N)
        for r in range(P):
            own = np.ones(N, dtype=bool)
            for k in range(j_sw):
                own &= ((loc >> q[k]) & 1) == ((r >> gb[k]) & 1)
            moves[r, own] = False
        assert np.all(touched[moves] == 1) and np.all(touched[~moves] == 0)
    else:
        assert skipped_chunks > 0 or n_chunks < 4
        assert touched.sum() < 0.6 * P * N


@pytest.mark.parametrize("g,seed,must_fuse", [(1, 5, True), (2, 1, True), (2, 2, False), (3, 2, False)])
def test_support_tracking_with_fused_swap_pass_items(monkeypatch, g, seed, must_fuse):
    """Same invariant on schedules of the default geometry (12-bit tiles) with 2, 4 and 8 virtual ranks: joint
    placement, plain and fused swap+pass items; the tracker must follow the 'pass first' / 'pass after' order of
    the fused kernel, and the final state must equal the oracle's."""
    from helpers import emulate_program
    from qsv.schedule import Support, plan_schedule, pack_passes
    n = 14 + g
    monkeypatch.setattr(schedule, "SWAP_MIN_POS", 6)
    steps = so.random_layered_steps(n, 10, seed=seed)
    c = circuit_from_steps(steps, n)
    c.sort()
    ops, src_pos, flip = lower(c, n_global=g, in_flip=True, free_swap=True, fuse_swaps=True)
    n_local, P = n - g, 1 << g
    basis = 0b1101 ^ flip
    full = np.zeros(1 << n, dtype=np.complex128)
    full[basis] = 1
    sup = Support(n, basis)
    idx = np.arange(1 << n)

    def run_pass(ps, full):
        packed = pack_passes([ps])
        shards = full.reshape(P, -1)
        return np.concatenate([emulate_program(packed, shards[r], rank_bits=r << n_local) for r in range(P)])

    def run_swap(swap_op, full):
        from helpers import swap_positions
        return swap_positions(full, swap_op, n_local)

    kinds = []
    for kind, item in plan_schedule(ops, n_local, fuse_swaps=True):
        kinds.append(kind)
        if kind == "pass":
            fixed, val = sup.pass_mask(item)
            assert np.all(full[(idx & fixed) != val] == 0)
            full = run_pass(item, full)
            sup.after_ops(item.ops)
        elif kind == "swap_pass":
            swap_op, ps, pass_first = item
            fixed, val = sup.pass_mask(ps)                  # what the fused kernel is told: pairs outside are skipped
            assert np.all(full[(idx & fixed) != val] == 0)
            if pass_first:
                full = run_swap(swap_op, run_pass(ps, full))
                sup.after_ops(ps.ops)
                sup.swap_op(swap_op, n_local)
            else:
                full = run_pass(ps, run_swap(swap_op, full))
                sup.swap_op(swap_op, n_local)
                sup.after_ops(ps.ops)
        else:
            assert item.kind == "global_swap"
            fixed, val = sup.fixed()                        # what k_swap_peer is told
            assert np.all(full[(idx & fixed) != val] == 0)
            full = run_swap(item, full)
            sup.swap_op(item, n_local)
    assert "swap_pass" in kinds or not must_fuse, kinds
    assert "swap_pass" in kinds or "op" in kinds, kinds
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[0b1101] = 1
    for M, wires, _ in steps:
        ref = so.apply_gate(ref, n, wires, M)
    assert np.max(np.abs(_permute_to_output(full, src_pos) - ref)) <= TOL


@pytest.mark.parametrize("g", [0, 2])
def test_support_tracking_state_is_zero_outside_the_claimed_support(g, monkeypatch):
    """schedule.Support: before every pass of a plan run on a basis state, every amplitude whose index disagrees
    with (fixed_mask, fixed_val) on a position OUTSIDE the pass's tile must be exactly zero - that is what lets the
    kernels skip those tiles (qsv_run_pass_rounds_sparse).  Checked on the numpy emulation of a (sharded) run."""
    from helpers import emulate_program, _emulate_whole_state_op
    from qsv.schedule import Support, plan_schedule, pack_passes
    n = 12
    monkeypatch.setattr(schedule, "SWAP_MIN_POS", 3)
    monkeypatch.setenv("QSV_TILE_BITS", "6")
    monkeypatch.setenv("QSV_LOW_BITS", "2")
    steps = so.random_layered_steps(n, 6, seed=77)
    c = circuit_from_steps(steps, n)
    c.sort()
    ops, src_pos, flip = lower(c, n_global=g, in_flip=True, free_swap=bool(g))
    n_local, P = n - g, 1 << g
    basis = 0b101101 ^ flip
    full = np.zeros(1 << n, dtype=np.complex128)
    full[basis] = 1
    sup = Support(n, basis)
    idx = np.arange(1 << n)
    sparse_passes = 0
    for kind, item in plan_schedule(ops, n_local, 6, 2):
        if kind == "pass":
            fixed, val = sup.pass_mask(item)
            assert not fixed & schedule._mask(list(range(item.low_bits)) + list(item.hi_pos))
            outside = (idx & fixed) != val
            assert np.all(full[outside] == 0), "non-zero amplitude outside the claimed support"
            sparse_passes += bool(fixed)
            packed = pack_passes([item])
            shards = full.reshape(P, -1)
            full = np.concatenate([emulate_program(packed, shards[r], rank_bits=r << n_local) for r in range(P)])
            sup.after_ops(item.ops)
        elif item.kind == "global_swap":
            from helpers import swap_positions
            fixed, val = sup.fixed()
            assert np.all(full[(idx & fixed) != val] == 0)
            full = swap_positions(full, item, n_local)
            sup.swap_op(item, n_local)
        else:
            raise AssertionError(item.kind)
    assert sparse_passes >= 2
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[0b101101] = 1
    for M, wires, _ in steps:
        ref = so.apply_gate(ref, n, wires, M)
    assert np.max(np.abs(_permute_to_output(full, src_pos) - ref)) <= TOL
