"""Multi-GPU parity check, launched with torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py

Every rank runs the sharded engine; the shards are all-gathered and rank 0 compares the full state with
the CPU oracle (amplitudes within 1e-12; permutation-only circuits bit-exact)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "quantum-circuits_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import statevec_oracle as so          # noqa: E402
from qsv import engine, lowering, workloads   # noqa: E402
from qsv.dist import Comm             # noqa: E402


def gather_full(sv, comm):
    parts = [torch.empty_like(sv.state) for _ in range(comm.world)]
    dist.all_gather(parts, sv.state)
    return torch.cat(parts).cpu().numpy()


def to_output_order(psi, src_pos):
    n = len(src_pos)
    z = np.arange(1 << n)
    y = np.zeros_like(z)
    for b in range(n):
        y |= ((z >> b) & 1) << src_pos[b]
    return psi[y]


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm(staging_bytes=1 << 16)          # small staging: exercises the chunked in-place exchange
    g = comm.global_bits
    ok = True

    def check(name, cond):
        nonlocal ok
        if rank == 0:
            print(("PASS " if cond else "FAIL ") + name, flush=True)
        ok = ok and bool(cond)

    # 1. layered circuits (H / X / phases / CNot / CZ on every wire -> global<->local swaps)
    from qsv import schedule
    for n, depth, seed, min_pos in ((10 + g, 4, 3, 9), (14 + g, 6, 4, 9), (14 + g, 6, 5, 2), (15 + g, 16, 6, 6)):
        c = workloads.random_layered_circuit(n, depth, seed=seed)
        c.sort()
        sv = engine.StateVector(n, comm=comm)
        ref = np.zeros(1 << n, dtype=np.complex128)
        ref[5] = 1
        for M, wires, _ in so.random_layered_steps(n, depth, seed=seed):
            ref = so.apply_gate(ref, n, wires, M)
        # re-layout to the top positions (works with the NCCL exchange too), then - with peer memory - the
        # free-position swaps the benchmarks use (victims exchanged where they are, X gates as an input flip)
        for free in ([False, True] if comm.peer_swap else [False]):
            schedule.SWAP_MIN_POS = min_pos
            ops, src_pos, flip = lowering.lower(c, n_global=g, in_flip=True, free_swap=free)
            schedule.SWAP_MIN_POS = 9
            swaps0 = comm.swaps
            sv.init_basis(5 ^ flip)
            sv.run_ops(ops)
            sv.synchronize()
            full = gather_full(sv, comm)
            err = np.max(np.abs(to_output_order(full, src_pos) - ref))
            check("layered n=%d depth=%d free_swap=%s min_pos=%d swaps=%d err=%.2e"
                  % (n, depth, free, min_pos, comm.swaps - swaps0, err), err <= 1e-12)
        # sampling (main/circuit.py:360-364): the exact rule in OUTPUT order - identical on all ranks, equal to the
        # brute-force bisect over the exact cumulative sums of the gathered device weights, i.e. to what ONE GPU
        # returns for the same state and draw, and equal to the oracle's float rule away from ties
        import helpers
        wparts = [torch.empty(1 << sv.n_local, dtype=torch.float64, device="cuda") for _ in range(comm.world)]
        dist.all_gather(wparts, sv.probs(None))
        w_phys = torch.cat(wparts).cpu().numpy()
        for u in (0.4321, 0.0, 0.87654321):
            z = sv.sample(u, src_pos)
            zs = torch.tensor([z], device="cuda")
            allz = [torch.empty_like(zs) for _ in range(comm.world)]
            dist.all_gather(allz, zs)
            want = helpers.exact_draw_bruteforce(w_phys, n, src_pos, u) if rank == 0 else z
            # (weights that are rounding noise of an exactly-zero amplitude - 1e-33 against outcomes of 1e-6 - lie below
            # the 2^-96 resolution of the exact rule and count as zero there: the float rule gets them as zero too,
            # otherwise u = 0 would return whichever noise entry the oracle's own rounding left first)
            w_out = np.abs(ref) ** 2
            w_out = np.where(w_out >= 2.0 ** -80, w_out, 0.0)
            check("sample n=%d u=%.4f z=%d (exact rule %d, float rule of the oracle %d)"
                  % (n, u, z, want, so.sample_index_np(w_out, u)),
                  all(int(t) == z for t in allz) and z == want and z == so.sample_index_np(w_out, u))

    # 1b. swap + pass fused into one kernel over peer memory (specialised kernels forced at this small size)
    if comm.peer_swap:
        os.environ["QSV_JIT"] = "force"
        fused_total = 0
        for n, depth, seed in ((16 + g, 16, 12), (17 + g, 24, 13)):
            c = workloads.random_layered_circuit(n, depth, seed=seed)
            c.sort()
            sv = engine.StateVector(n, comm=comm)
            ref = np.zeros(1 << n, dtype=np.complex128)
            ref[9] = 1
            for M, wires, _ in so.random_layered_steps(n, depth, seed=seed):
                ref = so.apply_gate(ref, n, wires, M)
            ops, src_pos, flip = lowering.lower(c, n_global=g, in_flip=True, free_swap=True, fuse_swaps=sv.fuse_swaps())
            fused0 = getattr(comm, "fused_swaps", 0)
            for rep in range(2):                   # twice: the flag epochs must keep the handshake correct
                sv.init_basis(9 ^ flip)
                sv.run_ops(ops)
                sv.synchronize()
            full = gather_full(sv, comm)
            err = np.max(np.abs(to_output_order(full, src_pos) - ref))
            fused_total += getattr(comm, "fused_swaps", 0) - fused0
            check("fused swap+pass n=%d depth=%d fused_kernels=%d err=%.2e"
                  % (n, depth, getattr(comm, "fused_swaps", 0) - fused0, err), err <= 1e-12)
            del sv
        # whether a given swap can carry a pass is the planner's call (victims outside the pass's tile); across the two
        # circuits at least one fused kernel must have run, or this section tested nothing
        check("fused swap+pass kernels executed: %d" % fused_total, fused_total > 0)
        # 1b'. the same circuits WITHOUT the clear of the register (run_from_basis on a sharded register): the shards
        #      are filled with NaN first; passes, support-aware swaps and fused swap+pass kernels up to the first full
        #      sweep run in the zero-fill form - any read from outside the support would poison the result
        fresh_groups = fresh_fused = 0
        for n, depth, seed in ((16 + g, 16, 12), (17 + g, 24, 13), (17 + g, 20, 21)):
            c = workloads.random_layered_circuit(n, depth, seed=seed)
            c.sort()
            sv = engine.StateVector(n, comm=comm)
            ref = np.zeros(1 << n, dtype=np.complex128)
            ref[0] = 1
            for M, wires, _ in so.random_layered_steps(n, depth, seed=seed):
                ref = so.apply_gate(ref, n, wires, M)
            ops, src_pos, flip = lowering.lower(c, n_global=g, in_flip=True, free_swap=True, fuse_swaps=sv.fuse_swaps())
            plan = sv.prepare(ops)
            n_fresh = sv._fresh_prefix(plan)
            kinds = [k for k, _, _, _ in sv._walk(plan, flip)][:n_fresh]
            fresh_groups += n_fresh
            fresh_fused += kinds.count("swap_pass")
            for rep in range(2):
                torch.view_as_real(sv.state).fill_(float("nan"))
                torch.cuda.synchronize()
                dist.barrier()
                sv.run_from_basis(flip, plan)
                sv.synchronize()
            full = gather_full(sv, comm)
            err = np.max(np.abs(to_output_order(full, src_pos) - ref)) if not np.any(np.isnan(full)) else float("nan")
            check("no clear, NaN-poisoned shards n=%d depth=%d zero-fill groups=%d %s err=%.2e"
                  % (n, depth, n_fresh, kinds, err), err <= 1e-12)
            del sv
        check("zero-fill launch groups executed: %d (fused swap+pass among them: %d)" % (fresh_groups, fresh_fused),
              fresh_groups > 0)
        os.environ.pop("QSV_JIT")

    # 1c. QFT as H + controlled-phase ladder on a sharded register (config 4): every wire gets one H, so every global
    #     qubit has to come down once; closed form of Matrix.QFT on a basis state (main/matrix.py:212-219, sign +i):
    #     psi[k] = exp(2 pi i x k / 2^n) / sqrt(2^n) in output order (the bit reversal sits in the output wiring)
    for n, jit in ((12 + g, None), (16 + g, "force")) if comm.peer_swap else ((12 + g, None),):
        if jit:
            os.environ["QSV_JIT"] = jit
        c = workloads.qft_ladder_circuit(n)
        c.sort()
        in_v = [(i + 1) % 2 for i in range(n)]
        x = lowering.basis_index(in_v)
        k = np.arange(1 << n, dtype=np.float64)
        ref = np.exp(2j * np.pi * ((x * k) % (1 << n)) / (1 << n)) / np.sqrt(float(1 << n))
        sv = engine.StateVector(n, comm=comm)
        for free in ([False, True] if comm.peer_swap else [False]):
            ops, src_pos, flip = lowering.lower(c, n_global=g, in_flip=True, free_swap=free,
                                                fuse_swaps=sv.fuse_swaps() if free else False)
            swaps0, fused0 = comm.swaps, getattr(comm, "fused_swaps", 0)
            sv.init_basis(x ^ flip)
            sv.run_ops(ops)
            sv.synchronize()
            full = gather_full(sv, comm)
            err = np.max(np.abs(to_output_order(full, src_pos) - ref))
            check("qft ladder n=%d jit=%s free_swap=%s swaps=%d fused=%d err=%.2e"
                  % (n, jit, free, comm.swaps - swaps0, getattr(comm, "fused_swaps", 0) - fused0, err), err <= 1e-12)
        del sv
        if jit:
            os.environ.pop("QSV_JIT")

    # 2. Grover with the dedicated oracle + diffusion kernels (all-reduce of the two means)
    ns, K, x = 9 + g, 5, 0b1011001 % (1 << (9 + g))
    c = workloads.grover_circuit(ns, x, K)
    c.sort()
    ops, src_pos = lowering.lower(c, n_global=g)
    sv = engine.StateVector(ns + 1, comm=comm)
    sv.init_basis(1)
    sv.run_ops(ops)
    sv.synchronize()
    full = gather_full(sv, comm)
    psi = so.apply_gate(so.initial_state([0] * ns + [1]), ns + 1, list(range(ns + 1)), so.kron_power(so.mat_H(), ns + 1))
    for _ in range(K):
        psi = so.grover_iteration_closed_form(psi, ns, x)
    err = np.max(np.abs(to_output_order(full, src_pos) - psi))
    check("grover n_search=%d K=%d err=%.2e" % (ns, K, err), err <= 1e-12)

    # 3. Shor circuit with the modexp kernel: x register partly on global qubits
    N, a = 21, 17
    c, q = workloads.shor_circuit(N, a, dense=False)
    c.sort()
    ops, src_pos = lowering.lower(c, n_global=g)
    sv = engine.StateVector(q, comm=comm)
    in_v = [0] * (q // 2) + [1] * (q // 2)
    sv.init_basis(lowering.basis_index(in_v))
    sv.run_ops(ops)
    sv.synchronize()
    full = gather_full(sv, comm)
    steps, _ = so.shor_steps(N, a)
    ref = so.evolve(steps, list(range(q)), in_v)
    err = np.max(np.abs(to_output_order(full, src_pos) - ref))
    check("shor N=21 a=17 err=%.2e" % err, err <= 1e-12)

    # 4. the public API under a one-process-per-GPU launch: Circuit.run / run_method2 shard the register over
    #    the ranks by themselves (qsv.dist.default_comm); weights in output order on every rank, one sample
    #    shared by all ranks (rank 0's uniform draw)
    import io
    import contextlib
    import random
    os.environ["QSV_SHARD_MIN_LOCAL"] = "8"
    n, depth, seed = 13 + g, 12, 9
    c = workloads.random_layered_circuit(n, depth, seed=seed)
    ref = so.initial_state([0] * n)
    for M, wires, _ in so.random_layered_steps(n, depth, seed=seed):
        ref = so.apply_gate(ref, n, wires, M)
    wref = so.weights(ref)
    with contextlib.redirect_stdout(io.StringIO()):
        w = np.array(c.run_method2([0] * n))
    st = lowering.last_run_stats()
    err = float(np.max(np.abs(w - wref)))
    from qsv.dist import default_comm
    dc = default_comm()
    check("API run_method2 sharded n=%d depth=%d (world=%d, peer_swap=%s, passes=%d) err=%.2e"
          % (n, depth, dc.world if dc else 1, bool(dc and dc.peer_swap), st["passes"], err),
          dc is not None and err <= 1e-12)
    random.seed(100 + rank)                      # different generators per rank: rank 0's draw must win
    with contextlib.redirect_stdout(io.StringIO()):
        out = c.run([0] * n, 2)
    random.seed(100)
    u0 = random.random()
    zbits = torch.tensor(out, device="cuda")
    allb = [torch.empty_like(zbits) for _ in range(comm.world)]
    dist.all_gather(allb, zbits)
    idx = lowering.basis_index(out)
    check("API run sharded: same outcome on all ranks, non-zero weight (idx=%d, w=%.3e)" % (idx, wref[idx]),
          all(bool((t == zbits).all()) for t in allb) and wref[idx] > 0)
    lowering.release_sharded_states()
    # the same call on ONE GPU (every rank on its own unsharded copy) under the same seed returns the same bits
    os.environ["QSV_SHARD"] = "0"
    random.seed(100)
    with contextlib.redirect_stdout(io.StringIO()):
        out1 = c.run([0] * n, 2)
    os.environ.pop("QSV_SHARD")
    check("API run: %d-GPU outcome == 1-GPU outcome under the same seed (%s)" % (comm.world, "".join(map(str, out1))),
          out1 == out)

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("dist_check: all passed")


if __name__ == "__main__":
    main()
