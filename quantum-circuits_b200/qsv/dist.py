"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch).

The register of n qubits is sharded by its top g = log2(P) physical positions: rank r holds the
2^(n-g) amplitudes whose top g index bits equal r.  Everything local, diagonal, or merely
controlled by a global qubit runs without communication (the rank index supplies the bit).  The only
exchange step is the global<->local qubit swap: the g global positions trade places with the top g
local positions, which is exactly an all-to-all of equal blocks — rank r sends its block j to rank j
and stores rank j's block r in slot j.  At 36 qubits on 8 GPUs a shard is 137 GB of 180 GB, so the
exchange is done in place in column chunks through a small staging buffer.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _cabi

DEFAULT_STAGING_BYTES = 1 << 29      # per staging buffer (two in flight per direction)


def plan_exchange(n_local, g, staging_bytes=DEFAULT_STAGING_BYTES):
    """Chunk plan of one global<->local swap.  The shard is a [P, S] matrix of amplitudes (P = 2^g
    blocks of S = 2^(n_local-g)); chunk k covers columns [k*C, (k+1)*C) of every block and is
    exchanged with one all-to-all of P*C amplitudes.  Returns (P, S, C, n_chunks)."""
    P = 1 << g
    S = 1 << (n_local - g)
    C_ = max(1, min(S, staging_bytes // (16 * max(1, P - 1))))
    C_ = 1 << (C_.bit_length() - 1)               # power of two: divides S
    return P, S, C_, S // C_


_DEFAULT_COMM = None


def default_comm():
    """The communicator the public API (main.circuit.Circuit.run / run_method2) shards over: the default
    process group when the program was launched with one process per GPU (torchrun, NCCL backend, 2^g ranks),
    else None.  QSV_SHARD=0 keeps every rank on its own unsharded copy."""
    global _DEFAULT_COMM
    import os
    if os.environ.get("QSV_SHARD", "1") == "0":
        return None
    if not (dist.is_available() and dist.is_initialized()):
        return None
    world = dist.get_world_size()
    if world < 2 or world & (world - 1) or dist.get_backend() != "nccl":
        return None
    if _DEFAULT_COMM is None or _DEFAULT_COMM.group is not dist.group.WORLD:
        _DEFAULT_COMM = Comm(dist.group.WORLD)
    return _DEFAULT_COMM


class Comm:
    """Thin wrapper over a torch.distributed process group of size 2^g."""

    def __init__(self, group=None, staging_bytes=DEFAULT_STAGING_BYTES):
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.global_bits = (self.world - 1).bit_length()
        if self.world != 1 << self.global_bits:
            raise RuntimeError("qsv.dist: world size must be a power of two")
        self.staging_bytes = staging_bytes
        self._staging = None
        self._symm = None                # (tensor, handle) of the symmetric-memory state, if enabled
        self.peer_swap = False
        self.bytes_sent = 0
        self.swaps = 0
        self.swap_events = []            # (start, end) CUDA events of every swap, for bench.py

    # ------------------------------------------------------------------ peer memory
    def alloc_state(self, n_amps, device):
        """State buffer for a sharded register.  With QSV_SWAP=peer (default on NCCL) the buffer is
        symmetric memory mapped into every rank of the node (torch.distributed._symmetric_memory), so the
        qubit swap can run as one kernel over NVLink peer pointers; QSV_SWAP=nccl, or any failure to map,
        falls back to a plain buffer + the chunked NCCL all-to-all."""
        import os
        want = os.environ.get("QSV_SWAP", "peer")
        if want == "peer" and self.world > 1 and dist.get_backend(self.group) == "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                raw = symm_mem.empty(2 * n_amps, dtype=torch.float64, device=device)
                hdl = symm_mem.rendezvous(raw, self.group)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                if len(ptrs) != self.world or any(p == 0 for p in ptrs):
                    raise RuntimeError("incomplete peer mapping")
                self._symm = (raw, hdl, ptrs)
                self.peer_swap = True
                return torch.view_as_complex(raw.view(-1, 2))
            except Exception as e:                      # noqa: BLE001 - stay on the NCCL path, but say so
                if self.rank == 0:
                    print("qsv.dist: symmetric memory unavailable (%s); using the NCCL all-to-all" % (e,), flush=True)
                self._symm = None
                self.peer_swap = False
        return torch.empty(n_amps, dtype=torch.complex128, device=device)

    def _moved_bytes(self, sv, positions, gbits, support):
        """Bytes this rank sends in a swap of local `positions` against global bits `gbits`: 16 B for every element
        that changes rank, (1 - 2^-j) of the shard, and - with a support hint - only for the pairs of which at least
        one side can be non-zero (the kernels skip the others).  Exact: only the swapped bits and the pinned bits
        decide, every other local bit is a factor of two."""
        nl, j = sv.n_local, len(positions)
        mask, val = support if support is not None else (0, 0)
        local_mask = (1 << nl) - 1
        swapped_g = sum(1 << (nl + b) for b in gbits)
        if ((self.rank << nl) ^ val) & mask & ~local_mask & ~swapped_g:
            return 0                    # a pinned global bit outside the swap: the whole shard is zero on both sides
        pos_mask = sum(1 << p for p in positions)
        free = nl - j - bin(mask & local_mask & ~pos_mask).count("1")
        rg = [(self.rank >> b) & 1 for b in gbits]

        def in_support(pbits, gvals):
            for k in range(j):
                gp = nl + gbits[k]
                if (mask >> positions[k]) & 1 and pbits[k] != (val >> positions[k]) & 1:
                    return False
                if (mask >> gp) & 1 and gvals[k] != (val >> gp) & 1:
                    return False
            return True

        count = 0
        for a in range(1 << j):
            ab = [(a >> k) & 1 for k in range(j)]
            if ab != rg and (in_support(ab, rg) or in_support(rg, ab)):
                count += 1
        return (16 * count) << free

    def _swap_peer(self, sv, n_swap, positions=None, gbits=None, support=None):
        raw, hdl, ptrs = self._symm
        arr = (C.c_uint64 * self.world)(*ptrs)
        pos = _cabi.int_array(positions) if positions is not None else None
        gb = _cabi.int_array(gbits) if gbits is not None else None
        sp_mask, sp_val = support if support is not None else (0, 0)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        hdl.barrier()                                   # every rank's earlier kernels are done
        ev[0].record()                                  # (after the barrier: a rank that arrived early is not timed waiting)
        _cabi.check(sv.lib.qsv_swap_peer_ex(sv._ptr(sv.state), arr, sv.n_local, self.world, self.rank, int(n_swap), pos,
                                            gb, int(sp_mask), int(sp_val), sv._stream()), "qsv_swap_peer_ex")
        ev[1].record()
        hdl.barrier()                                   # nobody touches its shard before all swaps landed
        self.swap_events.append(ev)
        self.bytes_sent += self._moved_bytes(sv, list(positions) if positions is not None else
                                             list(range(sv.n_local - n_swap, sv.n_local)),
                                             list(gbits) if gbits is not None else list(range(n_swap)), support)
        self.swaps += 1

    def _flag_arrays(self, device):
        """Peer-mapped flag arrays of the fused swap + pass kernel: per rank 16 x 1024 uint64, zeroed once."""
        if getattr(self, "_flags", None) is None:
            import torch.distributed._symmetric_memory as symm_mem
            raw = symm_mem.empty(16 * 1024, dtype=torch.int64, device=device)
            raw.zero_()
            hdl = symm_mem.rendezvous(raw, self.group)
            torch.cuda.synchronize()
            hdl.barrier()
            self._flags = (raw, hdl, [int(p) for p in hdl.buffer_ptrs])
            self._epoch = 0
        return self._flags

    def swap_pass(self, sv, positions, cpass, prog, tables_ptr, pass_first=0, gbits=None, support=None, fresh=None):
        """Global<->local swap of `positions` + the gate pass that follows it as ONE kernel over peer memory
        (qsv_run_pass_rounds_swap): tiles are pulled from the peers, transformed and stored locally; a per-tile flag
        handshake keeps the exchange in place.  One barrier before (earlier kernels of all ranks are done), none after."""
        if support is None:
            sv._basis = None
            sv._presum = None
        self._check_owner(sv)
        raw, hdl, ptrs = self._symm
        fraw, fhdl, fptrs = self._flag_arrays(sv.state.device)
        gb = list(gbits) if gbits is not None else list(range(len(positions)))
        sp_mask, sp_val = support if support is not None else (0, 0)
        self._epoch += 1
        arr = (C.c_uint64 * self.world)(*ptrs)
        farr = (C.c_uint64 * self.world)(*fptrs)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        hdl.barrier()
        ev[0].record()
        if fresh is not None:       # the register was not cleared before the run: zero-fill form (fresh = zf_mask, zf_val)
            _cabi.check(sv.lib.qsv_run_pass_rounds_swap_fresh(
                sv._ptr(sv.state), C.byref(cpass), C.c_void_p(prog.ctypes.data), int(prog.size), tables_ptr, arr, farr,
                self.world, len(positions), _cabi.int_array(positions), _cabi.int_array(gb), int(pass_first), self.rank,
                self._epoch << 32, int(sp_mask), int(sp_val), int(fresh[0]), int(fresh[1]), sv._stream()),
                "qsv_run_pass_rounds_swap_fresh")
        else:
            _cabi.check(sv.lib.qsv_run_pass_rounds_swap_ex(sv._ptr(sv.state), C.byref(cpass), C.c_void_p(prog.ctypes.data),
                                                           int(prog.size), tables_ptr, arr, farr, self.world,
                                                           len(positions), _cabi.int_array(positions), _cabi.int_array(gb),
                                                           int(pass_first), self.rank, self._epoch << 32, int(sp_mask),
                                                           int(sp_val), sv._stream()), "qsv_run_pass_rounds_swap_ex")
        ev[1].record()
        self.swap_events.append(ev)
        self.fused_swaps = getattr(self, "fused_swaps", 0) + 1
        self.bytes_sent += self._moved_bytes(sv, list(positions), gb, support)
        self.swaps += 1

    def _check_owner(self, sv):
        """The peer pointers on this communicator belong to ONE register (the last one alloc_state made): a kernel
        over peer memory launched for any other StateVector would read another register's shards."""
        if self._symm is None or sv.state.data_ptr() != self._symm[2][self.rank]:
            raise RuntimeError("qsv.dist: this StateVector does not own the communicator's symmetric-memory state "
                               "(a newer sharded register was allocated on the same Comm, or release_state() ran)")

    def release_state(self):
        """Drop the symmetric-memory state (the StateVector that used it must be dropped as well)."""
        self._symm = None
        self.peer_swap = False

    # ------------------------------------------------------------------ collectives
    def broadcast_float(self, value, src=0):
        """The same float64 on every rank (rank `src`'s): one uniform draw per run, shared by all shards."""
        t = torch.tensor([float(value)], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
        dist.broadcast(t, dist.get_global_rank(self.group, src), group=self.group)
        return float(t.item())

    def all_gather_local(self, t):
        """Concatenation of every rank's 1-D tensor in rank order (physical index order of a sharded register)."""
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def all_reduce_sum(self, t):
        """In-place sum of a complex128 / float64 tensor over all ranks (NCCL has no complex type)."""
        v = torch.view_as_real(t) if t.is_complex() else t
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_to_all(self, out, inp, async_op=False):
        """All-to-all of P equal blocks where a rank's OWN block is not transferred (it never moves in a
        qubit swap): `inp`/`out` hold P-1 blocks, one per peer in rank order.  gloo (CPU tests) has no
        all_to_all -> send/recv.  With async_op the NCCL work handle is returned (wait() orders the
        current stream after it)."""
        P = self.world
        blk = inp.numel() // (P - 1)
        if dist.get_backend(self.group) == "nccl":
            splits = [blk if j != self.rank else 0 for j in range(P)]
            return dist.all_to_all_single(out, inp, output_split_sizes=splits, input_split_sizes=splits,
                                          group=self.group, async_op=async_op)
        ins, outs = inp.chunk(P - 1), out.chunk(P - 1)
        peers = [j for j in range(P) if j != self.rank]
        reqs = []
        for i, peer in enumerate(peers):
            reqs.append(dist.isend(ins[i].contiguous(), dist.get_global_rank(self.group, peer), group=self.group))
            reqs.append(dist.irecv(outs[i], dist.get_global_rank(self.group, peer), group=self.group))
        for r in reqs:
            r.wait()
        return None

    # ------------------------------------------------------------------ qubit swap
    def exchange(self, state, n_local, pack, unpack):
        """Swap the g global positions with the top g local positions of `state` (complex128 tensor of
        2^n_local amplitudes), in place, chunk by chunk.  pack(staging, offset, C) gathers column chunk
        [offset, offset+C) of every block into the contiguous staging tensor; unpack scatters it back."""
        g = self.global_bits
        P, S, C_, n_chunks = plan_exchange(n_local, g, self.staging_bytes)
        need = (P - 1) * C_              # the rank's own block stays where it is
        if self._staging is None or self._staging[0].numel() < need or self._staging[0].device != state.device:
            self._staging = tuple(torch.empty(need, dtype=torch.complex128, device=state.device) for _ in range(4))
        s_in = [self._staging[0][:need], self._staging[1][:need]]
        s_out = [self._staging[2][:need], self._staging[3][:need]]
        ev = None
        if state.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        # two-deep pipeline: while chunk k is on the wire (NCCL stream), the compute stream unpacks chunk
        # k-1 and packs chunk k+1.  Staging buffers alternate; a buffer is reused only after the collective
        # that read / wrote it has been waited for.
        handles = [None, None]
        for k in range(n_chunks + 1):
            if k < n_chunks:
                b = k & 1
                pack(s_in[b], k * C_, C_)
                handles[b] = self.all_to_all(torch.view_as_real(s_out[b]).view(-1),
                                             torch.view_as_real(s_in[b]).view(-1), async_op=True)
            if k >= 1:
                pb = (k - 1) & 1
                if handles[pb] is not None:
                    handles[pb].wait()
                unpack(s_out[pb], (k - 1) * C_, C_)
        if ev is not None:
            ev[1].record()
            self.swap_events.append(ev)
        self.bytes_sent += 16 * (P - 1) * S
        self.swaps += 1

    def swap_stats(self):
        """Total device time of the recorded swaps (ms) and bytes sent per GPU; clears the record."""
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self.swap_events)
        n = len(self.swap_events)
        self.swap_events = []
        return n, ms

    def swap_global_local(self, sv, n_swap, positions=None, gbits=None, support=None):
        """Engine hook for a 'global_swap' op: the peer-memory kernel (any local positions, any subset of the global
        bits, support hint), or the chunked NCCL all-to-all with the CUDA pack/unpack kernels of libqsv (all global
        qubits against the top local positions only)."""
        if support is None:
            # called outside StateVector.execute(): whatever was known about the support / weights of the state is void
            sv._basis = None
            sv._presum = None
        if self.peer_swap and self._symm is not None and sv.state.data_ptr() == self._symm[2][self.rank]:
            return self._swap_peer(sv, n_swap, positions, gbits, support)
        if n_swap != self.global_bits or (gbits is not None and list(gbits) != list(range(n_swap))):
            raise RuntimeError("qsv.dist: the NCCL exchange swaps all %d global qubits at once; partial swaps need the "
                               "peer-memory kernel" % self.global_bits)
        top = list(range(sv.n_local - self.global_bits, sv.n_local))
        if positions is not None and list(positions) != top:
            raise RuntimeError("qsv.dist: the NCCL exchange swaps the top local positions only; lower the circuit "
                               "with free_swap=False when peer memory is unavailable")
        lib = sv.lib
        P = 1 << self.global_bits
        S = 1 << (sv.n_local - self.global_bits)

        r = self.rank
        base = sv.state.data_ptr()

        # the blocks below / above the rank's own block are packed as two runs of slices
        def pack(staging, offset, C_):
            sp = staging.data_ptr()
            if r > 0:
                _cabi.check(lib.qsv_pack_slices(C.c_void_p(base), C.c_void_p(sp), r, S, offset, C_, sv._stream()),
                            "qsv_pack_slices")
            if r < P - 1:
                _cabi.check(lib.qsv_pack_slices(C.c_void_p(base + 16 * (r + 1) * S), C.c_void_p(sp + 16 * r * C_),
                                                P - 1 - r, S, offset, C_, sv._stream()), "qsv_pack_slices")

        def unpack(staging, offset, C_):
            sp = staging.data_ptr()
            if r > 0:
                _cabi.check(lib.qsv_unpack_slices(C.c_void_p(base), C.c_void_p(sp), r, S, offset, C_, sv._stream()),
                            "qsv_unpack_slices")
            if r < P - 1:
                _cabi.check(lib.qsv_unpack_slices(C.c_void_p(base + 16 * (r + 1) * S), C.c_void_p(sp + 16 * r * C_),
                                                  P - 1 - r, S, offset, C_, sv._stream()), "qsv_unpack_slices")

        self.exchange(sv.state, sv.n_local, pack, unpack)

    # ------------------------------------------------------------------ measurement
    def all_reduce_int64(self, t):
        """In-place integer sum over the ranks (bucket arrays of the exact sampler: order-independent)."""
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def sample(self, sv, u, src_pos=None):
        """One outcome for the uniform draw u on a sharded register: StateVector.sample runs the exact radix descent
        in OUTPUT order, all-reducing the bucket sums of every level over the ranks, so every rank ends with the
        same index - the one an unsharded register returns for the same u (main/circuit.py:360-364)."""
        return sv.sample(u, src_pos)
