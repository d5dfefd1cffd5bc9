"""Device-side execution: owns the complex128 state in HBM and drives libqsv through the C-ABI.

torch is used for plumbing only (device memory, the current CUDA stream, torch.distributed); every
amplitude is touched by the hand-written kernels of csrc/.  There is no CPU path: constructing a
StateVector without CUDA raises RuntimeError.
"""
import ctypes as C
import numpy as np
import torch

from . import _cabi
from .schedule import Op, TILED_KINDS, Support, plan_passes, pack_passes, swap_fields, common_external_controls

SCAN_BITS_EXACT = 20     # up to 2^20 outcomes: one fully sequential scan (bit-exact with random.choices)
SAMPLE_BLOCK_BITS = 12   # larger states: block sums of 2^12 outcomes, then a sequential scan inside one block


PRESUM_MIN_BITS = 3      # the last pass weighs the state for the sampler if it can fix at least this many top output bits
PRESUM_SWEEP_MS = 4.7                        # the sampler's own first-level sweep of 2^30 amplitudes
PRESUM_STATIC_MS = (0.5, 0.6, 1.2, 2.0)      # sums variant over the plain pass, by bucket bits inside the tile (static form)
PRESUM_GENERAL_MS = (0.5, 0.7, 2.4, 6.3)     # ... general form
_NO_SAMPLE = object()    # "no sample follows" (None is a valid src_pos: the identity map)

_PLAN_CACHE = {}      # op-list fingerprint -> [("program", PackedProgram) | ("op", Op)] (host objects only)
_KERNELS_READY = set()   # fingerprints whose specialised kernels are already in libqsv's cache (skip the compile pool)


def _ops_fingerprint(ops, geometry):
    """Content hash of an op list (kinds, positions, control values, matrices, scalars, parameters) plus the
    tile geometry and the planner switches: equal fingerprints -> identical schedules and programs."""
    import hashlib
    import os
    h = hashlib.blake2b(digest_size=16)
    h.update(repr((geometry, os.environ.get("QSV_PLAN_SEARCH"), os.environ.get("QSV_REGISTER_KERNEL"),
                   os.environ.get("QSV_REG_BITS"), os.environ.get("QSV_REG_MIN_OPS"), os.environ.get("QSV_JIT"),
                   os.environ.get("QSV_JIT_MIN_QUBITS"), os.environ.get("QSV_TILE_BITS"),
                   os.environ.get("QSV_LOW_BITS"), os.environ.get("QSV_DEBUG_DROP"), os.environ.get("QSV_FILL"), os.environ.get("QSV_FUSE_SWAP"),
                   os.environ.get("QSV_TRIM_ROUNDS"), os.environ.get("QSV_PLAN_RESTARTS"),
                   os.environ.get("QSV_PLAN_NATIVE"), os.environ.get("QSV_SPARSE_START"))).encode())
    cached = getattr(ops, "digest", None)       # lowering.OpList: the digest the list was cached under
    if cached is not None:
        h.update(cached)
        return h.digest()
    for op in ops:
        h.update(repr((op.kind, tuple(op.targets), tuple(op.controls),
                       None if op.cvals is None else tuple(int(v) for v in op.cvals),
                       complex(op.scalar), sorted(op.params.items()) if op.params else None)).encode())
        if op.matrix is not None:
            h.update(np.ascontiguousarray(op.matrix, dtype=np.complex128).tobytes())
        if op.perm is not None:
            h.update(np.ascontiguousarray(op.perm).tobytes())
    return h.digest()


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("qsv: no CUDA device available - the state-vector engine has no CPU fallback")


class StateVector:
    """2^n_local complex128 amplitudes of one shard, plus the operations of the hot path.

    n        total qubits of the register
    comm     optional qsv.dist.Comm (rank, world size 2^g); the shard holds the amplitudes whose top g
             index bits equal the rank
    """

    def __init__(self, n, device=None, comm=None):
        _require_cuda()
        self.lib = _cabi.load()
        self.n = int(n)
        self.comm = comm
        self.g = comm.global_bits if comm is not None else 0
        self.rank = comm.rank if comm is not None else 0
        self.n_local = self.n - self.g
        if self.n_local < 1:
            raise RuntimeError("qsv: need at least one local qubit per shard")
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        if comm is not None and comm.world > 1:
            self.state = comm.alloc_state(1 << self.n_local, self.device)
        else:
            self.state = torch.empty(1 << self.n_local, dtype=torch.complex128, device=self.device)
        self._scratch = None
        self._keep = []       # host/device buffers that must outlive the async launches
        self.passes_run = 0
        self._basis = None
        self._presum = None       # (k, src_pos tuple): the sampler's buckets hold the sums of the top k output bits
        self._sampler = None
        self.d2h_bytes = 0
        self.tile_bits = None
        self.low_bits = None

    # ------------------------------------------------------------------ helpers
    @property
    def rank_bits(self):
        return self.rank << self.n_local

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ptr(self, t):
        return C.c_void_p(t.data_ptr())

    def synchronize(self):
        torch.cuda.current_stream(self.device).synchronize()
        self._keep.clear()

    # ------------------------------------------------------------------ state preparation
    def init_basis(self, index):
        """|index>: main/circuit.py:290-301."""
        _cabi.check(self.lib.qsv_init_basis(self._ptr(self.state), self.n_local, int(index), self.rank,
                                            self._stream()), "qsv_init_basis")
        self._presum = None
        self._basis = int(index)         # the next execute() knows the support of the state (schedule.Support)

    def init_columns(self, n, flip=0):
        """Every basis input of an n-qubit circuit at once: this register has 2n qubits, column c sits under the
        high half = c with the low half in |c ^ flip> (Circuit.get_subcircuit, main/circuit.py:122-157)."""
        if self.n != 2 * n or self.g:
            raise RuntimeError("qsv: init_columns needs an unsharded register of 2n qubits")
        _cabi.check(self.lib.qsv_init_columns(self._ptr(self.state), int(n), int(flip), self._stream()),
                    "qsv_init_columns")
        self._basis = None               # 2^n non-zero amplitudes: not a basis state, no support hint
        self._presum = None

    # ------------------------------------------------------------------ gate application
    def upload_program(self, ops):
        """Plan + pack a run of tiled ops and copy the image to the device.  Returns a handle for
        run_program (so benchmarks can separate host scheduling from device time)."""
        return self.upload_passes(plan_passes(ops, self.n_local, self.tile_bits, self.low_bits))

    def upload_passes(self, passes):
        packed = pack_passes(passes)
        self._prepare_kernels(packed)
        host = torch.from_numpy(packed.blob)
        dev = host.to(self.device, non_blocking=False)
        return (packed, dev)

    def _cpass(self, ps):
        cp = _cabi.QsvPass()
        cp.n_local, cp.tile_bits, cp.low_bits = ps.n_local, ps.tile_bits, ps.low_bits
        cp.n_hi = ps.tile_bits - ps.low_bits
        for j, p in enumerate(ps.hi_pos):
            cp.hi_pos[j] = p
        cp.rank_bits = self.rank_bits
        return cp

    def _prepare_kernels(self, packed):
        """Large states run pass-specialised kernels (csrc/qsv_jit.cu).  Generate + compile the ones this
        program needs now, on all host cores (NVRTC, ~0.3 s per kernel, cached by program bytes inside
        libqsv), so that the launches only look them up."""
        todo = []
        for ps, _, n_ops, extra in packed.passes:
            if n_ops >= 0:
                continue
            prog = extra[0]
            cp = self._cpass(ps)
            cp.reserved = int(prog[12])
            if self.lib.qsv_jit_wanted(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size)):
                todo.append((cp, prog))
        if not todo:
            return

        def compile_one(item):
            cp, prog = item
            return self.lib.qsv_jit_compile(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), None, 0, None)

        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4, 32)) as ex:
            list(ex.map(compile_one, todo))     # failures surface at launch time (with the fallback policy)

    def run_program(self, handle):
        packed, dev = handle
        self._basis = None          # the state is about to change outside execute(): the support hint is void
        for meta in packed.passes:
            self._launch_pass(meta, dev)
            self.passes_run += 1
        self._keep.append(dev)

    def _adopt(self, out):
        """Make `out` the state after an out-of-place op.  A symmetric-memory state keeps its buffer
        (peers hold its address): the result is copied back instead of swapping references."""
        self._basis = None
        self._presum = None
        if self._owns_symm():
            self.state.copy_(out)
        else:
            self.state, self._scratch = out, self.state

    def _owns_symm(self):
        """True when this register's buffer is the communicator's peer-mapped symmetric memory (a Comm serves one
        sharded register at a time; an older StateVector on the same Comm must not use its peer pointers)."""
        symm = getattr(self.comm, "_symm", None) if self.comm is not None else None
        return bool(symm is not None and getattr(self.comm, "peer_swap", False)
                    and self.state.data_ptr() == symm[2][self.rank])

    def _scratch_state(self):
        if self._scratch is None:
            self._scratch = torch.empty_like(self.state)
        return self._scratch

    def _run_whole_state_op(self, op, support=None):
        self._presum = None
        if support is None:
            self._basis = None      # called outside execute(): nothing is known about the support any more
        st = self._stream()
        if op.kind == "wide_dense":
            M = torch.from_numpy(np.ascontiguousarray(op.matrix, dtype=np.complex128)).to(self.device)
            out = self._scratch_state()
            _cabi.check(self.lib.qsv_apply_dense_wide(self._ptr(self.state), self._ptr(out), self.n_local,
                                                      len(op.targets), _cabi.int_array(op.targets), self._ptr(M),
                                                      st), "qsv_apply_dense_wide")
            self._adopt(out)
            self._keep.append(M)
        elif op.kind == "wide_perm":
            P = torch.from_numpy(np.ascontiguousarray(op.perm, dtype=np.uint32).view(np.int32)).to(self.device)
            out = self._scratch_state()
            _cabi.check(self.lib.qsv_apply_perm_wide(self._ptr(self.state), self._ptr(out), self.n_local,
                                                     len(op.targets), _cabi.int_array(op.targets), self._ptr(P),
                                                     st), "qsv_apply_perm_wide")
            self._adopt(out)
            self._keep.append(P)
        elif op.kind == "wide_diag":
            D = torch.from_numpy(np.ascontiguousarray(op.matrix, dtype=np.complex128)).to(self.device)
            _cabi.check(self.lib.qsv_apply_diag_wide(self._ptr(self.state), self.n_local, len(op.targets),
                                                     _cabi.int_array(op.targets), self._ptr(D), st),
                        "qsv_apply_diag_wide")
            self._keep.append(D)
        elif op.kind == "modexp":
            p = op.params          # controls = x register (MSB first, may be global), targets = y register
            _cabi.check(self.lib.qsv_modexp_involution(self._ptr(self.state), self.n_local, len(op.controls),
                                                       _cabi.int_array(op.controls), len(op.targets),
                                                       _cabi.int_array(op.targets), int(p["a"]), int(p["N"]),
                                                       self.rank_bits, st), "qsv_modexp_involution")
        elif op.kind == "oracle":
            self._oracle(op)
        elif op.kind == "diffusion":
            self._diffusion(op)
        elif op.kind == "global_swap":
            if self.comm is None:
                raise RuntimeError("qsv: global_swap op without a communicator")
            pos, gb = swap_fields(op, self.n_local)
            self.comm.swap_global_local(self, len(pos), op.params.get("positions") and pos, gb, support=support)
        else:
            raise RuntimeError("qsv: unknown op kind %r" % op.kind)

    def _oracle(self, op):
        """Grover oracle for one marked item (main/Grover.py:14-22): exchange the two amplitudes whose
        search-register bits equal x.  params: idx0, idx1 = global indices of the pair."""
        cvals = op.cvals if op.cvals is not None else (1,) * len(op.controls)
        i0 = 0
        for c, v in zip(op.controls, cvals):
            i0 |= (1 if v else 0) << c
        anc = op.targets[0]
        if anc >= self.n_local:
            raise RuntimeError("qsv: oracle pair straddles shards; the target qubit must be local")
        i1 = i0 | (1 << anc)
        if (i0 >> self.n_local) != self.rank:
            return
        m = (1 << self.n_local) - 1
        _cabi.check(self.lib.qsv_grover_oracle(self._ptr(self.state), i0 & m, i1 & m, self._stream()),
                    "qsv_grover_oracle")

    def _diffusion(self, op):
        """2|s><s| - I on all positions >= anc_bits (main/Grover.py:24-31,40-44)."""
        anc = int(op.params["anc_bits"])
        need = self.lib.qsv_diffusion_scratch_bytes(self.n_local, anc)
        scratch = torch.empty(max(16, need), dtype=torch.uint8, device=self.device)
        sums = torch.empty(1 << anc, dtype=torch.complex128, device=self.device)
        st = self._stream()
        _cabi.check(self.lib.qsv_diffusion_sums(self._ptr(self.state), self.n_local, anc, self._ptr(scratch),
                                                self._ptr(sums), st), "qsv_diffusion_sums")
        if self.comm is not None and self.comm.world > 1:
            self.comm.all_reduce_sum(sums)
        if sorted(op.targets) != list(range(anc, self.n)):
            raise RuntimeError("qsv: diffusion needs the search register on positions [%d, %d)" % (anc, self.n))
        _cabi.check(self.lib.qsv_diffusion_update(self._ptr(self.state), self.n_local, anc, self._ptr(sums),
                                                  self.n - anc, st), "qsv_diffusion_update")
        self._keep += [scratch, sums]

    def prepare(self, ops):
        """Host scheduling + upload: the op list becomes passes (planned across whole-state ops and qubit
        swaps, qsv.schedule.plan_schedule) whose programs are packed and copied to HBM, interleaved with the
        whole-state ops.  The returned plan can be executed any number of times.  Planning + packing is
        cached by the content of the op list (Shor's sampling loop and benchmarks re-run one circuit)."""
        key = _ops_fingerprint(ops, (self.n_local, self.tile_bits, self.low_bits, self.fuse_swaps(), self.g))
        host_plan = _PLAN_CACHE.get(key)
        if host_plan is None:
            from .schedule import plan_schedule_search
            host_plan = []
            batch = []
            for kind, item in plan_schedule_search(ops, self.n_local, self.tile_bits, self.low_bits,
                                                   fuse_swaps=self.fuse_swaps(), n_global=self.g):
                if kind == "pass":
                    batch.append(item)
                    continue
                if batch:
                    host_plan.append(("program", pack_passes(batch)))
                    batch = []
                if kind == "swap_pass":
                    swap_op, ps, pass_first = item
                    packed = pack_passes([ps], min_ops=0)   # a round program whatever its size: only those can ride

                    if len(packed.passes) == 1 and self._jit_runs(packed.passes[0]):
                        host_plan.append(("swap_pass", (swap_op, packed, pass_first)))
                    elif pass_first:            # not a single specialised kernel: the two steps on their own
                        host_plan.append(("program", packed))
                        host_plan.append(("op", swap_op))
                    else:
                        host_plan.append(("op", swap_op))
                        host_plan.append(("program", packed))
                    continue
                host_plan.append(("op", item))
            if batch:
                # the plan ends in a pass: a round program whatever its size, so that it can weigh the state for the
                # sampler (a lone one- or two-op pass stays on the shared-memory kernel, which is faster for it)
                host_plan.append(("program", pack_passes(batch, last_as_rounds=bool(host_plan) or len(batch) > 1)))
            if len(_PLAN_CACHE) >= 16:
                _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
            _PLAN_CACHE[key] = host_plan
        segments = []
        h2d = 0
        compiled = key in _KERNELS_READY        # this process has already generated + compiled this plan's kernels
        for kind, item in host_plan:
            if kind == "program":
                if not compiled:
                    self._prepare_kernels(item)
                dev = torch.from_numpy(item.blob).to(self.device, non_blocking=False)
                segments.append(("program", (item, dev)))
                h2d += int(item.blob.size)
                h2d += sum(int(m[3][0].size) for m in item.passes if m[2] < 0)     # by-value round programs
            elif kind == "swap_pass":
                swap_op, packed, pass_first = item
                dev = torch.from_numpy(packed.blob).to(self.device, non_blocking=False)
                h2d += int(packed.blob.size) + int(packed.passes[0][3][0].size)
                if self._compile_fused(packed, swap_fields(swap_op, self.n_local), pass_first):
                    segments.append(("swap_pass", (swap_op, packed, dev, pass_first)))
                else:
                    # the fused kernel cannot be built (same answer on every rank: the program is rank independent):
                    # run the two steps on their own rather than leave the peers waiting for a kernel that never starts
                    self._prepare_kernels(packed)
                    two = [("program", (packed, dev)), ("op", swap_op)]
                    segments.extend(two if pass_first else two[::-1])
            else:
                segments.append(("op", item))
        if len(_KERNELS_READY) >= 64:
            _KERNELS_READY.clear()
        _KERNELS_READY.add(key)
        return {"segments": segments, "h2d_bytes": h2d}

    def execute(self, plan, sample_src_pos=_NO_SAMPLE):
        """Run a prepared plan on the current state.  Right after init_basis the support of the state is known and
        the passes skip the tiles that are still zero (schedule.Support); any other state is treated as dense.
        sample_src_pos (a src_pos list or None for the identity map): a sample(u, src_pos) with that output wiring
        follows - if the plan ends in a pass-specialised kernel, that pass also adds up the weights the sampler's
        first level needs (qsv_run_pass_rounds_sums), saving its read of the whole state."""
        basis, self._basis = self._basis, None
        self._run_segments(plan, basis, 0, sample_src_pos)

    def _run_segments(self, plan, basis, n_fresh, sample_src_pos=_NO_SAMPLE):
        """The launches of a plan; the first n_fresh launch groups (passes, swaps, fused swap+pass kernels) run in
        the zero-fill form (run_from_basis: the register was not cleared)."""
        items = list(self._walk(plan, basis))
        presum = self._presum_plan(items, sample_src_pos)
        done = 0
        self.mark("init")
        for i, (kind, seg, sp, zf) in enumerate(items):
            if i:
                self.mark(items[i - 1][0] if items[i - 1][0] != "op" else items[i - 1][1].kind)
            if kind == "pass":
                meta, dev = seg
                fresh = zf if i < n_fresh else None
                if presum is not None and i == len(items) - 1:
                    self._launch_pass_sums(meta, dev, sp, fresh, presum)
                else:
                    self._launch_pass(meta, dev, sp, fresh)
                done += 1
                self.passes_run += 1
                self._keep.append(dev)
            elif kind == "swap_pass":
                self._run_swap_pass(seg, sp, zf if i < n_fresh else None)
            else:
                self._run_whole_state_op(seg, sp if sp is not None else (0, 0))
        if items:
            self.mark(items[-1][0] if items[-1][0] != "op" else items[-1][1].kind)
        self._basis = None
        self._presum = (presum[0], presum[2]) if presum is not None else None

    def mark(self, label):
        """Step timeline (bench.py, N > 1): with `trace` set to a list, an event is recorded on the stream after every
        launch group of a run - the gaps between consecutive events are what each group cost THIS rank, waits at the
        swaps' barriers included.  Off (None) by default: nothing is recorded."""
        tr = getattr(self, "trace", None)
        if tr is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.device))
            tr.append((label, ev))

    def _presum_plan(self, items, sample_src_pos):
        """(k, positions, src tuple) if the plan's last launch can weigh the state for a sample in output wiring
        sample_src_pos: the top k output bits sit on positions outside that pass's tile; None otherwise."""
        import os
        if sample_src_pos is _NO_SAMPLE or not items or os.environ.get("QSV_PRESUM", "1") == "0":
            return None
        sharded = self.comm is not None and self.comm.world > 1
        if not sharded and self.n_local <= SCAN_BITS_EXACT:
            return None                         # small registers sample with the sequential float scan
        kind, seg, sp, zf = items[-1]
        if kind != "pass" or not self._jit_runs(seg[0]):
            return None
        from .sampling import presum_bits
        ps = seg[0][0]
        if ps.tile_bits != 12:
            return None
        src = tuple(sample_src_pos) if sample_src_pos is not None else tuple(range(self.n))
        # What the sums cost on top of the plain pass (ms per 2^30 amplitudes, B200, bench.py step_timeline): the weights
        # themselves ~0.5, every bucket bit INSIDE the tile more - little when the kernel classifies with
        # generation-time constants (qsv_jit_sums_static), a lot when it reads every amplitude's bucket from its final
        # index.  What they save: the sampler's first sweep (4.7 ms); the next level still reads 2^-k of the state.
        # The cheapest of the four limits wins; none may beat the sampler's own sweep.
        prog = seg[0][3][0]
        cp = self._cpass(ps)
        cp.reserved = int(prog[12])
        static = self.lib.qsv_jit_sums_static(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size)) == 1
        inside_ms = PRESUM_STATIC_MS if static else PRESUM_GENERAL_MS
        tile = list(range(ps.low_bits)) + list(ps.hi_pos)
        tile_set = set(tile)
        best, k = PRESUM_SWEEP_MS, 0
        if os.environ.get("QSV_PRESUM_COST", "1") == "0":       # tests / A-B runs: as many bits as the kernel can take
            best = float("inf")
            inside_ms = (3.0, 2.0, 1.0, 0.0)
        for limit in range(4):
            kk = presum_bits(src, self.n, tile, max_inside=limit)
            if kk < min(PRESUM_MIN_BITS, self.n):
                continue
            n_in = sum(1 for b in range(self.n - kk, self.n) if src[b] in tile_set)
            ms = inside_ms[n_in] + PRESUM_SWEEP_MS * 2.0 ** -kk
            if ms < best - 1e-9:
                best, k = ms, kk
        if not k:
            return None
        positions = [src[b] for b in range(self.n - k, self.n)]          # least significant bucket bit first
        return k, positions, src

    def _launch_pass_sums(self, packed_meta, dev, support, fresh, presum):
        ps, table_off, n_ops, extra = packed_meta
        prog = extra[0]
        cp = self._cpass(ps)
        cp.max_dense_k = 1
        cp.reserved = int(prog[12])
        if self._sampler is None:
            from .sampling import ExactSampler
            self._sampler = ExactSampler(self.lib, self.device)
        sp = support if support is not None else (0, 0)
        zf = fresh if fresh is not None else (0, 0)
        k, positions, _ = presum
        _cabi.check(self.lib.qsv_run_pass_rounds_sums(self._ptr(self.state), C.byref(cp), C.c_void_p(prog.ctypes.data),
                                                      int(prog.size), C.c_void_p(dev.data_ptr() + table_off),
                                                      int(sp[0]), int(sp[1]), int(zf[0]), int(zf[1]), k,
                                                      _cabi.int_array(positions),
                                                      C.c_void_p(self._sampler.buckets.data_ptr()), self._stream()),
                    "qsv_run_pass_rounds_sums")

    def run_from_basis(self, index, plan, sample_src_pos=_NO_SAMPLE):
        """init_basis(index) + execute(plan), without clearing the register where the plan allows it.

        The reference builds |in_v> entry by entry (main/circuit.py:290-301); init_basis clears 16 * 2^n bytes for it
        (17 GB = 2.6 ms of a 63 ms step at 30 qubits).  A plan that starts with pass-specialised kernels and reaches a
        pass touching every tile does not need the clear: only the basis amplitude is written (qsv_init_basis_sparse),
        the passes up to and including the first full sweep run in the zero-fill form (qsv_run_pass_rounds_fresh: what
        lies outside the support of the state is taken as zero, never read as data) and write every tile they touch
        in full; from then on the register is an ordinary dense state.  Sharded registers (peer-memory state) do the
        same: support-aware swaps move what lies outside the support around untouched, fused swap+pass kernels take
        tiles pulled from outside the support as zeros (qsv_run_pass_rounds_swap_fresh).  Everything else - plans
        with whole-state ops or interpreter passes before the first full sweep, QSV_SPARSE_START=0, QSV_LAZY_INIT=0 -
        takes init_basis + execute."""
        n_fresh = self._fresh_prefix(plan)
        if not n_fresh:
            self.init_basis(index)
            self.execute(plan, sample_src_pos)
            return
        _cabi.check(self.lib.qsv_init_basis_sparse(self._ptr(self.state), self.n_local, int(index), self.rank,
                                                   self._stream()), "qsv_init_basis_sparse")
        self._run_segments(plan, int(index), n_fresh, sample_src_pos)

    def _fresh_prefix(self, plan):
        """Number of leading launch groups that must run in the zero-fill form when the register is NOT cleared: up to
        and including the first pass (or fused swap+pass) that touches every tile of every shard; 0 if the plan does
        not qualify (see run_from_basis).  On a sharded register the groups before it may also be support-aware swaps
        (they move whatever lies outside the support around as it is: it stays outside) and fused swap+pass kernels
        (qsv_run_pass_rounds_swap_fresh); QSV_LAZY_INIT_SHARDED=0 keeps the clear there."""
        import os
        from .schedule import sparse_start_enabled
        if os.environ.get("QSV_LAZY_INIT", "1") == "0" or not sparse_start_enabled():
            return 0
        sharded = self.comm is not None and self.comm.world > 1
        if sharded and (os.environ.get("QSV_LAZY_INIT_SHARDED", "1") == "0" or not self._owns_symm()):
            return 0
        key = ("_fresh", os.environ.get("QSV_JIT"), os.environ.get("QSV_JIT_MIN_QUBITS"), sharded)
        if key in plan:
            return plan[key]
        count, ok = 0, False
        for kind, seg, sp, zf in self._walk(plan, 0):
            if kind == "pass":
                if not self._jit_runs(seg[0]):
                    break
            elif kind == "swap_pass":
                if not sharded:
                    break
            elif not (sharded and kind == "op" and seg.kind == "global_swap"):
                break
            count += 1
            if kind != "op" and not sp[0]:
                ok = True               # this launch touches every tile: the register is fully written after it
                break
        plan[key] = count if ok else 0
        return plan[key]

    def _walk(self, plan, basis):
        """Segments of a plan in execution order with the support hint of every launch:
        ("pass", (meta, dev), (fixed_mask, fixed_val), (zf_mask, zf_val)) | ("swap_pass", seg, (mask, val) before the
        swap, the pass's tile bits excluded, (zf_mask, zf_val)) | ("op", op, (mask, val) before a global_swap else None,
        None).
        (zf_mask, zf_val): the support's fixed bits INSIDE the pass's tile, tile-local (zero-fill form)."""
        sup = Support(self.n, basis)
        for kind, seg in plan["segments"]:
            if kind == "program":
                packed, dev = seg
                for meta in packed.passes:
                    yield "pass", (meta, dev), sup.pass_mask(meta[0]), sup.tile_mask(meta[0])
                    sup.after_ops(meta[0].ops)
            elif kind == "swap_pass":
                swap_op, packed, dev, pass_first = seg
                yield "swap_pass", seg, sup.pass_mask(packed.passes[0][0]), sup.tile_mask(packed.passes[0][0])
                if pass_first:
                    sup.after_ops(packed.passes[0][0].ops)
                    sup.swap_op(swap_op, self.n_local)
                else:
                    sup.swap_op(swap_op, self.n_local)
                    sup.after_ops(packed.passes[0][0].ops)
            elif seg.kind == "global_swap":
                yield "op", seg, sup.fixed(), None
                sup.swap_op(seg, self.n_local)
            else:
                yield "op", seg, None, None
                sup.dense()

    def fuse_swaps(self):
        """True when a global<->local swap and the pass after it may run as ONE kernel over peer memory
        (csrc/qsv_jit.cu, fused variant): peer-mapped state, specialised kernels in use, QSV_FUSE_SWAP != 0."""
        import os
        if os.environ.get("QSV_FUSE_SWAP", "1") == "0":
            return False
        if not self._owns_symm() or self.n_local < 12 + self.g:
            return False
        mode = os.environ.get("QSV_JIT")
        if mode == "0":
            return False
        return mode == "force" or self.n_local >= int(os.environ.get("QSV_JIT_MIN_QUBITS", "28"))

    def _compile_fused(self, packed, positions, pass_first):
        ps, _, _, extra = packed.passes[0]
        prog = extra[0]
        cp = self._cpass(ps)
        cp.reserved = int(prog[12])
        pos, gb = positions
        rc = self.lib.qsv_jit_compile_swap_ex(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), len(pos),
                                              _cabi.int_array(pos), _cabi.int_array(gb), int(pass_first), None, 0, None)
        if rc != 0:
            import warnings
            warnings.warn("qsv: fused swap+pass kernel unavailable (%s); running swap and pass separately"
                          % self.lib.qsv_last_error().decode("utf-8", "replace"))
        return rc == 0

    def _jit_runs(self, meta):
        ps, _, n_ops, extra = meta
        if n_ops >= 0:
            return False
        prog = extra[0]
        cp = self._cpass(ps)
        cp.reserved = int(prog[12])
        return bool(self.lib.qsv_jit_wanted(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size)))

    def _run_swap_pass(self, seg, support=None, fresh=None):
        self._presum = None
        swap_op, packed, dev, pass_first = seg
        ps, table_off, n_ops, extra = packed.passes[0]
        prog = extra[0]
        cp = self._cpass(ps)
        cp.max_dense_k = 1
        cp.reserved = int(prog[12])
        pos, gb = swap_fields(swap_op, self.n_local)
        self.comm.swap_pass(self, pos, cp, prog, C.c_void_p(dev.data_ptr() + table_off), pass_first, gbits=gb,
                            support=support, fresh=fresh)
        self.passes_run += 1
        self._keep.append(dev)

    def run_ops(self, ops):
        """Apply ops (physical positions) in order; consecutive tiled ops share passes."""
        plan = self.prepare(ops)
        self.execute(plan)
        return plan

    # ------------------------------------------------------------------ instrumentation (bench.py)
    def _launch_pass(self, packed_meta, dev, support=None, fresh=None):
        self._presum = None
        ps, table_off, n_ops, maxk = packed_meta
        base = dev.data_ptr()
        cp = self._cpass(ps)
        if fresh is not None and n_ops >= 0:
            raise RuntimeError("qsv: zero-fill pass requested for a program the pass-specialised kernel does not run")
        if n_ops < 0:       # register-resident round program (host image, passed to the kernel by value)
            prog = maxk[0]
            cp.max_dense_k = 1
            cp.reserved = int(prog[12])          # register bits per round (program header word 3)
            if fresh is not None:
                sp = support if support is not None else (0, 0)
                _cabi.check(self.lib.qsv_run_pass_rounds_fresh(self._ptr(self.state), C.byref(cp),
                                                               C.c_void_p(prog.ctypes.data), int(prog.size),
                                                               C.c_void_p(base + table_off), int(sp[0]), int(sp[1]),
                                                               int(fresh[0]), int(fresh[1]), self._stream()),
                            "qsv_run_pass_rounds_fresh")
                return
            if support is not None and support[0]:
                _cabi.check(self.lib.qsv_run_pass_rounds_sparse(self._ptr(self.state), C.byref(cp),
                                                                C.c_void_p(prog.ctypes.data), int(prog.size),
                                                                C.c_void_p(base + table_off), int(support[0]),
                                                                int(support[1]), self._stream()),
                            "qsv_run_pass_rounds_sparse")
                return
            _cabi.check(self.lib.qsv_run_pass_rounds(self._ptr(self.state), C.byref(cp),
                                                     C.c_void_p(prog.ctypes.data), int(prog.size),
                                                     C.c_void_p(base + table_off), self._stream()),
                        "qsv_run_pass_rounds")
            return
        cp.max_dense_k = maxk
        # tiles that cannot hold work are not even enumerated: the support hint plus the external controls every op
        # of the pass shares (a lone CZ / Toffoli / controlled phase touches 1/4 of the tiles)
        mask, val = common_external_controls(ps)
        if support is not None and support[0]:
            if (support[1] ^ val) & support[0] & mask:
                return                      # the shared control contradicts the support: nothing to do
            mask, val = mask | support[0], val | support[1]
        _cabi.check(self.lib.qsv_run_pass_sparse(self._ptr(self.state), C.byref(cp), C.c_void_p(base),
                                                 C.c_void_p(base + table_off), n_ops, int(mask), int(val),
                                                 self._stream()), "qsv_run_pass_sparse")

    def time_passes(self, plan, repeats=1):
        """CUDA-event duration (ms) of every k_pass launch of the plan, on the launching stream."""
        stream = torch.cuda.current_stream(self.device)
        self._basis = None
        out = []
        for kind, seg in plan["segments"]:
            if kind != "program":
                continue
            packed, dev = seg
            for meta in packed.passes:
                best = []
                for _ in range(repeats):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    self._launch_pass(meta, dev)
                    e1.record(stream)
                    e1.synchronize()
                    best.append(e0.elapsed_time(e1))
                out.append(sum(best) / len(best))
        return out

    def time_plan(self, plan, repeats=1, basis=None):
        """CUDA-event duration of every launch group of the plan with its algorithmic bytes:
        [{'kind': 'pass' | op kind, 'ms': .., 'bytes': .., 'active': fraction of the shard's tiles the launch
        touches}].  With `basis` (the index the register is initialised to in a step) every pass is launched with
        the support hint it gets in execute(), so the list is what a step runs; without it every pass is a full
        sweep.  'global_swap' and fused swap+pass segments are skipped (they are timed by the communicator)."""
        from .schedule import algorithmic_bytes
        stream = torch.cuda.current_stream(self.device)
        self._basis = None
        out = []
        fr = iter(self.pass_active_fractions(plan))
        local_mask = (1 << self.n_local) - 1
        n_fresh = self._fresh_prefix(plan) if basis is not None else 0
        for i, (kind, seg, sp, zf) in enumerate(self._walk(plan, basis)):
            if kind == "pass":
                meta, dev = seg
                ms = []
                for _ in range(repeats):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    self._launch_pass(meta, dev, sp, zf if i < n_fresh else None)
                    e1.record(stream)
                    e1.synchronize()
                    ms.append(e0.elapsed_time(e1))
                active = next(fr)
                if sp is not None and sp[0]:
                    here = (self.rank_bits & sp[0] & ~local_mask) == (sp[1] & ~local_mask)
                    active *= Support.fraction(sp[0] & local_mask) if here else 0.0
                out.append({"kind": "pass", "ms": sum(ms) / len(ms), "bytes": 32.0 * (1 << self.n_local) * active,
                            "ops": len(meta[0].ops), "active": active})
            elif kind == "op" and seg.kind != "global_swap":
                ms = []
                for _ in range(repeats):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    self._run_whole_state_op(seg, (0, 0))
                    e1.record(stream)
                    e1.synchronize()
                    ms.append(e0.elapsed_time(e1))
                out.append({"kind": seg.kind, "ms": sum(ms) / len(ms), "bytes": float(algorithmic_bytes(seg, self.n_local)),
                            "ops": 1, "active": 1.0})
        return out

    def pass_active_fractions(self, plan):
        """Fraction of tiles each pass actually loads (tiles switched off by external controls are
        skipped): 1.0 if any op is unconditional, else the union bound over the ops' control masks."""
        out = []
        for kind, seg in plan["segments"]:
            if kind != "program":
                continue
            packed, _ = seg
            for ps, _, _, _ in packed.passes:
                frac = 0.0
                for op in ps.ops:
                    ext = [c for c in op.controls if c >= ps.n_local or ps.loc(c) < 0]
                    frac += 2.0 ** (-len(ext))
                out.append(min(1.0, frac))
        return out

    # ------------------------------------------------------------------ measurement
    def probs(self, src_pos=None):
        """float64 tensor of the local probabilities in OUTPUT order (main/circuit.py:329-337)."""
        out = torch.empty(1 << self.n_local, dtype=torch.float64, device=self.device)
        sp = _cabi.int_array(src_pos) if src_pos is not None else None
        _cabi.check(self.lib.qsv_probs(self._ptr(self.state), self.n_local, sp, self._ptr(out), self._stream()),
                    "qsv_probs")
        return out

    def _scan(self, sp, z0, z1, prefix, target):
        res = torch.empty(2, dtype=torch.int64, device=self.device)
        _cabi.check(self.lib.qsv_sample_scan(self._ptr(self.state), self.n_local, sp, z0, z1, float(prefix),
                                             float(target), self._ptr(res), self._stream()), "qsv_sample_scan")
        host = res.cpu().numpy()
        self.d2h_bytes += 16
        return int(host[0]), float(host.view(np.float64)[1])

    def sample(self, u, src_pos=None):
        """Outcome index of random.choices(states, weights) for the uniform draw u
        (main/circuit.py:360-364): bisect_right(accumulate(weights), u*total, 0, N-1) over the weights in OUTPUT order.

        Up to 2^20 outcomes on one GPU: the sequential float scan, one launch, bit-exact with CPython.  Larger and
        sharded registers: the same rule with exact fixed-point sums (qsv.sampling / csrc/qsv_sample.cu) - the
        outcome is independent of the shard count and differs from the float rule only when u*total lies within one
        rounding error of a cumulative sum.  One small D2H per draw (16 / 64 bytes)."""
        sharded = self.comm is not None and self.comm.world > 1
        if not sharded and self.n_local <= SCAN_BITS_EXACT:
            sp = _cabi.int_array(src_pos) if src_pos is not None else None
            res = torch.empty(2, dtype=torch.int64, device=self.device)
            _cabi.check(self.lib.qsv_sample_sequential(self._ptr(self.state), self.n_local, sp, float(u), self._ptr(res),
                                                       self._stream()), "qsv_sample_sequential")
            host = res.cpu().numpy()
            self.d2h_bytes += 16
            if not float(host.view(np.float64)[1]) > 0.0:
                raise ValueError("Total of weights must be greater than zero")
            return int(host[0])
        if self._sampler is None:
            from .sampling import ExactSampler
            self._sampler = ExactSampler(self.lib, self.device)
        presummed = 0
        if self._presum is not None:
            src = tuple(src_pos) if src_pos is not None else tuple(range(self.n))
            if self._presum[1] == src:
                presummed = self._presum[0]      # the last pass left the first level's bucket sums in the sampler
            self._presum = None                  # (consumed: the descent overwrites the bucket array)
        z, nbytes = self._sampler.draw(self.state, self.n, self.n_local, self.rank, src_pos, u,
                                       torch.cuda.current_stream(self.device).cuda_stream,
                                       self.comm.all_reduce_int64 if sharded else None, presummed=presummed)
        self.d2h_bytes += nbytes
        return z

    def block_sums(self, src_pos=None, block_bits=SAMPLE_BLOCK_BITS):
        block_bits = min(block_bits, self.n_local)
        out = torch.empty(1 << (self.n_local - block_bits), dtype=torch.float64, device=self.device)
        sp = _cabi.int_array(src_pos) if src_pos is not None else None
        _cabi.check(self.lib.qsv_block_sums(self._ptr(self.state), self.n_local, sp, block_bits, self._ptr(out),
                                            self._stream()), "qsv_block_sums")
        return out

    def amplitudes(self):
        """Local amplitudes as a numpy array (tests / small states only)."""
        return self.state.cpu().numpy()
