"""Host side of the exact sampler (csrc/qsv_sample.cu): the level plan of the radix descent and its driver.

random.choices (main/circuit.py:360-364) returns bisect_right(accumulate(weights), u * total, 0, N - 1) over the
weights in OUTPUT order.  Registers above 2^20 outcomes and every sharded register evaluate that rule with exact
fixed-point sums, descending over the output index from its most significant bit: a level pins up to 12 output bits
(bucket sums over the elements still compatible with the bits chosen so far -> all-reduce over the ranks -> select),
the last <= 12 bits are resolved by a leaf.  The plan depends only on the output wiring (src_pos) and the shard
geometry, never on the data; the chosen bits stay in device memory between the kernels.
"""
import ctypes as C
import math
import os

import torch

from . import _cabi

LEVEL_BITS = 12        # QSV_SAMPLE_MAX_LEVEL_BITS
RUN_BITS = 10          # the level kernel walks runs over the 10 lowest free local positions ...
MAX_LOW = 2            # ... and keeps per-thread accumulators for at most two bucket bits inside a run


PRESUM_MAX_INSIDE = 3   # kSumMaxInside of csrc/qsv_jit.cu


def presum_bits(src_pos, n, tile_positions, max_inside=PRESUM_MAX_INSIDE):
    """How many of the TOP output bits the last pass of a run can weigh for the sampler (qsv_run_pass_rounds_sums):
    the longest prefix, at most 12 bits, of which at most three sit on positions INSIDE that pass's tile (a bucket
    bit outside the tile is the same for a whole tile; every bit inside doubles the partial sums a tile keeps).
    max_inside / QSV_PRESUM_MAX_INSIDE lower the three."""
    src = list(src_pos) if src_pos is not None else list(range(n))
    tile = set(tile_positions)
    k = inside = 0
    max_inside = min(PRESUM_MAX_INSIDE, max_inside, int(os.environ.get("QSV_PRESUM_MAX_INSIDE", PRESUM_MAX_INSIDE)))
    while k < min(LEVEL_BITS, n):
        if src[n - 1 - k] in tile:
            if inside == max_inside:
                break
            inside += 1
        k += 1
    return k


def level_plan(src_pos, n, n_local, first_bits=0):
    """[("level" | "leaf", out_bits, positions, fixed_mask)] for an n-qubit register whose output bit b shows
    physical position src_pos[b] (positions >= n_local are global = rank bits).  out_bits / positions are listed
    from the LEAST significant bucket bit; fixed_mask = physical positions pinned before the step runs.
    first_bits > 0: the first level is exactly the top first_bits output bits (its bucket sums were produced by the
    last gate pass, presum_bits); the descent continues below them as usual."""
    src = list(src_pos) if src_pos is not None else list(range(n))
    if sorted(src) != list(range(n)):
        raise RuntimeError("qsv.sampling: src_pos is not a permutation of the %d positions" % n)
    steps = []
    fixed = 0
    b = n - 1
    if first_bits:
        bits = list(range(n - first_bits, n))
        steps.append(("level", bits, [src[x] for x in bits], 0))
        for x in bits:
            fixed |= 1 << src[x]
        b = n - 1 - first_bits
    while b + 1 > LEVEL_BITS:
        free_local = [p for p in range(n_local) if not (fixed >> p) & 1]
        run = set(free_local[:RUN_BITS])
        bits, n_low = [], 0
        while b >= 0 and len(bits) < LEVEL_BITS:
            p = src[b]
            if p in run:
                if n_low == MAX_LOW:
                    break
                n_low += 1
            bits.append(b)
            b -= 1
        bits.reverse()                                   # least significant bucket bit first
        steps.append(("level", bits, [src[x] for x in bits], fixed))
        for x in bits:
            fixed |= 1 << src[x]
    bits = list(range(0, b + 1))
    steps.append(("leaf", bits, [src[x] for x in bits], fixed))
    return steps


def u_fixed(u):
    """u in [0, 1) as (mantissa < 2^53, shift): u = mantissa * 2^-shift exactly."""
    if not 0.0 <= u < 1.0:
        raise ValueError("uniform draw outside [0, 1): %r" % (u,))
    if u == 0.0:
        return 0, 0
    m, e = math.frexp(u)               # u = m * 2^e, 0.5 <= m < 1
    return int(m * (1 << 53)), 53 - e


class ExactSampler:
    """Device buffers + the launch sequence of one draw.  `all_reduce(int64 tensor)` sums the bucket arrays over
    the ranks of a sharded register (None on one GPU)."""

    def __init__(self, lib, device):
        self.lib = lib
        self.device = device
        self.sel = torch.zeros(8, dtype=torch.int64, device=device)
        self.buckets = torch.empty(4 << LEVEL_BITS, dtype=torch.int64, device=device)

    def draw(self, state, n, n_local, rank, src_pos, u, stream, all_reduce=None, presummed=0):
        """Outcome index (output order) of the draw u; raises ValueError if the total weight is zero.
        Returns (z, d2h_bytes).  presummed = k > 0: self.buckets already holds the bucket sums of the top k output
        bits (left there by the last gate pass, qsv_run_pass_rounds_sums): the first level's sweep is skipped."""
        lib = self.lib
        rank_bits = int(rank) << n_local
        mant, shift = u_fixed(u)
        st = C.c_void_p(stream)
        sel, bk, sp = C.c_void_p(self.sel.data_ptr()), C.c_void_p(self.buckets.data_ptr()), C.c_void_p(state.data_ptr())
        _cabi.check(lib.qsv_sample_begin(sel, st), "qsv_sample_begin")
        first = 1
        for kind, out_bits, positions, fixed in level_plan(src_pos, n, n_local, first_bits=presummed):
            nb = len(out_bits)
            pos, ob = _cabi.int_array(positions), _cabi.int_array(out_bits)
            if not (first and presummed):
                fn = lib.qsv_sample_level if kind == "level" else lib.qsv_sample_leaf
                _cabi.check(fn(sp, n_local, rank_bits, int(fixed), nb, pos, sel, bk, st), "qsv_sample_" + kind)
            if all_reduce is not None:
                all_reduce(self.buckets[: 4 << nb])
            _cabi.check(lib.qsv_sample_select(bk, nb, pos, ob, first, mant, shift, sel, st), "qsv_sample_select")
            first = 0
        host = self.sel.cpu().tolist()                   # the one D2H of a draw: 64 bytes
        if host[6]:
            raise ValueError("Total of weights must be greater than zero")
        return int(host[1]) & ((1 << n) - 1), 64
