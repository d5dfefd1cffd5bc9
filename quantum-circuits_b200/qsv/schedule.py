"""Host-side scheduling: gates -> ops on physical index bits -> tile passes -> packed programs.

Pure host logic (numpy only, no GPU needed): everything here is O(gates) bookkeeping that the
reference does inside run_method2 with 2^n x 2^n matrices (main/circuit.py:303-327).

Vocabulary
  wire      logical qubit of the circuit, wire 0 = most-significant index bit (main/circuit.py:290-299)
  position  physical bit of the amplitude index, position 0 = least significant
  op        one primitive the kernels understand (dense / mcx / phase / diag / swap, or a whole-state op)
  pass      one sweep over the state: a tile (set of positions held in shared memory) + a program of ops
"""
from dataclasses import dataclass, field
import os
import numpy as np

from . import _cabi

TILED_KINDS = ("dense", "mcx", "phase", "diag", "swap")


@dataclass
class Op:
    kind: str                       # TILED_KINDS or 'wide_dense' | 'wide_perm' | 'wide_diag' | 'modexp' | 'oracle' | 'diffusion'
    targets: tuple = ()             # physical positions; targets[0] = gate MSB (= gate port 0)
    controls: tuple = ()            # physical positions the op is conditioned on
    cvals: object = None            # required value of each control (None = all 1)
    matrix: object = None           # dense: (2^k, 2^k) complex128; diag / wide_diag: (2^k,) complex128
    scalar: complex = 1.0           # phase
    perm: object = None             # wide_perm: uint32[2^k], perm[c] = row r with M[r][c] == 1
    params: dict = field(default_factory=dict)
    name: str = ""

    def nondiag_bits(self):
        if self.kind in ("dense", "mcx", "swap"):
            return tuple(self.targets)
        if self.kind in ("phase", "diag"):
            return ()
        return tuple(self.targets) + tuple(self.controls)

    def diag_bits(self):
        if self.kind in ("dense", "mcx", "swap"):
            return tuple(self.controls)
        if self.kind in ("phase", "diag"):
            return tuple(self.targets) + tuple(self.controls)
        return ()


# ------------------------------------------------------------------------------------------------
# gate classification
# ------------------------------------------------------------------------------------------------

def _as_array(matrix_rows):
    return np.array(matrix_rows, dtype=np.complex128)


def _is_diagonal(M):
    return not np.any(M - np.diag(np.diag(M)))


def _is_permutation(M):
    if np.any(M.imag) or np.any((M.real != 0) & (M.real != 1)):
        return False
    R = M.real
    return bool(np.all(R.sum(axis=0) == 1) and np.all(R.sum(axis=1) == 1))


def _perm_of(M):
    """perm[c] = r with M[r][c] == 1  (psi'[r] = sum_c M[r][c] psi[c])."""
    return np.argmax(M.real, axis=0).astype(np.uint32)


def _perm_as_mcx(perm, k):
    """Recognise X on one gate bit conditioned on fixed values of a subset of the other bits.
    Returns (target_bit, control_bits, control_values) in gate-index bit numbers (0 = LSB)."""
    D = 1 << k
    idx = np.arange(D, dtype=np.uint32)
    moved = idx[perm != idx]
    if moved.size == 0:
        return None
    x = int(perm[moved[0]] ^ moved[0])
    if x & (x - 1):
        return None
    t = x.bit_length() - 1
    ones = int(np.bitwise_and.reduce(moved))
    zeros = ~int(np.bitwise_or.reduce(moved)) & (D - 1)
    cmask = (ones | zeros) & ~x & (D - 1)
    cval = ones & cmask
    expect = np.where((idx & np.uint32(cmask)) == np.uint32(cval), idx ^ np.uint32(x), idx)
    if np.array_equal(expect, perm):
        cbits = [b for b in range(k) if (cmask >> b) & 1]
        return t, cbits, [(cval >> b) & 1 for b in cbits]
    return None


def _perm_as_xor(perm, k):
    """perm[c] == c ^ mask (tensor power of X): returns mask or None."""
    mask = int(perm[0])
    idx = np.arange(1 << k, dtype=np.uint32)
    return mask if np.array_equal(perm, idx ^ np.uint32(mask)) else None


def _kron_split(M, a=1):
    """Try M = A (2^a x 2^a) (x) B.  Returns (A, B) or None.  Exact up to ~1 ulp per entry."""
    D = M.shape[0]
    da = 1 << a
    h = D // da
    R = M.reshape(da, h, da, h).transpose(0, 2, 1, 3).reshape(da * da, h * h)
    i, j = np.unravel_index(np.argmax(np.abs(R)), R.shape)
    piv = R[i, j]
    if piv == 0:
        return None
    a = R[:, j] / piv
    b = R[i, :]
    tol = 4e-16 * max(1.0, float(np.abs(M).max()))
    if np.abs(np.outer(a, b) - R).max() > tol:
        return None
    A, B = a.reshape(da, da), b.reshape(h, h)
    # keep scalar factors out of identity blocks: (cA) (x) I rather than A (x) (cI)
    if _is_diagonal(B) and np.all(np.diag(B) == B[0, 0]) and B[0, 0] != 0:
        A, B = A * B[0, 0], np.eye(h, dtype=np.complex128)
    elif _is_diagonal(A) and np.all(np.diag(A) == A[0, 0]) and A[0, 0] != 0:
        A, B = np.eye(da, dtype=np.complex128), B * A[0, 0]
    return A, B


def _is_identity(M):
    return np.array_equal(M, np.eye(M.shape[0], dtype=M.dtype))


def classify_matrix(M, pos, name=""):
    """Turn one k-qubit gate matrix acting on physical positions pos (pos[0] = gate MSB) into ops.

    Follows the gate set of SURVEY §2a: diagonal -> phase/diag, permutation -> (multi-)controlled X or
    exact index permutation, tensor products -> per-factor ops, anything else dense.
    """
    k = len(pos)
    D = 1 << k
    M = np.asarray(M, dtype=np.complex128)
    assert M.shape == (D, D), "matrix does not match the number of target qubits"
    pos = tuple(int(p) for p in pos)

    def bitpos(b):          # gate-index bit b (0 = LSB) -> physical position
        return pos[k - 1 - b]

    if _is_diagonal(M):
        d = np.diag(M).copy()
        if np.all(d == 1):
            return []
        # drop qubits the diagonal does not depend on (U0 (x) I of main/Grover.py:31 ignores the ancilla)
        for b in range(k):
            v = d.reshape(1 << (k - 1 - b), 2, 1 << b)
            if np.array_equal(v[:, 0, :], v[:, 1, :]):
                keep = tuple(p for p in range(k) if p != k - 1 - b)
                return classify_matrix(np.diag(v[:, 0, :].reshape(-1)), tuple(pos[p] for p in keep), name)
        off = np.flatnonzero(d != 1)
        if off.size == 1:                             # Z, PhaseShift, CZ, controlled phase
            j = int(off[0])
            return [Op("phase", (), pos, cvals=tuple((j >> (k - 1 - p)) & 1 for p in range(k)),
                       scalar=complex(d[j]), name=name)]
        if k <= _cabi.QSV_MAX_DIAG_K:
            return [Op("diag", pos, (), matrix=d, name=name)]
        return [Op("wide_diag", pos, (), matrix=d, name=name)]

    if _is_permutation(M):
        perm = _perm_of(M)
        mask = _perm_as_xor(perm, k)
        if mask is not None:                          # X^(x)size
            return [Op("mcx", (bitpos(b),), (), name=name) for b in range(k) if (mask >> b) & 1]
        mcx = _perm_as_mcx(perm, k)
        if mcx is not None:                           # CNot, Toffoli, Grover oracle
            t, cbits, cv = mcx
            return [Op("mcx", (bitpos(t),), tuple(bitpos(b) for b in cbits), cvals=tuple(cv), name=name)]
        if k > _cabi.QSV_MAX_TILE_K:
            return [Op("wide_perm", pos, (), perm=perm, name=name)]
        # small generic permutation: a 0/1 dense matrix moves amplitudes exactly (1*x + 0*y == x)

    if k == 1:
        return [Op("dense", pos, (), matrix=M, name=name)]

    # tensor products (H(size), H^n (x) I and U0 (x) I of main/Grover.py:31,40): factor by factor
    for a in range(1, k):
        split = _kron_split(M, a)
        if split is not None:
            A, B = split
            ops = []
            if not _is_identity(A):
                ops += classify_matrix(A, pos[:a], name)
            if not _is_identity(B):
                ops += classify_matrix(B, pos[a:], name)
            return ops

    if k <= _cabi.QSV_MAX_TILE_K:
        return [Op("dense", pos, (), matrix=M, name=name)]
    return [Op("wide_dense", pos, (), matrix=M, name=name)]


# ------------------------------------------------------------------------------------------------
# pass planning
# ------------------------------------------------------------------------------------------------

def default_tile_bits():
    return int(os.environ.get("QSV_TILE_BITS", "12"))


def default_low_bits():
    return int(os.environ.get("QSV_LOW_BITS", "6"))


@dataclass
class Pass:
    n_local: int
    tile_bits: int
    low_bits: int
    hi_pos: tuple
    ops: list

    def loc(self, p):
        """tile-local position of physical position p, or -1 if outside the tile."""
        if p < self.low_bits:
            return p
        if p in self.hi_pos:
            return self.low_bits + self.hi_pos.index(p)
        return -1


def _mask(bits):
    m = 0
    for b in bits:
        m |= 1 << b
    return m


def plan_passes(ops, n_local, tile_bits=None, low_bits=None):
    """Greedy partition of a list of TILED ops into passes (each one read+write of the state).

    A pass's tile always holds positions 0..L-1 plus up to T-L higher positions.  Ops are taken in
    program order; an op joins the current pass if all its non-diagonal targets fit in the tile and
    it commutes past every op that had to be deferred (deferred ops block the bits they touch:
    non-diagonal use blocks everything, diagonal/control use blocks only non-diagonal use).
    """
    T = min(n_local, tile_bits if tile_bits is not None else default_tile_bits())
    T = min(T, _cabi.QSV_MAX_TILE_T)
    L = min(T, low_bits if low_bits is not None else default_low_bits())
    L = max(L, T - 8)
    remaining = list(ops)
    passes = []
    while remaining:
        hi = []
        free = T - L
        blocked_nd = 0
        blocked_d = 0
        chosen, rest = [], []
        for op in remaining:
            nd = list(op.nondiag_bits())
            if any(b >= n_local for b in nd):
                raise RuntimeError("op %r targets a global qubit; it must be swapped local first" % (op.name,))
            dg = list(op.diag_bits())
            nd_m, dg_m = _mask(nd), _mask(dg)
            ok = not (nd_m & (blocked_nd | blocked_d)) and not (dg_m & blocked_nd)
            if ok:
                new = [b for b in set(nd) if b >= L and b not in hi]
                if len(new) <= free:
                    hi.extend(new)
                    free -= len(new)
                    chosen.append(op)
                    continue
            blocked_nd |= nd_m
            blocked_d |= dg_m
            rest.append(op)
        if not chosen:
            raise RuntimeError("pass planner made no progress (op needs more than %d tile-local qubits)" % T)
        # fill the unused tile slots with the lowest free positions, avoiding control/diagonal bits
        # of the chosen ops so that those stay external (tiles they switch off are never loaded)
        avoid = 0
        for op in chosen:
            avoid |= _mask(op.diag_bits())
        for rnd in (0, 1):
            for b in range(L, n_local):
                if free == 0:
                    break
                if b in hi or (rnd == 0 and (avoid >> b) & 1):
                    continue
                hi.append(b)
                free -= 1
        passes.append(Pass(n_local, T, L, tuple(sorted(hi)), chosen))
        remaining = rest
    return passes


# ------------------------------------------------------------------------------------------------
# program packing (must mirror qsv_op_t in include/qsv.h)
# ------------------------------------------------------------------------------------------------

OP_DTYPE = np.dtype([
    ("kind", "<u4"), ("k", "<u4"), ("n_ins", "<u4"), ("flags", "<u4"),
    ("ext_ctrl_mask", "<u8"), ("ext_ctrl_val", "<u8"),
    ("loc_ctrl_val", "<u4"), ("payload_off", "<u4"), ("op_bytes", "<u4"), ("reserved", "<u4"),
    ("ins", "u1", (16,)), ("tgt", "u1", (8,)), ("scalar", "<f8", (2,)),
])
assert OP_DTYPE.itemsize == 88


def encode_op(op, ps):
    """Serialise one tiled op for pass `ps` (positions translated to tile-local bits)."""
    hdr = np.zeros((), dtype=OP_DTYPE)
    payload = b""
    ins = []
    ext_mask = 0
    loc_ctrl = 0
    ext_val = 0
    ctrls = tuple(op.controls)
    cvals = tuple(op.cvals) if op.cvals is not None else (1,) * len(ctrls)
    assert len(cvals) == len(ctrls)
    for c, v in zip(ctrls, cvals):
        l = ps.loc(c) if c < ps.n_local else -1
        if l >= 0:
            ins.append(l)
            loc_ctrl |= (1 << l) if v else 0
        else:
            ext_mask |= 1 << c
            ext_val |= (1 << c) if v else 0
    if op.kind == "dense":
        k = len(op.targets)
        if k > _cabi.QSV_MAX_TILE_K:
            raise RuntimeError("dense op wider than %d qubits cannot be tiled" % _cabi.QSV_MAX_TILE_K)
        M = np.ascontiguousarray(op.matrix, dtype=np.complex128)
        hdr["kind"], hdr["k"] = _cabi.QSV_OP_DENSE, k
        if not np.any(M.imag):
            hdr["flags"] = _cabi.QSV_OPF_REAL_MATRIX
        payload = M.tobytes()
        for p, t in enumerate(op.targets):
            l = ps.loc(t)
            assert l >= 0, "dense target outside the tile"
            hdr["tgt"][p] = l
            ins.append(l)
    elif op.kind == "mcx":
        hdr["kind"], hdr["k"] = _cabi.QSV_OP_MCX, 1
        l = ps.loc(op.targets[0])
        assert l >= 0, "mcx target outside the tile"
        hdr["tgt"][0] = l
        ins.append(l)
    elif op.kind == "swap":
        hdr["kind"], hdr["k"] = _cabi.QSV_OP_SWAP, 2
        for p, t in enumerate(op.targets):
            l = ps.loc(t)
            assert l >= 0, "swap target outside the tile"
            hdr["tgt"][p] = l
            ins.append(l)
    elif op.kind == "phase":
        hdr["kind"] = _cabi.QSV_OP_PHASE
        hdr["scalar"][0], hdr["scalar"][1] = op.scalar.real, op.scalar.imag
    elif op.kind == "diag":
        k = len(op.targets)
        if k > _cabi.QSV_MAX_DIAG_K:
            raise RuntimeError("diag op wider than %d qubits cannot be tiled" % _cabi.QSV_MAX_DIAG_K)
        hdr["kind"], hdr["k"] = _cabi.QSV_OP_DIAG, k
        payload = np.ascontiguousarray(op.matrix, dtype=np.complex128).tobytes()
        for p, t in enumerate(op.targets):
            l = ps.loc(t) if t < ps.n_local else -1
            hdr["tgt"][p] = l if l >= 0 else (0x80 | t)
    else:
        raise RuntimeError("op kind %r is not a tiled op" % op.kind)
    ins.sort()
    if len(set(ins)) != len(ins):
        raise RuntimeError("a qubit is used twice by one gate")
    if len(ins) > 16:
        raise RuntimeError("too many tile-local bits in one op")
    hdr["n_ins"] = len(ins)
    for i, l in enumerate(ins):
        hdr["ins"][i] = l
    hdr["ext_ctrl_mask"] = ext_mask
    hdr["ext_ctrl_val"] = ext_val
    hdr["loc_ctrl_val"] = loc_ctrl
    hdr["payload_off"] = _cabi.QSV_OP_HEADER_BYTES
    total = (_cabi.QSV_OP_HEADER_BYTES + len(payload) + 15) // 16 * 16
    hdr["op_bytes"] = total
    raw = hdr.tobytes() + b"\0" * (_cabi.QSV_OP_HEADER_BYTES - OP_DTYPE.itemsize) + payload
    return raw + b"\0" * (total - len(raw))


@dataclass
class PackedProgram:
    blob: np.ndarray        # uint8, the whole device image
    passes: list            # [(Pass, offsets_byte_offset, n_ops, max_dense_k)]


def pack_passes(passes):
    """Lay out [offset tables of every pass | ops of every pass] in one buffer; op offsets are
    relative to the start of the buffer."""
    tables = 0
    for ps in passes:
        tables += (4 * len(ps.ops) + 15) // 16 * 16
    chunks, meta, table_words = [], [], []
    cursor = tables
    tcur = 0
    for ps in passes:
        offs = []
        maxk = 0
        for op in ps.ops:
            raw = encode_op(op, ps)
            offs.append(cursor)
            chunks.append(raw)
            cursor += len(raw)
            if op.kind == "dense":
                maxk = max(maxk, len(op.targets))
        tbytes = (4 * len(offs) + 15) // 16 * 16
        tw = np.zeros(tbytes // 4, dtype="<u4")
        tw[: len(offs)] = offs
        table_words.append(tw)
        meta.append((ps, tcur, len(offs), maxk))
        tcur += tbytes
    blob = np.frombuffer(b"".join([t.tobytes() for t in table_words] + chunks), dtype=np.uint8).copy()
    assert blob.size == cursor and cursor < 2 ** 32
    return PackedProgram(blob, meta)


# ------------------------------------------------------------------------------------------------
# algorithmic bytes — numerator of the roofline
# ------------------------------------------------------------------------------------------------

def algorithmic_bytes(op, n_local):
    N = 1 << n_local
    if op.kind in ("dense", "wide_dense", "wide_perm", "swap"):
        return 32 * N >> len([c for c in op.controls if c < n_local])
    if op.kind == "mcx":
        return 32 * N >> len([c for c in op.controls if c < n_local])
    if op.kind == "phase":
        return 32 * N >> len([c for c in op.controls if c < n_local])
    if op.kind in ("diag", "wide_diag"):
        return 32 * N
    if op.kind == "diffusion":
        return 48 * N
    return 0
