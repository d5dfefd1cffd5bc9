"""Shared test helpers: rebuild fixture circuits with the mirror API, emulate packed tile programs."""
import numpy as np


def to_complex_rows(m):
    out = []
    for row in m:
        r = []
        for re, im in row:
            r.append(complex(re, im) if im != 0 else (int(re) if float(re).is_integer() else re))
        out.append(r)
    return out


def perm_to_rows(perm):
    D = len(perm)
    rows = [[0] * D for _ in range(D)]
    for col, r in enumerate(perm):
        rows[int(r)][col] = 1
    return rows


def build_circuit(case, arrays=None):
    """Re-create a serialised fixture circuit with quantum-circuits_b200/main (the mirror API)."""
    import main.circuit as mc
    from main.matrix import Matrix
    c = mc.Circuit(case["n"])
    gates = []
    for g in case["gates"]:
        cls = g["cls"]
        if cls == "Gate":
            if "perm_matrix" in g:
                rows = perm_to_rows(arrays["modexp_perm/" + g["perm_matrix"]])
            elif "matrix_key" in g:
                rows = [[complex(v) for v in row] for row in arrays[g["matrix_key"]]]
            else:
                rows = to_complex_rows(g["matrix"])
            obj = mc.Gate(Matrix(rows), g["name"], True)
        elif cls in ("CNot", "T"):
            obj = getattr(mc, cls)(g["name"])
        else:
            obj = getattr(mc, cls)(g["size"], g["name"])
        gates.append(c.add(obj))
    for src, sp, dst, dp in case["wires"]:
        c.add_wire(c if src < 0 else gates[src], sp, c if dst < 0 else gates[dst], dp)
    return c


def circuit_from_steps(steps, n):
    """Mirror-API circuit from oracle-style (matrix, wires[, label]) steps."""
    import main.circuit as mc
    from main.matrix import Matrix
    c = mc.Circuit(n)
    last = [(c, i) for i in range(n)]
    for k, st in enumerate(steps):
        M, wires = st[0], st[1]
        label = st[2] if len(st) > 2 else "G"
        rows = [[complex(v) for v in row] for row in np.asarray(M)]
        g = c.add(mc.Gate(Matrix(rows), "%s_%d" % (label, k), True))
        for p, w in enumerate(wires):
            c.add_wire(last[w][0], last[w][1], g, p)
            last[w] = (g, p)
    for w in range(n):
        c.add_wire(last[w][0], last[w][1], c, w)
    return c


# ------------------------------------------------------------------------------------------------
# numpy emulation of k_pass (csrc/qsv_tile.cu) working directly on the packed device image: checks
# plan_passes + encode_op + the tile index arithmetic on the CPU, where no GPU is available.
# ------------------------------------------------------------------------------------------------

def _insert_zero(v, p):
    return ((v >> p) << (p + 1)) | (v & ((1 << p) - 1))


def emulate_program(packed, psi, rank_bits=0):
    from qsv.schedule import OP_DTYPE
    from qsv import _cabi
    blob = packed.blob.tobytes()
    psi = psi.copy()
    for ps, table_off, n_ops, maxk in packed.passes:
        if n_ops < 0:
            psi = emulate_round_program(maxk[0], blob[table_off:], ps, psi, rank_bits)
            continue
        T, L = ps.tile_bits, ps.low_bits
        hi = list(ps.hi_pos)
        offs = np.frombuffer(blob, dtype="<u4", count=n_ops, offset=table_off)
        loc = np.arange(1 << T)
        gl_off = loc & ((1 << L) - 1)
        for j, p in enumerate(hi):
            gl_off = gl_off | (((loc >> (L + j)) & 1) << p)
        for t in range(1 << (ps.n_local - T)):
            base = t << L
            for p in hi:
                base = _insert_zero(base, p)
            gbase = base | rank_bits
            tile = psi[base + gl_off].copy()
            for o in offs:
                hdr = np.frombuffer(blob, dtype=OP_DTYPE, count=1, offset=int(o))[0]
                if (gbase & int(hdr["ext_ctrl_mask"])) != int(hdr["ext_ctrl_val"]):
                    continue
                k, n_ins = int(hdr["k"]), int(hdr["n_ins"])
                ins = [int(x) for x in hdr["ins"][:n_ins]]
                tgt = [int(x) for x in hdr["tgt"]]
                j = np.arange(1 << (T - n_ins))
                for p in ins:
                    j = _insert_zero(j, p)
                bidx = j | int(hdr["loc_ctrl_val"])
                pay = int(o) + int(hdr["payload_off"])
                kind = int(hdr["kind"])
                if kind == _cabi.QSV_OP_DENSE:
                    D = 1 << k
                    M = np.frombuffer(blob, dtype=np.complex128, count=D * D, offset=pay).reshape(D, D)
                    off = [sum(((c >> (k - 1 - p)) & 1) << tgt[p] for p in range(k)) for c in range(D)]
                    x = np.stack([tile[bidx | off[c]] for c in range(D)])
                    y = M @ x
                    for r in range(D):
                        tile[bidx | off[r]] = y[r]
                elif kind == _cabi.QSV_OP_MCX:
                    b = 1 << tgt[0]
                    a0, a1 = tile[bidx].copy(), tile[bidx | b].copy()
                    tile[bidx], tile[bidx | b] = a1, a0
                elif kind == _cabi.QSV_OP_SWAP:
                    ba, bb = 1 << tgt[0], 1 << tgt[1]
                    a0, a1 = tile[bidx | ba].copy(), tile[bidx | bb].copy()
                    tile[bidx | ba], tile[bidx | bb] = a1, a0
                elif kind == _cabi.QSV_OP_PHASE:
                    tile[bidx] = tile[bidx] * complex(hdr["scalar"][0], hdr["scalar"][1])
                elif kind == _cabi.QSV_OP_DIAG:
                    d = np.frombuffer(blob, dtype=np.complex128, count=1 << k, offset=pay)
                    di = np.zeros_like(bidx)
                    for p in range(k):
                        tb = tgt[p]
                        if tb & 0x80:
                            di |= ((gbase >> (tb & 0x7F)) & 1) << (k - 1 - p)
                        else:
                            di |= ((bidx >> tb) & 1) << (k - 1 - p)
                    tile[bidx] = tile[bidx] * d[di]
                else:
                    raise AssertionError("bad op kind %d" % kind)
            psi[base + gl_off] = tile
    return psi


def emulate_round_program(prog, tables, ps, psi, rank_bits=0, single_tile_gbase=None):
    """numpy emulation of k_pass_reg (csrc/qsv_tile_reg.cu) on the host round program.  With
    single_tile_gbase, psi is ONE tile (2^T amplitudes) whose global base index is that value."""
    from qsv.schedule import ROP_DTYPE, ROUND_DTYPE
    prog = bytes(prog)
    n_rounds, total, n_ops, RB = (int(v) for v in np.frombuffer(prog, dtype="<u4", count=4))
    assert total == len(prog) and RB in (3, 4)
    NS = 1 << RB
    rounds = np.frombuffer(prog, dtype=ROUND_DTYPE, count=n_rounds, offset=16)
    rops = np.frombuffer(prog, dtype=ROP_DTYPE, count=n_ops, offset=16 + 16 * int(n_rounds))
    T, L = ps.tile_bits, ps.low_bits
    hi = list(ps.hi_pos)
    loc = np.arange(1 << T)
    gl_off = loc & ((1 << L) - 1)
    for j, p in enumerate(hi):
        gl_off = gl_off | (((loc >> (L + j)) & 1) << p)
    psi = psi.copy()
    tid = np.arange(1 << (T - RB))

    def permute(tb, rb, ops, gbase, general):
        """Addresses [16, threads] after the permutation ops, computed the way the kernel does: per slot
        (general) or in XOR-linear form base ^ XOR d_q (perm_linear in csrc/qsv_tile_reg.cu)."""
        roff = [sum((1 << rb[q]) for q in range(len(rb)) if (r >> q) & 1) for r in range(1 << len(rb))]
        if general:
            a = np.stack([tb | roff[r] for r in range(1 << len(rb))])
            for o in ops:
                assert int(o["kind"]) == 19 and bool(int(o["flags2"]) & 1) == bool(int(o["ext_mask"]))
                if (gbase & int(o["ext_mask"])) != int(o["ext_val"]):
                    continue
                cm, cv, xm = int(o["tmask"]), int(o["tval"]), int(o["table_off"])
                assert xm and not (xm & cm)
                a = np.where((a & cm) == cv, a ^ xm, a)
            return a
        base = tb.copy()
        d = [np.full_like(tb, 1 << rb[q]) for q in range(len(rb))]
        for o in ops:
            assert int(o["kind"]) == 19 and bool(int(o["flags2"]) & 1) == bool(int(o["ext_mask"]))
            if (gbase & int(o["ext_mask"])) != int(o["ext_val"]):
                continue
            cm, cv, xm, q = int(o["tmask"]), int(o["tval"]), int(o["table_off"]), int(o["j"])
            assert xm and not (xm & cm)
            if q == 0xFF:
                assert int(o["flags"]) == 0
                base = np.where((base & cm) == cv, base ^ xm, base)
                continue
            pbit = 1 << int(o["rval"])
            assert cm & pbit and int(o["flags"]) == pbit
            cmu = cm & ~pbit
            f = np.where((base & cmu) == (cv & cmu), xm, 0)
            fd = np.where((d[q] & pbit) != 0, f, 0)
            base = np.where(((base ^ cv) & pbit) == 0, base ^ f, base)
            d[q] = d[q] ^ fd
        a = [base]
        for r in range(1, 1 << len(rb)):
            low = (r & -r).bit_length() - 1
            a.append(a[r & (r - 1)] ^ d[low])
        return np.stack(a)

    def tiles():
        if single_tile_gbase is not None:
            yield None, single_tile_gbase
            return
        for t in range(1 << (ps.n_local - T)):
            base = t << L
            for p in hi:
                base = _insert_zero(base, p)
            yield base, base | rank_bits

    for base, gbase in tiles():
        if not any((gbase & int(o["ext_mask"])) == int(o["ext_val"]) for o in rops):
            continue
        tile = psi.copy() if base is None else psi[base + gl_off].copy()
        for R in rounds:
            rb = [int(b) for b in R["rb"]][:RB]
            assert rb == sorted(rb) and len(set(rb)) == RB
            tb = tid.copy()
            for b in rb:
                tb = _insert_zero(tb, b)
            f0 = int(R["first_op"])
            npre, gen_pre = int(R["n_pre"]) & 0x7FFF, bool(int(R["n_pre"]) & 0x8000)
            npost, gen_post = int(R["n_post"]) & 0x7FFF, bool(int(R["n_post"]) & 0x8000)
            n_pt, nreg0 = (int(R["n_ops"]) >> 16) & 0x3FFF, int(R["n_ops"]) & 0xFFFF
            assert all(int(o["kind"]) == 18 for o in rops[f0 + npre: f0 + npre + n_pt])
            assert all(int(o["kind"]) < 18 or int(o["kind"]) == 20 for o in rops[f0 + npre + n_pt: f0 + npre + n_pt + nreg0])
            nreg = n_pt + nreg0          # the phase_t list commutes with the ordered list: emulate in sequence
            x = tile[permute(tb, rb, rops[f0: f0 + npre][::-1], gbase, gen_pre)]           # [16, threads]
            for o in rops[f0 + npre: f0 + npre + nreg]:
                assert bool(int(o["flags2"]) & 1) == bool(int(o["ext_mask"]))
                if (gbase & int(o["ext_mask"])) != int(o["ext_val"]):
                    continue
                tp = (tb & int(o["tmask"])) == int(o["tval"])
                kind, J = int(o["kind"]), int(o["j"])
                rm, rv = int(o["rmask"]), int(o["rval"])
                rsel = [r for r in range(NS) if (r & rm) == rv]
                # decode the kind byte exactly like apply_rop() in csrc/qsv_tile_reg.cu and cross-check
                # the specialisation against the generic (j, rmask, rval) fields
                if 0 <= kind < 8:
                    assert J == (kind & 3) and rm == 0 and bool(int(o["flags"]) & 1) == (kind >= 4)
                    m = o["m"].view(np.complex128)
                    assert J < RB
                    for r in range(NS):
                        if (r >> J) & 1:
                            continue
                        x0, x1 = x[r].copy(), x[r | (1 << J)].copy()
                        x[r] = np.where(tp, m[0] * x0 + m[1] * x1, x0)
                        x[r | (1 << J)] = np.where(tp, m[2] * x0 + m[3] * x1, x1)
                elif kind == 18 or 8 <= kind < 16 or kind == 16:
                    if kind == 18:
                        assert rm == 0
                    elif kind != 16:
                        jj, vv = (kind - 8) >> 1, (kind - 8) & 1
                        assert rm == 1 << jj and rv == vv << jj
                    sc = complex(o["m"][0], o["m"][1])
                    for r in rsel:
                        x[r] = np.where(tp, x[r] * sc, x[r])
                elif kind == 20:           # in-register (multi-)controlled X on register bit J
                    assert J < RB and not (rm >> J) & 1
                    for r in rsel:
                        if (r >> J) & 1:
                            continue
                        x0, x1 = x[r].copy(), x[r | (1 << J)].copy()
                        x[r] = np.where(tp, x1, x0)
                        x[r | (1 << J)] = np.where(tp, x0, x1)
                elif kind == 17:
                    k = int(o["k"])
                    tab = np.frombuffer(tables, dtype=np.complex128, count=1 << k, offset=int(o["table_off"]))
                    di = np.zeros_like(tb)
                    for p in range(k):
                        s_ = int(o["dsrc"][p])
                        if s_ == 0xFF:
                            continue
                        bit = ((gbase >> (s_ & 0x7F)) & 1) if (s_ & 0x80) else ((tb >> s_) & 1)
                        di |= bit << (k - 1 - p)
                    for r in rsel:
                        d = di.copy()
                        for q in range(RB):
                            if (r >> q) & 1:
                                d |= int(o["rc"][q])
                        x[r] = np.where(tp, x[r] * tab[d], x[r])
                else:
                    raise AssertionError("bad rop kind 0x%x" % kind)
            new_tile = tile.copy()
            dst = permute(tb, rb, rops[f0 + npre + nreg: f0 + npre + nreg + npost], gbase, gen_post)
            assert np.unique(dst).size == dst.size, "store addresses must be a permutation of the tile"
            new_tile[dst] = x
            tile = new_tile
        if base is None:
            return tile
        psi[base + gl_off] = tile
    return psi


def _emulate_whole_state_op(op, psi):
    idx = np.arange(psi.size)
    k = len(op.targets)
    c = np.zeros_like(idx)
    for p, pos in enumerate(op.targets):
        c |= ((idx >> pos) & 1) << (k - 1 - p)
    tmask = sum(1 << p for p in op.targets)

    def scatter(v):
        o = np.zeros_like(idx)
        for p, pos in enumerate(op.targets):
            o |= ((v >> (k - 1 - p)) & 1) << pos
        return o
    if op.kind == "wide_perm":
        out = np.empty_like(psi)
        out[(idx & ~tmask) | scatter(op.perm[c].astype(np.int64))] = psi
        return out
    if op.kind == "wide_diag":
        return psi * op.matrix[c]
    if op.kind == "wide_dense":
        out = np.zeros_like(psi)
        for col in range(1 << k):
            out += op.matrix[c, col] * psi[(idx & ~tmask) | scatter(np.full_like(idx, col))]
        return out
    raise AssertionError("emulation: unsupported op kind %s" % op.kind)


def emulate_ops(ops, psi, n_local, tile_bits=None, low_bits=None, rank_bits=0):
    """Run a list of qsv.schedule.Op on a numpy state the way engine.prepare/execute would: the schedule of
    qsv.schedule.plan_schedule, passes through the packed-program emulator, whole-state ops through direct
    numpy restatements."""
    from qsv.schedule import plan_schedule, pack_passes
    for kind, item in plan_schedule(ops, n_local, tile_bits, low_bits):
        if kind == "pass":
            psi = emulate_program(pack_passes([item]), psi, rank_bits)
        else:
            psi = _emulate_whole_state_op(item, psi)
    return psi


def swap_positions(full, swap_op, n_local):
    """Full state (physical order) after a 'global_swap' op: local position pos[k] trades places with global
    position n_local + gbits[k]."""
    from qsv.schedule import swap_fields
    pos, gb = swap_fields(swap_op, n_local)
    idx = np.arange(full.size)
    src = idx.copy()
    for p_loc, b_ in zip(pos, gb):
        gq = n_local + b_
        a, b = (idx >> p_loc) & 1, (idx >> gq) & 1
        src = (src & ~((1 << p_loc) | (1 << gq))) | (b << p_loc) | (a << gq)
    return full[src]


def emulate_sharded(ops, psi, n, g, tile_bits=None, low_bits=None):
    """Run PLACED ops (qsv.schedule.place) on P = 2^g virtual ranks held as rows of one numpy array, in the
    order of qsv.schedule.plan_schedule (passes may absorb ops from behind a qubit swap): passes go through the
    packed-program emulator with each rank's rank_bits, 'global_swap' is the exchange of the chosen local
    positions with the global bits.  Returns the full state in physical order."""
    from qsv.schedule import plan_schedule, pack_passes
    n_local = n - g
    P = 1 << g
    shards = psi.reshape(P, 1 << n_local).copy()
    for kind, item in plan_schedule(ops, n_local, tile_bits, low_bits):
        if kind == "pass":
            packed = pack_passes([item])
            for r in range(P):
                shards[r] = emulate_program(packed, shards[r], rank_bits=r << n_local)
        elif item.kind == "global_swap":
            if item.params.get("positions") is None and item.params.get("gbits") is None:
                t = shards.reshape(P, P, -1)            # [rank, top local field, rest]
                shards = np.ascontiguousarray(t.transpose(1, 0, 2)).reshape(P, -1)
            else:                                       # position pos[k] trades places with global bit gbits[k]
                shards = swap_positions(shards.reshape(-1), item, n_local).reshape(P, -1)
        else:
            for r in range(P):
                shards[r] = _emulate_whole_state_op(item, shards[r])
    return shards.reshape(-1)


# ------------------------------------------------------------------------------------------------
# pass-specialised kernels (csrc/qsv_jit.cu): the generated source is also valid C++; with
# -DQSV_JIT_HOST its round functions are driven sequentially over the thread ids, so the generator is
# checked against emulate_round_program in a container without a GPU.
# ------------------------------------------------------------------------------------------------

def jit_cpass(ps, rank_bits=0):
    from qsv import _cabi
    cp = _cabi.QsvPass()
    cp.n_local, cp.tile_bits, cp.low_bits = ps.n_local, ps.tile_bits, ps.low_bits
    cp.n_hi = ps.tile_bits - ps.low_bits
    for j, p in enumerate(ps.hi_pos):
        cp.hi_pos[j] = p
    cp.rank_bits = rank_bits
    cp.reserved = 4
    return cp


def jit_source(ps, prog):
    import ctypes as C
    from qsv import _cabi
    lib = _cabi.load()
    cp = jit_cpass(ps)
    need = C.c_size_t(0)
    _cabi.check(lib.qsv_jit_source(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), None, 0, C.byref(need)),
                "qsv_jit_source")
    buf = C.create_string_buffer(need.value)
    _cabi.check(lib.qsv_jit_source(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), buf, need.value,
                                   C.byref(need)), "qsv_jit_source")
    return buf.raw.decode()


def jit_source_sums(ps, prog, positions):
    import ctypes as C
    from qsv import _cabi
    lib = _cabi.load()
    cp = jit_cpass(ps)
    need = C.c_size_t(0)
    pos = _cabi.int_array(positions)
    _cabi.check(lib.qsv_jit_source_sums(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), len(positions), pos,
                                        None, 0, C.byref(need)), "qsv_jit_source_sums")
    buf = C.create_string_buffer(need.value)
    _cabi.check(lib.qsv_jit_source_sums(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), len(positions), pos,
                                        buf, need.value, C.byref(need)), "qsv_jit_source_sums")
    return buf.raw.decode()


def jit_host_run(ps, prog, tables, psi, rank_bits, workdir, single_tile_gbase=None, fixed_mask=0, fixed_val=0,
                 zf_mask=0, zf_val=0, sums=None):
    """Compile the generated source of one pass as host C++ and run it on psi (numpy complex128).  With
    single_tile_gbase, psi is one tile and only the rounds run (state sizes no host array can hold).
    sums = (bucket positions, uint64 array of 4 << len(positions) limbs): the sums variant of the kernel, which adds
    the weights of every tile it writes to the bucket array."""
    import ctypes as C
    import hashlib
    import os
    import subprocess
    src = jit_source(ps, prog) if sums is None else jit_source_sums(ps, prog, sums[0])
    tag = hashlib.sha1(src.encode()).hexdigest()[:16]
    cu, so = os.path.join(workdir, tag + ".cu"), os.path.join(workdir, tag + ".so")
    if not os.path.exists(so):
        with open(cu, "w") as f:
            f.write(src)
        r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-DQSV_JIT_HOST", "-x", "c++", cu,
                            "-o", so], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[:4000]
    h = C.CDLL(so)
    out = np.ascontiguousarray(psi, dtype=np.complex128).copy()
    tabs = np.frombuffer(bytes(tables) + b"\0" * 16, dtype=np.uint8).copy()
    if single_tile_gbase is not None:
        h.qsv_jit_host_tile(C.c_void_p(out.ctypes.data), C.c_uint64(single_tile_gbase), C.c_void_p(tabs.ctypes.data))
        return out
    if sums is not None:
        h.qsv_jit_host_set_buckets(C.c_void_p(sums[1].ctypes.data))
    h.qsv_jit_host_run(C.c_void_p(out.ctypes.data), C.c_uint64(rank_bits), C.c_uint64(1 << (ps.n_local - ps.tile_bits)),
                       C.c_void_p(tabs.ctypes.data), C.c_uint64(fixed_mask), C.c_uint64(fixed_val),
                       C.c_uint32(zf_mask), C.c_uint32(zf_val))
    return out


def jit_host_run_swap(ps, prog, tables, shards, positions, n_local, workdir, pass_first=0, gbits=None, support=(0, 0),
                      fresh=None):
    """Fused swap + pass on the host: `shards` = the shards BEFORE the swap (numpy complex128 arrays, one per rank).
    Returns the list of shards after swap + pass, one call of the generated qsv_jit_host_run_swap per rank.
    gbits: positions[k] trades places with global bit gbits[k] (partial swaps); support = (fixed_mask, fixed_val) of the
    state before the swap.  fresh = (zf_mask, zf_val): zero-fill form (qsv_run_pass_rounds_swap_fresh) - what lies outside
    the support in `shards` is never used as data."""
    import ctypes as C
    import hashlib
    import os
    import subprocess
    from qsv import _cabi
    lib = _cabi.load()
    g = len(positions)
    cp = jit_cpass(ps)
    pos = _cabi.int_array(positions)
    gb = _cabi.int_array(gbits if gbits is not None else list(range(g)))
    need = C.c_size_t(0)
    _cabi.check(lib.qsv_jit_source_swap_ex(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), g, pos, gb,
                                           pass_first, None, 0, C.byref(need)), "qsv_jit_source_swap_ex")
    buf = C.create_string_buffer(need.value)
    _cabi.check(lib.qsv_jit_source_swap_ex(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), g, pos, gb,
                                           pass_first, buf, need.value, C.byref(need)), "qsv_jit_source_swap_ex")
    src = buf.raw.decode()
    tag = "swap" + hashlib.sha1(src.encode()).hexdigest()[:16]
    cu, so = os.path.join(workdir, tag + ".cu"), os.path.join(workdir, tag + ".so")
    if not os.path.exists(so):
        with open(cu, "w") as f:
            f.write(src)
        r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-DQSV_JIT_HOST", "-x", "c++", cu,
                            "-o", so], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[:4000]
    h = C.CDLL(so)
    tabs = np.frombuffer(bytes(tables) + b"\0" * 16, dtype=np.uint8).copy()
    before = [np.ascontiguousarray(s_, dtype=np.complex128) for s_ in shards]
    ptrs = (C.c_void_p * len(before))(*[b.ctypes.data for b in before])
    out = []
    for rank in range(len(before)):
        # tiles the kernel skips keep what the shard held before (zeros, by the support invariant)
        dst = before[rank].copy() if support[0] else np.zeros(1 << n_local, dtype=np.complex128)
        h.qsv_jit_host_run_swap(C.c_void_p(dst.ctypes.data), ptrs, C.c_uint32(rank), C.c_uint64(rank << n_local),
                                C.c_uint64(1 << (n_local - ps.tile_bits)), C.c_void_p(tabs.ctypes.data),
                                C.c_uint64(support[0]), C.c_uint64(support[1]),
                                C.c_uint32(fresh[0] if fresh else 0), C.c_uint32(fresh[1] if fresh else 0),
                                C.c_uint32(1 if fresh is not None else 0))
        out.append(dst)
    return out


# ------------------------------------------------------------------------------------------------
# exact sampler (qsv/sampling.py + csrc/qsv_sample.cu): integer restatement on the host
# ------------------------------------------------------------------------------------------------

def weight_fix(w):
    """floor(w * 2^96) of a non-negative double, as the device computes it (weight_fix in csrc/qsv_sample.cu)."""
    from fractions import Fraction
    return int(Fraction(float(w)) * (1 << 96))


def exact_draw_bruteforce(weights_phys, n, src_pos, u):
    """bisect_right over the EXACT cumulative weights in output order: first z with cum(z) > floor(u * total)."""
    from qsv.sampling import u_fixed
    src = list(src_pos) if src_pos is not None else list(range(n))
    q = [weight_fix(w) for w in weights_phys]
    total = sum(q)
    mant, shift = u_fixed(u)
    target = (total * mant) >> shift
    cum = 0
    for z in range(1 << n):
        y = 0
        for b in range(n):
            y |= ((z >> b) & 1) << src[b]
        cum += q[y]
        if cum > target:
            return z
    return (1 << n) - 1


def exact_draw_emulated(weights_phys, n, g, src_pos, u, first_bits=0):
    """The radix descent of qsv.sampling.level_plan over 2^g virtual ranks, kernel by kernel (level / leaf bucket sums,
    integer all-reduce, select) with Python integers."""
    from qsv.sampling import level_plan, u_fixed
    n_local = n - g
    local = (1 << n_local) - 1
    q = [weight_fix(w) for w in weights_phys]
    mant, shift = u_fixed(u)
    fixed_val = z = 0
    target = None
    for step, (kind, out_bits, positions, fixed) in enumerate(level_plan(src_pos, n, n_local, first_bits=first_bits)):
        nb = len(out_bits)
        buckets = [0] * (1 << nb)
        if step == 0 and first_bits:
            assert nb == first_bits and fixed == 0      # bucket sums of the last gate pass: any positions, no run rule
        elif kind == "leaf":
            assert sorted(p for p in positions if p < n_local) == [p for p in range(n_local) if not (fixed >> p) & 1]
        else:
            run = [p for p in range(n_local) if not (fixed >> p) & 1][:10]
            assert sum(1 for p in positions if p in run) <= 2 and 1 <= nb <= 12
        for rank in range(1 << g):
            rank_bits = rank << n_local
            if (fixed_val ^ rank_bits) & fixed & ~local:
                continue
            for y in range(1 << n_local):
                if (y ^ fixed_val) & fixed & local:
                    continue
                full = y | rank_bits
                idx = 0
                for k, p in enumerate(positions):
                    idx |= ((full >> p) & 1) << k
                buckets[idx] += q[full]
        if target is None:
            total = sum(buckets)
            if total == 0:
                raise ValueError("Total of weights must be greater than zero")
            target = (total * mant) >> shift
        cum = 0
        found = (1 << nb) - 1
        for b_, v in enumerate(buckets):
            if cum + v > target:
                found = b_
                break
            cum += v
        target -= cum
        for k in range(nb):
            bit = (found >> k) & 1
            fixed_val |= bit << positions[k]
            z |= bit << out_bits[k]
    return z
