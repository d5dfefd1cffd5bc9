"""Compile, without a GPU, every kernel the bench configurations will ask NVRTC for: the plain pass kernels, the fused
swap+pass kernels and the sums variant of the last pass (with the bucket positions the engine's cost rule picks) of
30..36 qubits on 1, 2, 4, 8 GPUs.  A kernel that NVRTC / ptxas rejected would only show on the GPU box, in the middle of
a measurement.     python tools/compile_bench_plans.py [--configs 30:1,33:1,31:2,...]"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-circuits_b200"))
os.environ.setdefault("QSV_JIT", "force")
from qsv import _cabi, engine, lowering, sampling, schedule, workloads   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="30:1,33:1,31:2,34:2,32:4,35:4,33:8,36:8")
ap.add_argument("--depth", type=int, default=20)
a = ap.parse_args()
lib = _cabi.load()


def cpass(ps, prog):
    cp = _cabi.QsvPass()
    cp.n_local, cp.tile_bits, cp.low_bits, cp.n_hi = ps.n_local, ps.tile_bits, ps.low_bits, ps.tile_bits - ps.low_bits
    for j, p in enumerate(ps.hi_pos):
        cp.hi_pos[j] = p
    cp.reserved = int(prog[12])
    return cp


def check(rc, what):
    if rc != 0:
        raise SystemExit("FAILED %s: %s" % (what, lib.qsv_last_error().decode("utf-8", "replace")))


total = 0
for cfg in a.configs.split(","):
    n, gpus = (int(x) for x in cfg.split(":"))
    g = (gpus - 1).bit_length()
    nl = n - g
    t0 = time.time()
    c = workloads.random_layered_circuit(n, a.depth, seed=20261019)
    c.sort()
    ops, src, flip = lowering.lower(c, n_global=g, in_flip=True, free_swap=bool(g), fuse_swaps=bool(g))
    items = schedule.plan_schedule_search(ops, nl, fuse_swaps=bool(g), n_global=g)
    # packed the way StateVector.prepare packs them: runs of passes together, the last run with last_as_rounds
    groups, batch = [], []
    for kind, it in items:
        if kind == "pass":
            batch.append(it)
            continue
        if batch:
            groups.append(("program", schedule.pack_passes(batch)))
            batch = []
        if kind == "swap_pass":
            groups.append(("swap_pass", (it[0], schedule.pack_passes([it[1]], min_ops=0), it[2])))
    if batch:
        groups.append(("program", schedule.pack_passes(batch, last_as_rounds=bool(groups) or len(batch) > 1)))
    n_plain = n_fused = n_sums = 0
    need = C.c_size_t(0)
    for gi, (kind, item) in enumerate(groups):
        if kind == "swap_pass":
            swap_op, packed, first = item
            meta = packed.passes[0]
            assert len(packed.passes) == 1 and meta[2] < 0
            prog = meta[3][0]
            cp = cpass(meta[0], prog)
            pos, gb = schedule.swap_fields(swap_op, nl)
            check(lib.qsv_jit_compile_swap_ex(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), len(pos),
                                              _cabi.int_array(pos), _cabi.int_array(gb), int(first), None, 0,
                                              C.byref(need)), "%s fused swap+pass" % cfg)
            n_fused += 1
            continue
        for mi, meta in enumerate(item.passes):
            if meta[2] >= 0:
                continue                        # shared-memory kernel (compiled with the library)
            prog = meta[3][0]
            cp = cpass(meta[0], prog)
            check(lib.qsv_jit_compile(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), None, 0,
                                      C.byref(need)), "%s pass" % cfg)
            n_plain += 1
            if gi == len(groups) - 1 and mi == len(item.passes) - 1:
                # the sums variant, with the bits StateVector._presum_plan would take
                ps = meta[0]
                static = lib.qsv_jit_sums_static(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size)) == 1
                inside_ms = engine.PRESUM_STATIC_MS if static else engine.PRESUM_GENERAL_MS
                tile = list(range(ps.low_bits)) + list(ps.hi_pos)
                best, k = engine.PRESUM_SWEEP_MS, 0
                for limit in range(4):
                    kk = sampling.presum_bits(src, n, tile, max_inside=limit)
                    if kk < min(engine.PRESUM_MIN_BITS, n):
                        continue
                    n_in = sum(1 for b in range(n - kk, n) if src[b] in set(tile))
                    ms = inside_ms[n_in] + engine.PRESUM_SWEEP_MS * 2.0 ** -kk
                    if ms < best - 1e-9:
                        best, k = ms, kk
                if k:
                    positions = [src[b] for b in range(n - k, n)]
                    cp.max_dense_k = 1
                    check(lib.qsv_jit_compile_sums(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), k,
                                                   _cabi.int_array(positions), None, 0, C.byref(need)),
                          "%s sums variant (%d bits, static=%s)" % (cfg, k, static))
                    n_sums += 1
                    sums_note = "%d bits, %s form" % (k, "static" if static else "general")
                else:
                    sums_note = "none"
    total += n_plain + n_fused + n_sums
    print("%2d qubits on %d GPU(s): %2d launch groups, %2d pass kernels + %d fused swap+pass + %d sums (%s) compiled for "
          "sm_100a in %.1f s" % (n, gpus, len(items), n_plain, n_fused, n_sums, sums_note, time.time() - t0), flush=True)
print("all %d kernels compiled" % total)
