"""Dump the generated CUDA source (and, with --sass, the SASS of the NVRTC-compiled sm_100a cubin) of one pass of
the config-3 circuit.  Works without a GPU.   python tools/jit_dump.py [--qubits 30] [--pass 4] [--sass] [--out DIR]"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-circuits_b200"))
os.environ.setdefault("QSV_JIT", "force")
from qsv import _cabi, lowering, schedule, workloads   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=30)
ap.add_argument("--depth", type=int, default=20)
ap.add_argument("--pass", dest="k", type=int, default=4)
ap.add_argument("--sass", action="store_true")
ap.add_argument("--out", default="/tmp")
a = ap.parse_args()

lib = _cabi.load()
c = workloads.random_layered_circuit(a.qubits, a.depth, seed=20261019)
c.sort()
ops, _, _ = lowering.lower(c, in_flip=True)
passes = [it for kind, it in schedule.plan_schedule_search(ops, a.qubits) if kind == "pass"]      # the schedule bench.py runs
ps = passes[a.k]
packed = schedule.pack_passes([ps])
prog = packed.passes[0][3][0]
cp = _cabi.QsvPass()
cp.n_local, cp.tile_bits, cp.low_bits, cp.n_hi = ps.n_local, ps.tile_bits, ps.low_bits, ps.tile_bits - ps.low_bits
for j, p in enumerate(ps.hi_pos):
    cp.hi_pos[j] = p
cp.reserved = int(prog[12])
need = C.c_size_t(0)
_cabi.check(lib.qsv_jit_source(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), None, 0, C.byref(need)), "source")
buf = C.create_string_buffer(need.value + 1)
_cabi.check(lib.qsv_jit_source(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), buf, need.value, C.byref(need)), "source")
src_path = os.path.join(a.out, "qsv_jit_pass%d.cu" % a.k)
open(src_path, "wb").write(buf.raw[:need.value])
print("pass %d: %d ops, hi_pos %s -> %s (%d bytes)" % (a.k, len(ps.ops), ps.hi_pos, src_path, need.value))
if a.sass:
    _cabi.check(lib.qsv_jit_compile(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), None, 0, C.byref(need)), "compile")
    cub = C.create_string_buffer(need.value)
    _cabi.check(lib.qsv_jit_compile(C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size), cub, need.value, C.byref(need)), "compile")
    cub_path = os.path.join(a.out, "qsv_jit_pass%d.cubin" % a.k)
    open(cub_path, "wb").write(cub.raw)
    sass = subprocess.run(["cuobjdump", "-sass", cub_path], capture_output=True, text=True).stdout
    open(cub_path + ".sass", "w").write(sass)
    from collections import Counter
    ops_ = Counter(ln.split(";")[0].split()[1].split(".")[0] for ln in sass.splitlines()
                   if ln.strip().startswith("/*") and len(ln.split()) > 2 and ln.split()[1][0].isalpha())
    print("SASS: %d instructions; top: %s" % (sum(ops_.values()), ops_.most_common(14)))
    for m in ("UBLKCP", "SYNCS", "ELECT", "DFMA", "DMUL", "DADD", "HMMA", "UTCMMA"):
        print("  %-7s %d" % (m, sum(1 for ln in sass.splitlines() if (" " + m) in ln)))
