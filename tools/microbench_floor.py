"""Floor of the pass kernels at full size: ONE H gate per sweep through (a) the shared-memory kernel k_pass, (b) the
pass-specialised kernel storing from registers, (c) the pass-specialised kernel storing through TMA.  Run on a B200:
    N=30 python tools/microbench_floor.py            (CASE=a|b|c runs one variant only, for ncu)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-circuits_b200"))
from qsv import engine
from qsv.schedule import Op

n = int(os.environ.get("N", "30"))
s = 2 ** -0.5
H = np.array([[s, s], [s, -s]], dtype=np.complex128)
sv = engine.StateVector(n)
sv.init_basis(0)
case = os.environ.get("CASE", "abc")
reps = int(os.environ.get("REPS", "10"))


def run(label, ops, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    engine._PLAN_CACHE.clear()
    plan = sv.prepare(ops)
    for _ in range(3):
        sv.execute(plan)
    ms = sv.time_passes(plan, repeats=reps)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k)
        else:
            os.environ[k] = v
    nbytes = 32.0 * (1 << n)
    print("%-58s ms=%s  %s GB/s" % (label, [round(x, 3) for x in ms], [round(nbytes / x / 1e6) for x in ms]), flush=True)


for q in (n - 1, n // 2, 7):
    ops = [Op("dense", (q,), (), matrix=H)]
    if "a" in case:
        run("a: k_pass (smem, TMA in/out, 3 CTAs/SM)        H q%d" % q, ops, {"QSV_REGISTER_KERNEL": "0"})
    if "b" in case:
        run("b: qsv_jit_pass, registers -> STG               H q%d" % q, ops, {"QSV_REG_MIN_OPS": "0", "QSV_JIT": "force"})
    if "c" in case:
        run("c: qsv_jit_pass, TMA store (QSV_JIT_DIRECT_STORE=0) H q%d" % q, ops,
            {"QSV_REG_MIN_OPS": "0", "QSV_JIT": "force", "QSV_JIT_DIRECT_STORE": "0"})
ops4 = [Op("dense", (n - 1 - i,), (), matrix=H) for i in range(4)]
if "b" in case:
    run("b: qsv_jit_pass STG, 4 H one round", ops4, {"QSV_REG_MIN_OPS": "0", "QSV_JIT": "force"})
if "c" in case:
    run("c: qsv_jit_pass TMA store, 4 H one round", ops4, {"QSV_REG_MIN_OPS": "0", "QSV_JIT": "force", "QSV_JIT_DIRECT_STORE": "0"})
