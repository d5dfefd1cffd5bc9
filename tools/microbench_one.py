"""One micro-benchmark pass (for ncu): 32 thread-level phases + 1 H in one round."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-circuits_b200"))
from qsv import engine
from qsv.schedule import Op
n = int(os.environ.get("N", "26"))
s = 2 ** -0.5
H = np.array([[s, s], [s, -s]], dtype=np.complex128)
os.environ["QSV_REG_MIN_OPS"] = "0"
sv = engine.StateVector(n)
sv.init_basis(0)
which = os.environ.get("CASE", "phase_t")
if which == "phase_t":
    ops = [Op("phase", (), (5,), scalar=complex(s, s)) for i in range(32)] + [Op("dense", (n - 1,), (), matrix=H)]
else:
    ops = [Op("dense", (n - 1 - (i % 4),), (), matrix=H) for i in range(32)]
plan = sv.prepare(ops)
for _ in range(4):
    sv.execute(plan)
sv.synchronize()
print(sv.time_passes(plan, repeats=3))
