"""Micro-benchmarks of the pass kernels: cost per round and per op (run on a B200 via gpurun)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-circuits_b200"))
import torch
from qsv import engine, schedule
from qsv.schedule import Op

n = int(os.environ.get("N", "28"))
s = 2 ** -0.5
H = np.array([[s, s], [s, -s]], dtype=np.complex128)
G = np.array([[0.6, 0.8j], [0.8j, 0.6]], dtype=np.complex128)
sv = engine.StateVector(n)
sv.init_basis(0)
os.environ["QSV_REG_MIN_OPS"] = "0"

def run(label, ops, reps=3):
    plan = sv.prepare(ops)
    for _ in range(2):
        sv.execute(plan)
    ms = sv.time_passes(plan, repeats=reps)
    packed = plan["segments"][0][1][0]
    nr = [m[3][1] if m[2] < 0 else 0 for m in packed.passes]
    print("%-44s passes=%d rounds=%s ms=%s" % (label, len(ms), nr, [round(x, 3) for x in ms]), flush=True)

hi = [n - 1 - i for i in range(6)]          # 6 high qubits -> tile hi bits
run("1 op H (round kernel)", [Op("dense", (hi[0],), (), matrix=H)])
run("4 H on 4 qubits (1 round)", [Op("dense", (hi[i],), (), matrix=H) for i in range(4)])
run("32 H on 4 qubits (1 round)", [Op("dense", (hi[i % 4],), (), matrix=H) for i in range(32)])
run("32 complex dense on 4 qubits (1 round)", [Op("dense", (hi[i % 4],), (), matrix=G) for i in range(32)])
run("12 H on 12 qubits (3 rounds)", [Op("dense", (q,), (), matrix=H) for q in list(range(6)) + hi])
run("36 H on 12 qubits (3 rounds)", [Op("dense", (q,), (), matrix=H) for _ in range(3) for q in list(range(6)) + hi])
run("32 phase_t (1 round)", [Op("phase", (), (5,), scalar=complex(s, s)) for i in range(32)] + [Op("dense", (hi[0],), (), matrix=H)])
run("32 phase1 (1 round)", [Op("dense", (hi[0],), (), matrix=H)] + [Op("phase", (), (hi[0],), scalar=complex(s, s)) for i in range(32)])
run("32 cz thread bits (1 round)", [Op("dense", (hi[0],), (), matrix=H)] + [Op("phase", (), (3, 4), scalar=-1 + 0j) for i in range(32)])
run("8 CNot post perms (1 round)", [Op("dense", (hi[0],), (), matrix=H)] + [Op("mcx", (hi[1 + i % 3],), (i % 5,)) for i in range(8)])
run("H,CNot chain 6x (rounds)", sum([[Op("dense", (hi[i % 4],), (), matrix=H), Op("mcx", (hi[(i + 1) % 4],), (hi[i % 4],))] for i in range(6)], []))
os.environ["QSV_REGISTER_KERNEL"] = "0"
run("smem kernel: 1 op H", [Op("dense", (hi[0],), (), matrix=H)])
run("smem kernel: 4 H", [Op("dense", (hi[i],), (), matrix=H) for i in range(4)])
run("smem kernel: 32 H on 4 qubits", [Op("dense", (hi[i % 4],), (), matrix=H) for i in range(32)])
