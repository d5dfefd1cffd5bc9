"""Schedule of a BASELINE workload without a GPU: passes, rounds, swaps (plain / fused), the fraction of the shard
every pass touches when the register starts in a basis state (schedule.Support), and a modelled step time.

    python tools/plan_report.py --qubits 30                 # config 3 on one GPU
    python tools/plan_report.py --qubits 36 --gpus 8        # sharded: joint placement + fused swaps
    python tools/plan_report.py --qubits 33 --workload qft

Model (measured at 2^30 amplitudes per GPU, round 2: profiles/r02_bench_*): a full sweep costs max(5.2 ms, 1.75 ms per
round) - HBM-bound up to three rounds, bound by the SM beyond -, scaled by the touched fraction and by the shard size; a swap moves (1 - 2^-g) of the shard at 655 GB/s
and hides the pass fused into it."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-circuits_b200"))
os.environ.setdefault("QSV_JIT", "force")
from qsv import lowering, schedule, workloads   # noqa: E402
from qsv.schedule import Support                # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=30)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--depth", type=int, default=20)
ap.add_argument("--workload", default="layered", choices=["layered", "qft"])
ap.add_argument("--seed", type=int, default=20261019)
a = ap.parse_args()

g = (a.gpus - 1).bit_length()
n, n_local = a.qubits, a.qubits - g
c = workloads.random_layered_circuit(n, a.depth, seed=a.seed) if a.workload == "layered" else workloads.qft_ladder_circuit(n)
c.sort()
ops, src_pos, flip = lowering.lower(c, n_global=g, in_flip=True, free_swap=bool(g), fuse_swaps=bool(g))
in_bits = [0] * n if a.workload == "layered" else [(i + 1) % 2 for i in range(n)]
sup = Support(n, lowering.basis_index(in_bits) ^ flip)
scale = 2.0 ** (n_local - 30)
swap_ms = 16.0 * (1 << n_local) * (1 - 2.0 ** -g) / 655e9 * 1e3 if g else 0.0
local = (1 << n_local) - 1
total = dense = 0.0
print("%d qubits on %d GPU(s): %d gates -> %d ops" % (n, a.gpus, len(c.gates), len(ops)))
for kind, it in schedule.plan_schedule(ops, n_local, fuse_swaps=bool(g)):
    if kind == "pass":
        fixed, _ = sup.pass_mask(it)
        frac = Support.fraction(fixed & local)
        rounds = schedule.plan_rounds(it, reg_mcx=True, padded_layout=True)
        nr = len(rounds) if rounds else 0
        sweep = max(5.2, 1.75 * nr + (0.5 if nr == 3 else 0.0)) * scale
        total += 0.1 + sweep * frac
        dense += 0.1 + sweep
        rb = " ".join("".join("%x" % b for b in sorted(R.rb)) for R in rounds) if rounds else "shared-memory kernel"
        gfix = bin(fixed >> n_local).count("1")
        print("  pass  %3d gates  %d rounds [%s]  tile hi %s  touches %.3g of the shard on %d of %d ranks"
              % (len(it.ops), nr, rb, list(it.hi_pos), frac, a.gpus >> gfix, a.gpus))
        sup.after_ops(it.ops)
    elif kind == "swap_pass":
        sw, ps, first = it
        j = int(sw.params["n_swap"])
        swap_ms = 16.0 * (1 << n_local) * (1 - 2.0 ** -j) / 655e9 * 1e3
        total += swap_ms
        dense += swap_ms
        print("  SWAP %s fused with a pass of %d gates (%s)" % (sw.params["positions"], len(ps.ops), "pass first" if first else "pass after"))
        if first:
            sup.after_ops(ps.ops)
            sup.swap_op(sw, n_local)
        else:
            sup.swap_op(sw, n_local)
            sup.after_ops(ps.ops)
    else:
        pos, gb = schedule.swap_fields(it, n_local)
        j = len(pos)
        frac = min(1.0, 2.0 * Support.fraction(sup.fixed()[0] & local))
        one = 16.0 * (1 << n_local) * (1 - 2.0 ** -j) / 655e9 * 1e3
        total += 0.1 + one * frac
        dense += one
        print("  SWAP %s <-> global bits %s  moves %.3g of its pairs" % (pos, gb, frac))
        sup.swap_op(it, n_local)
print("modelled step: %.1f ms (%.0f gates/s); with full sweeps %.1f ms (%.0f gates/s)"
      % (total, len(c.gates) / total * 1e3, dense, len(c.gates) / dense * 1e3))
