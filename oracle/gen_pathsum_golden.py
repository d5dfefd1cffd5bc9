"""Generate tests/golden/pathsum_direct.json from the REFERENCE itself.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_pathsum_golden.py        # seconds; needs the reference at /root/reference (build container)

Circuits in which circuit inputs are wired STRAIGHT to circuit outputs, run through the unmodified reference's
Circuit.run_method1 (the Feynman path sum, main/circuit.py:221-276: such a wire lets only out == in through,
:240-252) and run_method2 for every basis input.  The fixtures pin the oracle's loop-for-loop restatement
(statevec_oracle.run_method1_path_sum) and the packed description the CUDA kernel evaluates (qsv.pathsum) on exactly
the branch the random golden circuits rarely take.  Circuits are stored structurally, results numerically.
"""
import itertools
import json
import os
import sys

REF = os.environ.get("QSV_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "pathsum_direct.json")
sys.path.insert(0, REF)

import main.circuit as rc               # noqa: E402  (the reference)

assert rc.__file__.startswith(REF), "must import the reference, got %s" % rc.__file__


def build(n, gates, wires):
    """gates: [(class name, ctor args)], wires: [src gate index or -1, src port, dst gate index or -1, dst port]."""
    c = rc.Circuit(n)
    objs = [c.add(getattr(rc, cls)(*args)) for cls, args in gates]
    for s, sp, d, dp in wires:
        c.add_wire(c if s < 0 else objs[s], sp, c if d < 0 else objs[d], dp)
    return c


CASES = [
    # one H, two direct wires that cross
    ("h_two_direct_crossing", 3, [("H", (1, "h"))], [[-1, 0, -1, 2], [-1, 1, 0, 0], [0, 0, -1, 0], [-1, 2, -1, 1]]),
    # a CNot whose outputs cross, two direct wires that cross
    ("cnot_crossing_two_direct", 4, [("CNot", ("cx",))],
     [[-1, 0, 0, 0], [-1, 1, 0, 1], [0, 0, -1, 1], [0, 1, -1, 0], [-1, 2, -1, 3], [-1, 3, -1, 2]]),
    # Toffoli, then H on its target: one internal wire, one direct wire in the middle of the register
    ("toffoli_h_one_direct", 4, [("T", ("t",)), ("H", (1, "h"))],
     [[-1, 0, 0, 0], [-1, 1, -1, 1], [-1, 2, 0, 1], [-1, 3, 0, 2], [0, 0, -1, 0], [0, 1, -1, 2], [0, 2, 1, 0],
      [1, 0, -1, 3]]),
    # H, CZ-like phase structure through H X H with a direct wire: two internal wires
    ("hxh_chain_one_direct", 2, [("H", (1, "h1")), ("X", (1, "x")), ("H", (1, "h2"))],
     [[-1, 0, 0, 0], [0, 0, 1, 0], [1, 0, 2, 0], [2, 0, -1, 1], [-1, 1, -1, 0]]),
]


def main():
    out = []
    for name, n, gates, wires in CASES:
        c = build(n, gates, wires)
        inputs = [list(v) for v in itertools.product((0, 1), repeat=n)]
        w1 = [[float(x) for x in c.run_method1(list(v))] for v in inputs]
        c2 = build(n, gates, wires)
        w2 = [[float(x) for x in c2.run_method2(list(v))] for v in inputs]
        out.append({"name": name, "n": n, "gates": [{"cls": cls, "args": list(args)} for cls, args in gates],
                    "wires": wires, "inputs": inputs, "weights_m1": w1, "weights_m2": w2})
        print(name, "ok: method 1 == method 2 within", max(abs(a - b) for r1, r2 in zip(w1, w2) for a, b in zip(r1, r2)))
    with open(OUT, "w") as f:
        json.dump({"reference": "MatejVitek/Quantum-Circuits", "cases": out}, f)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
