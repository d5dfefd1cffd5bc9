"""Generate tests/golden/* from the REFERENCE itself.  TEST INFRASTRUCTURE ONLY.

Run in the build container, where the reference is mounted read-only:

    python oracle/gen_golden.py            # ~3-4 minutes of single-core CPython

It imports the unmodified reference (`/root/reference/main/*.py`), runs its own
Circuit.run_method1 / run_method2 / run / Shor on small circuits and stores inputs + outputs as
fixtures.  Nothing of the reference's source is copied: circuits are serialised structurally
(gate class / matrix entries / wiring) and the reference's results numerically.  Final amplitudes —
which no reference function returns — are captured by spying on the reference's own
Matrix.__mul__: the last product of run_method2 (main/circuit.py:332) is the final state.

The GPU box has no /root/reference; tests there read only the fixtures written here.
"""
import io
import json
import math
import os
import random
import sys
import contextlib
import time

REF = os.environ.get("QSV_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

import numpy as np                      # noqa: E402
import main.circuit as rc               # noqa: E402  (the reference)
import main.matrix as rm                # noqa: E402
import statevec_oracle as so            # noqa: E402

assert rc.__file__.startswith(REF), "gen_golden must import the reference, got %s" % rc.__file__

_last_product = [None]
_orig_mul = rm.Matrix.__mul__


def _spy_mul(self, other):
    r = _orig_mul(self, other)
    _last_product[0] = r
    return r


def ref_weights_and_amps(c, in_v):
    """Reference run_method2 weights + the final state column it computed them from."""
    rm.Matrix.__mul__ = _spy_mul
    try:
        w = c.run_method2(list(in_v))
    finally:
        rm.Matrix.__mul__ = _orig_mul
    amps = [complex(r[0]) for r in _last_product[0].rows]
    return w, amps


def cplx(z):
    z = complex(z)
    return [z.real, z.imag]


_matrix_store = {}


def store_matrix(rows):
    """Gate matrices go to arrays.npz, de-duplicated by content (Grover repeats three of them 6x)."""
    import hashlib
    M = np.array(rows, dtype=np.complex128)
    key = "matrix/" + hashlib.sha1(M.tobytes()).hexdigest()[:16]
    _matrix_store[key] = M
    return key


def serialise(c, name):
    """Structural description of a reference circuit in its ORIGINAL insertion order (the order the
    reference's depth-first sort starts from)."""
    gates = []
    index = {}
    for i, g in enumerate(c.gates):
        index[id(g)] = i
        cls = type(g).__name__
        entry = {"cls": cls, "name": g.name, "size": len(g)}
        if cls == "Gate":
            entry["matrix_key"] = store_matrix(g.matrix.rows)
        gates.append(entry)
    wires = []
    for w in c.wires:
        src = -1 if w.left is c else index[id(w.left)]
        dst = -1 if w.right is c else index[id(w.right)]
        wires.append([src, w.lind, dst, w.rind])
    return {"name": name, "n": len(c), "gates": gates, "wires": wires}


def circuit_case(c, name, inputs, with_m1=False, arrays=None):
    case = serialise(c, name)
    before = {id(g): i for i, g in enumerate(c.gates)}
    c.sort()
    case["sorted_order"] = [before[id(g)] for g in c.gates]
    case["sorted_gate_names"] = [str(g) for g in c.gates]
    case["targets"] = [[c.startqbit(i, p) for p in range(len(g))] for i, g in enumerate(c.gates)]
    case["startqbits"] = c.startqbits()
    case["inputs"] = [list(v) for v in inputs]
    W, A, W1 = [], [], []
    for v in inputs:
        w, a = ref_weights_and_amps(c, v)
        W.append(w)
        A.append(a)
        if with_m1:
            W1.append(c.run_method1(list(v)))
    arrays[name + "/weights_m2"] = np.array(W, dtype=np.float64)
    arrays[name + "/amps"] = np.array(A, dtype=np.complex128)
    if with_m1:
        arrays[name + "/weights_m1"] = np.array(W1, dtype=np.float64)
    return case


def build_from_steps(steps, n):
    """Reference circuit from (matrix, wires, label) steps, each gate wired on its wires in order."""
    c = rc.Circuit(n)
    last = [(c, i) for i in range(n)]
    for k, (M, wires, label) in enumerate(steps):
        rows = [[(complex(v) if v.imag else float(v.real)) for v in row] for row in np.asarray(M)]
        g = c.add(rc.Gate(rm.Matrix(rows), "%s_%d" % (label, k), True))
        for p, w in enumerate(wires):
            c.add_wire(last[w][0], last[w][1], g, p)
            last[w] = (g, p)
    for w in range(n):
        c.add_wire(last[w][0], last[w][1], c, w)
    return c


def main():
    t0 = time.time()
    os.makedirs(OUT, exist_ok=True)
    arrays = {}
    cases = []
    meta = {"reference": "MatejVitek/Quantum-Circuits", "python": sys.version.split()[0]}

    # --- A: paper circuit (5), main/test.py:4-36 + truth table main/computational-tests.py:9-18
    import main.test as rt
    c = rt.create_test_circuit()
    ins = [[a, b, 0, 0, 0, 0] for a in (0, 1) for b in (0, 1)]
    case = circuit_case(c, "paper5", ins, with_m1=True, arrays=arrays)
    outs = []
    for v in ins:
        with contextlib.redirect_stdout(io.StringIO()):
            outs.append([c.run(list(v)), c.run(list(v), 2)])
    case["run_outputs"] = outs
    cases.append(case)
    print("paper5 done", time.time() - t0)

    # --- B: reference random_circuit generator under fixed seeds (main/circuit.py:406-445)
    for seed in range(1, 13):
        random.seed(seed)
        c = rc.Circuit.random_circuit()
        n = len(c)
        if n <= 4:
            ins = [list(v) for v in __import__("itertools").product((0, 1), repeat=n)]
        else:
            r = random.Random(1000 + seed)
            ins = [[0] * n, [1] * n] + [[r.randrange(2) for _ in range(n)] for _ in range(2)]
        case = circuit_case(c, "random_seed%d" % seed, ins, with_m1=(n <= 5), arrays=arrays)
        case["seed"] = seed
        cases.append(case)
    print("random circuits done", time.time() - t0)

    # --- C: Grover, main/Grover.py (7 qubits, 25 full-width gates)
    import main.Grover as rg
    case = circuit_case(rg.c, "grover7", [[0] * 6 + [1]], arrays=arrays)
    case["marked"] = rg.x
    cases.append(case)
    print("grover7 done", time.time() - t0)

    # --- D: Shor N=15 circuits for every admissible a, captured from main/Shor.py:34 itself
    import main.Shor as rs
    captured = {}
    orig_run = rc.Circuit.run

    class _Stop(Exception):
        pass

    def capture_run(self, in_v, method=None):
        captured["c"], captured["in_v"] = self, list(in_v)
        raise _Stop()

    shor_perms = {}
    for a in (2, 4, 7, 8, 11, 13, 14):
        rs.randint = lambda lo, hi, a=a: a
        rc.Circuit.run = capture_run
        try:
            rs.Shor(15)
        except _Stop:
            pass
        finally:
            rc.Circuit.run = orig_run
            rs.randint = random.randint
        c = captured["c"]
        U = [g for g in c.gates if g.name == "U"][0].matrix
        perm = [row_i for col in range(len(U)) for row_i in range(len(U)) if U[row_i][col] == 1]
        shor_perms["15_%d" % a] = perm
        if a in (2, 7, 13):
            case = circuit_case(c, "shor15_a%d" % a, [captured["in_v"]], with_m1=(a == 7), arrays=arrays)
            # store U compactly: the tests rebuild the dense matrix from the permutation
            for g in case["gates"]:
                if g["name"] == "U":
                    _matrix_store.pop(g.pop("matrix_key"), None)
                    g["perm_matrix"] = "15_%d" % a
            case["N"], case["a"] = 15, a
            cases.append(case)
        print("shor15 a=%d done" % a, time.time() - t0)

    # --- E: Shor_example (N=21, a=17, 10 qubits): modexp permutation + method-1 weights
    import main.Shor_example as rse
    U = [g for g in rse.c.gates if g.name == "U"][0].matrix
    shor_perms["21_17"] = [r for col in range(len(U)) for r in range(len(U)) if U[r][col] == 1]
    w1 = rse.c.run_method1([0] * 5 + [1] * 5)
    arrays["shor21_a17/weights_m1"] = np.array([w1], dtype=np.float64)
    meta["shor21"] = {"N": 21, "a": 17, "q": 10, "input": [0] * 5 + [1] * 5}
    print("shor21 done", time.time() - t0)
    for k, p in shor_perms.items():
        arrays["modexp_perm/" + k] = np.array(p, dtype=np.int64)

    # --- F: Shor(15) end to end under fixed seeds: every run() outcome and the factor list
    seq = {}
    for seed in (0, 1, 2):
        random.seed(seed)
        log = []

        def spy_run(self, in_v, method=None, log=log):
            with contextlib.redirect_stdout(io.StringIO()):
                out = orig_run(self, in_v, method)
            log.append(out)
            return out

        rc.Circuit.run = spy_run
        try:
            factors = rs.Shor(15)
        finally:
            rc.Circuit.run = orig_run
        seq[str(seed)] = {"runs": log, "factors": factors}
    meta["shor15_sequences"] = seq
    print("shor15 sequences done", time.time() - t0)

    # --- G: config-3 layered random circuits (generator restated in the oracle), small registers
    for n, depth in ((5, 3), (6, 2), (7, 2)):
        steps = so.random_layered_steps(n, depth)
        c = build_from_steps(steps, n)
        r = random.Random(n)
        ins = [[0] * n, [r.randrange(2) for _ in range(n)]]
        cases.append(circuit_case(c, "layered_n%d_d%d" % (n, depth), ins, arrays=arrays))
    print("layered done", time.time() - t0)

    # --- I: built-in multi-qubit gate classes on scattered wires (QFT, H(size), Y, SqrtNot, Z)
    c = rc.Circuit(5)
    h = c.add_gate(rc.H, 2)
    q = c.add_gate(rc.QFT, 3)
    y = c.add_gate(rc.Y)
    s = c.add_gate(rc.SqrtNot, 2)
    z = c.add_gate(rc.Z)
    c.add_wires(c, (4, 1), h, ())
    c.add_wire(c, 0, q, 2)
    c.add_wire(h, 0, q, 0)
    c.add_wire(c, 3, q, 1)
    c.add_wire(h, 1, y, 0)
    c.add_wire(c, 2, s, 0)
    c.add_wire(q, 1, s, 1)
    c.add_wire(q, 0, z, 0)
    c.add_wire(z, 0, c, 3)
    c.add_wire(q, 2, c, 0)
    c.add_wire(y, 0, c, 4)
    c.add_wires(s, (), c, (1, 2))
    r = random.Random(5)
    ins = [[0] * 5, [1, 0, 1, 1, 0], [1] * 5]
    cases.append(circuit_case(c, "builtin_mix", ins, with_m1=True, arrays=arrays))
    print("builtin mix done", time.time() - t0)

    # --- J: get_subcircuit (main/circuit.py:122-157) on the small random circuits + a crossed-wire one
    sub = {}
    for seed in range(1, 13):
        random.seed(seed)
        c = rc.Circuit.random_circuit()
        if len(c) <= 4:
            try:
                g = c.get_subcircuit("sub")
            except RuntimeError as e:        # the reference's own unitarity check can reject its result
                sub["random_seed%d" % seed] = {"n": len(c), "error": str(e)}
                continue
            arrays["subcircuit/random_seed%d" % seed] = np.array(g.matrix.rows, dtype=np.complex128)
            sub["random_seed%d" % seed] = {"n": len(c), "error": None}
    meta["subcircuit_cases"] = sub

    # --- H: Matrix factories / arithmetic (main/matrix.py)
    mats = {
        "H1": rm.Matrix.H(), "H3": rm.Matrix.H(3), "X2": rm.Matrix.X(2), "Y2": rm.Matrix.Y(2),
        "Z2": rm.Matrix.Z(2), "SqrtNot2": rm.Matrix.SqrtNot(2), "Phase0.3_2": rm.Matrix.PhaseShift(0.3, 2),
        "QFT1": rm.Matrix.QFT(1), "QFT3": rm.Matrix.QFT(3), "QFT4": rm.Matrix.QFT(4), "Cnot": rm.Matrix.Cnot(),
        "T": rm.Matrix.T(), "Perm201": rm.Matrix.Permutation([2, 0, 1]), "Id4": rm.Matrix.Id(4),
        "Zero2x3": rm.Matrix.Zero(2, 3),
    }
    A = rm.Matrix.QFT(2)
    B = rm.Matrix.SqrtNot(2)
    mats["QFT2*SqrtNot2"] = A * B
    mats["QFT2.conj"] = A.getConjugate()
    mats["QFT2.round3"] = A.Round(3)
    mats["tensor3"] = rm.tensor([rm.Matrix.H(), rm.Matrix.Y(), rm.Matrix.PhaseShift(1.1)])
    mats["matrix_product"] = rm.matrix_product([rm.Matrix.H(), rm.Matrix.Y(), rm.Matrix.SqrtNot()])
    v = rm.Matrix.vector([1, 0, 0, 0])
    mats["apply"] = v.apply([rm.Matrix.H(2), rm.Matrix.Cnot()])
    meta["matrices"] = {k: [[cplx(x) for x in row] for row in m.rows] for k, m in mats.items()}
    meta["matrix_int_entries"] = {k: all(isinstance(x, int) for row in m.rows for x in row) for k, m in mats.items()}
    meta["isUnitary"] = {"QFT3": rm.Matrix.QFT(3).isUnitary(), "T": rm.Matrix.T().isUnitary(),
                         "nonunitary": rm.Matrix([[1, 1], [0, 1]]).isUnitary()}
    meta["str_Cnot"] = str(rm.Matrix.Cnot())
    meta["repr_X"] = repr(rm.Matrix.X())

    # --- tests.txt (output artefact of Shor_test(32), main/tests.txt): factor multisets
    facts = {}
    with open(os.path.join(REF, "main", "tests.txt")) as f:
        for line in f:
            left, right = line.strip().split(":")
            facts[left.split()[-1]] = sorted(json.loads(right))
    meta["tests_txt_factors"] = facts

    # --- cost-model decisions of run() (main/circuit.py:343-351)
    meta["method_choice"] = {"grover7": so.choose_method(len(rg.c.gates), 7, len(rg.c.get_internal_wires())),
                             "shor21": so.choose_method(len(rse.c.gates), 10, len(rse.c.get_internal_wires()))}

    arrays.update(_matrix_store)
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump({"meta": meta, "cases": cases}, f)
    np.savez_compressed(os.path.join(OUT, "arrays.npz"), **arrays)
    print("wrote", OUT, "in %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
