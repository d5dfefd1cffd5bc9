"""The "run unchanged" claim and the .cir row.

* the reference's own scripts main/Grover.py, main/Shor.py, main/Shor_example.py, main/test.py import on top of the
  mirror package and their circuits lower to kernel ops (CPU; needs the reference checkout, skipped without it);
* circuits saved by the reference GUI (tests/golden/saved/*.cir.xz, copied from Resources/Saved by
  oracle/gen_saved_fixtures.py) unpickle into mirror objects and - on the GPU - run through the engine to the weights
  the unmodified reference computed for them;
* circuits pickled by the mirror load into the REFERENCE's classes and run there (CPU, needs the checkout)."""
import lzma
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quantum-circuits_b200")
SAVED = os.path.join(ROOT, "tests", "golden", "saved")
REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "main")), reason="reference checkout not present")


def _load_saved(name):
    import main.circuit as mc       # noqa: F401  (the pickles name main.circuit / main.matrix classes)
    with lzma.open(os.path.join(SAVED, name + ".cir.xz"), "rb") as f:
        return pickle.load(f)       # first pickle of the file = the Circuit (gui/scene.py:73-81)


@needs_ref
def test_reference_scripts_import_unchanged_over_the_mirror(golden):
    """A fresh interpreter with only the mirror on sys.path: main.circuit / main.matrix are ours, the algorithm
    scripts are the reference's files; their circuits lower without a dense full-width fallback."""
    code = textwrap.dedent("""
        import json, sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import main.circuit, main.matrix, main.Grover, main.Shor, main.Shor_example, main.test
        from qsv import lowering
        import statevec_oracle as so
        out = {"files": [main.circuit.__file__, main.matrix.__file__, main.Grover.__file__, main.Shor.__file__,
                         main.Shor_example.__file__, main.test.__file__]}
        g = main.Grover.c
        g.sort()
        ops, src = lowering.lower(g)
        out["grover_kinds"] = sorted({o.kind for o in ops}); out["grover_ops"] = len(ops)
        out["grover_w"] = so.run_method2(g, [0] * 6 + [1]).tolist()
        s = main.Shor_example.c
        s.sort()
        ops, src = lowering.lower(s)
        out["shor_kinds"] = sorted({o.kind for o in ops})
        t = main.test.create_test_circuit()
        out["paper_w"] = [so.run_method2(t, v).tolist() for v in ([0]*6, [0,1,0,0,0,0], [1,0,0,0,0,0], [1,1,0,0,0,0])]
        out["shor_fn"] = callable(main.Shor.Shor) and main.Shor.Matrix is main.matrix.Matrix
        print("RESULT" + json.dumps(out))
    """) % (os.path.join(ROOT, "oracle"), PKG)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1][6:])
    assert all(f.startswith(PKG) for f in out["files"][:2]) and all(f.startswith(REF) for f in out["files"][2:])
    assert set(out["grover_kinds"]) <= {"dense", "diag", "mcx", "phase"} and out["grover_ops"] < 200
    assert not {"wide_dense"} & set(out["shor_kinds"])          # modexp U -> exact permutation, QFT(5) -> dense tile op
    meta, cases, arrays = golden
    assert np.max(np.abs(np.array(out["grover_w"]) - arrays["grover7/weights_m2"][0])) <= 1e-12
    assert np.array_equal(np.array(out["paper_w"]), arrays["paper5/weights_m2"][:4])
    assert out["shor_fn"]


def test_saved_circuits_of_the_reference_gui_load_into_the_mirror():
    import main.circuit as mc
    from main.matrix import Matrix
    import statevec_oracle as so
    ref_w = np.load(os.path.join(SAVED, "weights.npz"))
    c = _load_saved("5")
    assert type(c) is mc.Circuit and len(c) == 6 and c.check()
    assert sorted(type(g).__name__ for g in c.gates) == ["CNot", "CNot", "CNot", "T", "X", "X", "X"]
    for v, w in zip(ref_w["5/inputs"], ref_w["5/weights_m2"]):
        assert np.array_equal(so.run_method2(c, [int(b) for b in v]), w)
    g = _load_saved("Grover")
    assert len(g) == 7 and len(g.gates) == 25 and isinstance(g.gates[1].matrix, Matrix)
    w = so.run_method2(g, [int(b) for b in ref_w["Grover/inputs"][0]])
    assert np.max(np.abs(w - ref_w["Grover/weights_m2"][0])) <= 1e-12


@needs_ref
def test_circuits_pickled_by_the_mirror_run_on_the_unmodified_reference(golden, tmp_path):
    """Mirror -> reference: the pickle names main.circuit.Circuit etc. with the reference's attribute layout
    (gates as lists of Gate objects whose matrix.rows are lists), so the reference's classes load it and its own
    run_method2 reproduces the golden weights."""
    import main.circuit as mc
    import helpers
    meta, cases, arrays = golden
    c = helpers.build_circuit(cases["paper5"], arrays)
    blob = tmp_path / "paper5_from_mirror.cir"
    with open(blob, "wb") as f:
        pickle.dump(c, f)
        pickle.dump({}, f)                      # the GUI appends its own data after the circuit
    g = mc.Circuit(3)                            # symbolic matrices must travel as plain rows
    h = g.add_gate(mc.H, 2)
    x = g.add(mc.Gate(mc.Matrix.PhaseShift(0.3), "P", True))
    g.add_wires(g, (0, 1), h, ())
    g.add_wire(g, 2, x, 0)
    g.add_wires(h, (), g, (0, 1))
    g.add_wire(x, 0, g, 2)
    blob2 = tmp_path / "h2_from_mirror.cir"
    with open(blob2, "wb") as f:
        pickle.dump(g, f)
    code = textwrap.dedent("""
        import json, pickle, sys
        sys.path.insert(0, %r)
        import main.circuit as rc
        assert rc.__file__.startswith(%r)
        c = pickle.load(open(%r, "rb"))
        assert type(c) is rc.Circuit and c.check()
        w = [c.run_method2(v) for v in ([0]*6, [0,1,0,0,0,0], [1,0,0,0,0,0], [1,1,0,0,0,0])]
        g = pickle.load(open(%r, "rb"))
        assert type(g.gates[0].matrix.rows) is list and len(g.gates[0].matrix.rows) in (2, 4)
        print("RESULT" + json.dumps({"w": w, "g": g.run_method2([0, 0, 1])}))
    """) % (REF, REF, str(blob), str(blob2))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1][6:])
    assert np.array_equal(np.array(out["w"]), arrays["paper5/weights_m2"][:4])
    gw = np.array(out["g"])
    assert abs(gw.sum() - 1.0) < 1e-12 and np.allclose(gw[[1, 3, 5, 7]], 0.25, atol=1e-12)


@pytest.mark.gpu
def test_saved_circuits_run_through_the_engine():
    """.cir files of the reference GUI -> mirror objects -> Circuit.run_method2 on the B200 engine: the weights the
    unmodified reference computed (fixtures), within 1e-12; permutation-only paper circuit bit-exact."""
    ref_w = np.load(os.path.join(SAVED, "weights.npz"))
    c = _load_saved("5")
    for v, w in zip(ref_w["5/inputs"], ref_w["5/weights_m2"]):
        assert np.array_equal(np.array(c.run_method2([int(b) for b in v])), w)
    g = _load_saved("Grover")
    w = np.array(g.run_method2([int(b) for b in ref_w["Grover/inputs"][0]]))
    assert np.max(np.abs(w - ref_w["Grover/weights_m2"][0])) <= 1e-12
    # and back: the circuit the engine just ran pickles to bytes the reference layout accepts (rows as lists)
    blob = pickle.dumps(g)
    g2 = pickle.loads(blob)
    assert type(g2.gates[1].matrix.rows) is list and len(g2.gates) == 25
