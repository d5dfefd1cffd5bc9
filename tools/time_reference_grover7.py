"""Time the UNMODIFIED reference on its own Grover script (main/Grover.py:10-48 builds the 7-qubit circuit at import,
then Circuit.run_method2, main/circuit.py:279-337) and write profiles/r02_reference_grover7.json.

Runs where the reference checkout is (QSV_REFERENCE_ROOT, default /root/reference) - i.e. in the build container, on
its CPU; the GPU box has no checkout, so bench.py reports this committed number next to the GPU arm's time for the same
circuit (`reference_grover7`).  Nothing of this repo is on the reference's path: sys.path holds the checkout only."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("QSV_REFERENCE_ROOT", "/root/reference")

CODE = r"""
import json, sys, time
sys.path.insert(0, %r)
t0 = time.perf_counter()
import main.Grover as G                      # builds F, U0, H^n (x) I and the 25-gate circuit
t1 = time.perf_counter()
assert G.__file__.startswith(%r) and G.c.__class__.__module__ == "main.circuit"
import main.circuit as rc
assert rc.__file__.startswith(%r)
in_v = [0] * G.n + [1]
w = G.c.run_method2(in_v)
t2 = time.perf_counter()
best = max(range(len(w)), key=lambda i: w[i])
print("RESULT" + json.dumps({"build_s": t1 - t0, "run_method2_s": t2 - t1, "qubits": len(G.c), "gates": len(G.c.gates),
                             "argmax": best, "p_marked_pair": w[2 * G.x] + w[2 * G.x + 1]}))
""" % (REF, REF, REF)


def main():
    if not os.path.isdir(os.path.join(REF, "main")):
        raise SystemExit("no reference checkout at %s" % REF)
    runs = []
    for _ in range(3):
        r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, timeout=1200)
        if r.returncode != 0:
            raise SystemExit(r.stderr[-2000:])
        runs.append(json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1][6:]))
    best = min(runs, key=lambda d: d["run_method2_s"])
    out = {"what": "unmodified reference (MatejVitek/Quantum-Circuits main/Grover.py + Circuit.run_method2), CPython %s, "
                   "one core of the build container (no GPU involved)" % sys.version.split()[0],
           "qubits": best["qubits"], "gates": best["gates"], "build_s": best["build_s"],
           "run_method2_s": best["run_method2_s"], "runs_run_method2_s": [d["run_method2_s"] for d in runs],
           "gates_per_sec": best["gates"] / best["run_method2_s"], "argmax": best["argmax"],
           "p_marked_pair": best["p_marked_pair"], "host": os.uname().nodename, "date": time.strftime("%Y-%m-%d")}
    path = os.path.join(ROOT, "profiles", "r02_reference_grover7.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
