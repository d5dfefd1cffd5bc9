"""TEST INFRASTRUCTURE.  Fixtures for the .cir compatibility row: the reference GUI saves a circuit as
a pickled Circuit followed by GUI data (gui/scene.py:73-81).  This script takes the reference's own saved files
(Resources/Saved/5.cir, Grover.cir - data, not source), stores them xz-compressed under tests/golden/saved/ and
records the weights the UNMODIFIED reference computes for the circuits inside them (run_method2, main/circuit.py:279-337).

    python oracle/gen_saved_fixtures.py          # needs the reference checkout at /root/reference
"""
import lzma
import os
import pickle
import sys
import time

import numpy as np

REF = os.environ.get("QSV_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "saved")


def main():
    sys.path.insert(0, REF)                      # the reference's own main package: its classes unpickle the files
    import main.circuit  # noqa: F401
    os.makedirs(OUT, exist_ok=True)
    weights = {}
    for name, inputs in (("5", [[0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0]]),
                         ("Grover", [[0, 0, 0, 0, 0, 0, 1]])):
        raw = open(os.path.join(REF, "Resources", "Saved", name + ".cir"), "rb").read()
        with lzma.open(os.path.join(OUT, name + ".cir.xz"), "wb", preset=9) as f:
            f.write(raw)
        c = pickle.loads(raw)
        t0 = time.time()
        weights[name + "/inputs"] = np.array(inputs, dtype=np.int64)
        weights[name + "/weights_m2"] = np.array([c.run_method2(list(v)) for v in inputs], dtype=np.float64)
        print("%s.cir: %d qubits, %d gates, reference run_method2 %.1f s" % (name, len(c), len(c.gates), time.time() - t0))
    np.savez_compressed(os.path.join(OUT, "weights.npz"), **weights)


if __name__ == "__main__":
    main()
