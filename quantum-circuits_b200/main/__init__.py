"""Drop-in `main` package: main.circuit and main.matrix are the B200-backed mirror; every other
module of the reference's main/ (Grover.py, Shor.py, Shor_example.py, test.py) is picked up
UNCHANGED from a reference checkout when one is present, by extending this package's search path.
Set QSV_REFERENCE_MAIN to the reference's main/ directory, or mount the reference at /root/reference.
"""
import os as _os

for _cand in (_os.environ.get("QSV_REFERENCE_MAIN"), "/root/reference/main"):
    if _cand and _os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)
        break
