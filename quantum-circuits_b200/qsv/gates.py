"""Structured gates the reference's dense-matrix API cannot express at 30+ qubits.

The reference describes the Grover oracle, the diffusion operator and Shor's modular exponentiation
as 2^n x 2^n matrices (main/Grover.py:14-31, main/Shor.py:54-83).  These classes are Gate objects
(usable with Circuit.add / add_wires like any other) whose matrix stays symbolic and whose lowering
goes straight to the dedicated kernels.  They live outside main.circuit on purpose: the GUI builds its
palette from every Gate subclass found in that module (gui/main_window.py:269-273).
"""
import cmath

import numpy as np

from main.circuit import Gate
from main.matrix import Matrix
from .schedule import Op


def _symbolic(dim, rows_fn):
    return Matrix._symbolic(("rows_fn", rows_fn), dim)


class Phase(Gate):
    """diag(1, e^{i theta}) — Matrix.PhaseShift (main/matrix.py:202-209) as a gate class."""

    def __init__(self, theta, name="P"):
        super().__init__(Matrix.PhaseShift(theta), name, True)


class T8(Phase):
    """pi/8 gate (the reference's class T is the Toffoli gate)."""

    def __init__(self, name="T8"):
        super().__init__(cmath.pi / 4, name)


class CZ(Gate):
    SIZE = 2

    def __init__(self, name="CZ"):
        super().__init__(Matrix([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, -1]]), name, True)


class CPhase(Gate):
    """Controlled phase diag(1,1,1,e^{i theta})."""
    SIZE = 2

    def __init__(self, theta, name="CP"):
        super().__init__(Matrix([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, cmath.e ** (1j * theta)]]),
                         name, True)


class GroverOracle(Gate):
    """F of main/Grover.py:14-22 for a search register of n_search qubits (ports 0..n_search-1, port 0
    most significant) and one ancilla (last port): X on the ancilla iff the register equals x."""

    def __init__(self, n_search, x, name="f"):
        self.n_search, self.x = int(n_search), int(x)
        if not 0 <= self.x < (1 << self.n_search):
            raise RuntimeError("Marked item does not fit the search register.")
        dim = 1 << (self.n_search + 1)

        def rows():
            r = Matrix.Id(dim).rows
            a, b = 2 * self.x, 2 * self.x + 1
            r[a][a] = r[b][b] = 0
            r[a][b] = r[b][a] = 1
            return r
        super().__init__(_symbolic(dim, rows), name, True)

    def _qsv_lower(self, wires, layout, transposed):
        n = self.n_search
        return [Op("oracle", (layout.pos[wires[n]],), tuple(layout.pos[w] for w in wires[:n]),
                   cvals=tuple((self.x >> (n - 1 - p)) & 1 for p in range(n)), name=self.name)]


class GroverDiffusion(Gate):
    """H^n U0 H^n = 2|s><s| - I on n_search qubits (main/Grover.py:24-31,40-44, without the (x) Id on the
    ancilla: wire it to the search register only)."""

    def __init__(self, n_search, name="d"):
        self.n_search = int(n_search)
        dim = 1 << self.n_search

        def rows():
            v = 2.0 / dim
            return [[v - (1 if i == j else 0) for j in range(dim)] for i in range(dim)]
        super().__init__(_symbolic(dim, rows), name, True)

    def _qsv_lower(self, wires, layout, transposed):
        n_total = layout.n
        pos = sorted(layout.pos[w] for w in wires)
        anc_bits = n_total - self.n_search
        if pos == list(range(anc_bits, n_total)) and anc_bits <= 3:
            return [Op("diffusion", tuple(pos), (), params={"anc_bits": anc_bits, "n_search": self.n_search},
                       name=self.name)]
        # general placement: H, reflection about |0..0>, H
        s = 2 ** (-0.5)
        Hm = np.array([[s, s], [s, -s]], dtype=np.complex128)
        hs = [Op("dense", (layout.pos[w],), (), matrix=Hm, name=self.name) for w in wires]
        allp = tuple(layout.pos[w] for w in wires)
        refl = [Op("phase", (), allp, cvals=(0,) * len(allp), scalar=-1 + 0j, name=self.name),
                Op("phase", (), (), scalar=-1 + 0j, name=self.name)]
        return hs + refl + hs


class ModExp(Gate):
    """Shor's modular-exponentiation unitary U of main/Shor.py:54-83 for modulus N and base a, on
    q = 2*bitlen(N) qubits: ports 0..q/2-1 hold x (port 0 most significant), the rest hold y."""

    def __init__(self, a, N, name="U"):
        self.a, self.N = int(a), int(N)
        self.q = (len(bin(self.N)) - 2) * 2
        dim = 1 << self.q

        def rows():
            h = self.q // 2
            ones = (1 << h) - 1
            r = Matrix.Zero(dim).rows
            for i in range(dim):
                x, y = i >> h, i & ones
                f = pow(self.a, x, self.N)
                j = i
                if f != 0:
                    if y == 0:
                        j = (x << h) | f
                    elif y == f:
                        j = x << h
                    elif f != ones and y == ones:
                        j = (x << h) | (ones & ~f)
                    elif f != ones and y == (ones & ~f):
                        j = (x << h) | ones
                r[j][i] = 1
            return r
        super().__init__(_symbolic(dim, rows), name, True)

    def _qsv_lower(self, wires, layout, transposed):
        h = self.q // 2
        return [Op("modexp", tuple(layout.pos[w] for w in wires[h:]), tuple(layout.pos[w] for w in wires[:h]),
                   params={"a": self.a, "N": self.N}, name=self.name)]
