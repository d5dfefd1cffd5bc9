"""Host-side mirror of the reference's dense-matrix helper module (main/matrix.py).

Same public names, argument meaning, return values and error texts as the reference
(`Matrix`, `MatrixError`, `tensor`, `matrix_product`; /root/reference/main/matrix.py:6-249), so that
`main/Grover.py`, `main/Shor.py` and the GUI keep working unchanged.  What differs is intent: these
objects only DESCRIBE gates.  The simulation never multiplies them into 2^n x 2^n operators; the
engine (qsv/) lowers each gate to index arithmetic inside CUDA kernels.  Tensor-power factories
(`Matrix.H(30)`, ...) therefore stay symbolic until somebody really asks for `.rows`.

Arithmetic that IS performed here (products, Kronecker products, gate constants) reproduces the
reference's floating-point evaluation order exactly, so matrices built by user scripts are
bit-identical to the ones the reference would build.
"""
import cmath
from itertools import product

# Largest symbolic matrix that may be expanded into a list of lists on request (entries).
MATERIALIZE_LIMIT = 1 << 24


def tensor(matrixlist):
    """Kronecker product of a list (>= 2) of matrices, left to right (reference main/matrix.py:6-11)."""
    acc = matrixlist[0].bintensor(matrixlist[1])
    for nxt in matrixlist[2:]:
        acc = acc.bintensor(nxt)
    return acc


def matrix_product(matrixlist):
    """M_k * ... * M_1 * Id  (reference main/matrix.py:13-18)."""
    acc = Matrix.Id(len(matrixlist[0]))
    for m in matrixlist:
        acc = m * acc
    return acc


class MatrixError(Exception):
    """ An exception class for Matrix """
    pass


class Matrix(object):
    """List-of-lists matrix.  `rows` is readable and writable exactly like the reference's attribute
    (callers assign into it: main/Grover.py:18, main/Shor.py:78-80)."""

    def __init__(self, rows):
        if rows == []:
            raise MatrixError('Please specify the rows of the matrix.')
        width = len(rows[0])
        for r in rows[1:]:
            if len(r) != width:
                raise MatrixError('Row dimensions are not consistent.')
        self.rows = rows

    # -- symbolic construction ---------------------------------------------------------------
    @classmethod
    def _symbolic(cls, spec, dim):
        """A matrix known by structure only.  spec = ('power', base_rows, size) for base^(x)size,
        ('qft', size), or ('rows_fn', callable returning the rows) for the engine's structured gates."""
        m = object.__new__(cls)
        m.__dict__['_spec'] = spec
        m.__dict__['_dim'] = dim
        return m

    @property
    def rows(self):
        d = self.__dict__
        if 'rows' not in d:
            d['rows'] = self._materialize()
        return d['rows']

    @rows.setter
    def rows(self, value):
        self.__dict__['rows'] = value
        self.__dict__.pop('_spec', None)   # once written, the matrix is whatever the caller made it

    def structure(self):
        """('power', base_rows, size) / ('qft', size) / ('kron', A, B) / ('id', size) while the matrix is still
        symbolic, else None."""
        d = self.__dict__
        return d.get('_spec') if 'rows' not in d else None

    def _materialize(self):
        spec = self.__dict__.get('_spec')
        if spec is None:
            raise MatrixError('Please specify the rows of the matrix.')
        dim = self.__dict__['_dim']
        if dim * dim > MATERIALIZE_LIMIT:
            raise MatrixError('Matrix of dimension %d is symbolic and too large to expand.' % dim)
        if spec[0] == 'power':
            _, base, size = spec
            acc = Matrix([list(r) for r in base])
            for _ in range(size - 1):
                acc = acc.bintensor(Matrix([list(r) for r in base]))
            return acc.rows
        if spec[0] == 'qft':
            return _qft_rows(spec[1])
        if spec[0] == 'kron':
            return _kron_rows(spec[1].rows, spec[2].rows)
        if spec[0] == 'id':
            rows = [[0] * dim for _ in range(dim)]
            for i in range(dim):
                rows[i][i] = 1
            return rows
        if spec[0] == 'rows_fn':
            return spec[1]()
        raise MatrixError('unknown symbolic matrix')

    def __getstate__(self):
        # pickles carry plain rows, like the reference's (gui/scene.py:73-81)
        return {'rows': self.rows}

    def __setstate__(self, state):
        self.__dict__.update(state)

    # -- container protocol ------------------------------------------------------------------
    def __len__(self):
        d = self.__dict__
        if 'rows' not in d and '_dim' in d:
            return d['_dim']
        return len(self.rows)

    def __getitem__(self, i):
        return self.rows[i]

    def __setitem__(self, i, item):
        self.rows[i] = item

    def __str__(self):
        return ''.join(''.join(str(v) + ' ' for v in row) + '\n' for row in self.rows)

    def __repr__(self):
        return "Matrix: \"%s\"" % (str(self.rows))

    def __eq__(self, M):
        return (M.rows == self.rows)

    __hash__ = None

    # -- arithmetic (same evaluation order as the reference, main/matrix.py:60-70) -------------
    def __mul__(self, M):
        if len(self.rows[0]) != len(M):
            raise MatrixError('Matrix dimensions must agree.')
        columns = [list(col) for col in zip(*M.rows)]
        out = []
        for row in self.rows:
            out.append([sum([a * b for a, b in zip(row, col)]) for col in columns])
        return Matrix(out)

    def Round(self, places):
        """Real parts rounded to `places` (reference main/matrix.py:73-78: imaginary parts are dropped)."""
        n = len(self)
        out = Matrix.Zero(n)
        for i in range(n):
            src = self.rows[i]
            for j in range(n):
                out[i][j] = round(src[j].real, places)
        return out

    def transpose(self):
        self.rows = [list(col) for col in zip(*self.rows)]

    def getConjugate(self):
        """Conjugate transpose (reference main/matrix.py:83-88)."""
        n = len(self)
        out = Matrix.Zero(n)
        for i in range(n):
            src = self.rows[i]
            for j in range(len(out.rows[0])):
                out[j][i] = (src[j]).conjugate()
        return out

    def isUnitary(self):
        """(self * self^dagger).Round(12) == Id  (reference main/matrix.py:90-94; Round keeps real parts only).
        The reference evaluates the O(n^3) product entry by entry in Python - 80 % of its run time at 8 qubits
        (SURVEY 8a-4).  Numeric matrices take the same test through one BLAS product; the entries differ from the
        reference's sums by rounding (1e-16), far inside the 12-decimal comparison."""
        spec = self.structure()
        if spec is not None and spec[0] in ('kron', 'power', 'qft', 'id'):
            # decided on the structure: a Kronecker product / tensor power is unitary iff its factors are (up to
            # scalars, which the reference's test would reject in either factor just the same for |c| != 1)
            if spec[0] == 'kron':
                return spec[1].isUnitary() and spec[2].isUnitary()
            if spec[0] == 'power':
                return Matrix([list(r) for r in spec[1]]).isUnitary()
            return True
        fast = _is_unitary_fast(self.rows)
        if fast is not None:
            return fast
        return (self * (self.getConjugate())).Round(12) == Matrix.Id(len(self))

    def bintensor(self, M):
        """self (x) M  (reference main/matrix.py:98-112); entry = self[a][b] * M[k][l].
        A product with a symbolic factor, or one of 2^12 rows and more, stays symbolic (('kron', A, B)): the lowering
        takes it factor by factor and nobody pays for 4^q entries (`tensor([Matrix.H(n), Matrix.Id(2)])` of
        main/Grover.py:42-44 at n = 20 would be 4 * 10^12 of them).  Reading or writing `.rows` expands it, with
        the reference's entry order and arithmetic."""
        if isinstance(M, Matrix) and (self.structure() is not None or M.structure() is not None
                                      or len(self) * len(M) >= LAZY_KRON_DIM) \
                and _is_square(self) and _is_square(M):
            return Matrix._symbolic(('kron', self, M), len(self) * len(M))
        return Matrix(_kron_rows(self.rows, M.rows))

    def signature(self):
        """Hashable description of the matrix without expanding symbolic ones (the lowering cache keys on it)."""
        d = self.__dict__
        if 'rows' in d:
            return tuple(map(tuple, d['rows']))
        spec = d.get('_spec')
        if spec[0] == 'kron':
            return ('kron', spec[1].signature(), spec[2].signature())
        if spec[0] == 'power':
            return ('power', tuple(map(tuple, spec[1])), spec[2])
        if spec[0] == 'rows_fn':
            raise TypeError('structured gate: not describable by content')
        return tuple(spec)

    def apply(self, matrices):
        if len(self.rows[0]) > 1:
            raise MatrixError("This method is reserved for vectors.")
        v = Matrix.vector([self[i][0] for i in range(len(self))])
        for N in matrices:
            v = N * v
        return v

    # -- factories (reference main/matrix.py:131-249) ---------------------------------------------
    @classmethod
    def vector(cls, column):
        return cls([[v] for v in column])

    @classmethod
    def Zero(cls, nrows, ncolumns=0):
        if ncolumns == 0:
            ncolumns = nrows
        return cls([[0] * ncolumns for _ in range(nrows)])

    @classmethod
    def Id(cls, size):
        if size >= LAZY_KRON_DIM:
            return cls._symbolic(('id', size), size)      # expanded when somebody reads or writes .rows
        rows = [[0] * size for _ in range(size)]
        for i in range(size):
            rows[i][i] = 1
        return cls(rows)

    @classmethod
    def _power(cls, base, size):
        if size == 1:
            return cls([list(r) for r in base])
        return cls._symbolic(('power', [list(r) for r in base], size), 2 ** size)

    @classmethod
    def H(cls, size=1):
        s = 2 ** (-0.5)
        return cls._power([[s, s], [s, -s]], size)

    @classmethod
    def X(cls, size=1):
        return cls._power([[0, 1], [1, 0]], size)

    @classmethod
    def Y(cls, size=1):
        return cls._power([[0, -1j], [0 + 1j, 0]], size)

    @classmethod
    def Z(cls, size=1):
        return cls._power([[1, 0], [0, -1]], size)

    @classmethod
    def SqrtNot(cls, size=1):
        return cls._power([[0.5 * (1 + 1j), 0.5 * (1 - 1j)], [0.5 * (1 - 1j), 0.5 * (1 + 1j)]], size)

    @classmethod
    def PhaseShift(cls, phase, size=1):
        return cls._power([[1, 0], [0, cmath.e ** (1j * (phase))]], size)

    @classmethod
    def QFT(cls, size=1):
        if size <= 5:
            return cls(_qft_rows(size))
        return cls._symbolic(('qft', size), 2 ** size)

    @classmethod
    def Permutation(cls, permutation):
        """Qubit-reordering matrix: P[y][z] = 1 where bit i of z is bit permutation[i] of y
        (reference main/matrix.py:222-238).  The engine never builds these; kept for API parity."""
        n = len(permutation)
        P = cls.Zero(2 ** n)
        for bits in product(range(2), repeat=n):
            y = 0
            z = 0
            for i in range(n):
                y = (y << 1) | bits[i]
                z = (z << 1) | bits[permutation[i]]
            P[y][z] = 1
        return P

    @classmethod
    def Cnot(cls):
        return cls([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])

    @classmethod
    def T(cls):
        """Toffoli (the reference calls it T, main/matrix.py:245-249)."""
        m = cls.Id(8)
        m.rows[6] = [0, 0, 0, 0, 0, 0, 0, 1]
        m.rows[7] = [0, 0, 0, 0, 0, 0, 1, 0]
        return m


LAZY_KRON_DIM = 1 << 12      # Kronecker products of this many rows and more are kept as a pair of factors


def _is_square(m):
    d = m.__dict__
    if 'rows' not in d:
        return True                  # symbolic matrices are square by construction
    rows = d['rows']
    return len(rows) > 0 and all(len(r) == len(rows) for r in rows)


def _kron_rows(arows, mrows):
    out = []
    for arow in arows:
        for mrow in mrows:
            line = []
            for a in arow:
                line.extend([a * m for m in mrow])
            out.append(line)
    return out


def _is_unitary_fast(rows):
    """numpy form of the unitarity test, or None when the rows are not a plain numeric square matrix."""
    try:
        import numpy as np
        a = np.array(rows, dtype=np.complex128)
    except (TypeError, ValueError):
        return None
    if a.ndim != 2 or a.shape[0] != a.shape[1] or a.shape[0] < 16:
        return None                              # small matrices: the reference's own arithmetic is instant
    p = a @ a.conj().T
    return bool(np.array_equal(np.round(p.real, 12), np.eye(a.shape[0])))


def _qft_rows(size):
    """n^-1/2 * e^(+2 pi i * (i*j) / n), i*j NOT reduced mod n — the reference's exact expression
    (main/matrix.py:212-219), so the constants handed to the kernels are the reference's doubles."""
    n = 2 ** size
    rows = [[0] * n for _ in range(n)]
    for i in range(n):
        for j in range(n):
            rows[i][j] = (n ** (-0.5)) * (cmath.e ** ((i * j) * 1j * 2 * cmath.pi / n))
    return rows
