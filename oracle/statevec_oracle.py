"""CPU oracle for the state-vector hot path.  TEST INFRASTRUCTURE ONLY.

A numpy restatement of what the reference (MatejVitek/Quantum-Circuits) computes on the path
Circuit.run / Circuit.run_method2 (main/circuit.py:279-366) and of the operators its algorithm
scripts feed that path (main/Grover.py:14-31, main/Shor.py:54-83, main/matrix.py:152-249).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker.  The product (quantum-circuits_b200/) never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against vectors
produced by the reference itself (oracle/gen_golden.py imports /root/reference and writes
tests/golden/*.json|npz): weights, amplitudes (replayed with the reference's own Matrix ops),
sampled outcome sequences under fixed seeds, the modexp matrices and the Matrix factories.

Each function cites the reference lines it follows.  State layout: numpy complex128 vector of 2^n
amplitudes, wire 0 = most-significant index bit (main/circuit.py:290-299).
"""
import bisect
import cmath
import itertools
import numpy as np


# ---------------------------------------------------------------------------------------------
# circuit graph -> (matrix, target wires) list        (host logic of the reference, restated)
# ---------------------------------------------------------------------------------------------

def startqbit(circuit, gate_index, port):
    """Input wire whose qubit reaches `port` of gate `gate_index` — main/circuit.py:380-389."""
    gate = circuit.gates[gate_index]
    j = port
    wire = gate.in_wires[j]
    while wire.left is not circuit:
        gate = wire.left
        j = wire.lind
        wire = gate.in_wires[j]
    return wire.lind


def startqbits(circuit):
    """For every output wire, the input wire whose qubit it shows — main/circuit.py:391-403."""
    out = []
    for i in range(len(circuit)):
        wire = circuit.output[i]
        while wire.left is not circuit:
            wire = wire.left.in_wires[wire.lind]
        out.append(wire.lind)
    return out


def gate_matrix(gate):
    """gate.matrix as a complex128 array (Gate.get_matrix, main/circuit.py:504-511)."""
    return np.array(gate.matrix.rows, dtype=np.complex128)


def trace(circuit):
    """Sorted circuit -> ([(matrix, [target wires])...], startqbits).  main/circuit.py:280,303-305."""
    circuit.sort()
    steps = []
    for gi, gate in enumerate(circuit.gates):
        steps.append((gate_matrix(gate), [startqbit(circuit, gi, p) for p in range(len(gate))]))
    return steps, startqbits(circuit)


# ---------------------------------------------------------------------------------------------
# state evolution
# ---------------------------------------------------------------------------------------------

def initial_state(in_v):
    """Tensor product of e0/e1 columns — main/circuit.py:290-301, main/matrix.py:6-11."""
    n = len(in_v)
    idx = 0
    for b in in_v:
        idx = (idx << 1) | int(b)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[idx] = 1.0
    return psi


def bring_to_front_permutation(n, targets):
    """The qubit order used for one gate: targets first, the displaced qubits swapped into the holes —
    main/circuit.py:306-313."""
    perm = list(range(n))
    for i, y in enumerate(targets):
        x = perm[i]
        if x != y:
            j = perm.index(y)
            perm[i], perm[j] = y, x
    return perm


def apply_gate(psi, n, targets, M):
    """psi <- P1 * (M (x) Id) * P2 * psi — main/circuit.py:315-327 with Matrix.Permutation
    (main/matrix.py:222-238): after P2 the qubit at position i is perm[i]; M acts on the first k
    positions with port 0 as its most-significant bit; P1 restores input-wire order."""
    k = len(targets)
    perm = bring_to_front_permutation(n, targets)
    t = psi.reshape((2,) * n)
    t = np.transpose(t, perm)                      # P2 * psi: new axis i carries qubit perm[i]
    t = t.reshape(1 << k, -1)
    t = np.asarray(M, dtype=np.complex128) @ t      # (M (x) Id) * psi
    t = t.reshape((2,) * n)
    t = np.transpose(t, np.argsort(perm))          # P1 * psi
    return np.ascontiguousarray(t).reshape(-1)


def output_relabel(psi, n, sq):
    """psi_out[z] = psi[y] with bit i of z = bit sq[i] of y — main/circuit.py:329-332 (P1^T * psi)."""
    t = psi.reshape((2,) * n)
    return np.ascontiguousarray(np.transpose(t, sq)).reshape(-1)


def evolve(steps, sq, in_v, transposed=False):
    """Final amplitudes in output order.  transposed=True applies M^T for every gate, which is what
    the path sum of run_method1 computes (it indexes gate.matrix[in][out], main/circuit.py:264-269)."""
    n = len(in_v)
    psi = initial_state(in_v)
    for M, targets in steps:
        psi = apply_gate(psi, n, targets, M.T if transposed else M)
    return output_relabel(psi, n, sq)


def weights(psi):
    """|psi_x|^2 as abs(z)**2 — main/circuit.py:334-337."""
    return np.abs(psi) ** 2


def run_method2(circuit, in_v):
    """Circuit.run_method2 — main/circuit.py:279-337."""
    validate_input(circuit, in_v)
    steps, sq = trace(circuit)
    return weights(evolve(steps, sq, in_v))


def run_method1(circuit, in_v):
    """Weights of Circuit.run_method1 (main/circuit.py:221-276) obtained from the transposed gates."""
    validate_input(circuit, in_v)
    steps, sq = trace(circuit)
    return weights(evolve(steps, sq, in_v, transposed=True))


def run_method1_path_sum(circuit, in_v):
    """Circuit.run_method1 as the reference writes it (main/circuit.py:221-276), loop for loop: every output
    assignment, every assignment of the internal wires, the product over circuit.gates of matrix[i][j] with i / j read
    from the in- / out-wire values (first wire most significant), |Sum|^2.  Pure Python: small circuits only."""
    validate_input(circuit, in_v)
    n = len(circuit)
    value = {}
    for i in range(n):
        value[id(circuit.input[i])] = in_v[i]
    int_wires = circuit.get_internal_wires()
    mats = [gate_matrix(g) for g in circuit.gates]
    w = np.zeros(1 << n)
    for index, out in enumerate(itertools.product(range(2), repeat=n)):
        clash = False
        for i in range(n):                                   # a wire from an input straight to output i
            if circuit.output[i].left is circuit and value[id(circuit.output[i])] != out[i]:
                clash = True
                break
            value[id(circuit.output[i])] = out[i]
        if clash:
            for i in range(n):                               # (the reference keeps input wires' values: restore)
                value[id(circuit.input[i])] = in_v[i]
            continue
        total = 0
        for assign in itertools.product(range(2), repeat=len(int_wires)):
            for wire, v in zip(int_wires, assign):
                value[id(wire)] = v
            prod = 1
            for g, M in zip(circuit.gates, mats):
                i = j = 0
                for wire in g.in_wires:
                    i = (i << 1) | value[id(wire)]
                for wire in g.out_wires:
                    j = (j << 1) | value[id(wire)]
                if M[i][j] == 0:
                    prod = 0
                    break
                prod = prod * M[i][j]
            total += prod
        w[index] = abs(total) ** 2
    return w


def validate_input(circuit, in_v):
    """main/circuit.py:282-288."""
    if len(in_v) != len(circuit):
        raise RuntimeError("Input length does not match the size of the circuit.")
    for i in in_v:
        if not (i == 0 or i == 1):
            raise RuntimeError("Invalid input - values must be 0 or 1.")


# ---------------------------------------------------------------------------------------------
# sampling
# ---------------------------------------------------------------------------------------------

def sample_index(w, u):
    """Index chosen by random.choices(states, weights) for the uniform draw u — main/circuit.py:360-364
    with CPython 3.12 random.choices: sequential accumulate, total = cum[-1],
    bisect_right(cum, u*total, 0, N-1)."""
    cum = list(itertools.accumulate(float(x) for x in w))
    total = cum[-1] + 0.0
    if total <= 0.0:
        raise ValueError("Total of weights must be greater than zero")
    return bisect.bisect_right(cum, u * total, 0, len(cum) - 1)


def sample_index_np(w, u):
    """Same rule as sample_index with numpy primitives (np.cumsum adds strictly left to right, so the
    prefix sums are bit-identical to itertools.accumulate); for registers too large for Python loops."""
    cum = np.cumsum(np.asarray(w, dtype=np.float64))
    total = float(cum[-1])
    if total <= 0.0:
        raise ValueError("Total of weights must be greater than zero")
    return int(np.searchsorted(cum[:-1], u * total, side="right"))


def index_bits(index, n):
    """Outcome as the bit list run() returns (wire 0 first) — main/circuit.py:360-364."""
    return [(index >> (n - 1 - i)) & 1 for i in range(n)]


def choose_method(n_gates, n_qubits, n_internal_wires):
    """Cost model of Circuit.run — main/circuit.py:343-351.  Returns 1 or 2."""
    n1 = n_gates * (2 ** (n_qubits + n_internal_wires))
    n2 = (3 * n_gates) * (2 ** (2 * n_qubits))
    return 1 if n1 <= n2 else 2


# ---------------------------------------------------------------------------------------------
# gate constants                                              main/matrix.py:152-249
# ---------------------------------------------------------------------------------------------

def mat_H():
    s = 2 ** (-0.5)
    return np.array([[s, s], [s, -s]], dtype=np.complex128)


def mat_X():
    return np.array([[0, 1], [1, 0]], dtype=np.complex128)


def mat_Y():
    return np.array([[0, -1j], [1j, 0]], dtype=np.complex128)


def mat_Z():
    return np.array([[1, 0], [0, -1]], dtype=np.complex128)


def mat_SqrtNot():
    return np.array([[0.5 * (1 + 1j), 0.5 * (1 - 1j)], [0.5 * (1 - 1j), 0.5 * (1 + 1j)]], dtype=np.complex128)


def mat_PhaseShift(phase):
    return np.array([[1, 0], [0, cmath.e ** (1j * phase)]], dtype=np.complex128)


def mat_QFT(size):
    """main/matrix.py:212-219: n^-1/2 e^{+2 pi i (i j)/n}, exponent not reduced mod n."""
    n = 2 ** size
    M = np.empty((n, n), dtype=np.complex128)
    for i in range(n):
        for j in range(n):
            M[i, j] = (n ** (-0.5)) * (cmath.e ** ((i * j) * 1j * 2 * cmath.pi / n))
    return M


def mat_Cnot():
    return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)


def mat_Toffoli():
    M = np.eye(8, dtype=np.complex128)
    M[6, 6] = M[7, 7] = 0
    M[6, 7] = M[7, 6] = 1
    return M


def kron_power(M, size):
    """Left-fold Kronecker power like tensor() — main/matrix.py:6-11,98-112."""
    out = M
    for _ in range(size - 1):
        out = np.kron(out, M)
    return out


# ---------------------------------------------------------------------------------------------
# Grover operators                                             main/Grover.py:14-31
# ---------------------------------------------------------------------------------------------

def grover_oracle_matrix(n_search, x):
    """F: identity except 2x <-> 2x+1 — main/Grover.py:14-22."""
    D = 2 ** (n_search + 1)
    F = np.eye(D, dtype=np.complex128)
    a, b = 2 * x, 2 * x + 1
    F[a, a] = F[b, b] = 0
    F[a, b] = F[b, a] = 1
    return F


def grover_u0_matrix(n_search):
    """U0 (x) Id(2), U0 = diag(+1, -1, ..., -1) — main/Grover.py:24-31."""
    d = -np.ones(2 ** n_search)
    d[0] = 1
    return np.kron(np.diag(d), np.eye(2)).astype(np.complex128)


def grover_steps(n_search, x, iterations=None):
    """The gate list of main/Grover.py:33-48 as (matrix, targets) steps on n_search+1 wires."""
    import math
    n = n_search + 1
    allw = list(range(n))
    if iterations is None:
        iterations = int((math.pi / 4) * math.sqrt(2 ** n_search))
    Hn = kron_power(mat_H(), n)
    HI = np.kron(kron_power(mat_H(), n_search), np.eye(2))
    steps = [(Hn, allw)]
    for _ in range(iterations):
        steps += [(grover_oracle_matrix(n_search, x), allw), (HI, allw), (grover_u0_matrix(n_search), allw), (HI, allw)]
    return steps


def grover_iteration_closed_form(psi, n_search, x):
    """One oracle + diffusion step without building matrices (for registers too large for
    grover_steps): swap 2x<->2x+1, then psi[i,a] <- 2*mean_a - psi[i,a] (H U0 H = 2|s><s| - I)."""
    out = psi.copy().reshape(-1, 2)
    out[x, 0], out[x, 1] = psi.reshape(-1, 2)[x, 1], psi.reshape(-1, 2)[x, 0]
    mean = out.sum(axis=0) / (2 ** n_search)
    return (2 * mean[None, :] - out).reshape(-1)


# ---------------------------------------------------------------------------------------------
# Shor modular exponentiation                                  main/Shor.py:54-83
# ---------------------------------------------------------------------------------------------

def modexp_partner(i, q, a, N):
    """Image of basis index i under U (an involution), closed form of the reference's loops:
    i = (x << h) | y, h = q/2, f = a^x mod N; (x,0)<->(x,f), (x,2^h-1)<->(x,~f), rest fixed."""
    h = q // 2
    ones = (1 << h) - 1
    x, y = i >> h, i & ones
    f = pow(a, x, N)
    nf = ones & ~f
    if f != 0:
        if y == 0:
            return (x << h) | f
        if y == f:
            return (x << h)
        if f != ones:
            if y == ones:
                return (x << h) | nf
            if y == nf:
                return (x << h) | ones
    return i


def modexp_permutation(q, a, N):
    return np.array([modexp_partner(i, q, a, N) for i in range(1 << q)], dtype=np.int64)


def modexp_matrix_reference_loops(q, a, N):
    """The 2^q x 2^q 0/1 matrix exactly as the reference's double loop builds it (main/Shor.py:54-83,
    same as main/Shor_example.py:11-40), restated with integer bit operations.  O(4^q): small q only."""
    h = q // 2
    ones = (1 << h) - 1
    D = 1 << q
    U = np.zeros((D, D), dtype=np.int64)
    for i in range(D):
        i1, i2 = i >> h, i & ones
        f = pow(a, i1, N)
        for j in range(D):
            j1, j2 = j >> h, j & ones
            if i2 == 0 and i1 == j1 and f == j2:
                U[i, j] = 1
                U[j, i] = 1
                break
            elif i2 == ones and i1 == j1:
                if (f ^ j2) & ones == ones and f <= ones:       # every bit of j2 differs from f
                    U[i, j] = 1
                    U[j, i] = 1
                    break
    for r in range(D):
        if not U[r].any():
            U[r, r] = 1
    return U


def shor_steps(N, a):
    """Gate list of main/Shor.py:85-94 as (matrix, targets) steps: H^(q/2) on the first register,
    U on everything, QFT(q/2) on the first register."""
    q = (len(bin(N)) - 2) * 2
    h = q // 2
    perm = modexp_permutation(q, a, N)
    U = np.zeros((1 << q, 1 << q), dtype=np.complex128)
    U[perm, np.arange(1 << q)] = 1
    first = list(range(h))
    return [(kron_power(mat_H(), h), first), (U, list(range(q))), (mat_QFT(h), first)], q


# ---------------------------------------------------------------------------------------------
# synthetic workloads
# ---------------------------------------------------------------------------------------------

def mat_CZ():
    return np.diag([1, 1, 1, -1]).astype(np.complex128)


def random_layered_steps(n, depth, seed=20261019):
    """Config-3 generator: per layer one 1-qubit gate on every wire from {H, X, T8=phase(pi/4),
    P=phase(uniform)}, then a random perfect matching of wires with CNot or CZ on each pair."""
    import math
    import random
    rng = random.Random(seed)
    steps = []
    for _ in range(depth):
        for w in range(n):
            c = rng.randrange(4)
            if c == 0:
                steps.append((mat_H(), [w], "H"))
            elif c == 1:
                steps.append((mat_X(), [w], "X"))
            elif c == 2:
                steps.append((mat_PhaseShift(math.pi / 4), [w], "T8"))
            else:
                steps.append((mat_PhaseShift(rng.uniform(0, 2 * math.pi)), [w], "P"))
        order = list(range(n))
        rng.shuffle(order)
        for i in range(0, n - 1, 2):
            a, b = order[i], order[i + 1]
            if rng.randrange(2) == 0:
                steps.append((mat_Cnot(), [a, b], "CNot"))
            else:
                steps.append((mat_CZ(), [a, b], "CZ"))
    return steps


def qft_ladder_steps(n):
    """Config-4 gate list: H + controlled-phase ladder; equals mat_QFT(n) followed by a bit reversal
    of the wires (the engine folds that reversal into its output map)."""
    import math
    steps = []
    for i in range(n):
        steps.append((mat_H(), [i], "H"))
        for j in range(i + 1, n):
            theta = 2 * math.pi / (1 << (j - i + 1))
            steps.append((np.diag([1, 1, 1, cmath.e ** (1j * theta)]).astype(np.complex128), [j, i], "CP"))
    return steps
