"""Workloads of BASELINE.json's configs, built through the (mirror of the) reference's circuit API.

Grover and Shor restate what the reference's algorithm scripts do (main/Grover.py, main/Shor.py) so
that the same circuits exist on the GPU box, where the reference checkout is absent; the random
layered circuit and the QFT ladder are the synthetic configs 3 and 4 of SURVEY §8d.
"""
import fractions
import math
import random
from math import gcd

import main.circuit as mc
from main.matrix import Matrix, tensor
from . import gates as xg


# ---- config 3: random H / X / T8 / phase layers + CNot / CZ matchings ----------------------------

def _layered_specs(n, depth, seed):
    """Gate list of the config-3 generator: [(kind, wires, angle)], kind in H / X / T8 / P / CNot / CZ."""
    rng = random.Random(seed)
    specs = []
    for _ in range(depth):
        for w in range(n):
            r = rng.randrange(4)
            if r == 0:
                specs.append(("H", (w,), None))
            elif r == 1:
                specs.append(("X", (w,), None))
            elif r == 2:
                specs.append(("T8", (w,), math.pi / 4))
            else:
                specs.append(("P", (w,), rng.uniform(0, 2 * math.pi)))
        order = list(range(n))
        rng.shuffle(order)
        for i in range(0, n - 1, 2):
            specs.append(("CNot" if rng.randrange(2) == 0 else "CZ", (order[i], order[i + 1]), None))
    return specs


def _circuit_from_specs(n, specs):
    c = mc.Circuit(n)
    last = [(c, i) for i in range(n)]
    for k, (kind, wires, angle) in enumerate(specs):
        name = "%s_%d" % (kind, k)
        if kind == "H":
            g = mc.H(1, name)
        elif kind == "X":
            g = mc.X(1, name)
        elif kind in ("T8", "P"):
            g = xg.Phase(angle, name)
        elif kind == "CNot":
            g = mc.CNot(name)
        else:
            g = xg.CZ(name)
        c.add(g)
        for p, w in enumerate(wires):
            c.add_wire(last[w][0], last[w][1], g, p)
            last[w] = (g, p)
    for w in range(n):
        c.add_wire(last[w][0], last[w][1], c, w)
    return c


def random_layered_circuit(n, depth, seed=20261019):
    """Per layer: one 1-qubit gate on every wire drawn from {H, X, phase(pi/4), phase(uniform)}, then a
    random perfect matching of the wires with CNot or CZ on each pair: n + n//2 gates per layer."""
    return _circuit_from_specs(n, _layered_specs(n, depth, seed))


def random_layered_roundtrip_circuit(n, depth, seed=20261019):
    """The config-3 circuit followed by its inverse (gates in reverse order, phases conjugated; H, X, CNot
    and CZ are their own inverses): maps every basis state to itself.  A size-independent correctness
    check for registers far beyond the oracle's reach (bench.py `check`, 30-36 qubits)."""
    fwd = _layered_specs(n, depth, seed)
    inv = [(kind, wires, None if angle is None else -angle) for kind, wires, angle in reversed(fwd)]
    return _circuit_from_specs(n, fwd + inv)


# ---- config 4: QFT as H + controlled-phase ladder ------------------------------------------------

def qft_ladder_circuit(n):
    """n H gates and n(n-1)/2 controlled phases (sign +i like Matrix.QFT, main/matrix.py:212-219).  The
    closing bit reversal is expressed in the output wiring: output wire i shows qubit n-1-i."""
    c = mc.Circuit(n)
    last = [(c, i) for i in range(n)]
    k = 0
    for i in range(n):
        g = c.add(mc.H(1, "H_%d" % i))
        c.add_wire(last[i][0], last[i][1], g, 0)
        last[i] = (g, 0)
        for j in range(i + 1, n):
            g = c.add(xg.CPhase(2 * math.pi / (1 << (j - i + 1)), "CP_%d" % k))
            k += 1
            c.add_wire(last[j][0], last[j][1], g, 0)
            c.add_wire(last[i][0], last[i][1], g, 1)
            last[j], last[i] = (g, 0), (g, 1)
    for i in range(n):
        c.add_wire(last[n - 1 - i][0], last[n - 1 - i][1], c, i)
    return c


# ---- Grover ---------------------------------------------------------------------------------------

def grover_circuit_dense(n_search=6, x=13, iterations=None):
    """The circuit of main/Grover.py:10-48, dense full-width operators included (small registers)."""
    n = n_search
    c = mc.Circuit(n + 1)
    dim = 2 ** (n + 1)
    F = Matrix.Id(dim)
    F[2 * x][2 * x] = F[2 * x + 1][2 * x + 1] = 0
    F[2 * x][2 * x + 1] = F[2 * x + 1][2 * x] = 1
    U = Matrix.Zero(2 ** n)
    for i in range(2 ** n):
        U[i][i] = 1 if i == 0 else -1
    U = tensor([U, Matrix.Id(2)])
    h1 = c.add_gate(mc.H, n + 1)
    c.add_wires(c, (), h1, ())
    times = int((math.pi / 4) * math.sqrt(2 ** n)) if iterations is None else iterations
    for t in range(times):
        u1 = c.add(mc.Gate(F, 'f' + str(t), True))
        c.add_wires(h1, (), u1, ())
        h2 = c.add(mc.Gate(tensor([Matrix.H(n), Matrix.Id(2)]), 'h1' + str(t), True))
        c.add_wires(u1, (), h2, ())
        u2 = c.add(mc.Gate(U, 'u0' + str(t), True))
        c.add_wires(h2, (), u2, ())
        h3 = c.add(mc.Gate(tensor([Matrix.H(n), Matrix.Id(2)]), 'h2' + str(t), True))
        c.add_wires(u2, (), h3, ())
        h1 = h3
    c.add_wires(h1, (), c, ())
    return c


def grover_circuit(n_search, x, iterations):
    """Same algorithm with structured gates (config 5: 35 + 1 qubits): H on everything, then
    `iterations` x [oracle, diffusion].  Input [0]*n_search + [1]."""
    n = n_search
    c = mc.Circuit(n + 1)
    h = c.add_gate(mc.H, n + 1)
    c.add_wires(c, (), h, ())
    prev_search = [(h, i) for i in range(n)]
    prev_anc = (h, n)
    for t in range(iterations):
        f = c.add(xg.GroverOracle(n, x, 'f' + str(t)))
        for i in range(n):
            c.add_wire(prev_search[i][0], prev_search[i][1], f, i)
        c.add_wire(prev_anc[0], prev_anc[1], f, n)
        d = c.add(xg.GroverDiffusion(n, 'd' + str(t)))
        for i in range(n):
            c.add_wire(f, i, d, i)
        prev_search = [(d, i) for i in range(n)]
        prev_anc = (f, n)
    for i in range(n):
        c.add_wire(prev_search[i][0], prev_search[i][1], c, i)
    c.add_wire(prev_anc[0], prev_anc[1], c, n)
    return c


# ---- Shor -----------------------------------------------------------------------------------------

def shor_circuit(N, a, dense=True):
    """H^(q/2) -> U -> QFT(q/2) on q = 2*bitlen(N) qubits (main/Shor.py:50-94).  dense=True hands the
    engine the same 2^q x 2^q 0/1 matrix object the reference builds; dense=False the structured gate."""
    q = (len(bin(N)) - 2) * 2
    h = q // 2
    first, second = list(range(h)), list(range(h, q))
    c = mc.Circuit(q)
    hg = c.add_gate(mc.H, h)
    c.add_wires(c, first, hg, first)
    ug = xg.ModExp(a, N, 'U')
    if dense:
        ug = mc.Gate(Matrix(ug.matrix.rows), 'U', True)
    c.add(ug)
    c.add_wires(hg, first, ug, first)
    c.add_wires(c, second, ug, second)
    c.add_wires(ug, second, c, second)
    qg = c.add(mc.Gate(Matrix.QFT(h), 'QFT', True))
    c.add_wires(ug, first, qg, first)
    c.add_wires(qg, first, c, first)
    return c, q


def is_prime(n):
    if n in (2, 3):
        return True
    return all(n % i for i in range(2, int(n ** 0.5) + 1))


def closest_fraction(value, maxd):
    """Smallest-denominator fraction within 1/maxd of value (the reference's `approx`, main/Shor.py:16-23)."""
    d = 1
    x = fractions.Fraction.from_float(value).limit_denominator(d)
    while abs(value - x) >= 1 / maxd:
        d += 1
        x = fractions.Fraction.from_float(value).limit_denominator(d)
    return x


def shor_factor(N, dense=True, on_run=None):
    """Shor's algorithm with the reference's control flow and RNG consumption (main/Shor.py:34-123):
    randint for the base, gcd shortcut, sample until the first register is non-zero, continued
    fractions, recursion on the two factors."""
    found = []
    while N != 1:
        if is_prime(N):
            found.append(N)
            break
        a = random.randint(2, N - 1)
        g = gcd(a, N)
        if g != 1:
            found = found + shor_factor(g, dense, on_run)
            N = N // g
            continue
        c, q = shor_circuit(N, a, dense)
        h = q // 2
        in_v = [0] * h + [1] * h
        r = 0
        while r == 0:
            out = c.run(in_v)
            if on_run is not None:
                on_run(a, out)
            r = int(''.join(str(b) for b in out[:h]), 2)
        r = closest_fraction(r / (2 ** h), 2 ** h).denominator
        if r % 2 == 1:
            continue
        if a ** (r // 2) % N == N - 1:
            continue
        m1 = gcd(a ** (r // 2) + 1, N)
        m2 = gcd(a ** (r // 2) - 1, N)
        m = m1 if m1 > m2 else m2
        found = found + shor_factor(m, dense, on_run)
        N = N // m
    return found
