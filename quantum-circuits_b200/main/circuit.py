"""Host-side mirror of the reference's circuit module (main/circuit.py) on top of the B200 engine.

Public surface, argument meaning, mutation behaviour and error texts follow the reference
(/root/reference/main/circuit.py:9-545: Wire, Circuit, Gate and the eight gate classes), so the
reference's own Grover.py / Shor.py / test.py / GUI run on top of this module unchanged.  The graph
half (build, validate, sort, trace qubits) is small host logic.  The simulation half is not: run(),
run_method2() and run_method1() lower the sorted circuit to a flat op list (qsv.lowering) and hand
it to the CUDA state-vector engine (qsv.engine) — no permutation matrices, no Kronecker products,
no CPU fallback.

The star-import surface of the reference module is kept too (`from main.circuit import *` gives
Matrix, tensor, Counter, product, choices, choice, randint, log, modf, uuid4).
"""
from main.matrix import Matrix, tensor
from collections import Counter
from itertools import product
from random import choices, choice, randint
from math import log, modf
from uuid import uuid4
import random as _random


class Wire(object):
    """Directed connection: output port `lind` of `left` -> input port `rind` of `right`
    (either end may be the circuit itself).  Reference: main/circuit.py:9-24."""

    def __init__(self, left, right, lind, rind):
        self.left = left
        self.right = right
        self.lind = lind
        self.rind = rind
        self.value = None

    def __str__(self):
        return str(self.left) + " --> " + str(self.right)

    def is_internal(self):
        return isinstance(self.left, Gate) and isinstance(self.right, Gate)


import weakref as _weakref
from itertools import chain as _chain
from operator import attrgetter as _attrgetter
_W_LEFT, _W_LIND, _W_RIGHT = _attrgetter("left"), _attrgetter("lind"), _attrgetter("right")
_SORT_CACHE = _weakref.WeakKeyDictionary()      # Circuit -> [(structure before, sorted order as indices, structure after)]


class Circuit(object):
    """Reference: main/circuit.py:27-445."""

    def __init__(self, size):
        self.input = [None] * size
        self.output = [None] * size
        self.gates = []
        self.ordered_gates = []
        self.wires = []
        self.cls_ids = Counter()

    def __len__(self):
        return len(self.input)

    def __iter__(self):
        return iter(self.gates)

    # ---------------------------------------------------------------- building
    def add(self, component):
        self.gates.append(component)
        if component.parent is not None:
            raise RuntimeError("Component is already added to a circuit.")
        self.cls_ids[component.name] += 1
        component.set_parent(self)
        return component

    def add_gate(self, t, *args):
        return self.add(t(*args))

    def add_wire(self, from_component, from_port, to_component, to_port):
        if from_component not in self.gates and from_component is not self:
            raise RuntimeError("Cannot add wire: Start point is not a valid component.")
        if to_component not in self.gates and to_component is not self:
            raise RuntimeError("Cannot add wire: End point is not a valid component.")
        if from_port >= len(from_component):
            raise RuntimeError("Cannot add wire: {0} does not have {1} output ports.".format(from_component, from_port+1))
        if to_port >= len(to_component):
            raise RuntimeError("Cannot add wire: {0} does not have {1} input ports.".format(to_component, to_port+1))

        w = Wire(from_component, to_component, from_port, to_port)
        self.wires.append(w)
        src = from_component.out_wires if isinstance(from_component, Gate) else from_component.input
        dst = to_component.in_wires if isinstance(to_component, Gate) else to_component.output
        src[from_port] = w
        dst[to_port] = w
        return w

    def add_wires(self, from_component, from_ports, to_component, to_ports):
        if from_ports == 'all' or len(from_ports) == 0:
            from_ports = tuple(range(len(from_component)))
        if to_ports == 'all' or len(to_ports) == 0:
            to_ports = tuple(range(len(to_component)))
        if len(from_ports) != len(to_ports):
            raise RuntimeError("Cannot add wires: Number of output ports does not match number of input ports.")
        return [self.add_wire(from_component, a, to_component, b) for a, b in zip(from_ports, to_ports)]

    def get_internal_wires(self):
        return [w for w in self.wires if w.is_internal()]

    def remove_gate(self, g):
        if not g:
            return
        for w in g.in_wires + g.out_wires:
            self.remove_wire(w)
        self.gates.remove(g)

    def remove_wire(self, w):
        if not w:
            return
        if w.left is self:
            self.input[w.lind] = None
        elif w.left:
            w.left.out_wires[w.lind] = None
        if w.right is self:
            w.right.output[w.rind] = None
        elif w.right:
            w.right.in_wires[w.rind] = None
        self.wires.remove(w)

    # ---------------------------------------------------------------- whole-circuit unitary
    def get_subcircuit(self, name):
        """Compress the circuit into one Gate (reference main/circuit.py:122-157).  Column c of the
        unitary is the engine's final state for basis input c."""
        from qsv import lowering
        return Gate(Matrix(lowering.circuit_unitary_rows(self)), name)

    # ---------------------------------------------------------------- ordering / validation
    def visit(self, gate):
        """Depth-first post-order from `gate`, prepending finished gates to ordered_gates
        (reference main/circuit.py:160-171).  Iterative, same visiting order."""
        if gate.tempmarked:
            raise RuntimeError("Cannot sort - the circuit contains a cycle.")
        if gate.marked:
            return
        stack = [(gate, 0)]
        gate.tempmarked = True
        while stack:
            g, port = stack.pop()
            descended = False
            while port < len(g.out_wires):
                nxt = g.out_wires[port].right
                port += 1
                if nxt is self:
                    continue
                if nxt.tempmarked:
                    raise RuntimeError("Cannot sort - the circuit contains a cycle.")
                if not nxt.marked:
                    stack.append((g, port))
                    nxt.tempmarked = True
                    stack.append((nxt, 0))
                    descended = True
                    break
            if not descended:
                g.marked = True
                g.tempmarked = False
                self.ordered_gates.insert(0, g)

    def sort(self):
        self._sort_structure()

    def _structure(self):
        """Everything check() / sort() / the lowering read of the GRAPH (not of the gate matrices) as one tuple of tuples:
        the gate order, per gate where its in-wires come from and where its out-wires lead, the output wiring.  None
        while the circuit is unfinished (a missing wire).  run() is called over and over on one circuit (Shor's
        sampling loop, benchmarks): an unchanged tuple means the previous sort's result still holds."""
        gates = self.gates
        try:
            ins = [g.in_wires for g in gates]
            outs = [g.out_wires for g in gates]
            win = list(_chain.from_iterable(ins))
            wout = list(_chain.from_iterable(outs))
            # (C-level loops over the ~3 wires per gate: attrgetter raises AttributeError on a missing wire, None)
            return (len(gates), tuple(map(id, gates)), tuple(map(len, ins)), tuple(map(len, outs)),
                    tuple(map(id, map(_W_LEFT, win))), tuple(map(_W_LIND, win)), tuple(map(id, map(_W_RIGHT, wout))),
                    tuple(map(id, map(_W_LEFT, self.output))), tuple(map(_W_LIND, self.output)),
                    tuple(map(id, map(_W_RIGHT, self.input))))
        except AttributeError:          # a port without a wire (None): let check() report it
            return None

    def _sort_structure(self):
        """sort() of the reference (main/circuit.py:173-184: check, depth-first post-order, gates re-ordered), returning
        the structure tuple of the SORTED circuit (None if it cannot be described).  The result is remembered per
        structure (a few entries per circuit, module-level and weakly keyed: nothing is added to the instance, its
        pickles stay what the reference writes; entries hold integers only): the same graph sorts to the same order
        without walking it again."""
        before = self._structure()
        cache = _SORT_CACHE.get(self)
        if before is not None and cache is not None:
            for fp, order, after in cache:
                if fp == before:
                    gates = self.gates
                    self.gates = [gates[i] for i in order]
                    return after
        if not self.check():
            raise RuntimeError("Cannot sort - please finish building the circuit first.")
        for gate in self.gates:
            if not gate.marked:
                self.visit(gate)
        self.gates = self.ordered_gates
        self.ordered_gates = []
        for gate in self.gates:
            gate.marked = False
            gate.tempmarked = False
        after = self._structure()
        if before is not None and after is not None:
            if cache is None:
                cache = _SORT_CACHE[self] = []
            if len(cache) >= 4:
                cache.pop(0)
            # the order as a permutation of the gate list it started from: no reference to a gate (a gate points back at
            # its circuit, and an entry that kept the key alive would never leave the weak table)
            position = {gid: i for i, gid in enumerate(before[1])}
            cache.append((before, [position[id(g)] for g in self.gates], after))
        return after

    def check(self):
        return (
            all(w is not None for w in self.input) and
            all(w is not None for w in self.output) and
            all(w is not None for g in self for w in g.in_wires) and
            all(w is not None for g in self for w in g.out_wires) and
            not self.contains_cycle()
        )

    def contains_cycle(self):
        """True iff the gate graph has a directed cycle (reference main/circuit.py:199-217)."""
        WHITE, GREY, BLACK = 0, 1, 2
        colour = {}                     # keyed by id(): Gate.__hash__ goes through the uuid (three Python calls)
        for root in self.gates:
            if colour.get(id(root), WHITE) != WHITE:
                continue
            colour[id(root)] = GREY
            stack = [(root, 0)]
            while stack:
                g, port = stack.pop()
                descended = False
                while port < len(g.out_wires):
                    w = g.out_wires[port]
                    port += 1
                    if w is None or w.right is self:
                        continue
                    c = colour.get(id(w.right), WHITE)
                    if c == GREY:
                        return True
                    if c == WHITE:
                        stack.append((g, port))
                        colour[id(w.right)] = GREY
                        stack.append((w.right, 0))
                        descended = True
                        break
                if not descended:
                    colour[id(g)] = BLACK
        return False

    # ---------------------------------------------------------------- simulation (the hot path)
    def _validate_input(self, in_v):
        if len(in_v) != len(self):
            raise RuntimeError("Input length does not match the size of the circuit.")
        for i in in_v:
            if i == 0 or i == 1:
                pass
            else:
                raise RuntimeError("Invalid input - values must be 0 or 1.")

    def run_method1(self, in_v):
        """Reference: Feynman path sum (main/circuit.py:221-276).  Circuits with few paths (qsv.pathsum.wanted:
        2^(n + internal wires) within the limit, QSV_METHOD1=pathsum|engine overrides) are summed path by path on
        the device, as the reference defines the method; the rest get the same weights from the state-vector
        engine - the path sum indexes gate.matrix[in][out], i.e. applies the TRANSPOSED gate matrices, which is
        reproduced."""
        self._validate_input(in_v)
        for i in range(len(self)):
            (self.input[i]).value = in_v[i]
        from qsv import lowering, pathsum
        if pathsum.wanted(self):
            return pathsum.weights(self, in_v).tolist()
        return lowering.simulate_weights(self, in_v, transposed=True, presorted=False)

    def run_method2(self, in_v):
        """State-vector evolution (reference main/circuit.py:279-337) on the GPU engine."""
        structure = self._sorted()
        self._validate_input(in_v)
        from qsv import lowering
        return lowering.simulate_weights(self, in_v, transposed=False, presorted=True, structure=structure)

    def _sorted(self):
        """self.sort(); the structure tuple of the sorted circuit when our own sort() did it (the lowering then skips
        re-deriving what every gate acts on for a graph it has seen), None for a subclass's sort()."""
        if type(self).sort is Circuit.sort:
            return self._sort_structure()
        self.sort()
        return None

    def run(self, in_v, method=None):
        """One measurement outcome as a list of bits, wire 0 first (reference main/circuit.py:341-366).
        The method selector and its prints are kept; both methods execute on the engine.  Exactly one
        random.random() is drawn per call, like random.choices does."""
        if not method:
            int_wires = self.get_internal_wires()
            n1 = len(self.gates)*(2**(len(self)+len(int_wires)))
            n2 = (3*len(self.gates))*(2**(2*len(self)))
            if n1 <= n2:
                print("Running method 1...")
                method = 1
            else:
                print("Running method 2...")
                method = 2

        structure = None
        if method == 1:
            self._validate_input(in_v)
            for i in range(len(self)):
                (self.input[i]).value = in_v[i]
            transposed, presorted = True, False
        elif method == 2:
            structure = self._sorted()
            self._validate_input(in_v)
            transposed, presorted = False, True
        else:
            raise RuntimeError("Please choose one of the avalible methods: 1 or 2")

        from qsv import lowering
        index = lowering.simulate_sample(self, in_v, _random.random, transposed=transposed, presorted=presorted,
                                         structure=structure)
        n = len(self)
        return [(index >> (n - 1 - i)) & 1 for i in range(n)]

    # ---------------------------------------------------------------- qubit tracing
    def endqbit(self, gate_index, outwire_index):
        gate = self.gates[gate_index]
        j = outwire_index
        wire = gate.out_wires[j]
        while gate.out_wires[j].right != self:
            gate = gate.out_wires[j].right
            j = wire.rind
            wire = gate.out_wires[j]
        return wire.rind

    def startqbit(self, gate_index, inwire_index):
        gate = self.gates[gate_index]
        j = inwire_index
        wire = gate.in_wires[j]
        while gate.in_wires[j].left != self:
            gate = gate.in_wires[j].left
            j = wire.lind
            wire = gate.in_wires[j]
        return wire.lind

    def startqbits(self):
        out = []
        for i in range(len(self)):
            wire = self.output[i]
            while wire.left != self:
                gate = wire.left
                wire = gate.in_wires[wire.lind]
            out.append(wire.lind)
        return out

    @staticmethod
    def random_circuit(size=0, gates=0):
        """Random circuit over {X,Y,Z,H,CNot,T}; consumes the global `random` stream in the same order
        as the reference generator (main/circuit.py:406-445)."""
        if size == 0:
            size = randint(3, 6)
        n_gates = randint(1, size) if gates == 0 else gates

        c = Circuit(size)
        sources = [c]
        free_ports = [list(range(size))]

        def take():
            j = randint(0, len(sources)-1)
            k = randint(0, len(free_ports[j])-1)
            comp, port = sources[j], free_ports[j][k]
            del free_ports[j][k]
            if free_ports[j] == []:
                del free_ports[j]
                del sources[j]
            return comp, port

        for i in range(n_gates):
            label = "Gate "+str(i)
            gate = c.add(choice([X(1, label), Y(1, label), Z(1, label), H(1, label), CNot(label), T(label)]))
            for in_wire in range(len(gate)):
                comp, port = take()
                c.add_wire(comp, port, gate, in_wire)
            sources.append(gate)
            free_ports.append(list(range(len(gate))))

        for i in range(len(c)):
            comp, port = take()
            c.add_wire(comp, port, c, i)
        return c


class Gate(object):
    """Reference: main/circuit.py:448-511."""
    SIZE = 1

    def __init__(self, mtrx, name, bypass_error_check=False):
        if not bypass_error_check:
            symbolic = getattr(mtrx, "structure", None) is not None and mtrx.structure() is not None
            if not symbolic and len(mtrx) != len(mtrx.rows[0]):          # (symbolic matrices are square by construction)
                raise RuntimeError("Gates should be represented by square matrices.")
            if not mtrx.isUnitary():
                raise RuntimeError("Gates should be represented by unitary matrices.")
            if modf(log(len(mtrx), 2))[0] != 0:
                raise RuntimeError("Invalid gate dimension.")

        size = int(log(len(mtrx), 2))

        self.matrix = mtrx
        self.in_wires = [None] * size
        self.out_wires = [None] * size
        self.name = name

        self.marked = False
        self.tempmarked = False

        self.parent = None
        self.uuid = None
        self.id = None

    def __len__(self):
        return len(self.in_wires)

    def __str__(self):
        s = self.name
        if self.parent is not None and self.parent.cls_ids[self.name] > 1:
            s += str(self.id)
        return s

    def __eq__(self, other):
        if isinstance(other, self.__class__):
            return self.uuid == other.uuid
        return NotImplemented

    def __ne__(self, other):
        if isinstance(other, self.__class__):
            return not self == other
        return NotImplemented

    def __hash__(self):
        return hash(self.uuid)

    def set_parent(self, parent):
        self.parent = parent
        self.id = parent.cls_ids[self.name]
        self.uuid = uuid4()
        while any(g.uuid == self.uuid for g in parent if g is not self):
            self.uuid = uuid4()

    def get_matrix(self):
        """Fresh copy of the gate matrix (reference main/circuit.py:504-511)."""
        dim = 2**(len(self))
        return Matrix([[self.matrix[i][j] for j in range(dim)] for i in range(dim)])


class X(Gate):
    def __init__(self, size=1, name="X"):
        super().__init__(Matrix.X(size), name, True)


class Y(Gate):
    def __init__(self, size=1, name="Y"):
        super().__init__(Matrix.Y(size), name, True)


class Z(Gate):
    def __init__(self, size=1, name="Z"):
        super().__init__(Matrix.Z(size), name, True)


class H(Gate):
    def __init__(self, size=1, name="H"):
        super().__init__(Matrix.H(size), name, True)


class SqrtNot(Gate):
    def __init__(self, size=1, name="SqrtNot"):
        super().__init__(Matrix.SqrtNot(size), name, True)


class QFT(Gate):
    def __init__(self, size=1, name="QFT"):
        super().__init__(Matrix.QFT(size), name, True)


class CNot(Gate):
    SIZE = 2

    def __init__(self, name="CNot"):
        super().__init__(Matrix.Cnot(), name, True)


class T(Gate):
    SIZE = 3

    def __init__(self, name="T"):
        super().__init__(Matrix.T(), name, True)
