"""Host side of the Feynman path sum (csrc/qsv_pathsum.cu): Circuit.run_method1 of the reference,
main/circuit.py:221-276, as the reference defines it - a sum over the assignments of the internal wires of products of
gate-matrix entries - evaluated on the device.

The reference's cost model (main/circuit.py:343-351) sends a circuit to method 1 when gates * 2^(n + internal wires)
is small; `feasible` applies a hard limit on top (the sum is exponential in the internal wires by definition).  Larger
circuits take the state-vector engine with transposed matrices, which gives the same weights (the path sum indexes
gate.matrix[in][out])."""
import ctypes as C
import os

import numpy as np

from . import _cabi

MAX_LOG2_PATHS = 34          # n + internal wires: 2^34 path products (x gates) is seconds on one B200
DEFAULT_LOG2_PATHS = 26      # run_method1 takes the path sum by itself up to here (milliseconds); above, the engine
MAX_GATE_QUBITS = 12         # a gate's matrix is read as a dense 2^k x 2^k table


def feasible(circuit, log2_paths=MAX_LOG2_PATHS):
    n = len(circuit)
    return (1 <= n <= 30 and n + len(circuit.get_internal_wires()) <= log2_paths
            and all(len(g.in_wires) <= MAX_GATE_QUBITS for g in circuit.gates))


def wanted(circuit):
    """Does Circuit.run_method1 sum paths (True) or evolve the state with transposed matrices (False)?
    QSV_METHOD1=pathsum: whenever feasible; =engine: never; unset: up to 2^DEFAULT_LOG2_PATHS paths."""
    mode = os.environ.get("QSV_METHOD1", "")
    if mode == "engine":
        return False
    return feasible(circuit, MAX_LOG2_PATHS if mode == "pathsum" else DEFAULT_LOG2_PATHS)


def describe(circuit, in_v):
    """Plain arrays for qsv_path_sum: (wires [n_wires, 2] uint8, gate_hdr [n_gates, 3] uint32, ports uint32, mats
    complex128, direct [n_direct, 2] uint32, n_internal).  Gates in circuit.gates order (the product commutes)."""
    n = len(circuit)
    internal = circuit.get_internal_wires()
    w_int = {id(w): j for j, w in enumerate(internal)}
    n_int = len(internal)
    wire_id = {}
    wires = []

    def wid(w):
        k = id(w)
        if k not in wire_id:
            wire_id[k] = len(wires)
            if w.left is circuit:                       # fed by a circuit input: constant (also when it is a direct wire)
                wires.append((0, int(in_v[w.lind])))
            elif w.right is circuit:                    # a circuit output: bit of the output assignment
                wires.append((1, n - 1 - w.rind))
            else:                                       # internal: bit of the assignment (first wire = slowest digit)
                wires.append((2, n_int - 1 - w_int[k]))
        return wire_id[k]

    hdr, ports, mats = [], [], []
    moff = 0
    for g in circuit.gates:
        k = len(g.in_wires)
        M = np.array(g.matrix.rows, dtype=np.complex128)
        if M.shape != (1 << k, 1 << k):
            raise RuntimeError("Gates should be represented by square matrices.")
        hdr.append((k, moff, len(ports)))
        ports.extend([wid(w) for w in g.in_wires] + [wid(w) for w in g.out_wires])
        mats.append(M.reshape(-1))
        moff += M.size
    direct = [(int(in_v[w.lind]), n - 1 - w.rind) for w in circuit.input if w.right is circuit]
    if not wires:
        wires.append((0, 0))
    return (np.array(wires, dtype=np.uint8).reshape(-1, 2),
            np.array(hdr, dtype=np.uint32).reshape(-1, 3),
            np.array(ports if ports else [0], dtype=np.uint32),
            np.concatenate(mats) if mats else np.zeros(1, dtype=np.complex128),
            np.array(direct, dtype=np.uint32).reshape(-1, 2), n_int)


def emulate(desc, n):
    """The kernel's arithmetic in numpy loops (tests: the packing is checked on the CPU against the goldens)."""
    wires, hdr, ports, mats, direct, n_int = desc
    out_w = np.zeros(1 << n)
    for out in range(1 << n):
        if any(((out >> int(pos)) & 1) != int(bit) for bit, pos in direct):
            continue
        total = 0j
        for a in range(1 << n_int):
            prod = 1 + 0j
            for k, moff, poff in hdr:
                k, moff, poff = int(k), int(moff), int(poff)
                i = j = 0
                for p in range(k):
                    for which, base in ((0, poff + p), (1, poff + k + p)):
                        kind, idx = wires[ports[base]]
                        v = int(idx) if kind == 0 else ((out if kind == 1 else a) >> int(idx)) & 1
                        if which == 0:
                            i = (i << 1) | v
                        else:
                            j = (j << 1) | v
                m = mats[moff + (i << k) + j]
                if m == 0:
                    prod = 0
                    break
                prod *= m
            total += prod
        out_w[out] = abs(total) ** 2
    return out_w


def weights(circuit, in_v, device=None):
    """run_method1's weights (list index = output assignment, wire 0 most significant) as a float64 numpy array (one D2H of 2^n doubles)."""
    import torch
    from . import engine
    engine._require_cuda()
    lib = _cabi.load()
    dev = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
    wires, hdr, ports, mats, direct, n_int = describe(circuit, in_v)
    n = len(circuit)
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in
         (wires, hdr, ports, mats.view(np.float64), direct if direct.size else np.zeros((1, 2), dtype=np.uint32))]
    out = torch.empty(1 << n, dtype=torch.float64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(lib.qsv_path_sum(C.c_void_p(t[0].data_ptr()), int(wires.shape[0]), C.c_void_p(t[1].data_ptr()),
                                 C.c_void_p(t[2].data_ptr()), C.c_void_p(t[3].data_ptr()), int(hdr.shape[0]), n, n_int,
                                 C.c_void_p(t[4].data_ptr()), int(direct.shape[0]), C.c_void_p(out.data_ptr()), st),
                "qsv_path_sum")
    torch.cuda.current_stream(dev).synchronize()
    return out.cpu().numpy()
