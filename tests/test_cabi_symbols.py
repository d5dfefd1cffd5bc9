"""The C-ABI library builds, loads and exports every symbol include/qsv.h declares (no compute
calls: there is no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from qsv import _cabi
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "qsv.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(qsv_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in include/qsv.h"
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.qsv_abi_version() == 1
    assert lib.qsv_launch_count() == 0


def test_struct_layouts_match_header():
    from qsv import _cabi, schedule
    assert ctypes.sizeof(_cabi.QsvPass) == 48
    assert schedule.OP_DTYPE.itemsize == 88 and schedule.OP_DTYPE.itemsize <= _cabi.QSV_OP_HEADER_BYTES
    offs = {n: schedule.OP_DTYPE.fields[n][1] for n in schedule.OP_DTYPE.names}
    assert offs == {"kind": 0, "k": 4, "n_ins": 8, "flags": 12, "ext_ctrl_mask": 16, "ext_ctrl_val": 24,
                    "loc_ctrl_val": 32, "payload_off": 36, "op_bytes": 40, "reserved": 44, "ins": 48, "tgt": 64,
                    "scalar": 72}


def test_argument_validation_without_gpu():
    """Host-side validation runs before any CUDA call, so error paths are testable on the CPU."""
    from qsv import _cabi
    lib = _cabi.load()
    rc = lib.qsv_apply_cx(None, 4, 0, None, 1, 0, None)
    assert rc != 0 and b"NULL" in lib.qsv_last_error()
    buf = ctypes.c_void_p(16)
    rc = lib.qsv_apply_cx(buf, 4, 0, None, 9, 0, None)
    assert rc != 0 and b"not a local qubit" in lib.qsv_last_error()
    rc = lib.qsv_apply_dense(buf, 4, 6, _cabi.int_array([0, 1, 2, 3, 4, 5]), 0, None,
                             (ctypes.c_double * 2)(), 0, None)
    assert rc != 0 and b"qsv_apply_dense_wide" in lib.qsv_last_error()


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from qsv import engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.StateVector(3)
    import main.circuit as mc
    c = mc.Circuit(1)
    x = c.add_gate(mc.X)
    c.add_wire(c, 0, x, 0)
    c.add_wire(x, 0, c, 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.run([0], 2)


def test_round_programs_pass_host_validation():
    """Every round program the scheduler emits must get past qsv_run_pass_rounds' host-side checks; without
    a GPU the call then fails inside the CUDA runtime, never in the argument validation."""
    import torch
    if torch.cuda.is_available():
        return
    import ctypes as C
    import statevec_oracle as so
    from helpers import circuit_from_steps
    from qsv import _cabi, schedule
    from qsv.lowering import lower
    lib = _cabi.load()
    n = 14
    c = circuit_from_steps(so.random_layered_steps(n, 6, seed=2), n)
    c.sort()
    ops, _ = lower(c)
    packed = schedule.pack_passes(schedule.plan_passes(ops, n, 12, 6))
    seen = 0
    for ps, off, n_ops, extra in packed.passes:
        if n_ops >= 0:
            continue
        prog = extra[0]
        cp = _cabi.QsvPass()
        cp.n_local, cp.tile_bits, cp.low_bits, cp.n_hi = ps.n_local, ps.tile_bits, ps.low_bits, ps.tile_bits - ps.low_bits
        for j, p in enumerate(ps.hi_pos):
            cp.hi_pos[j] = p
        cp.reserved = int(prog[12])
        rc = lib.qsv_run_pass_rounds(C.c_void_p(16), C.byref(cp), C.c_void_p(prog.ctypes.data), int(prog.size),
                                     C.c_void_p(16), None)
        msg = lib.qsv_last_error().decode()
        assert rc != 0 and ("cuda" in msg.lower() or "kernel launch failed" in msg), msg
        seen += 1
    assert seen >= 1


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host cores only (the oracle port): one JSON line with the contract's
    keys, usable on a box without a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-qubits", "12"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gates/sec" and line["unit"] == "gates/sec"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "gates/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
