"""First Circuit.run of a fresh process (bench.py's `e2e_cold`): builds the config-3 circuit, creates the CUDA context,
then times the first and the second `Circuit.run(in_v, 2)`.  Whether the on-disk caches (cubins of the pass kernels,
placement, schedule: QSV_CACHE_DIR) are warm is the caller's business - bench.py runs this once after its own e2e runs
(warm) and once with an empty cache directory (cold: NVRTC + schedule search included)."""
import contextlib
import ctypes
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantum-circuits_b200"))


def main():
    n, depth, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import random
    import torch
    from qsv import _cabi, lowering, workloads
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")                 # CUDA context + allocator
    torch.cuda.synchronize()
    lib = _cabi.load()
    t0 = time.perf_counter()
    c = workloads.random_layered_circuit(n, depth, seed=seed)
    t_build = time.perf_counter() - t0
    random.seed(1)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            out = c.run([0] * n, 2)
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    jc, jh, js = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_double(0)
    lib.qsv_jit_stats(ctypes.byref(jc), ctypes.byref(jh), ctypes.byref(js))
    print("RESULT" + json.dumps({"first_call_ms": times[0], "second_call_ms": times[1], "third_call_ms": times[2],
                                 "circuit_build_ms": t_build * 1e3, "kernels_compiled": int(jc.value),
                                 "kernels_from_disk": int(lib.qsv_jit_disk_hits()), "compile_cpu_s": float(js.value),
                                 "outcome_bits": len(out)}))


if __name__ == "__main__":
    main()
