#!/usr/bin/env python
"""Benchmark of the state-vector hot path (BASELINE.json metric: gates/sec at 30 q on 1 GPU and at
30+log2(N) q sharded over N GPUs; fraction of the HBM roofline per gate pass).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--qubits Q] [--depth D]
    python bench.py --impl reference ...      # CPU arm: the oracle port timed on the host cores

One "step" = one complete run of the config-3 circuit (random H / X / T8 / phase layers with CNot / CZ
matchings, depth 20 -> 45 gates per layer at 30 q) on a fresh |0...0> state: state preparation, every
gate pass, and one measurement sample.  `value` times that with the packed program already in HBM;
`e2e` times Circuit.run(in_v, 2) through the public API (host scheduling, H2D of the program, D2H of
the sampled index inside the timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "quantum-circuits_b200"),):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "gates/sec"
SEED = 20261019


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--qubits", type=int, default=0, help="total qubits (default 30 + log2(gpus))")
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-qubits", type=int, default=22, help="register size of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gate-pass", action="store_true", help="skip the single-gate-pass roofline sweep")
    ap.add_argument("--workload", default="layered", choices=["layered", "qft", "grover"],
                    help="layered = BASELINE config 3 (headline); qft = config 4 (H + controlled-phase ladder); "
                         "grover = config 5 (oracle + diffusion kernels)")
    ap.add_argument("--iterations", type=int, default=32, help="Grover iterations (config 5)")
    ap.add_argument("--no-cold", action="store_true", help="skip e2e_cold (fresh-process first call) and the "
                    "reference Grover-7 leg")
    ap.add_argument("--no-check", action="store_true", help="skip the in-run correctness check (norm, hint on/off, "
                                                            "circuit^-1 o circuit round trip)")
    ap.add_argument("--no-big", action="store_true",
                    help="skip the 137 GB-per-GPU block (33 qubits on one GPU, 33+log2(N) on N: 36 on 8 = north star)")
    ap.add_argument("--big-steps", type=int, default=3)
    ap.add_argument("--cpu-threads", type=int, default=0,
                    help="host threads of the CPU arm (default: min(host cores, 32), the same at every --gpus)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def mark(self):
        """Start of the timed region: nvidia-smi has been streaming since before the warm-up (its start-up takes longer
        than a short timed region), only the samples from here on count."""
        self.t0 = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        inside = [ln for t, ln in self.lines if t >= t0]
        if not inside and self.lines:
            inside = [self.lines[-1][1]]            # region shorter than one sampling interval: the sample next to it
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def traffic_from_profile(kind, n_local):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel for a DENSE sweep (every
    tile touched), from the committed `ncu --set full` capture (profiles/*_traffic.json: bytes per amplitude),
    scaled to this state size.  Returns (bytes per dense launch, bytes per amplitude, source)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                per_amp = json.load(f)[kind]["dram_bytes_per_amplitude"]
            return per_amp * (1 << n_local), per_amp, "profiles/" + name
        except (OSError, KeyError, ValueError):
            continue
    return None, None, None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores (bounded sample of the same workload)
# ------------------------------------------------------------------------------------------------

def cpu_threads(requested=0):
    """Host threads of the CPU arm: the same number whatever --gpus is (torchrun exports OMP_NUM_THREADS=1 to its
    workers, which used to throttle the N>1 reference arm to one thread)."""
    return int(requested) if requested else max(1, min(os.cpu_count() or 1, 32))


def cpu_baseline(n_cpu, depth=1, budget_s=25.0, threads=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    want = cpu_threads(threads)
    if "numpy" not in sys.modules:      # BLAS pools read these when the library loads
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = str(want)
    import numpy as np
    import statevec_oracle as so
    limiter = None
    try:
        import threadpoolctl
        limiter = threadpoolctl.threadpool_limits(limits=want)
    except Exception:
        pass
    steps = so.random_layered_steps(n_cpu, depth, seed=SEED)
    psi = so.initial_state([0] * n_cpu)
    done = 0
    t0 = time.perf_counter()
    for M, wires, _ in steps:
        psi = so.apply_gate(psi, n_cpu, wires, M)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    w = so.weights(psi)
    so.sample_index_np(w, 0.5)
    try:
        import threadpoolctl
        threads = max([p["num_threads"] for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        threads = 1
    if limiter is not None:
        limiter.restore_original_limits()
    return {"value": done / dt, "unit": METRIC, "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
            "sample": "oracle port (numpy restatement of Circuit.run_method2): config-3 generator, %d qubits, "
                      "first %d gates of layer 1 (the reference itself needs 86 s/gate at 10 qubits)" % (n_cpu, done)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself is not
    present on the GPU box) on the host cores, same metric/unit.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args.cpu_qubits, 1, budget_s=20.0, threads=args.cpu_threads)
        if i >= args.warmup:
            vals.append(base["value"])
        if i == 0 and args.warmup > 0:
            continue
    v = sum(vals) / len(vals)
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": METRIC, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
            "config": {"workload": "config-3 random H/X/T8/phase + CNot/CZ layered circuit, CPU sample at %d qubits"
                                   % args.cpu_qubits, "seed": SEED},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))



# ------------------------------------------------------------------------------------------------
# in-run correctness check of the layered workload (any N, any size: properties, not an oracle)
# ------------------------------------------------------------------------------------------------

def layered_check(sv, comm, n, depth, plan, in_index, src_pos, g, roundtrip=True):
    """Three size-independent properties, evaluated on the state the timed steps produce:
    (1) the all-reduced norm is 1; (2) the run with the support hint (tiles known to be zero are skipped) and the
    run without it give bit-identical block sums of |psi|^2; (3) circuit followed by its inverse, lowered, placed
    and planned like the timed circuit, returns to the input basis state (amplitude 1 at index 0 in output order,
    and the sampler returns 0)."""
    import torch
    import torch.distributed as dist
    from qsv import lowering, workloads
    world = comm.world if comm is not None else 1
    rank = comm.rank if comm is not None else 0

    def allsum(x):
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    def allmin(x):
        t = torch.tensor([int(x)], device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    # the timed path: no clear of the register where the leading passes can zero-fill - run here on a register that
    # was deliberately filled with NaN, so that anything read from outside the support would poison the sums
    n_fresh = sv._fresh_prefix(plan)
    if n_fresh:
        torch.view_as_real(sv.state).fill_(float("nan"))
    sv.run_from_basis(in_index, plan)
    sums_hint = sv.block_sums(None, 12).clone()
    norm = allsum(sums_hint.sum().item())
    old = os.environ.get("QSV_SPARSE_START")
    os.environ["QSV_SPARSE_START"] = "0"
    sv.init_basis(in_index)
    sv.execute(plan)
    sums_dense = sv.block_sums(None, 12)
    if old is None:
        os.environ.pop("QSV_SPARSE_START")
    else:
        os.environ["QSV_SPARSE_START"] = old
    identical = bool(allmin(int(torch.equal(sums_hint, sums_dense))))
    # the sample of the timed step (first level of the sampler's descent summed by the last gate pass) against the
    # sample the sampler finds on its own, for three draws
    same = True
    presummed = 0
    for u in (0.37, 0.0, 0.93):
        sv.run_from_basis(in_index, plan, sample_src_pos=src_pos)
        presummed = max(presummed, sv._presum[0] if sv._presum is not None else 0)
        z_fused = sv.sample(u, src_pos)
        sv.run_from_basis(in_index, plan)
        same = same and (z_fused == sv.sample(u, src_pos))
    same = bool(allmin(int(same)))
    out = {"norm_error": abs(norm - 1.0), "hint_on_off_block_sums_bit_identical": identical, "tolerance": 1e-10,
           "zero_fill_passes_on_nan_poisoned_register": int(n_fresh),
           "sample_with_sums_of_the_last_pass_equals_plain_sampler": same, "output_bits_summed_by_the_last_pass": presummed}
    ok = abs(norm - 1.0) <= 1e-10 and identical and same
    if roundtrip:
        rt = workloads.random_layered_roundtrip_circuit(n, depth, seed=SEED)
        rt.sort()
        ops, rt_src, flip = lowering.lower(rt, n_global=g if comm is not None else 0, in_flip=True,
                                           free_swap=bool(comm is not None and comm.peer_swap),
                                           fuse_swaps=sv.fuse_swaps())
        rt_plan = sv.prepare(ops)
        sv.init_basis(0 ^ flip)
        sv.execute(rt_plan)
        a0 = sv.state[0:1].clone() if rank == 0 else torch.zeros(1, dtype=torch.complex128, device="cuda")
        if world > 1:
            dist.all_reduce(torch.view_as_real(a0))
        a0 = complex(a0[0].item())
        rt_norm = allsum(sv.block_sums(None, 12).sum().item())
        z = sv.sample(0.37, rt_src)
        out.update({"roundtrip_gates": len(rt.gates), "roundtrip_amp0_error": abs(a0 - 1.0),
                    "roundtrip_norm_error": abs(rt_norm - 1.0), "roundtrip_sampled_index": int(z)})
        ok = ok and abs(a0 - 1.0) <= 1e-10 and abs(rt_norm - 1.0) <= 1e-10 and int(z) == 0
        del rt_plan
    out["ok"] = bool(ok)
    return out


def run_big_shard(args, comm, g, world, rank, barrier):
    """The 137 GB-per-GPU point: config-3 generator at 33 qubits on one GPU and 33 + log2(N) qubits on N GPUs
    (36 qubits / 1.1 TB on 8 = BASELINE configs[4] sizing).  Timed like the headline (K steps between
    barriers, CUDA events, max over ranks), with and without the support hint, plus the in-run check."""
    import torch
    import torch.distributed as dist
    from qsv import engine, lowering, workloads
    n = 33 + g
    need = (16 << 33) + (2 << 30)
    torch.cuda.empty_cache()
    free = torch.cuda.mem_get_info()[0]
    t = torch.tensor([free], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    free = int(t.item())
    if free < need:
        return {"qubits": n, "skipped": "needs %.1f GB free per GPU, %.1f GB available" % (need / 1e9, free / 1e9)}
    circ = workloads.random_layered_circuit(n, args.depth, seed=SEED)
    circ.sort()
    n_gates = len(circ.gates)
    t0 = time.perf_counter()
    sv = engine.StateVector(n, comm=comm)
    ops, src_pos, flip = lowering.lower(circ, n_global=g if comm is not None else 0, in_flip=True,
                                        free_swap=bool(comm is not None and comm.peer_swap),
                                        fuse_swaps=sv.fuse_swaps())
    plan = sv.prepare(ops)
    t_prep = time.perf_counter() - t0
    in_index = 0 ^ flip

    def step():
        sv.run_from_basis(in_index, plan, sample_src_pos=src_pos)
        return sv.sample(0.37, src_pos)

    def timed(k):
        for _ in range(2):
            step()
        if comm is not None:
            comm.swap_stats()
        b0 = comm.bytes_sent if comm is not None else 0
        f0 = getattr(comm, "fused_swaps", 0) if comm is not None else 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0 = sv.passes_run
        barrier()
        e0.record()
        for _ in range(k):
            step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        sw = None
        if comm is not None:
            n_sw, sw_ms = comm.swap_stats()
            out_b = comm.bytes_sent - b0
            gbs = out_b / (sw_ms * 1e-3) / 1e9 if sw_ms > 0 else 0.0
            sw = {"swaps_per_step": n_sw / k, "ms_per_step": sw_ms / k, "bytes_out_per_gpu_per_step": out_b / k,
                  "achieved_GBps_out_per_gpu": gbs, "frac_of_nvlink": gbs / 770.0,
                  "fused_with_pass_per_step": (getattr(comm, "fused_swaps", 0) - f0) / k}
        if world > 1:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / k, (sv.passes_run - p0) // k, sw

    k = max(1, args.big_steps)
    ms_hint, passes, sw_hint = timed(k)
    os.environ["QSV_SPARSE_START"] = "0"
    ms_dense, _, sw_dense = timed(k)
    os.environ.pop("QSV_SPARSE_START")
    check = None if args.no_check else layered_check(sv, comm, n, args.depth, plan, in_index, src_pos, g)
    out = {"qubits": n, "gates": n_gates, "state_bytes_per_gpu": 16 << sv.n_local, "steps": k, "warmup": 2,
           "ms_per_step": ms_hint, "circuit_gates_per_sec": n_gates / (ms_hint * 1e-3),
           "value": n_gates * world / (ms_hint * 1e-3), "passes_per_step": passes,
           "dense_start": {"ms_per_step": ms_dense, "circuit_gates_per_sec": n_gates / (ms_dense * 1e-3)},
           "prepare_wall_s": t_prep, "check": check,
           "workload": "config-3 generator, %d qubits, depth %d, %d gates, complex128, one sample per step; "
                       "2^33 amplitudes (137.4 GB) per GPU" % (n, args.depth, n_gates)}
    if sw_hint is not None:
        out["swap"] = sw_hint
        out["dense_start"]["swap"] = sw_dense
    del plan
    sv.state = None
    del sv
    if comm is not None:
        comm.release_state()
    torch.cuda.empty_cache()
    return out

# ------------------------------------------------------------------------------------------------
# cold start of the public API, and the unmodified reference on its own Grover script
# ------------------------------------------------------------------------------------------------

def _run_json(cmd, env=None, timeout=900):
    import subprocess
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if r.returncode != 0:
        return {"error": (r.stderr or r.stdout)[-400:]}
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")]
    return json.loads(lines[-1][6:]) if lines else {"error": "no RESULT line"}


def cold_start(n, depth, e2e_ms):
    """`e2e_cold`: the first Circuit.run(in_v, 2) of a FRESH process on the same circuit (tools/cold_run.py), with the
    on-disk caches this process has just filled (cubins, schedule) and with an empty cache directory (NVRTC and the
    schedule search included).  The steady-state `e2e` excludes all of this."""
    import tempfile
    cmd = [sys.executable, os.path.join(ROOT, "tools", "cold_run.py"), str(n), str(depth), str(SEED)]
    warm = _run_json(cmd)
    env = dict(os.environ, QSV_CACHE_DIR=tempfile.mkdtemp(prefix="qsv-cold-"))
    cold = _run_json(cmd, env=env)
    out = {"unit": "ms per Circuit.run(in_v, 2), first call of a fresh process (CUDA context already created)",
           "disk_cache_warm": warm, "disk_cache_empty": cold, "steady_state_e2e_ms": e2e_ms}
    if "first_call_ms" in warm:
        out["value"] = warm["first_call_ms"]
        out["ratio_to_e2e"] = warm["first_call_ms"] / e2e_ms
    return out


def reference_grover7():
    """The unmodified reference on its own script (main/Grover.py + Circuit.run_method2: 7 qubits, 25 dense gates)
    against the same circuit through the mirror API on the GPU.  The reference runs live when a checkout is present
    (QSV_REFERENCE_ROOT / /root/reference - the build container); on the GPU box the committed measurement of
    tools/time_reference_grover7.py (profiles/r02_reference_grover7.json) is quoted and labelled as such."""
    import numpy as np
    import torch
    from qsv import workloads
    ref_root = os.environ.get("QSV_REFERENCE_ROOT", "/root/reference")
    ref, src = None, None
    if os.path.isdir(os.path.join(ref_root, "main")):
        import subprocess
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_reference_grover7.py")],
                           capture_output=True, text=True, timeout=1800)
        src = "live run of the checkout at %s" % ref_root
    try:
        with open(os.path.join(ROOT, "profiles", "r02_reference_grover7.json")) as f:
            ref = json.load(f)
        src = src or "profiles/r02_reference_grover7.json (measured in the build container; no checkout on this box)"
    except (OSError, ValueError):
        return None
    g = workloads.grover_circuit_dense(6, 13)
    in_v = [0] * 6 + [1]
    t0 = time.perf_counter()
    w = g.run_method2(in_v)
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    k = 20
    t0 = time.perf_counter()
    for _ in range(k):
        w = g.run_method2(in_v)
    torch.cuda.synchronize()
    steady = (time.perf_counter() - t0) / k
    p = w[2 * 13] + w[2 * 13 + 1]
    return {"circuit": "main/Grover.py:10-48 (n = 6 search qubits + ancilla, x = 13, 6 iterations, 25 dense gates)",
            "reference": {"run_method2_s": ref["run_method2_s"], "gates_per_sec": ref["gates_per_sec"],
                          "p_marked_pair": ref["p_marked_pair"], "source": src, "what": ref["what"]},
            "gpu": {"api": "Circuit.run_method2 through the mirror (workloads.grover_circuit_dense)",
                    "first_call_s": first, "steady_s": steady, "gates_per_sec": len(g.gates) / steady,
                    "p_marked_pair": p},
            "p_marked_pair_abs_diff": abs(p - ref["p_marked_pair"]), "tolerance": 1e-12,
            "ok": bool(abs(p - ref["p_marked_pair"]) <= 1e-12),
            "speedup_steady": ref["run_method2_s"] / steady, "argmax_equal": bool(int(np.argmax(w)) == ref["argmax"])}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from qsv.dist import Comm
        comm = Comm(dist.group.WORLD)
    g = (world - 1).bit_length()
    assert world == 1 << g, "number of GPUs must be a power of two"
    n = args.qubits if args.qubits else 30 + g

    from qsv import _cabi, engine, lowering, workloads, schedule
    lib = _cabi.load()

    # ---- build + lower the workload once (host), upload the packed programs (resident in HBM)
    t_host0 = time.perf_counter()
    check = None
    if args.workload == "layered":
        circ = workloads.random_layered_circuit(n, args.depth, seed=SEED)
        in_bits = [0] * n
        wl_name = ("config-3: random H/X/T8/phase layers + CNot/CZ matchings, %d qubits, depth %d" % (n, args.depth))
    elif args.workload == "qft":
        circ = workloads.qft_ladder_circuit(n)
        in_bits = [(i + 1) % 2 for i in range(n)]
        wl_name = "config-4: QFT as H + controlled-phase ladder (sign +i, bit reversal in the output map), %d qubits" % n
    else:
        x_marked = 0x5A5A5A5A5 & ((1 << (n - 1)) - 1)
        circ = workloads.grover_circuit(n - 1, x_marked, args.iterations)
        in_bits = [0] * (n - 1) + [1]
        wl_name = ("config-5: Grover, %d search qubits + ancilla, H on all then %d x [oracle, diffusion]"
                   % (n - 1, args.iterations))
    circ.sort()
    n_gates = len(circ.gates)
    sv = engine.StateVector(n, comm=comm)      # allocates the shard (symmetric memory when peers can map it)
    t_host1 = time.perf_counter()
    # uncontrolled X gates are pushed back to the input by the lowering: they become `flip` of the start index;
    # with the peer-memory swap kernel, global<->local swaps exchange the victims where they are (free_swap)
    ops, src_pos, flip = lowering.lower(circ, n_global=g if comm is not None else 0, in_flip=True,
                                        free_swap=bool(comm is not None and comm.peer_swap),
                                        fuse_swaps=sv.fuse_swaps())
    t_host = (t_host1 - t_host0) - 0.0 + (time.perf_counter() - t_host1)
    t_prep0 = time.perf_counter()
    plan = sv.prepare(ops)          # pass planning, packing, H2D, and (large states) kernel specialisation
    t_prepare = time.perf_counter() - t_prep0
    in_index = lowering.basis_index(in_bits) ^ flip

    def step(u=0.37):
        # = init_basis + execute + sample, as Circuit.run does it: the 16 * 2^n byte clear is skipped where the leading
        # passes can zero-fill, and the last pass adds up the weights the sampler's first level needs
        sv.run_from_basis(in_index, plan, sample_src_pos=src_pos)
        return sv.sample(u, src_pos)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events, max over ranks
    sampler.mark()
    lib.qsv_reset_launch_count()
    passes0 = sv.passes_run
    if comm is not None:
        comm.swap_stats()
        bytes0 = comm.bytes_sent
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = int(lib.qsv_launch_count())
    passes_per_step = (sv.passes_run - passes0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    swap = None
    if comm is not None:
        n_sw, sw_ms = comm.swap_stats()
        out_bytes = comm.bytes_sent - bytes0
        nv_peak = 770.0      # measured peer-copy GB/s per direction per GPU (B200_PROFILING.md); 900 nominal
        gbs = out_bytes / (sw_ms * 1e-3) / 1e9 if sw_ms > 0 else 0.0
        swap = {"swaps_per_step": n_sw / args.steps, "ms_per_step": sw_ms / args.steps,
                "bytes_out_per_gpu_per_swap": out_bytes / max(1, n_sw), "achieved_GBps_out_per_gpu": gbs,
                "nvlink_peak_GBps": nv_peak, "frac_of_nvlink": gbs / nv_peak,
                "fused_with_pass_per_step": getattr(comm, "fused_swaps", 0) / max(1, args.steps + max(args.warmup, 3)),
                "path": ("k_swap_peer over NVLink peer memory; swaps that carry a gate pass run as the fused "
                         "qsv_jit_pass swap+pass kernel (TMA pull from the peers + per-tile flag handshake)")
                if comm.peer_swap else "NCCL all_to_all_single, chunked in place",
                "note": "device time of the swap kernels on rank 0, from the barrier that starts them to their end "
                        "(peer kernel, incl. the fused pass where one rides along) / incl. pack+unpack (NCCL)"}
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    # the same K steps with the support hint switched off: every pass sweeps the whole state, as if the register
    # did not start in a basis state (reported next to `value`, which is what the API path runs)
    os.environ["QSV_SPARSE_START"] = "0"
    step()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(args.steps):
        step()
    ev3.record()
    barrier()
    ms_dense = ev2.elapsed_time(ev3)
    os.environ.pop("QSV_SPARSE_START")
    if world > 1:
        t = torch.tensor([ms_dense], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dense = float(t.item())
    ms_dense_per_step = ms_dense / args.steps
    circuit_gps = n_gates / (ms_per_step * 1e-3)
    # whole-job aggregate: every gate of the sharded circuit updates the shards of all N ranks, so the units all
    # ranks processed per step are gates x N (one unit = one gate applied to one 2^n_local-amplitude shard).  At
    # N=1 this IS the circuit's gates/sec; at N>1 (weak scaling: 2^30 amplitudes per GPU, 30+log2(N) qubits) it
    # grows with N like any aggregate throughput, while `circuit_gates_per_sec` is the raw rate of the big circuit.
    value = circuit_gps * world

    if args.workload == "grover":
        # closed form: amplitude of the marked pair after K iterations = +-sin((2K+1) theta)/sqrt(2)
        import math
        val = torch.zeros(2, dtype=torch.complex128, device="cuda")
        for i, idx in enumerate(((x_marked << 1), (x_marked << 1) | 1)):       # output order: ancilla = last wire = LSB
            y = 0
            for b in range(n):
                y |= ((idx >> b) & 1) << src_pos[b]          # the layout the swaps left (output bit b <- position)
            if (y >> sv.n_local) == (comm.rank if comm else 0):
                val[i] = sv.state[y & ((1 << sv.n_local) - 1)]
        if world > 1:
            v = torch.view_as_real(val)
            dist.all_reduce(v)
        theta = math.asin(2.0 ** (-(n - 1) / 2))
        expect = math.sin((2 * args.iterations + 1) * theta) / math.sqrt(2)
        err = max(abs(complex(val[0]) - expect), abs(complex(val[1]) + expect))
        check = {"marked_pair_amplitude_error": err, "expected": expect, "tolerance": 1e-12, "ok": bool(err < 1e-12)}

    if args.workload == "qft" and not args.no_check:
        # closed form of Matrix.QFT on a basis state (main/matrix.py:212-219, sign +i): every outcome has weight 2^-n
        # and psi[k] = exp(2 pi i x k / 2^n) / sqrt(2^n) in output order; sampled at 16 output indices, plus the norm
        import cmath
        import random as _random
        step()
        x_in = lowering.basis_index(in_bits)
        rng = _random.Random(7)
        worst = 0.0
        for _ in range(16):
            k = rng.getrandbits(n)
            y = 0
            for b in range(n):
                y |= ((k >> b) & 1) << src_pos[b]
            val = torch.zeros(1, dtype=torch.complex128, device="cuda")
            if (y >> sv.n_local) == (comm.rank if comm else 0):
                loc = y & ((1 << sv.n_local) - 1)
                val = sv.state[loc: loc + 1].clone()
            if world > 1:
                dist.all_reduce(torch.view_as_real(val))
            expect = cmath.exp(2j * cmath.pi * ((x_in * k) % (1 << n)) / (1 << n)) * 2.0 ** (-n / 2)
            worst = max(worst, abs(complex(val[0]) - expect) * 2.0 ** (n / 2))
        nrm = torch.tensor([float(sv.block_sums(None, 12).sum())], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(nrm)
        check = {"relative_amplitude_error_16_samples": worst, "norm_error": abs(float(nrm.item()) - 1.0),
                 "tolerance": 1e-10, "ok": bool(worst < 1e-10 and abs(float(nrm.item()) - 1.0) < 1e-10)}

    # ---- roofline of the dominant kernel (k_pass): per-launch CUDA-event timing of every pass
    peak, peak_src = measured_peak_gbs()
    n_local = sv.n_local
    timed = sv.time_plan(plan, repeats=2, basis=in_index)      # every launch as a step runs it (support hints)
    by_kind = {}
    for t in timed:
        k = by_kind.setdefault(t["kind"], {"ms": 0.0, "bytes": 0.0, "launch_groups": 0, "ops": 0})
        k["ms"] += t["ms"]; k["bytes"] += t["bytes"]; k["launch_groups"] += 1; k["ops"] += t["ops"]
    dom = max(by_kind, key=lambda k: by_kind[k]["ms"]) if by_kind else "pass"
    d = by_kind.get(dom, {"ms": 0.0, "bytes": 0.0, "launch_groups": 0, "ops": 0})
    # load balance of the stand-alone passes: every rank's sum of pass times for one step (with the support hint some
    # ranks hold nothing of the state before the first swaps); idle share = (slowest - fastest rank) / step time
    balance = None
    if world > 1:
        mine = torch.tensor([sum(t["ms"] for t in timed if t["kind"] == "pass")], device="cuda", dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [float(x.item()) for x in allr]
        balance = {"pass_ms_per_rank": [round(v, 3) for v in per_rank], "max_ms": max(per_rank), "min_ms": min(per_rank),
                   "idle_share_of_step": (max(per_rank) - min(per_rank)) / ms_per_step if ms_per_step else None}
    # timeline of ONE step as every rank saw it: events between the launch groups of a real step (waits at the swaps'
    # barriers land in the group that follows them); "init" = clear / basis store, "sample" = the draw
    step()
    barrier()
    sv.trace = []
    sv.mark("start")
    step()
    sv.mark("sample")
    torch.cuda.synchronize()
    tr, sv.trace = sv.trace, None
    mine = [(b[0], round(a[1].elapsed_time(b[1]), 3)) for a, b in zip(tr, tr[1:])]
    allt = [mine]
    if world > 1:
        allt = [None] * world
        dist.all_gather_object(allt, mine)
    timeline = {"labels": [x[0] for x in mine], "ms_per_rank": [[x[1] for x in r] for r in allt]}
    achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    import ctypes
    jc, jh, js = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_double(0)
    lib.qsv_jit_stats(ctypes.byref(jc), ctypes.byref(jh), ctypes.byref(js))
    jit = {"kernels_compiled": int(jc.value), "compile_cpu_s": float(js.value), "prepare_wall_s": t_prepare,
           "note": "pass-specialised kernels are generated + compiled once per distinct pass program (NVRTC, all host "
                   "cores) and cached; prepare_wall_s is paid once per circuit, outside the timed steps"}
    pass_kernel = ("qsv_jit_pass (pass-specialised register-resident tiled TMA gate pass, csrc/qsv_jit.cu)"
                   if jc.value else "qsv::k_pass_reg / k_pass (tiled TMA gate pass)")
    names = {"pass": pass_kernel,
             "diffusion": "qsv::k_diff_partial + k_diff_update (Grover diffusion)"}
    tr_bytes, tr_per_amp, tr_src = traffic_from_profile(dom, n_local)
    roofline = {"bound": "hbm", "kernel": names.get(dom, dom), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": tr_bytes,
                "traffic_note": (None if tr_bytes is None else
                                 "DRAM read+write bytes of ONE DENSE sweep (every tile touched) of this kernel, from the "
                                 "ncu --set full capture in %s scaled to 2^%d amplitudes: %.3f B/amplitude = %.3f x the "
                                 "algorithmic 32 B/amplitude (no re-reads); compare with algorithmic_bytes_per_dense_launch,"
                                 " not with the per-launch average below, which includes launches that skip zero tiles"
                                 % (tr_src, n_local, tr_per_amp, tr_per_amp / 32.0)),
                "algorithmic_bytes_per_dense_launch": 32.0 * (1 << n_local),
                "peak_source": peak_src, "launches_timed": d["launch_groups"],
                "avg_launch_ms": d["ms"] / max(1, d["launch_groups"]),
                "algorithmic_bytes_per_launch": d["bytes"] / max(1, d["launch_groups"]),
                "share_of_step": d["ms"] / ms_per_step if ms_per_step else None,
                "gates_per_launch": d["ops"] / max(1, d["launch_groups"]),
                "per_launch": [[round(t["ms"], 3), t["ops"]] + ([round(t["active"], 6)] if t["active"] < 1.0 else [])
                               for t in timed if t["kind"] == dom][:64],
                "per_launch_columns": "ms, gates, [fraction of the shard's tiles touched, if < 1: the register starts in "
                                      "a basis state and early passes skip tiles that are still zero]",
                "other_kernels": {k: {"ms_per_step": v["ms"], "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] else 0.0}
                                  for k, v in by_kind.items() if k != dom}}

    gate_pass = None
    if not args.no_gate_pass and rank == 0 and world == 1:
        gate_pass = single_gate_sweep(sv, peak)

    if args.workload == "layered" and not args.no_check:
        check = layered_check(sv, comm, n, args.depth, plan, in_index, src_pos, g)

    # ---- end to end through the public API: Circuit.run(in_v, 2) with host inputs
    in_v = in_bits
    e2e = None
    if world == 1 and (16 << n) * 2 < 170e9:
        import io
        import contextlib
        import random
        random.seed(1)
        del sv
        torch.cuda.empty_cache()
        for _ in range(2):
            with contextlib.redirect_stdout(io.StringIO()):
                circ.run(in_v, 2)
        torch.cuda.synchronize()
        k_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            with contextlib.redirect_stdout(io.StringIO()):
                out = circ.run(in_v, 2)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k_e2e
        st = lowering.last_run_stats()
        e2e = {"value": n_gates / dt, "unit": METRIC, "h2d_bytes_per_step": st["h2d_bytes"],
               "d2h_bytes_per_step": st["d2h_bytes"], "ms_per_step": dt * 1e3,
               "api": "main.circuit.Circuit.run(in_v, 2)"}

    e2e_cold = ref_grover = None
    if world == 1 and e2e is not None and args.workload == "layered" and not args.no_cold:
        e2e_cold = cold_start(n, args.depth, e2e["ms_per_step"])
        ref_grover = reference_grover7()

    if world > 1 and (16 << n_local) * 2 < 170e9 and args.workload == "layered":
        # the same public call under the one-process-per-GPU launch: Circuit.run shards the register over the
        # ranks by itself (qsv.dist.default_comm); every rank makes the call, the slowest rank's time counts
        import io
        import contextlib
        import random
        random.seed(1)
        sv = plan = timed = None
        comm.release_state()
        torch.cuda.empty_cache()
        for _ in range(2):
            with contextlib.redirect_stdout(io.StringIO()):
                circ.run(in_v, 2)
        k_e2e = max(2, min(args.steps, 5))
        if os.environ.get("QSV_BENCH_PROFILE_E2E") and rank == 0:      # where the host time of a sharded run() goes
            import cProfile
            import pstats
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(k_e2e):
                with contextlib.redirect_stdout(io.StringIO()):
                    circ.run(in_v, 2)
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
        elif os.environ.get("QSV_BENCH_PROFILE_E2E"):
            for _ in range(k_e2e):
                with contextlib.redirect_stdout(io.StringIO()):
                    circ.run(in_v, 2)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            with contextlib.redirect_stdout(io.StringIO()):
                out = circ.run(in_v, 2)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k_e2e
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        st = lowering.last_run_stats()
        e2e = {"value": n_gates * world / dt, "unit": METRIC, "circuit_gates_per_sec": n_gates / dt,
               "h2d_bytes_per_step": st["h2d_bytes"] * world, "d2h_bytes_per_step": st["d2h_bytes"] * world,
               "ms_per_step": dt * 1e3,
               "api": "main.circuit.Circuit.run(in_v, 2) on every rank (register sharded by qsv.dist.default_comm)"}
        lowering.release_sharded_states()

    # ---- the 137 GB-per-GPU point (33 qubits on one GPU / 33 + log2(N) on N; 36 qubits on 8 = the north star)
    big = None
    if args.workload == "layered" and not args.no_big and not args.qubits:
        sv = plan = timed = None
        if comm is not None:
            comm.release_state()
        lowering.release_sharded_states()
        engine._PLAN_CACHE.clear()
        torch.cuda.empty_cache()
        big = run_big_shard(args, comm, g, world, rank, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128 (f64 pairs)", "data": "synthetic",
        "config": {"workload": wl_name + ", %d gates, complex128, one sample per step" % n_gates,
                   "qubits": n, "depth": args.depth, "gates": n_gates, "seed": SEED,
                   "state_bytes_per_gpu": 16 << n_local, "passes_per_step": passes_per_step,
                   "cache": "state (%.1f GB/GPU) is far larger than the 126 MB L2; no flush needed"
                            % ((16 << n_local) / 1e9),
                   "sharding": "top %d qubits across %d GPUs" % (g, world) if world > 1 else "none",
                   "host_lowering_s": t_host, "jit": jit},
        "roofline": roofline, "gpu_launches": launches, "clocks": clocks, "e2e": e2e,
        "circuit_gates_per_sec": circuit_gps,
        "dense_start": {"value": n_gates * world / (ms_dense_per_step * 1e-3), "ms_per_step": ms_dense_per_step,
                        "note": "same steps with QSV_SPARSE_START=0: no support hint, every pass sweeps the whole state"},
        "amplitude_updates_per_sec": circuit_gps * float(1 << n),
    }
    if world > 1:
        line["config"]["value_definition"] = (
            "value = gates x GPUs / s: units all ranks processed (a gate of the %d-qubit circuit updates all %d "
            "shards of 2^%d amplitudes); circuit_gates_per_sec = %d gates / step time is the raw rate of the "
            "sharded circuit" % (n, world, n_local, n_gates))
    if e2e_cold is not None:
        line["e2e_cold"] = e2e_cold
    if ref_grover is not None:
        line["reference_grover7"] = ref_grover
    if gate_pass is not None:
        line["gate_pass"] = gate_pass
    if swap is not None:
        line["swap"] = swap
    if balance is not None:
        line["rank_balance"] = balance
    line["step_timeline"] = timeline
    if check is not None:
        line["check"] = check
    if big is not None:
        key = ("north_star_36q" if big["qubits"] == 36 else
               "one_gpu_33q" if world == 1 else "big_shard_%dq" % big["qubits"])
        line[key] = big
        line["big_shard_key"] = key
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.cpu_qubits, threads=args.cpu_threads)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def single_gate_sweep(sv, peak):
    """One gate per pass, the figure the north star's '>= 70 % of HBM roofline per gate pass' refers to:
    algorithmic bytes / CUDA-event time of that single launch."""
    import numpy as np
    import torch
    from qsv.schedule import Op
    s = 2 ** -0.5
    Hm = np.array([[s, s], [s, -s]], dtype=np.complex128)
    n = sv.n_local
    N = 1 << n
    cases = [("H q0 (lsb)", Op("dense", (0,), (), matrix=Hm), 32 * N),
             ("H q%d (mid)" % (n // 2), Op("dense", (n // 2,), (), matrix=Hm), 32 * N),
             ("H q%d (msb)" % (n - 1), Op("dense", (n - 1,), (), matrix=Hm), 32 * N),
             ("X q%d" % (n - 2), Op("mcx", (n - 2,), ()), 32 * N),
             ("CNot c%d t%d" % (n - 1, 3), Op("mcx", (3,), (n - 1,)), 16 * N),
             ("CZ %d,%d" % (n - 1, n - 2), Op("phase", (), (n - 1, n - 2), scalar=-1 + 0j), 8 * N),
             ("phase q%d" % (n - 1), Op("phase", (), (n - 1,), scalar=complex(s, s)), 16 * N)]
    out = []
    for label, op, nbytes in cases:
        plan = sv.prepare([op])
        for _ in range(2):
            sv.execute(plan)
        ms = sv.time_passes(plan, repeats=5)
        t = sum(ms) / len(ms)
        gbs = nbytes / (t * 1e-3) / 1e9
        out.append({"gate": label, "ms": t, "algorithmic_GBps": gbs, "frac_of_peak": gbs / peak})
    return out


if __name__ == "__main__":
    main()
