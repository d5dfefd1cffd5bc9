#!/bin/bash
# QFT (config 4) and Grover (config 5) bench lines at N GPUs.   usage: tools/run_ngpu_workloads.sh N tag [qft_qubits] [grover_qubits]
N=$1; tag=$2; QQ=${3:-0}; GQ=${4:-0}
out=gpurun_out
run() {
  name=$1; shift
  port=$((29500 + RANDOM % 400))
  if [ "$N" = "1" ]; then
    ( timeout 900 python bench.py --gpus 1 "$@" > $out/${tag}_${name}_${N}gpu.json 2> $out/${tag}_${name}_${N}gpu.err; echo "rc=$?" >> $out/${tag}_${name}_${N}gpu.err )
  else
    ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > $out/${tag}_${name}_${N}gpu.json 2> $out/${tag}_${name}_${N}gpu.err; echo "rc=$?" >> $out/${tag}_${name}_${N}gpu.err )
  fi
  tail -c 300 $out/${tag}_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/${tag}_${name}_${N}gpu.json") if l.startswith("{")][-1])
    print("$name", d["config"]["qubits"], "q value", d["value"], "circuit gates/s", d["circuit_gates_per_sec"], "ms", d["ms_per_step"], "check", d.get("check"), "roofline", d["roofline"]["kernel"][:30], d["roofline"]["frac"])
except Exception as e:
    print("$name: no bench line:", e)
PY
}
qa=""; [ "$QQ" != "0" ] && qa="--qubits $QQ"
ga=""; [ "$GQ" != "0" ] && ga="--qubits $GQ"
run qft --workload qft $qa --no-cpu-baseline --no-gate-pass --no-big
run grover --workload grover $ga --steps 3 --no-cpu-baseline --no-gate-pass --no-big
