#!/bin/bash
# N-GPU round: sharded parity (tests/dist_check.py) and the bench line at N GPUs.   usage: tools/run_ngpu.sh N tag [bench args]
N=$1; tag=$2; shift 2
out=gpurun_out
port=$((29500 + RANDOM % 400))
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tests/dist_check.py > $out/${tag}_dist_check_${N}gpu.log 2>&1; echo "dist_check rc=$?" >> $out/${tag}_dist_check_${N}gpu.log )
grep -E "PASS|FAIL|rc=|all passed" $out/${tag}_dist_check_${N}gpu.log | tail -40
port=$((29500 + RANDOM % 400))
( timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err; echo "bench rc=$?" >> $out/${tag}_bench_${N}gpu.err )
tail -c 1500 $out/${tag}_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/${tag}_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "dense", d["dense_start"]["ms_per_step"], "check", d.get("check"))
    print("swap", d.get("swap"))
    k=d.get("big_shard_key"); print(k, json.dumps(d.get(k)) if k else None)
except Exception as e:
    print("no bench line:", e)
PY
