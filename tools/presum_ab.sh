#!/bin/bash
# A/B of the sampler sums fused into the last gate pass on one GPU: step time and the step timeline of every variant
# (QSV_SUM_STATIC: buckets from generation-time constants; QSV_PRESUM_MAX_INSIDE: bucket bits allowed inside the tile;
# QSV_PRESUM=0: the sampler sweeps the state itself).   usage (on the GPU box): tools/presum_ab.sh tag
tag=${1:-ab}
out=gpurun_out
common="--steps 10 --warmup 3 --no-big --no-cold --no-gate-pass --no-cpu-baseline"
run() {
  name=$1; shift
  env "$@" python bench.py $common > $out/${tag}_presum_$name.json 2> $out/${tag}_presum_$name.err
  python - <<PY
import json
d=json.loads([l for l in open("$out/${tag}_presum_$name.json") if l.startswith("{")][-1])
t=d["step_timeline"]
print("$name", "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "last", t["ms_per_rank"][0][-3:], "sum", round(sum(t["ms_per_rank"][0]),2), "bits", d["check"].get("output_bits_summed_by_the_last_pass"), "ok", d["check"]["ok"])
PY
}
run default QSV_PRESUM=1
run general QSV_SUM_STATIC=0
if [ -z "$SHORT" ]; then
run general_in2 QSV_SUM_STATIC=0 QSV_PRESUM_MAX_INSIDE=2
run off QSV_PRESUM=0
fi
