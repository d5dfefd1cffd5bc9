#!/bin/bash
# usage: tools/ab_env.sh <tag> "<ENV=.. ENV=..>" ["<ENV..>" ...]  - headline bench (no big block, no CPU arm) once per setting
out=gpurun_out
tag=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gate-pass --no-big --no-cold \
     > $out/${tag}_$i.json 2> $out/${tag}_$i.err
  echo "[$envs] rc=$?"
  python - $out/${tag}_$i.json <<'P'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print(round(d["value"]), round(d["ms_per_step"], 2), "dense", round(d["dense_start"]["ms_per_step"], 2), "ok", d["check"]["ok"],
              "e2e", round(d["e2e"]["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 3), "prep", round(d["config"]["jit"]["prepare_wall_s"], 2),
              [p[0] for p in d["roofline"]["per_launch"]])
P
done
