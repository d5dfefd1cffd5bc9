#!/bin/bash
# A/B of the tile's low-bit count (run length of the TMA copies): fewer low bits -> more free tile positions per pass
out=gpurun_out
for L in 6 5 4; do
  QSV_LOW_BITS=$L timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gate-pass --no-big \
     > $out/r02n_L$L.json 2> $out/r02n_L$L.err
  echo "L=$L rc=$?"
  python - $out/r02n_L$L.json <<'P'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print(d["value"], d["ms_per_step"], d["dense_start"]["ms_per_step"], d["check"]["ok"], [p[0] for p in d["roofline"]["per_launch"]])
P
done
