for w in 4 16 32 64 128; do echo "== waves $w"; QSV_JIT_WAVES=$w timeout 300 python bench.py --no-cpu-baseline --no-gate-pass --no-big --no-check 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['dense_start']['ms_per_step'], [x[0] for x in d['roofline']['per_launch']])
"; done
