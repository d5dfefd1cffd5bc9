for w in 1 2 4 8 16 32 64 128; do echo "== waves $w"; QSV_JIT_WAVES=$w CASE=b N=30 REPS=5 timeout 200 python tools/microbench_floor.py 2>&1 | grep -v "^$"; done
