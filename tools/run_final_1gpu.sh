#!/bin/bash
# Final 1-GPU campaign: GPU tests, smoke, headline bench, QFT / Grover workloads, ncu launch list + one full capture.
tag=${1:-r02f}
out=gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log )
( timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log )
( timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err )
( timeout 600 python bench.py --workload qft --no-cpu-baseline --no-gate-pass > $out/${tag}_qft.json 2> $out/${tag}_qft.err; echo "qft rc=$?" >> $out/${tag}_qft.err )
( timeout 600 python bench.py --workload grover --no-cpu-baseline --no-gate-pass > $out/${tag}_grover.json 2> $out/${tag}_grover.err; echo "grover rc=$?" >> $out/${tag}_grover.err )
( timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_reference_arm.json 2> $out/${tag}_reference_arm.err )
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gate-pass --no-big --no-check --no-cold"
$cmd > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:qsv_jit_pass -s 81 -c 3 -o $out/${tag}_prof $cmd > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_pytest.log; tail -2 $out/${tag}_smoke.log; tail -c 300 $out/${tag}_bench.err; tail -c 200 $out/${tag}_qft.err; tail -c 200 $out/${tag}_grover.err
