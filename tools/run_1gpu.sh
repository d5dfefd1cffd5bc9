#!/bin/bash
# 1-GPU round: GPU test suite, smoke, headline bench (with the in-run check and the 33-qubit block)
tag=${1:-r02b}
out=gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log )
( timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log )
( timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err )
tail -3 $out/${tag}_pytest.log; tail -2 $out/${tag}_smoke.log; tail -c 600 $out/${tag}_bench.err
