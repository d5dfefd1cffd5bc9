#!/bin/bash
# shared-memory bank conflicts of the full-sweep passes, default thread-bit mapping vs QSV_JIT_BANKSEARCH=1
out=gpurun_out
M=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum
cmd="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gate-pass --no-big --no-check --no-cold"
$cmd > $out/r02v_plain.log 2>&1 || exit 1
for bs in 0 1; do
  QSV_JIT_BANKSEARCH=$bs ncu --metrics $M --clock-control none -k regex:qsv_jit_pass -s 51 -c 9 --csv --log-file $out/r02v_banks_$bs.csv $cmd > $out/r02v_ncu_$bs.log 2>&1
  echo "banksearch=$bs rc=$?"
done
