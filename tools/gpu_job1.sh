#!/bin/bash
# round 2, GPU visit 1: parity tests with the row-scaled relative metric, phase profile, executed-flop counters
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -s 2>&1 | grep -E "worst relative|passed|failed|Error|error|assert" | tail -80 | tee gpurun_out/pytest_gpu_r2a.log
timeout 300 python tools/time_kino.py 2>&1 | tail -6 | tee gpurun_out/time_base.txt
HIPPOPT_B200_LIB=$PWD/hippopt_b200/variants/libhb_phase.so timeout 300 python tools/phase_profile.py 2>&1 | tail -30 | tee gpurun_out/phase_profile.txt
timeout 900 ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'kino_|reduce_f' -s 9 -c 3 --csv --log-file gpurun_out/flops_r2a.csv python tools/time_kino.py > gpurun_out/flops_run.log 2>&1
tail -3 gpurun_out/flops_run.log
