#!/bin/bash
set -x
mkdir -p gpurun_out
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,sm__cycles_elapsed.avg,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem
timeout 600 ncu --metrics $M --clock-control none -k regex:'kino_|reduce_f' -s 9 -c 3 --csv --log-file gpurun_out/counters.csv python tools/time_kino.py > gpurun_out/counters_run.log 2>&1
timeout 900 python bench.py --steps 100 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench_r2.json
tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference_r2.json
