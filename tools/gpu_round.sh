#!/bin/bash
# One GPU-box visit (round 2): parity tests, smoke, bench + reference arm, ncu launch list of the bench command,
# ncu --set full captures of the two evaluation kernels, per-kernel counters for the roofline.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kino_kin_kernel -s 4 -c 1 -f -o gpurun_out/prof_kin \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_kin.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kino_contact_kernel -s 4 -c 1 -f -o gpurun_out/prof_contact \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_contact.log 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,sm__cycles_elapsed.avg,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem
timeout 600 ncu --metrics $M --clock-control none -k regex:'kino_|reduce_f' -s 9 -c 3 --csv --log-file gpurun_out/counters.csv python tools/time_kino.py > gpurun_out/counters_run.log 2>&1
ls -la gpurun_out | tail -15
