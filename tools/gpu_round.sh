#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list, ncu full capture of the top kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kino_kin_kernel -s 4 -c 2 -f -o gpurun_out/prof_kin \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_kin.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kino_contact_kernel -s 4 -c 1 -f -o gpurun_out/prof_contact \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_contact.log 2>&1
# rows f1-f3: plan set-up rate + interpolation kernel timing, periodic-step plans with the reference's IPOPT options
timeout 300 python tools/check_periodic_step.py -k 2>&1 | tail -4 | tee gpurun_out/setup.txt
timeout 400 python tools/check_periodic_step.py -b 64 -s -i 300 --ref-options --fz 1.2258 2>&1 | tail -2 | tee gpurun_out/periodic_step.txt
ls -la gpurun_out
