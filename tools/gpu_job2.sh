#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2b.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_r2b.log
timeout 300 python tools/callback_latency.py 2>&1 | tail -2 | tee gpurun_out/callback_latency.txt
timeout 300 python tools/time_kino.py 2>&1 | tail -5 | tee gpurun_out/time_r2b.txt
