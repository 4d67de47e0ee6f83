#!/bin/bash
set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2c.log
timeout 900 python bench.py --steps 100 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench_r2.json | cut -c1-300
tail -3 gpurun_out/bench.err
