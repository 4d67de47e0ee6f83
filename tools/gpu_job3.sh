#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -6
HB_DEBUG_SCHED=1 timeout 300 python tools/time_kino.py 2>&1 | grep -E "packed|hess|jac" | head -6
HB_SWEEP_UNTYPED=1 HB_DEBUG_SCHED=1 timeout 300 python tools/time_kino.py 2>&1 | grep -E "packed|f\+g\+grad\+jac\+hess|^hess" | head -4
