#!/bin/bash
echo "== shipped (17)"; timeout 300 python tools/time_kino.py 2>&1 | grep -E "^f\+g\+grad\+jac\+hess"
for v in scat8 scat12 scat25; do
  echo "== $v"
  HIPPOPT_B200_LIB=$PWD/hippopt_b200/variants/libhb_$v.so timeout 300 python tools/time_kino.py 2>&1 | grep -E "^f\+g\+grad\+jac\+hess"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden or edge or against_oracle" 2>&1 | tail -2
