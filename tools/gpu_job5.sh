#!/bin/bash
for v in fake2 fka fkb fkd; do
  echo "== $v"
  HIPPOPT_B200_LIB=$PWD/hippopt_b200/variants/libhb_$v.so timeout 300 python tools/time_kino.py 2>&1 | grep -E "^f\+g\+grad\+jac\+hess|^hess"
done
