#!/bin/bash
echo "== lbfgs, ref options, the reference's guess 100 N / mass (53.6 kg), 256 plans"
timeout 900 python tools/check_periodic_step.py -b 256 -s -i 400 --ref-options --lbfgs --fz 1.8657 2>&1 | tail -3
echo "== lbfgs, tol 1e-6 (no acceptable level), fz=100/mass, 64 plans"
timeout 900 python tools/check_periodic_step.py -b 64 -s -i 600 --lbfgs --fz 1.8657 -t 1e-6 2>&1 | tail -3
