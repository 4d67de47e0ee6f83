"""Cycles per phase of the two kinodynamic evaluation kernels (developer tool).

Needs the instrumented build: tools/build_variant.sh phase -DHB_PHASE_CLOCK, then
HIPPOPT_B200_LIB=hippopt_b200/variants/libhb_phase.so python tools/phase_profile.py
The instrumentation adds a warp barrier and an atomic per phase, so the shares are what to read, not the sum."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from hippopt_b200 import _capi  # noqa: E402
from hippopt_b200.evaluator import ALL, KinoEvaluator  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch  # noqa: E402

KIN = ["inputs staged", "FK by depth", "body quantities, sums, frames", "g rows, costs", "composite moments",
       "direction tangents", "Jacobian columns staged", "Jacobian scatter + grad_f", "seeds", "primal adjoint pass",
       "cross terms", "packed tangent sweep", "root chain rule", "Hessian scatter"]
CON = ["inputs staged", "g rows, per-point terms", "least-squares rows", "f, grad_f", "Jacobian values + scatter",
       "Hessian terms", "Hessian scatter"]

model = synthetic_ergocub()
ev = KinoEvaluator(model, KinoSettings(horizon=30))
B = 1024
x, p, lam, sigma = kino_batch(ev.layout, model, B, seed=2)
d = torch.device("cuda:0")
X, P, L, S = (torch.tensor(a, device=d) for a in (x, p, lam, sigma))
lib = _capi.lib()
buf = (ctypes.c_ulonglong * 64)()
for _ in range(3):
    ev.eval(ALL, X, P, L, S)
lib.hb_debug_phase_read(buf)
reps = 5
for _ in range(reps):
    ev.eval(ALL, X, P, L, S)
lib.hb_debug_phase_read(buf)
warps = B * 30 * reps
for name, labels, off in (("kino_kin_kernel<true>", KIN, 0), ("kino_contact_kernel<0>", CON, 32)):
    vals = [buf[off + i] / warps for i in range(len(labels))]
    tot = sum(vals)
    print(f"== {name}: {tot:.0f} cycles per warp (knot-eval)")
    for lab, v in zip(labels, vals):
        print(f"  {v:9.0f}  {100 * v / tot:5.1f}%  {lab}")
