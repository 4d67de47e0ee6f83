"""Small workloads for compute-sanitizer (memcheck / racecheck / initcheck): every kernel once, tiny batch."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import ALL, F, G, JAC_G, HostPipeline, KinoEvaluator, PoseEvaluator, ToyEvaluator  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.kkt import lu_factor, lu_solve  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch, pose_batch  # noqa: E402

d = torch.device("cuda:0")
model = synthetic_ergocub()
for st in (KinoSettings(horizon=3, final_state_constraint=True, periodicity_constraint=True),
           KinoSettings(horizon=2, terrain="smooth_steps", n_terrain_params=10)):
    ev = KinoEvaluator(model, st)
    x, p, lam, sigma = kino_batch(ev.layout, model, 3, seed=1, noise=0.1)
    t = [torch.tensor(a, device=d) for a in (x, p, lam, sigma)]
    for mask in (ALL, F | G, JAC_G):
        ev.eval(mask, *t)
    torch.cuda.synchronize()
pev = PoseEvaluator(model)
t = [torch.tensor(a, device=d) for a in pose_batch(pev.layout, model, 5, seed=1)]
pev.eval(ALL, *t)
tev = ToyEvaluator(6, "trapezoid")
tev.eval(ALL, torch.zeros((2, tev.n_x), dtype=torch.float64, device=d), torch.ones((2, tev.n_p), dtype=torch.float64, device=d),
         torch.ones((2, tev.m), dtype=torch.float64, device=d), torch.ones(2, dtype=torch.float64, device=d))
A = torch.randn(3, 37, 37, dtype=torch.float64, device=d)
Fm, piv, info = lu_factor(A.clone())
lu_solve(Fm, piv, torch.randn(3, 37, 5, dtype=torch.float64, device=d))
A = torch.randn(2, 337, 337, dtype=torch.float64, device=d)  # the KKT stage-block size: 22 panels, both register variants
Fm, piv, info = lu_factor(A.clone())
lu_solve(Fm, piv, torch.randn(2, 337, 40, dtype=torch.float64, device=d))
# interpolation kernel: a plan with swings, per-instance phases, both outputs, a warp with fewer than three items
from hippopt_b200.initial_guess import periodic_step_phases  # noqa: E402
from hippopt_b200.interpolators import humanoid_state_interpolator  # noqa: E402

B = 7
ph = periodic_step_phases(np.linspace(0.1, 0.3, B), 1.0)
s0, s1 = torch.randn(B, 105, dtype=torch.float64, device=d), torch.randn(B, 105, dtype=torch.float64, device=d)
for q in (s0, s1):
    q[:, 75:79] /= q[:, 75:79].norm(dim=1, keepdim=True)
xo = torch.zeros((B, 189 * 12 + 6), dtype=torch.float64, device=d)
humanoid_state_interpolator(s0, s1, ph, 10, 0.1, x_out=xo, knot0=1)
torch.cuda.synchronize()
print("sanitize workload done")
