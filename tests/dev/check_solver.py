"""Developer check: batched interior-point solves of the toy OCP (closed form known) and the pose finder."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import PoseEvaluator, ToyEvaluator  # noqa: E402
from hippopt_b200.ipsolver import BatchedInteriorPoint  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import pose_batch  # noqa: E402
from oracle import toy  # noqa: E402

d = torch.device("cuda:0")
verbose = "-v" in sys.argv
# ---- toy OCP (test_multiple_shooting.py:253-353), randomised g, x0, v0
N, dt, B = 100, 0.01, 64
ev = ToyEvaluator(N, "euler", dt)
rng = np.random.default_rng(0)
p = np.stack([rng.uniform(-12, -8, B), rng.uniform(0.5, 1.5, B), rng.uniform(-1, 1, B)], axis=1)
p[0] = [-9.81, 1.0, 0.0]
lb, ub = ev.bounds(p)
sol = BatchedInteriorPoint(ev, tol=1e-8, verbose=verbose)
x0 = torch.zeros((B, ev.n_x), dtype=torch.float64, device=d)
t0 = time.perf_counter()
out = sol.solve(x0, torch.tensor(p, device=d), lb, ub)
torch.cuda.synchronize()
dt_s = time.perf_counter() - t0
exact = np.stack([toy.closed_form_solution(N, dt, *p[i]) for i in range(B)])
err = np.abs(out.values.cpu().numpy() - exact).max(axis=1)
print(f"toy: {int(out.success.sum())}/{B} converged, iters max {int(out.iterations.max())}, max |x - closed form| {err.max():.2e}, "
      f"{B / dt_s:.1f} solves/s ({dt_s:.2f} s, {out.evaluations} batched evals)")
# ---- pose finder
model = synthetic_ergocub()
ev = PoseEvaluator(model)
B = 256
x, p, lam, sigma = pose_batch(ev.layout, model, B, seed=1, noise=0.02)
lb, ub = ev.bounds(p)
sol = BatchedInteriorPoint(ev, tol=1e-8, max_iter=300, verbose=verbose)
t0 = time.perf_counter()
try:
    out = sol.solve(torch.tensor(x, device=d), torch.tensor(p, device=d), lb, ub)
    torch.cuda.synchronize()
    dt_s = time.perf_counter() - t0
    print(f"pose: {int(out.success.sum())}/{B} converged, iters median {int(out.iterations.median())} max {int(out.iterations.max())}, "
          f"kkt err max over converged {out.kkt_error[out.success].max().item():.2e}, {B / dt_s:.1f} solves/s ({dt_s:.2f} s, "
          f"{out.evaluations} batched evals), cost median {out.cost_value.median().item():.4f}")
except Exception as e:
    print("pose failed:", e)
