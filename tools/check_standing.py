"""Developer check of rows f1 + f2 end to end: pose-finder solves (dense KKT) produce statically balanced
poses; each becomes the initial state, the reference and the initial guess of a kinodynamic OCP ("keep
standing", hippopt_b200.workloads.standing_problem), which is then solved with the stage-wise KKT sweep.
Reports convergence honestly.   usage: check_standing.py [-n HORIZON] [-i MAX_ITER] [-v]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import KinoEvaluator, PoseEvaluator  # noqa: E402
from hippopt_b200.ipsolver import BatchedInteriorPoint, OptiFailure  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import pose_batch, standing_problem  # noqa: E402

d = torch.device("cuda:0")
verbose = "-v" in sys.argv
N = int(sys.argv[sys.argv.index("-n") + 1]) if "-n" in sys.argv else 4
iters = int(sys.argv[sys.argv.index("-i") + 1]) if "-i" in sys.argv else 300
model = synthetic_ergocub()

pev = PoseEvaluator(model)
B0 = 256
x, p, lam, sigma = pose_batch(pev.layout, model, B0, seed=1, noise=0.02)
lb, ub = pev.bounds(p)
t0 = time.perf_counter()
out = BatchedInteriorPoint(pev, tol=1e-8, max_iter=300).solve(torch.tensor(x, device=d), torch.tensor(p, device=d), lb, ub)
torch.cuda.synchronize()
pose = out.values.cpu().numpy()[out.success.cpu().numpy()]
B = pose.shape[0]
print(f"pose finder: {B}/{B0} converged in {time.perf_counter() - t0:.1f} s")

periodic = "-p" in sys.argv  # BASELINE config 4's structure: final-state constraint + periodicity rows
ev = KinoEvaluator(model, KinoSettings(horizon=N, final_state_constraint=periodic, periodicity_constraint=periodic))
lay = ev.layout
pk, x0 = standing_problem(lay, model, pose)
lbk, ubk = lay.bounds(pk)
g0 = ev.eval(4, torch.tensor(x0, device=d), torch.tensor(pk, device=d))["g"].cpu().numpy()
viol = np.maximum(lbk - g0, 0) + np.maximum(g0 - ubk, 0)
print(f"standing guess: max constraint violation {viol.max():.2e}")
sol = BatchedInteriorPoint(ev, tol=1e-6, max_iter=iters, verbose=verbose, kkt="stage", delta_c=1e-9, mu_init=1e-3)
t0 = time.perf_counter()
try:
    res = sol.solve(torch.tensor(x0, device=d), torch.tensor(pk, device=d), lbk, ubk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    okk = res.success.cpu().numpy()
    gs = ev.eval(4, res.values, torch.tensor(pk, device=d))["g"].cpu().numpy()
    vs = (np.maximum(lbk - gs, 0) + np.maximum(gs - ubk, 0))[okk]
    print(f"standing OCP (N={N}, n_x={lay.n_x}, m={lay.m}): {int(okk.sum())}/{B} converged to 1e-6 in <= {iters} iterations "
          f"(median {int(res.iterations[res.success].median()) if okk.any() else -1}), {dt:.1f} s wall, {sol.kkt_seconds:.1f} s in "
          f"the KKT sweep, {res.evaluations} batched evaluations; constraint violation of the solutions {vs.max():.1e}, "
          f"cost median {res.cost_value[res.success].median().item():.3e}")
except OptiFailure as e:
    print(f"standing OCP: {e} ({time.perf_counter() - t0:.1f} s, {sol.kkt_seconds:.1f} s in the KKT sweep)")
