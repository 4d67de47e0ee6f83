"""Developer check: weight-shift OCP by continuation.  The interior-point driver converges reliably from almost
feasible starts (tools/check_standing.py) but not from the interpolated guess of a 30 mm CoM transfer
(tools/check_transfer.py).  Here the final state is moved in small steps, each OCP warm-started from the previous
solution.   usage: check_transfer_continuation.py [-n HORIZON] [-d SHIFT_M] [-k STEPS] [-t TOL]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import KinoEvaluator, PoseEvaluator  # noqa: E402
from hippopt_b200.ipsolver import BatchedInteriorPoint, OptiFailure  # noqa: E402
from hippopt_b200.kino_layout import COM, NZ, KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import pose_batch, transfer_problem  # noqa: E402

d = torch.device("cuda:0")
N = int(sys.argv[sys.argv.index("-n") + 1]) if "-n" in sys.argv else 10
shift = float(sys.argv[sys.argv.index("-d") + 1]) if "-d" in sys.argv else 0.03
K = int(sys.argv[sys.argv.index("-k") + 1]) if "-k" in sys.argv else 10
tol = float(sys.argv[sys.argv.index("-t") + 1]) if "-t" in sys.argv else 1e-5
B0 = 64
model = synthetic_ergocub()
pev = PoseEvaluator(model)
po_p = pev.layout.po
xa, pa, _, _ = pose_batch(pev.layout, model, B0, seed=1, noise=0.02)
lb, ub = pev.bounds(pa)
pip = BatchedInteriorPoint(pev, tol=1e-8, max_iter=300)
A = pip.solve(torch.tensor(xa, device=d), torch.tensor(pa, device=d), lb, ub)
alive = A.success.cpu().numpy()
a = A.values.cpu().numpy()
ev = KinoEvaluator(model, KinoSettings(horizon=N, final_state_constraint=True))
lay = ev.layout
pose_t, x_prev = A.values.clone(), None
t0 = time.perf_counter()
total_iters = 0
com_y = None
for step in range(1, K + 1):
    pb = pa.copy()
    pb[:, po_p.ref + po_p.ST_COM + 1] += shift * step / K
    Bt = pip.solve(pose_t, torch.tensor(pb, device=d), lb, ub)
    alive &= Bt.success.cpu().numpy()
    pose_t = Bt.values.clone()
    pk, x0 = transfer_problem(lay, model, a, pose_t.cpu().numpy())
    lbk, ubk = lay.bounds(pk)
    guess = torch.tensor(x0, device=d) if x_prev is None else x_prev
    sol = BatchedInteriorPoint(ev, tol=tol, max_iter=200, kkt="stage", delta_c=1e-9, mu_init=1e-3)
    try:
        res = sol.solve(guess, torch.tensor(pk, device=d), lbk, ubk)
    except OptiFailure as e:
        print(f"step {step}: {e}")
        break
    okk = res.success.cpu().numpy()
    alive &= okk
    x_prev = res.values.clone()
    total_iters += int(res.iterations.max())
    com_y = res.values[:, [NZ * k + COM + 1 for k in range(N)]].cpu().numpy()
    print(f"step {step:2d}/{K}: CoM shift {1e3 * shift * step / K:5.1f} mm, {int(okk.sum())}/{B0} converged this step, "
          f"{int(alive.sum())} alive, iterations median {int(res.iterations.median())}")
torch.cuda.synchronize()
print(f"weight-shift by continuation (N={N}, {1e3 * shift:.0f} mm in {K} steps): {int(alive.sum())}/{B0} instances converged at every "
      f"step, {time.perf_counter() - t0:.1f} s")
if alive.any() and com_y is not None:
    i = int(np.nonzero(alive)[0][0])
    print("CoM y over the horizon, instance", i, np.array2string(com_y[i], precision=4))
