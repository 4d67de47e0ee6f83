"""Developer check: weight-shift OCP -- initial and final states are two pose-finder solutions with the same
feet and a CoM reference displaced sideways; the kinodynamic OCP (final-state constraint on) has to move the
robot from one to the other (hippopt_b200.workloads.transfer_problem).  Guess = linear interpolation.
usage: check_transfer.py [-n HORIZON] [-d SHIFT_M] [-i MAX_ITER] [-v]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import KinoEvaluator, PoseEvaluator  # noqa: E402
from hippopt_b200.ipsolver import BatchedInteriorPoint, OptiFailure  # noqa: E402
from hippopt_b200.kino_layout import COM, NPT, NZ, KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import pose_batch, transfer_problem  # noqa: E402

d = torch.device("cuda:0")
N = int(sys.argv[sys.argv.index("-n") + 1]) if "-n" in sys.argv else 10
shift = float(sys.argv[sys.argv.index("-d") + 1]) if "-d" in sys.argv else 0.03
iters = int(sys.argv[sys.argv.index("-i") + 1]) if "-i" in sys.argv else 300
B0 = 64
model = synthetic_ergocub()
pev = PoseEvaluator(model)
po_p = pev.layout.po
xa, pa, _, _ = pose_batch(pev.layout, model, B0, seed=1, noise=0.02)
pb = pa.copy()
pb[:, po_p.ref + po_p.ST_COM + 1] += shift  # CoM reference displaced along y
lb, ub = pev.bounds(pa)
ip = BatchedInteriorPoint(pev, tol=1e-8, max_iter=300)
A = ip.solve(torch.tensor(xa, device=d), torch.tensor(pa, device=d), lb, ub)
Bsol = ip.solve(A.values.clone(), torch.tensor(pb, device=d), lb, ub)  # warm start from pose A
ok = (A.success & Bsol.success).cpu().numpy()
a, b = A.values.cpu().numpy()[ok], Bsol.values.cpu().numpy()[ok]
print(f"poses: {int(ok.sum())}/{B0} pairs; CoM moved by {np.abs(b[:, 78:81] - a[:, 78:81]).max(axis=0)}; feet moved by "
      f"{max(np.abs(b[:, 6 * i:6 * i + 3] - a[:, 6 * i:6 * i + 3]).max() for i in range(NPT)):.1e}")

ev = KinoEvaluator(model, KinoSettings(horizon=N, final_state_constraint=True))
lay = ev.layout
pk, x0 = transfer_problem(lay, model, a, b)
nB = a.shape[0]
lbk, ubk = lay.bounds(pk)
g0 = ev.eval(4, torch.tensor(x0, device=d), torch.tensor(pk, device=d))["g"].cpu().numpy()
viol = np.maximum(lbk - g0, 0) + np.maximum(g0 - ubk, 0)
print(f"interpolated guess: max constraint violation {viol.max():.2e}")
tol = float(sys.argv[sys.argv.index("-t") + 1]) if "-t" in sys.argv else 1e-6
dc = float(sys.argv[sys.argv.index("-c") + 1]) if "-c" in sys.argv else 1e-9
sol = BatchedInteriorPoint(ev, tol=tol, max_iter=iters, verbose="-v" in sys.argv, kkt="stage", delta_c=dc, mu_init=1e-3)
t0 = time.perf_counter()
try:
    res = sol.solve(torch.tensor(x0, device=d), torch.tensor(pk, device=d), lbk, ubk)
    torch.cuda.synchronize()
    okk = res.success.cpu().numpy()
    xs = res.values.cpu().numpy()[okk]
    gs = ev.eval(4, res.values, torch.tensor(pk, device=d))["g"].cpu().numpy()
    vs = (np.maximum(lbk - gs, 0) + np.maximum(gs - ubk, 0))[okk]
    com_y = xs[:, [NZ * k + COM + 1 for k in range(N)]]
    print(f"weight-shift OCP (N={N}, shift {shift} m): {int(okk.sum())}/{nB} converged to {tol:g}, iterations median "
          f"{int(res.iterations[res.success].median()) if okk.any() else -1}, {time.perf_counter() - t0:.1f} s; constraint violation "
          f"{vs.max() if okk.any() else float('nan'):.1e}; CoM y of instance 0 over the horizon: "
          f"{np.array2string(com_y[0], precision=4) if okk.any() else '-'}")
except OptiFailure as e:
    print(f"weight-shift OCP: {e} ({time.perf_counter() - t0:.1f} s)")
