"""Developer check: BASELINE config 3 as posed (single step on flat ground, main_single_step_flat_ground.py) for a batch
of step lengths, solved with the options that main hands to IPOPT (:107-131, limited-memory Hessian).
usage: check_single_step.py [-b BATCH] [-i MAX_ITER] [-v] [--final]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import KinoEvaluator, PoseEvaluator  # noqa: E402
from hippopt_b200.initial_guess import single_step_problem  # noqa: E402
from hippopt_b200.ipsolver import BatchedInteriorPoint, OptiFailure  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402


def arg(flag, default):
    return type(default)(sys.argv[sys.argv.index(flag) + 1]) if flag in sys.argv else default


B, iters, N = arg("-b", 64), arg("-i", 600), 30
d = torch.device("cuda:0")
model = synthetic_ergocub()
pev = PoseEvaluator(model)
ev = KinoEvaluator(model, KinoSettings(horizon=N, final_state_constraint="--final" in sys.argv))
L = np.random.default_rng(7).uniform(0.2, 0.4, B)
gs = single_step_problem(model, pev, ev, L)
P = torch.tensor(gs.parameters, device=d)
lb, ub = ev.layout.bounds(gs.parameters)
opts = {"max_iter": iters, "hessian_approximation": "limited-memory", "tol": 1e-3, "dual_inf_tol": 1000.0, "compl_inf_tol": 1e-2,
        "constr_viol_tol": 1e-4, "acceptable_tol": 1e0, "acceptable_iter": 2, "acceptable_compl_inf_tol": 1.0,
        "acceptable_obj_change_tol": 1e0, "nlp_scaling_method": "gradient-based"}
sol = BatchedInteriorPoint(ev, kkt="stage", delta_c=1e-9, mu_init=1e-1, ipopt_options=opts, verbose="-v" in sys.argv)
t0 = time.perf_counter()
try:
    res = sol.solve(gs.x0, P, lb, ub)
    torch.cuda.synchronize()
    ok = res.success.cpu().numpy()
    g = ev.eval(4, res.values, P)["g"].cpu().numpy()
    viol = (np.maximum(lb - g, 0) + np.maximum(g - ub, 0))[ok]
    z = res.values.cpu().numpy()[ok][:, :189 * N].reshape(-1, N, 189)
    print(f"single-step OCP (config 3, n_x={ev.n_x}, m={ev.m}): {int(ok.sum())}/{B} converged ({int(res.acceptable.sum())} at the "
          f"acceptable level) in a median of {int(res.iterations[res.success].median()) if ok.any() else -1} iterations, "
          f"{time.perf_counter() - t0:.1f} s; constraint violation max {viol.max() if ok.any() else float('nan'):.1e}; right foot "
          f"travels {np.median(z[:, -1, 15 * 4 + 6] - z[:, 0, 15 * 4 + 6]) if ok.any() else float('nan'):.3f} m (median; step length median "
          f"{np.median(L[ok]) if ok.any() else float('nan'):.3f}), CoM advances {np.median(z[:, -1, 180] - z[:, 0, 180]) if ok.any() else float('nan'):.3f} m; "
          f"cost median {res.cost_value[res.success].median().item() if ok.any() else float('nan'):.3e}")
except OptiFailure as e:
    print(f"single-step OCP: {e} ({time.perf_counter() - t0:.1f} s)")
