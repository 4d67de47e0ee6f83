"""Developer check of row f3 (+ f1/f2 on top of it): the set-up of main_periodic_step.py for a batch of step
lengths -- keyframe poses by the batched pose finder, initial guess by the device interpolator -- and, with -s, an
attempt to solve the resulting periodic-step OCPs (config 4's structure) with the stage-wise KKT sweep.
Reports what happens, converged or not.   usage: check_periodic_step.py [-b BATCH] [-n HORIZON] [-s] [-i MAX_ITER] [-v] [--fz FORCE] [--lmin L] [--lmax L] [-m MU_INIT] [-t TOL] [--ref-options]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import KinoEvaluator, PoseEvaluator  # noqa: E402
from hippopt_b200.initial_guess import periodic_step_guess  # noqa: E402
from hippopt_b200.ipsolver import BatchedInteriorPoint, OptiFailure  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402


def arg(flag, default):
    return type(default)(sys.argv[sys.argv.index(flag) + 1]) if flag in sys.argv else default


B, N, iters = arg("-b", 256), arg("-n", 30), arg("-i", 300)
Lmin, Lmax = arg("--lmin", 0.1), arg("--lmax", 0.3)
d = torch.device("cuda:0")
model = synthetic_ergocub()
pev = PoseEvaluator(model)
ev = KinoEvaluator(model, KinoSettings(horizon=N, final_state_constraint=True, periodicity_constraint=True))
lay = ev.layout
L = np.random.default_rng(5).uniform(Lmin, Lmax, B)
periodic_step_guess(model, pev, ev, L[:8])  # warm-up (library load, allocator)
torch.cuda.synchronize()
t0 = time.perf_counter()
gs = periodic_step_guess(model, pev, ev, L, force_z=arg("--fz", 100.0))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ok = gs.ok.cpu().numpy()
print(f"periodic-step set-up for {B} instances (step length U({Lmin}, {Lmax}) m, horizon {N}): {int(ok.sum())}/{B} instances "
      f"have all three keyframe poses ({3 * B} pose-finder solves, median {int(gs.pose_iterations.median())} iterations); "
      f"poses + interpolation + parameters in {dt:.2f} s = {B / dt:.0f} set-ups/s")
P = torch.tensor(gs.parameters, device=d)
lb, ub = lay.bounds(gs.parameters)
g0 = ev.eval(4, gs.x0, P)["g"].cpu().numpy()
viol = np.maximum(lb - g0, 0) + np.maximum(g0 - ub, 0)
print(f"guess: max constraint violation median {np.median(viol.max(axis=1)):.3f}, worst {viol.max():.3f} "
      f"(zero velocities, planned feet: the guess is not meant to be feasible)")
if "-k" in sys.argv:  # the interpolation kernel alone, at config 4's batch (4096 instances x 30 knots)
    from hippopt_b200.initial_guess import periodic_step_phases
    from hippopt_b200.interpolators import humanoid_state_interpolator

    nb = 4096
    rep = (nb + B - 1) // B
    k0 = gs.keyframes[0].repeat(rep, 1)[:nb].contiguous()
    k1 = gs.keyframes[1].repeat(rep, 1)[:nb].contiguous()
    ph = periodic_step_phases(np.resize(L, nb), N * 0.1)
    xo = torch.zeros((nb, lay.n_x), dtype=torch.float64, device=d)
    for both in (False, True):
        for _ in range(3):
            humanoid_state_interpolator(k0, k1, ph, N, 0.1, x_out=xo if both else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(20):
            humanoid_state_interpolator(k0, k1, ph, N, 0.1, x_out=xo if both else None)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 20
        ms = e0.elapsed_time(e1) / 20
        written = nb * N * (105 + (81 if both else 0)) * 8
        print(f"interpolator, {nb} instances x {N} points, states{' + decision vector' if both else ''}: {ms * 1e3:.0f} us per "
              f"call on the stream ({wall * 1e3:.2f} ms wall with the host-side schedule and phase upload), "
              f"{written / 1e6:.0f} MB written")
if "-s" in sys.argv:
    sel = np.nonzero(ok)[0]
    # --ref-options: the termination / scaling options main_periodic_step.py:111-134 hands to IPOPT
    ref_opts = {"tol": 1e-3, "dual_inf_tol": 1000.0, "compl_inf_tol": 1e-2, "constr_viol_tol": 1e-4, "acceptable_tol": 10,
                "acceptable_iter": 2, "acceptable_compl_inf_tol": 1000.0, "acceptable_obj_change_tol": 1e0,
                "nlp_scaling_method": "gradient-based"} if "--ref-options" in sys.argv else None
    if "--lbfgs" in sys.argv:  # hessian_approximation = limited-memory, the reference's own setting (:116)
        ref_opts = dict(ref_opts or {}, hessian_approximation="limited-memory", limited_memory_max_history=arg("--history", 6))
    if "--no-resto" in sys.argv:
        ref_opts = dict(ref_opts or {}, hb_restoration=False)
    sol = BatchedInteriorPoint(ev, tol=arg("-t", 1e-6), max_iter=iters, verbose="-v" in sys.argv, kkt="stage",
                               delta_c=1e-9, mu_init=arg("-m", 1e-1), ipopt_options=ref_opts)
    if "--callback" in sys.argv:  # the planner's criterion (humanoid_kinodynamic/planner.py:57-63)
        from hippopt_b200 import opti_callback

        sol.callback_criterion = opti_callback.BestCost() & opti_callback.AcceptablePrimalInfeasibility(arg("--callback", 1e-2))
    t0 = time.perf_counter()
    try:
        res = sol.solve(gs.x0[sel], P[sel], lb[sel], ub[sel])
        torch.cuda.synchronize()
        n_ok = int(res.success.sum())
        print(f"periodic-step OCP (n_x={lay.n_x}, m={lay.m}): {n_ok}/{len(sel)} converged ({int(res.acceptable.sum())} of them at "
              f"IPOPT's acceptable level{', reference options' if ref_opts else ''}) in <= {iters} iterations "
              f"(median {int(res.iterations[res.success].median()) if n_ok else -1}), {time.perf_counter() - t0:.1f} s, "
              f"KKT error median {res.kkt_error.median().item():.2e}, best {res.kkt_error.min().item():.2e}")
        if sol.callback_criterion is not None:
            used = (res.callback_iteration >= 0).cpu().numpy()
            if used.any():
                gu = ev.eval(4, res.values, P[sel])["g"].cpu().numpy()
                vu = (np.maximum(lb[sel] - gu, 0) + np.maximum(gu - ub[sel], 0))[used]
                print(f"  callback: {int(used.sum())} of the {len(sel) - n_ok} unconverged plans return an iterate saved by BestCost & "
                      f"AcceptablePrimalInfeasibility (iterations {res.callback_iteration[res.callback_iteration >= 0].tolist()}, "
                      f"constraint violation max {vu.max():.1e})")
            else:
                print(f"  callback: none of the {len(sel) - n_ok} unconverged plans has a saved iterate")
        if n_ok:
            okk = res.success.cpu().numpy()
            gsol = ev.eval(4, res.values, P[sel])["g"].cpu().numpy()
            vs = (np.maximum(lb[sel] - gsol, 0) + np.maximum(gsol - ub[sel], 0))[okk]
            xs = res.values.cpu().numpy()[okk]
            zk = xs[:, :189 * N].reshape(-1, N, 189)
            lift = zk[:, :, 6 + 2].max(axis=1)  # height of the first left contact point over the plan
            moved = zk[:, -1, 6] - zk[:, 0, 6]
            print(f"  solutions: constraint violation max {vs.max():.1e}; left foot travels {np.median(moved):.3f} m (median; planned "
                  f"{np.median(L[sel][okk]):.3f}) and is lifted by {np.median(lift) * 1e3:.1f} mm (median, max over the knots); "
                  f"cost median {res.cost_value[res.success].median().item():.3e}; {res.evaluations} batched evaluations, "
                  f"{sol.kkt_seconds:.1f} s in the KKT sweep; {sol.restoration_entries} entries into the restoration phase")
    except OptiFailure as e:
        print(f"periodic-step OCP: {e} ({time.perf_counter() - t0:.1f} s)")
