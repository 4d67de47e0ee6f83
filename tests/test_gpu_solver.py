"""Row f1: the batched interior-point driver, with the evaluator in the loop, reaches the
solution the reference's own end-to-end test asserts for config 1
(/root/reference/test/test_multiple_shooting.py:253-353: Euler recursion, foo = 0 / 5 / 6)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_toy_ocp_solves_reach_the_reference_solution(built_library):
    from hippopt_b200.evaluator import ToyEvaluator
    from hippopt_b200.ipsolver import BatchedInteriorPoint
    from oracle import toy

    N, dt, B = 100, 0.01, 12
    ev = ToyEvaluator(N, "euler", dt)
    rng = np.random.default_rng(0)
    p = np.stack([rng.uniform(-12, -8, B), rng.uniform(0.5, 1.5, B), rng.uniform(-1, 1, B)], axis=1)
    p[0] = [-9.81, 1.0, 0.0]  # the reference test's own numbers
    lb, ub = ev.bounds(p)
    dev = torch.device("cuda:0")
    out = BatchedInteriorPoint(ev, tol=1e-8).solve(torch.zeros((B, ev.n_x), dtype=torch.float64, device=dev),
                                                   torch.tensor(p, device=dev), lb, ub)
    assert bool(out.success.all())
    x = out.values.cpu().numpy()
    exact = np.stack([toy.closed_form_solution(N, dt, *p[i]) for i in range(B)])
    assert np.abs(x - exact).max() < 1e-7  # pytest.approx's default 1e-6 relative in the reference test
    assert out.cost_value.cpu().numpy() == pytest.approx(3 * (98 * 25.0 + 36.0) * np.ones(B), rel=1e-8)
    # IPOPT sign convention: multipliers of the active lower bounds foo >= 5 are negative (-2 foo = -10)
    lam = out.constraint_multipliers.cpu().numpy()
    o = 2 * (N - 1) + 2 + 2 * N
    assert lam[:, o:o + 3 * (N - 2)] == pytest.approx(-10.0, abs=1e-5)
    assert int(out.iterations.max()) < 30


def test_ipopt_termination_and_scaling_options(built_library):
    """The IPOPT options the reference's mains pass (`main_periodic_step.py:111-134`): acceptable-level
    termination, the extra desired-level tolerances and gradient-based objective scaling, with IPOPT's meaning."""
    from hippopt_b200.evaluator import ToyEvaluator
    from hippopt_b200.ipsolver import BatchedInteriorPoint

    N, dt, B = 100, 0.01, 6
    ev = ToyEvaluator(N, "euler", dt)
    p = np.tile([-9.81, 1.0, 0.0], (B, 1))
    lb, ub = ev.bounds(p)
    dev = torch.device("cuda:0")
    x0, P = torch.zeros((B, ev.n_x), dtype=torch.float64, device=dev), torch.tensor(p, device=dev)
    tight = BatchedInteriorPoint(ev, tol=1e-8).solve(x0, P, lb, ub)
    assert not bool(tight.acceptable.any())
    # an unreachable desired tolerance: stops at the acceptable level after acceptable_iter iterations in a row
    loose = BatchedInteriorPoint(ev, ipopt_options={"tol": 1e-300, "acceptable_tol": 1e-2, "acceptable_iter": 2,
                                                    "max_iter": 60}).solve(x0, P, lb, ub)
    assert bool(loose.success.all()) and bool(loose.acceptable.all())
    assert int(loose.iterations.max()) < int(tight.iterations.max())
    assert float(loose.kkt_error.max()) <= 1e-2 and (loose.values - tight.values).abs().max() < 1e-1
    # a desired-level side condition that cannot hold keeps an otherwise converged instance running
    never = BatchedInteriorPoint(ev, tol=1e-8, max_iter=40, ipopt_options={"constr_viol_tol": -1.0, "acceptable_tol": 1e-6,
                                                                         "acceptable_iter": 3})
    out = never.solve(x0, P, lb, ub)
    assert bool(out.acceptable.all())  # ... until the acceptable level takes it
    # gradient-based scaling (the objective gradient at x0 = 0 is zero here, at a shifted start it is ~ 6 N): same
    # solution, and cost / multipliers are reported unscaled
    x1 = torch.full_like(x0, 40.0)
    a = BatchedInteriorPoint(ev, tol=1e-8).solve(x1, P, lb, ub)
    b = BatchedInteriorPoint(ev, tol=1e-8, ipopt_options={"nlp_scaling_method": "gradient-based"}).solve(x1, P, lb, ub)
    assert bool(a.success.all()) and bool(b.success.all())
    assert (a.values - b.values).abs().max() < 1e-6
    assert b.cost_value.cpu().numpy() == pytest.approx(a.cost_value.cpu().numpy(), rel=1e-8)
    o = 2 * (N - 1) + 2 + 2 * N
    assert b.constraint_multipliers[:, o:o + 3 * (N - 2)].cpu().numpy() == pytest.approx(-10.0, abs=1e-4)


def test_solver_reports_failure_like_the_reference(built_library):
    """`OptiFailure` when nothing converges (base/opti_solver.py:28-37, test_optimization_problem.py:257-267)."""
    from hippopt_b200.evaluator import ToyEvaluator
    from hippopt_b200.ipsolver import BatchedInteriorPoint, OptiFailure

    ev = ToyEvaluator(10, "euler", 0.01)
    p = np.array([[-9.81, 1.0, 0.0]])
    lb, ub = ev.bounds(p)
    dev = torch.device("cuda:0")
    with pytest.raises(OptiFailure):
        BatchedInteriorPoint(ev, tol=1e-12, max_iter=2).solve(torch.zeros((1, ev.n_x), dtype=torch.float64, device=dev),
                                                              torch.tensor(p, device=dev), lb, ub)


def test_callback_criterion_with_the_cuda_evaluator(built_library):
    """`OptiSolver(callback_criterion=...)` (opti_solver.py:451-520) on device tensors: a solve cut off after a few
    iterations returns, per instance, the iterate its criterion saved instead of raising."""
    from hippopt_b200 import opti_callback
    from hippopt_b200.evaluator import G, ToyEvaluator
    from hippopt_b200.ipsolver import BatchedInteriorPoint

    ev = ToyEvaluator(20, "euler", 0.01)
    B = 5
    p = np.tile([-9.81, 1.0, 0.0], (B, 1))
    lb, ub = ev.bounds(p)
    dev = torch.device("cuda:0")
    x0, P = torch.zeros((B, ev.n_x), dtype=torch.float64, device=dev), torch.tensor(p, device=dev)
    crit = opti_callback.BestCost() | opti_callback.BestPrimalInfeasibility()
    out = BatchedInteriorPoint(ev, tol=1e-12, max_iter=4, callback_criterion=crit).solve(x0, P, lb, ub)
    assert not bool(out.success.any()) and (out.callback_iteration >= 0).all() and (out.callback_iteration <= 3).all()
    g = ev.eval(G, out.values, P)["g"].cpu().numpy()
    assert np.isfinite(g).all() and torch.isfinite(out.cost_value).all()
    assert torch.equal(crit.rhs.best_primal_infeasibility.isfinite(), torch.ones(B, dtype=torch.bool, device=dev))


@pytest.mark.parametrize("periodic", [False, True])  # True: config 4's structure (final state + periodicity rows)
def test_standing_ocp_solves_with_the_stage_kkt(model, built_library, periodic):
    """Rows f1 + f2 end to end on the real problem: pose-finder solutions (dense KKT) become "keep standing"
    kinodynamic OCPs, solved with the stage-wise KKT sweep and the batched LU kernels.  No reference
    trajectory exists for this (IPOPT is not available), so the checks are intrinsic: the guess is feasible
    by construction, and what the solver returns satisfies the constraints and its own KKT test."""
    from hippopt_b200.evaluator import G, KinoEvaluator, PoseEvaluator
    from hippopt_b200.ipsolver import BatchedInteriorPoint
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import pose_batch, standing_problem

    dev = torch.device("cuda:0")
    pev = PoseEvaluator(model)
    x, p, lam, sigma = pose_batch(pev.layout, model, 96, seed=1, noise=0.02)
    lb, ub = pev.bounds(p)
    out = BatchedInteriorPoint(pev, tol=1e-8, max_iter=300).solve(torch.tensor(x, device=dev), torch.tensor(p, device=dev), lb, ub)
    pose = out.values.cpu().numpy()[out.success.cpu().numpy()]
    assert pose.shape[0] >= 90  # 96/96 with the f-type step acceptance (36 % without: profiles/r01/solver_v8.txt)
    assert int(out.iterations.max()) <= 150
    ev = KinoEvaluator(model, KinoSettings(horizon=4, final_state_constraint=periodic, periodicity_constraint=periodic))
    lay = ev.layout
    pk, x0 = standing_problem(lay, model, pose)
    lbk, ubk = lay.bounds(pk)
    P = torch.tensor(pk, device=dev)
    g0 = ev.eval(G, torch.tensor(x0, device=dev), P)["g"].cpu().numpy()
    assert (np.maximum(lbk - g0, 0) + np.maximum(g0 - ubk, 0)).max() < 1e-9  # feasible guess
    sol = BatchedInteriorPoint(ev, tol=1e-6, max_iter=300, kkt="stage", delta_c=1e-9, mu_init=1e-3)
    res = sol.solve(torch.tensor(x0, device=dev), P, lbk, ubk)
    ok = res.success.cpu().numpy()
    assert ok.sum() >= int(0.9 * pose.shape[0]), f"{int(ok.sum())} of {pose.shape[0]} standing OCPs converged"
    assert int(res.iterations[res.success].median()) <= 60
    assert float(res.kkt_error[res.success].max()) <= 1e-6
    gs = ev.eval(G, res.values, P)["g"].cpu().numpy()[ok]
    assert (np.maximum(lbk[ok] - gs, 0) + np.maximum(gs - ubk[ok], 0)).max() < 1e-5
    assert torch.isfinite(res.cost_value[res.success]).all()
    # run to run: every product with jac_g / hess_l has a fixed summation order (hb_ccs_group_mul) and the kernels
    # write every output slot once, so a second solve reproduces the first bit for bit
    again = BatchedInteriorPoint(ev, tol=1e-6, max_iter=300, kkt="stage", delta_c=1e-9, mu_init=1e-3).solve(
        torch.tensor(x0, device=dev), P, lbk, ubk)
    assert torch.equal(again.iterations, res.iterations) and torch.equal(again.values, res.values)
    assert torch.equal(again.constraint_multipliers, res.constraint_multipliers)


def test_sparse_products_match_dense(model, built_library):
    """hb_ccs_group_mul (J x, J^T lam, x^T H x) against dense matrices built from the same CCS arrays."""
    from hippopt_b200.evaluator import ALL, KinoEvaluator
    from hippopt_b200.ipsolver import SparseOps
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import kino_batch

    dev = torch.device("cuda:0")
    ev = KinoEvaluator(model, KinoSettings(horizon=3, final_state_constraint=True, periodicity_constraint=True))
    x, p, lam, sigma = kino_batch(ev.layout, model, 3, seed=6, noise=0.1)
    out = ev.eval(ALL, *(torch.tensor(a, device=dev) for a in (x, p, lam, sigma)))
    ops = SparseOps(ev.n_x, ev.m, ev.jac_sparsity(), ev.hess_sparsity(), dev)
    assert ops.native
    J, Hm = ops.dense_jac(out["jac"]), ops.dense_hess(out["hess"])
    g = torch.Generator().manual_seed(1)
    v = torch.randn((3, ev.n_x), generator=g, dtype=torch.float64).to(dev)
    w = torch.randn((3, ev.m), generator=g, dtype=torch.float64).to(dev)
    assert torch.allclose(ops.J_mul(out["jac"], v), torch.bmm(J, v[:, :, None])[:, :, 0], rtol=1e-12, atol=1e-12)
    assert torch.allclose(ops.Jt_mul(out["jac"], w), torch.bmm(J.transpose(1, 2), w[:, :, None])[:, :, 0], rtol=1e-12, atol=1e-12)
    q = (v * torch.bmm(Hm, v[:, :, None])[:, :, 0]).sum(1)
    assert torch.allclose(ops.W_quad(out["hess"], v), q, rtol=1e-12)
    assert torch.equal(ops.J_mul(out["jac"], v), ops.J_mul(out["jac"], v))


def test_b200solver_behind_the_optimization_solver_interface(model, built_library):
    """Row a30 / (b): the casadi-free solver object (hippopt_b200/plugin.py) drives a batch of standing OCPs through
    the 16-method interface and returns per-name cost values and multipliers (opti_solver.py:522-537); a second
    solve warm-started from the first one's solution (main_single_step_flat_ground.py:120-125) needs fewer iterations."""
    from hippopt_b200 import naming, plugin
    from hippopt_b200.evaluator import F, KinoEvaluator, PoseEvaluator
    from hippopt_b200.ipsolver import BatchedInteriorPoint
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import pose_batch, standing_problem

    dev = torch.device("cuda:0")
    pev = PoseEvaluator(model)
    x, p, _, _ = pose_batch(pev.layout, model, 8, seed=4, noise=0.02)
    lb, ub = pev.bounds(p)
    pose = BatchedInteriorPoint(pev, tol=1e-8, max_iter=300).solve(torch.tensor(x, device=dev), torch.tensor(p, device=dev), lb, ub)
    pose = pose.values.cpu().numpy()[pose.success.cpu().numpy()][:4]
    B = pose.shape[0]
    st = KinoSettings(horizon=4)
    ev = KinoEvaluator(model, st)
    pk, x0 = standing_problem(ev.layout, model, pose)
    opts = {"tol": 1e-6, "max_iter": 300, "mu_init": 1e-3}
    s = plugin.B200Solver(model=model, settings=st, batch=B, evaluator=ev, options_solver=opts,
                          options_plugin={"expand": True, "detect_simple_bounds": True})
    s.generate_optimization_objects({"x": x0, "p": pk})
    s.register_problem("standing")
    s.solve()
    vals, costs, mult = s.get_values(), s.get_cost_values(), s.get_constraint_multipliers()
    assert len(vals) == B and len(costs) == B and len(mult) == B
    vec = s.get_solution_vectors()
    f = ev.eval(F, torch.tensor(vec["x"], device=dev), torch.tensor(pk, device=dev))["f"].cpu().numpy()
    assert s.get_cost_value() == pytest.approx(f, rel=1e-12)
    for b in range(B):
        assert sum(costs[b].values()) == pytest.approx(f[b], rel=1e-12)       # the named costs add up to f
        assert set(mult[b]) == set(naming.constraint_rows(ev.layout))
        assert np.array_equal(mult[b]["unitary_quaternion[2]"], vec["lam_g"][b][naming.constraint_rows(ev.layout)["unitary_quaternion[2]"]])
        assert np.array_equal(vals[b]["system"][3]["kinematics"]["joints"]["positions"], vec["x"][b, 189 * 3 + 157:189 * 3 + 180])
    cold_iters = int(s._last_output.iterations.max())
    # warm start: previous solution and multipliers as the guess
    w = plugin.B200Solver(model=model, settings=st, batch=B, evaluator=ev,
                          options_solver=dict(opts, warm_start_init_point="yes", mu_init=1e-7))  # mu the cold solve ended with: tol / 10
    w.generate_optimization_objects({"x": vec["x"], "p": pk, "lam_g": vec["lam_g"]})
    w.solve()
    assert int(w._last_output.iterations.max()) <= max(3, cold_iters // 3)
    assert np.abs(w.get_solution_vectors()["x"] - vec["x"]).max() < 1e-4


def test_oracle_cache_over_hb_eval_host(model, built_library):
    """The shim's evaluation path with the real library: `HostEvaluator` (hb_eval_host, one instance, pinned
    buffers) behind the x-keyed `OracleCache` -- values equal the device-pointer path bit for bit, two launches per
    iterate."""
    from hippopt_b200 import plugin
    from hippopt_b200.evaluator import ALL, KinoEvaluator
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import kino_batch

    ev = KinoEvaluator(model, KinoSettings(horizon=5))
    x, p, lam, sigma = kino_batch(ev.layout, model, 1, seed=8, noise=0.1)
    dev = torch.device("cuda:0")
    ref = {k: v.cpu().numpy()[0] for k, v in ev.eval(ALL, *(torch.tensor(a, device=dev) for a in (x, p, lam, sigma))).items()}
    host = plugin.HostEvaluator(ev)
    host.set_parameters(p[0])
    cache = plugin.OracleCache(host, host.masks)
    for name in ("f", "grad_f", "g", "jac", "g", "f"):
        assert np.array_equal(np.ravel(cache.get(name, x[0])), np.ravel(ref[name])), name
    assert np.array_equal(cache.hess(x[0], lam[0], sigma[0]), ref["hess"])
    assert cache.launches == {"first_order": 1, "hess": 1}
    cache.get("g", x[0] + 1e-3)
    assert cache.launches["first_order"] == 2


def test_periodic_step_plans_with_the_references_own_options(model, built_library):
    """BASELINE config 4 as posed (rows f1 + f2 + f3): periodic walking step plans set up like main_periodic_step.py
    (keyframe poses, interpolated guess, planned force 100 N per point divided by the mass as the planner does) and
    solved with the options the reference hands to IPOPT (:111-134) -- hessian_approximation = limited-memory, tol 1e-3,
    acceptable level -- converge, satisfy the constraints and make the planned step."""
    from hippopt_b200.evaluator import G, KinoEvaluator, PoseEvaluator
    from hippopt_b200.initial_guess import periodic_step_guess
    from hippopt_b200.ipsolver import BatchedInteriorPoint
    from hippopt_b200.kino_layout import KinoSettings

    dev = torch.device("cuda:0")
    N, B = 30, 24
    pev = PoseEvaluator(model)
    ev = KinoEvaluator(model, KinoSettings(horizon=N, final_state_constraint=True, periodicity_constraint=True))
    L = np.random.default_rng(5).uniform(0.1, 0.3, B)
    gs = periodic_step_guess(model, pev, ev, L, force_z=100.0, mass_normalised=True)
    assert bool(gs.ok.all())
    lb, ub = ev.layout.bounds(gs.parameters)
    opts = {"tol": 1e-3, "dual_inf_tol": 1000.0, "compl_inf_tol": 1e-2, "constr_viol_tol": 1e-4, "acceptable_tol": 10,
            "acceptable_iter": 2, "acceptable_compl_inf_tol": 1000.0, "acceptable_obj_change_tol": 1e0,
            "nlp_scaling_method": "gradient-based", "max_iter": 400, "hessian_approximation": "limited-memory"}
    ip = BatchedInteriorPoint(ev, kkt="stage", delta_c=1e-9, mu_init=1e-1, ipopt_options=opts)
    P = torch.tensor(gs.parameters, device=dev)
    res = ip.solve(gs.x0, P, lb, ub)
    ok = res.success.cpu().numpy()
    assert ok.sum() >= int(0.9 * B), f"{int(ok.sum())} of {B} periodic-step plans converged"
    assert int(res.iterations[res.success].median()) <= 150
    g = ev.eval(G, res.values, P)["g"].cpu().numpy()[ok]
    assert (np.maximum(lb[ok] - g, 0) + np.maximum(g - ub[ok], 0)).max() <= 1e-4  # the reference's constr_viol_tol
    z = res.values.cpu().numpy()[ok][:, :189 * N].reshape(-1, N, 189)
    travelled = z[:, -1, 6] - z[:, 0, 6]  # x of the first left contact point, first to last knot
    assert np.abs(travelled - L[ok]).max() < 5e-3
    assert z[:, :, 8].max(axis=1).min() > 0.01  # the swing foot leaves the ground


def test_b200solver_on_the_pose_finder_template(model, built_library):
    """Config 2 behind the same 16-method interface: a batch of static pose problems, per-name cost values (the 28 named
    costs of humanoid_pose_finder/planner.py add up to the objective), multipliers per named constraint, the `state` tree."""
    from hippopt_b200 import naming, plugin
    from hippopt_b200.evaluator import F, PoseEvaluator
    from hippopt_b200.workloads import pose_batch

    dev = torch.device("cuda:0")
    pev = PoseEvaluator(model)
    B = 6
    x, p, _, _ = pose_batch(pev.layout, model, B, seed=5, noise=0.02)
    s = plugin.B200Solver(model=model, batch=B, evaluator=pev, kkt="dense", options_solver={"tol": 1e-8, "max_iter": 300})
    s.generate_optimization_objects({"x": x, "p": p})
    s.register_problem("pose")
    s.solve()
    vals, costs, mult = s.get_values(), s.get_cost_values(), s.get_constraint_multipliers()
    vec = s.get_solution_vectors()
    f = pev.eval(F, torch.tensor(vec["x"], device=dev), torch.tensor(p, device=dev))["f"].cpu().numpy()
    assert s.get_cost_value() == pytest.approx(f, rel=1e-12)
    rows = naming.constraint_rows(pev.layout)
    for b in range(B):
        assert len(costs[b]) == 28 and sum(costs[b].values()) == pytest.approx(f[b], rel=1e-11)
        assert "state.contact_points.right[3].f_regularization" in costs[b] and "frame_rotation_error" in costs[b]
        assert set(mult[b]) == set(rows) and "centroidal_momentum_dynamics" in mult[b]
        assert np.array_equal(mult[b]["joint_position_bounds"], vec["lam_g"][b][rows["joint_position_bounds"]])
        assert np.array_equal(vals[b]["state"]["kinematics"]["joints"]["positions"], vec["x"][b, 55:78])
        assert np.array_equal(vals[b]["references"]["state"]["com"], p[b, pev.layout.po.ref + 102:pev.layout.po.ref + 105])


def test_b200solver_on_the_toy_template(built_library):
    """Config 1 (the mass-falling OCP of the reference's own test) behind the 16-method interface: no name table, so the
    values are the flat vectors and the multipliers one block; the solution is the test's closed form."""
    from hippopt_b200 import plugin
    from hippopt_b200.evaluator import ToyEvaluator
    from oracle import toy

    N, dt, B = 100, 0.01, 3
    ev = ToyEvaluator(N, "euler", dt)
    p = np.tile([-9.81, 1.0, 0.0], (B, 1))
    s = plugin.B200Solver(batch=B, evaluator=ev, kkt="dense", options_solver={"tol": 1e-8})
    s.generate_optimization_objects({"x": np.zeros((B, ev.n_x)), "p": p})
    s.register_problem("mass_falling")
    s.solve()
    exact = toy.closed_form_solution(N, dt, -9.81, 1.0, 0.0)
    for b in range(B):
        assert np.abs(s.get_values()[b]["x"] - exact).max() < 1e-7
        assert s.get_cost_values()[b] == {} and list(s.get_constraint_multipliers()[b]) == ["g"]
    assert s.get_cost_value() == pytest.approx(3 * (98 * 25.0 + 36.0) * np.ones(B), rel=1e-8)


def test_solution_sensitivity_on_the_pose_finder(model, built_library):
    """Row f4, the `to_function` path: d(solution) / d(parameter) from one more KKT solve against central differences of
    complete re-solves, for a reference (desired CoM height), a physical parameter (friction) and a joint reference."""
    from hippopt_b200 import plugin
    from hippopt_b200.evaluator import PoseEvaluator
    from hippopt_b200.workloads import pose_batch

    pev = PoseEvaluator(model)
    po = pev.layout.po
    B = 4
    x, p, _, _ = pose_batch(pev.layout, model, B, seed=6, noise=0.02)
    opts = {"tol": 1e-9, "max_iter": 400}

    def solved(pp, guess=None):
        s = plugin.B200Solver(model=model, batch=B, evaluator=pev, kkt="dense", options_solver=opts)
        s.generate_optimization_objects({"x": x if guess is None else guess, "p": pp})
        s.solve()
        return s, s._last_output.success.cpu().numpy()

    s0, ok = solved(p)
    xs = s0.get_solution_vectors()["x"]
    idx = [po.ref + 104, po.mu, po.ref + 79 + 14]  # desired CoM height, static friction, a knee reference
    dx = s0.get_solution_sensitivity(idx)
    assert dx.shape == (B, pev.n_x, 3)
    for c, j in enumerate(idx):
        h = 1e-3 * max(1.0, abs(p[0, j]))  # the re-solves are exact to ~1e-9: a smaller step would amplify that noise
        pp, pm = p.copy(), p.copy()
        pp[:, j] += h
        pm[:, j] -= h
        (sp, okp), (sm, okm) = solved(pp, xs), solved(pm, xs)
        use = ok & okp & okm
        assert use.sum() >= B - 1
        fd = (sp.get_solution_vectors()["x"] - sm.get_solution_vectors()["x"])[use] / (2 * h)
        scale = max(np.abs(fd).max(), 1e-3)
        assert np.abs(dx[use][:, :, c] - fd).max() <= 5e-3 * scale, (c, np.abs(dx[use][:, :, c] - fd).max(), scale)
    assert np.abs(dx[ok][:, :, 0]).max() > 1e-2  # the pose does move with the desired CoM height


def test_solution_sensitivity_stage_backend_matches_dense(model, built_library):
    """The same derivative through the block-tridiagonal sweep (kkt='stage': fused assembly, several right-hand sides)
    and through one dense solve per instance, on standing OCPs of the kinodynamic planner."""
    from hippopt_b200 import plugin
    from hippopt_b200.evaluator import KinoEvaluator, PoseEvaluator
    from hippopt_b200.ipsolver import BatchedInteriorPoint
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.sensitivity import solution_sensitivity
    from hippopt_b200.workloads import pose_batch, standing_problem

    dev = torch.device("cuda:0")
    pev = PoseEvaluator(model)
    x, p, _, _ = pose_batch(pev.layout, model, 6, seed=4, noise=0.02)
    lb, ub = pev.bounds(p)
    pose = BatchedInteriorPoint(pev, tol=1e-8, max_iter=300).solve(torch.tensor(x, device=dev), torch.tensor(p, device=dev), lb, ub)
    pose = pose.values.cpu().numpy()[pose.success.cpu().numpy()][:3]
    st = KinoSettings(horizon=4)
    ev = KinoEvaluator(model, st)
    pk, x0 = standing_problem(ev.layout, model, pose)
    s = plugin.B200Solver(model=model, settings=st, batch=pose.shape[0], evaluator=ev,
                          options_solver={"tol": 1e-7, "max_iter": 300, "mu_init": 1e-3})
    s.generate_optimization_objects({"x": x0, "p": pk})
    s.solve()
    sv = s.get_solution_vectors()
    po = ev.layout.po
    idx = [po.dt, po.mu, po.refs0 + 55 * 2 + po.R_JR + 3]  # time step, friction, a joint reference of knot 2
    args = [torch.as_tensor(sv[k], device=dev) for k in ("x", "lam_g", "p")]
    # the standing OCP has redundant equality rows and directions its objective barely sees: both systems carry the
    # solver's own kind of regularisation (Hessian shift, -delta_c on the multiplier block)
    d_stage, _, _ = solution_sensitivity(ev, *args, ev.layout.bounds, idx, kkt="stage", delta=1e-4, delta_c=1e-8)
    d_dense, _, _ = solution_sensitivity(ev, *args, ev.layout.bounds, idx, kkt="dense", delta=1e-4, delta_c=1e-8)
    scale = d_dense.abs().amax(dim=(1, 2), keepdim=True).clamp(min=1e-6)
    assert torch.isfinite(d_stage).all() and float(scale.max()) < 1e6
    assert ((d_stage - d_dense).abs() / scale).max().item() < 1e-5
