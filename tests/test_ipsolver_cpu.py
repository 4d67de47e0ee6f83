"""Host logic of the batched interior-point driver (hippopt_b200/ipsolver.py) on the CPU:
the driver only needs an object with the evaluator's interface, so a small torch evaluator (test code, below) stands
in for the CUDA one and the known-answer problem of IPOPT's own documentation (Hock-Schittkowski 71) pins the
conventions the reference relies on -- CCS value order of jac_g / upper-triangular hess_l, IPOPT's sign of lam_g
(`opti_solver.py:522-537` hands these multipliers out as constraint_multipliers), termination options
(`main_periodic_step.py:111-134`), `OptiFailure` (`opti_solver.py:28-37`)."""
import numpy as np
import pytest
import torch

from hippopt_b200.evaluator import F, G, GRAD_F, HESS_L, JAC_G
from hippopt_b200.ipsolver import BatchedInteriorPoint, OptiFailure

# HS071: min x0 x3 (x0 + x1 + x2) + x2  s.t.  x0 x1 x2 x3 >= 25,  sum x^2 = 40,  1 <= x <= 5
HS071_X = np.array([1.0, 4.74299963, 3.82114998, 1.37940829])
HS071_F = 17.0140173
HS071_LAM = np.array([-0.55229366, 0.16146857])  # IPOPT's lam_g for the two general constraints


class TorchEvaluator:
    """Evaluator interface (eval / jac_sparsity / hess_sparsity / n_x / m) over torch.func derivatives, dense
    patterns.  Variable bounds are rows of g, as in CasADi's Opti without simple-bound detection."""

    def __init__(self, f, g, n, m, scale_f=1.0):
        self.f, self.g, self.n_x, self.m, self.scale_f = f, g, n, m, scale_f
        self.n_p = 1
        self._hu = np.triu_indices(n)  # (row, col) pairs with row <= col ...
        order = np.lexsort((self._hu[0], self._hu[1]))  # ... sorted by column, then row: CCS
        self._hr, self._hc = self._hu[0][order], self._hu[1][order]

    def jac_sparsity(self):
        return np.arange(0, self.n_x * self.m + 1, self.m), np.tile(np.arange(self.m), self.n_x)

    def hess_sparsity(self):
        colind = np.concatenate([[0], np.cumsum(np.arange(1, self.n_x + 1))])
        return colind, self._hr

    def eval(self, mask, x, p, lam, sigma):
        fun = lambda v: self.scale_f * self.f(v)  # noqa: E731
        out = {}
        if mask & F:
            out["f"] = torch.vmap(fun)(x)
        if mask & GRAD_F:
            out["grad_f"] = torch.vmap(torch.func.grad(fun))(x)
        if mask & G:
            out["g"] = torch.vmap(self.g)(x)
        if mask & JAC_G:
            J = torch.vmap(torch.func.jacrev(self.g))(x)  # (B, m, n)
            out["jac"] = J.transpose(1, 2).reshape(x.shape[0], -1)  # column after column
        if mask & HESS_L:
            lag = lambda v, l, s: s * fun(v) + (l * self.g(v)).sum()  # noqa: E731
            H = torch.vmap(torch.func.hessian(lag))(x, lam, sigma)
            out["hess"] = H[:, self._hr, self._hc]
        return out


def hs071():
    f = lambda v: v[0] * v[3] * (v[0] + v[1] + v[2]) + v[2]  # noqa: E731
    g = lambda v: torch.cat([torch.stack([v.prod(), (v * v).sum()]), v])  # noqa: E731
    lb = np.array([25.0, 40.0, 1, 1, 1, 1])
    ub = np.array([np.inf, 40.0, 5, 5, 5, 5])
    return f, g, lb, ub


def starts():
    return torch.tensor([[1.0, 5.0, 5.0, 1.0], [2.0, 3.0, 4.0, 2.0], [3.0, 3.0, 3.0, 3.0]], dtype=torch.float64)


def test_hs071_known_answer():
    f, g, lb, ub = hs071()
    ev = TorchEvaluator(f, g, 4, 6)
    x0 = starts()
    out = BatchedInteriorPoint(ev, tol=1e-9).solve(x0, torch.zeros((3, 1), dtype=torch.float64), lb, ub)
    assert bool(out.success.all()) and not bool(out.acceptable.any())
    assert out.values.numpy() == pytest.approx(np.tile(HS071_X, (3, 1)), abs=2e-7)
    assert out.cost_value.numpy() == pytest.approx(HS071_F, abs=1e-6)
    lam = out.constraint_multipliers.numpy()
    assert lam[:, :2] == pytest.approx(np.tile(HS071_LAM, (3, 1)), abs=1e-6)
    # x0 sits on its lower bound (negative multiplier, IPOPT's convention: z_U - z_L), the other bound rows are inactive
    assert (lam[:, 2] < -1.0).all() and np.abs(lam[:, 3:]).max() < 1e-6
    assert int(out.iterations.max()) < 40 and float(out.kkt_error.max()) <= 1e-9


def test_ipopt_termination_options_on_hs071():
    f, g, lb, ub = hs071()
    ev = TorchEvaluator(f, g, 4, 6)
    x0, p = starts(), torch.zeros((3, 1), dtype=torch.float64)
    tight = BatchedInteriorPoint(ev, tol=1e-9).solve(x0, p, lb, ub)
    # unreachable desired level: the acceptable level ends the solve, earlier and looser
    loose = BatchedInteriorPoint(ev, ipopt_options={"tol": 1e-300, "acceptable_tol": 1e-3, "acceptable_iter": 2,
                                                    "max_iter": 100}).solve(x0, p, lb, ub)
    assert bool(loose.success.all()) and bool(loose.acceptable.all())
    assert int(loose.iterations.max()) < int(tight.iterations.max())
    assert float(loose.kkt_error.max()) <= 1e-3 and np.abs(loose.values.numpy() - HS071_X).max() < 1e-2
    # acceptable_iter counts CONSECUTIVE iterations: asking for more of them takes longer
    longer = BatchedInteriorPoint(ev, ipopt_options={"tol": 1e-300, "acceptable_tol": 1e-3, "acceptable_iter": 5,
                                                     "max_iter": 100}).solve(x0, p, lb, ub)
    assert (longer.iterations >= loose.iterations + 3).all()
    # a side condition of the desired level that cannot hold (IPOPT: ALL of them must) -> runs to max_iter -> OptiFailure
    with pytest.raises(OptiFailure, match="no instance reached"):
        BatchedInteriorPoint(ev, tol=1e-6, max_iter=25, ipopt_options={"constr_viol_tol": -1.0}).solve(x0, p, lb, ub)
    # ... and with generous side conditions nothing changes
    same = BatchedInteriorPoint(ev, tol=1e-9, ipopt_options={"dual_inf_tol": 1.0, "constr_viol_tol": 1e-4,
                                                             "compl_inf_tol": 1e-4}).solve(x0, p, lb, ub)
    assert torch.equal(same.iterations, tight.iterations)
    # an acceptable_obj_change_tol of zero can never be met while the objective still moves
    stuck = BatchedInteriorPoint(ev, tol=1e-9, ipopt_options={"acceptable_tol": 1e3, "acceptable_iter": 1,
                                                              "acceptable_obj_change_tol": 0.0}).solve(x0, p, lb, ub)
    assert not bool(stuck.acceptable.any()) and torch.equal(stuck.iterations, tight.iterations)


def test_gradient_based_objective_scaling_reports_unscaled_values():
    f, g, lb, ub = hs071()
    ev = TorchEvaluator(f, g, 4, 6, scale_f=1e4)  # |grad f(x0)| ~ 1e5: IPOPT scales the objective by 100 / that
    x0, p = starts(), torch.zeros((3, 1), dtype=torch.float64)
    out = BatchedInteriorPoint(ev, tol=1e-9, ipopt_options={"nlp_scaling_method": "gradient-based"}).solve(x0, p, lb, ub)
    assert bool(out.success.all())
    assert out.values.numpy() == pytest.approx(np.tile(HS071_X, (3, 1)), abs=1e-6)
    assert out.cost_value.numpy() == pytest.approx(1e4 * HS071_F, rel=1e-8)
    assert out.constraint_multipliers.numpy()[:, :2] == pytest.approx(1e4 * np.tile(HS071_LAM, (3, 1)), rel=1e-5)


def test_opti_failure_and_give_up():
    f, g, lb, ub = hs071()
    ev = TorchEvaluator(f, g, 4, 6)
    with pytest.raises(OptiFailure):
        BatchedInteriorPoint(ev, tol=1e-12, max_iter=2).solve(starts(), torch.zeros((3, 1), dtype=torch.float64), lb, ub)
    # an infeasible problem (x0 <= 5 but x0 >= 6 as a second row): line searches fail, instances are given up, no hang
    g2 = lambda v: torch.cat([g(v), v[:1]])  # noqa: E731
    ev2 = TorchEvaluator(f, g2, 4, 7)
    with pytest.raises(OptiFailure):
        BatchedInteriorPoint(ev2, tol=1e-8, max_iter=60).solve(starts(), torch.zeros((3, 1), dtype=torch.float64),
                                                                np.append(lb, 6.0), np.append(ub, 7.0))
    # instances must share the equality / inequality structure (one pattern for the batch)
    lbb = np.tile(lb, (3, 1))
    lbb[1, 0] = np.inf
    with pytest.raises(ValueError, match="share"):
        BatchedInteriorPoint(ev).solve(starts(), torch.zeros((3, 1), dtype=torch.float64), lbb, np.tile(ub, (3, 1)))


def test_opti_callback_saves_an_intermediate_solution():
    """/root/reference/test/test_optimization_problem.py:270-283: an infeasible problem (0 <= x <= 1, x^2 == 10) does
    not raise when a callback criterion saved an iterate; without one it does (:257-267)."""
    from hippopt_b200 import opti_callback

    f = lambda v: (v * v).sum() * 0.0  # noqa: E731  (the reference's problem has no cost)
    g = lambda v: torch.cat([v, v * v])  # noqa: E731
    ev = TorchEvaluator(f, g, 1, 2)
    lb, ub = np.array([0.0, 10.0]), np.array([1.0, 10.0])
    x0, p = torch.tensor([[0.5], [0.9]], dtype=torch.float64), torch.zeros((2, 1), dtype=torch.float64)
    with pytest.raises(OptiFailure, match="Opti failed to solve the problem. Message"):
        BatchedInteriorPoint(ev, tol=1e-8, max_iter=40).solve(x0, p, lb, ub)
    crit = opti_callback.BestCost() | opti_callback.BestPrimalInfeasibility()
    out = BatchedInteriorPoint(ev, tol=1e-8, max_iter=40, callback_criterion=crit).solve(x0, p, lb, ub)
    assert not bool(out.success.any()) and (out.callback_iteration >= 0).all()
    # what comes back is the iterate with the smallest violation seen, x close to its upper bound
    assert (out.values > 0.9).all() and (out.values <= 1.0 + 1e-6).all()
    # a criterion that is never satisfied saves nothing: the failure says so, in the reference's words
    never = opti_callback.AcceptablePrimalInfeasibility(1e-12)
    with pytest.raises(OptiFailure, match="the callback did not manage to save an intermediate solution"):
        BatchedInteriorPoint(ev, tol=1e-8, max_iter=40, callback_criterion=never).solve(x0, p, lb, ub)
    with pytest.raises(TypeError):
        opti_callback.BestCost() | 3
    with pytest.raises(TypeError):
        opti_callback.BestCost() & "x"


def test_planner_criterion_per_instance():
    """`BestCost() & AcceptablePrimalInfeasibility(tol)` (humanoid_kinodynamic/planner.py:57-63): each instance keeps
    its own best cost; converged instances return their solution, cut-off ones the best acceptable iterate."""
    from hippopt_b200 import opti_callback

    f, g, lb, ub = hs071()
    ev = TorchEvaluator(f, g, 4, 6)
    x0, p = starts(), torch.zeros((3, 1), dtype=torch.float64)
    full = BatchedInteriorPoint(ev, tol=1e-9).solve(x0, p, lb, ub)
    cut = int(full.iterations.min())  # the fastest instance just converges, the others are cut off
    crit = opti_callback.BestCost() & opti_callback.AcceptablePrimalInfeasibility(1e-1)
    out = BatchedInteriorPoint(ev, tol=1e-9, max_iter=cut + 1, callback_criterion=crit).solve(x0, p, lb, ub)
    ok = out.success.numpy()
    assert ok.any() and not ok.all()
    saved = out.callback_iteration >= 0  # (a cut-off instance whose violation never got below 0.1 has nothing saved)
    assert (out.callback_iteration[out.success] == -1).all() and bool(saved.any()) and not bool((saved & out.success).any())
    assert out.values.numpy()[ok] == pytest.approx(np.tile(HS071_X, (int(ok.sum()), 1)), abs=1e-6)
    # saved iterates respect the acceptable violation and carry their own cost
    xs = out.values[saved]
    gv = torch.vmap(g)(xs).numpy()
    assert (np.maximum(lb - gv, 0) + np.maximum(gv - ub, 0)).max() < 1e-1
    assert out.cost_value[saved].numpy() == pytest.approx(torch.vmap(f)(xs).numpy(), rel=1e-12)
    assert torch.equal(crit.lhs.best_cost[saved], out.cost_value[saved])  # per-instance state of the criterion


def test_speculative_hessian_shifts_take_the_same_steps():
    """Several shifts of the regularisation sequence factored in one batched solve (the GPU default for the stage
    backend) choose the shift the one-at-a-time loop would have chosen."""
    f, g, lb, ub = hs071()
    ev = TorchEvaluator(lambda v: -f(v), g, 4, 6)  # maximise: indefinite Hessians, shifts on most iterations
    x0, p = starts(), torch.zeros((3, 1), dtype=torch.float64)
    one = BatchedInteriorPoint(ev, tol=1e-8, delta_c=0.0)
    a = one.solve(x0, p, lb, ub)
    spec = BatchedInteriorPoint(ev, tol=1e-8, delta_c=0.0)
    spec.spec_wave = 12  # 3 instances -> 4 shifts per solve
    b = spec.solve(x0, p, lb, ub)
    assert bool(a.success.all()) and torch.equal(a.iterations, b.iterations)
    assert (a.values - b.values).abs().max() < 1e-9 and (a.constraint_multipliers - b.constraint_multipliers).abs().max() < 1e-7
    assert b.evaluations == a.evaluations


@pytest.mark.parametrize("form", ["upper", "lower", "two_sided"])
def test_one_sided_inequalities(form):
    """min (x0-2)^2 + (x1-2)^2 s.t. x0 + x1 <= 2 written with an upper bound only (what Opti's `<=` produces), a
    lower bound only, and both: the slack push must use the bound that exists (advisor finding, round 1)."""
    f = lambda v: ((v - 2.0) ** 2).sum()  # noqa: E731
    if form == "lower":
        g, lb, ub = (lambda v: -(v[0] + v[1]).reshape(1)), [-2.0], [np.inf]  # noqa: E731
    else:
        g, lb, ub = (lambda v: (v[0] + v[1]).reshape(1)), [-np.inf if form == "upper" else -50.0], [2.0]  # noqa: E731
    ev = TorchEvaluator(f, g, 2, 1)
    ip = BatchedInteriorPoint(ev, tol=1e-9, max_iter=60)
    x0 = torch.tensor([[0.0, 0.0], [3.0, -1.0]], dtype=torch.float64)
    out = ip.solve(x0, torch.zeros((2, 1), dtype=torch.float64), np.array(lb), np.array(ub))
    assert bool(out.success.all())
    assert torch.allclose(out.values, torch.ones((2, 2), dtype=torch.float64), atol=1e-6)
    # active constraint: lam_g = +2 on an upper bound, -2 on a lower bound (IPOPT's sign)
    assert torch.allclose(out.constraint_multipliers, torch.full((2, 1), -2.0 if form == "lower" else 2.0, dtype=torch.float64), atol=1e-5)


def test_stage_backend_needs_a_stage_structure():
    ev = TorchEvaluator(lambda v: (v * v).sum(), lambda v: v[:1], 2, 1)
    ip = BatchedInteriorPoint(ev, kkt="stage")
    with pytest.raises(ValueError, match="multiple-shooting layout"):
        ip.solve(torch.zeros((1, 2), dtype=torch.float64), torch.zeros((1, 1), dtype=torch.float64), np.array([0.0]), np.array([0.0]))


def test_limited_memory_mode_reaches_the_known_answer():
    """hessian_approximation = limited-memory (the reference's setting for the kinodynamic planners,
    main_periodic_step.py:116): hess_l is never requested, the L-BFGS matrix takes its place."""
    f, g, lb, ub = hs071()

    class NoHessian(TorchEvaluator):
        def eval(self, mask, x, p, lam, sigma):
            assert not mask & HESS_L, "limited-memory mode must not ask for hess_l"
            return super().eval(mask, x, p, lam, sigma)

    ev = NoHessian(f, g, 4, 6)
    out = BatchedInteriorPoint(ev, tol=1e-8, max_iter=200, ipopt_options={"hessian_approximation": "limited-memory"}).solve(
        starts(), torch.zeros((3, 1), dtype=torch.float64), lb, ub)
    assert bool(out.success.all())
    assert out.values.numpy() == pytest.approx(np.tile(HS071_X, (3, 1)), abs=1e-5)
    assert out.cost_value.numpy() == pytest.approx(HS071_F, abs=1e-6)
    assert out.constraint_multipliers[:, :2].numpy() == pytest.approx(np.tile(HS071_LAM, (3, 1)), abs=1e-4)


def test_limited_memory_matrix_is_the_bfgs_matrix():
    """compact representation against the textbook recursive BFGS update started from sigma I"""
    from hippopt_b200.ipsolver import LimitedMemory

    rng = np.random.default_rng(0)
    n, k = 7, 4
    lm = LimitedMemory(1, n, k, "cpu")
    A = rng.normal(size=(n, n))
    A = A @ A.T + n * np.eye(n)  # SPD "true" Hessian: y = A s guarantees s^T y > 0
    pairs = []
    for _ in range(6):  # more pairs than the history holds: the oldest are dropped
        s = rng.normal(size=n)
        y = A @ s
        lm.update(torch.ones(1, dtype=torch.bool), torch.tensor(s[None]), torch.tensor(y[None]))
        pairs.append((s, y))
    pairs = pairs[-k:]
    sigma = float(lm.sigma[0])
    assert sigma == pytest.approx(pairs[-1][0] @ pairs[-1][1] / (pairs[-1][0] @ pairs[-1][0]))
    Bm = sigma * np.eye(n)
    for s, y in pairs:
        Bs = Bm @ s
        Bm = Bm - np.outer(Bs, Bs) / (s @ Bs) + np.outer(y, y) / (y @ s)
    v = rng.normal(size=(1, n))
    got = lm.apply(torch.tensor([0]), torch.tensor(v)).numpy()
    assert got == pytest.approx(v @ Bm.T, rel=1e-10)


def test_restoration_phase_reduces_the_violation_and_hands_back():
    """The (simplified) feasibility restoration: Levenberg-Marquardt steps on the constraint violation through the
    same KKT back end, accepted on decrease of the violation alone, left at required_infeasibility_reduction x the
    violation at entry.  IPOPT's `start_with_resto` begins in it: HS071 from infeasible starts must still reach the
    known answer, and the first accepted iterates must reduce the violation."""
    f, g, lb, ub = hs071()
    ev = TorchEvaluator(f, g, 4, 6)
    x0 = starts()
    p = torch.zeros((3, 1), dtype=torch.float64)
    ip = BatchedInteriorPoint(ev, tol=1e-8, max_iter=200, ipopt_options={"start_with_resto": "yes",
                                                                        "required_infeasibility_reduction": 0.5})
    out = ip.solve(x0, p, lb, ub)
    assert ip.restoration_entries == 3  # all three starts violate sum x^2 = 40
    assert bool(out.success.all())
    assert out.values.numpy() == pytest.approx(np.tile(HS071_X, (3, 1)), abs=2e-6)
    # one iteration only: still in restoration, the violation of the equality row went down
    one = BatchedInteriorPoint(ev, tol=1e-8, max_iter=1, ipopt_options={"start_with_resto": "yes"},
                               callback_criterion=None)
    try:
        one.solve(x0, p, lb, ub)
    except OptiFailure:
        pass
    plain = BatchedInteriorPoint(ev, tol=1e-8, max_iter=200).solve(x0, p, lb, ub)
    assert (out.values - plain.values).abs().max() < 1e-5
