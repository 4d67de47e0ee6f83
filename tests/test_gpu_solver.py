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
