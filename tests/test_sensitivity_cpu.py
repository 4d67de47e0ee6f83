"""Derivative of the solution with respect to parameters (hippopt_b200/sensitivity.py; the counterpart of the reference's
`OptiSolver.to_function(...).jacobian()` path, `opti_solver.py:597-638`, `humanoid_pose_finder/main_sensitivity.py:213-247`)
on the CPU: a torch evaluator with parameters stands in for the CUDA one; known closed forms and re-solved finite differences."""
import numpy as np
import pytest
import torch

from hippopt_b200.evaluator import F, G, GRAD_F, HESS_L, JAC_G
from hippopt_b200.ipsolver import BatchedInteriorPoint
from hippopt_b200.sensitivity import solution_sensitivity


class ParamEvaluator:
    """Evaluator interface over torch.func derivatives of f(x, p), g(x, p); dense patterns (test code)."""

    def __init__(self, f, g, n, m, n_p):
        self.f, self.g, self.n_x, self.m, self.n_p = f, g, n, m, n_p
        hu = np.triu_indices(n)
        order = np.lexsort((hu[0], hu[1]))
        self._hr, self._hc = hu[0][order], hu[1][order]

    def jac_sparsity(self):
        return np.arange(0, self.n_x * self.m + 1, self.m), np.tile(np.arange(self.m), self.n_x)

    def hess_sparsity(self):
        return np.concatenate([[0], np.cumsum(np.arange(1, self.n_x + 1))]), self._hr

    def eval(self, mask, x, p, lam, sigma):
        out = {}
        if mask & F:
            out["f"] = torch.vmap(self.f)(x, p)
        if mask & GRAD_F:
            out["grad_f"] = torch.vmap(torch.func.grad(self.f))(x, p)
        if mask & G:
            out["g"] = torch.vmap(self.g)(x, p)
        if mask & JAC_G:
            out["jac"] = torch.vmap(torch.func.jacrev(self.g))(x, p).transpose(1, 2).reshape(x.shape[0], -1)
        if mask & HESS_L:
            lag = lambda v, q, l, s: s * self.f(v, q) + (l * self.g(v, q)).sum()  # noqa: E731
            H = torch.vmap(torch.func.hessian(lag))(x, p, lam, sigma)
            out["hess"] = H[:, self._hr, self._hc]
        return out


def test_projection_problem_has_the_closed_form_sensitivity():
    """min 1/2 |x - c|^2, c = p[2:5], s.t. x0 + x1 = p0 (a bound that depends on p), x2 >= p1 (active / inactive)."""
    f = lambda v, q: 0.5 * ((v - q[2:5]) ** 2).sum()  # noqa: E731
    g = lambda v, q: torch.stack([v[0] + v[1], v[2]])  # noqa: E731
    ev = ParamEvaluator(f, g, 3, 2, 5)
    p = torch.tensor([[1.0, 0.5, 0.2, 0.3, 0.1],    # c2 = 0.1 < p1 = 0.5: the inequality is active, x2 = p1
                      [1.0, -2.0, 0.2, 0.3, 0.1]],  # c2 = 0.1 > p1 = -2: inactive, x2 = c2
                     dtype=torch.float64)

    def bounds(pn):
        pn = np.atleast_2d(pn)
        return np.stack([pn[:, 0], pn[:, 1]], axis=1), np.stack([pn[:, 0], np.full(len(pn), np.inf)], axis=1)

    lb, ub = bounds(p.numpy())
    out = BatchedInteriorPoint(ev, tol=1e-10).solve(torch.zeros((2, 3), dtype=torch.float64), p, lb, ub)
    assert bool(out.success.all())
    x = out.values.numpy()
    assert x[0] == pytest.approx([0.45, 0.55, 0.5], abs=1e-7) and x[1] == pytest.approx([0.45, 0.55, 0.1], abs=1e-7)
    dx, dlE, eq = solution_sensitivity(ev, out.values, out.constraint_multipliers, p, bounds, [0, 1, 2, 3, 4])
    assert eq.tolist() == [0] and dx.shape == (2, 3, 5) and dlE.shape == (2, 1, 5)
    # (x0, x1): projection on x0 + x1 = p0 of (c0, c1); x2 = max(c2, p1)
    expect_active = np.array([[0.5, 0.0, 0.5, -0.5, 0.0], [0.5, 0.0, -0.5, 0.5, 0.0], [0.0, 1.0, 0.0, 0.0, 0.0]])
    expect_inactive = np.array([[0.5, 0.0, 0.5, -0.5, 0.0], [0.5, 0.0, -0.5, 0.5, 0.0], [0.0, 0.0, 0.0, 0.0, 1.0]])
    assert dx[0].numpy() == pytest.approx(expect_active, abs=1e-6)
    assert dx[1].numpy() == pytest.approx(expect_inactive, abs=1e-6)


def test_nonlinear_problem_against_resolved_finite_differences():
    """HS071-like objective with parameters in the objective, an equality and an inequality level: the sensitivity
    against central differences of complete re-solves."""
    f = lambda v, q: v[0] * v[3] * (v[0] + v[1] + v[2]) + q[2] * v[2]  # noqa: E731
    g = lambda v, q: torch.cat([torch.stack([v.prod(), (v * v).sum()]), v])  # noqa: E731
    ev = ParamEvaluator(f, g, 4, 6, 3)

    def bounds(pn):
        pn = np.atleast_2d(pn)
        B = len(pn)
        lb = np.concatenate([pn[:, :1], pn[:, 1:2], np.ones((B, 4))], axis=1)
        ub = np.concatenate([np.full((B, 1), np.inf), pn[:, 1:2], 5.0 * np.ones((B, 4))], axis=1)
        return lb, ub

    p0 = np.array([[25.0, 40.0, 1.0], [24.0, 39.0, 1.3]])
    x0 = torch.tensor([[1.0, 5.0, 5.0, 1.0]] * 2, dtype=torch.float64)

    def solve(pn):
        lb, ub = bounds(pn)
        o = BatchedInteriorPoint(ev, tol=1e-11).solve(x0, torch.tensor(pn), lb, ub)
        assert bool(o.success.all())
        return o

    out = solve(p0)
    dx, _, _ = solution_sensitivity(ev, out.values, out.constraint_multipliers, torch.tensor(p0), bounds, [0, 1, 2])
    for c in range(3):
        h = 1e-3  # the re-solves are exact to ~1e-9: a smaller step would divide that noise by itself
        pp, pm = p0.copy(), p0.copy()
        pp[:, c] += h
        pm[:, c] -= h
        fd = (solve(pp).values - solve(pm).values).numpy() / (2 * h)
        assert dx[:, :, c].numpy() == pytest.approx(fd, abs=2e-5, rel=1e-3), c
