"""oracle/toy.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the mass-falling multiple-shooting OCP the reference solves in
`/root/reference/test/test_multiple_shooting.py:210-353` (config 1), on oracle/nlp.py:
row order = the test's ``add_dynamics`` / ``add_constraint`` / ``add_expression_to_horizon`` call
order; integrator steps follow `integrators/forward_euler.py:31-33` and
`integrators/implicit_trapezoid.py:31-37`; dynamics added with ``mode=minimize`` become
``sumsqr(lhs - rhs)`` costs (`base/problem.py:118-122`).

x = [masses[0][0..N-1].{x,v}, masses[1][..], masses[2][..], foo[0..N-1](3)], p = [g, x0, v0]
(x0, v0 are the test's ``initial_position`` / ``initial_velocity`` constants made runtime data).
"""
from __future__ import annotations

import numpy as np

from . import sx
from .nlp import NLP, Template

INF = float("inf")


def euler_step(x0, f0, f1, dt):
    """forward_euler.py:31-33."""
    return x0 + dt * f0


def trapezoid_step(x0, f0, f1, dt):
    """implicit_trapezoid.py:31-37."""
    return x0 + 0.5 * dt * (f0 + f1)


def build(N: int = 100, integrator: str = "euler", dt: float = 0.01) -> NLP:
    step = {"euler": euler_step, "trapezoid": trapezoid_step}[integrator]
    n_x = 9 * N
    nlp = NLP(n_x, 3)
    G, X0, V0 = n_x, n_x + 1, n_x + 2

    def X(j, i):
        return j * 2 * N + 2 * i

    def V(j, i):
        return X(j, i) + 1

    def FOO(i):
        return [6 * N + 3 * i + c for c in range(3)]

    xk, vk, xn, vn, g = (sx.sym(n) for n in ("xk", "vk", "xn", "vn", "g"))
    # dynamics of MassFallingState.get_dynamics(): x_dot = v, v_dot = g
    rx = xn - step(xk, vk, vn, dt)
    rv = vn - step(vk, g, g, dt)
    t_both = Template("mass_dyn", [xk, vk, xn, vn, g], [rx, rv])
    t_x = Template("x_dyn", [xk, vk, xn, vn], [rx])
    t_v = Template("v_dyn", [vk, vn, g], [rv])
    a, b = sx.sym("a"), sx.sym("b")
    t_ic = Template("ic", [a, b], [a], lb=[b], ub=[b])
    t_cost_sq = Template("sq", [a, b], [sx.sq(a - b)])
    t_cx = Template("x_dyn_cost", [xk, vk, xn, vn], [sx.sq(rx)])
    t_cv = Template("v_dyn_cost", [vk, vn, g], [sx.sq(rv)])
    foo = sx.syms("foo", 3)
    t_foo_ge = Template("foo_ge", list(foo), list(foo), lb=[5.0] * 3, ub=[INF] * 3)
    t_foo_0 = Template("foo_eq0", list(foo), list(foo), lb=[0.0] * 3, ub=[0.0] * 3)
    t_foo_6 = Template("foo_eq6", list(foo), list(foo), lb=[6.0] * 3, ub=[6.0] * 3)
    t_foo_cost = Template("foo_cost", list(foo), [sx.sumsqr(foo)])

    for i in range(N - 1):  # test :266-270
        nlp.subject_to(t_both, [X(0, i), V(0, i), X(0, i + 1), V(0, i + 1), G], f"dot(masses[0])[{i + 1}]")
    nlp.subject_to(t_ic, [X(0, 0), X0], "initial_position")  # :272-275
    nlp.subject_to(t_ic, [V(0, 0), V0], "initial_velocity")
    nlp.minimize(t_cost_sq, [X(1, 0), X0])  # :277-285 (mode=minimize)
    nlp.minimize(t_cost_sq, [V(1, 0), V0])
    for i in range(N - 1):
        nlp.minimize(t_cx, [X(1, i), V(1, i), X(1, i + 1), V(1, i + 1)])
        nlp.minimize(t_cv, [V(1, i), V(1, i + 1), G])
    nlp.subject_to(t_ic, [X(2, 0), X0], "initial_condition_simple_x")  # :287-293
    for i in range(N - 1):
        nlp.subject_to(t_x, [X(2, i), V(2, i), X(2, i + 1), V(2, i + 1)], f"dot(masses[2].x)[{i + 1}]")
    nlp.subject_to(t_ic, [V(2, 0), V0], "initial_condition_simple_v")  # :295-301
    for i in range(N - 1):
        nlp.subject_to(t_v, [V(2, i), V(2, i + 1), G], f"dot(masses[2].v)[{i + 1}]")
    for i in range(1, N):  # :303-305
        nlp.subject_to(t_foo_ge, FOO(i), f"foo_ge[{i}]")
    nlp.subject_to(t_foo_0, FOO(0), "foo_initial")  # :307-308
    nlp.subject_to(t_foo_6, FOO(N - 1), "foo_final")
    for i in range(N):  # :310-314
        nlp.minimize(t_foo_cost, FOO(i))
    return nlp


def closed_form_solution(N: int, dt: float, g: float, x0: float, v0: float) -> np.ndarray:
    """The solution the reference's test asserts (:336-353): explicit Euler recursion, foo = 0/5/6."""
    x = np.zeros(9 * N)
    pos, vel = x0, v0
    for i in range(N):
        for j in range(3):
            x[j * 2 * N + 2 * i] = pos
            x[j * 2 * N + 2 * i + 1] = vel
        x[6 * N + 3 * i:6 * N + 3 * i + 3] = 0.0 if i == 0 else (6.0 if i == N - 1 else 5.0)
        pos += dt * vel
        vel += dt * g
    return x
