"""The CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md section 8(c)): PlanarTerrain exact outputs, one-step integrators, the mass-falling
Euler recursion.  CPU only."""
import math

import numpy as np
import pytest

from oracle import expressions as ex
from oracle import sx, toy


def _eval(exprs, syms, vals):
    return sx.Tape(list(exprs), list(syms)).eval(np.asarray(vals, dtype=float)[None])[0]


def test_planar_terrain_known_answers():
    """/root/reference/test/test_planar_terrain.py:7-40: exact (==) outputs at (0, 0, 0.5)."""
    p = sx.syms("p", 3)
    t = ex.PlanarTerrain()
    point = [0.0, 0.0, 0.5]
    assert _eval([t.height(p)], p, point)[0] == 0.5
    assert (_eval(t.normal(p), p, point) == np.array([0.0, 0.0, 1.0])).all()
    assert (_eval(t.orientation(p).ravel(), p, point).reshape(3, 3) == np.eye(3)).all()
    # transform_function: [R, (p_x, p_y, p_z - h); 0 0 0 1] == eye(4) at that point
    h = _eval([t.height(p)], p, point)[0]
    assert (np.array([point[0], point[1], point[2] - h]) == 0.0).all()
    # the normal is a structural constant: n . f collapses to f_z (planar_terrain.py:25)
    f = sx.syms("f", 3)
    assert ex.normal_force_component(t, p, f) is f[2]


@pytest.mark.parametrize("step", [toy.euler_step, toy.trapezoid_step])
def test_one_step_integrators(step):
    """/root/reference/test/test_integrators.py:69-117: x_dot = lam x, x = 0.5, lam = 1, dt = 0.005,
    x0 = xf, against x exp(lam dt) to rel 1e-4."""
    x, lam, dt = 0.5, 1.0, 0.005
    integrated = step(x, lam * x, lam * x, dt)
    assert float(integrated) == pytest.approx(x * math.exp(lam * dt), rel=1e-4)
    assert float(integrated) == pytest.approx(x * (1 + lam * dt), rel=1e-15)


@pytest.mark.parametrize("integrator", ["euler"])
def test_mass_falling_closed_form_is_a_kkt_point(integrator):
    """/root/reference/test/test_multiple_shooting.py:253-353: horizon 100, dt 0.01, x0 = 1, v0 = 0,
    g = -9.81; the asserted solution (Euler recursion, foo = 0 / 5 / 6) must be feasible and stationary
    for the oracle's restatement of that NLP."""
    N, dt, g, x0, v0 = 100, 0.01, -9.81, 1.0, 0.0
    nlp = toy.build(N, integrator, dt)
    assert (nlp.n_x, nlp.m) == (900, 703)
    x = toy.closed_form_solution(N, dt, g, x0, v0)[None]
    p = np.array([[g, x0, v0]])
    gv = nlp.eval_g(x, p)
    lb, ub = nlp.eval_bounds(p)
    assert np.all(gv >= lb - 1e-12) and np.all(gv <= ub + 1e-12)
    # cost: mass 1 follows the dynamics exactly, so only sumsqr(foo) remains
    assert nlp.eval_f(x, p)[0] == pytest.approx(3 * (98 * 25.0 + 36.0), rel=1e-14)
    grad = nlp.eval_grad_f(x, p)[0]
    J = nlp.dense_jac(x, p)[0]
    lam, *_ = np.linalg.lstsq(J.T, -grad, rcond=None)
    assert np.abs(J.T @ lam + grad).max() < 1e-9
    # multipliers of the active ``foo >= 5`` rows must push inwards (lower bound active: lam <= 0)
    o = 2 * (N - 1) + 2 + 2 * N
    assert np.all(lam[o:o + 3 * (N - 2)] <= 1e-9)
    assert lam[o] == pytest.approx(-10.0)


def test_mass_falling_trapezoid_rows():
    nlp = toy.build(8, "trapezoid", 0.1)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, nlp.n_x))
    p = np.array([[-9.81, 1.0, 0.0], [-8.0, 0.5, 0.3]])
    g = nlp.eval_g(x, p)
    # first row: x1 - (x0 + dt/2 (v0 + v1))
    assert g[:, 0] == pytest.approx(x[:, 2] - (x[:, 0] + 0.05 * (x[:, 1] + x[:, 3])))
    assert g[:, 1] == pytest.approx(x[:, 3] - (x[:, 1] + 0.05 * (p[:, 0] + p[:, 0])))
