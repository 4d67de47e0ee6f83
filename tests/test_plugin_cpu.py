"""The solver back end behind hippopt's interface (hippopt_b200/plugin.py) without CasADi, hippopt or a GPU:
* the CasADi shim `B200OptiSolver.solve()` is executed against recording stand-ins of the two modules
  (tests/fakes), with the CPU oracle playing both the Opti graph and the kernel evaluator;
* the casadi-free `B200Solver` exposes the 16 methods of `OptimizationSolver`
  (/root/reference/src/hippopt/base/optimization_solver.py:24-96) with the reference's error behaviour."""
import inspect

import numpy as np
import pytest

from hippopt_b200 import naming, plugin
from hippopt_b200.kino_layout import KinoLayout, KinoSettings
from hippopt_b200.workloads import kino_batch
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "fakes"))
import fake_casadi as fcs  # noqa: E402
import fake_hippopt as fhp  # noqa: E402

MASKS = {"f": 1, "grad_f": 2, "g": 4, "jac": 8, "hess": 16}


@pytest.fixture(scope="module")
def small(model):
    """kinodynamic OCP with N = 2: layout + oracle + one instance's vectors."""
    from oracle import kinodynamic as kd

    lay = KinoLayout(model, KinoSettings(horizon=2))
    nlp, _ = kd.build(model, kd.Settings(horizon=2))
    x, p, lam, sigma = kino_batch(lay, model, 1, seed=11, noise=0.05)
    lbg, ubg = lay.bounds(p)
    return lay, nlp, x[0], p[0], lbg[0], ubg[0]


class OracleHost:
    """test double of HostEvaluator: same call, numbers from the CPU oracle"""

    def __init__(self, nlp, p):
        self.nlp, self.p, self.n_calls = nlp, p, 0

    def __call__(self, mask, x, lam, sigma):
        self.n_calls += 1
        out, nlp, X, P = {}, self.nlp, x[None], self.p[None]
        if mask & 1:
            out["f"] = nlp.eval_f(X, P)
        if mask & 2:
            out["grad_f"] = nlp.eval_grad_f(X, P)[0]
        if mask & 4:
            out["g"] = nlp.eval_g(X, P)[0]
        if mask & 8:
            out["jac"] = nlp.eval_jac(X, P)[0]
        if mask & 16:
            out["hess"] = nlp.eval_hess(X, P, lam[None], np.asarray(sigma))[0]
        return out


def make_solver(small, options_plugin=None, options_solver=None, perturb_f=0.0):
    lay, nlp, x0, p, lbg, ubg = small
    opti = fcs.Opti(lambda x, pp: nlp.eval_f(x[None], pp[None]) * (1.0 + perturb_f), lambda x, pp: nlp.eval_g(x[None], pp[None])[0],
                    x0, p, lbg, ubg)
    host = OracleHost(nlp, p)
    cls = plugin.make_opti_solver(fcs, fhp, evaluator_factory=lambda match: (lay, host, MASKS))
    s = cls(opti, options_plugin=options_plugin, options_solver=options_solver)
    # two variables and one parameter in creation order, two named costs, two named constraints (one flipped)
    s._variables_map = {opti.variable_slice(0, 189): "system[0]", opti.variable_slice(189, lay.n_x): "rest"}
    s._parameters_map = {opti.parameter_slice(0, 24): "descriptors"}
    s._cost_expressions = {"twice_f": opti.expression(lambda x, pp: 2.0 * nlp.eval_f(x[None], pp[None])),
                           "x0_squared": opti.expression(lambda x, pp: x[:1] ** 2)}
    s._constraint_expressions = {"first_rows": opti.constraint(0, 3), "flipped_rows": opti.constraint(3, 7, flipped=True)}
    s._cost = fcs.MX(0)
    return s, opti, host


def test_shim_runs_ipopt_through_callbacks(small):
    lay, nlp, x0, p, lbg, ubg = small
    fcs.RECORD.clear()
    s, opti, host = make_solver(small, options_plugin={"expand": True, "detect_simple_bounds": True, "print_time": False},
                                options_solver={"tol": 1e-3, "max_iter": 7})
    s.solve()
    assert not s.stock_solve_called and opti.minimized is s._cost
    bridge = s.last_bridge
    # options: what cannot be forwarded is stripped, the rest and the IPOPT dict go through, oracles are Callbacks
    opts = bridge.options_passed
    assert "expand" not in opts and "detect_simple_bounds" not in opts and opts["print_time"] is False
    assert opts["ipopt"] == {"tol": 1e-3, "max_iter": 7} and opts["calc_lam_p"] is False
    for k in ("grad_f", "jac_g", "hess_lag"):
        assert isinstance(opts[k], fcs.Callback)
    # detect_simple_bounds: IPOPT sees the reduced row set and bounds on x
    rows, cols = lay.simple_bound_rows()
    solve_rec = [r for r in fcs.RECORD if r[0] == "solve"][0][1]
    assert solve_rec["m"] == lay.m - len(rows) and solve_rec["finite_lbx"] > 0
    assert opts["jac_g"].get_sparsity_out(1).size1() == lay.m - len(rows)
    assert opts["hess_lag"].get_sparsity_out(0).nnz() == len(lay.hess_row)
    # two launches per iterate: one first-order evaluation per distinct x (+ the verification probe), one hess_l
    iters = fcs._Nlpsol.iterations
    assert bridge.cache.launches == {"first_order": iters, "hess": iters}  # the probe's x is iterate 0's x: cached
    assert bridge.cache.calls["f"] == 2 * iters and bridge.cache.calls["g"] == 2 * iters
    assert host.n_calls == 2 * iters
    # outputs of opti_solver.py:522-537
    xs = s.last_bridge.solver._x
    assert np.array_equal(s._output_solution["system[0]"], xs[:189]) and np.array_equal(s._output_solution["rest"], xs[189:])
    assert np.array_equal(s._output_solution["descriptors"], p[:24])
    assert s._cost_values["twice_f"] == pytest.approx(2.0 * float(nlp.eval_f(xs[None], p[None])[0]), rel=1e-14)
    assert s._cost_values["x0_squared"] == pytest.approx(xs[0] ** 2)
    lam_full = bridge.red.lam_full(s.last_bridge.solver._lam)
    assert np.array_equal(s._constraint_values["first_rows"], lam_full[0:3])
    assert np.array_equal(s._constraint_values["flipped_rows"], -lam_full[3:7])  # Opti's own sign rule is applied


def test_limited_memory_needs_no_hessian_callback(small):
    s, _, host = make_solver(small, options_solver={"hessian_approximation": "limited-memory"})
    s.solve()
    assert "hess_lag" not in s.last_bridge.options_passed
    assert s.last_bridge.cache.launches["hess"] == 0 and host.n_calls == fcs._Nlpsol.iterations


def test_failure_and_fallback(small):
    s, _, _ = make_solver(small)
    fcs._Nlpsol.succeed = False
    try:
        with pytest.raises(fhp.OptiFailure, match="Infeasible_Problem_Detected"):
            s.solve()
    finally:
        fcs._Nlpsol.succeed = True
    # a graph that is not the template's (f differs): the probe notices and the stock path runs
    s, _, _ = make_solver(small, perturb_f=1e-6)
    s.solve()
    assert s.stock_solve_called
    s, _, _ = make_solver(small, perturb_f=1e-6)
    s.strict = True
    with pytest.raises(plugin.TemplateMismatch):
        s.solve()
    s, _, _ = make_solver(small)
    s._free_parameters = ["dt"]
    with pytest.raises(ValueError, match="The following parameters are not set"):
        s.solve()


def test_row_reduction(small):
    lay, nlp, x0, p, lbg, ubg = small
    red = plugin.RowReduction(lay, True)
    rows, cols = lay.simple_bound_rows()
    assert red.m_reduced == lay.m - len(rows)
    lr, ur, lbx, ubx = red.bounds(lbg, ubg)
    assert np.array_equal(lr, np.delete(lbg, rows)) and np.array_equal(ur, np.delete(ubg, rows))
    for r, j in zip(rows, cols):
        assert lbx[j] >= lbg[r] and ubx[j] <= ubg[r]
    free = np.setdiff1d(np.arange(lay.n_x), cols)
    assert np.all(np.isinf(lbx[free])) and np.all(np.isinf(ubx[free]))
    # the reduced Jacobian is the full one without the simple rows, still in compressed-column order
    J = nlp.dense_jac(x0[None], p[None])[0]
    vals = nlp.eval_jac(x0[None], p[None])[0]
    Jr = np.zeros((red.m_reduced, lay.n_x))
    col = np.repeat(np.arange(lay.n_x), np.diff(red.jac_colind))
    Jr[red.jac_row, col] = red.jac(vals)
    assert np.array_equal(Jr, np.delete(J, rows, axis=0))
    lam_r = np.arange(red.m_reduced, dtype=float) + 1.0
    full = red.lam_full(lam_r)
    assert np.array_equal(full[red.general], lam_r) and np.all(full[rows] == 0.0)
    off = plugin.RowReduction(lay, False)
    assert off.m_reduced == lay.m and np.array_equal(off.jac_keep, np.arange(len(lay.jac_row)))


def test_b200solver_implements_the_abstract_interface(model):
    """Method names and argument names of optimization_solver.py:24-96 (+ cost_function, which
    MultipleShootingSolver calls: multiple_shooting_solver.py:906-907)."""
    expected = {
        "generate_optimization_objects": ["input_structure"], "get_optimization_objects": [],
        "get_optimization_structure": [], "register_problem": ["problem"], "get_problem": [],
        "set_initial_guess": ["initial_guess"], "get_initial_guess": [], "solve": [], "get_values": [],
        "get_cost_value": [], "add_cost": ["input_cost", "name"], "add_constraint": ["input_constraint", "name"],
        "get_cost_expressions": [], "get_constraint_expressions": [], "get_cost_values": [],
        "get_constraint_multipliers": [], "cost_function": [],
    }
    for name, args in expected.items():
        fn = getattr(plugin.B200Solver, name)
        got = [a for a in inspect.signature(fn).parameters if a not in ("self", "kwargs")]
        assert got[:len(args)] == args, name
    s = plugin.B200Solver(model=model, settings=KinoSettings(horizon=3), batch=2)
    with pytest.raises(plugin.ProblemNotRegisteredException, match="No problem has been registered."):
        s.get_problem()
    with pytest.raises(plugin.SolutionNotAvailableException, match="No solution is available"):
        s.get_values()
    with pytest.raises(plugin.SolutionNotAvailableException):
        s.get_cost_value()
    with pytest.raises(ValueError, match="neither an optimization object nor a list"):
        s.generate_optimization_objects([1, 2])
    lay = KinoLayout(model, KinoSettings(horizon=3))
    x, p, _, _ = kino_batch(lay, model, 2, seed=1)
    objs = s.generate_optimization_objects({"x": x, "p": p})
    assert s.get_optimization_objects() is objs and set(s.get_optimization_structure()) == {"x", "p"}
    assert np.array_equal(s.get_initial_guess()["x"], x)
    assert set(s.get_constraint_expressions()) == set(naming.constraint_rows(lay))
    assert set(s.get_cost_expressions()) == set(naming.cost_slots(lay))
    with pytest.raises(ValueError, match="does not match"):
        s.set_initial_guess({"x": x[:, :-1], "p": p})
    s.register_problem("problem")
    assert s.get_problem() == "problem"
    # optional expressions of the template: the reference's ExpressionType switch (planner.py:417, 923)
    assert "final_state_expression" not in s.get_constraint_expressions()
    s.add_constraint(plugin.TemplateExpression("final_state_expression", "constraint"))
    assert "final_state_expression" in s.get_constraint_expressions() and s._layout().m == lay.m + 105
    with pytest.raises(ValueError, match="The constraint final_state_expression is already present."):
        s.add_constraint(plugin.TemplateExpression("final_state_expression", "constraint"))
    with pytest.raises(ValueError, match="Only the template's own"):
        s.add_cost("not an expression", name="c")


def test_install_needs_the_reference_environment():
    with pytest.raises(ImportError, match="casadi"):
        plugin.install(None, None)
