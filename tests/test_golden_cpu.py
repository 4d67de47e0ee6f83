"""The oracle reproduces the committed golden fixtures (guards the fixtures against oracle drift) and
the product layout's pattern equals the fixtures' pattern.  CPU only."""
import glob
import os

import numpy as np
import pytest

from tests_support import golden_files

from hippopt_b200.kino_layout import KinoLayout, KinoSettings
from oracle import kinodynamic as kd
from oracle import toy

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", golden_files("kino"))
def test_kino_fixture(model, path):
    d = np.load(path)
    from oracle import expressions as ex

    N, fin, per = int(d["horizon"]), bool(d["final"]), bool(d["periodicity"])
    smooth = bool(d["smooth"]) if "smooth" in d else False
    extra = dict(terrain=ex.TwoSmoothSteps(), terrain_params=10) if smooth else {}
    nlp, _ = kd.build(model, kd.Settings(horizon=N, final_state_constraint=fin, periodicity_constraint=per, **extra))
    if os.path.basename(path).startswith("casadi_"):
        # a dump of CasADi's own numbers (tools/dump_casadi_golden.py): THE pin of the oracle -- all five outputs to
        # north_star's 1e-10 relative, patterns identical
        from oracle import parity

        assert np.array_equal(nlp.jac_structure()[0], d["jac_colind"]) and np.array_equal(nlp.jac_structure()[1], d["jac_row"])
        assert np.array_equal(nlp.hess_structure()[0], d["hess_colind"]) and np.array_equal(nlp.hess_structure()[1], d["hess_row"])
        parity.check_all(parity.reference_outputs(nlp, d["x"], d["p"], d["lam"], d["sigma"]), {k: d[k] for k in d.files},
                         (d["jac_colind"], d["jac_row"]), (d["hess_colind"], d["hess_row"]), d["x"])
    else:
        assert nlp.eval_g(d["x"], d["p"]) == pytest.approx(d["g"], rel=1e-13, abs=1e-13)
        assert nlp.eval_f(d["x"], d["p"]) == pytest.approx(d["f"], rel=1e-13)
    lay = KinoLayout(model, KinoSettings(horizon=N, final_state_constraint=fin, periodicity_constraint=per,
                                         terrain="smooth_steps" if smooth else "planar",
                                         n_terrain_params=10 if smooth else 0))
    assert np.array_equal(lay.jac_colind, d["jac_colind"]) and np.array_equal(lay.jac_row, d["jac_row"])
    assert np.array_equal(lay.hess_colind, d["hess_colind"]) and np.array_equal(lay.hess_row, d["hess_row"])
    lb, ub = lay.bounds(d["p"])
    assert np.array_equal(lb, d["lbg"]) and np.array_equal(ub, d["ubg"])


@pytest.mark.parametrize("path", golden_files("toy"))
def test_toy_fixture(path):
    d = np.load(path)
    integrator = "euler" if "euler" in path else "trapezoid"
    nlp = toy.build(int(d["horizon"]), integrator, float(d["dt"]))
    assert nlp.eval_g(d["x"], d["p"]) == pytest.approx(d["g"], rel=1e-14, abs=1e-14)
    assert nlp.eval_hess(d["x"], d["p"], d["lam"], d["sigma"]) == pytest.approx(d["hess"], rel=1e-14)
