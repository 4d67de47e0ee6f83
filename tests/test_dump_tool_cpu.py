"""tools/dump_casadi_golden.py on a stand-in module tree: the CasADi-facing core writes a file with the golden
fixtures' key layout (plus what IPOPT sees after detect_simple_bounds), and the parity tests' loader picks such a
file up.  With the real modules the same code dumps CasADi's own numbers (the builder's container has none)."""
import importlib.util
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "fakes"))
import fake_casadi as fcs  # noqa: E402

from hippopt_b200 import plugin  # noqa: E402
from hippopt_b200.kino_layout import KinoLayout, KinoSettings  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _tool():
    spec = importlib.util.spec_from_file_location("dump_casadi_golden", os.path.join(ROOT, "tools", "dump_casadi_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_dump_has_the_fixture_layout_and_is_picked_up(model, tmp_path):
    from oracle import kinodynamic as kd

    fx = dict(np.load(os.path.join(GOLD, "kino_n3_flat.npz")))
    N = int(fx["horizon"])
    lay = KinoLayout(model, KinoSettings(horizon=N))
    nlp, _ = kd.build(model, kd.Settings(horizon=N))
    lbg, ubg = lay.bounds(fx["p"])
    opti = fcs.Opti(lambda x, p: nlp.eval_f(x[None], p[None]), lambda x, p: nlp.eval_g(x[None], p[None])[0],
                    fx["x"][0], fx["p"][0], lbg[0], ubg[0])
    opti.attach_oracle(nlp, plugin.RowReduction(lay, True))
    tool = _tool()
    out = tool.dump_nlp(fcs, opti, fx["x"], fx["p"], fx["lam"], fx["sigma"])
    # same keys as the committed fixture (settings keys are copied by main())
    need = {"x", "p", "lam", "sigma", "f", "grad_f", "g", "jac", "hess", "jac_colind", "jac_row", "hess_colind",
            "hess_row", "lbg", "ubg"}
    assert need <= set(out)
    for k in need:
        assert out[k].shape == fx[k].shape, k
        assert np.allclose(out[k], fx[k], rtol=1e-13, atol=1e-13, equal_nan=True), k  # the stand-in IS the oracle
    # what IPOPT sees with {"expand", "detect_simple_bounds"}
    rows, _ = lay.simple_bound_rows()
    assert int(out["sb_m"]) == lay.m - len(rows) and len(out["sb_g_rows"]) == lay.m - len(rows)
    assert out["sb_jac_colind"][-1] == len(out["sb_jac_row"]) < len(fx["jac_row"])
    assert np.array_equal(out["sb_hess_row"], fx["hess_row"])
    assert np.isfinite(out["sb_lbx"]).sum() > 0
    # picked up by the parity tests' loader under the casadi_ prefix
    for k in ("horizon", "final", "periodicity", "smooth"):
        out[k] = fx[k]
    path = tmp_path / "casadi_kino_n3_flat.npz"
    np.savez_compressed(path, **out)
    from tests_support import golden_files

    found = golden_files("kino", extra_dirs=[str(tmp_path)])
    assert str(path) in found and any(os.path.basename(f) == "kino_n3_flat.npz" for f in found)


def test_main_reports_the_missing_environment(capsys):
    assert _tool().main([]) == 2
    assert "reference environment" in capsys.readouterr().out


def test_per_instance_dumps_merge_into_one_fixture():
    """Stairs fixtures are dumped one planner per instance (the reference bakes the step dimensions into the graph):
    batched keys stack, patterns must agree."""
    import importlib.util
    import os

    import numpy as np
    import pytest

    spec = importlib.util.spec_from_file_location(
        "dump_casadi_golden", os.path.join(os.path.dirname(__file__), "..", "tools", "dump_casadi_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    def part(v):
        return {"x": np.full((1, 3), v), "p": np.full((1, 2), v), "lam": np.zeros((1, 2)), "sigma": np.ones(1),
                "f": np.array([v]), "grad_f": np.full((1, 3), v), "g": np.zeros((1, 2)), "jac": np.full((1, 4), v),
                "hess": np.full((1, 5), v), "lbg": np.zeros((1, 2)), "ubg": np.zeros((1, 2)),
                "jac_colind": np.array([0, 2, 3, 4]), "jac_row": np.array([0, 1, 0, 1]), "sb_m": np.int64(2)}

    out = mod.merge_instances([part(1.0), part(2.0)])
    assert out["x"].shape == (2, 3) and out["f"].tolist() == [1.0, 2.0] and out["jac"][1, 0] == 2.0
    assert np.array_equal(out["jac_row"], [0, 1, 0, 1]) and int(out["sb_m"]) == 2
    bad = part(3.0)
    bad["jac_row"] = np.array([0, 1, 1, 1])
    with pytest.raises(ValueError, match="jac_row"):
        mod.merge_instances([part(1.0), bad])
