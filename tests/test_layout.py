"""The product's layout compiler against the oracle's independent derivation (dependency propagation
on the restated graph): x / p / g sizes, bounds, and the CCS sparsity of jac_g and hess_l must be
identical.  CPU only."""
import numpy as np
import pytest

from hippopt_b200.kino_layout import KinoLayout, KinoSettings
from oracle import kinodynamic as kd


@pytest.mark.parametrize("N,fin,per", [(3, False, False), (4, True, True), (2, True, False)])
def test_patterns_match_oracle(model, N, fin, per):
    lay = KinoLayout(model, KinoSettings(horizon=N, final_state_constraint=fin, periodicity_constraint=per))
    nlp, _ = kd.build(model, kd.Settings(horizon=N, final_state_constraint=fin, periodicity_constraint=per))
    assert (lay.n_x, lay.n_p, lay.m) == (nlp.n_x, nlp.n_p, nlp.m)
    colind, row, _, _ = nlp.jac_structure()
    assert np.array_equal(colind, lay.jac_colind) and np.array_equal(row, lay.jac_row)
    hcol, hrow, _, _ = nlp.hess_structure()
    assert np.array_equal(hcol, lay.hess_colind) and np.array_equal(hrow, lay.hess_row)
    assert np.all(lay.hess_row <= lay.hess_col)  # upper triangle
    P = np.random.default_rng(0).uniform(0.5, 1.5, (2, lay.n_p))
    lb, ub = lay.bounds(P)
    olb, oub = nlp.eval_bounds(P)
    assert np.array_equal(lb, olb) and np.array_equal(ub, oub)


def test_sizes_of_the_reference_configs(model):
    """n_x = 189 N + 6, n_p = 79 N + 326; row counts of configs 3 and 4."""
    lay = KinoLayout(model, KinoSettings(horizon=30))
    assert (lay.n_x, lay.n_p) == (5676, 2696)
    assert lay.m == 8142
    lay4 = KinoLayout(model, KinoSettings(horizon=30, final_state_constraint=True, periodicity_constraint=True))
    assert lay4.m == 8142 + 105 + 84 - 6
    # every local entry maps to a distinct slot and every slot is produced exactly once
    for l in (lay, lay4):
        slots = np.concatenate([l.jc_map.ravel(), l.jk_map.ravel()])
        slots = slots[slots >= 0]
        assert len(slots) == l.nnz_j and len(np.unique(slots)) == l.nnz_j
        hs = np.concatenate([l.hc_map.ravel(), l.hk_map.ravel(), l.hk2_map.ravel()])
        hs = hs[hs >= 0]
        assert len(hs) == l.nnz_h and len(np.unique(hs)) == l.nnz_h


def test_knot_columns_are_contiguous_ccs_segments(model):
    """Column-knot ownership: all Jacobian / Hessian slots a warp writes lie in its knot's segment."""
    lay = KinoLayout(model, KinoSettings(horizon=5))
    for k in range(5):
        lo, hi = lay.jac_colind[189 * k], lay.jac_colind[189 * (k + 1)]
        s = np.concatenate([lay.jc_map[k], lay.jk_map[k]])
        s = s[s >= 0]
        extra = s[(s < lo) | (s >= hi)]
        # only the six -1 entries of the momentum initial condition live in the trailing columns
        assert len(extra) == (6 if k == 0 else 0)
        lo, hi = lay.hess_colind[189 * k], lay.hess_colind[189 * (k + 1)]
        s = np.concatenate([lay.hc_map[k], lay.hk_map[k], lay.hk2_map[k]])
        s = s[s >= 0]
        assert np.all((s >= lo) & (s < hi))
