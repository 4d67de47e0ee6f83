"""The parity metric itself (oracle/parity.py): exact data passes, a relative perturbation of 1e-9 of one entry
fails whatever the entry's magnitude relative to 1, and a perturbation far below a row's scale passes."""
import os

import numpy as np
import pytest

from oracle import parity

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    d = np.load(os.path.join(GOLD, "kino_n3_flat.npz"))
    return {k: d[k] for k in d.files}


def _patterns(d):
    return (d["jac_colind"], d["jac_row"]), (d["hess_colind"], d["hess_row"])


def test_exact_data_passes(gold):
    jp, hp = _patterns(gold)
    worst = parity.check_all(gold, gold, jp, hp, gold["x"])
    assert set(worst) == {"f", "grad_f", "g", "jac", "hess"} and max(worst.values()) == 0.0


@pytest.mark.parametrize("key", ["jac", "hess", "grad_f", "g", "f"])
def test_relative_perturbation_of_a_large_entry_fails(gold, key):
    """An entry that dominates its row carries no floor: 1e-9 relative is caught even if |entry| << 1 or >> 1."""
    jp, hp = _patterns(gold)
    got = {k: np.array(v, dtype=float, copy=True) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v
           for k, v in gold.items()}
    a = np.atleast_2d(got[key])
    ref = np.atleast_2d(gold[key])
    if key == "jac":
        S = parity.jac_row_scale(ref, jp[1], gold["g"].shape[1])[:, jp[1]]
        e = int(np.argmax((np.abs(ref[0]) == S[0]) & (np.abs(ref[0]) < 0.5) & (ref[0] != 0)))  # row leader below 1
        assert 0 < abs(ref[0, e]) < 0.5
    elif key == "f":
        e = 0
    else:
        e = int(np.argmax(np.abs(ref[0])))
    a[0, e] *= 1.0 + 1e-9
    got[key] = a.reshape(gold[key].shape)
    with pytest.raises(AssertionError, match="max relative error"):
        parity.check_all(got, gold, jp, hp, gold["x"])


def test_small_entry_is_judged_against_its_row(gold):
    """A nearly cancelling entry may move by 1e-12 of its row's scale, but not by 1e-9 of it."""
    jp, hp = _patterns(gold)
    ref = gold["jac"]
    S = parity.jac_row_scale(ref, jp[1], gold["g"].shape[1])[:, jp[1]]
    ratio = np.where(ref[0] != 0, np.abs(ref[0]) / S[0], 1.0)
    e = int(np.argmin(ratio))
    assert ratio[e] < 1e-2
    got = dict(gold)
    a = ref.copy()
    a[0, e] += 1e-12 * S[0, e]
    got["jac"] = a
    parity.check_all(got, gold, jp, hp, gold["x"], keys=("jac",))
    a[0, e] += 1e-9 * S[0, e]
    with pytest.raises(AssertionError):
        parity.check_all(got, gold, jp, hp, gold["x"], keys=("jac",))


def test_nan_is_an_error(gold):
    jp, hp = _patterns(gold)
    got = dict(gold)
    a = gold["hess"].copy()
    a[1, 7] = np.nan
    got["hess"] = a
    with pytest.raises(AssertionError):
        parity.check_all(got, gold, jp, hp, gold["x"], keys=("hess",))
