"""Expression names (row f4, hippopt_b200/naming.py): against the names, call kinds and `apply_to_first_elements`
flags parsed from the reference planner's source (tests/golden/reference_expression_names.json, made by
tests/dev/parse_reference_names.py) and against the oracle's recording of the same calls."""
import json
import os
import re

import numpy as np
import pytest

from hippopt_b200 import naming
from hippopt_b200.kino_layout import KinoLayout, KinoSettings

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_expression_names.json")


def _pattern(name: str) -> tuple[str, list[int], bool]:
    """full name -> (reference pattern, knot indices, has "{j}")."""
    knots = [int(k) for k in re.findall(r"\[(\d+)\](?=\{|$)", name)]
    base = re.sub(r"\[\d+\](\{\d+\})?$", "", name)
    base = re.sub(r"^system\.contact_points\.(left|right)\[\d\]\.", "<point>.", base)
    base = re.sub(r"^(left|right)_yaw", "<foot_name>_yaw", base)
    return base, knots, name.endswith("}")


@pytest.fixture(scope="module")
def reference():
    return {e["name"]: e for e in json.load(open(GOLD))["expressions"]}


@pytest.mark.parametrize("fin,per", [(False, False), (True, True)])
def test_names_match_the_reference_planner(model, reference, fin, per):
    N = 5
    lay = KinoLayout(model, KinoSettings(horizon=N, final_state_constraint=fin, periodicity_constraint=per))
    cons, costs = naming.constraint_rows(lay), naming.cost_slots(lay)
    seen: dict[str, set] = {}
    for full in list(cons) + list(costs):
        base, knots, gen = _pattern(full)
        assert base in reference, f"{full}: {base} is not a name= argument of the reference planner"
        ref = reference[base]
        is_cost = full in costs
        if "minimize" in ref["mode"]:
            assert is_cost, full
        elif "settings" not in ref["mode"]:
            assert not is_cost, full
        assert gen == (ref["call"] == "add_dynamics"), full  # "{j}": generator expressions only
        if ref["call"] == "add_expression":
            assert knots == [], full
        seen.setdefault(base, set()).update(knots)
    expected = set(reference)
    if not fin:
        expected.discard("final_state_expression")
    if not per:
        expected.discard("periodicity_expression")
    assert set(seen) == expected
    for base, knots in seen.items():
        ref = reference[base]
        if ref["call"] == "add_expression_to_horizon":
            first = 0 if ref["apply_to_first_elements"] == "True" else 1
            assert knots == set(range(first, N)), base
        elif ref["call"] == "add_dynamics":
            # [0] = initial condition, dropped for the momentum when the periodicity expression is on (planner.py:580-584)
            want = set(range(1, N)) if (per and base == "centroidal_momentum_dynamics") else set(range(N))
            assert knots == want, base


def test_rows_partition_g_and_match_the_oracle(model):
    from oracle import kinodynamic as kd

    for fin, per in ((False, False), (True, False), (True, True)):
        lay = KinoLayout(model, KinoSettings(horizon=4, final_state_constraint=fin, periodicity_constraint=per))
        cons = naming.constraint_rows(lay)
        allrows = np.concatenate(list(cons.values()))
        assert np.array_equal(allrows, np.arange(lay.m))  # every row named once, in the reference's order
        nlp, _ = kd.build(model, kd.Settings(horizon=4, final_state_constraint=fin, periodicity_constraint=per))
        ref = {n: np.arange(o, o + r) for n, o, r in nlp.constraint_names}
        assert list(cons) == list(ref)
        assert all(np.array_equal(cons[k], ref[k]) for k in ref)
        assert set(naming.cost_slots(lay)) == set(nlp.cost_names) and len(nlp.cost_names) == len(set(nlp.cost_names))


def test_output_to_dict_nesting(model):
    """`Output.to_dict()` (problem.py:58-79) nests the flat names by their dots."""
    from hippopt_b200.solution import nest_by_dots

    lay = KinoLayout(model, KinoSettings(horizon=3))
    lam = np.arange(lay.m, dtype=float)
    flat = naming.constraint_multipliers(lay, lam)
    nested = nest_by_dots(flat)
    got = nested["system"]["contact_points"]["left[1]"]["p_dcc[2]"]
    assert np.array_equal(got, lam[naming.constraint_rows(lay)["system.contact_points.left[1].p_dcc[2]"]])
    assert "unitary_quaternion[1]" in nested


def test_output_to_dict_and_mat_file(model, tmp_path):
    """Output.to_dict() keys of the reference (problem.py:72-79) and the .mat dump of main_periodic_step.py:503-513."""
    from scipy.io import loadmat

    from hippopt_b200 import solution
    from hippopt_b200._capi import H
    from hippopt_b200.workloads import kino_batch

    lay = KinoLayout(model, KinoSettings(horizon=3, final_state_constraint=True))
    x, p, lam, _ = kino_batch(lay, model, 1, seed=3)
    terms = np.random.default_rng(0).uniform(size=(lay.N, H["HB_COST_TERMS"]))
    out = solution.make_output(lay, x[0], p[0], lam[0], 12.5, terms)
    d = out.to_dict()
    assert set(d) == {"values", "cost_value", "cost_values", "constraint_multipliers"}
    v = d["values"]
    assert len(v["system"]) == 3 and np.array_equal(v["system"][2]["kinematics"]["joints"]["positions"],
                                                    x[0, 189 * 2 + 157:189 * 2 + 180])
    assert np.array_equal(v["system"][1]["contact_points"]["right"][0]["f"], x[0, 189 + 15 * 4 + 9:189 + 15 * 4 + 12])
    assert np.array_equal(v["initial_state"]["centroidal_momentum"], x[0, 567:573])
    assert v["dt"] == p[0, lay.po.dt] and v["references"][1]["feet"]["desired_swing_height"] == p[0, lay.po.refs0 + 55 + 10]
    assert d["cost_values"]["joint_positions_error[2]"] == terms[2, H["HB_CT_JOINTS"]]
    assert d["cost_values"]["system"]["contact_points"]["left[3]"]["f_regularization[1]"] == terms[1, H["HB_CT_FRATIO0"] + 3]
    path = str(tmp_path / "plan.mat")
    solution.save_mat(path, out, guess=v)
    back = loadmat(path, simplify_cells=True)
    assert back["output"]["cost_value"] == 12.5
    assert np.allclose(back["output"]["values"]["system"][2]["com"], x[0, 189 * 2 + 180:189 * 2 + 183])
    assert np.allclose(back["output"]["constraint_multipliers"]["final_state_expression"],
                       lam[0][naming.constraint_rows(lay)["final_state_expression"]])
    assert back["output"]["cost_values"]["joint_positions_error_2"] == terms[2, H["HB_CT_JOINTS"]]


# ------------------------------------------------------------------------------------------------ pose finder
POSE_GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_pose_expression_names.json")


def test_pose_finder_names_match_the_reference_planner(model):
    """Static pose finder (config 2): a plain hp.Problem -- names carry no [k] suffix and live under `state.`; every
    name= argument of humanoid_pose_finder/planner.py that the default settings use is present exactly once per point."""
    from hippopt_b200.pose_layout import PoseLayout
    from hippopt_b200.workloads import pose_batch

    reference = {e["name"]: e for e in json.load(open(POSE_GOLD))["expressions"]}
    lay = PoseLayout(model)
    x, p, _, _ = pose_batch(lay, model, 1, seed=3)
    cons = naming.constraint_rows(lay)
    costs = naming.pose_cost_values(lay, model, x[0], p[0])
    assert sorted(np.concatenate(list(cons.values())).tolist()) == list(range(lay.m))  # every row of g has a name
    seen: dict[str, int] = {}
    for full in list(cons) + list(costs):
        assert "[" not in full.replace("left[", "").replace("right[", ""), full  # no horizon index
        base = re.sub(r"^state\.contact_points\.(left|right)\[\d\]\.", "<point>.", full)
        assert base in reference, f"{full}: {base} is not a name= argument of the reference pose finder"
        ref = reference[base]
        if ref["call"] == "add_cost":
            assert full in costs, full
        elif ref["call"] == "add_constraint":
            assert full in cons, full
        seen[base] = seen.get(base, 0) + 1
    # hands are skipped by the defaults of planner.py:79-92 (ExpressionType.skip); everything else is used
    assert set(seen) == set(reference) - {"left_hand_position_error", "right_hand_position_error"}
    for base, n in seen.items():
        assert n == (8 if base.startswith("<point>.") else 1), base


def test_pose_finder_named_costs_against_oracle(model):
    """`get_cost_values()` of the pose finder: the 28 named costs in recording order against the oracle's per-application
    values, and their sum against f."""
    from hippopt_b200.pose_layout import PoseLayout
    from hippopt_b200.workloads import pose_batch
    from oracle import pose_finder as pf

    lay = PoseLayout(model)
    x, p, _, _ = pose_batch(lay, model, 4, seed=9)
    nlp, _ = pf.build(model)
    terms, f = nlp.eval_cost_terms(x, p), nlp.eval_f(x, p)
    for b in range(4):
        vals = np.array(list(naming.pose_cost_values(lay, model, x[b], p[b]).values()))
        assert vals.shape == (28,)
        assert np.abs(vals - terms[b]).max() <= 1e-12 * np.abs(terms[b]).max()
        assert abs(vals.sum() - f[b]) <= 1e-12 * abs(f[b])
