"""Solution I/O of the kinodynamic planner.

What `Problem.solve()` returns in the reference is an `Output` (`/root/reference/src/hippopt/base/problem.py:28-79`):
the solved `Variables` tree, the cost value, one value per named cost and one multiplier array per named
constraint, and `Output.to_dict()` is what the mains dump to a `.mat` file
(`turnkey_planners/humanoid_kinodynamic/main_periodic_step.py:503-513`).  Here the same structure is rebuilt for
one instance of a batched solve from the flat vectors the device works on: `x` / `p` are cut along SURVEY.md
Appendix B.1 / B.2 into the field names of `variables.py:121-301`, names come from hippopt_b200/naming.py.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import naming
from .kino_layout import COM, FD, F, H as ZH, NJ, NPT, NZ, P, PB, Q, QD, S, SD, U, V, VB


def nest_by_dots(flat: dict) -> dict:
    """`Output.to_dict`'s flatten_to_nested_dict (problem.py:58-71): "a.b.c" -> d["a"]["b"]["c"]."""
    nested: dict = {}
    for key, value in flat.items():
        keys = key.split(".")
        assert all(isinstance(k, str) and len(k) > 0 for k in keys)
        d = nested
        for k in keys[:-1]:
            d = d.setdefault(k, {})
        d[keys[-1]] = value
    return nested


def _point(z, i, desc):
    o = 15 * i
    return {"p": z[o + P:o + P + 3].copy(), "f": z[o + F:o + F + 3].copy(), "v": z[o + V:o + V + 3].copy(),
            "f_dot": z[o + FD:o + FD + 3].copy(), "u_v": z[o + U:o + U + 3].copy(),
            "descriptor": {"position_in_foot_frame": desc.copy()}}


def _state_block(blk):
    """105-double initial / final state block (Appendix B.2): 8 x (p, f, descriptor), base position, quaternion,
    joints, com."""
    pts = [{"p": blk[9 * i:9 * i + 3].copy(), "f": blk[9 * i + 3:9 * i + 6].copy(),
            "descriptor": {"position_in_foot_frame": blk[9 * i + 6:9 * i + 9].copy()}} for i in range(NPT)]
    return {"contact_points": {"left": pts[:4], "right": pts[4:]},
            "kinematics": {"base": {"position": blk[72:75].copy(), "quaternion_xyzw": blk[75:79].copy()},
                           "joints": {"positions": blk[79:79 + NJ].copy()}},
            "com": blk[102:105].copy()}


def pose_values_dict(layout, x: np.ndarray, p: np.ndarray) -> dict:
    """The pose finder's `Variables` tree (`humanoid_pose_finder/planner.py:226-300`): `state` (a HumanoidState: contact
    points with p, f and descriptor, base pose, joints, com) from x, the parameters and `references` from p
    (SURVEY.md Appendix B.4)."""
    x, p = np.asarray(x, dtype=np.float64).ravel(), np.asarray(p, dtype=np.float64).ravel()
    assert x.shape == (layout.n_x,) and p.shape == (layout.n_p,)
    po = layout.po
    desc = p[po.desc0:po.desc0 + 24].reshape(NPT, 3)
    pts = [{"p": x[6 * i:6 * i + 3].copy(), "f": x[6 * i + 3:6 * i + 6].copy(),
            "descriptor": {"position_in_foot_frame": desc[i].copy()}} for i in range(NPT)]
    state = {"contact_points": {"left": pts[:4], "right": pts[4:]},
             "kinematics": {"base": {"position": x[48:51].copy(), "quaternion_xyzw": x[51:55].copy()},
                            "joints": {"positions": x[55:55 + NJ].copy()}},
             "com": x[78:81].copy()}
    return {
        "state": state, "mass": p[po.mass], "parametric_link_length_multipliers": p[po.plm],
        "parametric_link_densities": p[po.pld], "gravity": p[po.gravity:po.gravity + 6].copy(),
        "references": {"state": _state_block(p[po.ref:po.ref + 105]),
                       "frame_quaternion_xyzw": p[po.ref_fq:po.ref_fq + 4].copy(),
                       "left_hand_position": p[po.ref_lhand:po.ref_lhand + 3].copy(),
                       "right_hand_position": p[po.ref_rhand:po.ref_rhand + 3].copy()},
        "relaxed_complementarity_epsilon": p[po.eps], "static_friction": p[po.mu],
        "maximum_joint_positions": p[po.max_s:po.max_s + NJ].copy(),
        "minimum_joint_positions": p[po.min_s:po.min_s + NJ].copy(),
        "left_hand_position_in_frame": p[po.lhand_in_frame:po.lhand_in_frame + 3].copy(),
        "right_hand_position_in_frame": p[po.rhand_in_frame:po.rhand_in_frame + 3].copy(),
    }


def values_dict(layout, x: np.ndarray, p: np.ndarray) -> dict:
    """`output.values.to_dict(flatten=False)`: the `Variables` tree (variables.py:254-301) as nested dicts / lists,
    variables taken from x and parameters from p."""
    if naming.is_pose_layout(layout):
        return pose_values_dict(layout, x, p)
    if not hasattr(layout, "po"):  # a template without a field table (toy OCP): the flat vectors
        return {"x": np.asarray(x, dtype=np.float64).copy(), "p": np.asarray(p, dtype=np.float64).copy()}
    x, p = np.asarray(x, dtype=np.float64).ravel(), np.asarray(p, dtype=np.float64).ravel()
    assert x.shape == (layout.n_x,) and p.shape == (layout.n_p,)
    po = layout.po
    system, refs = [], []
    for k in range(layout.N):
        z = x[NZ * k:NZ * (k + 1)]
        desc = p[po.desc0 + 24 * k:po.desc0 + 24 * (k + 1)].reshape(NPT, 3)
        pts = [_point(z, i, desc[i]) for i in range(NPT)]
        system.append({
            "contact_points": {"left": pts[:4], "right": pts[4:]},
            "kinematics": {"base": {"position": z[PB:PB + 3].copy(), "quaternion_xyzw": z[Q:Q + 4].copy(),
                                    "linear_velocity": z[VB:VB + 3].copy(),
                                    "quaternion_velocity_xyzw": z[QD:QD + 4].copy()},
                           "joints": {"positions": z[S:S + NJ].copy(), "velocities": z[SD:SD + NJ].copy()}},
            "com": z[COM:COM + 3].copy(), "centroidal_momentum": z[ZH:ZH + 6].copy()})
        r = p[po.refs0 + 55 * k:po.refs0 + 55 * (k + 1)]
        refs.append({
            "feet": {"left": {"points": [{"desired_force_ratio": r[po.R_RATIO_L + i]} for i in range(4)],
                              "yaw": r[po.R_YAW_L]},
                     "right": {"points": [{"desired_force_ratio": r[po.R_RATIO_R + i]} for i in range(4)],
                               "yaw": r[po.R_YAW_R]},
                     "desired_swing_height": r[po.R_SWING],
                     "centroid_weights": r[po.R_CW:po.R_CW + 3].copy(), "centroid": r[po.R_CC:po.R_CC + 3].copy()},
            "com_linear_velocity": r[po.R_COMV:po.R_COMV + 3].copy(),
            "desired_frame_quaternion_xyzw": r[po.R_FQ:po.R_FQ + 4].copy(),
            "base_quaternion_xyzw": r[po.R_BQ:po.R_BQ + 4].copy(),
            "base_quaternion_xyzw_velocity": r[po.R_BQV:po.R_BQV + 4].copy(),
            "joint_regularization": r[po.R_JR:po.R_JR + NJ].copy()})
    init = _state_block(p[po.init:po.init + 105])
    init["centroidal_momentum"] = x[layout.h_init:layout.h_init + 6].copy()  # a variable (variables.py:240)
    out = {
        "system": system, "mass": p[po.mass], "parametric_link_length_multipliers": p[po.plm],
        "parametric_link_densities": p[po.pld], "initial_state": init, "final_state": _state_block(p[po.final:po.final + 105]),
        "dt": p[po.dt], "gravity": p[po.gravity:po.gravity + 6].copy(), "planar_dcc_height_multiplier": p[po.kt],
        "dcc_gain": p[po.k_bs], "dcc_epsilon": p[po.eps], "static_friction": p[po.mu],
        "maximum_velocity_control": p[po.max_u:po.max_u + 3].copy(),
        "maximum_force_derivative": p[po.max_fd:po.max_fd + 3].copy(), "maximum_angular_momentum": p[po.max_L],
        "minimum_com_height": p[po.min_com_h], "minimum_feet_lateral_distance": p[po.min_feet_d],
        "maximum_feet_relative_height": p[po.max_feet_h],
        "maximum_joint_positions": p[po.max_s:po.max_s + NJ].copy(), "minimum_joint_positions": p[po.min_s:po.min_s + NJ].copy(),
        "maximum_joint_velocities": p[po.max_sd:po.max_sd + NJ].copy(),
        "minimum_joint_velocities": p[po.min_sd:po.min_sd + NJ].copy(), "references": refs,
    }
    return out


@dataclasses.dataclass
class Output:
    """One instance's `hippopt.Output` (problem.py:28-79)."""

    values: dict                                   # nested Variables tree (values_dict)
    cost_value: float
    cost_values: dict                              # {expression name: value}
    constraint_multipliers: dict                   # {expression name: multipliers}

    def to_dict(self) -> dict:
        return {"values": self.values, "cost_value": self.cost_value, "cost_values": nest_by_dots(self.cost_values),
                "constraint_multipliers": nest_by_dots(self.constraint_multipliers)}


def make_output(layout, x, p, lam_g, cost_value, cost_terms) -> Output:
    """Assemble the `Output` of one instance: x / p / lam_g are that instance's vectors, `cost_terms` its
    [N][HB_COST_TERMS] table from `KinoEvaluator.cost_terms` (hb_eval_cost_terms)."""
    return Output(values=values_dict(layout, x, p), cost_value=float(cost_value),
                  cost_values=naming.cost_values(layout, cost_terms),
                  constraint_multipliers=naming.constraint_multipliers(layout, lam_g))


def _mat_safe(obj):
    """scipy.io.savemat needs MATLAB field names ([A-Za-z][A-Za-z0-9_]*, at most 31 characters are kept by old MAT
    versions; the reference uses hdf5storage, which maps arbitrary keys) and no empty lists: "[k]" -> "_k",
    "{j}" -> "_j"; lists of dicts become MATLAB cell arrays (object arrays)."""
    import re

    if isinstance(obj, dict):
        out = {}
        for k, v in obj.items():
            key = re.sub(r"[\[\{](\d+)[\]\}]", r"_\1", str(k))
            key = re.sub(r"[^A-Za-z0-9_]", "_", key)
            if not key[0].isalpha():
                key = "x" + key
            out[key] = _mat_safe(v)
        return out
    if isinstance(obj, (list, tuple)):
        arr = np.empty(len(obj), dtype=object)
        for i, v in enumerate(obj):
            arr[i] = _mat_safe(v)
        return arr
    return np.asarray(obj, dtype=np.float64)


def save_mat(path: str, output: Output, guess: dict | None = None) -> None:
    """The dump of main_periodic_step.py:503-513: {"output": output.to_dict(), "guess": guess tree} as a .mat file
    (MATLAB v5 through scipy; long_field_names for the expression names)."""
    from scipy.io import savemat

    mdict = {"output": _mat_safe(output.to_dict())}
    if guess is not None:
        mdict["guess"] = _mat_safe(guess)
    savemat(path, mdict, long_field_names=True, do_compression=True)
