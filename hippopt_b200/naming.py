"""Expression names of the kinodynamic planner, attached to the rows of g and to the cost slots of the kernels.

`OptiSolver.solve()` hands back one value per NAMED cost and one multiplier array per NAMED constraint
(`/root/reference/src/hippopt/base/opti_solver.py:522-537`); `Output.to_dict()` (`base/problem.py:58-79`) nests them
by the dots in the name.  The names are built by three layers of the reference:

* the planner's `name=` arguments (`turnkey_planners/humanoid_kinodynamic/planner.py:215-930`); contact-point
  expressions use the flattened symbol name of the point's variable, e.g. ``system.contact_points.left[2].p``
  (`base/multiple_shooting_solver.py:292-337`, list elements get ``[i]``: `:395-401`);
* `add_expression_to_horizon` appends ``[k]`` (`:810`), `add_dynamics` ``[0]`` for the initial condition (`:696-697`)
  and ``[k]`` for the defect between knots k-1 and k (`:728`);
* `Problem.add_constraint / add_cost` append ``{j}`` when the expression is a generator (`base/problem.py:105-110,
  140-145`): the dynamics are generators over their state variables -- one per `add_dynamics` call of the planner.

Pure index arithmetic on `KinoLayout` (no kernels): `constraint_rows` maps a name to its rows of g / lam_g,
`cost_slots` to (knot, slot) of `hb_eval_cost_terms`.
"""
from __future__ import annotations

import numpy as np

from ._capi import H

NPT = 8


def point_symbol(i: int) -> str:
    """Flattened symbol prefix of contact point i (0-3 left foot, 4-7 right)."""
    return f"system.contact_points.{'left' if i < 4 else 'right'}[{i % 4}]"


# layout family -> (reference base name, kind); kind "dyn0": initial condition of add_dynamics, "dyn": its defects,
# "horizon": add_expression_to_horizon, "single": Problem.add_expression called once
_POINT_FAMILIES = [
    ("f_ic", "{pt}.f_dynamics", "dyn0"), ("f_dyn", "{pt}.f_dynamics", "dyn"),
    ("p_ic", "{pt}.p_dynamics", "dyn0"), ("p_dyn", "{pt}.p_dynamics", "dyn"),
    ("planar", "{pt}.p_planar_complementarity", "horizon"), ("dcc", "{pt}.p_dcc", "horizon"),
    ("height", "{pt}.p_height", "horizon"), ("normal", "{pt}.f_normal", "horizon"),
    ("friction", "{pt}.f_friction", "horizon"), ("u_bounds", "{pt}.u_v_bounds", "horizon"),
    ("fd_bounds", "{pt}.f_dot_bounds", "horizon"), ("fk", "{pt}.p_kinematics_consistency", "horizon"),
]
_ROBOT_FAMILIES = [
    ("pb_ic", "base_position_dynamics", "dyn0"), ("pb_dyn", "base_position_dynamics", "dyn"),
    ("q_ic", "base_quaternion_dynamics", "dyn0"), ("q_dyn", "base_quaternion_dynamics", "dyn"),
    ("s_ic", "joint_position_dynamics", "dyn0"), ("s_dyn", "joint_position_dynamics", "dyn"),
    ("com_ic", "com_dynamics", "dyn0"), ("com_dyn", "com_dynamics", "dyn"),
    ("h_ic", "centroidal_momentum_dynamics", "dyn0"), ("h_dyn", "centroidal_momentum_dynamics", "dyn"),
    ("unit_quat", "unitary_quaternion", "horizon"), ("com_kin", "com_kinematics_consistency", "horizon"),
    ("mom_kin", "centroidal_momentum_kinematics_consistency", "horizon"),
    ("L_bounds", "angular_momentum_bounds", "horizon"), ("com_height", "minimum_com_height", "horizon"),
    ("feet_dist", "minimum_feet_distance", "horizon"), ("s_bounds", "joint_position_bounds", "horizon"),
    ("sd_bounds", "joint_velocity_bounds", "horizon"), ("final", "final_state_expression", "single"),
    ("feet_relh", "maximum_feet_relative_height", "horizon"), ("periodicity", "periodicity_expression", "single"),
]


def _full_name(base: str, kind: str, k: int) -> str:
    if kind == "dyn0":
        return f"{base}[0]{{0}}"
    if kind == "dyn":
        return f"{base}[{k}]{{0}}"
    if kind == "horizon":
        return f"{base}[{k}]"
    return base


def constraint_rows(layout) -> dict[str, np.ndarray]:
    """name -> rows of g (and of lam_g), in the reference's `subject_to` order (= increasing row index)."""
    if is_pose_layout(layout):
        return _pose_constraint_rows(layout)
    if not hasattr(layout, "fam"):  # a template without a name table (toy OCP): one anonymous block
        return {"g": np.arange(layout.m)}
    out: dict[str, np.ndarray] = {}
    fams = [(f"pt{i}.{fam}", base.format(pt=point_symbol(i)), kind) for i in range(NPT)
            for fam, base, kind in _POINT_FAMILIES]
    fams += list(_ROBOT_FAMILIES)
    for fam, base, kind in fams:
        if fam not in layout.fam:
            continue
        first, rows, k0, k1 = layout.fam[fam]
        for k in range(k0, k1 + 1):
            name = _full_name(base, kind, k)
            assert name not in out, name
            out[name] = np.arange(first + (k - k0) * rows, first + (k - k0 + 1) * rows)
    # recording order of the reference = row order
    return dict(sorted(out.items(), key=lambda kv: kv[1][0]))


# cost slot (include/hippopt_b200.h, HB_CT_*) -> (reference base name, first knot)
def _cost_table():
    t = []
    for i in range(NPT):
        pt = point_symbol(i)
        t += [(H["HB_CT_SWING0"] + i, f"{pt}.p_swing_height_regularization", 1),
              (H["HB_CT_UV0"] + i, f"{pt}.u_v_regularization", 1),
              (H["HB_CT_FDOT0"] + i, f"{pt}.f_dot_regularization", 1),
              (H["HB_CT_FRATIO0"] + i, f"{pt}.f_regularization", 1)]
    t += [(H["HB_CT_COM_VELOCITY"], "com_velocity_error", 0), (H["HB_CT_CENTROID"], "contacts_centroid_cost", 1),
          (H["HB_CT_YAW_LEFT"], "left_yaw_regularization", 1), (H["HB_CT_YAW_RIGHT"], "right_yaw_regularization", 1),
          (H["HB_CT_FRAME_QUAT"], "frame_quaternion_error", 1), (H["HB_CT_BASE_QUAT"], "base_quaternion_error", 1),
          (H["HB_CT_BASE_QUAT_VEL"], "base_quaternion_velocity_error", 0), (H["HB_CT_JOINTS"], "joint_positions_error", 1)]
    return t


def cost_names(layout) -> list[str]:
    """Names of the template's cost expressions, in recording order."""
    if is_pose_layout(layout):
        names = ["base_quaternion_error", "frame_rotation_error", "com_position_error", "joint_positions_error"]
        for foot in range(2):
            pts = [pose_point_symbol(4 * foot + i) for i in range(4)]
            names += [f"{pt}.f_average_regularization" for pt in pts]
            for pt in pts:
                names += [f"{pt}.p_regularization", f"{pt}.f_regularization"]
        return names
    if not hasattr(layout, "fam"):
        return []
    return list(cost_slots(layout))


def cost_slots(layout) -> dict[str, tuple[int, int]]:
    """name -> (knot, slot) into the [N][HB_COST_TERMS] table `hb_eval_cost_terms` writes."""
    out = {}
    for slot, base, k0 in _cost_table():
        for k in range(k0, layout.N):
            out[f"{base}[{k}]"] = (k, slot)
    return out


def cost_values(layout, terms: np.ndarray) -> dict[str, float]:
    """One instance's [N][HB_COST_TERMS] table -> {name: value} (`OptiSolver.get_cost_values`)."""
    terms = np.asarray(terms).reshape(layout.N, H["HB_COST_TERMS"])
    return {name: float(terms[k, s]) for name, (k, s) in cost_slots(layout).items()}


def constraint_multipliers(layout, lam_g: np.ndarray) -> dict[str, np.ndarray]:
    """One instance's lam_g -> {name: multipliers} (`OptiSolver.get_constraint_multipliers`)."""
    lam_g = np.asarray(lam_g).ravel()
    return {name: lam_g[rows].copy() for name, rows in constraint_rows(layout).items()}


# ------------------------------------------------------------------------------------------------ pose finder
# The static pose finder (`turnkey_planners/humanoid_pose_finder/planner.py`) is a plain `hp.Problem`: names carry no
# "[k]" suffix, and its variables live under `state` (`Variables.state: HumanoidState`, planner.py:226-229).
def pose_point_symbol(i: int) -> str:
    return f"state.contact_points.{'left' if i < 4 else 'right'}[{i % 4}]"


_POSE_POINT_FAMILIES = [("complementarity", "{pt}.p_complementarity"), ("height", "{pt}.p_height"),
                        ("normal", "{pt}.f_normal"), ("friction", "{pt}.f_friction"),
                        ("fk", "{pt}.p_kinematics_consistency")]
_POSE_ROBOT_FAMILIES = [("unit_quat", "unitary_quaternion"), ("com_kin", "com_kinematics_consistency"),
                        ("balance", "centroidal_momentum_dynamics"), ("s_bounds", "joint_position_bounds")]


def is_pose_layout(layout) -> bool:
    return getattr(layout, "N", None) == 1 and "balance" in getattr(layout, "fam", {})


def _pose_constraint_rows(layout) -> dict[str, np.ndarray]:
    out = {}
    for i in range(NPT):
        for fam, base in _POSE_POINT_FAMILIES:
            first, rows, _, _ = layout.fam[f"pt{i}.{fam}"]
            out[base.format(pt=pose_point_symbol(i))] = np.arange(first, first + rows)
    for fam, base in _POSE_ROBOT_FAMILIES:
        first, rows, _, _ = layout.fam[fam]
        out[base] = np.arange(first, first + rows)
    return dict(sorted(out.items(), key=lambda kv: kv[1][0]))


def _rot_from_quat(q):
    """R = I + 2 w [v]x + 2 [v]x^2 of an xyzw quaternion (as given: callers normalise where the planner does)."""
    v, w = q[:3], q[3]
    S = np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
    return np.eye(3) + 2.0 * w * S + 2.0 * S @ S


def _frame_rotation(model, frame: str, qn, s):
    """Orientation of a frame in the inertial frame: base rotation, then joint_rot @ Rodrigues(axis, s) down the chain."""
    body, R_f, _ = model.frames[frame]
    R = _rot_from_quat(qn)
    for b in model.chain_to_root(body):
        a = model.joint_axis[b]
        K = np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])
        th = s[b - 1]
        R = R @ model.joint_rot[b] @ (np.eye(3) + np.sin(th) * K + (1.0 - np.cos(th)) * K @ K)
    return R @ R_f


def pose_cost_values(layout, model, x, p) -> dict[str, float]:
    """Values of the pose finder's named costs at one solution, in the order the planner records them
    (planner.py:523-594, 751-788): `OptiSolver.get_cost_values()` of that problem.  Evaluated on the host from the
    solution (a report made once per solve, not part of the evaluation path); their sum is the objective the kernels
    return."""
    from .pose_layout import XCOM, XQ, XS  # local: naming has no other dependency on the layouts

    x, p = np.asarray(x, dtype=np.float64).ravel(), np.asarray(p, dtype=np.float64).ravel()
    st, po = layout.st, layout.po
    q, s, com = x[XQ:XQ + 4], x[XS:XS + len(model.joint_names)], x[XCOM:XCOM + 3]
    ref = p[po.ref:po.ref + 105]
    qd, s_ref, com_ref = ref[po.ST_Q:po.ST_Q + 4], ref[po.ST_S:po.ST_S + len(s)], ref[po.ST_COM:po.ST_COM + 3]
    out: dict[str, float] = {}
    # base quaternion: (qd^-1 (x) q) - identity, xyzw Hamilton product with the conjugate of qd (quaternion.py:54-85)
    a = np.array([-qd[0], -qd[1], -qd[2], qd[3]])
    av, aw, bv, bw = a[:3], a[3], q[:3], q[3]
    e = np.concatenate([aw * bv + bw * av + np.cross(av, bv), [aw * bw - av @ bv - 1.0]])
    out["base_quaternion_error"] = st.base_quaternion_cost_multiplier * float(e @ e)
    # frame orientation: (trace(R_frame R(q_desired)^T) - 3)^2 with the normalised base quaternion (kinematics.py:444-448)
    fq = p[po.ref_fq:po.ref_fq + 4]
    E = _frame_rotation(model, st.frame_quaternion_cost_frame, q / np.linalg.norm(q), s) @ _rot_from_quat(fq).T
    out["frame_rotation_error"] = st.desired_frame_quaternion_cost_multiplier * float((np.trace(E) - 3.0) ** 2)
    out["com_position_error"] = st.com_regularization_cost_multiplier * float((com - com_ref) @ (com - com_ref))
    es = s - s_ref
    out["joint_positions_error"] = st.joint_regularization_cost_multiplier * float(
        es @ (np.asarray(st.joint_regularization_cost_weights) * es))
    for foot in range(2):
        pts = [4 * foot + i for i in range(4)]
        forces = [x[6 * i + 3:6 * i + 6] for i in pts]
        mean = 0.25 * (forces[0] + forces[1] + forces[2] + forces[3])
        for i, f in zip(pts, forces):
            out[f"{pose_point_symbol(i)}.f_average_regularization"] = (
                st.average_force_regularization_cost_multiplier * float((f - mean) @ (f - mean)))
        for i, f in zip(pts, forces):
            pt = pose_point_symbol(i)
            dp = x[6 * i:6 * i + 3] - ref[9 * i:9 * i + 3]
            df = f - ref[9 * i + 3:9 * i + 6]
            out[f"{pt}.p_regularization"] = st.point_position_regularization_cost_multiplier * float(dp @ dp)
            out[f"{pt}.f_regularization"] = st.force_regularization_cost_multiplier * float(df @ df)
    return out
