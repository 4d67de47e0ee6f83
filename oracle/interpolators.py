"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's state interpolators
(/root/reference/src/hippopt/robot_planning/utilities/interpolators.py), one instance at a time, appending state
after state the way the reference does -- deliberately NOT the schedule-table formulation of the product
(hippopt_b200/interpolators.py + csrc/interp.cu), so that the two can disagree.

PARITY UNPINNED against the reference itself: interpolators.py imports casadi and liecasadi, neither of which is
installed here, and the reference ships no test or golden vector for these functions.  liecasadi (pinned by the
reference's setup.cfg as `liecasadi`, no version) supplies `Quaternion.slerp_step` and `SO3.act`; their
published definitions are restated below.  Anchors: the reference's own call site (main_periodic_step.py:367-451,
reproduced in tests/test_interpolators_cpu.py), closed-form properties (end points, unit norm, constant
angular rate) and an independent implementation: scipy's `Slerp` (same arc whenever q0 . q1 > 0) and
`Rotation.as_matrix` agree with `quaternion_slerp` / `rotation_matrix` to 1e-12 / 1e-14.

A state is a dict: p (8, 3), f (8, 3), base_position (3), base_quaternion (4), joints (n), com (3).
A phase is a dict: position (3), quaternion (4), mid_position / mid_quaternion (or None), force (3),
activation_time, deactivation_time (or None).
"""
from __future__ import annotations

import copy
import math

import numpy as np


def linear_interpolator(initial, final, number_of_points):
    """interpolators.py:24-50."""
    initial, final = np.asarray(initial, dtype=np.float64), np.asarray(final, dtype=np.float64)
    if initial.shape != final.shape:
        raise ValueError(f"Initial value has shape {initial.shape}, but final value has shape {final.shape}.")
    return [(1 - t) * initial + t * final for t in np.linspace(0.0, 1.0, number_of_points)]


def slerp_step(q1, q2, t):
    """liecasadi Quaternion.slerp_step [ext]: (sin((1 - t) a) q1 + sin(t a) q2) / sin(a), a = acos(q1 . q2)."""
    angle = math.acos(float(np.dot(q1, q2)))
    return (math.sin((1.0 - t) * angle) * q1 + math.sin(t * angle) * q2) / math.sin(angle)


def quaternion_slerp(initial, final, number_of_points):
    """interpolators.py:53-77: slerp unless the angle is below 1e-6 (then, or when acos fails, the initial one)."""
    initial, final = np.asarray(initial, dtype=np.float64), np.asarray(final, dtype=np.float64)
    dot = float(np.dot(initial, final))
    angle = math.acos(dot) if abs(dot) <= 1.0 else math.nan
    out = []
    for t in np.linspace(0.0, 1.0, number_of_points):
        out.append(slerp_step(initial, final, float(t)) if abs(angle) > 1e-6 else initial.copy())
    return out


def transform_interpolator(initial, final, number_of_points):
    """interpolators.py:80-103; a transform is (translation, quaternion)."""
    lin = linear_interpolator(initial[0], final[0], number_of_points)
    rot = quaternion_slerp(initial[1], final[1], number_of_points)
    return [(lin[i], rot[i]) for i in range(number_of_points)]


def rotation_matrix(q):
    """liecasadi SO3.as_matrix [ext] = SURVEY.md A.1: I + 2 w [v]x + 2 [v]x^2."""
    x, y, z, w = q
    S = np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])
    return np.eye(3) + 2.0 * w * S + 2.0 * S @ S


def foot_state(descriptor, transform, force):
    """FootContactState.from_parent_frame_transform (variables/contacts.py:103-127) + the force assignment of
    append_stance_phase / append_swing_phase: (p (n_pts, 3), f (n_pts, 3))."""
    R = rotation_matrix(transform[1])
    p = np.stack([transform[0] + R @ np.asarray(d, dtype=np.float64) for d in descriptor])
    return p, np.tile(np.asarray(force, dtype=np.float64), (len(descriptor), 1))


def foot_contact_state_interpolator(phases, descriptor, number_of_points, dt, t0=0.0):
    """interpolators.py:106-309."""
    assert len(phases) > 0 and number_of_points > 0 and dt > 0.0
    end_time = t0 + dt * number_of_points
    ph = copy.deepcopy(phases)
    first, final = ph[0], ph[-1]
    if first["activation_time"] is None:
        d = first["deactivation_time"] if first["deactivation_time"] is not None else t0
        first["activation_time"] = min(d, t0) - dt
    if first["activation_time"] > t0:
        raise ValueError("The first phase activation time is after the start time.")
    if any(q["activation_time"] is None for q in ph):
        raise ValueError("A phase has no activation time, but is not the first phase.")
    if final["deactivation_time"] is None:
        final["deactivation_time"] = max(end_time, final["activation_time"]) + dt
    if final["deactivation_time"] < end_time:
        raise ValueError("The Last phase deactivation time is before the end time.")
    for n, q in enumerate(ph):
        if q["deactivation_time"] is None:
            raise ValueError("A phase has no deactivation time, but is not the last phase.")
        if q["activation_time"] > q["deactivation_time"]:
            raise ValueError("A phase has an activation time greater than its deactivation time.")
        if n + 1 < len(ph) and q["deactivation_time"] > ph[n + 1]["activation_time"]:
            raise ValueError("A phase has a deactivation time greater than the activation time of the next phase.")

    out = []

    def stance(q, points):
        for _ in range(points):
            out.append(foot_state(descriptor, (q["position"], q["quaternion"]), q["force"]))

    def swing(a, b, points):
        full = int(np.ceil((b["activation_time"] - a["deactivation_time"]) / dt))
        if a["mid_position"] is None:
            a["mid_position"] = (np.asarray(a["position"]) + np.asarray(b["position"])) / 2
            a["mid_quaternion"] = b["quaternion"]
        mid = (a["mid_position"], a["mid_quaternion"])
        n_up = min(round(full / 2), points)
        for tr in transform_interpolator((a["position"], a["quaternion"]), mid, n_up):
            out.append(foot_state(descriptor, tr, np.zeros(3)))
        if points - n_up == 0:
            return
        for tr in transform_interpolator(mid, (b["position"], b["quaternion"]), points - n_up):
            out.append(foot_state(descriptor, tr, np.zeros(3)))

    if len(ph) == 1 or first["deactivation_time"] >= end_time:
        stance(first, number_of_points)
        return out
    i, activation = 0, first["activation_time"]
    while activation < t0:
        if ph[i]["deactivation_time"] > t0:
            break
        i += 1
        activation = ph[i]["activation_time"]
    if activation > t0:
        new_t0 = ph[i - 1]["deactivation_time"] - dt
        advance = int(np.ceil((t0 - new_t0) / dt))
        longer = foot_contact_state_interpolator(ph, descriptor, number_of_points + advance, dt, new_t0)
        return longer[advance:]
    remaining = number_of_points
    guard = 0
    while i < len(ph) - 1:
        guard += 1
        assert guard < 10 * number_of_points + 10, "the reference would loop forever on these phases"
        a, b = ph[i], ph[i + 1]
        n_st = min(int(np.ceil((a["deactivation_time"] - max(a["activation_time"], t0)) / dt)), remaining)
        stance(a, n_st)
        remaining -= n_st
        if remaining == 0:
            return out
        n_sw = min(int(np.ceil((b["activation_time"] - a["deactivation_time"]) / dt)), remaining)
        if n_sw == 0:
            continue
        swing(a, b, n_sw)
        remaining -= n_sw
        if remaining == 0:
            return out
        i += 1
    stance(final, remaining)
    return out


def humanoid_state_interpolator(initial_state, final_state, contact_phases, contact_descriptor, number_of_points, dt,
                                t0=0.0):
    """interpolators.py:396-448 with feet_contact_points_interpolator (:312-337),
    floating_base_system_state_interpolator (:340-393) and the CoM (:419-423).
    contact_phases = (left phases, right phases), contact_descriptor = (left points, right points)."""
    left = foot_contact_state_interpolator(contact_phases[0], contact_descriptor[0], number_of_points, dt, t0)
    right = foot_contact_state_interpolator(contact_phases[1], contact_descriptor[1], number_of_points, dt, t0)
    assert len(left) == len(right) == number_of_points
    pos = linear_interpolator(initial_state["base_position"], final_state["base_position"], number_of_points)
    quat = quaternion_slerp(initial_state["base_quaternion"], final_state["base_quaternion"], number_of_points)
    if len(initial_state["joints"]) != len(final_state["joints"]):
        raise ValueError("Initial and final state have a different number of joints.")
    joints = linear_interpolator(initial_state["joints"], final_state["joints"], number_of_points)
    com = linear_interpolator(initial_state["com"], final_state["com"], number_of_points)
    return [{"p": np.concatenate([left[k][0], right[k][0]]), "f": np.concatenate([left[k][1], right[k][1]]),
             "base_position": pos[k], "base_quaternion": quat[k], "joints": joints[k], "com": com[k]}
            for k in range(number_of_points)]


def state_block(state, descriptor):
    """The product's (82 + n_joints) state block (hippopt_b200/kino_layout.py ParamOffsets.st_pt / ST_*)."""
    d = np.concatenate([np.asarray(descriptor[0], dtype=np.float64), np.asarray(descriptor[1], dtype=np.float64)])
    pts = np.concatenate([state["p"], state["f"], d], axis=1).ravel()
    return np.concatenate([pts, state["base_position"], state["base_quaternion"], state["joints"], state["com"]])
