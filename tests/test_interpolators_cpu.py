"""Row f3 on the CPU: the oracle's interpolators (oracle/interpolators.py) against closed-form properties, and the
product's host-side phase bookkeeping (hippopt_b200.interpolators.foot_contact_schedule, one table for the whole
batch) against the oracle's append-as-you-go restatement of
/root/reference/src/hippopt/robot_planning/utilities/interpolators.py:106-309 on the reference's own call site
(turnkey_planners/humanoid_kinodynamic/main_periodic_step.py:367-451) and on randomised phase timings."""
import numpy as np
import pytest

from hippopt_b200.interpolators import (SWING_DOWN, SWING_UP, FootContactPhaseDescriptor, foot_contact_schedule)
from oracle import interpolators as oi

DESCRIPTOR = [[0.08, 0.03, 0.0], [0.08, -0.03, 0.0], [-0.08, -0.03, 0.0], [-0.08, 0.03, 0.0]]


def rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def periodic_step_phases(horizon, step_length=0.6):
    """main_periodic_step.py:367-412 as oracle phase dicts: (left, right)."""
    ident = np.array([0.0, 0.0, 0.0, 1.0])
    f = np.array([0.0, 0.0, 100.0])

    def phase(pos, mid, act, dea):
        return {"position": np.array(pos), "quaternion": ident, "mid_position": None if mid is None else np.array(mid),
                "mid_quaternion": None if mid is None else ident, "force": f, "activation_time": act,
                "deactivation_time": dea}

    left = [phase([0.0, 0.1, 0.0], [step_length / 2, 0.1, 0.05], None, horizon / 6.0),
            phase([step_length, 0.1, 0.0], None, horizon / 3.0, None)]
    right = [phase([step_length / 2, -0.1, 0.0], [step_length, -0.1, 0.05], None, horizon * 2.0 / 3.0),
             phase([1.5 * step_length, -0.1, 0.0], None, horizon * 5.0 / 6.0, None)]
    return left, right


def to_product(phases):
    return [FootContactPhaseDescriptor(position=p["position"], quaternion_xyzw=p["quaternion"],
                                       mid_swing_position=p["mid_position"], mid_swing_quaternion_xyzw=p["mid_quaternion"],
                                       force=p["force"], activation_time=p["activation_time"],
                                       deactivation_time=p["deactivation_time"]) for p in phases]


def evaluate_schedule(table, phases):
    """Numpy evaluation of a schedule table with the oracle's transform_interpolator (what csrc/interp.cu does)."""
    out = []
    for kind, a, b, j, n in table.tolist():
        A, B = phases[a], phases[b]
        if kind in (SWING_UP, SWING_DOWN):
            if A["mid_position"] is None:
                mid = ((A["position"] + B["position"]) / 2, B["quaternion"])
            else:
                mid = (A["mid_position"], A["mid_quaternion"])
            ends = ((A["position"], A["quaternion"]), mid) if kind == SWING_UP else (mid, (B["position"], B["quaternion"]))
            out.append(oi.foot_state(DESCRIPTOR, oi.transform_interpolator(*ends, n)[j], np.zeros(3)))
        else:
            out.append(oi.foot_state(DESCRIPTOR, (A["position"], A["quaternion"]), A["force"]))
    return out


def same(a, b):
    assert len(a) == len(b)
    for (pa, fa), (pb, fb) in zip(a, b):
        assert np.array_equal(pa, pb) and np.array_equal(fa, fb)


def test_slerp_is_the_great_arc_at_constant_rate():
    rng = np.random.default_rng(0)
    for _ in range(20):
        q0, q1 = rand_quat(rng), rand_quat(rng)
        qs = oi.quaternion_slerp(q0, q1, 9)
        assert np.allclose(qs[0], q0, atol=1e-15) and np.allclose(qs[-1], q1, atol=1e-14)
        assert np.allclose([np.linalg.norm(q) for q in qs], 1.0, atol=1e-14)
        total = np.arccos(np.dot(q0, q1))
        steps = [np.arccos(np.clip(np.dot(qs[i], qs[i + 1]), -1, 1)) for i in range(8)]
        assert np.allclose(steps, total / 8, atol=1e-7)  # acos near 1 loses half the digits
    # below the reference's 1e-6 threshold, and for a dot product rounded above 1: the initial quaternion
    q0 = rand_quat(rng)
    assert all(np.array_equal(q, q0) for q in oi.quaternion_slerp(q0, q0, 4))
    assert all(np.array_equal(q, q0) for q in oi.quaternion_slerp(q0, q0 * (1 + 1e-15), 3))


def test_linear_interpolator_end_points_and_shape_error():
    a, b = np.arange(5.0), np.arange(5.0) ** 2
    out = oi.linear_interpolator(a, b, 7)
    assert np.array_equal(out[0], a) and np.array_equal(out[-1], b) and np.allclose(out[3], (a + b) / 2)
    assert len(oi.linear_interpolator(a, b, 1)) == 1 and np.array_equal(oi.linear_interpolator(a, b, 1)[0], a)
    with pytest.raises(ValueError):
        oi.linear_interpolator(a, b[:4], 3)


def test_rotation_matrix_is_a_rotation():
    rng = np.random.default_rng(1)
    R = oi.rotation_matrix(rand_quat(rng))
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-14) and np.isclose(np.linalg.det(R), 1.0)
    assert np.allclose(oi.rotation_matrix([0, 0, np.sin(0.3), np.cos(0.3)]) @ [1, 0, 0], [np.cos(0.6), np.sin(0.6), 0])


@pytest.mark.parametrize("n, dt", [(30, 0.1), (15, 0.1), (31, 0.07)])
def test_schedule_on_the_periodic_step_call_site(n, dt):
    """The two halves of main_periodic_step.py:433-451 (second half with t0 in the middle of the plan)."""
    horizon = n * dt
    for foot in periodic_step_phases(horizon):
        half = n // 2
        for pts, t0 in ((half, 0.0), (n - half, half * dt), (n, 0.0)):
            ref = oi.foot_contact_state_interpolator(foot, DESCRIPTOR, pts, dt, t0)
            table = foot_contact_schedule(to_product(foot), pts, dt, t0)
            assert table.shape == (pts, 5) and table.dtype == np.int32
            same(evaluate_schedule(table, foot), ref)
    # the shape of the left foot's plan: stance, swing up, swing down, stance of the second phase
    kinds = foot_contact_schedule(to_product(periodic_step_phases(3.0)[0]), 30, 0.1)[:, 0].tolist()
    assert kinds == [0] * 5 + [1] * 2 + [2] * 3 + [0] * 20


def random_phases(rng, n_phases, t_start, span):
    cuts = np.sort(rng.uniform(t_start, t_start + span, 2 * n_phases))
    if n_phases > 1 and rng.random() < 0.3:  # touching phases: the swing in between has no points
        cuts[2] = cuts[1]
    phases = []
    for i in range(n_phases):
        phases.append({"position": rng.normal(size=3), "quaternion": rand_quat(rng),
                       "mid_position": rng.normal(size=3) if rng.random() < 0.5 else None, "force": rng.normal(size=3),
                       "activation_time": float(cuts[2 * i]), "deactivation_time": float(cuts[2 * i + 1])})
        phases[-1]["mid_quaternion"] = rand_quat(rng) if phases[-1]["mid_position"] is not None else None
    if rng.random() < 0.7:
        phases[0]["activation_time"] = None
    if rng.random() < 0.7:
        phases[-1]["deactivation_time"] = None
    return phases


def test_schedule_on_random_phase_timings():
    rng = np.random.default_rng(7)
    raised = agreed = 0
    for _ in range(400):
        n_ph = int(rng.integers(1, 5))
        dt = float(rng.choice([0.05, 0.1, 0.13]))
        pts = int(rng.integers(1, 40))
        t0 = float(rng.uniform(0.0, 1.5)) if rng.random() < 0.6 else 0.0
        phases = random_phases(rng, n_ph, -0.5, 4.0)
        try:
            ref = oi.foot_contact_state_interpolator(phases, DESCRIPTOR, pts, dt, t0)
        except ValueError:
            with pytest.raises(ValueError):
                foot_contact_schedule(to_product(phases), pts, dt, t0)
            raised += 1
            continue
        same(evaluate_schedule(foot_contact_schedule(to_product(phases), pts, dt, t0), phases), ref)
        agreed += 1
    assert raised > 20 and agreed > 100  # both outcomes exercised


def test_reference_errors():
    ph = to_product(periodic_step_phases(3.0)[0])
    ph[0].activation_time = 0.5
    with pytest.raises(ValueError, match="first phase activation time"):
        foot_contact_schedule(ph, 10, 0.1)
    ph = to_product(periodic_step_phases(3.0)[0])
    ph[1].activation_time = None
    with pytest.raises(ValueError, match="no activation time"):
        foot_contact_schedule(ph, 10, 0.1)
    ph = to_product(periodic_step_phases(3.0)[0])
    ph[1].deactivation_time = 2.0
    with pytest.raises(ValueError, match="before the end time"):
        foot_contact_schedule(ph, 30, 0.1)
    ph = to_product(periodic_step_phases(3.0)[0])
    ph[0].deactivation_time = 1.5
    with pytest.raises(ValueError, match="greater than the activation time of the next phase"):
        foot_contact_schedule(ph, 30, 0.1)
    with pytest.raises(ValueError, match="given together"):
        FootContactPhaseDescriptor(mid_swing_position=np.zeros(3))


def golden_interp():
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "interp_periodic_step.npz"))

    def t(v):
        return None if np.isnan(v) else float(v)

    feet = {}
    for side in ("left", "right"):
        B = g[f"{side}_position"].shape[0]
        feet[side] = [[{"position": g[f"{side}_position"][b, i], "quaternion": g[f"{side}_quaternion"][b, i],
                        "force": g[f"{side}_force"][b, i],
                        "mid_position": g[f"{side}_mid_position"][b] if i == 0 else None,
                        "mid_quaternion": g[f"{side}_mid_quaternion"][b] if i == 0 else None,
                        "activation_time": t(g[f"{side}_times"][i, 0]), "deactivation_time": t(g[f"{side}_times"][i, 1])}
                       for i in range(2)] for b in range(B)]
    return g, feet


def test_oracle_reproduces_the_interpolator_fixture():
    """tests/golden/interp_periodic_step.npz (tests/dev/make_golden.py interp) guards the fixture against oracle drift;
    a dump from the reference itself can replace it key for key."""
    g, feet = golden_interp()
    desc = (g["descriptor"].tolist(),) * 2
    N, dt = int(g["n_points"]), float(g["dt"])
    for h, (k0, k1, pts) in enumerate(((0, 1, N // 2), (1, 2, N - N // 2))):
        for b in range(g["key_0"].shape[0]):
            def unblock(v):
                pts9 = v[:72].reshape(8, 9)
                return {"p": pts9[:, :3], "f": pts9[:, 3:6], "base_position": v[72:75], "base_quaternion": v[75:79],
                        "joints": v[79:102], "com": v[102:105]}
            out = oi.humanoid_state_interpolator(unblock(g[f"key_{k0}"][b]), unblock(g[f"key_{k1}"][b]),
                                                 (feet["left"][b], feet["right"][b]), desc, pts, dt, float(g[f"t0_{h}"]))
            got = np.stack([oi.state_block(s, desc) for s in out])
            assert np.array_equal(got, g[f"states_{h}"][b])


def test_periodic_step_phases_and_pose_references(model):
    """Host side of hippopt_b200.initial_guess: the contact phases of main_periodic_step.py:365-412 per instance and
    the pose-finder references of compute_state (:192-258)."""
    from hippopt_b200.initial_guess import DESIRED_JOINTS_DEG, periodic_step_phases, pose_problem
    from hippopt_b200.pose_layout import PoseLayout, PoseSettings
    from hippopt_b200.workloads import FOOT_CORNERS

    L = np.array([0.6, 0.2])
    ph = periodic_step_phases(L, 3.0)
    ref_l, ref_r = periodic_step_phases_oracle(3.0)  # this file's transcription of the reference's call site
    for mine, ref in ((ph.left, ref_l), (ph.right, ref_r)):
        for a, b in zip(mine, ref):  # instance 0 is the reference's own step length
            assert np.array_equal(a.position[0], b["position"]) and np.array_equal(a.force, b["force"])
            assert a.activation_time == b["activation_time"] and a.deactivation_time == b["deactivation_time"]
            assert (a.mid_swing_position is None) == (b["mid_position"] is None)
            if b["mid_position"] is not None:
                assert np.array_equal(a.mid_swing_position[0], b["mid_position"])
    assert np.allclose(ph.left[1].position[1], [0.2, 0.1, 0.0]) and np.allclose(ph.right[1].position[1], [0.3, -0.1, 0.0])
    lay = PoseLayout(model, PoseSettings())
    lp, rp = ph.left[1].position, ph.right[0].position  # the middle keyframe (compute_middle_state, :292-308)
    x, p = pose_problem(lay, model, lp, rp)
    po = lay.po
    com = p[:, po.ref + po.ST_COM:po.ref + po.ST_COM + 3]
    assert np.allclose(com[:, :2], ((lp + rp) / 2)[:, :2]) and np.allclose(com[:, 2], 0.7)
    assert np.allclose(p[:, po.ref + po.ST_S:po.ref + po.ST_S + 23], np.deg2rad(DESIRED_JOINTS_DEG))
    for i in range(8):
        foot = lp if i < 4 else rp
        assert np.allclose(p[:, po.ref + 9 * i:po.ref + 9 * i + 3], foot + FOOT_CORNERS[i % 4])
    assert x.shape == (2, lay.n_x) and np.array_equal(x[:, 78:81], com)


def periodic_step_phases_oracle(horizon):
    return periodic_step_phases(horizon)


def test_slerp_and_rotation_against_scipy():
    """An independent implementation as anchor (the reference's own dependency, liecasadi, is not installable here):
    scipy's Slerp follows the shorter arc, liecasadi's formula the arc of acos(q0 . q1) -- the same one when the dot
    product is positive; quaternion -> matrix must agree always (both xyzw)."""
    from scipy.spatial.transform import Rotation, Slerp

    rng = np.random.default_rng(3)
    checked = 0
    for _ in range(40):
        q0, q1 = rand_quat(rng), rand_quat(rng)
        assert np.allclose(oi.rotation_matrix(q0), Rotation.from_quat(q0).as_matrix(), atol=1e-14)
        if np.dot(q0, q1) <= 0.05:
            continue
        ts = np.linspace(0.0, 1.0, 7)
        ref = Slerp([0.0, 1.0], Rotation.from_quat([q0, q1]))(ts).as_quat()
        got = np.stack(oi.quaternion_slerp(q0, q1, 7))
        ref = ref * np.sign((ref * got).sum(axis=1, keepdims=True))  # q and -q are the same rotation
        assert np.allclose(got, ref, atol=1e-12)
        checked += 1
    assert checked >= 10
