"""Row f3 on the GPU: `hb_interpolate_humanoid_states` (csrc/interp.cu, one kernel for the whole batch, driven by
the host-side schedule) against the oracle's instance-by-instance, state-by-state restatement of
/root/reference/src/hippopt/robot_planning/utilities/interpolators.py:396-448.
Linear parts are bit-exact (same two products and one sum); slerp goes through acos / sin of two different maths
libraries: 1e-13 absolute on unit quaternions and on the contact points they rotate."""
import numpy as np
import pytest
import torch

from oracle import interpolators as oi
from test_interpolators_cpu import DESCRIPTOR, periodic_step_phases, rand_quat, random_phases, to_product

pytestmark = pytest.mark.gpu

NJ = 23
SLERP_TOL = 1e-13


def random_state(rng, q=None):
    return {"p": rng.normal(size=(8, 3)), "f": rng.normal(size=(8, 3)), "base_position": rng.normal(size=3),
            "base_quaternion": rand_quat(rng) if q is None else q, "joints": rng.uniform(-1, 1, NJ), "com": rng.normal(size=3)}


def batch_phases(per_instance):
    """list over instances of oracle phase lists -> product descriptors with (B, .) arrays."""
    from hippopt_b200.interpolators import FootContactPhaseDescriptor

    out = []
    for i, first in enumerate(per_instance[0]):
        col = [inst[i] for inst in per_instance]
        has_mid = first["mid_position"] is not None
        out.append(FootContactPhaseDescriptor(
            position=np.stack([c["position"] for c in col]), quaternion_xyzw=np.stack([c["quaternion"] for c in col]),
            mid_swing_position=np.stack([c["mid_position"] for c in col]) if has_mid else None,
            mid_swing_quaternion_xyzw=np.stack([c["mid_quaternion"] for c in col]) if has_mid else None,
            force=np.stack([c["force"] for c in col]), activation_time=first["activation_time"],
            deactivation_time=first["deactivation_time"]))
    return out


def check(states, ref_blocks):
    """states (B, n, ns) from the kernel, ref_blocks the oracle's; lerp entries exact, slerp-dependent ones close."""
    got = states.cpu().numpy()
    assert got.shape == ref_blocks.shape
    exact = np.r_[72:75, 79:79 + NJ + 3]  # base position, joints, com
    exact = np.concatenate([exact] + [np.r_[9 * i + 3:9 * i + 9] for i in range(8)])  # forces, descriptors
    assert np.array_equal(got[..., exact], ref_blocks[..., exact])
    assert np.abs(got - ref_blocks).max() < SLERP_TOL


def test_periodic_step_guess_matches_the_oracle(built_library):
    """The reference's call site (main_periodic_step.py:433-451): two halves, the second starting mid-plan."""
    from hippopt_b200.interpolators import FeetContactPhasesDescriptor, humanoid_state_interpolator

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    B, n, dt = 5, 30, 0.1
    left, right = periodic_step_phases(n * dt)
    desc = (DESCRIPTOR, DESCRIPTOR)
    phases = FeetContactPhasesDescriptor(left=to_product(left), right=to_product(right))
    half = n // 2
    for pts, t0 in ((half, 0.0), (n - half, half * dt)):
        ini = [random_state(rng) for _ in range(B)]
        fin = [random_state(rng) for _ in range(B)]
        ref = np.stack([np.stack([oi.state_block(s, desc) for s in
                                  oi.humanoid_state_interpolator(ini[b], fin[b], (left, right), desc, pts, dt, t0)])
                        for b in range(B)])
        I = torch.tensor(np.stack([oi.state_block(s, desc) for s in ini]), device=dev)
        F = torch.tensor(np.stack([oi.state_block(s, desc) for s in fin]), device=dev)
        check(humanoid_state_interpolator(I, F, phases, pts, dt, t0), ref)


def test_per_instance_phases_and_guess_scatter(built_library):
    """Transforms, mid-swing transforms and forces that differ per instance; the same values land in the
    decision vector of the kinodynamic NLP (`x_out`), everything else in it untouched."""
    from hippopt_b200.interpolators import FeetContactPhasesDescriptor, humanoid_state_interpolator
    from hippopt_b200.kino_layout import COM, NZ, PB, Q, S, F as ZF, P as ZP

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    B, desc = 7, (DESCRIPTOR, [[0.1, 0.02, 0.0], [0.1, -0.02, 0.01], [-0.05, -0.02, 0.0], [-0.05, 0.02, 0.0]])
    done = 0
    while done < 6:
        n_ph, pts = int(rng.integers(1, 4)), int(rng.integers(1, 35))
        dt, t0 = float(rng.choice([0.05, 0.1])), float(rng.uniform(0, 1.0))
        feet = []
        for _ in range(2):
            state = rng.bit_generator.state
            per = []
            for b in range(B):  # same timings (and the same mid-swing presence) for every instance
                rng.bit_generator.state = state
                tmpl = random_phases(rng, n_ph, -0.5, 4.0)
                r2 = np.random.default_rng(1000 * done + b)
                for ph in tmpl:
                    ph["position"], ph["quaternion"], ph["force"] = r2.normal(size=3), rand_quat(r2), r2.normal(size=3)
                    if ph["mid_position"] is not None:
                        ph["mid_position"], ph["mid_quaternion"] = r2.normal(size=3), rand_quat(r2)
                per.append(tmpl)
            feet.append(per)
        try:
            for f in feet:
                oi.foot_contact_state_interpolator(f[0], DESCRIPTOR, pts, dt, t0)
        except (ValueError, AssertionError):
            continue
        ini = [random_state(rng) for _ in range(B)]
        fin = [random_state(rng) for _ in range(B)]
        fin[0]["base_quaternion"] = ini[0]["base_quaternion"].copy()  # below the 1e-6 threshold: initial quaternion
        ref = np.stack([np.stack([oi.state_block(s, desc) for s in
                                  oi.humanoid_state_interpolator(ini[b], fin[b], (feet[0][b], feet[1][b]), desc, pts, dt, t0)])
                        for b in range(B)])
        I = torch.tensor(np.stack([oi.state_block(s, desc) for s in ini]), device=dev)
        Fn = torch.tensor(np.stack([oi.state_block(s, desc) for s in fin]), device=dev)
        phases = FeetContactPhasesDescriptor(left=batch_phases(feet[0]), right=batch_phases(feet[1]))
        knot0, n_x = 2, NZ * (pts + 3) + 6
        x = torch.full((B, n_x), -7.0, dtype=torch.float64, device=dev)
        states = humanoid_state_interpolator(I, Fn, phases, pts, dt, t0, x_out=x, knot0=knot0)
        check(states, ref)
        assert np.array_equal(states[0, :, 75:79].cpu().numpy(), np.tile(ini[0]["base_quaternion"], (pts, 1)))
        xs, st = x.cpu().numpy(), states.cpu().numpy()
        touched = np.zeros(n_x, dtype=bool)
        for k in range(pts):
            z = NZ * (knot0 + k)
            for i in range(8):
                assert np.array_equal(xs[:, z + 15 * i + ZP:z + 15 * i + ZP + 3], st[:, k, 9 * i:9 * i + 3])
                assert np.array_equal(xs[:, z + 15 * i + ZF:z + 15 * i + ZF + 3], st[:, k, 9 * i + 3:9 * i + 6])
                touched[z + 15 * i + ZP:z + 15 * i + ZP + 3] = touched[z + 15 * i + ZF:z + 15 * i + ZF + 3] = True
            for zo, so, ln in ((PB, 72, 3), (Q, 75, 4), (S, 79, NJ), (COM, 79 + NJ, 3)):
                assert np.array_equal(xs[:, z + zo:z + zo + ln], st[:, k, so:so + ln])
                touched[z + zo:z + zo + ln] = True
        assert (xs[:, ~touched] == -7.0).all()
        done += 1


def test_interpolator_argument_errors(built_library):
    from hippopt_b200.interpolators import FeetContactPhasesDescriptor, humanoid_state_interpolator

    dev = torch.device("cuda:0")
    left, right = periodic_step_phases(3.0)
    phases = FeetContactPhasesDescriptor(left=to_product(left), right=to_product(right))
    I = torch.zeros((3, 82 + NJ), dtype=torch.float64, device=dev)
    with pytest.raises(ValueError, match="Initial value has shape"):  # interpolators.py:36-44
        humanoid_state_interpolator(I, I[:, :-1], phases, 10, 0.1)
    with pytest.raises(ValueError, match="x_out"):
        humanoid_state_interpolator(I, I, phases, 10, 0.1, x_out=torch.zeros((2, 4000), dtype=torch.float64, device=dev))
    from hippopt_b200._capi import EvaluationError
    with pytest.raises(EvaluationError, match="x_stride must cover"):
        humanoid_state_interpolator(I, I, phases, 10, 0.1, x_out=torch.zeros((3, 500), dtype=torch.float64, device=dev))


def test_periodic_step_setup_pipeline(model, built_library):
    """main_periodic_step.py:365-478 for a batch: keyframes from the batched pose finder, guess from the device
    interpolator, written into the decision vector and the parameters of the kinodynamic NLP."""
    from hippopt_b200.evaluator import KinoEvaluator, PoseEvaluator
    from hippopt_b200.initial_guess import periodic_step_guess
    from hippopt_b200.kino_layout import COM, NZ, PB, Q, S, KinoSettings, F as ZF, P as ZP
    from hippopt_b200.workloads import FOOT_CORNERS

    B, N = 24, 12
    L = np.random.default_rng(5).uniform(0.1, 0.3, B)
    ev = KinoEvaluator(model, KinoSettings(horizon=N, final_state_constraint=True, periodicity_constraint=True))
    gs = periodic_step_guess(model, PoseEvaluator(model), ev, L)
    assert int(gs.ok.sum()) >= int(0.9 * B)
    x, key, po = gs.x0.cpu().numpy(), gs.keyframes.cpu().numpy(), ev.layout.po
    assert np.isfinite(x).all() and x.shape == (B, ev.layout.n_x)
    z0, zl = x[:, :NZ], x[:, NZ * (N - 1):NZ * N]
    # the base, joints and CoM go from the initial keyframe to the final one through the middle one
    for zo, so, ln in ((PB, 72, 3), (S, 79, NJ), (COM, 79 + NJ, 3)):
        assert np.array_equal(z0[:, zo:zo + ln], key[0][:, so:so + ln])
        assert np.array_equal(zl[:, zo:zo + ln], key[2][:, so:so + ln])
        zh = x[:, NZ * (N // 2):NZ * (N // 2 + 1)]
        assert np.array_equal(zh[:, zo:zo + ln], key[1][:, so:so + ln])
    qn = np.stack([np.linalg.norm(x[:, NZ * k + Q:NZ * k + Q + 4], axis=1) for k in range(N)])
    assert np.abs(qn - 1.0).max() < 1e-12
    # feet where the phases put them: left foot starts at x = 0 and ends at x = L, right foot from L/2 to 3L/2;
    # planned force on the ground, none in the air
    for i in range(8):
        x_first, x_last = (0.0 * L, L) if i < 4 else (L / 2, 1.5 * L)
        y = 0.1 if i < 4 else -0.1
        assert np.allclose(z0[:, 15 * i + ZP:15 * i + ZP + 3], FOOT_CORNERS[i % 4] + np.stack([x_first, y + 0 * L, 0 * L], 1), atol=1e-15)
        if i < 4:  # (the right foot is still in the air at the last knot of this short horizon)
            assert np.allclose(zl[:, 15 * i + ZP:15 * i + ZP + 3], FOOT_CORNERS[i % 4] + np.stack([x_last, y + 0 * L, 0 * L], 1), atol=1e-15)
    pz = np.stack([x[:, NZ * k + ZP + 2] for k in range(N)])  # height of the first left point over the knots
    fz = np.stack([x[:, NZ * k + ZF + 2] for k in range(N)])
    assert (pz.max(axis=0) > 0.03).all() and (fz[pz > 1e-12] == 0.0).all() and set(np.unique(fz)) == {0.0, 100.0}
    # parameters: initial / final state = first / last keyframe; joint regularisation = the guess
    p = gs.parameters
    assert np.array_equal(p[:, po.init:po.init + 105], key[0]) and np.array_equal(p[:, po.final:po.final + 105], key[2])
    for k in (0, N // 2, N - 1):
        r = po.refs0 + 55 * k
        assert np.array_equal(p[:, r + po.R_JR:r + po.R_JR + NJ], x[:, NZ * k + S:NZ * k + S + NJ])
    # and the evaluator accepts the pair
    g = ev.eval(4, gs.x0, torch.tensor(p, device=gs.x0.device))["g"]
    assert torch.isfinite(g).all()


def test_kernel_matches_the_interpolator_fixture(built_library):
    """Against the committed fixture (no oracle at run time): per-instance rotated foot transforms, both halves."""
    from hippopt_b200.interpolators import FeetContactPhasesDescriptor, FootContactPhaseDescriptor, humanoid_state_interpolator
    from test_interpolators_cpu import golden_interp

    g, _ = golden_interp()
    dev = torch.device("cuda:0")

    def phases(side):
        def t(v):
            return None if np.isnan(v) else float(v)
        return [FootContactPhaseDescriptor(
            position=g[f"{side}_position"][:, i], quaternion_xyzw=g[f"{side}_quaternion"][:, i], force=g[f"{side}_force"][:, i],
            mid_swing_position=g[f"{side}_mid_position"] if i == 0 else None,
            mid_swing_quaternion_xyzw=g[f"{side}_mid_quaternion"] if i == 0 else None,
            activation_time=t(g[f"{side}_times"][i, 0]), deactivation_time=t(g[f"{side}_times"][i, 1])) for i in range(2)]

    ph = FeetContactPhasesDescriptor(left=phases("left"), right=phases("right"))
    N, dt = int(g["n_points"]), float(g["dt"])
    for h, (k0, k1, pts) in enumerate(((0, 1, N // 2), (1, 2, N - N // 2))):
        out = humanoid_state_interpolator(torch.tensor(g[f"key_{k0}"], device=dev), torch.tensor(g[f"key_{k1}"], device=dev),
                                          ph, pts, dt, float(g[f"t0_{h}"]))
        check(out, g[f"states_{h}"])
