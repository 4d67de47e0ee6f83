"""Initial states, final states and initial guesses for a batch of periodic-step plans, on the device
(SURVEY.md 8(f) row f3).  The batched counterpart of the set-up part of
/root/reference/src/hippopt/turnkey_planners/humanoid_kinodynamic/main_periodic_step.py:

* contact phase guesses of a step of length L per instance (:365-412),
* three pose-finder solves per instance -- initial, middle and final double-support keyframes
  (compute_state / compute_initial_state / compute_middle_state / compute_final_state, :192-327) -- as ONE batch of
  3 B instances of the interior-point driver (ipsolver.py) over the pose-finder evaluator,
* `humanoid_state_interpolator` over the two halves (:433-451) written straight into the decision vectors
  (interpolators.py -> csrc/interp.cu),
* the references of get_references (:330-352) and the initial / final state parameters (:470-478).

Everything numeric runs on the GPU; the per-instance Python loop of the reference (one IPOPT solve after the other)
is gone.  How far the interior-point driver gets on the plans built here (some converge, many do not) is recorded
in DESIGN.md section 9, row (f), and profiles/r01/solver_v11.txt."""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

from .interpolators import FeetContactPhasesDescriptor, FootContactPhaseDescriptor, humanoid_state_interpolator
from .ipsolver import BatchedInteriorPoint
from .kino_layout import NJ, NPT
from .workloads import FOOT_CORNERS, GRAVITY, kino_parameters

DESIRED_JOINTS_DEG = [7, 0.12, -0.01, 12, 7, -12, 40.769, 12, 7, -12, 40.769, 5.76, 1.61, -0.31, -31.64, -20.52, -1.52,
                      5.76, 1.61, -0.31, -31.64, -20.52, -1.52]  # main_periodic_step.py:199-225


def periodic_step_phases(step_length: np.ndarray, horizon_time: float, foot_y: float = 0.1,
                         swing_height: float = 0.05, force_z: float = 100.0) -> FeetContactPhasesDescriptor:
    """main_periodic_step.py:365-412 with a step length per instance (the reference uses 0.6 for ergoCub)."""
    L = np.asarray(step_length, dtype=np.float64).reshape(-1)
    B = L.shape[0]
    ident = np.array([0.0, 0.0, 0.0, 1.0])

    def pos(x, y, z):
        return np.stack([x, np.full(B, y), np.full(B, z)], axis=1)

    def phase(p, mid, act, dea):
        return FootContactPhaseDescriptor(position=p, quaternion_xyzw=ident, mid_swing_position=mid,
                                          mid_swing_quaternion_xyzw=None if mid is None else ident,
                                          force=np.array([0.0, 0.0, force_z]), activation_time=act, deactivation_time=dea)

    T = horizon_time
    return FeetContactPhasesDescriptor(
        left=[phase(pos(0 * L, foot_y, 0.0), pos(L / 2, foot_y, swing_height), None, T / 6.0),
              phase(pos(L, foot_y, 0.0), None, T / 3.0, None)],
        right=[phase(pos(L / 2, -foot_y, 0.0), pos(L, -foot_y, swing_height), None, T * 2.0 / 3.0),
               phase(pos(1.5 * L, -foot_y, 0.0), None, T * 5.0 / 6.0, None)])


def pose_problem(lay, model, left_position: np.ndarray, right_position: np.ndarray, com_height: float = 0.7):
    """Pose-finder references of compute_state (main_periodic_step.py:192-258) for feet flat on the ground at the
    given positions (identity rotations): contact points from the foot transforms, CoM above the middle of the
    feet, identity base / frame quaternions, the desired joint configuration.  Returns (x guess, p)."""
    po = lay.po
    B = left_position.shape[0]
    p = np.zeros((B, lay.n_p))
    for i in range(NPT):
        p[:, po.desc0 + 3 * i:po.desc0 + 3 * i + 3] = FOOT_CORNERS[i % 4]
    p[:, po.mass] = model.total_mass()
    p[:, po.gravity:po.gravity + 6] = GRAVITY
    feet = (left_position, right_position)
    for i in range(NPT):
        o = po.ref + 9 * i
        p[:, o:o + 3] = feet[i // 4] + FOOT_CORNERS[i % 4]
        p[:, o + 5] = 9.80665 / 8.0
        p[:, o + 6:o + 9] = FOOT_CORNERS[i % 4]
    com = (left_position + right_position) / 2.0
    com[:, 2] = com_height
    p[:, po.ref + po.ST_PB:po.ref + po.ST_PB + 3] = com + [0.0, 0.0, 0.05]
    p[:, po.ref + po.ST_Q + 3] = 1.0
    p[:, po.ref + po.ST_S:po.ref + po.ST_S + NJ] = np.deg2rad(DESIRED_JOINTS_DEG)
    p[:, po.ref + po.ST_COM:po.ref + po.ST_COM + 3] = com
    p[:, po.ref_fq + 3] = 1.0
    p[:, po.eps], p[:, po.mu] = 1e-4, 0.3
    p[:, po.max_s:po.max_s + NJ] = 1.5
    p[:, po.min_s:po.min_s + NJ] = -1.5
    x = np.zeros((B, lay.n_x))
    for i in range(NPT):
        x[:, 6 * i:6 * i + 6] = p[:, po.ref + 9 * i:po.ref + 9 * i + 6]
    x[:, 48:81] = p[:, po.ref + po.ST_PB:po.ref + po.ST_COM + 3]
    return x, p


def state_blocks(pose: torch.Tensor) -> torch.Tensor:
    """(B, 81) pose-finder solutions [8 x (p, f), p_b, q, s, com] -> (B, 105) state blocks [8 x (p, f, descriptor), ...]."""
    B = pose.shape[0]
    corners = torch.as_tensor(np.tile(FOOT_CORNERS, (2, 1)), dtype=pose.dtype, device=pose.device)
    pts = torch.cat([pose[:, :48].reshape(B, NPT, 6), corners.expand(B, NPT, 3)], dim=2).reshape(B, 72)
    return torch.cat([pts, pose[:, 48:81]], dim=1).contiguous()


@dataclasses.dataclass
class PeriodicStepGuess:
    parameters: np.ndarray  # (B, n_p) of the kinodynamic NLP
    x0: torch.Tensor  # (B, n_x) initial guess on the device
    ok: torch.Tensor  # (B,) all three keyframe poses converged
    keyframes: torch.Tensor  # (3, B, 105) initial, middle, final state blocks
    pose_iterations: torch.Tensor  # (3 B,)


def periodic_step_guess(model, pose_evaluator, kino_evaluator, step_length: np.ndarray, tol: float = 1e-8,
                        max_iter: int = 300, force_z: float = 100.0, mass_normalised: bool = False) -> PeriodicStepGuess:
    """main_periodic_step.py:365-478 for a batch of step lengths (see the module docstring).  force_z: the planned
    contact force of the phases, 100 N per point in the reference's main.  The planner divides every force of the
    guess by the robot's mass before it reaches the NLP (`planner.py:932-980`, `_apply_mass_regularization`, called
    from `set_initial_guess`): mass_normalised=True does the same, so that force_z = 100 IS the reference's guess."""
    if mass_normalised:
        force_z = force_z / float(model.total_mass())
    lay, pl = kino_evaluator.layout, pose_evaluator.layout
    N, po = lay.N, lay.po
    L = np.asarray(step_length, dtype=np.float64).reshape(-1)
    B = L.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    dt = 0.1
    phases = periodic_step_phases(L, N * dt, force_z=force_z)
    # keyframes: (left phase, right phase) of compute_initial_state / compute_middle_state / compute_final_state
    lp = np.concatenate([phases.left[0].position, phases.left[1].position, phases.left[1].position])
    rp = np.concatenate([phases.right[0].position, phases.right[0].position, phases.right[1].position])
    xq, pq = pose_problem(pl, model, lp, rp)
    lb, ub = pose_evaluator.bounds(pq)
    out = BatchedInteriorPoint(pose_evaluator, tol=tol, max_iter=max_iter).solve(
        torch.tensor(xq, device=dev), torch.tensor(pq, device=dev), lb, ub)  # OptiFailure if no keyframe pose converges at all
    key = state_blocks(out.values).view(3, B, 105)
    ok = out.success.view(3, B).all(dim=0)

    x0 = torch.zeros((B, lay.n_x), dtype=torch.float64, device=dev)
    half = N // 2
    humanoid_state_interpolator(key[0], key[1], phases, half, dt, x_out=x0, states_out=False)
    second = humanoid_state_interpolator(key[1], key[2], phases, N - half, dt, t0=half * dt, x_out=x0, knot0=half)
    first = humanoid_state_interpolator(key[0], key[1], phases, half, dt)
    guess = torch.cat([first, second], dim=1).cpu().numpy()  # (B, N, 105): the joint references below

    p = kino_parameters(lay, model, B, np.random.default_rng(0), spread=0.0)
    k0, k2 = key[0].cpu().numpy(), key[2].cpu().numpy()
    n_state = po.ST_COM + 3
    p[:, po.init:po.init + n_state] = k0
    p[:, po.final:po.final + n_state] = k2
    for k in range(N):
        r = po.refs0 + 55 * k  # get_references, main_periodic_step.py:330-352
        p[:, r + po.R_CW:r + po.R_CW + 3] = [100.0, 100.0, 10.0]
        p[:, r + po.R_CC] = L / 2.0  # 0.3 for the reference's 0.6 m step
        p[:, r + po.R_CC + 1:r + po.R_CC + 3] = 0.0
        p[:, r + po.R_COMV:r + po.R_COMV + 3] = [0.1, 0.0, 0.0]
        p[:, r + po.R_YAW_L] = p[:, r + po.R_YAW_R] = 0.0
        p[:, r + po.R_FQ:r + po.R_FQ + 4] = [0.0, 0.0, 0.0, 1.0]
        p[:, r + po.R_BQ:r + po.R_BQ + 4] = [0.0, 0.0, 0.0, 1.0]
        p[:, r + po.R_BQV:r + po.R_BQV + 4] = 0.0
        p[:, r + po.R_JR:r + po.R_JR + NJ] = guess[:, k, 79:79 + NJ]
    return PeriodicStepGuess(parameters=p, x0=x0, ok=ok, keyframes=key, pose_iterations=out.iterations)


def single_step_problem(model, pose_evaluator, kino_evaluator, step_length: np.ndarray, tol: float = 1e-8,
                        max_iter: int = 300) -> PeriodicStepGuess:
    """BASELINE config 3, main_single_step_flat_ground.py:188-356, for a batch of step lengths: the initial state
    (feet side by side, `compute_initial_state` :188-262) and the final state (right foot one step ahead, CoM half
    a step ahead, `compute_final_state` :265-338) from 2 B pose-finder solves, and the references of
    `get_references` (:341-356: contact centroid one step ahead with weights (100, 100, 10), joint regularisation
    towards the final joints, CoM velocity 0.1 m/s forward), scaled from the main's 0.3 m step to L.

    The main hands IPOPT no initial guess (the planner's default variables); here the guess holds the initial state
    over the horizon with zero velocities -- the same "standing" start in the layout of the decision vector."""
    lay, pl = kino_evaluator.layout, pose_evaluator.layout
    N, po = lay.N, lay.po
    L = np.asarray(step_length, dtype=np.float64).reshape(-1)
    B = L.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    zero = np.zeros(B)

    def pos(x, y):
        return np.stack([x, np.full(B, y), zero], axis=1)

    lp = np.concatenate([pos(zero, 0.1), pos(zero, 0.1)])
    rp = np.concatenate([pos(zero, -0.1), pos(L, -0.1)])
    xq, pq = pose_problem(pl, model, lp, rp)
    lb, ub = pose_evaluator.bounds(pq)
    out = BatchedInteriorPoint(pose_evaluator, tol=tol, max_iter=max_iter).solve(
        torch.tensor(xq, device=dev), torch.tensor(pq, device=dev), lb, ub)
    key = state_blocks(out.values).view(2, B, 105)
    ok = out.success.view(2, B).all(dim=0)
    k0, k2 = key[0].cpu().numpy(), key[1].cpu().numpy()
    p = kino_parameters(lay, model, B, np.random.default_rng(0), spread=0.0)
    n_state = po.ST_COM + 3
    p[:, po.init:po.init + n_state] = k0
    p[:, po.final:po.final + n_state] = k2
    for k in range(N):
        r = po.refs0 + 55 * k
        p[:, r + po.R_CW:r + po.R_CW + 3] = [100.0, 100.0, 10.0]
        p[:, r + po.R_CC] = L
        p[:, r + po.R_CC + 1:r + po.R_CC + 3] = 0.0
        p[:, r + po.R_COMV:r + po.R_COMV + 3] = [0.1, 0.0, 0.0]
        p[:, r + po.R_YAW_L] = p[:, r + po.R_YAW_R] = 0.0
        p[:, r + po.R_FQ:r + po.R_FQ + 4] = [0.0, 0.0, 0.0, 1.0]
        p[:, r + po.R_BQ:r + po.R_BQ + 4] = [0.0, 0.0, 0.0, 1.0]
        p[:, r + po.R_BQV:r + po.R_BQV + 4] = 0.0
        p[:, r + po.R_JR:r + po.R_JR + NJ] = k2[:, po.ST_S:po.ST_S + NJ]
    # guess: the initial state at every knot (points p, f; base; joints; CoM), velocities and momentum zero
    from .kino_layout import COM, F, NZ, P, PB, Q, S

    x0 = np.zeros((B, lay.n_x))
    for k in range(N):
        z = x0[:, NZ * k:NZ * (k + 1)]
        for i in range(NPT):
            z[:, 15 * i + P:15 * i + P + 3] = k0[:, 9 * i:9 * i + 3]
            z[:, 15 * i + F:15 * i + F + 3] = k0[:, 9 * i + 3:9 * i + 6]
        z[:, PB:PB + 3] = k0[:, po.ST_PB:po.ST_PB + 3]
        z[:, Q:Q + 4] = k0[:, po.ST_Q:po.ST_Q + 4]
        z[:, S:S + NJ] = k0[:, po.ST_S:po.ST_S + NJ]
        z[:, COM:COM + 3] = k0[:, po.ST_COM:po.ST_COM + 3]
    return PeriodicStepGuess(parameters=p, x0=torch.tensor(x0, device=dev), ok=ok,
                             keyframes=torch.stack([key[0], key[1], key[1]]), pose_iterations=out.iterations)
