"""Batched state interpolators: the reference's initial-guess / reference-trajectory
generator `humanoid_state_interpolator` (/root/reference/src/hippopt/robot_planning/utilities/interpolators.py:396-448)
for many instances at once, evaluated by one CUDA kernel (csrc/interp.cu through ``hb_interpolate_humanoid_states``).

Split of the work:
* the phase bookkeeping of ``foot_contact_state_interpolator`` (interpolators.py:106-169 validation, :231-309
  sequencing) only looks at activation / deactivation times, the number of points, ``dt`` and ``t0``.  It is done
  here, once for the whole batch, by :func:`foot_contact_schedule`, which returns for every point what the
  reference would have appended there (a stance of phase a / sample j of n of the first or second half of the
  swing a -> b) and raises the reference's ``ValueError``s;
* everything numeric (linear interpolation, slerp, contact points from the foot transform) runs on the device,
  per instance: the transforms, mid-swing transforms and forces of the phases may differ between instances.

States are "state blocks" of ``82 + n_joints`` doubles, the layout the kinodynamic NLP uses for its initial /
final state parameters (kino_layout.ParamOffsets.st_pt, ST_*): 8 x (p, f, position_in_foot_frame), base position,
base quaternion (xyzw), joint positions, CoM.  There is no CPU path: without the CUDA library the call fails.
"""
from __future__ import annotations

import ctypes
import dataclasses
import math

import numpy as np
import torch

from . import _capi

STANCE, SWING_UP, SWING_DOWN = 0, 1, 2
PHASE_RECORD = 17


@dataclasses.dataclass
class FootContactPhaseDescriptor:
    """`FootContactPhaseDescriptor` (robot_planning/variables/contacts.py:143-167) with the SE3 members spelled
    out: arrays of shape (3,) / (4,), or (B, 3) / (B, 4) when they differ between instances."""

    position: np.ndarray = None
    quaternion_xyzw: np.ndarray = None
    mid_swing_position: np.ndarray | None = None
    mid_swing_quaternion_xyzw: np.ndarray | None = None
    force: np.ndarray = None
    activation_time: float | None = None
    deactivation_time: float | None = None

    def __post_init__(self) -> None:  # contacts.py:150-167 defaults: identity transform, zero force
        self.position = np.zeros(3) if self.position is None else np.asarray(self.position, dtype=np.float64)
        self.quaternion_xyzw = (np.array([0.0, 0.0, 0.0, 1.0]) if self.quaternion_xyzw is None
                                else np.asarray(self.quaternion_xyzw, dtype=np.float64))
        self.force = np.zeros(3) if self.force is None else np.asarray(self.force, dtype=np.float64)
        if (self.mid_swing_position is None) != (self.mid_swing_quaternion_xyzw is None):
            raise ValueError("mid_swing_position and mid_swing_quaternion_xyzw must be given together.")


@dataclasses.dataclass
class FeetContactPhasesDescriptor:  # contacts.py:170-177
    left: list[FootContactPhaseDescriptor] = dataclasses.field(default_factory=list)
    right: list[FootContactPhaseDescriptor] = dataclasses.field(default_factory=list)


def foot_contact_schedule(phases: list[FootContactPhaseDescriptor], number_of_points: int, dt: float,
                          t0: float = 0.0) -> np.ndarray:
    """What `foot_contact_state_interpolator` (interpolators.py:106-309) appends at each of the
    ``number_of_points`` points, as an int32 table (number_of_points, 5): kind, a, b, j, n.

    kind STANCE: transform and force of phase a.  SWING_UP / SWING_DOWN: zero force, transform =
    transform_interpolator(a.transform -> a.mid_swing_transform, n)[j]  /  (a.mid_swing_transform -> b.transform, n)[j].
    """
    assert len(phases) > 0
    act = [ph.activation_time for ph in phases]
    dea = [ph.deactivation_time for ph in phases]
    return np.asarray(_schedule(act, dea, number_of_points, dt, t0), dtype=np.int32).reshape(number_of_points, 5)


def _schedule(act: list, dea: list, n_pts: int, dt: float, t0: float) -> list[tuple]:
    assert n_pts > 0
    assert dt > 0.0
    end_time = t0 + dt * n_pts
    act, dea = list(act), list(dea)
    if act[0] is None:  # :118-124
        act[0] = min(dea[0] if dea[0] is not None else t0, t0) - dt
    if act[0] > t0:
        raise ValueError(f"The first phase activation time ({act[0]}) is after the start time ({t0}).")
    for i, a in enumerate(act):
        if a is None:
            raise ValueError(f"Phase {i} has no activation time, but is not the first phase.")
    last = len(act) - 1
    if dea[last] is None:  # :140-144
        dea[last] = max(end_time, act[last]) + dt
    if dea[last] < end_time:
        raise ValueError(f"The Last phase deactivation time ({dea[last]}) is before the end time ({end_time}, "
                         f"computed from the inputs).")
    for i in range(last + 1):
        if dea[i] is None:
            raise ValueError(f"Phase {i} has no deactivation time, but is not the last phase.")
        if act[i] > dea[i]:
            raise ValueError(f"Phase {i} has an activation time ({act[i]}) greater than its deactivation time ({dea[i]}).")
        if i < last and dea[i] > act[i + 1]:
            raise ValueError(f"Phase {i} has a deactivation time ({dea[i]}) greater than the activation time of the "
                             f"next phase ({act[i + 1]}).")

    # sequencing (:231-309)
    out: list[tuple] = []
    if last == 0 or dea[0] >= end_time:  # :231-233
        return [(STANCE, 0, 0, 0, 1)] * n_pts
    i = 0
    while act[i] < t0 and dea[i] <= t0:  # :235-241: skip the phases that ended before t0
        i += 1
    if act[i] > t0:
        # t0 falls into a swing (:243-254): the reference interpolates from one step before the lift-off (with
        # the completed times, validated again) and drops the points before t0, so the swing keeps its shape
        new_t0 = dea[i - 1] - dt
        advance = int(math.ceil((t0 - new_t0) / dt))
        return _schedule(act, dea, n_pts + advance, dt, new_t0)[advance:]
    remaining = n_pts
    while i < last:
        stance = min(int(math.ceil((dea[i] - max(act[i], t0)) / dt)), remaining)  # :261-264
        out += [(STANCE, i, i, 0, 1)] * stance
        remaining -= stance
        if remaining == 0:
            return out
        swing = min(int(math.ceil((act[i + 1] - dea[i]) / dt)), remaining)  # :272-276
        if swing == 0:
            # the reference `continue`s without advancing (:278-279) and appends the same stance again; with a
            # zero-length stance that loop never ends, which is reported here instead
            if stance <= 0:
                raise ValueError(f"Phase {i} and the swing after it have no points: the interpolation cannot advance.")
            continue
        full = int(math.ceil((act[i + 1] - dea[i]) / dt))  # append_swing_phase :188-190
        up = min(round(full / 2), swing)  # :204 (Python's round: half to even)
        out += [(SWING_UP, i, i + 1, j, up) for j in range(up)]
        out += [(SWING_DOWN, i, i + 1, j, swing - up) for j in range(swing - up)]
        remaining -= swing
        if remaining == 0:
            return out
        i += 1
    return out + [(STANCE, last, last, 0, 1)] * remaining  # :307-309


def _phase_records(phases: list[FootContactPhaseDescriptor], device) -> tuple[torch.Tensor, int]:
    """(B or 1, n_phases, 17) device tensor of transform, mid-swing transform and force per phase; a missing
    mid-swing transform is the reference's default (:192-202): translation half way to the next phase's, rotation
    of the next phase (the last phase has no swing after it: its mid-swing slot repeats its own transform)."""
    def arr(v, n):
        v = np.asarray(v, dtype=np.float64)
        if v.shape[-1] != n or v.ndim > 2:
            raise ValueError(f"expected an array of shape ({n},) or (B, {n}), got {v.shape}")
        return v.reshape(-1, n)

    cols = []
    for i, ph in enumerate(phases):
        pos, quat, force = arr(ph.position, 3), arr(ph.quaternion_xyzw, 4), arr(ph.force, 3)
        if ph.mid_swing_position is not None:
            mpos, mquat = arr(ph.mid_swing_position, 3), arr(ph.mid_swing_quaternion_xyzw, 4)
        elif i + 1 < len(phases):
            nxt = phases[i + 1]
            mpos, mquat = (pos + arr(nxt.position, 3)) / 2, arr(nxt.quaternion_xyzw, 4)
        else:
            mpos, mquat = pos, quat
        cols.append([pos, quat, mpos, mquat, force])
    B = max(a.shape[0] for c in cols for a in c)
    rec = np.empty((B, len(phases), PHASE_RECORD))
    for i, c in enumerate(cols):
        o = 0
        for a in c:
            if a.shape[0] not in (1, B):
                raise ValueError(f"phase arrays have batch sizes {a.shape[0]} and {B}")
            rec[:, i, o:o + a.shape[1]] = a
            o += a.shape[1]
    return torch.as_tensor(rec, device=device), B


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def humanoid_state_interpolator(initial_state: torch.Tensor, final_state: torch.Tensor,
                                contact_phases: FeetContactPhasesDescriptor, number_of_points: int, dt: float,
                                t0: float = 0.0, x_out: torch.Tensor | None = None, knot0: int = 0,
                                states_out: bool = True) -> torch.Tensor | None:
    """`humanoid_state_interpolator` (interpolators.py:396-448) for a batch.

    initial_state, final_state: (B, 82 + n_joints) float64 CUDA state blocks (the contact point descriptors --
    ``contact_descriptor`` in the reference -- are the position_in_foot_frame entries of ``initial_state``).
    Returns the (B, number_of_points, 82 + n_joints) interpolated states; with ``x_out`` (B, n_x) the same values
    are also written into the decision vector of the kinodynamic NLP at knots knot0 ..."""
    if (initial_state.dim() != 2 or initial_state.shape != final_state.shape or initial_state.dtype != torch.float64
            or not initial_state.is_cuda or not final_state.is_cuda):
        raise ValueError(f"Initial value has shape {tuple(initial_state.shape)}, but final value has shape "
                         f"{tuple(final_state.shape)} (float64 CUDA state blocks of equal shape are required).")
    B, ns = initial_state.shape
    nj = ns - 82
    if nj < 0:
        raise ValueError(f"a state block has at least 82 entries, got {ns}")
    dev = initial_state.device
    sched = np.stack([foot_contact_schedule(contact_phases.left, number_of_points, dt, t0),
                      foot_contact_schedule(contact_phases.right, number_of_points, dt, t0)])
    rec_l, bl = _phase_records(contact_phases.left, dev)
    rec_r, br = _phase_records(contact_phases.right, dev)
    for nb in (bl, br):
        if nb not in (1, B):
            raise ValueError(f"the phases describe {nb} instances, the states {B}")
    d_sched = torch.as_tensor(sched, device=dev)
    states = torch.empty((B, number_of_points, ns), dtype=torch.float64, device=dev) if states_out else None
    if x_out is not None and (not x_out.is_cuda or x_out.dtype != torch.float64 or x_out.dim() != 2
                              or x_out.shape[0] != B or x_out.stride(1) != 1):
        raise ValueError("x_out must be a (B, n_x) float64 CUDA tensor with unit inner stride")
    ini, fin = initial_state.contiguous(), final_state.contiguous()
    st = torch.cuda.current_stream(dev).cuda_stream
    _capi.check(_capi.lib().hb_interpolate_humanoid_states(
        B, number_of_points, nj, _ptr(ini), _ptr(fin), _ptr(d_sched),
        _ptr(rec_l), rec_l.shape[1], 0 if bl == 1 else rec_l.stride(0),
        _ptr(rec_r), rec_r.shape[1], 0 if br == 1 else rec_r.stride(0),
        _ptr(states), _ptr(x_out), x_out.stride(0) if x_out is not None else 0, knot0, ctypes.c_void_p(st)),
        "hb_interpolate_humanoid_states")
    return states
