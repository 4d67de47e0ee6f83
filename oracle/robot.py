"""oracle/robot.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement, on the SX engine of oracle/sx.py, of the floating-base quantities the
reference obtains from adam-robotics and liecasadi [ext: neither is vendored under
/root/reference nor installed here; both un-pinned in `/root/reference/setup.cfg:52-78`]:

  * ``liecasadi.SO3.from_quat(xyzw).as_matrix()`` / ``SE3.from_position_quaternion``
    (`expressions/kinematics.py:49-51,165-167,251-253`)      -> :func:`quat_to_rot`
  * ``kindyn.forward_kinematics_fun(frame)(H_b, s)`` (`kinematics.py:249`) -> :func:`body_transforms`
    / :func:`frame_transform`  (H_frame = H_b * prod_j H_j(s_j), revolute
    H_j = [R_rpy * Rodrigues(axis, s), xyz])
  * ``kindyn.CoM_position_fun()`` (`kinematics.py:163`)       -> :func:`com_position`
  * ``kindyn.centroidal_momentum_matrix_fun()`` (`kinematics.py:47`) -> :func:`centroidal_momentum_matrix`
    restating adam's published route: composite-rigid-body algorithm in body-fixed
    representation (Featherstone), conversion of the floating-base rows to the MIXED velocity
    representation (linear velocity of the base origin and angular velocity both in inertial
    coordinates), then the change of pole to the centre of mass (Orin & Goswami).

The model is passed in as plain arrays (``hippopt_b200.robot_model.RobotModel`` fields); no
product *code* is used here.  Spatial vectors are ordered [linear; angular].
parity status: UNPINNED (no adam install to compare with); cross-checked in tests/ against
first-principles sums over bodies written independently with numpy.
"""
from __future__ import annotations

import numpy as np

from . import sx
from .sx import SX


def quat_to_rot(q):
    """R = I + 2 w [v]x + 2 [v]x^2 for an xyzw quaternion (not normalised here)."""
    v = q[:3]
    w = q[3]
    S = sx.skew(v)
    S2 = sx.matmul(S, S)
    R = sx.eye(3)
    for i in range(3):
        for j in range(3):
            R[i, j] = R[i, j] + 2.0 * w * S[i, j] + 2.0 * S2[i, j]
    return R


def quat_normalize(q):
    n = sx.norm2(q)
    return sx.vec(*[qi / n for qi in q])


def rodrigues(axis, s):
    """Rotation about a constant unit axis: I + sin(s) [a]x + (1 - cos(s)) [a]x^2."""
    A = sx.lift(np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]]))
    A2 = sx.matmul(A, A)
    sn, cs = sx.sin(s), sx.cos(s)
    R = sx.eye(3)
    for i in range(3):
        for j in range(3):
            R[i, j] = R[i, j] + sn * A[i, j] + (1.0 - cs) * A2[i, j]
    return R


def homogeneous(R, t):
    H = sx.eye(4)
    for i in range(3):
        for j in range(3):
            H[i, j] = R[i, j]
        H[i, 3] = sx._wrap(t[i])
    return H


def joint_transform(model, b: int, s):
    """parent-frame <- child-frame transform of the joint moving body b (b >= 1)."""
    R = sx.matmul(sx.lift(model.joint_rot[b]), rodrigues(model.joint_axis[b], s))
    return homogeneous(R, sx.lift(model.joint_xyz[b]))


def body_transforms(model, H_b, s):
    """World transform of every reduced body: H_0 = H_b, H_b' = H_parent * H_joint(s)."""
    H = [None] * model.n_bodies
    H[0] = H_b
    for b in range(1, model.n_bodies):
        H[b] = sx.matmul(H[int(model.parent[b])], joint_transform(model, b, s[b - 1]))
    return H


def frame_transform(model, H_b, s, frame: str):
    """adam ``forward_kinematics_fun(frame)``: only the chain root -> frame is multiplied."""
    body, Rf, tf = model.frames[frame]
    H = H_b
    for b in model.chain_to_root(body):
        H = sx.matmul(H, joint_transform(model, b, s[b - 1]))
    return sx.matmul(H, homogeneous(sx.lift(Rf), sx.lift(tf)))


def com_position(model, H_b, s):
    """x_com = sum_l m_l (H_l [c_l; 1])_{0:3} / M."""
    H = body_transforms(model, H_b, s)
    M = model.total_mass()
    acc = sx.zeros(3)
    for b in range(model.n_bodies):
        c = sx.vec(*sx.lift(model.com[b]), 1.0)
        pc = sx.matmul(H[b], c)
        for i in range(3):
            acc[i] = acc[i] + model.mass[b] * pc[i]
    return sx.vec(*[acc[i] / M for i in range(3)])


def _spatial_inertia(m, c, I_c):
    """6x6 spatial inertia about the body-frame origin, [lin; ang] ordering."""
    S = np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])
    out = np.zeros((6, 6))
    out[:3, :3] = m * np.eye(3)
    out[:3, 3:] = -m * S
    out[3:, :3] = m * S
    out[3:, 3:] = I_c - m * S @ S
    return out


def _motion_transform(R, r):
    """child <- parent velocity transform for H = (R, r) mapping child coords to parent coords."""
    Rt = np.empty((3, 3), dtype=object)
    for i in range(3):
        for j in range(3):
            Rt[i, j] = R[j, i]
    X = sx.zeros(6, 6)
    RtS = sx.matmul(Rt, sx.skew(r))
    for i in range(3):
        for j in range(3):
            X[i, j] = Rt[i, j]
            X[i, 3 + j] = -RtS[i, j]
            X[3 + i, 3 + j] = Rt[i, j]
    return X


def _transpose(A):
    n, m = A.shape
    out = np.empty((m, n), dtype=object)
    for i in range(n):
        for j in range(m):
            out[j, i] = A[i, j]
    return out


def centroidal_momentum_matrix(model, H_b, s):
    """A_G (6 x (6+n)), mixed representation, momentum expressed in G[A] (pole at the CoM,
    inertial orientation): h_G = A_G [pb_dot; omega_world; s_dot]."""
    nb = model.n_bodies
    n = model.n_joints
    X = [None] * nb  # body <- parent
    Ic = [sx.lift(_spatial_inertia(model.mass[b], model.com[b], model.inertia[b])) for b in range(nb)]
    for b in range(1, nb):
        Hj = joint_transform(model, b, s[b - 1])
        X[b] = _motion_transform(Hj[:3, :3], Hj[:3, 3])
    for b in range(nb - 1, 0, -1):
        p = int(model.parent[b])
        contrib = sx.matmul(_transpose(X[b]), sx.matmul(Ic[b], X[b]))
        for i in range(6):
            for j in range(6):
                Ic[p][i, j] = Ic[p][i, j] + contrib[i, j]
    A_body = sx.zeros(6, 6 + n)
    for i in range(6):
        for j in range(6):
            A_body[i, j] = Ic[0][i, j]
    for b in range(1, nb):
        phi = sx.vec(0.0, 0.0, 0.0, *sx.lift(model.joint_axis[b]))
        F = sx.matmul(Ic[b], phi)
        j = b
        while j > 0:
            F = sx.matmul(_transpose(X[j]), F)
            j = int(model.parent[j])
        for i in range(6):
            A_body[i, 6 + b - 1] = F[i]
    # body-fixed -> mixed: V_body = blkdiag(R^T, R^T) V_mixed ; h_mixed = blkdiag(R, R) h_body
    R = H_b[:3, :3]
    RR = sx.zeros(6, 6)
    for i in range(3):
        for j in range(3):
            RR[i, j] = R[i, j]
            RR[3 + i, 3 + j] = R[i, j]
    A_base = sx.matmul(A_body[:, :6], _transpose(RR))
    A_mixed = sx.zeros(6, 6 + n)
    tmp = sx.matmul(RR, A_base)
    for i in range(6):
        for j in range(6):
            A_mixed[i, j] = tmp[i, j]
    tmp = sx.matmul(RR, A_body[:, 6:])
    for i in range(6):
        for j in range(n):
            A_mixed[i, 6 + j] = tmp[i, j]
    # pole: base origin -> centre of mass.  m [c]x sits in the lower-left block of Ic[0].
    M = model.total_mass()
    c_body = sx.vec(Ic[0][5, 1] / M, Ic[0][3, 2] / M, Ic[0][4, 0] / M)
    d = sx.matmul(R, c_body)  # x_com - p_b in inertial coordinates
    Sd = sx.skew(d)
    A_G = sx.zeros(6, 6 + n)
    shift = sx.matmul(Sd, A_mixed[:3, :])
    for j in range(6 + n):
        for i in range(3):
            A_G[i, j] = A_mixed[i, j]
            A_G[3 + i, j] = A_mixed[3 + i, j] - shift[i, j]
    return A_G


def base_pose(p_b, q_b):
    """liecasadi ``SE3.from_position_quaternion(p, q).as_matrix()`` (q assumed unit)."""
    return homogeneous(quat_to_rot(q_b), p_b)
