"""oracle/expressions.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (on oracle/sx.py graphs) of the reference's expression factories
`/root/reference/src/hippopt/robot_planning/expressions/*.py` and terrain descriptors
`/root/reference/src/hippopt/robot_planning/utilities/{planar_terrain,terrain_descriptor,
smooth_terrain,terrain_sum}.py`.  Each function cites the lines it follows.  Functions take and
return numpy object arrays of SX.
"""
from __future__ import annotations

import math

import numpy as np

from . import robot, sx
from .sx import SX


# ------------------------------------------------------------------ quaternion.py
def quaternion_xyzw_normalization(q):
    """quaternion.py:5-22 (liecasadi ``Quaternion.normalize``: q / ||q||_2)."""
    return robot.quat_normalize(q)


def quaternion_velocity_to_right_trivialized_angular_velocity(q, q_dot):
    """quaternion.py:25-51, line 42: 2(-q_dot_w q_i + q_w q_dot_i - q_dot_i x q_i)."""
    q_w, q_i = q[3], q[:3]
    qd_w, qd_i = q_dot[3], q_dot[:3]
    c = sx.cross(qd_i, q_i)
    return sx.vec(*[2.0 * (-(qd_w * q_i[k]) + q_w * qd_i[k] - c[k]) for k in range(3)])


def _quat_mul(a, b):
    """Hamilton product of xyzw quaternions."""
    av, aw, bv, bw = a[:3], a[3], b[:3], b[3]
    c = sx.cross(av, bv)
    v = [aw * bv[k] + bw * av[k] + c[k] for k in range(3)]
    w = aw * bw - sx.dot(av, bv)
    return sx.vec(*v, w)


def quaternion_xyzw_error(q, qd):
    """quaternion.py:54-85: (qd^-1 (x) q) - identity, liecasadi SO3 product/inverse [ext]:
    the inverse of a (unit) rotation quaternion is its conjugate."""
    qd_inv = sx.vec(-qd[0], -qd[1], -qd[2], qd[3])
    e = _quat_mul(qd_inv, q)
    return sx.vec(e[0], e[1], e[2], e[3] - 1.0)


# ------------------------------------------------------------------ terrains
class Terrain:
    """terrain_descriptor.py:8-138."""

    def height(self, p) -> SX:
        raise NotImplementedError

    def normal(self, p):
        """terrain_descriptor.py:45-61: gradient of h, normalised."""
        grad = sx.gradient(self.height(p), list(p))
        n = sx.norm2(grad)
        return sx.vec(*[g / n for g in grad])

    def orientation(self, p):
        """terrain_descriptor.py:63-80: columns [x, y, n]."""
        n = self.normal(p)
        y = sx.cross(n, sx.vec(1.0, 0.0, 0.0))
        x = sx.cross(y, n)
        xn = sx.norm2(x)
        x = sx.vec(*[xi / xn for xi in x])
        y = sx.cross(n, x)
        R = sx.zeros(3, 3)
        for i in range(3):
            R[i, 0], R[i, 1], R[i, 2] = x[i], y[i], n[i]
        return R


class PlanarTerrain(Terrain):
    """planar_terrain.py:6-41: h = p_z, n = eye(3)[:, 2], R = eye(3) (structural constants)."""

    def height(self, p):
        return p[2]

    def normal(self, p):
        return sx.vec(0.0, 0.0, 1.0)

    def orientation(self, p):
        return sx.eye(3)


class SmoothStep(Terrain):
    """smooth_terrain.py:201-227 (height), 266-336 (``step``): a smooth box of ``height``
    centred at ``origin`` with footprint ``length`` x ``width`` rotated by ``yaw``:
    h = z - exp(-g^(2 s)) * height,  g = (2 x_t / l)^(2 e) + (2 y_t / w)^(2 e),
    (x_t, y_t) = R_z(yaw)^T (p - origin)_{xy}.  Every parameter may be an SX symbol, which is how
    the evaluator makes the step height runtime data (BASELINE.json config 5)."""

    def __init__(self, length, width, height, origin=(0.0, 0.0, 0.0), yaw=0.0,
                 edge_exponent=5.0, sharpness=10.0):
        self.length, self.width, self.h = length, width, height
        self.origin, self.yaw = origin, yaw
        self.e, self.s = float(edge_exponent), float(sharpness)

    def height(self, p):
        dx = p[0] - self.origin[0]
        dy = p[1] - self.origin[1]
        c, s = sx.cos(self.yaw), sx.sin(self.yaw)
        xt = c * dx + s * dy
        yt = c * dy - s * dx
        g = sx.powc(2.0 * xt / self.length, 2.0 * self.e) + sx.powc(2.0 * yt / self.width, 2.0 * self.e)
        zt = sx.exp(-sx.powc(g, 2.0 * self.s)) * self.h
        return p[2] - (zt + self.origin[2])


class TerrainSum(Terrain):
    """terrain_sum.py:19-38: h = h_lhs + h_rhs - p_z."""

    def __init__(self, lhs: Terrain, rhs: Terrain):
        self.lhs, self.rhs = lhs, rhs

    def height(self, p):
        return self.lhs.height(p) + self.rhs.height(p) - p[2]


class TwoSmoothSteps(Terrain):
    """The stairs of `main_walking_on_stairs.py:18-28`: ``SmoothTerrain.step(...) +
    SmoothTerrain.step(...)`` (a TerrainSum).  The reference bakes length / width / height / origin
    into the graph as Python floats (`smooth_terrain.py:271,306-308`); here they are the ten runtime
    parameters (l, w, h, ox, oy) x 2 so that a batch can randomise the step heights."""

    N_PARAMS = 10

    def __init__(self, params=None):
        self.params = params

    def with_params(self, tp):
        assert len(tp) == self.N_PARAMS
        return TwoSmoothSteps(tp)

    def _sum(self):
        tp = self.params
        a = SmoothStep(tp[0], tp[1], tp[2], origin=(tp[3], tp[4], 0.0))
        b = SmoothStep(tp[5], tp[6], tp[7], origin=(tp[8], tp[9], 0.0))
        return TerrainSum(a, b)

    def height(self, p):
        return self._sum().height(p)

    @staticmethod
    def stairs_parameters(step_length=0.9, width=0.8, height=0.1):
        """Numeric values of `main_walking_on_stairs.py:18-28,397-403` (length = step_length / 2)."""
        L = step_length / 2.0
        return np.array([2 * L, width, height, 1.5 * L, 0.0, 0.9 * L, width, height, 2 * L, 0.0])


def jtimes(expr, wrt, direction):
    """``cs.jtimes(expr, wrt, v)`` for a scalar or vector expr: J(expr, wrt) @ v."""
    exprs = [expr] if isinstance(expr, SX) else list(expr)
    out = []
    for e in exprs:
        if e.op == sx.OP_CONST:
            out.append(sx.const(0.0))
        else:
            g = sx.gradient(e, list(wrt))
            out.append(sx.dot(g, direction))
    return out[0] if isinstance(expr, SX) else sx.vec(*out)


# ------------------------------------------------------------------ complementarity.py
def dcc_planar_complementarity(terrain: Terrain, p, kt, u_p):
    """complementarity.py:6-41, lines 30-32: R_t diag(tau, tau, 1) u, tau = tanh(kt h(p))."""
    tau = sx.tanh(kt * terrain.height(p))
    scaled = sx.vec(tau * u_p[0], tau * u_p[1], u_p[2])
    return sx.matmul(terrain.orientation(p), scaled)


def dcc_complementarity_margin(terrain: Terrain, p, f, v, f_dot, k_bs, eps):
    """complementarity.py:44-110, lines 68-89."""
    h = terrain.height(p)
    n = terrain.normal(p)
    h_dot = jtimes(h, p, v)
    n_dot = jtimes(n, p, v)
    normal_force = sx.dot(n, f)
    normal_force_derivative = sx.dot(n, f_dot)
    complementarity = h * normal_force
    csi = h_dot * normal_force + h * sx.dot(f, n_dot) + h * normal_force_derivative
    return eps - k_bs * complementarity - csi


def relaxed_complementarity_margin(terrain: Terrain, p, f, eps):
    """complementarity.py:113-155, lines 131-138."""
    return eps - terrain.height(p) * sx.dot(terrain.normal(p), f)


# ------------------------------------------------------------------ contacts.py
def normal_force_component(terrain: Terrain, p, f):
    """contacts.py:6-33, line 24."""
    return sx.dot(terrain.normal(p), f)


def friction_cone_square_margin(terrain: Terrain, p, f, mu):
    """contacts.py:36-75, lines 54-66: [-1, -1, mu^2] . (R_t^T f)^2."""
    R = terrain.orientation(p)
    fc = [sx.dot(R[:, k], f) for k in range(3)]
    return -sx.sq(fc[0]) - sx.sq(fc[1]) + sx.sq(mu) * sx.sq(fc[2])


def contact_points_centroid(points):
    """contacts.py:78-116, lines 101-107."""
    acc = sx.zeros(3)
    for pt in points:
        for k in range(3):
            acc[k] = acc[k] + pt[k]
    return sx.vec(*[acc[k] / float(len(points)) for k in range(3)])


def contact_points_yaw_alignment_error(p0, p1, yaw):
    """contacts.py:119-141, line 132."""
    return -sx.sin(yaw) * (p1[0] - p0[0]) + sx.cos(yaw) * (p1[1] - p0[1])


def swing_height_heuristic(terrain: Terrain, p, v, hd):
    """contacts.py:144-175, lines 158-166."""
    R = terrain.orientation(p)
    pv = [sx.dot(R[:, k], v) for k in range(2)]
    return 0.5 * (sx.sq(terrain.height(p) - hd) + (sx.sq(pv[0]) + sx.sq(pv[1])))


# ------------------------------------------------------------------ centroidal.py
def centroidal_dynamics_with_point_forces(gravity, com, points, forces, mass=1.0):
    """centroidal.py:4-73, lines 62-64 (``assume_unitary_mass=True`` -> m = 1.0)."""
    out = sx.vec(*[mass * gi for gi in gravity])
    for p, f in zip(points, forces):
        arm = sx.vec(*[p[k] - com[k] for k in range(3)])
        c = sx.cross(arm, f)
        for k in range(3):
            out[k] = out[k] + f[k]
            out[3 + k] = out[3 + k] + c[k]
    return out


# ------------------------------------------------------------------ kinematics.py
def point_position_from_kinematics(model, frame, pb, qb, s, p_parent):
    """kinematics.py:217-308, lines 249-265 (qb is supposed normalised)."""
    H = robot.frame_transform(model, robot.base_pose(pb, qb), s, frame)
    return sx.vec(*[sx.dot(H[i, :3], p_parent) + H[i, 3] for i in range(3)])


def center_of_mass_position_from_kinematics(model, pb, qb, s):
    """kinematics.py:134-214, lines 163-167,197."""
    return robot.com_position(model, robot.base_pose(pb, qb), s)


def centroidal_momentum_from_kinematics(model, pb, qb, s, pb_dot, qb_dot, s_dot):
    """kinematics.py:11-131, lines 47-69,108: A_G(H_b, s) @ [pb_dot; omega; s_dot]."""
    A = robot.centroidal_momentum_matrix(model, robot.base_pose(pb, qb), s)
    omega = quaternion_velocity_to_right_trivialized_angular_velocity(qb, qb_dot)
    nu = sx.vec(*pb_dot, *omega, *s_dot)
    return sx.matmul(A, nu)


def frames_relative_position(model, reference_frame, target_frame, s):
    """kinematics.py:311-394, lines 337-367 (identity base pose)."""
    Hb = sx.eye(4)
    Hr = robot.frame_transform(model, Hb, s, reference_frame)
    Ht = robot.frame_transform(model, Hb, s, target_frame)
    out = []
    for i in range(3):
        # R_ref^T t_tgt + (-R_ref^T t_ref)
        a = sx.dot(Hr[:3, i], Ht[:3, 3])
        b = -sx.dot(Hr[:3, i], Hr[:3, 3])
        out.append(a + b)
    return sx.vec(*out)


def rotation_error_from_kinematics(model, frame, pb, qb, s, qd):
    """kinematics.py:397-491, lines 444-448: R_frame(qb, s) @ R(qd)^T."""
    H = robot.frame_transform(model, robot.base_pose(pb, qb), s, frame)
    Rd = robot.quat_to_rot(qd)
    E = sx.zeros(3, 3)
    for i in range(3):
        for j in range(3):
            E[i, j] = sx.dot(H[i, :3], Rd[j, :])
    return E
