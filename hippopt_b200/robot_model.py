"""Reduced floating-base robot model used by the kinematics kernels.

The reference obtains its kinematic tree from ``adam.casadi.KinDynComputations(urdfstring=,
joints_name_list=, root_link=, ...)`` (`/root/reference/src/hippopt/turnkey_planners/
humanoid_kinodynamic/planner.py:43-49`) on the ergoCub URDF
(`main_single_step_flat_ground.py:17-19`, 23 actuated joints `:22-46`).  That URDF is not
available offline, so :func:`synthetic_ergocub_urdf` writes a floating-base tree of identical
topology and DoF (root_link; 3-joint torso chain -> chest; two 4-joint arms hanging from the
chest; two 6-joint legs hanging from the root, each ending in a fixed ``*_sole`` frame; a few
extra fixed/un-listed joints that must be lumped) as a URDF *string*, and
:func:`RobotModel.from_urdf` reduces any such URDF exactly the way the reference's call does:
joints that are not in ``joints_name_list`` are frozen at zero and their child links are lumped
into the parent body [ext: adam's model reduction].

Everything here is host-side data preparation (numpy only); the arrays end up in the device
constant block of the kinodynamic kernels (csrc/kinodynamic.cu).
"""
from __future__ import annotations

import dataclasses
import xml.etree.ElementTree as ET

import numpy as np

ERGOCUB_JOINTS = [
    "torso_pitch", "torso_roll", "torso_yaw",
    "l_shoulder_pitch", "l_shoulder_roll", "l_shoulder_yaw", "l_elbow",
    "r_shoulder_pitch", "r_shoulder_roll", "r_shoulder_yaw", "r_elbow",
    "l_hip_pitch", "l_hip_roll", "l_hip_yaw", "l_knee", "l_ankle_pitch", "l_ankle_roll",
    "r_hip_pitch", "r_hip_roll", "r_hip_yaw", "r_knee", "r_ankle_pitch", "r_ankle_roll",
]  # main_single_step_flat_ground.py:22-46


def rpy_to_matrix(rpy) -> np.ndarray:
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def _skew(v) -> np.ndarray:
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


@dataclasses.dataclass
class RobotModel:
    """Bodies are numbered 0 (root) .. n_joints; body i>0 is moved by joint i-1.

    ``parent[i]``      parent body of body i (``-1`` for the root), always ``< i``;
    ``joint_rot[i]``   constant rotation parent-frame <- joint-frame (URDF origin rpy);
    ``joint_xyz[i]``   joint origin in the parent frame;
    ``joint_axis[i]``  unit revolute axis in the child frame;
    ``mass/com/inertia[i]`` lumped inertial of body i, ``com`` and ``inertia`` (about the
                       CoM) expressed in the body frame;
    ``frames[name]``   ``(body, R, xyz)`` of a frame rigidly attached to a body.
    """

    joint_names: list[str]
    parent: np.ndarray
    joint_rot: np.ndarray
    joint_xyz: np.ndarray
    joint_axis: np.ndarray
    mass: np.ndarray
    com: np.ndarray
    inertia: np.ndarray
    frames: dict[str, tuple[int, np.ndarray, np.ndarray]]
    body_names: list[str]

    @property
    def n_joints(self) -> int:
        return len(self.joint_names)

    @property
    def n_bodies(self) -> int:
        return len(self.parent)

    def total_mass(self) -> float:
        return float(self.mass.sum())

    def chain_to_root(self, body: int) -> list[int]:
        out = []
        while body > 0:
            out.append(body)
            body = int(self.parent[body])
        return out[::-1]

    def subtree_mask(self) -> np.ndarray:
        """``m[j, l]`` is True when body l is in the subtree rooted at body j."""
        n = self.n_bodies
        m = np.zeros((n, n), dtype=bool)
        for l in range(n):
            b = l
            while b >= 0:
                m[b, l] = True
                b = int(self.parent[b])
        return m

    # ---------------------------------------------------------------------------------
    @staticmethod
    def from_urdf(urdfstring: str, joints_name_list: list[str], root_link: str = "root_link",
                  frames: list[str] | None = None) -> "RobotModel":
        root = ET.fromstring(urdfstring)
        links = {}
        for link in root.findall("link"):
            inertial = link.find("inertial")
            if inertial is None:
                links[link.get("name")] = (0.0, np.zeros(3), np.eye(3), np.zeros((3, 3)))
                continue
            org = inertial.find("origin")
            xyz = np.array([float(v) for v in (org.get("xyz", "0 0 0") if org is not None else "0 0 0").split()])
            rpy = [float(v) for v in (org.get("rpy", "0 0 0") if org is not None else "0 0 0").split()]
            m = float(inertial.find("mass").get("value"))
            it = inertial.find("inertia")
            ixx, ixy, ixz, iyy, iyz, izz = (float(it.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz"))
            inr = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
            links[link.get("name")] = (m, xyz, rpy_to_matrix(rpy), inr)
        children: dict[str, list] = {}
        for joint in root.findall("joint"):
            org = joint.find("origin")
            xyz = np.array([float(v) for v in (org.get("xyz", "0 0 0") if org is not None else "0 0 0").split()])
            rpy = [float(v) for v in (org.get("rpy", "0 0 0") if org is not None else "0 0 0").split()]
            axis_el = joint.find("axis")
            axis = np.array([float(v) for v in axis_el.get("xyz").split()]) if axis_el is not None else np.array([1.0, 0, 0])
            children.setdefault(joint.find("parent").get("link"), []).append(
                (joint.get("name"), joint.get("type"), joint.find("child").get("link"), xyz, rpy_to_matrix(rpy), axis)
            )
        for name in joints_name_list:
            if not any(j[0] == name and j[1] in ("revolute", "continuous") for js in children.values() for j in js):
                raise ValueError(f"joint {name} is not a revolute joint of the URDF")
        order = {n: i for i, n in enumerate(joints_name_list)}
        nb = len(joints_name_list) + 1
        parent = -np.ones(nb, dtype=np.int32)
        jrot = np.tile(np.eye(3), (nb, 1, 1))
        jxyz = np.zeros((nb, 3))
        jaxis = np.zeros((nb, 3))
        jaxis[:, 2] = 1.0
        mass = np.zeros(nb)
        mcom = np.zeros((nb, 3))  # m * com accumulated in the body frame
        parts: list[list] = [[] for _ in range(nb)]  # (m, com_in_body, R_in_body, inertia_at_com)
        frame_out: dict[str, tuple[int, np.ndarray, np.ndarray]] = {}
        body_names = [""] * nb
        wanted = set(frames) if frames is not None else None

        def visit(link_name: str, body: int, R: np.ndarray, t: np.ndarray) -> None:
            # (R, t): pose of `link_name` in the frame of reduced body `body`
            m, c, Rc, inr = links[link_name]
            if m > 0.0:
                parts[body].append((m, R @ c + t, R @ Rc, inr))
            if wanted is None or link_name in wanted:
                frame_out[link_name] = (body, R.copy(), t.copy())
            for (jn, jt, child, xyz, Rj, axis) in children.get(link_name, []):
                if jn in order:
                    b = order[jn] + 1
                    parent[b] = body
                    jrot[b] = R @ Rj
                    jxyz[b] = R @ xyz + t
                    jaxis[b] = axis / np.linalg.norm(axis)
                    body_names[b] = child
                    visit(child, b, np.eye(3), np.zeros(3))
                else:  # fixed, or a movable joint frozen at zero -> lump into this body
                    visit(child, body, R @ Rj, R @ xyz + t)

        body_names[0] = root_link
        visit(root_link, 0, np.eye(3), np.zeros(3))
        if np.any(parent[1:] < 0):
            raise ValueError("some listed joints are not reachable from the root link")
        inertia = np.zeros((nb, 3, 3))
        com = np.zeros((nb, 3))
        for b in range(nb):
            mass[b] = sum(p[0] for p in parts[b])
            if mass[b] <= 0.0:
                raise ValueError(f"body {body_names[b]} has no mass after lumping")
            com[b] = sum(p[0] * p[1] for p in parts[b]) / mass[b]
            for (m, c, R, inr) in parts[b]:
                d = c - com[b]
                inertia[b] += R @ inr @ R.T - m * _skew(d) @ _skew(d)
        # bodies must be topologically ordered (parent < child) for the device recursions;
        # the joint order of joints_name_list is kept, so verify rather than permute.
        for b in range(1, nb):
            if parent[b] >= b:
                raise ValueError("joints_name_list must list every joint after its ancestors")
        return RobotModel(list(joints_name_list), parent, jrot, jxyz, jaxis, mass, com, inertia,
                          frame_out, body_names)


# -------------------------------------------------------------------------------------
def synthetic_ergocub_urdf(seed: int = 7) -> str:
    """URDF string of a synthetic humanoid with ergoCub's reduced topology (see module doc)."""
    rng = np.random.default_rng(seed)

    def jitter(v, s):
        return np.asarray(v, dtype=np.float64) + rng.normal(0.0, s, 3)

    links: list[str] = []
    joints: list[str] = []

    def add_link(name, mass, com, size):
        if mass is None:
            links.append(f'  <link name="{name}"/>')
            return
        com = jitter(com, 0.005)
        rpy = rng.normal(0.0, 0.15, 3)
        sx, sy, sz = (abs(s) + 0.02 for s in jitter(size, 0.005))
        ixx = mass * (sy * sy + sz * sz) / 12.0
        iyy = mass * (sx * sx + sz * sz) / 12.0
        izz = mass * (sx * sx + sy * sy) / 12.0
        ixy, ixz, iyz = (0.05 * np.sqrt(a * b) * rng.uniform(-1, 1) for a, b in ((ixx, iyy), (ixx, izz), (iyy, izz)))
        links.append(
            f'  <link name="{name}">\n    <inertial>\n'
            f'      <origin xyz="{com[0]:.17g} {com[1]:.17g} {com[2]:.17g}" rpy="{rpy[0]:.17g} {rpy[1]:.17g} {rpy[2]:.17g}"/>\n'
            f'      <mass value="{mass:.17g}"/>\n'
            f'      <inertia ixx="{ixx:.17g}" ixy="{ixy:.17g}" ixz="{ixz:.17g}" iyy="{iyy:.17g}" iyz="{iyz:.17g}" izz="{izz:.17g}"/>\n'
            f'    </inertial>\n  </link>'
        )

    def add_joint(name, jtype, parent, child, xyz, axis=None, rpy_sigma=0.05, exact=False):
        xyz = np.asarray(xyz, dtype=np.float64) if exact else jitter(xyz, 0.003)
        rpy = np.zeros(3) if exact else rng.normal(0.0, rpy_sigma, 3)
        s = f'  <joint name="{name}" type="{jtype}">\n    <parent link="{parent}"/>\n    <child link="{child}"/>\n'
        s += f'    <origin xyz="{xyz[0]:.17g} {xyz[1]:.17g} {xyz[2]:.17g}" rpy="{rpy[0]:.17g} {rpy[1]:.17g} {rpy[2]:.17g}"/>\n'
        if axis is not None:
            a = jitter(axis, 0.04)
            a /= np.linalg.norm(a)
            s += f'    <axis xyz="{a[0]:.17g} {a[1]:.17g} {a[2]:.17g}"/>\n'
            s += '    <limit lower="-1.5" upper="1.5" effort="100" velocity="10"/>\n'
        s += "  </joint>"
        joints.append(s)

    X, Y, Z = [1, 0, 0], [0, 1, 0], [0, 0, 1]
    add_link("root_link", 8.0, [0, 0, 0.02], [0.2, 0.25, 0.15])
    # torso chain
    add_link("torso_1", 1.0, [0, 0, 0.02], [0.08, 0.08, 0.06])
    add_joint("torso_pitch", "revolute", "root_link", "torso_1", [0, 0, 0.08], Y)
    add_link("torso_2", 1.0, [0, 0, 0.02], [0.08, 0.08, 0.06])
    add_joint("torso_roll", "revolute", "torso_1", "torso_2", [0, 0, 0.04], X)
    add_link("chest", 9.0, [0, 0, 0.15], [0.2, 0.3, 0.3])
    add_joint("torso_yaw", "revolute", "torso_2", "chest", [0, 0, 0.04], Z)
    # head: neck joints exist in the URDF but are not in joints_name_list -> lumped into chest
    add_link("neck_1", 0.5, [0, 0, 0.02], [0.05, 0.05, 0.05])
    add_joint("neck_pitch", "revolute", "chest", "neck_1", [0, 0, 0.33], Y)
    add_link("head", 2.5, [0.01, 0, 0.08], [0.15, 0.15, 0.2])
    add_joint("neck_yaw", "revolute", "neck_1", "head", [0, 0, 0.04], Z)
    for side, sgn in (("l", 1.0), ("r", -1.0)):
        add_link(f"{side}_shoulder_1", 0.8, [0, 0, 0], [0.06, 0.06, 0.06])
        add_joint(f"{side}_shoulder_pitch", "revolute", "chest", f"{side}_shoulder_1", [0, sgn * 0.11, 0.25], Y)
        add_link(f"{side}_shoulder_2", 0.8, [0, 0, 0], [0.06, 0.06, 0.06])
        add_joint(f"{side}_shoulder_roll", "revolute", f"{side}_shoulder_1", f"{side}_shoulder_2", [0, sgn * 0.05, 0], X)
        add_link(f"{side}_upperarm", 1.6, [0, 0, -0.08], [0.07, 0.07, 0.2])
        add_joint(f"{side}_shoulder_yaw", "revolute", f"{side}_shoulder_2", f"{side}_upperarm", [0, 0, -0.03], Z)
        add_link(f"{side}_forearm", 1.0, [0, 0, -0.08], [0.06, 0.06, 0.18])
        add_joint(f"{side}_elbow", "revolute", f"{side}_upperarm", f"{side}_forearm", [0, 0, -0.19], Y)
        add_link(f"{side}_hand", 0.6, [0, 0, -0.04], [0.05, 0.08, 0.1])
        add_joint(f"{side}_wrist_yaw", "revolute", f"{side}_forearm", f"{side}_hand", [0, 0, -0.19], Z)
    for side, sgn in (("l", 1.0), ("r", -1.0)):
        add_link(f"{side}_hip_1", 1.2, [0, 0, 0], [0.08, 0.08, 0.08])
        add_joint(f"{side}_hip_pitch", "revolute", "root_link", f"{side}_hip_1", [0, sgn * 0.08, -0.06], Y)
        add_link(f"{side}_hip_2", 1.2, [0, 0, 0], [0.08, 0.08, 0.08])
        add_joint(f"{side}_hip_roll", "revolute", f"{side}_hip_1", f"{side}_hip_2", [0, sgn * 0.02, 0], X)
        add_link(f"{side}_upper_leg", 4.0, [0, 0, -0.15], [0.12, 0.12, 0.3])
        add_joint(f"{side}_hip_yaw", "revolute", f"{side}_hip_2", f"{side}_upper_leg", [0, 0, -0.05], Z)
        add_link(f"{side}_lower_leg", 2.8, [0, 0, -0.14], [0.09, 0.09, 0.3])
        add_joint(f"{side}_knee", "revolute", f"{side}_upper_leg", f"{side}_lower_leg", [0, 0, -0.3], Y)
        add_link(f"{side}_ankle_1", 0.8, [0, 0, 0], [0.06, 0.06, 0.06])
        add_joint(f"{side}_ankle_pitch", "revolute", f"{side}_lower_leg", f"{side}_ankle_1", [0, 0, -0.3], Y)
        add_link(f"{side}_ankle_2", 1.0, [0.02, 0, -0.03], [0.2, 0.09, 0.05])
        add_joint(f"{side}_ankle_roll", "revolute", f"{side}_ankle_1", f"{side}_ankle_2", [0, 0, -0.02], X)
        add_link(f"{side}_sole", None, None, None)
        add_joint(f"{side}_sole_fixed_joint", "fixed", f"{side}_ankle_2", f"{side}_sole", [0.03, 0, -0.06], None, exact=True)
    return '<?xml version="1.0"?>\n<robot name="synthetic_ergocub">\n' + "\n".join(links) + "\n" + "\n".join(joints) + "\n</robot>\n"


def synthetic_ergocub(seed: int = 7) -> RobotModel:
    return RobotModel.from_urdf(
        synthetic_ergocub_urdf(seed), ERGOCUB_JOINTS, "root_link", frames=["l_sole", "r_sole", "chest", "root_link"]
    )
