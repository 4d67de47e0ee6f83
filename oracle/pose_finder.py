"""oracle/pose_finder.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the static-pose NLP of
`/root/reference/src/hippopt/turnkey_planners/humanoid_pose_finder/planner.py:303-413, 444-788`
(BASELINE config 2; the only reference configuration solved with IPOPT's exact Hessian,
`humanoid_pose_finder/main.py:101`).  Defaults of `planner.py:79-92`: CoM and point positions as costs,
hand tasks skipped, PlanarTerrain.

x (81) = contact_points.left[0..3].{p, f}, right[0..3].{p, f}, base position, base quaternion (xyzw),
         joint positions, com                                      
p (202) = 8 descriptors (24), mass, plm, pld, gravity(6), references.state (105, same layout as the
         kinodynamic initial state), references.frame_quaternion_xyzw(4), left/right hand position (6),
         relaxed_complementarity_epsilon, static_friction, max/min joint positions (46),
         left/right hand position in frame (6)
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import expressions as ex
from . import sx
from .kinodynamic import INF, NJ, NPT, _cat, _l
from .nlp import NLP, Template

# offsets in x
XP = lambda i: 6 * i          # noqa: E731
XF = lambda i: 6 * i + 3      # noqa: E731
XPB, XQ, XS, XCOM, NX = 48, 51, 55, 78, 81


@dataclasses.dataclass
class Settings:
    """`humanoid_pose_finder/main.py:75-98`."""

    terrain: ex.Terrain = dataclasses.field(default_factory=ex.PlanarTerrain)
    foot_frames: tuple = ("l_sole", "r_sole")
    frame_quaternion_cost_frame: str = "chest"
    joint_regularization_cost_weights: np.ndarray = dataclasses.field(
        default_factory=lambda: np.concatenate([0.1 * np.ones(3), 10.0 * np.ones(8), np.ones(12)])
    )
    base_quaternion_cost_multiplier: float = 50.0
    desired_frame_quaternion_cost_multiplier: float = 100.0
    joint_regularization_cost_multiplier: float = 0.1
    force_regularization_cost_multiplier: float = 0.2
    com_regularization_cost_multiplier: float = 10.0
    average_force_regularization_cost_multiplier: float = 10.0
    point_position_regularization_cost_multiplier: float = 100.0


class Layout:
    def __init__(self):
        self.n_x = NX
        o = NX
        self.desc0 = o
        o += 24
        self.mass, self.plm, self.pld = o, o + 1, o + 2
        o += 3
        self.gravity = o
        o += 6
        self.ref = o  # references.state: per point (p, f, desc), pb, q, s, com
        o += 105
        self.ref_fq = o
        o += 4
        self.ref_lhand, self.ref_rhand = o, o + 3
        o += 6
        self.eps, self.mu = o, o + 1
        o += 2
        self.max_s, self.min_s = o, o + NJ
        o += 2 * NJ
        self.lhand_in_frame, self.rhand_in_frame = o, o + 3
        o += 6
        self.n_p = o - NX

    @staticmethod
    def rng(a, n):
        return np.arange(a, a + n)

    def ref_pt(self, i, what):
        return self.rng(self.ref + 9 * i + {"p": 0, "f": 3, "desc": 6}[what], 3)

    def ref_state(self, what):
        o, n = {"pb": (72, 3), "q": (75, 4), "s": (79, NJ), "com": (102, 3)}[what]
        return self.rng(self.ref + o, n)


def build(model, st: Settings | None = None) -> tuple[NLP, Layout]:
    st = st or Settings()
    lay = Layout()
    nlp = NLP(lay.n_x, lay.n_p)
    terrain = st.terrain
    p, f = sx.syms("p", 3), sx.syms("f", 3)
    mass, eps, mu = sx.sym("mass"), sx.sym("eps"), sx.sym("mu")
    fm = sx.vec(*[f[i] * mass for i in range(3)])  # point.f * variables.mass (planner.py:716-724)
    t_compl = Template("complementarity", _l(p, f, [mass, eps]), [ex.relaxed_complementarity_margin(terrain, p, fm, eps)],
                       lb=[0.0], ub=[INF])
    t_height = Template("height", _l(p), [terrain.height(p)], lb=[0.0], ub=[INF])
    t_normal = Template("normal", _l(p, f), [ex.normal_force_component(terrain, p, f)], lb=[0.0], ub=[INF])
    t_fric = Template("friction", _l(p, f, [mu]), [ex.friction_cone_square_margin(terrain, p, f, mu)], lb=[0.0], ub=[INF])
    pb, q, s = sx.syms("pb", 3), sx.syms("q", 4), sx.syms("s", NJ)
    qn = ex.quaternion_xyzw_normalization(q)
    p_parent = sx.syms("p_parent", 3)
    t_fk = {}
    for frame in st.foot_frames:
        fk = ex.point_position_from_kinematics(model, frame, pb, qn, s, p_parent)
        t_fk[frame] = Template("fk_" + frame, _l(p, pb, q, s, p_parent), [p[i] - fk[i] for i in range(3)])
    kin = _cat(lay.rng(XPB, 3), lay.rng(XQ, 4), lay.rng(XS, NJ))
    for i in range(NPT):  # planner.py:379-393
        frame = st.foot_frames[0] if i < 4 else st.foot_frames[1]
        xp, xf = lay.rng(XP(i), 3), lay.rng(XF(i), 3)
        nlp.subject_to(t_compl, _cat(xp, xf, lay.mass, lay.eps), f"pt{i}.complementarity")
        nlp.subject_to(t_height, xp, f"pt{i}.height")
        nlp.subject_to(t_normal, _cat(xp, xf), f"pt{i}.normal")
        nlp.subject_to(t_fric, _cat(xp, xf, lay.mu), f"pt{i}.friction")
        nlp.subject_to(t_fk[frame], _cat(xp, kin, lay.rng(lay.desc0 + 3 * i, 3)), f"pt{i}.fk")
    # _add_kinematics_constraints (planner.py:444-521)
    nlp.subject_to(Template("unitary_quaternion", _l(q), [sx.sumsqr(q)], lb=[1.0], ub=[1.0]), lay.rng(XQ, 4),
                   "unitary_quaternion")
    com = sx.syms("com", 3)
    ck = ex.center_of_mass_position_from_kinematics(model, pb, qn, s)
    nlp.subject_to(Template("com_kinematics_consistency", _l(com, pb, q, s), [com[i] - ck[i] for i in range(3)]),
                   _cat(lay.rng(XCOM, 3), kin), "com_kinematics_consistency")
    pts = [sx.syms(f"p{i}", 3) for i in range(NPT)]
    fs = [sx.syms(f"f{i}", 3) for i in range(NPT)]
    grav = sx.syms("grav", 6)
    hdot = ex.centroidal_dynamics_with_point_forces(grav, com, pts, fs)
    nlp.subject_to(Template("centroidal_momentum_dynamics", _l(grav, com, *pts, *fs), list(hdot)),
                   _cat(lay.rng(lay.gravity, 6), lay.rng(XCOM, 3), *[lay.rng(XP(i), 3) for i in range(NPT)],
                        *[lay.rng(XF(i), 3) for i in range(NPT)]), "centroidal_momentum_dynamics")
    lo, hi = sx.syms("lo", NJ), sx.syms("hi", NJ)
    nlp.subject_to(Template("joint_position_bounds", _l(s, lo, hi), list(s), lb=list(lo), ub=list(hi)),
                   _cat(lay.rng(XS, NJ), lay.rng(lay.min_s, NJ), lay.rng(lay.max_s, NJ)), "joint_position_bounds")
    # _add_kinematics_regularization (planner.py:523-594)
    qd = sx.syms("qd", 4)
    nlp.minimize(Template("base_quaternion_error", _l(q, qd), [sx.sumsqr(ex.quaternion_xyzw_error(q, qd))]),
                 _cat(lay.rng(XQ, 4), lay.ref_state("q")), st.base_quaternion_cost_multiplier)
    E = ex.rotation_error_from_kinematics(model, st.frame_quaternion_cost_frame, pb, qn, s, qd)
    nlp.minimize(Template("frame_rotation_error", _l(pb, q, s, qd), [sx.sq((E[0, 0] + E[1, 1] + E[2, 2]) - 3.0)]),
                 _cat(kin, lay.rng(lay.ref_fq, 4)), st.desired_frame_quaternion_cost_multiplier)
    cref = sx.syms("cref", 3)
    nlp.minimize(Template("com_position_error", _l(com, cref), [sx.sumsqr([com[i] - cref[i] for i in range(3)])]),
                 _cat(lay.rng(XCOM, 3), lay.ref_state("com")), st.com_regularization_cost_multiplier)
    sref = sx.syms("sref", NJ)
    w = st.joint_regularization_cost_weights
    acc = sx.const(0.0)
    for i in range(NJ):  # e^T diag(w) e, planner.py:584-588
        e = s[i] - sref[i]
        acc = acc + (e * float(w[i])) * e
    nlp.minimize(Template("joint_positions_error", _l(s, sref), [acc]), _cat(lay.rng(XS, NJ), lay.ref_state("s")),
                 st.joint_regularization_cost_multiplier)
    # _add_foot_regularization x 2 (planner.py:751-788)
    f4 = [sx.syms(f"g{i}", 3) for i in range(4)]
    ssum = [f4[0][c] + f4[1][c] + f4[2][c] + f4[3][c] for c in range(3)]
    t_avg = [Template(f"average_force{i}", _l(*f4), [sx.sumsqr([f4[i][c] - 0.25 * ssum[c] for c in range(3)])])
             for i in range(4)]
    a3, b3 = sx.syms("a", 3), sx.syms("b", 3)
    t_sq = Template("sq3", _l(a3, b3), [sx.sumsqr([a3[i] - b3[i] for i in range(3)])])
    for foot in range(2):
        base = 4 * foot
        for i in range(4):
            nlp.minimize(t_avg[i], _cat(*[lay.rng(XF(base + j), 3) for j in range(4)]),
                         st.average_force_regularization_cost_multiplier)
        for i in range(4):
            nlp.minimize(t_sq, _cat(lay.rng(XP(base + i), 3), lay.ref_pt(base + i, "p")),
                         st.point_position_regularization_cost_multiplier)
            nlp.minimize(t_sq, _cat(lay.rng(XF(base + i), 3), lay.ref_pt(base + i, "f")),
                         st.force_regularization_cost_multiplier)
    return nlp, lay
