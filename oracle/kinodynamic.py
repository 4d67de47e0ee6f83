"""oracle/kinodynamic.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the NLP that `/root/reference/src/hippopt/turnkey_planners/
humanoid_kinodynamic/planner.py:27-176` hands to CasADi: same decision-vector layout
(`variables.py:121-301`), same ``subject_to`` / ``minimize`` call
order (`planner.py:124-176`, Appendix B.3) and the same canonical (g, lbg, ubg) forms CasADi's
``Opti`` derives from each constraint [ext]: ``lhs == rhs`` with both sides depending on
decision variables -> ``lhs - rhs`` in [0, 0]; one side parametric -> the other side bounded by
it; ``e >= c`` -> ``e`` in [c, inf]; ``Opti_bounded(l, e, u)`` -> ``e`` in [l, u].

Built on oracle/{sx,nlp,robot,expressions}.py.  parity status: UNPINNED against CasADi.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import expressions as ex
from . import sx
from .nlp import NLP, Template

INF = float("inf")

# offsets inside one knot of x (Appendix B.1)
NZ = 189
V, FD, P, F, U = 0, 3, 6, 9, 12
VB, QD, PB, Q, SD, S, COM, H = 120, 123, 127, 130, 134, 157, 180, 183
NJ = 23
NPT = 8


@dataclasses.dataclass
class Settings:
    """Numeric settings of `main_single_step_flat_ground.py:54-104` (config 3) by default."""

    horizon: int = 30
    terrain: ex.Terrain = dataclasses.field(default_factory=ex.PlanarTerrain)
    terrain_params: int = 0  # number of runtime terrain parameters appended to p
    final_state_constraint: bool = False  # ExpressionType.subject_to vs skip (planner.py:417)
    periodicity_constraint: bool = False  # planner.py:923
    foot_frames: tuple = ("l_sole", "r_sole")
    frame_quaternion_cost_frame: str = "chest"
    joint_regularization_cost_weights: np.ndarray = dataclasses.field(
        default_factory=lambda: np.concatenate([0.1 * np.ones(3), 10.0 * np.ones(8), np.ones(12)])
    )
    com_linear_velocity_cost_weights: tuple = (10.0, 0.1, 1.0)
    contacts_centroid_cost_multiplier: float = 100.0
    com_linear_velocity_cost_multiplier: float = 1.0
    desired_frame_quaternion_cost_multiplier: float = 90.0
    base_quaternion_cost_multiplier: float = 50.0
    base_quaternion_velocity_cost_multiplier: float = 0.001
    joint_regularization_cost_multiplier: float = 0.1
    force_regularization_cost_multiplier: float = 10.0
    foot_yaw_regularization_cost_multiplier: float = 2000.0
    swing_foot_height_cost_multiplier: float = 1000.0
    contact_velocity_control_cost_multiplier: float = 5.0
    contact_force_control_cost_multiplier: float = 0.0001
    # rectangular_foot(...) (variables/contacts.py:38-65): top-left, bottom-left, bottom-right, top-right
    yaw_bottom_right: int = 2
    yaw_top_right: int = 3
    yaw_top_left: int = 0


class Layout:
    """Index arithmetic of x and p (Appendix B.1 / B.2); p slots are offset by n_x (w = [x; p])."""

    def __init__(self, N: int, n_terrain: int = 0):
        self.N = N
        self.n_x = NZ * N + 6
        o = self.n_x
        self.desc0 = o
        o += 24 * N
        self.mass, self.plm, self.pld = o, o + 1, o + 2
        o += 3
        self.init = o
        o += 105
        self.final = o
        o += 105
        self.dt = o
        o += 1
        self.gravity = o
        o += 6
        self.kt, self.k_bs, self.eps, self.mu = o, o + 1, o + 2, o + 3
        o += 4
        self.max_u = o
        o += 3
        self.max_fd = o
        o += 3
        self.max_L, self.min_com_h, self.min_feet_d, self.max_feet_h = o, o + 1, o + 2, o + 3
        o += 4
        self.max_s, self.min_s, self.max_sd, self.min_sd = o, o + NJ, o + 2 * NJ, o + 3 * NJ
        o += 4 * NJ
        self.refs0 = o
        o += 55 * N
        self.terrain = o
        o += n_terrain
        self.n_p = o - self.n_x
        self.h_init = NZ * N  # initial_state.centroidal_momentum: a *variable* (variables.py:240)

    def x(self, k, off, n=1):
        return np.arange(NZ * k + off, NZ * k + off + n)

    def pt(self, k, i):
        return self.x(k, 15 * i, 15)

    def desc(self, k, i):
        return np.arange(self.desc0 + 24 * k + 3 * i, self.desc0 + 24 * k + 3 * i + 3)

    # initial/final state blocks: per point (p, f, descriptor), base position, quaternion, joints, com
    def state_pt(self, base, i, what):
        o = base + 9 * i + {"p": 0, "f": 3, "desc": 6}[what]
        return np.arange(o, o + 3)

    def state(self, base, what):
        o, n = {"pb": (72, 3), "q": (75, 4), "s": (79, NJ), "com": (102, 3)}[what]
        return np.arange(base + o, base + o + n)

    def ref(self, k, what):
        o, n = {
            "ratio_l": (0, 4), "yaw_l": (4, 1), "ratio_r": (5, 4), "yaw_r": (9, 1), "swing_h": (10, 1),
            "cw": (11, 3), "cc": (14, 3), "comv": (17, 3), "fq": (20, 4), "bq": (24, 4), "bqv": (28, 4),
            "jr": (32, NJ),
        }[what]
        return np.arange(self.refs0 + 55 * k + o, self.refs0 + 55 * k + o + n)

    def rng(self, start, n=1):
        return np.arange(start, start + n)


def _cat(*arrs):
    return np.concatenate([np.atleast_1d(np.asarray(a, dtype=np.int64)) for a in arrs])


def build(model, st: Settings) -> tuple[NLP, Layout]:
    N = st.horizon
    lay = Layout(N, st.terrain_params)
    nlp = NLP(lay.n_x, lay.n_p)
    terrain = st.terrain
    tpar = sx.syms("tp", st.terrain_params)
    if st.terrain_params:
        terrain = terrain.with_params(tpar)
    tp_idx = lay.rng(lay.terrain, st.terrain_params)
    knots_all = range(N)
    knots_1 = range(1, N)

    # ---------------------------------------------------------------- generic linear pieces
    def ic_template(n, name):
        x0 = sx.syms("x0", n)
        ini = sx.syms("ini", n)
        return Template(name, _l(x0, ini), list(x0), lb=list(ini), ub=list(ini))

    def ic_template_var(n, name):  # both sides are decision variables
        x0 = sx.syms("x0", n)
        ini = sx.syms("ini", n)
        return Template(name, _l(x0, ini), [x0[i] - ini[i] for i in range(n)])

    def trapezoid_linear(n, name):
        """x_next == x_k + 0.5 dt (xdot_k + xdot_next) (implicit_trapezoid.py:31-37)."""
        xk, xn, dk, dn = sx.syms("xk", n), sx.syms("xn", n), sx.syms("dk", n), sx.syms("dn", n)
        dt = sx.sym("dt")
        rows = [xn[i] - (xk[i] + 0.5 * dt * (dk[i] + dn[i])) for i in range(n)]
        return Template(name, _l(xk, xn, dk, dn, [dt]), rows)

    def add_linear_dynamics(n, name, state_off, rate_off, init_idx, init_is_var=False):
        t_ic = (ic_template_var if init_is_var else ic_template)(n, name + "_ic")
        # names: multiple_shooting_solver.py:696-697 (x0_name = name + "[0]"), :728 (name + "[k]"); the dynamics
        # are handed to Problem.add_expression as a generator over the state variables, which appends "{j}"
        # (problem.py:105-110, 140-145) -- one state variable per call here, hence "{0}"
        if init_idx is not None:
            nlp.subject_to(t_ic, _cat(lay.x(0, state_off, n), init_idx), name + "[0]{0}")
        t = trapezoid_linear(n, name)
        for k in range(N - 1):
            nlp.subject_to(
                t,
                _cat(lay.x(k, state_off, n), lay.x(k + 1, state_off, n), lay.x(k, rate_off, n),
                     lay.x(k + 1, rate_off, n), lay.dt),
                f"{name}[{k + 1}]{{0}}",
            )

    # ---------------------------------------------------------------- per-point templates
    pt = sx.syms("pt", 15)
    pv, pfd, pp, pf, pu = pt[V:V + 3], pt[FD:FD + 3], pt[P:P + 3], pt[F:F + 3], pt[U:U + 3]
    kt, k_bs, eps, mu, mass = (sx.sym(n) for n in ("kt", "k_bs", "eps", "mu", "mass"))
    max_u, max_fd = sx.syms("max_u", 3), sx.syms("max_fd", 3)
    hd = sx.sym("hd")

    dcc_planar = ex.dcc_planar_complementarity(terrain, pp, kt, pu)
    t_planar = Template("planar_complementarity", _l(pt, [kt], tpar), [pv[i] - dcc_planar[i] for i in range(3)])
    t_dcc = Template("dcc", _l(pt, [k_bs, eps], tpar),
                     [ex.dcc_complementarity_margin(terrain, pp, pf, pv, pfd, k_bs, eps)], lb=[0.0], ub=[INF])
    t_height = Template("height", _l(pt, tpar), [terrain.height(pp)], lb=[0.0], ub=[INF])
    t_normal = Template("normal", _l(pt, tpar), [ex.normal_force_component(terrain, pp, pf)], lb=[0.0], ub=[INF])
    t_friction = Template("friction", _l(pt, [mu], tpar), [ex.friction_cone_square_margin(terrain, pp, pf, mu)],
                          lb=[0.0], ub=[INF])
    t_ubounds = Template("u_v_bounds", _l(pt, max_u), list(pu), lb=[-m for m in max_u], ub=list(max_u))
    t_fdbounds = Template("f_dot_bounds", _l(pt, [mass], max_fd), [pfd[i] * mass for i in range(3)],
                          lb=[-m for m in max_fd], ub=list(max_fd))
    t_swing = Template("swing", _l(pt, [hd], tpar), [ex.swing_height_heuristic(terrain, pp, pv, hd)])
    t_ureg = Template("u_v_reg", _l(pt), [sx.sumsqr(pu)])
    t_fdreg = Template("f_dot_reg", _l(pt), [sx.sumsqr(pfd)])

    pb, q, s = sx.syms("pb", 3), sx.syms("q", 4), sx.syms("s", NJ)
    vb, qd, sd = sx.syms("vb", 3), sx.syms("qd", 4), sx.syms("sd", NJ)
    qn = ex.quaternion_xyzw_normalization(q)  # planner.py:88-93
    p_parent = sx.syms("p_parent", 3)
    t_fk = {}
    for frame in st.foot_frames:
        fk = ex.point_position_from_kinematics(model, frame, pb, qn, s, p_parent)
        t_fk[frame] = Template("fk_" + frame, _l(pt, pb, q, s, p_parent), [pp[i] - fk[i] for i in range(3)])

    def kin_idx(k):
        return _cat(lay.x(k, PB, 3), lay.x(k, Q, 4), lay.x(k, S, NJ))

    # ---------------------------------------------------------------- emission (planner.py:124-147)
    for i in range(NPT):
        frame = st.foot_frames[0] if i < 4 else st.foot_frames[1]
        # flattened symbol names (multiple_shooting_solver.py:292-337, 395-401): lists are flattened with "[k]"
        name = f"system.contact_points.{'left' if i < 4 else 'right'}[{i % 4}]"
        # _add_point_dynamics (planner.py:721-744): dot(f) = f_dot, dot(p) = v
        add_linear_dynamics(3, name + ".f_dynamics", 15 * i + F, 15 * i + FD, lay.state_pt(lay.init, i, "f"))
        add_linear_dynamics(3, name + ".p_dynamics", 15 * i + P, 15 * i + V, lay.state_pt(lay.init, i, "p"))
        # _add_contact_point_feasibility (planner.py:634-719)
        for k in knots_all:
            nlp.subject_to(t_planar, _cat(lay.pt(k, i), lay.kt, tp_idx), f"{name}.p_planar_complementarity[{k}]")
        for k in knots_all:
            nlp.subject_to(t_dcc, _cat(lay.pt(k, i), lay.k_bs, lay.eps, tp_idx), f"{name}.p_dcc[{k}]")
        for k in knots_1:
            nlp.subject_to(t_height, _cat(lay.pt(k, i), tp_idx), f"{name}.p_height[{k}]")
        for k in knots_1:
            nlp.subject_to(t_normal, _cat(lay.pt(k, i), tp_idx), f"{name}.f_normal[{k}]")
        for k in knots_1:
            nlp.subject_to(t_friction, _cat(lay.pt(k, i), lay.mu, tp_idx), f"{name}.f_friction[{k}]")
        for k in knots_all:
            nlp.subject_to(t_ubounds, _cat(lay.pt(k, i), lay.rng(lay.max_u, 3)), f"{name}.u_v_bounds[{k}]")
        for k in knots_all:
            nlp.subject_to(t_fdbounds, _cat(lay.pt(k, i), lay.mass, lay.rng(lay.max_fd, 3)),
                           f"{name}.f_dot_bounds[{k}]")
        # _add_contact_kinematic_consistency (planner.py:590-632)
        for k in knots_1:
            nlp.subject_to(t_fk[frame], _cat(lay.pt(k, i), kin_idx(k), lay.desc(k, i)), f"{name}.p_kinematics_consistency[{k}]")
        # _add_contact_point_regularization (planner.py:855-895)
        for k in knots_1:
            nlp.minimize(t_swing, _cat(lay.pt(k, i), lay.ref(k, "swing_h"), tp_idx),
                         st.swing_foot_height_cost_multiplier, f"{name}.p_swing_height_regularization[{k}]")
        for k in knots_1:
            nlp.minimize(t_ureg, lay.pt(k, i), st.contact_velocity_control_cost_multiplier,
                         f"{name}.u_v_regularization[{k}]")
        for k in knots_1:
            nlp.minimize(t_fdreg, lay.pt(k, i), st.contact_force_control_cost_multiplier,
                         f"{name}.f_dot_regularization[{k}]")

    # ---------------------------------------------------------------- _add_robot_dynamics (:522-588)
    add_linear_dynamics(3, "base_position_dynamics", PB, VB, lay.state(lay.init, "pb"))
    add_linear_dynamics(4, "base_quaternion_dynamics", Q, QD, lay.state(lay.init, "q"))
    add_linear_dynamics(NJ, "joint_position_dynamics", S, SD, lay.state(lay.init, "s"))
    add_linear_dynamics(3, "com_dynamics", COM, H, lay.state(lay.init, "com"))

    # centroidal momentum dynamics: dot(h) = g + sum_i [f_i; (p_i - x) x f_i]
    def hdyn_inputs(tag):
        return dict(com=sx.syms("com" + tag, 3), p=[sx.syms(f"p{tag}{i}", 3) for i in range(NPT)],
                    f=[sx.syms(f"f{tag}{i}", 3) for i in range(NPT)])

    a, b = hdyn_inputs("a"), hdyn_inputs("b")
    grav = sx.syms("grav", 6)
    dt = sx.sym("dt")
    hk, hn = sx.syms("hk", 6), sx.syms("hn", 6)
    Fa = ex.centroidal_dynamics_with_point_forces(grav, a["com"], a["p"], a["f"])
    Fb = ex.centroidal_dynamics_with_point_forces(grav, b["com"], b["p"], b["f"])
    rows = [hn[i] - (hk[i] + 0.5 * dt * (Fa[i] + Fb[i])) for i in range(6)]
    t_hdyn = Template("centroidal_momentum_dynamics",
                      _l(hk, hn, a["com"], *a["p"], *a["f"], b["com"], *b["p"], *b["f"], grav, [dt]), rows)

    def hdyn_idx(k):
        return _cat(lay.x(k, COM, 3), *[lay.x(k, 15 * i + P, 3) for i in range(NPT)],
                    *[lay.x(k, 15 * i + F, 3) for i in range(NPT)])

    if not st.periodicity_constraint:  # planner.py:580-584
        nlp.subject_to(ic_template_var(6, "h_ic"), _cat(lay.x(0, H, 6), lay.rng(lay.h_init, 6)),
                       "centroidal_momentum_dynamics[0]{0}")
    for k in range(N - 1):
        nlp.subject_to(t_hdyn, _cat(lay.x(k, H, 6), lay.x(k + 1, H, 6), hdyn_idx(k), hdyn_idx(k + 1),
                                    lay.rng(lay.gravity, 6), lay.dt), f"centroidal_momentum_dynamics[{k + 1}]{{0}}")

    # ---------------------------------------------------------------- _add_kinematics_constraints (:266-425)
    t_unit = Template("unitary_quaternion", _l(q), [sx.sumsqr(q)], lb=[1.0], ub=[1.0])
    for k in knots_1:
        nlp.subject_to(t_unit, lay.x(k, Q, 4), f"unitary_quaternion[{k}]")

    comv = sx.syms("com", 3)
    com_kin = ex.center_of_mass_position_from_kinematics(model, pb, qn, s)
    t_com = Template("com_kinematics_consistency", _l(comv, pb, q, s), [comv[i] - com_kin[i] for i in range(3)])
    for k in knots_1:
        nlp.subject_to(t_com, _cat(lay.x(k, COM, 3), kin_idx(k)), f"com_kinematics_consistency[{k}]")

    hv = sx.syms("h", 6)
    h_kin = ex.centroidal_momentum_from_kinematics(model, pb, qn, s, vb, qd, sd)
    t_mom = Template("centroidal_momentum_kinematics_consistency", _l(hv, pb, q, s, vb, qd, sd, [mass]),
                     [hv[3 + i] - h_kin[3 + i] / mass for i in range(3)])
    for k in knots_all:
        nlp.subject_to(t_mom, _cat(lay.x(k, H, 6), kin_idx(k), lay.x(k, VB, 3), lay.x(k, QD, 4), lay.x(k, SD, NJ),
                                   lay.mass), f"centroidal_momentum_kinematics_consistency[{k}]")

    max_L = sx.sym("max_L")
    t_Lb = Template("angular_momentum_bounds", _l(hv, [mass, max_L]), [hv[3 + i] * mass for i in range(3)],
                    lb=[-max_L] * 3, ub=[max_L] * 3)
    for k in knots_all:
        nlp.subject_to(t_Lb, _cat(lay.x(k, H, 6), lay.mass, lay.max_L), f"angular_momentum_bounds[{k}]")

    min_h = sx.sym("min_com_h")
    t_comh = Template("minimum_com_height", _l(comv, [min_h], tpar), [terrain.height(comv)], lb=[min_h], ub=[INF])
    for k in knots_1:
        nlp.subject_to(t_comh, _cat(lay.x(k, COM, 3), lay.min_com_h, tp_idx), f"minimum_com_height[{k}]")

    min_d = sx.sym("min_feet_d")
    rel = ex.frames_relative_position(model, st.foot_frames[1], st.foot_frames[0], s)
    t_feet = Template("minimum_feet_distance", _l(s, [min_d]), [rel[1]], lb=[min_d], ub=[INF])
    for k in knots_1:
        nlp.subject_to(t_feet, _cat(lay.x(k, S, NJ), lay.min_feet_d), f"minimum_feet_distance[{k}]")

    lo, hi = sx.syms("lo", NJ), sx.syms("hi", NJ)
    t_sb = Template("joint_position_bounds", _l(s, lo, hi), list(s), lb=list(lo), ub=list(hi))
    for k in knots_1:
        nlp.subject_to(t_sb, _cat(lay.x(k, S, NJ), lay.rng(lay.min_s, NJ), lay.rng(lay.max_s, NJ)),
                       f"joint_position_bounds[{k}]")
    t_sdb = Template("joint_velocity_bounds", _l(sd, lo, hi), list(sd), lb=list(lo), ub=list(hi))
    for k in knots_all:
        nlp.subject_to(t_sdb, _cat(lay.x(k, SD, NJ), lay.rng(lay.min_sd, NJ), lay.rng(lay.max_sd, NJ)),
                       f"joint_velocity_bounds[{k}]")

    if st.final_state_constraint:
        # alphabetical leaf order of HumanoidState.to_list() (optimization_object.py:305-306):
        # com, contact_points.left[i].{descriptor, f, p}, ...right[i]..., kinematics.base.position,
        # .quaternion_xyzw, kinematics.joints.positions
        lhs, rhs = [lay.x(N - 1, COM, 3)], [lay.state(lay.final, "com")]
        for i in range(NPT):
            lhs += [lay.desc(N - 1, i), lay.x(N - 1, 15 * i + F, 3), lay.x(N - 1, 15 * i + P, 3)]
            rhs += [lay.state_pt(lay.final, i, "desc"), lay.state_pt(lay.final, i, "f"),
                    lay.state_pt(lay.final, i, "p")]
        lhs += [lay.x(N - 1, PB, 3), lay.x(N - 1, Q, 4), lay.x(N - 1, S, NJ)]
        rhs += [lay.state(lay.final, "pb"), lay.state(lay.final, "q"), lay.state(lay.final, "s")]
        nlp.subject_to(ic_template(105, "final_state_expression"), _cat(*lhs, *rhs), "final_state_expression")

    # ---------------------------------------------------------------- _add_kinematics_regularization (:427-520)
    cref = sx.syms("ref", 3)
    w = st.com_linear_velocity_cost_weights
    e = [hv[i] - cref[i] for i in range(3)]
    t_comvel = Template("com_velocity_error", _l(hv, cref), [_wquad(e, w)])
    for k in knots_all:
        nlp.minimize(t_comvel, _cat(lay.x(k, H, 6), lay.ref(k, "comv")), st.com_linear_velocity_cost_multiplier,
                     f"com_velocity_error[{k}]")

    qdes = sx.syms("qdes", 4)
    E = ex.rotation_error_from_kinematics(model, st.frame_quaternion_cost_frame, pb, qn, s, qdes)
    t_frame = Template("frame_quaternion_error", _l(pb, q, s, qdes), [sx.sq((E[0, 0] + E[1, 1] + E[2, 2]) - 3.0)])
    for k in knots_1:
        nlp.minimize(t_frame, _cat(kin_idx(k), lay.ref(k, "fq")), st.desired_frame_quaternion_cost_multiplier,
                     f"frame_quaternion_error[{k}]")

    t_bq = Template("base_quaternion_error", _l(q, qdes), [sx.sumsqr(ex.quaternion_xyzw_error(q, qdes))])
    for k in knots_1:
        nlp.minimize(t_bq, _cat(lay.x(k, Q, 4), lay.ref(k, "bq")), st.base_quaternion_cost_multiplier,
                     f"base_quaternion_error[{k}]")

    t_bqv = Template("base_quaternion_velocity_error", _l(qd, qdes), [sx.sumsqr([qd[i] - qdes[i] for i in range(4)])])
    for k in knots_all:
        nlp.minimize(t_bqv, _cat(lay.x(k, QD, 4), lay.ref(k, "bqv")), st.base_quaternion_velocity_cost_multiplier,
                     f"base_quaternion_velocity_error[{k}]")

    # joint regularisation, planner.py:506-520.  ``diag(w) * error`` is an ELEMENT-WISE product of
    # an n x n matrix with an n x 1 vector; CasADi repeats the vector horizontally [ext], giving
    # the n x n matrix M_ij = s_dot_i + delta_ij w_i e_i, so that
    # sumsqr(M) = sum_i [(n - 1) s_dot_i^2 + (s_dot_i + w_i e_i)^2]   (SURVEY.md A.11).
    jr = sx.syms("jr", NJ)
    wj = st.joint_regularization_cost_weights
    acc = sx.const(0.0)
    for i in range(NJ):
        for j in range(NJ):
            term = sd[i] + (float(wj[i]) * (s[i] - jr[i]) if i == j else 0.0)
            acc = acc + sx.sq(term)
    t_joint = Template("joint_positions_error", _l(s, sd, jr), [acc])
    for k in knots_1:
        nlp.minimize(t_joint, _cat(lay.x(k, S, NJ), lay.x(k, SD, NJ), lay.ref(k, "jr")),
                     st.joint_regularization_cost_multiplier, f"joint_positions_error[{k}]")

    # ---------------------------------------------------------------- _add_contact_centroids_expressions (:215-264)
    pts = [sx.syms(f"p{i}", 3) for i in range(NPT)]
    lc = ex.contact_points_centroid(pts[:4])
    rc = ex.contact_points_centroid(pts[4:])
    max_fh = sx.sym("max_feet_h")
    t_relh = Template("maximum_feet_relative_height", _l(*pts, [max_fh]), [lc[2] - rc[2]], lb=[-max_fh], ub=[max_fh])

    def pts_idx(k):
        return _cat(*[lay.x(k, 15 * i + P, 3) for i in range(NPT)])

    for k in knots_1:
        nlp.subject_to(t_relh, _cat(pts_idx(k), lay.max_feet_h), f"maximum_feet_relative_height[{k}]")
    cc, cw = sx.syms("cc", 3), sx.syms("cw", 3)
    ce = [cc[i] - 0.5 * (lc[i] + rc[i]) for i in range(3)]
    t_cent = Template("contacts_centroid_cost", _l(*pts, cc, cw), [_wquad(ce, cw)])
    for k in knots_1:
        nlp.minimize(t_cent, _cat(pts_idx(k), lay.ref(k, "cc"), lay.ref(k, "cw")),
                     st.contacts_centroid_cost_multiplier, f"contacts_centroid_cost[{k}]")

    # ---------------------------------------------------------------- _add_foot_regularization x2 (:746-853)
    fs = [sx.syms(f"f{i}", 3) for i in range(4)]
    alpha = sx.sym("alpha")
    ssum = [fs[0][c] + fs[1][c] + fs[2][c] + fs[3][c] for c in range(3)]
    t_ratio = []
    for i in range(4):
        t_ratio.append(Template(f"force_ratio{i}", _l(*fs, [alpha]),
                                [sx.sumsqr([fs[i][c] - alpha * ssum[c] for c in range(3)])]))
    p0, p1, p2 = sx.syms("pa", 3), sx.syms("pb_", 3), sx.syms("pc", 3)
    yaw = sx.sym("yaw")
    fwd = ex.contact_points_yaw_alignment_error(p0, p1, yaw)
    side = ex.contact_points_yaw_alignment_error(p1, p2, yaw + float(np.pi / 2))
    t_yaw = Template("yaw_regularization", _l(p0, p1, p2, [yaw]), [0.5 * (sx.sq(fwd) + sx.sq(side))])
    for foot in range(2):
        base = 4 * foot
        for i in range(4):
            for k in knots_1:
                nlp.minimize(t_ratio[i], _cat(*[lay.x(k, 15 * (base + j) + F, 3) for j in range(4)],
                                              lay.ref(k, "ratio_l" if foot == 0 else "ratio_r")[i]),
                             st.force_regularization_cost_multiplier,
                             f"system.contact_points.{'left' if foot == 0 else 'right'}[{i}].f_regularization[{k}]")
        for k in knots_1:
            nlp.minimize(t_yaw, _cat(lay.x(k, 15 * (base + st.yaw_bottom_right) + P, 3),
                                     lay.x(k, 15 * (base + st.yaw_top_right) + P, 3),
                                     lay.x(k, 15 * (base + st.yaw_top_left) + P, 3),
                                     lay.ref(k, "yaw_l" if foot == 0 else "yaw_r")),
                         st.foot_yaw_regularization_cost_multiplier,
                         f"{'left' if foot == 0 else 'right'}_yaw_regularization[{k}]")

    # ---------------------------------------------------------------- _add_periodicity_expression (:897-930)
    if st.periodicity_constraint:
        first, last = [], []
        for i in range(NPT):
            first += [lay.x(0, 15 * i + U, 3), lay.x(0, 15 * i + FD, 3)]
            last += [lay.x(N - 1, 15 * i + U, 3), lay.x(N - 1, 15 * i + FD, 3)]
        first += [lay.x(0, H, 6), lay.x(0, VB, 3), lay.x(0, QD, 4), lay.x(0, SD, NJ)]
        last += [lay.x(N - 1, H, 6), lay.x(N - 1, VB, 3), lay.x(N - 1, QD, 4), lay.x(N - 1, SD, NJ)]
        n = sum(len(a) for a in first)
        nlp.subject_to(ic_template_var(n, "periodicity_expression"), _cat(*first, *last), "periodicity_expression")

    return nlp, lay


def _l(*groups):
    out = []
    for g in groups:
        out.extend(list(g))
    return out


def _wquad(e, w):
    """e^T diag(w) e, accumulated left to right like a CasADi matrix product."""
    acc = sx.const(0.0)
    for i in range(len(e)):
        acc = acc + (e[i] * sx._wrap(w[i])) * e[i]
    return acc
