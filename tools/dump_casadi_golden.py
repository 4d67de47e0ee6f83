"""Dump the REFERENCE's own numbers for the golden fixtures: run this where `casadi`, `adam`, `liecasadi` and
`hippopt` are importable (they are not in the build container, which is why the oracle is "parity unpinned" for the
kinematic rows).

For every committed fixture tests/golden/kino_*.npz the script builds the same problem through the reference's own
code -- `hippopt.turnkey_planners.humanoid_kinodynamic.planner.Planner(settings)` (planner.py:27-176) on the synthetic
ergoCub URDF the tests use (hippopt_b200.robot_model.synthetic_ergocub_urdf); planar terrain for configs 3 / 4, and for
the stairs fixtures (config 5) the sum of two `SmoothTerrain.step` boxes of main_walking_on_stairs.py:18-28, one planner
per instance because the reference bakes the step dimensions into the graph where this library reads them from ten
parameters (those ten are stripped from p before the call).  (The toy fixtures toy_*.npz are pinned differently: the
reference's own test holds their known answers, tests/test_golden_cpu.py.)  It
takes the fixture's x / p / lam / sigma, evaluates the five nlpsol oracle functions CasADi generates (nlp_f,
nlp_grad_f, nlp_g, nlp_jac_g, nlp_hess_l) and writes tests/golden/casadi_<fixture>.npz with THE SAME KEYS
(x, p, lam, sigma, f, grad_f, g, jac, hess, jac_colind, jac_row, hess_colind, hess_row, lbg, ubg + the settings keys).
It also records what IPOPT really sees with the mains' plugin options {"expand": True, "detect_simple_bounds": True}:
  sb_g_rows          rows of g that stay general constraints (the others became lbx / ubx)
  sb_jac_colind/row  pattern of nlp_jac_g after the reduction, sb_hess_colind/row of nlp_hess_l
  sb_lbx / sb_ubx    the detected simple bounds
The parity tests pick casadi_*.npz up automatically (tests/test_gpu_parity.py::test_kino_golden,
tests/test_golden_cpu.py): the CUDA path and the oracle are then compared with CasADi's numbers directly.

The CasADi-facing core (`dump_nlp`) takes the modules as arguments, so that its plumbing is unit-tested here with a
recording stand-in (tests/test_dump_tool_cpu.py).
"""
from __future__ import annotations

import argparse
import glob
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def _ccs(sp):
    return np.asarray(sp.colind(), dtype=np.int64), np.asarray(sp.row(), dtype=np.int64)


def _nz(dm):
    """structural non-zeros of a DM in compressed-column order"""
    return np.asarray(dm.nonzeros() if hasattr(dm, "nonzeros") else dm.nz, dtype=np.float64).ravel()


def dump_nlp(cs, opti, x, p, lam, sigma, plugin_options=None) -> dict:
    """Evaluate CasADi's generated nlpsol oracle functions of the baked Opti problem on batches x (B, n_x),
    p (B, n_p), lam (B, m), sigma (B,).  Two solver objects are built: without plugin options (the full row set the
    kernels emit) and with the mains' {"expand", "detect_simple_bounds"} (what IPOPT is handed)."""
    x, p, lam, sigma = (np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in (x, p, lam, np.atleast_1d(sigma)))
    sigma = sigma.ravel()
    nlp = {"x": opti.x, "p": opti.p, "f": opti.f, "g": opti.g}
    full = cs.nlpsol("hb_dump_full", "ipopt", nlp, {})
    f_fun, gradf_fun = full.get_function("nlp_f"), full.get_function("nlp_grad_f")
    g_fun, jac_fun, hess_fun = full.get_function("nlp_g"), full.get_function("nlp_jac_g"), full.get_function("nlp_hess_l")
    out: dict = {"x": x, "p": p, "lam": lam, "sigma": sigma}
    jc, jr = _ccs(jac_fun.sparsity_out(1))
    hc, hr = _ccs(hess_fun.sparsity_out(0))
    out.update(jac_colind=jc, jac_row=jr, hess_colind=hc, hess_row=hr)
    B = x.shape[0]
    cols = {k: [] for k in ("f", "grad_f", "g", "jac", "hess", "lbg", "ubg")}
    for b in range(B):
        cols["f"].append(float(np.asarray(f_fun(x[b], p[b])).ravel()[0]))
        cols["grad_f"].append(np.asarray(gradf_fun(x[b], p[b])[1], dtype=np.float64).ravel())
        cols["g"].append(np.asarray(g_fun(x[b], p[b]), dtype=np.float64).ravel())
        cols["jac"].append(_nz(jac_fun(x[b], p[b])[1]))
        cols["hess"].append(_nz(hess_fun(x[b], p[b], sigma[b], lam[b])))
        bounds = cs.Function("hb_bounds", [opti.p], [opti.lbg, opti.ubg])(p[b])
        cols["lbg"].append(np.asarray(bounds[0], dtype=np.float64).ravel())
        cols["ubg"].append(np.asarray(bounds[1], dtype=np.float64).ravel())
    out.update({k: np.asarray(v) for k, v in cols.items()})
    # what IPOPT sees with the mains' plugin options
    opts = dict(plugin_options if plugin_options is not None else {"expand": True, "detect_simple_bounds": True})
    red = cs.nlpsol("hb_dump_reduced", "ipopt", nlp, opts)
    rj, rh = red.get_function("nlp_jac_g"), red.get_function("nlp_hess_l")
    sjc, sjr = _ccs(rj.sparsity_out(1))
    shc, shr = _ccs(rh.sparsity_out(0))
    out.update(sb_jac_colind=sjc, sb_jac_row=sjr, sb_hess_colind=shc, sb_hess_row=shr,
               sb_m=np.int64(rj.sparsity_out(1).size1()))
    if hasattr(red, "simple_bounds"):  # stand-in / future CasADi accessor; the real one exposes them through stats
        rows, lbx, ubx = red.simple_bounds(p[0])
        out.update(sb_g_rows=np.asarray(rows, dtype=np.int64), sb_lbx=np.asarray(lbx), sb_ubx=np.asarray(ubx))
    return out


def merge_instances(parts: list) -> dict:
    """Per-instance dumps (one planner each) -> one dump: batched keys are stacked, the patterns must agree."""
    out = dict(parts[0])
    if len(parts) == 1:
        return out
    batched = ("x", "p", "lam", "sigma", "f", "grad_f", "g", "jac", "hess", "lbg", "ubg")
    for k, v in parts[0].items():
        if k in batched:
            out[k] = np.concatenate([np.asarray(q[k]) for q in parts], axis=0)
        else:
            for q in parts[1:]:
                if not np.array_equal(np.asarray(q[k]), np.asarray(v)):
                    raise ValueError(f"instances disagree on '{k}': the sparsity pattern depends on the terrain")
    return out


# ------------------------------------------------------------------------------------------------------ planners
def build_kinodynamic_opti(fixture: dict, terrain_params=None):
    """The reference's planner for one fixture -> (casadi module, baked Opti).  Needs the reference environment.
    terrain_params: (l, w, height, ox, oy) x 2 of the two smooth steps (one instance of a stairs fixture)."""
    import casadi as cs
    import hippopt as hp
    import hippopt.robot_planning as hp_rp
    from hippopt.turnkey_planners.humanoid_kinodynamic import planner as kp
    from hippopt.turnkey_planners.humanoid_kinodynamic import settings as ks

    from hippopt_b200.robot_model import ERGOCUB_JOINTS, synthetic_ergocub_urdf
    from hippopt_b200.workloads import FOOT_CORNERS

    N = int(fixture["horizon"])
    st = ks.Settings()
    urdf = tempfile.NamedTemporaryFile("w", suffix=".urdf", delete=False)
    urdf.write(synthetic_ergocub_urdf())
    urdf.close()
    st.robot_urdf = urdf.name
    st.joints_name_list = list(ERGOCUB_JOINTS)
    st.root_link = "root_link"
    st.horizon_length = N
    st.time_step = 0.1
    st.contact_points = hp_rp.FeetContactPointDescriptors()
    st.contact_points.left = [hp_rp.ContactPointDescriptor(foot_frame="l_sole", position_in_foot_frame=np.array(c))
                              for c in FOOT_CORNERS]
    st.contact_points.right = [hp_rp.ContactPointDescriptor(foot_frame="r_sole", position_in_foot_frame=np.array(c))
                               for c in FOOT_CORNERS]
    st.integrator = hp.ImplicitTrapezoid
    st.terrain = hp_rp.PlanarTerrain()
    if bool(fixture.get("smooth", False)):
        if terrain_params is None:
            raise ValueError("a stairs fixture needs the terrain parameters of one instance")
        tp = np.asarray(terrain_params, dtype=np.float64).reshape(2, 5)
        steps = [hp_rp.SmoothTerrain.step(length=float(t[0]), width=float(t[1]), height=float(t[2]),
                                          position=np.array([t[3], t[4], 0.0])) for t in tp]
        st.terrain = steps[0] + steps[1]  # TerrainSum (utilities/terrain_sum.py)
    st.desired_frame_quaternion_cost_frame_name = "chest"
    st.final_state_expression_type = hp.ExpressionType.subject_to if bool(fixture["final"]) else hp.ExpressionType.skip
    st.periodicity_expression_type = (hp.ExpressionType.subject_to if bool(fixture["periodicity"])
                                      else hp.ExpressionType.skip)
    # cost multipliers / weights: the defaults of hippopt_b200.kino_layout.KinoSettings are those of
    # main_single_step_flat_ground.py:54-104, set them the same way here
    from hippopt_b200.kino_layout import KinoSettings

    d = KinoSettings()
    for name in ("contacts_centroid_cost_multiplier", "com_linear_velocity_cost_multiplier",
                 "desired_frame_quaternion_cost_multiplier", "base_quaternion_cost_multiplier",
                 "base_quaternion_velocity_cost_multiplier", "joint_regularization_cost_multiplier",
                 "force_regularization_cost_multiplier", "foot_yaw_regularization_cost_multiplier",
                 "swing_foot_height_cost_multiplier", "contact_velocity_control_cost_multiplier",
                 "contact_force_control_cost_multiplier"):
        setattr(st, name, getattr(d, name))
    st.com_linear_velocity_cost_weights = list(d.com_linear_velocity_cost_weights)
    st.joint_regularization_cost_weights = np.asarray(d.joint_regularization_cost_weights)
    st.casadi_opti_options = {}
    st.casadi_solver_options = {}
    planner = kp.Planner(settings=st)
    solver = planner.optimization_solver
    solver._cost = solver._cost if solver._cost is not None else cs.MX(0)
    solver._solver.minimize(solver._cost)
    return cs, solver._solver


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--out", default=GOLD)
    ap.add_argument("--fixtures", default="kino_*.npz")
    args = ap.parse_args(argv)
    try:
        import casadi  # noqa: F401
        import hippopt  # noqa: F401
    except ImportError as err:
        print(f"this script needs the reference environment (casadi, adam, liecasadi, hippopt): {err}")
        return 2
    for path in sorted(glob.glob(os.path.join(GOLD, args.fixtures))):
        fx = dict(np.load(path))
        smooth = bool(fx.get("smooth", False))
        n_tp = 10 if smooth else 0
        parts = []
        for b in (range(fx["x"].shape[0]) if smooth else [None]):
            cs, opti = build_kinodynamic_opti(fx, fx["p"][b, -n_tp:] if smooth else None)
            if (opti.nx, opti.np + n_tp, opti.ng) != (fx["x"].shape[1], fx["p"].shape[1], fx["g"].shape[1]):
                print(f"{os.path.basename(path)}: dimensions differ (Opti {opti.nx}/{opti.np}/{opti.ng}) -- layout "
                      "mismatch, SURVEY.md Appendix B needs revisiting")
                return 1
            sel = slice(None) if b is None else slice(b, b + 1)
            p_ref = fx["p"][sel, :fx["p"].shape[1] - n_tp]
            parts.append(dump_nlp(cs, opti, fx["x"][sel], p_ref, fx["lam"][sel], fx["sigma"][sel]))
        out = merge_instances(parts)
        out["p"] = fx["p"]  # the fixture's own parameter vector (with the terrain block), so that the keys line up
        for k in ("horizon", "final", "periodicity", "smooth", "dt"):
            if k in fx:
                out[k] = fx[k]
        dst = os.path.join(args.out, "casadi_" + os.path.basename(path))
        np.savez_compressed(dst, **out)
        print("wrote", dst)
    return 0


if __name__ == "__main__":
    sys.exit(main())
