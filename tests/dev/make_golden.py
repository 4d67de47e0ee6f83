"""Generate the golden fixtures under tests/golden/ from the CPU oracle (run in this container:
`python tests/dev/make_golden.py`).  The oracle itself is pinned by tests/test_oracle_*.py; these files
freeze its outputs so that the GPU parity tests do not depend on rebuilding the graphs, and so that a
later CasADi dump can replace them file-for-file (same keys).  Lives under tests/ because it imports
oracle/ (test infrastructure only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hippopt_b200.kino_layout import KinoLayout, KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch  # noqa: E402
from oracle import kinodynamic as kd  # noqa: E402
from oracle import toy  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def kino(name, N, fin, per, B, seed, noise, smooth=False):
    from oracle import expressions as ex

    model = synthetic_ergocub()
    lay = KinoLayout(model, KinoSettings(horizon=N, final_state_constraint=fin, periodicity_constraint=per,
                                         terrain="smooth_steps" if smooth else "planar",
                                         n_terrain_params=10 if smooth else 0))
    x, p, lam, sigma = kino_batch(lay, model, B, seed=seed, noise=noise)
    sigma = np.linspace(0.5, 1.5, B)
    extra = dict(terrain=ex.TwoSmoothSteps(), terrain_params=10) if smooth else {}
    nlp, _ = kd.build(model, kd.Settings(horizon=N, final_state_constraint=fin, periodicity_constraint=per, **extra))
    jc, jr, _, _ = nlp.jac_structure()
    hc, hr, _, _ = nlp.hess_structure()
    lb, ub = nlp.eval_bounds(p)
    np.savez_compressed(
        os.path.join(OUT, name), horizon=N, final=fin, periodicity=per, smooth=smooth, x=x, p=p, lam=lam, sigma=sigma,
        f=nlp.eval_f(x, p), grad_f=nlp.eval_grad_f(x, p), g=nlp.eval_g(x, p), jac=nlp.eval_jac(x, p),
        hess=nlp.eval_hess(x, p, lam, sigma), jac_colind=jc, jac_row=jr, hess_colind=hc, hess_row=hr, lbg=lb, ubg=ub)


def toy_case(name, N, integrator, dt, B, seed):
    rng = np.random.default_rng(seed)
    nlp = toy.build(N, integrator, dt)
    x = rng.normal(size=(B, nlp.n_x))
    p = np.stack([rng.uniform(-12, -8, B), rng.uniform(0.5, 1.5, B), rng.uniform(-1, 1, B)], axis=1)
    lam = rng.normal(size=(B, nlp.m))
    sigma = rng.uniform(0.5, 2.0, B)
    jc, jr, _, _ = nlp.jac_structure()
    hc, hr, _, _ = nlp.hess_structure()
    lb, ub = nlp.eval_bounds(p)
    np.savez_compressed(
        os.path.join(OUT, name), horizon=N, dt=dt, x=x, p=p, lam=lam, sigma=sigma, f=nlp.eval_f(x, p),
        grad_f=nlp.eval_grad_f(x, p), g=nlp.eval_g(x, p), jac=nlp.eval_jac(x, p), hess=nlp.eval_hess(x, p, lam, sigma),
        jac_colind=jc, jac_row=jr, hess_colind=hc, hess_row=hr, lbg=lb, ubg=ub)


def interp_case(name, B, N, dt, seed):
    """humanoid_state_interpolator on the timings of main_periodic_step.py:367-412 with random per-instance
    foot transforms: the two halves of the guess (:433-451).  None times are stored as NaN."""
    from oracle import interpolators as oi

    rng = np.random.default_rng(seed)

    def quat():
        q = rng.normal(size=4)
        return q / np.linalg.norm(q)

    T = N * dt
    times = {"left": [(None, T / 6.0), (T / 3.0, None)], "right": [(None, T * 2.0 / 3.0), (T * 5.0 / 6.0, None)]}
    desc = [[0.08, 0.03, 0.0], [0.08, -0.03, 0.0], [-0.08, -0.03, 0.0], [-0.08, 0.03, 0.0]]
    feet = {side: [[{"position": rng.normal(size=3), "quaternion": quat(), "force": rng.normal(size=3),
                     "mid_position": rng.normal(size=3) if i == 0 else None, "mid_quaternion": quat() if i == 0 else None,
                     "activation_time": a, "deactivation_time": d} for i, (a, d) in enumerate(times[side])]
                   for _ in range(B)] for side in ("left", "right")}

    def state():
        return {"p": rng.normal(size=(8, 3)), "f": rng.normal(size=(8, 3)), "base_position": rng.normal(size=3),
                "base_quaternion": quat(), "joints": rng.uniform(-1, 1, 23), "com": rng.normal(size=3)}

    keys = [[state() for _ in range(B)] for _ in range(3)]
    half = N // 2
    out = {"n_points": N, "dt": dt, "descriptor": np.asarray(desc)}
    for side in ("left", "right"):
        for fld in ("position", "quaternion", "force"):
            out[f"{side}_{fld}"] = np.stack([[ph[fld] for ph in inst] for inst in feet[side]])
        out[f"{side}_mid_position"] = np.stack([inst[0]["mid_position"] for inst in feet[side]])
        out[f"{side}_mid_quaternion"] = np.stack([inst[0]["mid_quaternion"] for inst in feet[side]])
        out[f"{side}_times"] = np.asarray([[np.nan if v is None else v for v in t] for t in times[side]])
    for h, (k0, k1, pts, t0) in enumerate(((0, 1, half, 0.0), (1, 2, N - half, half * dt))):
        res = [oi.humanoid_state_interpolator(keys[k0][b], keys[k1][b], (feet["left"][b], feet["right"][b]), (desc, desc),
                                              pts, dt, t0) for b in range(B)]
        out[f"states_{h}"] = np.stack([np.stack([oi.state_block(s, (desc, desc)) for s in r]) for r in res])
        out[f"t0_{h}"] = t0
    for k in range(3):
        out[f"key_{k}"] = np.stack([oi.state_block(s, (desc, desc)) for s in keys[k]])
    np.savez_compressed(os.path.join(OUT, name), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["interp"]:  # only the interpolator fixture (the others stay byte-identical)
        interp_case("interp_periodic_step.npz", 3, 30, 0.1, 41)
        raise SystemExit(0)
    kino("kino_n3_flat.npz", 3, False, False, 2, 21, 0.1)
    kino("kino_n4_periodic.npz", 4, True, True, 2, 22, 0.3)
    kino("kino_n3_stairs.npz", 3, True, False, 2, 23, 0.03, smooth=True)
    toy_case("toy_n6_euler.npz", 6, "euler", 0.01, 3, 31)
    toy_case("toy_n7_trapezoid.npz", 7, "trapezoid", 0.05, 3, 32)
    interp_case("interp_periodic_step.npz", 3, 30, 0.1, 41)
    print(sorted(os.listdir(OUT)))
