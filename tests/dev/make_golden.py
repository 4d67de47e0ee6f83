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


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    kino("kino_n3_flat.npz", 3, False, False, 2, 21, 0.1)
    kino("kino_n4_periodic.npz", 4, True, True, 2, 22, 0.3)
    kino("kino_n3_stairs.npz", 3, True, False, 2, 23, 0.03, smooth=True)
    toy_case("toy_n6_euler.npz", 6, "euler", 0.01, 3, 31)
    toy_case("toy_n7_trapezoid.npz", 7, "trapezoid", 0.05, 3, 32)
    print(sorted(os.listdir(OUT)))
