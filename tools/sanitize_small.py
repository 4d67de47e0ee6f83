"""Small workloads for compute-sanitizer (memcheck / racecheck / initcheck): every kernel once, tiny batch."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import ALL, F, G, JAC_G, HostPipeline, KinoEvaluator, PoseEvaluator, ToyEvaluator  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.kkt import lu_factor, lu_solve  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch, pose_batch  # noqa: E402

d = torch.device("cuda:0")
model = synthetic_ergocub()
for st in (KinoSettings(horizon=3, final_state_constraint=True, periodicity_constraint=True),
           KinoSettings(horizon=2, terrain="smooth_steps", n_terrain_params=10)):
    ev = KinoEvaluator(model, st)
    x, p, lam, sigma = kino_batch(ev.layout, model, 3, seed=1, noise=0.1)
    t = [torch.tensor(a, device=d) for a in (x, p, lam, sigma)]
    for mask in (ALL, F | G, JAC_G):
        ev.eval(mask, *t)
    torch.cuda.synchronize()
pev = PoseEvaluator(model)
t = [torch.tensor(a, device=d) for a in pose_batch(pev.layout, model, 5, seed=1)]
pev.eval(ALL, *t)
tev = ToyEvaluator(6, "trapezoid")
tev.eval(ALL, torch.zeros((2, tev.n_x), dtype=torch.float64, device=d), torch.ones((2, tev.n_p), dtype=torch.float64, device=d),
         torch.ones((2, tev.m), dtype=torch.float64, device=d), torch.ones(2, dtype=torch.float64, device=d))
A = torch.randn(3, 37, 37, dtype=torch.float64, device=d)
Fm, piv, info = lu_factor(A.clone())
lu_solve(Fm, piv, torch.randn(3, 37, 5, dtype=torch.float64, device=d))
A = torch.randn(2, 337, 337, dtype=torch.float64, device=d)  # the KKT stage-block size: 22 panels, both register variants
Fm, piv, info = lu_factor(A.clone())
lu_solve(Fm, piv, torch.randn(2, 337, 40, dtype=torch.float64, device=d))
A = torch.randn(1, 700, 700, dtype=torch.float64, device=d)  # one CTA per SM in the staged substitution (fewer columns per CTA)
Fm, piv, info = lu_factor(A.clone())
lu_solve(Fm, piv, torch.randn(1, 700, 19, dtype=torch.float64, device=d))
# interpolation kernel: a plan with swings, per-instance phases, both outputs, a warp with fewer than three items
from hippopt_b200.initial_guess import periodic_step_phases  # noqa: E402
from hippopt_b200.interpolators import humanoid_state_interpolator  # noqa: E402

B = 7
ph = periodic_step_phases(np.linspace(0.1, 0.3, B), 1.0)
s0, s1 = torch.randn(B, 105, dtype=torch.float64, device=d), torch.randn(B, 105, dtype=torch.float64, device=d)
for q in (s0, s1):
    q[:, 75:79] /= q[:, 75:79].norm(dim=1, keepdim=True)
xo = torch.zeros((B, 189 * 12 + 6), dtype=torch.float64, device=d)
humanoid_state_interpolator(s0, s1, ph, 10, 0.1, x_out=xo, knot0=1)
torch.cuda.synchronize()
# round 2: named cost values, deterministic sparse products, the CasADi external-function ABI, one L-BFGS solver iteration
import ctypes  # noqa: E402

from hippopt_b200 import _capi  # noqa: E402
from hippopt_b200.ipsolver import BatchedInteriorPoint, SparseOps  # noqa: E402

for st in (KinoSettings(horizon=3), KinoSettings(horizon=2, terrain="smooth_steps", n_terrain_params=10)):
    ev = KinoEvaluator(model, st)
    x, p, lam, sigma = kino_batch(ev.layout, model, 3, seed=1, noise=0.1)
    X, P, L, S = (torch.tensor(a, device=d) for a in (x, p, lam, sigma))
    ev.cost_terms(X, P)
    out = ev.eval(ALL, X, P, L, S)
    ops = SparseOps(ev.n_x, ev.m, ev.jac_sparsity(), ev.hess_sparsity(), d)
    ops.J_mul(out["jac"], X)
    ops.Jt_mul(out["jac"], L)
    ops.W_quad(out["hess"], X)
    torch.cuda.synchronize()
ev = KinoEvaluator(model, KinoSettings(horizon=3))
x, p, lam, sigma = kino_batch(ev.layout, model, 1, seed=2, noise=0.1)
lib = _capi.lib()
assert lib.hb_external_bind(ev._h) == 0
dp = ctypes.POINTER(ctypes.c_double)
gj = [np.zeros(ev.m), np.zeros(ev.nnz_j)]
arg = (dp * 2)(x[0].ctypes.data_as(dp), p[0].ctypes.data_as(dp))
res = (dp * 2)(gj[0].ctypes.data_as(dp), gj[1].ctypes.data_as(dp))
lib.hb_nlp_jac_g.restype = ctypes.c_int
assert lib.hb_nlp_jac_g(arg, res, None, None, 0) == 0
hs = np.zeros(ev.nnz_h)
arg4 = (dp * 4)(x[0].ctypes.data_as(dp), p[0].ctypes.data_as(dp), sigma.ctypes.data_as(dp), lam[0].ctypes.data_as(dp))
res1 = (dp * 1)(hs.ctypes.data_as(dp))
lib.hb_nlp_hess_l.restype = ctypes.c_int
assert lib.hb_nlp_hess_l(arg4, res1, None, None, 0) == 0
lib.hb_external_bind(None)
lb, ub = ev.bounds(p)
try:
    BatchedInteriorPoint(ev, kkt="stage", delta_c=1e-9, ipopt_options={"hessian_approximation": "limited-memory", "max_iter": 3,
                                                                       "tol": 1e-3}).solve(torch.tensor(x, device=d), torch.tensor(p, device=d), lb, ub)
except Exception as exc:  # noqa: BLE001 -- three iterations do not converge: OptiFailure is the expected outcome
    print("solver:", type(exc).__name__)
torch.cuda.synchronize()
print("sanitize workload done")
