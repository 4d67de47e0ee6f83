"""Developer check of the stage-wise KKT solve (row f2) on evaluated kinodynamic values:
   1. N = 4: against a dense solve of the same matrix;
   2. N = 30, B instances: residual of the full system through sparse products, and Newton-system solves / s;
   3. (-s) a batched interior-point solve of a short-horizon kinodynamic OCP with the stage backend."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import ALL, KinoEvaluator  # noqa: E402
from hippopt_b200.ipsolver import BatchedInteriorPoint, SparseOps  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.kkt import StageKKT  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch  # noqa: E402

d = torch.device("cuda:0")
model = synthetic_ergocub()


def system(N, B, seed=3):
    ev = KinoEvaluator(model, KinoSettings(horizon=N))
    x, p, lam, sigma = kino_batch(ev.layout, model, B, seed=seed, noise=0.05)
    lb, ub = ev.bounds(p)
    kkt, eq, ine = StageKKT.for_evaluator(ev, lb[0], ub[0], device=d)
    t = [torch.tensor(a, device=d) for a in (x, p, lam, np.ones(B))]
    out = ev.eval(ALL, *t)
    torch.cuda.synchronize()
    g = torch.Generator(device="cpu").manual_seed(seed)
    sig = (torch.rand((B, len(ine)), generator=g, dtype=torch.float64) * 10.0).to(d)
    delta = torch.full((B,), 1e-2, dtype=torch.float64, device=d)
    rx = torch.randn((B, ev.n_x), generator=g, dtype=torch.float64).to(d)
    rE = torch.randn((B, len(eq)), generator=g, dtype=torch.float64).to(d)
    return ev, kkt, eq, ine, out, sig, delta, rx, rE


# ---- 1. dense comparison
ev, kkt, eq, ine, out, sig, delta, rx, rE = system(4, 4)
dc = 1e-9
dx, dl = kkt.solve(out["hess"], out["jac"], sig, delta, dc, rx, rE)
jc, jr = ev.jac_sparsity()
hc, hr = ev.hess_sparsity()
K = kkt.dense_matrix(out["hess"], out["jac"], sig, delta, dc, eq, ine, jc, jr, hc, hr)
ref = torch.linalg.solve(K, torch.cat([rx, rE], dim=1))
u = torch.cat([dx, dl], dim=1)
print(f"N=4: stage vs dense solve, max rel diff {((u - ref).abs().amax(1) / ref.abs().amax(1)).max().item():.2e}; "
      f"cond(K) ~ {torch.linalg.cond(K[0]).item():.1e}")

# ---- 2. full size
B = int(sys.argv[sys.argv.index("-b") + 1]) if "-b" in sys.argv else 256
ev, kkt, eq, ine, out, sig, delta, rx, rE = system(30, B)
ops = SparseOps(ev.n_x, ev.m, ev.jac_sparsity(), ev.hess_sparsity(), d)
iE, iI = torch.as_tensor(eq, device=d), torch.as_tensor(ine, device=d)
lb_, ub_ = ev.bounds(kino_batch(ev.layout, model, 1, seed=3, noise=0.05)[1])
kkt_torch, _, _ = StageKKT.for_evaluator(ev, lb_[0], ub_[0], device=d, linalg="torch")
for name, solver in (("torch.linalg (cuSOLVER)", kkt_torch), ("hb_lu kernels", kkt)):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dx, dl = solver.solve(out["hess"], out["jac"], sig, delta, dc, rx, rE)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"  {name:26s} {dt * 1e3:7.1f} ms per batched Newton-system solve = {B / dt:.0f} KKT solves/s")
hv, jv = out["hess"], out["jac"]
Jdx = ops.J_mul(jv, dx)
lam_full = torch.zeros((B, ev.m), dtype=torch.float64, device=d)
lam_full[:, iE] = dl
lam_full[:, iI] = sig * Jdx[:, iI]
# W dx through the quadratic form's gradient: (W dx) = sum over entries
Wdx = torch.zeros_like(dx)
Wdx.index_add_(1, ops.hr, hv * dx[:, ops.hc])
off = ops.hr != ops.hc
Wdx.index_add_(1, ops.hc[off], hv[:, off] * dx[:, ops.hr[off]])
r_x = Wdx + delta[:, None] * dx + ops.Jt_mul(jv, lam_full) - rx
r_E = Jdx[:, iE] - dc * dl - rE
scale = max(1.0, float(dx.abs().max()), float(dl.abs().max()))
print(f"N=30, B={B}: block size {kkt.nb}, residual max |r_x| {r_x.abs().max().item():.2e} |r_E| {r_E.abs().max().item():.2e} "
      f"(solution scale {scale:.1e}); {dt * 1e3:.1f} ms per batched Newton-system solve = {B / dt:.0f} KKT solves/s")

# ---- 3. interior-point solve with the stage backend
if "-s" in sys.argv:
    N, B = 6, 8
    ev = KinoEvaluator(model, KinoSettings(horizon=N))
    x, p, lam, sigma = kino_batch(ev.layout, model, B, seed=5, noise=0.0)
    lb, ub = ev.bounds(p)
    sol = BatchedInteriorPoint(ev, tol=1e-6, max_iter=int(sys.argv[sys.argv.index("-s") + 1]), verbose=True, kkt="stage",
                               delta_c=1e-9)
    try:
        res = sol.solve(torch.tensor(x, device=d), torch.tensor(p, device=d), lb, ub)
        print(f"kino N={N}: {int(res.success.sum())}/{B} converged, iterations {res.iterations.tolist()}, "
              f"kkt error {res.kkt_error.tolist()}")
    except Exception as e:  # OptiFailure: report, do not hide
        print("kino solve:", e)
