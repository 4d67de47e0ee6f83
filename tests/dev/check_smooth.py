"""Developer check: smooth-terrain (config 5) CUDA path vs the CPU oracle."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import ALL, KinoEvaluator  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch  # noqa: E402
from oracle import expressions as ex  # noqa: E402
from oracle import kinodynamic as kd  # noqa: E402

model = synthetic_ergocub()
for (N, fin, noise) in ((3, True, 0.05), (4, False, 0.15)):
    st = KinoSettings(horizon=N, terrain="smooth_steps", n_terrain_params=10, final_state_constraint=fin)
    ev = KinoEvaluator(model, st)
    x, p, lam, sigma = kino_batch(ev.layout, model, 3, seed=17, noise=noise)
    sigma = np.array([1.0, 0.3, 2.0])
    nlp, _ = kd.build(model, kd.Settings(horizon=N, terrain=ex.TwoSmoothSteps(), terrain_params=10,
                                         final_state_constraint=fin))
    d = torch.device("cuda:0")
    out = ev.eval(ALL, *(torch.tensor(a, device=d) for a in (x, p, lam, sigma)))
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    ref = {"f": nlp.eval_f(x, p), "grad_f": nlp.eval_grad_f(x, p), "g": nlp.eval_g(x, p),
           "jac": nlp.eval_jac(x, p), "hess": nlp.eval_hess(x, p, lam, sigma)}
    print(f"N={N} final={fin}")
    for k in ("f", "g", "grad_f", "jac", "hess"):
        err = np.abs(res[k] - ref[k]) / np.maximum(1.0, np.abs(ref[k]))
        print(f"  {k:7s} max rel err {err.max():.3e}  (max |ref| {np.abs(ref[k]).max():.3e})")
        if err.max() > 1e-9:
            bad = np.argwhere(err > 1e-9)
            lay = ev.layout
            for idx in bad[:10]:
                extra = ""
                if k == "jac":
                    extra = f" row={nlp.row_names[lay.jac_row[idx[1]]]} col={lay.jac_col[idx[1]] % 189}"
                if k == "hess":
                    extra = f" ({lay.hess_row[idx[1]] % 189},{lay.hess_col[idx[1]] % 189}) knot {lay.hess_col[idx[1]] // 189}"
                if k == "g":
                    extra = f" {nlp.row_names[idx[1]]}"
                print("     ", idx.tolist(), res[k][tuple(idx)], ref[k][tuple(idx)], extra)
st = KinoSettings(horizon=50, terrain="smooth_steps", n_terrain_params=10, final_state_constraint=True)
ev = KinoEvaluator(model, st)
B = 512
x, p, lam, sigma = kino_batch(ev.layout, model, B, seed=4)
d = torch.device("cuda:0")
X, P, L, S = (torch.tensor(a, device=d) for a in (x, p, lam, sigma))
for _ in range(3):
    ev.eval(ALL, X, P, L, S)
torch.cuda.synchronize()
ev.profile(True)
t0 = time.perf_counter()
for _ in range(10):
    ev.eval(ALL, X, P, L, S)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
ms, n = ev.profile_read()
print(f"config 5 (N=50, B={B}): {dt * 1e3:.3f} ms -> {B * 50 / dt:.3e} knot-evals/s", {k: v / n for k, v in ms.items()})
