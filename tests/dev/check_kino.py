"""Developer check: CUDA kinodynamic evaluator vs the CPU oracle on a small problem (run on a GPU box)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hippopt_b200.evaluator import ALL, KinoEvaluator, probe_fp64_tflops  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch  # noqa: E402
from oracle import kinodynamic as kd  # noqa: E402


def relerr(a, b):
    return float(np.max(np.abs(a - b) / (1e-12 + np.maximum(1.0, np.abs(b)))))


def main():
    model = synthetic_ergocub()
    for (N, fin, per, noise) in ((4, False, False, 0.05), (5, True, True, 0.2)):
        st = KinoSettings(horizon=N, final_state_constraint=fin, periodicity_constraint=per)
        ev = KinoEvaluator(model, st)
        lay = ev.layout
        B = 3
        x, p, lam, sigma = kino_batch(lay, model, B, seed=11, noise=noise)
        sigma = np.array([1.0, 0.5, 2.0])
        nlp, _ = kd.build(model, kd.Settings(horizon=N, final_state_constraint=fin, periodicity_constraint=per))
        d = torch.device("cuda:0")
        out = ev.eval(ALL, torch.tensor(x, device=d), torch.tensor(p, device=d), torch.tensor(lam, device=d),
                      torch.tensor(sigma, device=d))
        torch.cuda.synchronize()
        res = {k: v.cpu().numpy() for k, v in out.items()}
        ref = {
            "f": nlp.eval_f(x, p), "grad_f": nlp.eval_grad_f(x, p), "g": nlp.eval_g(x, p),
            "jac": nlp.eval_jac(x, p), "hess": nlp.eval_hess(x, p, lam, sigma),
        }
        print(f"N={N} final={fin} per={per}")
        for k in ("f", "g", "grad_f", "jac", "hess"):
            e = relerr(res[k], ref[k])
            print(f"  {k:7s} max rel err {e:.3e}  (max |ref| {np.abs(ref[k]).max():.3e})")
            if e > 1e-9:
                bad = np.argwhere(np.abs(res[k] - ref[k]) / np.maximum(1.0, np.abs(ref[k])) > 1e-9)
                print("    first mismatches:", bad[:12].tolist())
                for idx in bad[:12]:
                    print("     ", idx.tolist(), res[k][tuple(idx)], ref[k][tuple(idx)])
    print("fp64 probe TFLOP/s:", probe_fp64_tflops())
    # quick timing at config-3 size
    st = KinoSettings(horizon=30)
    ev = KinoEvaluator(model, st)
    B = 1024
    x, p, lam, sigma = kino_batch(ev.layout, model, B, seed=2)
    d = torch.device("cuda:0")
    X, Pm, Lm, Sg = (torch.tensor(a, device=d) for a in (x, p, lam, sigma))
    for mask, name in ((ALL, "f+g+grad+jac+hess"), (ALL & ~16, "f+g+grad+jac"), (1 | 4, "f+g")):
        for _ in range(3):
            ev.eval(mask, X, Pm, Lm, Sg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            ev.eval(mask, X, Pm, Lm, Sg)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"  {name:20s} {dt * 1e3:8.3f} ms  -> {B * 30 / dt:.3e} knot-evals/s")


if __name__ == "__main__":
    main()
