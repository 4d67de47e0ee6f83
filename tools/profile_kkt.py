"""Where a stage-wise KKT solve (hippopt_b200.kkt.StageKKT.solve, config 3, B instances) spends its GPU time:
torch.profiler totals per kernel.   usage: profile_kkt.py [-b BATCH]"""
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from hippopt_b200.evaluator import ALL, KinoEvaluator  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.kkt import StageKKT  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch  # noqa: E402

d = torch.device("cuda:0")
B = int(sys.argv[sys.argv.index("-b") + 1]) if "-b" in sys.argv else 296
R = int(sys.argv[sys.argv.index("-r") + 1]) if "-r" in sys.argv else 1  # right-hand sides (13: limited-memory mode)
model = synthetic_ergocub()
ev = KinoEvaluator(model, KinoSettings(horizon=30))
x, p, lam, sigma = kino_batch(ev.layout, model, B, seed=3, noise=0.05)
lb, ub = ev.bounds(p)
kkt, eq, ine = StageKKT.for_evaluator(ev, lb[0], ub[0], device=d)
out = ev.eval(ALL, *[torch.tensor(a, device=d) for a in (x, p, lam, np.ones(B))])
g = torch.Generator(device="cpu").manual_seed(3)
sig = (torch.rand((B, len(ine)), generator=g, dtype=torch.float64) * 10.0).to(d)
delta = torch.full((B,), 1e-2, dtype=torch.float64, device=d)
rx = torch.randn((B, ev.n_x), generator=g, dtype=torch.float64).to(d)
rE = torch.randn((B, len(eq)), generator=g, dtype=torch.float64).to(d)
if R > 1:
    rx = torch.randn((B, ev.n_x, R), generator=g, dtype=torch.float64).to(d)
    rE = torch.randn((B, len(eq), R), generator=g, dtype=torch.float64).to(d)
for _ in range(2):
    kkt.solve(out["hess"], out["jac"], sig, delta, 1e-9, rx, rE)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    kkt.solve(out["hess"], out["jac"], sig, delta, 1e-9, rx, rE)
    torch.cuda.synchronize()
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
kkt.solve(out["hess"], out["jac"], sig, delta, 1e-9, rx, rE)
torch.cuda.synchronize()
print(f"wall {1e3 * (time.perf_counter() - t0):.1f} ms for {B} instances, {R} right-hand side(s)")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
