"""What a CPU-side IPOPT sees through the Callback shim: latency of ONE instance's evaluations through
hb_eval_host (HostEvaluator + OracleCache, hippopt_b200/plugin.py) -- first-order set and hess_l, config-3 size."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from hippopt_b200 import plugin  # noqa: E402
from hippopt_b200.evaluator import KinoEvaluator  # noqa: E402
from hippopt_b200.kino_layout import KinoSettings  # noqa: E402
from hippopt_b200.robot_model import synthetic_ergocub  # noqa: E402
from hippopt_b200.workloads import kino_batch  # noqa: E402

model = synthetic_ergocub()
ev = KinoEvaluator(model, KinoSettings(horizon=30))
x, p, lam, sigma = kino_batch(ev.layout, model, 1, seed=3)
host = plugin.HostEvaluator(ev)
host.set_parameters(p[0])
cache = plugin.OracleCache(host, host.masks)
rng = np.random.default_rng(0)
xs = [x[0] + 1e-3 * rng.normal(size=x.shape[1]) for _ in range(220)]
for xi in xs[:20]:
    cache.get("f", xi)
    cache.hess(xi, lam[0], 1.0)
t_first, t_hess = [], []
for xi in xs[20:]:
    t0 = time.perf_counter()
    for name in ("f", "grad_f", "g", "jac"):
        cache.get(name, xi)
    t1 = time.perf_counter()
    cache.hess(xi, lam[0], 1.0)
    t2 = time.perf_counter()
    t_first.append(t1 - t0)
    t_hess.append(t2 - t1)
f, h = np.array(t_first) * 1e6, np.array(t_hess) * 1e6
print(f"B=1 callback latency (config 3, N=30): f+grad_f+g+jac_g median {np.median(f):.0f} us (p90 {np.percentile(f, 90):.0f}), "
      f"hess_l median {np.median(h):.0f} us (p90 {np.percentile(h, 90):.0f}); one IPOPT iterate = {np.median(f) + np.median(h):.0f} us "
      f"= {1e6 / (np.median(f) + np.median(h)):.0f} iterates/s")
