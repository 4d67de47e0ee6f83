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

# ---- the compiled path: CasADi's external-function ABI (no Python inside the iterate besides these ctypes calls)
import ctypes  # noqa: E402

from hippopt_b200 import _capi  # noqa: E402

lib = _capi.lib()
assert lib.hb_external_bind(ev._h) == 0
dp = ctypes.POINTER(ctypes.c_double)
out_f, out_g, out_grad = np.zeros(1), np.zeros(ev.m), np.zeros(ev.n_x)
out_j, out_h = np.zeros(ev.nnz_j), np.zeros(ev.nnz_h)
sig = np.ones(1)
for fn in ("hb_nlp_f", "hb_nlp_grad_f", "hb_nlp_g", "hb_nlp_jac_g", "hb_nlp_hess_l"):
    getattr(lib, fn).restype = ctypes.c_int
res_f = (dp * 1)(out_f.ctypes.data_as(dp))
res_gf = (dp * 2)(out_f.ctypes.data_as(dp), out_grad.ctypes.data_as(dp))
res_g = (dp * 1)(out_g.ctypes.data_as(dp))
res_j = (dp * 2)(out_g.ctypes.data_as(dp), out_j.ctypes.data_as(dp))
res_h = (dp * 1)(out_h.ctypes.data_as(dp))
t_first, t_hess = [], []
for i, xi in enumerate(xs):
    xi = np.ascontiguousarray(xi)
    arg = (dp * 2)(xi.ctypes.data_as(dp), p[0].ctypes.data_as(dp))
    arg4 = (dp * 4)(xi.ctypes.data_as(dp), p[0].ctypes.data_as(dp), sig.ctypes.data_as(dp), lam[0].ctypes.data_as(dp))
    t0 = time.perf_counter()
    lib.hb_nlp_f(arg, res_f, None, None, 0)
    lib.hb_nlp_grad_f(arg, res_gf, None, None, 0)
    lib.hb_nlp_g(arg, res_g, None, None, 0)
    lib.hb_nlp_jac_g(arg, res_j, None, None, 0)
    t1 = time.perf_counter()
    lib.hb_nlp_hess_l(arg4, res_h, None, None, 0)
    t2 = time.perf_counter()
    if i >= 20:
        t_first.append(t1 - t0)
        t_hess.append(t2 - t1)
lib.hb_external_bind(None)
f, h = np.array(t_first) * 1e6, np.array(t_hess) * 1e6
print(f"B=1 external-function ABI (hb_nlp_*): f, grad_f, g, jac_g calls median {np.median(f):.0f} us (p90 {np.percentile(f, 90):.0f}), "
      f"hess_l median {np.median(h):.0f} us (p90 {np.percentile(h, 90):.0f}); one IPOPT iterate = {np.median(f) + np.median(h):.0f} us "
      f"= {1e6 / (np.median(f) + np.median(h)):.0f} iterates/s")
