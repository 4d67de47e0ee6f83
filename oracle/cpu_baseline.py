"""oracle/cpu_baseline.py -- MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline / reference arm).

Times the CPU oracle (the "port" of the reference's CasADi evaluation, oracle/{sx,nlp,kinodynamic}.py)
on the host cores for f, grad_f, g, jac_g and hess_l of a bounded sample of instances.  Graph
construction (done once per problem structure by the reference as well) is excluded from the timed
region.  Two back ends:

  * C virtual machine (oracle/sxvm.c via oracle/cvm.py), OpenMP over instances, all cores -- used when
    oracle/_build/libsxvm.so exists (built by ``__graft_entry__.build()``);
  * numpy tape interpreter, one worker process per core -- fallback.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_STATE = {}


def _build(model, horizon, final, periodicity):
    from . import kinodynamic as kd

    nlp, _ = kd.build(model, kd.Settings(horizon=horizon, final_state_constraint=final,
                                         periodicity_constraint=periodicity))
    # warm the lazily-built tapes / gradient graphs / scatter plans on one dummy instance
    x = np.zeros((1, nlp.n_x))
    x[:, 133::189] = 1.0  # unit quaternions
    p = np.ones((1, nlp.n_p))
    _evaluate(nlp, x, p, np.zeros((1, nlp.m)), np.ones(1), True)
    return nlp


def _evaluate(nlp, x, p, lam, sigma, with_hess):
    nlp.eval_f(x, p)
    nlp.eval_grad_f(x, p)
    nlp.eval_g(x, p)
    nlp.eval_jac(x, p)
    if with_hess:
        nlp.eval_hess(x, p, lam, sigma)


def _init(model, horizon, final, periodicity):
    _STATE["nlp"] = _build(model, horizon, final, periodicity)


def _work(args):
    x, p, lam, sigma, with_hess = args
    t0 = time.perf_counter()
    _evaluate(_STATE["nlp"], x, p, lam, sigma, with_hess)
    return time.perf_counter() - t0


class OraclePool:
    def __init__(self, model, horizon, final=False, periodicity=False, cores=None, backend=None):
        from . import cvm

        self.horizon = horizon
        avail = len(os.sched_getaffinity(0))
        if backend is None:
            backend = "c" if cvm.available() else "numpy"
        self.backend = backend
        if backend == "c":
            self.cores = cores or avail
            cvm.enable(self.cores)
            self.nlp = _build(model, horizon, final, periodicity)
            self.pool = None
            self.description = "C SX virtual machine (oracle/sxvm.c), OpenMP over instances"
        else:
            self.cores = cores or avail
            self.pool = mp.get_context("fork").Pool(self.cores, initializer=_init,
                                                    initargs=(model, horizon, final, periodicity))
            self.pool.map(abs, range(self.cores))  # wait for the initialisers
            self.description = "numpy SX virtual machine, one process per core"

    def step(self, x, p, lam, sigma, with_hess=True):
        """Evaluate the sample once over all cores; returns (wall seconds, knot-evals)."""
        n = x.shape[0]
        t0 = time.perf_counter()
        if self.pool is None:
            _evaluate(self.nlp, x, p, lam, sigma, with_hess)
        else:
            parts = np.array_split(np.arange(n), min(self.cores, n))
            self.pool.map(_work, [(x[i], p[i], lam[i], sigma[i], with_hess) for i in parts])
        return time.perf_counter() - t0, n * self.horizon

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
        else:
            from . import cvm

            cvm.disable()
