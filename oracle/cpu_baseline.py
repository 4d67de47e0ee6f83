"""oracle/cpu_baseline.py -- MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline / reference arm).

Times the CPU oracle (the "port" of the reference's CasADi evaluation, oracle/{sx,nlp,kinodynamic}.py)
on the host cores: one worker process per core, each evaluating f, grad_f, g, jac_g and hess_l for its
share of a bounded sample of instances.  Graph construction (done once per problem structure by the
reference as well) is excluded from the timed region."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_STATE = {}


def _init(model, horizon, final, periodicity):
    from . import kinodynamic as kd

    nlp, _ = kd.build(model, kd.Settings(horizon=horizon, final_state_constraint=final,
                                         periodicity_constraint=periodicity))
    # warm the lazily-built tapes / gradient graphs on one instance-sized dummy call
    x = np.zeros((1, nlp.n_x))
    x[:, 133::189] = 1.0  # unit quaternions
    p = np.ones((1, nlp.n_p))
    nlp.eval_g(x, p), nlp.eval_f(x, p), nlp.eval_grad_f(x, p), nlp.eval_jac(x, p)
    nlp.eval_hess(x, p, np.zeros((1, nlp.m)), 1.0)
    _STATE["nlp"] = nlp


def _work(args):
    x, p, lam, sigma, with_hess = args
    nlp = _STATE["nlp"]
    t0 = time.perf_counter()
    nlp.eval_f(x, p)
    nlp.eval_grad_f(x, p)
    nlp.eval_g(x, p)
    nlp.eval_jac(x, p)
    if with_hess:
        nlp.eval_hess(x, p, lam, sigma)
    return time.perf_counter() - t0


class OraclePool:
    def __init__(self, model, horizon, final=False, periodicity=False, cores=None):
        self.cores = cores or len(os.sched_getaffinity(0))
        self.horizon = horizon
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_init,
                                                initargs=(model, horizon, final, periodicity))
        self.pool.map(abs, range(self.cores))  # wait for the initialisers

    def step(self, x, p, lam, sigma, with_hess=True):
        """Evaluate the sample once over all cores; returns (wall seconds, knot-evals)."""
        n = x.shape[0]
        parts = np.array_split(np.arange(n), min(self.cores, n))
        t0 = time.perf_counter()
        self.pool.map(_work, [(x[i], p[i], lam[i], sigma[i], with_hess) for i in parts])
        dt = time.perf_counter() - t0
        return dt, n * self.horizon

    def close(self):
        self.pool.close()
        self.pool.join()
