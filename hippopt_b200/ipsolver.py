"""Batched primal-dual interior-point driver.

What sits directly above the evaluation path in the reference is IPOPT (`opti.solve()`,
`/root/reference/src/hippopt/base/opti_solver.py:479`), one serial solve at a time.  Here B independent
instances advance in lock-step on the GPU: every iteration makes ONE batched `hb_eval` call (f, grad_f,
g, jac_g, hess_l for all instances) and one batched KKT solve for the instances that still need a step.
The algorithm is the textbook line-search barrier method IPOPT implements (Waechter & Biegler 2006)
reduced to what a batched solver needs: slacks on the inequality rows, monotone (Fiacco-McCormick)
barrier update, fraction-to-boundary rule, l1 merit function with Armijo backtracking plus filter-type
and f-type step acceptance, and a Levenberg-type Hessian shift when the reduced Hessian is not positive
along the step (several shifts of the sequence per sweep while the batch is below a wave of the LU
kernels).  There is no restoration phase: an instance whose line search fails `max_fail` times in a row
is given up.

KKT back ends: "dense" (torch.linalg, one (n_x + m_E)-square solve per instance: toy OCP, pose finder)
and "stage" (hippopt_b200.kkt.StageKKT on the batched LU kernels: the kinodynamic OCPs).  Termination
and scaling follow IPOPT's options (`ipopt_options`), failures the reference's `OptiFailure` and callback
criteria (`callback_criterion`, hippopt_b200.opti_callback).

Conventions follow IPOPT / CasADi: L = sigma f + lam^T g, lam > 0 on an active upper bound.
"""
from __future__ import annotations

import dataclasses
import time

import numpy as np
import torch

from .evaluator import ALL, F, G


class OptiFailure(Exception):
    """Mirror of `hippopt.OptiFailure` (`base/opti_solver.py:28-37`): raised when no instance converged."""

    def __init__(self, message: str, callback_used: bool = False):
        callback_info = " and the callback did not manage to save an intermediate solution" if callback_used else ""
        super().__init__(f"Opti failed to solve the problem{callback_info}. Message: {message}")


@dataclasses.dataclass
class BatchedOutput:
    """Per-instance counterpart of `hippopt.Output` (`base/problem.py:28-79`)."""

    values: torch.Tensor                  # (B, n_x) primal solution
    cost_value: torch.Tensor              # (B,)
    constraint_multipliers: torch.Tensor  # (B, m)   lam_g
    success: torch.Tensor                 # (B,) bool
    iterations: torch.Tensor              # (B,) int
    kkt_error: torch.Tensor               # (B,) scaled optimality error at exit
    evaluations: int = 0                  # batched hb_eval calls made
    acceptable: torch.Tensor = None       # (B,) bool: stopped at IPOPT's "acceptable level" (part of `success`)
    callback_iteration: torch.Tensor = None  # (B,) int: >= 0 where a FAILED instance returns the iterate its callback
    #                                          criterion saved at that iteration (opti_solver.py:478-520), else -1


class SparseOps:
    """Products with jac_g / hess_l straight from the CCS value arrays the kernels write (one pattern shared by
    all instances).  On a CUDA device they run through `hb_ccs_group_mul` (one thread per output element, fixed
    summation order: bit-reproducible); on the CPU (tests with a torch stand-in evaluator) through `index_add_`,
    which is sequential there."""

    def __init__(self, n_x, m, jac_sparsity, hess_sparsity, device):
        colind, row = jac_sparsity
        self.n, self.m = n_x, m
        row, colind = np.asarray(row), np.asarray(colind)
        col = np.repeat(np.arange(n_x), np.diff(colind))
        self.jr = torch.as_tensor(row, dtype=torch.long, device=device)
        self.jc = torch.as_tensor(col, dtype=torch.long, device=device)
        hcolind, hrow = hess_sparsity
        hrow = np.asarray(hrow)
        hcol = np.repeat(np.arange(n_x), np.diff(np.asarray(hcolind)))
        self.hr = torch.as_tensor(hrow, dtype=torch.long, device=device)
        self.hc = torch.as_tensor(hcol, dtype=torch.long, device=device)
        self.hw = torch.where(self.hr == self.hc, 1.0, 2.0).to(torch.float64)  # upper triangle stored once
        self.native = torch.device(device).type == "cuda"
        if self.native:
            def group(keys, n_out, entry, idx):
                order = np.argsort(keys, kind="stable")
                ptr = np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=n_out))])
                i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=device)  # noqa: E731
                return i32(ptr), i32(entry[order]), i32(idx[order])

            e = np.arange(len(row))
            self._by_row = group(row, m, e, col)          # J x
            self._by_col = group(col, n_x, e, row)        # J^T lam
            eh = np.arange(len(hrow))
            off = hrow != hcol                            # H x over the mirrored upper triangle
            self._hess = group(np.concatenate([hrow, hcol[off]]), n_x, np.concatenate([eh, eh[off]]),
                               np.concatenate([hcol, hrow[off]]))

    def _mul(self, tables, vals, x, n_out):
        from . import _capi

        vals, x = vals.contiguous(), x.contiguous()
        out = torch.empty((vals.shape[0], n_out), dtype=torch.float64, device=vals.device)
        ptr, entry, idx = tables
        stream = torch.cuda.current_stream(vals.device).cuda_stream
        vp = lambda t: __import__("ctypes").c_void_p(t.data_ptr())  # noqa: E731
        _capi.check(_capi.lib().hb_ccs_group_mul(vp(vals), vp(ptr), vp(entry), vp(idx), None, vp(x), vp(out), n_out,
                                                 x.shape[1], vals.shape[1], vals.shape[0],
                                                 __import__("ctypes").c_void_p(stream)), "hb_ccs_group_mul")
        return out

    def J_mul(self, vals, x):
        if self.native:
            return self._mul(self._by_row, vals, x, self.m)
        out = torch.zeros((vals.shape[0], self.m), dtype=vals.dtype, device=vals.device)
        return out.index_add_(1, self.jr, vals * x[:, self.jc])

    def Jt_mul(self, vals, lam):
        if self.native:
            return self._mul(self._by_col, vals, lam, self.n)
        out = torch.zeros((vals.shape[0], self.n), dtype=vals.dtype, device=vals.device)
        return out.index_add_(1, self.jc, vals * lam[:, self.jr])

    def W_quad(self, hvals, x):
        if self.native:
            return (self._mul(self._hess, hvals, x, self.n) * x).sum(dim=1)
        return (hvals * x[:, self.hr] * x[:, self.hc] * self.hw).sum(dim=1)

    def dense_jac(self, vals):
        J = torch.zeros((vals.shape[0], self.m, self.n), dtype=vals.dtype, device=vals.device)
        J[:, self.jr, self.jc] = vals
        return J

    def dense_hess(self, vals):
        H = torch.zeros((vals.shape[0], self.n, self.n), dtype=vals.dtype, device=vals.device)
        H[:, self.hr, self.hc] = vals
        H[:, self.hc, self.hr] = vals
        return H


class DenseKKT:
    """One dense (n_x + m_E)-square solve per instance (torch.linalg): the small NLPs (toy OCP, pose finder)."""

    def __init__(self, ops: SparseOps, iE, iI):
        self.ops, self.iE, self.iI = ops, iE, iI
        self.K = None

    def solve(self, hess_vals, jac_vals, sigma_I, delta, delta_c, rhs_x, rhs_E):
        ops = self.ops
        B, n, mE = hess_vals.shape[0], ops.n, len(self.iE)
        J = ops.dense_jac(jac_vals)
        JE, JI = J[:, self.iE, :], J[:, self.iI, :]
        K = torch.zeros((B, n + mE, n + mE), dtype=hess_vals.dtype, device=hess_vals.device)
        K[:, :n, :n] = ops.dense_hess(hess_vals) + torch.einsum("bin,bi,bik->bnk", JI, sigma_I, JI)
        K[:, :n, :n] += delta[:, None, None] * torch.eye(n, dtype=K.dtype, device=K.device)
        K[:, :n, n:] = JE.transpose(1, 2)
        K[:, n:, :n] = JE
        if mE and delta_c:
            K[:, n:, n:] = -delta_c * torch.eye(mE, dtype=K.dtype, device=K.device)
        self.K = K
        multi = rhs_x.dim() == 3  # (B, n, R) right-hand sides: the low-rank (L-BFGS) correction rides along
        rhs = torch.cat([rhs_x, rhs_E], dim=1)
        # a singular instance gets NaN (-> larger shift for THAT instance); the others keep their solution
        sol, info = torch.linalg.solve_ex(K, rhs if multi else rhs[:, :, None])
        sol = torch.where((info != 0)[:, None, None], torch.full_like(sol, float("nan")), sol)
        if not multi:
            sol = sol[:, :, 0]
        return sol[:, :n], sol[:, n:]


class LimitedMemory:
    """IPOPT's `hessian_approximation = limited-memory` (the setting of every kinodynamic main of the reference,
    e.g. main_periodic_step.py:116): the Hessian of the Lagrangian is replaced by an L-BFGS matrix in compact form
    (Byrd, Nocedal, Schnabel 1994), B = sigma I - U M^-1 U^T with U = [sigma S, Y],
    M = [[sigma S^T S, L], [L^T, -D]], L / D the strictly lower / diagonal part of S^T Y, one memory per instance.
    The KKT matrix is then K0 - [U; 0] M^-1 [U; 0]^T with a DIAGONAL Hessian block in K0, and the step comes from
    one sweep with 1 + 2k right-hand sides (Sherman-Morrison-Woodbury).  Pairs with s^T y <= sqrt(eps) |s| |y| are
    skipped (IPOPT's limited_memory_update_type = bfgs); sigma = s^T y / s^T s of the newest pair
    (limited_memory_initialization = scalar1)."""

    def __init__(self, B, n, k, device):
        self.k = k
        self.S = torch.zeros((B, n, k), dtype=torch.float64, device=device)
        self.Y = torch.zeros((B, n, k), dtype=torch.float64, device=device)
        self.sigma = torch.ones(B, dtype=torch.float64, device=device)
        self.count = torch.zeros(B, dtype=torch.long, device=device)
        self.skipped = torch.zeros(B, dtype=torch.long, device=device)

    def update(self, active, s, y):
        sy, ss, yy = (s * y).sum(1), (s * s).sum(1), (y * y).sum(1)
        ok = active & (ss > 0) & (sy > 1.4901161193847656e-08 * torch.sqrt(ss * yy))
        self.skipped = torch.where(active & ~ok, self.skipped + 1, torch.where(ok, torch.zeros_like(self.skipped), self.skipped))
        # IPOPT resets the memory after limited_memory_max_skipping = 2 consecutive skips
        reset = self.skipped >= 2
        self.S = torch.where(reset[:, None, None], torch.zeros_like(self.S), self.S)
        self.Y = torch.where(reset[:, None, None], torch.zeros_like(self.Y), self.Y)
        self.count = torch.where(reset, torch.zeros_like(self.count), self.count)
        self.sigma = torch.where(reset, torch.ones_like(self.sigma), self.sigma)
        self.skipped = torch.where(reset, torch.zeros_like(self.skipped), self.skipped)
        S_new = torch.cat([self.S[:, :, 1:], s[:, :, None]], dim=2)
        Y_new = torch.cat([self.Y[:, :, 1:], y[:, :, None]], dim=2)
        self.S = torch.where(ok[:, None, None], S_new, self.S)
        self.Y = torch.where(ok[:, None, None], Y_new, self.Y)
        self.count = torch.where(ok, torch.clamp(self.count + 1, max=self.k), self.count)
        self.sigma = torch.where(ok, sy / torch.clamp(ss, min=1e-300), self.sigma)

    def matrices(self, idx):
        """sigma (b,), U (b, n, 2k), M (b, 2k, 2k) of the instances idx; unused pairs are zero columns with +-1 on M's
        diagonal."""
        S, Y, sg, k = self.S[idx], self.Y[idx], self.sigma[idx], self.k
        U = torch.cat([sg[:, None, None] * S, Y], dim=2)
        StY = torch.bmm(S.transpose(1, 2), Y)
        L = torch.tril(StY, diagonal=-1)
        D = torch.diagonal(StY, dim1=1, dim2=2)
        M = torch.zeros((S.shape[0], 2 * k, 2 * k), dtype=S.dtype, device=S.device)
        M[:, :k, :k] = sg[:, None, None] * torch.bmm(S.transpose(1, 2), S)
        M[:, :k, k:] = L
        M[:, k:, :k] = L.transpose(1, 2)
        M[:, k:, k:] = -torch.diag_embed(D)
        used = torch.arange(k, device=S.device)[None, :] >= (k - self.count[idx])[:, None]  # newest pairs sit at the end
        pad = torch.cat([(~used).to(S.dtype), -(~used).to(S.dtype)], dim=1)
        M = M + torch.diag_embed(pad)
        return sg, U, M

    def apply(self, idx, v):
        """B v for the instances idx."""
        sg, U, M = self.matrices(idx)
        t = torch.linalg.solve(M, torch.bmm(U.transpose(1, 2), v[:, :, None]))
        return sg[:, None] * v - torch.bmm(U, t)[:, :, 0]

    def kkt_solve(self, backend, idx, zero_h, jv, Sig, delta, dc, rhs_x, rhs_E):
        sg, U, M = self.matrices(idx)
        R = U.shape[2]
        RX = torch.cat([rhs_x[:, :, None], U], dim=2)
        RE = torch.cat([rhs_E[:, :, None], torch.zeros((rhs_E.shape[0], rhs_E.shape[1], R), dtype=U.dtype, device=U.device)], dim=2)
        ZX, ZE = backend.solve(zero_h, jv, Sig, sg + delta, dc, RX, RE)
        G = M - torch.bmm(U.transpose(1, 2), ZX[:, :, 1:])
        w, info = torch.linalg.solve_ex(G, torch.bmm(U.transpose(1, 2), ZX[:, :, :1]))
        w = torch.where((info != 0)[:, None, None], torch.full_like(w, float("nan")), w)
        dx = ZX[:, :, 0] + torch.bmm(ZX[:, :, 1:], w)[:, :, 0]
        dl = ZE[:, :, 0] + torch.bmm(ZE[:, :, 1:], w)[:, :, 0]
        return dx, dl


class BatchedInteriorPoint:
    def __init__(self, ev, tol: float = 1e-8, max_iter: int = 300, mu_init: float = 0.1, kappa_eps: float = 10.0,
                 kappa_mu: float = 0.2, theta_mu: float = 1.5, tau_min: float = 0.99, eta: float = 1e-4,
                 max_backtrack: int = 16, delta_min: float = 1e-8, delta_max: float = 1e8, exact_inertia: bool = False,
                 verbose: bool = False, kkt: str = "dense", delta_c: float = 1e-11, f_type: bool = True,
                 max_fail: int = 8, ipopt_options: dict | None = None, callback_criterion=None):
        """kkt: "dense" (one dense factorisation per instance) or "stage" (block-tridiagonal sweep over the knots,
        hippopt_b200.kkt.StageKKT -- the multiple-shooting OCPs of the kinodynamic planner).

        ipopt_options: the subset of IPOPT's termination / scaling options the reference's mains pass through
        `casadi_solver_options` (e.g. main_periodic_step.py:111-134), with IPOPT's meaning: "tol", "max_iter",
        "dual_inf_tol", "constr_viol_tol", "compl_inf_tol" (desired level: scaled error <= tol AND the three unscaled
        measures below their tolerances), "acceptable_tol", "acceptable_iter", "acceptable_dual_inf_tol",
        "acceptable_constr_viol_tol", "acceptable_compl_inf_tol", "acceptable_obj_change_tol" (stop after
        acceptable_iter consecutive iterations at the acceptable level), "nlp_scaling_method" ("gradient-based":
        objective scaled by min(1, nlp_scaling_max_gradient / |grad f(x0)|_inf); constraints are not rescaled;
        "none", the default here).  Unknown keys are ignored, as options of parts this driver does not have.

        callback_criterion: a hippopt_b200.opti_callback criterion (the reference's `OptiSolver(callback_criterion=)`,
        opti_solver.py:109,451-520): after every iteration the iterate of each unfinished instance whose criterion
        is satisfied is saved, and an instance that fails returns that iterate instead of its last one; OptiFailure
        is raised only if no instance converged AND none has a saved iterate."""
        self.callback_criterion = callback_criterion
        # instances x shifts factored per sweep (stage backend on a GPU: one CTA per SM); 0 disables the speculation
        self.spec_wave = 0
        self.speculative_shifts = True
        o = dict(ipopt_options or {})
        tol, max_iter = float(o.get("tol", tol)), int(o.get("max_iter", max_iter))
        self.dual_inf_tol, self.constr_viol_tol = o.get("dual_inf_tol"), o.get("constr_viol_tol")
        self.compl_inf_tol = o.get("compl_inf_tol")
        self.acceptable_tol = o.get("acceptable_tol")  # None: no acceptable-level termination
        self.acceptable_iter = int(o.get("acceptable_iter", 15))
        self.acceptable_dual_inf_tol = float(o.get("acceptable_dual_inf_tol", 1e10))
        self.acceptable_constr_viol_tol = float(o.get("acceptable_constr_viol_tol", 1e-2))
        self.acceptable_compl_inf_tol = float(o.get("acceptable_compl_inf_tol", 1e-2))
        self.acceptable_obj_change_tol = float(o.get("acceptable_obj_change_tol", 1e20))
        # warm start (IPOPT's warm_start_* options, main_single_step_flat_ground.py:120-125): used when solve() is
        # given multipliers of a previous solution
        self.warm_start = str(o.get("warm_start_init_point", "no")) == "yes"
        self.ws_slack_push = float(o.get("warm_start_slack_bound_push", 1e-3))
        self.ws_slack_frac = float(o.get("warm_start_slack_bound_frac", 1e-3))
        self.ws_mult_push = float(o.get("warm_start_mult_bound_push", 1e-3))
        if "mu_init" in o:
            mu_init = float(o["mu_init"])
        # hessian_approximation = limited-memory (main_periodic_step.py:116): L-BFGS instead of hess_l
        # feasibility restoration (simplified, see _restoration below): entered after `resto_after` failed line searches
        # in a row, left once the violation fell to required_infeasibility_reduction x its value at entry (IPOPT's
        # option, 0.9 by default; the reference's mains set 0.8)
        # Off by default: on every workload of this repository the filter / f-type acceptance and the Levenberg shifts
        # never fail twice in a row (profiles/r02/solver_v15.txt), so the phase is exercised only when asked for
        # ("hb_restoration": True, or IPOPT's "start_with_resto": "yes", which begins in it).
        self.start_with_resto = str(o.get("start_with_resto", "no")) == "yes"
        self.restoration = self.start_with_resto or bool(o.get("hb_restoration", False))
        self.resto_after = int(o.get("hb_restoration_after", 2))
        self.resto_reduction = float(o.get("required_infeasibility_reduction", 0.9))
        self.limited_memory = str(o.get("hessian_approximation", "exact")) == "limited-memory"
        self.lm_history = int(o.get("limited_memory_max_history", 6))
        self.obj_scaling = o.get("nlp_scaling_method", "none") == "gradient-based"
        self.scaling_max_gradient = float(o.get("nlp_scaling_max_gradient", 100.0))
        self.ev = ev
        self.exact_inertia = exact_inertia
        self.tol, self.max_iter, self.mu_init = tol, max_iter, mu_init
        self.kappa_eps, self.kappa_mu, self.theta_mu = kappa_eps, kappa_mu, theta_mu
        self.tau_min, self.eta, self.max_backtrack = tau_min, eta, max_backtrack
        self.delta_min, self.delta_max = delta_min, delta_max
        self.verbose = verbose
        if kkt not in ("dense", "stage"):
            raise ValueError("kkt must be 'dense' or 'stage'")
        self.kkt_kind, self.delta_c, self.f_type, self.max_fail = kkt, delta_c, f_type, max_fail
        self.kkt_seconds = 0.0

    # ------------------------------------------------------------------ solve
    def solve(self, x0: torch.Tensor, p: torch.Tensor, lbg, ubg, lam0=None) -> BatchedOutput:
        """lam0 (B, m), optional: constraint multipliers of a previous solution (IPOPT sign).  With them the solve is
        warm-started as IPOPT does with warm_start_init_point = yes: slacks pushed by warm_start_slack_bound_push / _frac
        only, equality multipliers taken over, bound multipliers = the given ones floored at
        warm_start_mult_bound_push.  (Set "mu_init" low as well: a warm start from a solution with mu = 0.1 walks
        back out along the central path.)"""
        ev, dev = self.ev, x0.device
        B, n, m = x0.shape[0], ev.n_x, ev.m
        lbg = torch.as_tensor(np.broadcast_to(np.asarray(lbg, dtype=np.float64), (B, m)).copy(), device=dev)
        ubg = torch.as_tensor(np.broadcast_to(np.asarray(ubg, dtype=np.float64), (B, m)).copy(), device=dev)
        eq = (lbg[0] == ubg[0])
        free = torch.isinf(lbg[0]) & torch.isinf(ubg[0])
        ine = ~eq & ~free
        if not (torch.equal((lbg == ubg), eq.expand(B, m)) and torch.equal(torch.isinf(lbg) & torch.isinf(ubg), free.expand(B, m))):
            raise ValueError("all instances must share the equality / inequality structure of their bounds")
        iE, iI = torch.nonzero(eq).ravel(), torch.nonzero(ine).ravel()
        mE, mI = len(iE), len(iI)
        ops = SparseOps(n, m, ev.jac_sparsity(), ev.hess_sparsity(), dev)
        if self.kkt_kind == "stage":
            from .kkt import StageKKT

            lay = getattr(ev, "layout", None)
            if lay is None or not hasattr(lay, "N") or not hasattr(lay, "knot_size"):
                raise ValueError("kkt='stage' needs an evaluator with a multiple-shooting layout (layout.N knots of "
                                 "layout.knot_size variables, e.g. KinoEvaluator); use kkt='dense' for "
                                 f"{type(ev).__name__}")
            jc_, jr_ = ev.jac_sparsity()
            hc_, hr_ = ev.hess_sparsity()
            backend = StageKKT(n, m, lay.N, lay.knot_size, jc_, jr_, hc_, hr_, iE.cpu().numpy(), iI.cpu().numpy(),
                               device=dev)
            if dev.type == "cuda" and self.speculative_shifts:
                self.spec_wave = torch.cuda.get_device_properties(dev).multi_processor_count
        else:
            backend = DenseKKT(ops, iE, iI)
        lb, ub = lbg[:, iI], ubg[:, iI]
        hasL, hasU = torch.isfinite(lb), torch.isfinite(ub)
        lbE = lbg[:, iE]
        x = x0.clone().contiguous()
        ones = torch.ones(B, dtype=torch.float64, device=dev)
        zeros_m = torch.zeros((B, m), dtype=torch.float64, device=dev)

        obj_scale = ones  # IPOPT's gradient-based objective scaling (set below); sigma of the Hessian carries it

        first_order = ALL & ~16 if self.limited_memory else ALL  # IPOPT never asks for hess_l with L-BFGS

        def evaluate(xx, lam=None, full=True):
            out = ev.eval(first_order if full else (F | G), xx, p, lam if lam is not None else zeros_m, obj_scale)
            out = {k: v.clone() for k, v in out.items()}
            if self.obj_scaling:
                out["f"] = out["f"] * obj_scale
                if "grad_f" in out:
                    out["grad_f"] = out["grad_f"] * obj_scale[:, None]
            return out

        if self.obj_scaling:
            g0 = ev.eval(ALL, x, p, zeros_m, ones)["grad_f"].abs().amax(dim=1)
            obj_scale = torch.clamp(self.scaling_max_gradient / torch.clamp(g0, min=1e-300), max=1.0).contiguous()

        big = 1e300
        lbs = torch.where(hasL, lb, torch.full_like(lb, -big))
        ubs = torch.where(hasU, ub, torch.full_like(ub, big))

        out = evaluate(x, full=False)
        n_eval = 2 if self.obj_scaling else 1
        gI = out["g"][:, iI]
        # push the slacks strictly inside their bounds (IPOPT bound_push / bound_frac)
        # (separate pushes per side, each only where that bound exists: a one-sided row must not inherit the
        # 1e300 stand-in of its missing bound)
        warm = lam0 is not None
        # warm start: IPOPT pushes by warm_start_slack_bound_push whatever mu_init is, which throws an active row of
        # the previous solution (slack ~ mu / z) far back inside; capped here at a tenth of mu_init so that the
        # start keeps the complementarity the previous solve ended with
        push, frac = ((min(self.ws_slack_push, 0.1 * self.mu_init), min(self.ws_slack_frac, 0.1 * self.mu_init))
                      if warm else (1e-2, 1e-2))
        width = torch.where(hasL & hasU, ub - lb, torch.full_like(lb, float("inf")))
        pL = torch.minimum(push * torch.clamp(lb.abs(), min=1.0), frac * width)
        pU = torch.minimum(push * torch.clamp(ub.abs(), min=1.0), frac * width)
        s = torch.where(hasL, torch.maximum(gI, lb + pL), gI)
        s = torch.where(hasU, torch.minimum(s, ub - pU), s)
        mu = torch.full((B,), self.mu_init, dtype=torch.float64, device=dev)
        zL = torch.where(hasL, mu[:, None] / (s - lbs), torch.zeros_like(s))
        zU = torch.where(hasU, mu[:, None] / (ubs - s), torch.zeros_like(s))
        lamE = torch.zeros((B, mE), dtype=torch.float64, device=dev)
        if warm:
            lam0 = torch.as_tensor(np.asarray(lam0.cpu() if torch.is_tensor(lam0) else lam0, dtype=np.float64), device=dev)
            lam0 = lam0.expand(B, m) * obj_scale[:, None]
            lamE = lam0[:, iE].clone()
            lI = lam0[:, iI]
            # IPOPT floors the given bound multipliers at warm_start_mult_bound_push; here the floor is capped at
            # 0.1 mu / slack, so that a constraint that is inactive at the previous solution does not come back with a
            # complementarity of push * slack >> mu (which costs IPOPT-style warm starts a dozen iterations)
            floorU = torch.minimum(torch.full_like(s, self.ws_mult_push), 0.1 * mu[:, None] / torch.clamp(ubs - s, min=1e-300))
            floorL = torch.minimum(torch.full_like(s, self.ws_mult_push), 0.1 * mu[:, None] / torch.clamp(s - lbs, min=1e-300))
            zU = torch.where(hasU, torch.maximum(lI, floorU), torch.zeros_like(s))
            zL = torch.where(hasL, torch.maximum(-lI, floorL), torch.zeros_like(s))
        delta = torch.zeros(B, dtype=torch.float64, device=dev)
        nu = torch.ones(B, dtype=torch.float64, device=dev)
        done = torch.zeros(B, dtype=torch.bool, device=dev)
        stalled = torch.zeros(B, dtype=torch.bool, device=dev)  # given up: max_fail line-search failures in a row
        fails = torch.zeros(B, dtype=torch.long, device=dev)
        resto = torch.zeros(B, dtype=torch.bool, device=dev)      # instances in the restoration phase
        theta_entry = torch.zeros(B, dtype=torch.float64, device=dev)
        self.restoration_entries = 0
        acceptable = torch.zeros(B, dtype=torch.bool, device=dev)
        acc_count = torch.zeros(B, dtype=torch.long, device=dev)
        f_prev = torch.full((B,), float("inf"), dtype=torch.float64, device=dev)
        iters = torch.zeros(B, dtype=torch.long, device=dev)
        best_it = torch.full((B,), -1, dtype=torch.long, device=dev)
        if self.callback_criterion is not None:
            self.callback_criterion.reset(B, dev)
            best_x, best_lam = x.clone(), zeros_m.clone()
            best_cost = torch.full((B,), float("inf"), dtype=torch.float64, device=dev)
        err0 = torch.full((B,), float("inf"), dtype=torch.float64, device=dev)
        def rows_I(v):  # (B, m_I) values on the inequality rows -> (B, m) with zeros elsewhere
            full = torch.zeros((B, m), dtype=torch.float64, device=dev)
            full[:, iI] = v
            return full

        def barrier(fv, ss, muv):
            t = torch.where(hasL, torch.log(ss - lbs), torch.zeros_like(ss)) + torch.where(hasU, torch.log(ubs - ss), torch.zeros_like(ss))
            return fv - muv * t.sum(dim=1)

        for it in range(self.max_iter):
            lam = zeros_m.clone()
            lam[:, iE] = lamE
            lam[:, iI] = zU - zL
            out = evaluate(x, lam)
            n_eval += 1
            fv, grad, g = out["f"], out["grad_f"], out["g"]
            jv = out["jac"]
            if self.limited_memory:
                if it == 0:
                    lm = LimitedMemory(B, n, self.lm_history, dev)
                    hv = torch.zeros((B, ev.nnz_h if hasattr(ev, "nnz_h") else len(ev.hess_sparsity()[1])),
                                     dtype=torch.float64, device=dev)
                else:  # secant pair of the step just taken, with the CURRENT multipliers on both sides
                    y_lm = (grad + ops.Jt_mul(jv, lam)) - (grad_prev + ops.Jt_mul(jv_prev, lam))
                    lm.update(moved_prev, x - x_prev, y_lm)
                x_prev, grad_prev, jv_prev = x.clone(), grad.clone(), jv.clone()
            else:
                hv = out["hess"]
            cE = g[:, iE] - lbE
            cI = g[:, iI] - s
            rd = grad + ops.Jt_mul(jv, lam)
            dL, dU = s - lbs, ubs - s
            compL = torch.where(hasL, dL * zL, torch.zeros_like(s))
            compU = torch.where(hasU, dU * zU, torch.zeros_like(s))

            def linf(t):
                return t.abs().amax(dim=1) if t.shape[1] else torch.zeros(B, dtype=torch.float64, device=dev)

            # IPOPT's scaled optimality error E_mu
            # (Waechter & Biegler eq. 6: s_d from all multipliers -- constraint multipliers of the slack formulation
            # and bound multipliers --, s_c from the bound multipliers alone; s_max = 100)
            n_z = int(hasL[0].sum() + hasU[0].sum())
            sd = torch.clamp((lam.abs().sum(1) + zL.sum(1) + zU.sum(1)) / max(1, m + n_z) / 100.0, min=1.0)
            sc = torch.clamp((zL.sum(1) + zU.sum(1)) / max(1, n_z) / 100.0, min=1.0)
            prim = torch.maximum(linf(cE), linf(cI))

            def emu(muv):
                cL = torch.where(hasL, compL - muv[:, None], torch.zeros_like(s))
                cU = torch.where(hasU, compU - muv[:, None], torch.zeros_like(s))
                return torch.maximum(torch.maximum(linf(rd) / sd, prim), torch.maximum(linf(cL), linf(cU)) / sc)

            err0 = torch.where(done, err0, emu(torch.zeros_like(mu)))
            # IPOPT's termination tests: desired level = scaled error AND (optionally) the unscaled measures;
            # acceptable level = the looser set, acceptable_iter times in a row
            dual_u, compl_u = linf(rd) / obj_scale, torch.maximum(linf(compL), linf(compU))
            newly = (~done) & (err0 <= self.tol)
            for limit, val in ((self.dual_inf_tol, dual_u), (self.constr_viol_tol, prim), (self.compl_inf_tol, compl_u)):
                if limit is not None:
                    newly &= val <= float(limit)
            if self.acceptable_tol is not None:
                acc = ((~done) & (err0 <= float(self.acceptable_tol)) & (dual_u <= self.acceptable_dual_inf_tol)
                       & (prim <= self.acceptable_constr_viol_tol) & (compl_u <= self.acceptable_compl_inf_tol)
                       & ((fv - f_prev).abs() / torch.clamp(fv.abs(), min=1.0) <= self.acceptable_obj_change_tol))
                acc_count = torch.where(acc, acc_count + 1, torch.zeros_like(acc_count))
                newly_acc = (acc_count >= self.acceptable_iter) & ~newly
                acceptable |= newly_acc
                newly |= newly_acc
            f_prev = torch.where(done, f_prev, fv)
            if self.callback_criterion is not None:
                # SaveBestUnsolvedVariablesCallback.call (opti_callback.py:342-373), per instance: runs for every
                # instance that had not finished before this iteration, the converging iterate included
                cost_u = fv / obj_scale
                sat = self.callback_criterion.satisfied(cost_u, prim) & ~done & ~stalled & torch.isfinite(cost_u)
                self.callback_criterion.update(sat, cost_u, prim)
                best_x = torch.where(sat[:, None], x, best_x)
                best_lam = torch.where(sat[:, None], lam, best_lam)
                best_cost = torch.where(sat, cost_u, best_cost)
                best_it = torch.where(sat, torch.full_like(best_it, it), best_it)
            done |= newly
            if self.verbose and it % getattr(self, "verbose_every", 10) == 0:
                print(f"it {it:3d} done {int(done.sum())}/{B} err0 med {err0.median().item():.2e} max {err0.max().item():.2e} "
                      f"mu med {mu.median().item():.1e} delta med {delta.median().item():.1e} max {delta.max().item():.1e} "
                      f"| dual med {(linf(rd) / sd).median().item():.2e} prim med {prim.median().item():.2e}")
            inactive = done | stalled
            if bool(inactive.all()):
                break
            iters += (~inactive).long()
            # barrier parameter update (possibly several reductions in one go)
            for _ in range(4):
                dec = (~done) & (emu(mu) <= self.kappa_eps * mu) & (mu > self.tol / 10.0)
                if not bool(dec.any()):
                    break
                mu = torch.where(dec, torch.clamp(torch.minimum(self.kappa_mu * mu, mu ** self.theta_mu), min=self.tol / 10.0), mu)
            tau = torch.clamp(1.0 - mu, min=self.tau_min)
            SigL = torch.where(hasL, zL / dL, torch.zeros_like(s))
            SigU = torch.where(hasU, zU / dU, torch.zeros_like(s))
            Sig = SigL + SigU
            lamhat = torch.where(hasU, mu[:, None] / dU, torch.zeros_like(s)) - torch.where(hasL, mu[:, None] / dL, torch.zeros_like(s))
            rhs_x = -(grad + ops.Jt_mul(jv, rows_I(lamhat + Sig * cI)))
            # solve with a per-instance Levenberg shift until the step has positive curvature
            dx = torch.zeros_like(x)
            lamE_new = lamE.clone()
            if it == 0 and self.start_with_resto:
                theta_entry = cE.abs().sum(1) + cI.abs().sum(1)
                resto = (theta_entry > self.tol) & ~inactive
                self.restoration_entries += int(resto.sum())
            need = ~inactive & ~resto
            def next_delta(dv):  # IPOPT's growth of the Hessian perturbation (kappa_plus = 8, first trial 1e-4)
                return torch.clamp(torch.maximum(dv * 8.0, torch.full_like(dv, 1e-4)), max=self.delta_max)

            for attempt in range(12):
                # delta_c only once a plain solve has failed (rank-deficient J_E); the stage-wise sweep always
                # carries it (its pivot blocks are the stage KKT matrices, not the whole one)
                dc = self.delta_c if (attempt > 0 or self.kkt_kind == "stage") else 0.0
                if dev.type == "cuda":
                    torch.cuda.synchronize(dev)
                t_k = time.perf_counter()
                # only the instances that still need a step: converged ones (and those whose step was
                # accepted in an earlier attempt) do not ride along -- the factorisation is what costs
                idx = torch.nonzero(need).ravel()
                # speculative shifts: the LU kernels are latency-bound below one CTA per SM, so while the active
                # instances fill less than half a wave the next S - 1 shifts of the sequence are factored in the
                # SAME sweep and the first one that passes the test is taken -- the shifts tried and the step
                # chosen are those of the one-at-a-time loop, only the number of sweeps drops
                S = max(1, min(4, self.spec_wave // max(1, idx.numel()))) if self.spec_wave else 1
                if self.limited_memory:
                    S = 1  # the L-BFGS matrix is positive definite: shifts are the exception, not worth speculating on
                if S > 1:
                    cand = [delta[idx]]
                    for _ in range(S - 1):
                        cand.append(next_delta(cand[-1]))
                    rep = idx.repeat(S)
                    dl_all = torch.cat(cand)
                    hv_r, jv_r, Sig_r, cI_r = hv[rep], jv[rep], Sig[rep], cI[rep]
                    dxa, lama = backend.solve(hv_r, jv_r, Sig_r, dl_all, dc, rhs_x[rep], -cE[rep])
                    if dev.type == "cuda":
                        torch.cuda.synchronize(dev)
                    self.kkt_seconds += time.perf_counter() - t_k
                    dsa = ops.J_mul(jv_r, dxa)[:, iI] + cI_r
                    curv = ops.W_quad(hv_r, dxa) + dl_all * (dxa * dxa).sum(1) + (Sig_r * dsa * dsa).sum(1)
                    oka = (torch.isfinite(dxa).all(dim=1) & torch.isfinite(lama).all(dim=1)
                           & (curv > 1e-12 * (dxa * dxa).sum(1))).view(S, -1)
                    first = torch.where(oka.any(dim=0), oka.to(torch.int8).argmax(dim=0), torch.full_like(idx, -1))
                    got = first >= 0
                    row = torch.clamp(first, min=0) * idx.numel() + torch.arange(idx.numel(), device=dev)
                    sel = idx[got]
                    dx[sel], lamE_new[sel] = dxa[row[got]], lama[row[got]]
                    delta[sel] = dl_all[row[got]]
                    delta[idx[~got]] = next_delta(cand[-1][~got])
                    need = need.clone()
                    need[sel] = False
                    if not bool(need.any()):
                        break
                    continue
                if self.limited_memory:
                    sub = lm.kkt_solve(backend, idx, hv[idx], jv[idx], Sig[idx], delta[idx], dc, rhs_x[idx], -cE[idx])
                    dxt = torch.zeros_like(x)
                    lamt = torch.zeros_like(lamE)
                    dxt[idx], lamt[idx] = sub[0], sub[1]
                elif idx.numel() == B:
                    dxt, lamt = backend.solve(hv, jv, Sig, delta, dc, rhs_x, -cE)
                else:
                    sub = backend.solve(hv[idx], jv[idx], Sig[idx], delta[idx], dc, rhs_x[idx], -cE[idx])
                    dxt = torch.zeros_like(x)
                    lamt = torch.zeros_like(lamE)
                    dxt[idx], lamt[idx] = sub[0], sub[1]
                if dev.type == "cuda":
                    torch.cuda.synchronize(dev)
                self.kkt_seconds += time.perf_counter() - t_k
                dst = ops.J_mul(jv, dxt)[:, iI] + cI
                # inertia test (IPOPT's criterion): the KKT matrix must have exactly n positive and mE negative
                # eigenvalues, i.e. the reduced Hessian is positive definite on the null space of J_E
                if self.exact_inertia and attempt < 11 and self.kkt_kind == "dense":
                    inertia_ok = torch.zeros(B, dtype=torch.bool, device=dev)
                    inertia_ok[idx] = (torch.linalg.eigvalsh(backend.K) < 0).sum(dim=1) == mE
                else:  # cheap proxy: positive curvature of the barrier Lagrangian along the step
                    if self.limited_memory:
                        allidx = torch.arange(B, device=dev)
                        wq = (dxt * lm.apply(allidx, dxt)).sum(1)
                    else:
                        wq = ops.W_quad(hv, dxt)
                    curv = wq + delta * (dxt * dxt).sum(1) + (Sig * dst * dst).sum(1)
                    inertia_ok = curv > 1e-12 * (dxt * dxt).sum(1)
                ok = torch.isfinite(dxt).all(dim=1) & torch.isfinite(lamt).all(dim=1) & inertia_ok
                take = need & ok
                dx = torch.where(take[:, None], dxt, dx)
                lamE_new = torch.where(take[:, None], lamt, lamE_new)
                need = need & ~ok
                if not bool(need.any()):
                    break
                delta = torch.where(need, next_delta(delta), delta)
            no_step = need & ~inactive  # no shift of the sequence gave a usable step: counts as a failed line search
            # ---- feasibility restoration.  IPOPT switches to a second NLP (min ||c||_1 + zeta ||D(x - x_r)||^2 with
            # its own slacks) when the line search cannot make progress; here the instances in that state take
            # Levenberg-Marquardt steps on the violation instead: the same stage-wise KKT matrix with the Hessian
            # replaced by lambda I and -I in the (2,2) block,
            #   [[lambda I + J_I^T Sigma J_I, J_E^T], [J_E, -I]] [dx; y] = [-J_I^T (lamhat + Sigma c_I); -c_E],
            # i.e. (lambda I + J_E^T J_E + J_I^T Sigma J_I) dx = -J_E^T c_E - J_I^T(...): a damped Gauss-Newton step
            # that ignores the objective; accepted on decrease of the violation alone (below).
            ridx = torch.nonzero(resto & ~inactive).ravel()
            if ridx.numel():
                lam_r = torch.clamp(1e-2 * torch.ones_like(delta[ridx]) * (1.0 + delta[ridx]), max=1e2)
                rhs_r = -ops.Jt_mul(jv[ridx], rows_I(lamhat + Sig * cI)[ridx])
                t_k = time.perf_counter()
                dxr, lamr = backend.solve(torch.zeros_like(hv[ridx]), jv[ridx], Sig[ridx], lam_r, 1.0, rhs_r, -cE[ridx])
                if dev.type == "cuda":
                    torch.cuda.synchronize(dev)
                self.kkt_seconds += time.perf_counter() - t_k
                okr = torch.isfinite(dxr).all(dim=1)
                dx[ridx] = torch.where(okr[:, None], dxr, torch.zeros_like(dxr))
                lamE_new[ridx] = lamE[ridx]  # multipliers are not updated by a feasibility step
                no_step = no_step.clone()
                no_step[ridx] = ~okr
            ds = ops.J_mul(jv, dx)[:, iI] + cI
            lamI_new = lamhat + Sig * ds
            zL_new = torch.where(hasL, mu[:, None] / dL - SigL * ds, torch.zeros_like(s))
            zU_new = torch.where(hasU, mu[:, None] / dU + SigU * ds, torch.zeros_like(s))

            # fraction to the boundary
            def max_step(val, dval, t):
                r = torch.where(dval < 0, -t[:, None] * val / dval, torch.full_like(val, float("inf")))
                return torch.clamp(r.amin(dim=1), max=1.0) if val.shape[1] else torch.ones(B, dtype=torch.float64, device=dev)

            a_p = torch.minimum(max_step(torch.where(hasL, dL, torch.full_like(s, big)), ds, tau),
                                max_step(torch.where(hasU, dU, torch.full_like(s, big)), -ds, tau))
            a_d = torch.minimum(max_step(torch.where(hasL, zL, torch.full_like(s, big)), zL_new - zL, tau),
                                max_step(torch.where(hasU, zU, torch.full_like(s, big)), zU_new - zU, tau))
            # l1 merit function and Armijo backtracking (batched: every trial evaluates all instances)
            cnorm = cE.abs().sum(1) + cI.abs().sum(1)
            lam_all = torch.cat([lamE_new, lamI_new], dim=1)
            nu = torch.maximum(nu, lam_all.abs().amax(dim=1) + 1.0) if lam_all.shape[1] else nu
            dbar = (grad * dx).sum(1) - mu * (torch.where(hasL, ds / dL, torch.zeros_like(s)) - torch.where(hasU, ds / dU, torch.zeros_like(s))).sum(1)
            dphi = dbar - nu * cnorm
            bar0 = barrier(fv, s, mu)
            phi0 = bar0 + nu * cnorm
            if it == 0:
                cnorm0 = cnorm.clone()
            alpha = a_p.clone()
            accepted = inactive.clone()
            x_new, s_new = x.clone(), s.clone()
            blocked = no_step
            ct_acc = cnorm.clone()  # violation of the accepted trial point (restoration bookkeeping)
            for bt in range(self.max_backtrack):
                xt = x + alpha[:, None] * dx
                st = s + alpha[:, None] * ds
                ot = evaluate(xt, full=False)
                n_eval += 1
                ct = (ot["g"][:, iE] - lbE).abs().sum(1) + (ot["g"][:, iI] - st).abs().sum(1)
                phit = barrier(ot["f"], st, mu) + nu * ct
                armijo = phit <= phi0 + self.eta * alpha * torch.minimum(dphi, torch.zeros_like(dphi)) + 1e-13 * phi0.abs()
                # filter-type acceptance against the current iterate (Waechter & Biegler, eq. 18): enough
                # progress in the constraint violation OR in the barrier objective
                bart = barrier(ot["f"], st, mu)
                filt = (ct <= (1.0 - 1e-5) * cnorm) | (bart <= bar0 - 1e-5 * cnorm)
                # f-type step (Waechter & Biegler, eq. 19-20): at an almost feasible iterate with a descent
                # direction for the barrier objective, ask for Armijo decrease of THAT alone, letting the
                # violation grow to a small multiple of theta_min.  Without it the l1 merit function
                # rejects the full step whenever the curved constraints (quaternion, kinematics) give back a
                # second-order violation -- the Maratos effect: 16 evaluations per iteration, tiny steps.
                theta_min = 1e-4 * torch.clamp(cnorm0, min=1.0)
                ftype = self.f_type & (cnorm <= theta_min) & (dbar < 0)
                acc_f = ftype & (bart <= bar0 + self.eta * alpha * dbar) & (ct <= 10.0 * theta_min)
                good = torch.isfinite(phit) & (armijo | (filt & (cnorm > theta_min)) | acc_f)
                # restoration: sufficient decrease of the constraint violation is the only criterion
                good = torch.where(resto, torch.isfinite(ct) & (ct <= (1.0 - 1e-4 * alpha) * cnorm), good)
                take = good & ~accepted & ~blocked
                x_new = torch.where(take[:, None], xt, x_new)
                s_new = torch.where(take[:, None], st, s_new)
                ct_acc = torch.where(take, ct, ct_acc)
                accepted |= take
                if bool((accepted | blocked).all()):
                    break
                alpha = torch.where(accepted, alpha, alpha * 0.5)
            moved = accepted & ~inactive
            moved_prev = moved.clone()
            # instances whose line search failed get a larger Hessian shift and try again; after max_fail
            # failures in a row an instance is given up (it would otherwise cost 16 evaluations and a KKT
            # solve per iteration until max_iter: there is no restoration phase to send it to)
            failed = ~accepted
            fails = torch.where(failed, fails + 1, torch.zeros_like(fails))
            if self.restoration:
                # leave: the violation fell enough (or the instance is feasible to the tolerance again)
                ct_now = ct_acc
                leave = resto & ((ct_now <= self.resto_reduction * theta_entry) | (ct_now <= self.tol))
                if bool(leave.any()):
                    resto = resto & ~leave
                    fails = torch.where(leave, torch.zeros_like(fails), fails)
                    delta = torch.where(leave, torch.zeros_like(delta), delta)
                    lamE = torch.where(leave[:, None], torch.zeros_like(lamE), lamE)  # stale multipliers: start afresh
                    nu = torch.where(leave, torch.ones_like(nu), nu)
                enter = (~resto) & (~inactive) & (fails >= self.resto_after) & (cnorm > self.tol)
                if bool(enter.any()):
                    resto = resto | enter
                    theta_entry = torch.where(enter, cnorm, theta_entry)
                    fails = torch.where(enter, torch.zeros_like(fails), fails)
                    self.restoration_entries += int(enter.sum())
            stalled |= fails >= self.max_fail
            delta = torch.where(failed, torch.clamp(torch.maximum(delta * 8.0, torch.full_like(delta, 1e-4)), max=self.delta_max),
                                torch.where(moved, delta / 3.0, delta))
            delta = torch.where(delta < self.delta_min, torch.zeros_like(delta), delta)
            am = torch.where(moved, alpha, torch.zeros_like(alpha))[:, None]
            ad = torch.where(moved, a_d, torch.zeros_like(a_d))[:, None]
            x, s = x_new.contiguous(), s_new
            lamE = lamE + am * (lamE_new - lamE)
            zL = zL + ad * (zL_new - zL)
            zU = zU + ad * (zU_new - zU)
            # keep the bound multipliers in the IPOPT safeguard band around mu / slack
            dLn, dUn = s - lbs, ubs - s
            ks = 1e10
            zL = torch.where(hasL, torch.minimum(torch.maximum(zL, mu[:, None] / (ks * dLn)), ks * mu[:, None] / dLn), zL)
            zU = torch.where(hasU, torch.minimum(torch.maximum(zU, mu[:, None] / (ks * dUn)), ks * mu[:, None] / dUn), zU)

        lam = zeros_m.clone()
        lam[:, iE] = lamE
        lam[:, iI] = zU - zL
        final = evaluate(x, full=False)
        n_eval += 1
        used = (~done) & (best_it >= 0)  # failed instances with an iterate saved by the callback criterion
        if not bool(done.any()) and not bool(used.any()):
            raise OptiFailure(f"no instance reached tol={self.tol} in {self.max_iter} iterations "
                              f"(best error {err0.min().item():.3e})", callback_used=self.callback_criterion is not None)
        # undo the objective scaling in what is reported (IPOPT: f / s_f, lam_g / s_f)
        cost, lam_out = final["f"] / obj_scale, lam / obj_scale[:, None]
        if bool(used.any()):  # opti_solver.py:478-520: the saved intermediate solution instead of the last iterate
            x = torch.where(used[:, None], best_x, x)
            cost = torch.where(used, best_cost, cost)
            lam_out = torch.where(used[:, None], best_lam / obj_scale[:, None], lam_out)
        return BatchedOutput(values=x, cost_value=cost, constraint_multipliers=lam_out, success=done, iterations=iters,
                             kkt_error=err0, evaluations=n_eval, acceptable=acceptable,
                             callback_iteration=torch.where(used, best_it, torch.full_like(best_it, -1)))
