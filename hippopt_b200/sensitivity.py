"""Derivative of the solution with respect to parameters: the device-side counterpart of the reference's
differentiable-solver path.

`OptiSolver.to_function` (`/root/reference/src/hippopt/base/opti_solver.py:597-638`) turns a whole solve into a
`cs.Function` so that `Function.jacobian()` yields d(solution) / d(parameter)
(`turnkey_planners/humanoid_pose_finder/main_sensitivity.py:213-247`).  What CasADi does there is the implicit-function
theorem on the optimality conditions of the solved problem; the same linear system is what this library already factors
at every iteration, so the derivative is one more KKT solve with one right-hand side per parameter:

    [ W + J_I^T Sigma J_I    J_E^T ] [ dx     ]     [ d_p grad_x L  +  J_I^T (Sigma d_p g_I - Sigma_L d_p lb_I - Sigma_U d_p ub_I) ]
    [ J_E                    0     ] [ dlam_E ] = - [ d_p g_E - d_p b_E                                                           ]

with W = hess_l(x*, lam*), Sigma = z_L / (g_I - lb) + z_U / (ub - g_I) from the multipliers of the inequality rows
(lam_I = z_U - z_L, IPOPT's sign) -- large on active rows, which the system then treats as equalities, negligible on
inactive ones -- and the parameter derivatives of grad_x L = grad_f + J^T lam, of g and of the bounds taken by central
differences of BATCHED evaluations (2 per parameter and instance, all in one `eval` call; the bounds are affine in p).
The reference's own use differentiates with respect to link lengths of the parametric robot model, which this library
does not evaluate; every parameter of the parameter vector p of the three template problems is
available instead (references, gains, friction, limits, initial / final states, ...).
"""
from __future__ import annotations

import numpy as np
import torch

from .evaluator import F, G, GRAD_F, HESS_L, JAC_G
from .ipsolver import DenseKKT, SparseOps


def solution_sensitivity(ev, x, lam_g, p, bounds, param_idx, kkt: str = "dense", rel_step: float = 1e-6,
                         delta_c: float = 0.0, sigma_cap: float = 1e8, delta: float = 0.0):
    """d x* / d p[:, param_idx] at a solved batch.

    ev          evaluator (`eval`, `jac_sparsity`, `hess_sparsity`, `n_x`, `m`); x (B, n_x), lam_g (B, m), p (B, n_p)
                the solution, its multipliers (IPOPT's sign) and the parameters, tensors on the evaluator's device
    bounds      callable p (numpy, (B', n_p)) -> (lbg, ubg) numpy arrays (B', m) or (m,)
    param_idx   the k parameters to differentiate with respect to
    kkt         "dense" (one (n_x + m_E)-square solve per instance) or "stage" (block-tridiagonal sweep, CUDA)
    sigma_cap   Sigma is capped at sigma_cap x (largest |hess_l| entry of the instance): an active row of a converged
                interior-point solve has z / slack ~ 1e11, which enforces the row to 1e-11 but costs the linear solve as
                many digits; at 1e8 both errors are ~1e-8 of the solution's derivative
    delta       shift added to the Hessian block (and delta_c subtracted on the multiplier block), as the solver's inertia
                correction does: for problems whose solution is not locally unique (redundant equality rows, directions
                the objective does not see -- the kinodynamic OCPs have both) the unshifted system is singular and the
                derivative is defined only up to that regularisation
    Returns (dx (B, n_x, k), dlam_E (B, m_E, k), eq_rows)."""
    param_idx = [int(j) for j in np.atleast_1d(param_idx)]
    k = len(param_idx)
    if k == 0:
        raise ValueError("param_idx is empty")
    B, n, m = x.shape[0], ev.n_x, ev.m
    dev = x.device
    p_np = p.detach().cpu().numpy()
    lbg, ubg = (np.broadcast_to(np.asarray(a, dtype=np.float64), (B, m)) for a in bounds(p_np))
    eq = lbg[0] == ubg[0]
    live = ~(np.isinf(lbg[0]) & np.isinf(ubg[0]))
    iE_np, iI_np = np.nonzero(eq)[0], np.nonzero(~eq & live)[0]
    iE, iI = torch.as_tensor(iE_np, device=dev), torch.as_tensor(iI_np, device=dev)
    ops = SparseOps(n, m, ev.jac_sparsity(), ev.hess_sparsity(), dev)
    ones = torch.ones(B, dtype=torch.float64, device=dev)
    base = ev.eval(G | JAC_G | HESS_L, x.contiguous(), p.contiguous(), lam_g.contiguous(), ones)
    base = {key: v.clone() for key, v in base.items()}
    # barrier diagonal of the inequality rows from the multipliers and the distances to the bounds
    gI = base["g"][:, iI]
    lb = torch.as_tensor(lbg[:, iI_np], device=dev)
    ub = torch.as_tensor(ubg[:, iI_np], device=dev)
    lI = lam_g[:, iI]
    cap = (sigma_cap * torch.clamp(base["hess"].abs().amax(dim=1), min=1.0))[:, None]
    tiny = 1e-300
    sigL = torch.where(torch.isfinite(lb), torch.clamp(-lI, min=0.0) / torch.clamp(gI - lb, min=tiny), torch.zeros_like(gI))
    sigU = torch.where(torch.isfinite(ub), torch.clamp(lI, min=0.0) / torch.clamp(ub - gI, min=tiny), torch.zeros_like(gI))
    sigL, sigU = torch.minimum(sigL, cap), torch.minimum(sigU, cap)
    sig = sigL + sigU
    # parameter derivatives by central differences, every (instance, parameter, sign) in ONE batched evaluation
    steps = np.array([rel_step * max(1.0, float(np.abs(p_np[:, j]).max())) for j in param_idx])
    P_pert = np.repeat(p_np[:, None, None, :], k, axis=1).repeat(2, axis=2)  # (B, k, 2, n_p)
    for c, j in enumerate(param_idx):
        P_pert[:, c, 0, j] += steps[c]
        P_pert[:, c, 1, j] -= steps[c]
    P_flat = P_pert.reshape(B * k * 2, -1)
    rep = torch.arange(B, device=dev).repeat_interleave(2 * k)
    zeros_m = torch.zeros((B * k * 2, m), dtype=torch.float64, device=dev)
    pert = ev.eval(F | GRAD_F | G | JAC_G, x[rep].contiguous(), torch.as_tensor(P_flat, device=dev).contiguous(), zeros_m,
                   torch.ones(B * k * 2, dtype=torch.float64, device=dev))
    gradL = pert["grad_f"] + ops.Jt_mul(pert["jac"], lam_g[rep].contiguous())
    gradL = gradL.view(B, k, 2, n)
    gp = pert["g"].view(B, k, 2, m)
    h = torch.as_tensor(steps, device=dev)[None, :, None]
    d_gradL = (gradL[:, :, 0, :] - gradL[:, :, 1, :]) / (2.0 * h)        # (B, k, n)
    d_g = (gp[:, :, 0, :] - gp[:, :, 1, :]) / (2.0 * h)                  # (B, k, m)
    lbp, ubp = (np.broadcast_to(np.asarray(a, dtype=np.float64), (B * k * 2, m)).reshape(B, k, 2, m) for a in bounds(P_flat))
    with np.errstate(invalid="ignore"):  # inf - inf on rows without that bound: masked below
        d_lb = (lbp[:, :, 0, :] - lbp[:, :, 1, :]) / (2.0 * steps[None, :, None])
        d_ub = (ubp[:, :, 0, :] - ubp[:, :, 1, :]) / (2.0 * steps[None, :, None])
    d_lb = torch.as_tensor(np.nan_to_num(d_lb, nan=0.0, posinf=0.0, neginf=0.0), device=dev)
    d_ub = torch.as_tensor(np.nan_to_num(d_ub, nan=0.0, posinf=0.0, neginf=0.0), device=dev)
    # right-hand sides
    tI = sig[:, None, :] * d_g[:, :, iI] - sigL[:, None, :] * d_lb[:, :, iI] - sigU[:, None, :] * d_ub[:, :, iI]  # (B, k, mI)
    lam_like = torch.zeros((B * k, m), dtype=torch.float64, device=dev)
    lam_like[:, iI] = tI.reshape(B * k, -1)
    jrep = torch.arange(B, device=dev).repeat_interleave(k)
    JtT = ops.Jt_mul(base["jac"][jrep].contiguous(), lam_like).view(B, k, n)
    rhs_x = -(d_gradL + JtT).transpose(1, 2).contiguous()                          # (B, n, k)
    rhs_E = -(d_g[:, :, iE] - d_lb[:, :, iE]).transpose(1, 2).contiguous()         # (B, mE, k): lb = ub on these rows
    delta = torch.full((B,), float(delta), dtype=torch.float64, device=dev)
    if kkt == "dense":
        backend = DenseKKT(ops, iE, iI)
    elif kkt == "stage":
        from .kkt import StageKKT

        lay = ev.layout
        jc_, jr_ = ev.jac_sparsity()
        hc_, hr_ = ev.hess_sparsity()
        backend = StageKKT(n, m, lay.N, lay.knot_size, jc_, jr_, hc_, hr_, iE_np, iI_np, device=dev)
        delta_c = delta_c or 1e-9  # the stage blocks are the stage KKT matrices: keep their (2,2) block regular
    else:
        raise ValueError("kkt must be 'dense' or 'stage'")
    dx, dlE = backend.solve(base["hess"], base["jac"], sig, delta, delta_c, rhs_x, rhs_E)
    return dx, dlE, iE_np
