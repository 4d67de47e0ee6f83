"""oracle/opcount.py -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Algorithmic operation counts of the NLP functions, per knot, from the oracle's tapes -- the stand-in
SURVEY.md 8(d) prescribes for CasADi's ``Function.n_instructions()`` when CasADi is not available:

  values    : instructions of the f and g tapes;
  Jacobian  : instructions a symbolic forward-mode sweep emits (what ``nlp_jac_g`` would contain):
              for every instruction, one derivative instruction group per decision variable it
              structurally depends on (zero tangents are pruned at graph construction, as SX does);
  Hessian   : the reverse-mode gradient-of-the-Lagrangian tape plus the same forward-mode count on it
              (forward-over-reverse, what ``nlp_hess_l`` contains up to colouring/CSE).

Unary/binary instructions count as 1 flop; the per-direction tangent cost is 1 for +,-, 3 for a product
of two dependent operands (1 if only one operand depends on x), 4 for a quotient, 2 for unary functions.
"""
from __future__ import annotations

import numpy as np

from . import sx


def _popcount(m: int) -> int:
    return bin(m).count("1")


def forward_mode_ops(outputs, inputs, seed_idx) -> int:
    """Flops of a pruned symbolic forward sweep of ``outputs`` w.r.t. inputs[seed_idx]."""
    pos = {int(i): d for d, i in enumerate(seed_idx)}
    index = {s.id: i for i, s in enumerate(inputs)}
    bits: dict[int, int] = {}
    total = 0
    for node in sx.topo(outputs):
        if node.op == sx.OP_SYM:
            i = index.get(node.id)
            bits[node.id] = (1 << pos[i]) if i in pos else 0
            continue
        if node.op == sx.OP_CONST:
            bits[node.id] = 0
            continue
        ma = bits[node.a.id]
        mb = bits[node.b.id] if node.b is not None else 0
        m = ma | mb
        bits[node.id] = m
        if not m:
            continue
        if node.op in (sx.OP_ADD, sx.OP_SUB):
            total += _popcount(ma & mb) + 0 * _popcount(m)
            total += 0 if node.op == sx.OP_ADD else _popcount(mb & ~ma)  # negation of a lone rhs tangent
        elif node.op == sx.OP_MUL:
            total += 3 * _popcount(ma & mb) + _popcount(ma ^ mb)
        elif node.op == sx.OP_DIV:
            total += 4 * _popcount(m)
        elif node.op == sx.OP_NEG:
            total += _popcount(m)
        else:
            total += 2 + _popcount(m)  # derivative factor once, then one product per direction
    return total


def per_knot_counts(nlp, knot_of_app, n_x, interior_knot: int) -> dict:
    """Counts attributable to one interior knot: every row/cost application whose *last* bound
    variable lies in that knot (defects of interval k-1 -> k belong to knot k)."""
    vals = jac = hess = 0
    apps = [(t, b, False) for (t, b, _) in nlp.row_apps] + [(t, b, True) for (t, b, _) in nlp.cost_apps]
    seen = {}
    by_template: dict[str, int] = {}
    for (t, binding, is_cost) in apps:
        if knot_of_app(binding) != interior_knot:
            continue
        key = (id(t), is_cost)
        if key not in seen:
            xin = [i for i in range(len(t.inputs)) if binding[i] < n_x]
            v = t.tape.n_ops
            j = 0 if is_cost else forward_mode_ops(t.rows, t.inputs, xin)
            lam, g, gtape, pat = t.lagrangian_gradient()
            h = gtape.n_ops + forward_mode_ops(g, t.inputs + lam, xin)
            seen[key] = (v, j, h)
        v, j, h = seen[key]
        vals += v
        jac += j
        hess += h
        by_template[t.name] = by_template.get(t.name, 0) + v + j + h
    return {"values": vals, "jacobian": jac, "hessian": hess, "total": vals + jac + hess, "by_template": by_template}


KINEMATICS_TEMPLATES = (
    "fk_", "com_kinematics_consistency", "centroidal_momentum_kinematics_consistency", "minimum_feet_distance",
    "unitary_quaternion", "frame_quaternion_error", "base_quaternion_error", "base_quaternion_velocity_error",
    "joint_positions_error",
)


def kinematics_share(counts: dict) -> int:
    """Flops of the templates evaluated by kino_kin_kernel (the rest belongs to kino_contact_kernel)."""
    return sum(v for k, v in counts["by_template"].items() if k.startswith(KINEMATICS_TEMPLATES))


def kinodynamic_counts(model, horizon: int = 30, **settings) -> dict:
    from . import kinodynamic as kd

    nlp, lay = kd.build(model, kd.Settings(horizon=horizon, **settings))

    def knot_of(binding):
        xs = binding[binding < lay.h_init]
        return int(xs.max() // kd.NZ) if len(xs) else -1

    out = per_knot_counts(nlp, knot_of, nlp.n_x, horizon // 2)
    out["horizon"] = horizon
    return out
