"""Per-instance counterparts of the reference's callback criteria
(/root/reference/src/hippopt/base/opti_callback.py:24-307) for the batched interior-point driver.

In the reference an `OptiCallback` runs after every IPOPT iteration; when its criterion is satisfied the criterion's
state is updated and the iterate is saved (`SaveBestUnsolvedVariablesCallback.call`, :342-373); if IPOPT then fails,
`OptiSolver.solve()` hands out the saved iterate instead of raising (`opti_solver.py:478-520`).  The kinodynamic
planner uses `BestCost() & AcceptablePrimalInfeasibility(tol)` (`humanoid_kinodynamic/planner.py:57-63`).

Here every criterion keeps one state value per instance; `satisfied(cost, inf_pr)` returns a (B,) mask and
`update(mask, cost, inf_pr)` touches only the instances of the mask -- each instance behaves as the reference's
scalar object would in its own solve.  `cost` is the (unscaled) objective of the iterate, `inf_pr` its constraint
violation (IPOPT's `inf_pr` column, :207/:248)."""
from __future__ import annotations

import torch


class CallbackCriterion:
    def reset(self, batch: int, device) -> None:
        raise NotImplementedError

    def satisfied(self, cost: torch.Tensor, inf_pr: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def update(self, mask: torch.Tensor, cost: torch.Tensor, inf_pr: torch.Tensor) -> None:
        raise NotImplementedError

    def __or__(self, other: "CallbackCriterion") -> "CombinedCallbackCriterion":  # opti_callback.py:44-53
        if not isinstance(other, CallbackCriterion):
            raise TypeError(other)
        return OrCombinedCallbackCriterion(lhs=self, rhs=other)

    def __and__(self, other: "CallbackCriterion") -> "CombinedCallbackCriterion":  # :55-64
        if not isinstance(other, CallbackCriterion):
            raise TypeError(other)
        return AndCombinedCallbackCriterion(lhs=self, rhs=other)


def _inf(batch, device):
    return torch.full((batch,), float("inf"), dtype=torch.float64, device=device)


class BestCost(CallbackCriterion):  # :78-115
    def reset(self, batch, device):
        self.best_cost = _inf(batch, device)

    def satisfied(self, cost, inf_pr):
        return cost < self.best_cost

    def update(self, mask, cost, inf_pr):
        self.best_cost = torch.where(mask, cost, self.best_cost)


class AcceptableCost(CallbackCriterion):  # :118-159
    def __init__(self, acceptable_cost: float = float("inf")):
        self.acceptable_cost = acceptable_cost

    def reset(self, batch, device):
        self.best_acceptable_cost = _inf(batch, device)

    def satisfied(self, cost, inf_pr):
        return cost < self.acceptable_cost

    def update(self, mask, cost, inf_pr):
        self.best_acceptable_cost = torch.where(mask & (cost < self.best_acceptable_cost), cost, self.best_acceptable_cost)


class AcceptablePrimalInfeasibility(CallbackCriterion):  # :162-207
    def __init__(self, acceptable_primal_infeasibility: float = float("inf")):
        self.acceptable_primal_infeasibility = acceptable_primal_infeasibility

    def reset(self, batch, device):
        self.best_acceptable_primal_infeasibility = _inf(batch, device)

    def satisfied(self, cost, inf_pr):
        return inf_pr < self.acceptable_primal_infeasibility

    def update(self, mask, cost, inf_pr):
        self.best_acceptable_primal_infeasibility = torch.where(
            mask & (inf_pr < self.best_acceptable_primal_infeasibility), inf_pr, self.best_acceptable_primal_infeasibility)


class BestPrimalInfeasibility(CallbackCriterion):  # :210-248
    def reset(self, batch, device):
        self.best_primal_infeasibility = _inf(batch, device)

    def satisfied(self, cost, inf_pr):
        return inf_pr < self.best_primal_infeasibility

    def update(self, mask, cost, inf_pr):
        self.best_primal_infeasibility = torch.where(mask, inf_pr, self.best_primal_infeasibility)


class CombinedCallbackCriterion(CallbackCriterion):  # :251-287: reset and update reach both sides
    def __init__(self, lhs: CallbackCriterion, rhs: CallbackCriterion):
        self.lhs, self.rhs = lhs, rhs

    def reset(self, batch, device):
        self.lhs.reset(batch, device)
        self.rhs.reset(batch, device)

    def update(self, mask, cost, inf_pr):
        self.lhs.update(mask, cost, inf_pr)
        self.rhs.update(mask, cost, inf_pr)


class OrCombinedCallbackCriterion(CombinedCallbackCriterion):  # :290-297
    def satisfied(self, cost, inf_pr):
        return self.lhs.satisfied(cost, inf_pr) | self.rhs.satisfied(cost, inf_pr)


class AndCombinedCallbackCriterion(CombinedCallbackCriterion):  # :300-307
    def satisfied(self, cost, inf_pr):
        return self.lhs.satisfied(cost, inf_pr) & self.rhs.satisfied(cost, inf_pr)
