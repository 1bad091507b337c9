"""Lagrange multipliers of the PPO-Lagrangian update (`lagrange.py:27-88`, parameters `alg_parameters.py:94-108`)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class LagrangianConfig:
    init_value: float = 1.0      # LagrangianParameters.INIT_VALUE
    upper_bound: float = 20.0    # UPPER_BOUND
    lr: float = 5e-2             # LR (vanilla)
    kp: float = 0.1              # KP, KI, KD (PID)
    ki: float = 0.01
    kd: float = 0.01
    cost_moving_avg_alpha: float = 0.95
    delta_moving_avg_alpha: float = 0.95


class Lagrangian:
    """lambda = softplus(theta), theta trained by Adam on  -theta * (ep_cost - limit), clamped to [0, upper]
    (`lagrange.py:27-53`)."""

    def __init__(self, cost_limit: float, cfg: LagrangianConfig = LagrangianConfig()):
        self.cost_limit, self.cfg = float(cost_limit), cfg
        self.theta = torch.tensor(max(0.0, cfg.init_value), dtype=torch.float32, requires_grad=True)
        self.opt = torch.optim.Adam([self.theta], lr=cfg.lr)

    def value(self) -> float:
        return float(F.softplus(self.theta).detach())

    def update(self, ep_cost_avg: float) -> None:
        self.opt.zero_grad()
        (-self.theta * (float(ep_cost_avg) - self.cost_limit)).backward()
        self.opt.step()
        with torch.no_grad():
            self.theta.clamp_(0.0, self.cfg.upper_bound)

    def state_dict(self):
        return {"theta": self.theta.detach().clone(), "opt": self.opt.state_dict()}

    def load_state_dict(self, sd):
        with torch.no_grad():
            self.theta.copy_(sd["theta"])
        self.opt.load_state_dict(sd["opt"])


class PIDLagrangian:
    """PID controller on the cost excess with exponentially smoothed P and D terms (`lagrange.py:55-88`)."""

    def __init__(self, cost_limit: float, cfg: LagrangianConfig = LagrangianConfig()):
        self.cost_limit, self.cfg = float(cost_limit), cfg
        self.i_term = max(0.0, cfg.init_value)
        self.lam = 0.0
        self.delta_avg = 0.0
        self.cost_avg = 0.0
        self.cost_avg_prev = 0.0

    def value(self) -> float:
        return self.lam

    def update(self, ep_cost_avg: float) -> None:
        c = self.cfg
        delta = float(ep_cost_avg) - self.cost_limit
        self.delta_avg *= c.delta_moving_avg_alpha
        self.delta_avg += (1 - c.delta_moving_avg_alpha) * delta
        self.cost_avg *= c.cost_moving_avg_alpha
        self.cost_avg += (1 - c.cost_moving_avg_alpha) * float(ep_cost_avg)
        d_term = max(0.0, self.cost_avg - self.cost_avg_prev)
        self.i_term = max(0.0, self.i_term + delta * c.ki)
        self.lam = max(0.0, c.kp * self.delta_avg + self.i_term + c.kd * d_term)
        self.cost_avg_prev = self.cost_avg

    def state_dict(self):
        return {k: getattr(self, k) for k in ("i_term", "lam", "delta_avg", "cost_avg", "cost_avg_prev")}

    def load_state_dict(self, sd):
        for k, v in sd.items():
            setattr(self, k, float(v))


def make_lagrangian(kind: int, cost_limit: float, cfg: LagrangianConfig = LagrangianConfig()):
    """`lagrange.get_lagrangian` (`lagrange.py:21-25`): 0 = vanilla, 1 = PID."""
    if kind == 0:
        return Lagrangian(cost_limit, cfg)
    if kind == 1:
        return PIDLagrangian(cost_limit, cfg)
    raise ValueError(f"unknown lagrangian type {kind}")
