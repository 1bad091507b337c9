"""PPO-Lagrangian minibatch loss of the reference learner (`model.py:78-199`), restated as a pure function.

The one structural change: the reference normalises advantages over the minibatch it is handed (`model.py:106-108`).
When a global minibatch is split over ranks, every rank must normalise with the GLOBAL mean / unbiased std, otherwise
the averaged gradient differs from the single-process one; `normalize_advantages` therefore takes the statistics from
an all-reduce of (count, sum, sum of squares) when a process group is given, and every `mean` of the
loss is expressed as local_sum / global_count so that summing rank gradients reproduces the global-minibatch gradient.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class PPOConfig:
    """TrainingParameters (`alg_parameters.py:53-91`)."""
    lr: float = 1e-5
    gamma: float = 0.95
    lam: float = 0.95
    clip_range: float = 0.2
    max_grad_norm: float = 10.0
    entropy_coef: float = 0.01
    value_coef: float = 0.08
    valid_coef: float = 0.5
    cost_value_coef: float = 0.0
    cost_coef: float = 0.0
    cost_limit_per_agent: float = 5.0
    n_epochs: int = 10
    n_steps: int = 256
    minibatch_size: int = 256
    minus_adv_with_cadv: bool = True
    lagrangian_type: int = 0


def _global_stats(x: torch.Tensor, group) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(count, mean, unbiased std) of x over all ranks of `group` (or locally when group is None)."""
    xd = x.double()
    s = torch.stack([torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device), xd.sum(), (xd * xd).sum()])
    if group is not None:
        dist.all_reduce(s, group=group)
    n, mean = s[0], s[1] / s[0]
    var = (s[2] - n * mean * mean) / (n - 1)
    return n, mean, var.clamp_min(0).sqrt()


def normalize_advantages(x: torch.Tensor, group=None) -> torch.Tensor:
    """`(x - x.mean()) / (x.std() + 1e-6)` (`model.py:106`), statistics over the global minibatch."""
    if group is None:
        return (x - x.mean()) / (x.std() + 1e-6)
    _, mean, std = _global_stats(x, group)
    return (x - mean.to(x.dtype)) / (std.to(x.dtype) + 1e-6)


def ppo_lagrange_loss(out, *, returns, cost_returns, old_v, old_cv, actions, old_ps, train_valid, lagrangian: float,
                      cfg: PPOConfig, group=None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """out: PolicyOutput for the minibatch; every other tensor is [B, N] (or [B, N, 5]).  Returns (loss, stats).

    With `group`, `loss` is this rank's SHARE of the global-minibatch loss: summing the ranks' losses (and hence
    gradients) gives `model.py:153-158` evaluated on the concatenated minibatch."""
    clip = cfg.clip_range
    adv = normalize_advantages(returns - old_v, group)
    cadv = normalize_advantages(cost_returns - old_cv, group)
    if cfg.minus_adv_with_cadv:                                              # model.py:111-113
        adv = (adv - lagrangian * cadv) / (lagrangian + 1.0)

    n_local = float(returns.numel())
    n_global = n_local
    if group is not None:
        t = torch.tensor(n_local, dtype=torch.float64, device=returns.device)
        dist.all_reduce(t, group=group)
        n_global = float(t)

    def gmean(x, per_row=1):            # mean over the global minibatch, expressed as a local share
        return x.sum() / (n_global * per_row)

    a = actions.long().unsqueeze(-1)
    new_p = out.policy.gather(-1, a).squeeze(-1)
    old_p = old_ps.gather(-1, a).squeeze(-1)
    ratio = torch.exp(torch.log(new_p.clamp(1e-6, 1.0)) - torch.log(old_p.clamp(1e-6, 1.0)))     # model.py:119
    entropy = gmean(-(out.policy * torch.log(out.policy.clamp(1e-6, 1.0))).sum(-1))             # :121

    def clipped_value_loss(new, old, target):                                                    # :124-136
        new = new.squeeze(-1)
        clipped = old + (new - old).clamp(-clip, clip)
        return gmean(torch.maximum((new - target) ** 2, (clipped - target) ** 2))
    critic_loss = clipped_value_loss(out.value, old_v, returns)
    cost_critic_loss = clipped_value_loss(out.cost_value, old_cv, cost_returns)

    policy_loss = gmean(torch.minimum(adv * ratio, adv * ratio.clamp(1.0 - clip, 1.0 + clip)))   # :139-143
    sig = out.policy_sig
    valid_loss = -gmean(torch.log(sig.clamp(1e-6, 1.0 - 1e-6)) * train_valid +
                        torch.log((1 - sig).clamp(1e-6, 1.0 - 1e-6)) * (1 - train_valid), per_row=sig.shape[-1])  # :146-148
    cost_loss = gmean(ratio * cadv)                                                              # :155
    loss = (-policy_loss - entropy * cfg.entropy_coef + cfg.value_coef * critic_loss + cfg.valid_coef * valid_loss
            + cfg.cost_value_coef * cost_critic_loss + cfg.cost_coef * lagrangian * cost_loss)   # :159-164
    with torch.no_grad():
        stats = dict(all_loss=loss, policy_loss=policy_loss, policy_entropy=entropy, critic_loss=critic_loss,
                     valid_loss=valid_loss, cost_critic_loss=cost_critic_loss, cost_loss=cost_loss,
                     clipfrac=gmean(((ratio - 1.0).abs() > clip).float()), advantage=gmean(adv),
                     cost_advantage=gmean(cadv))
        stats = {k: v.detach() for k, v in stats.items()}
    return loss, stats


def reduce_stats(stats: Dict[str, torch.Tensor], group=None) -> Dict[str, float]:
    """Sums the per-rank shares into the global-minibatch statistics."""
    keys = sorted(stats)
    t = torch.stack([stats[k].double() for k in keys])
    if group is not None:
        dist.all_reduce(t, group=group)
    return {k: float(v) for k, v in zip(keys, t)}
