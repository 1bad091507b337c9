"""Batched evaluation on fixed episodes — `evaluate.evaluate` (`evaluate.py:168-316`) with every episode running as one
world of the GPU vector env instead of one Python env after the other."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from ..vec_env import BatchedMapfGym, sample_actions
from .policy import ScrimpPolicy


@torch.no_grad()
def evaluate_fixed_episodes(policy: ScrimpPolicy, scenario, *, max_steps: int = 256, greedy: bool = False,
                            model_name: str = "model", seed: int = 1234, device=None, amp_dtype=None,
                            frames_for_world: Optional[int] = None) -> Dict:
    """Runs `max_steps` steps of all episodes (`EvalParameters.MAX_STEPS`, `alg_parameters.py:7`) and returns
    {"per_episode": {...arrays [W]...}, "metrics": the reference's `all_metrics` keys for this model
    (`evaluate.py:297-312`), "frames": list of uint8 frames of one world if requested}.
    Actions: argmax of the policy when `greedy`, else sampled on device (`Model.evaluate`, `model.py:43-62`)."""
    env = BatchedMapfGym(scenario, device=device, seed=seed)
    W, N = env.W, env.N
    obs, vec = env.getAllObservations()
    reward_sum = torch.zeros(W, device=env.device, dtype=torch.float64)
    cost_sum = torch.zeros(W, device=env.device, dtype=torch.float64)
    frames = []

    def frame():
        if frames_for_world is None:
            return
        from ..episode_io import render_world
        w = frames_for_world
        st = env.state()
        sc = scenario.to_host() if hasattr(scenario, "to_host") else scenario
        tick = int(env_tick[0]) % int(sc.hlen[w])
        frames.append(render_world(sc.obst[w], st["pos"][w].cpu().numpy(), st["goal"][w].cpu().numpy(),
                                   tuple(sc.htrace[w, tick, :2])))
    env_tick = [0]
    frame()
    for t in range(max_steps):
        with torch.autocast(env.device.type, dtype=amp_dtype, enabled=amp_dtype is not None):
            out = policy(obs, vec)
        ps = out.policy.float()
        actions = ps.argmax(dim=-1).to(torch.int8) if greedy else sample_actions(ps.contiguous(), seed=seed, draw=t)
        so, obs, vec = env.step_observe(actions)
        reward_sum += so.reward.sum(dim=1, dtype=torch.float64)
        cost_sum += so.cost.sum(dim=1, dtype=torch.float64)
        env_tick[0] += 1
        frame()
    c = env.counters().cpu().numpy()          # totalGoals, shadowGoals, staticCollide, humanCollide, agentCollide, violations
    per = dict(episodeReward=reward_sum.cpu().numpy(), episodeCostReward=cost_sum.cpu().numpy(), totalGoals=c[:, 0],
               shadowGoals=c[:, 1], staticCollide=c[:, 2], humanCollide=c[:, 3], agentCollide=c[:, 4],
               constraintViolations=c[:, 5], err=env.state()["err"].cpu().numpy())
    metrics = {}
    for key, val in (("hc", per["humanCollide"]), ("cv", per["constraintViolations"]), ("ecr", per["episodeCostReward"]),
                     ("goals", per["totalGoals"])):
        v = np.asarray(val, dtype=np.float64)
        mean_pa, std_pa = v.mean() / N, v.std() / N
        metrics[f"{model_name}/{key}_per_agent/mean"] = float(mean_pa)
        metrics[f"{model_name}/{key}_per_agent/std"] = float(std_pa)
        metrics[f"{model_name}/{key}_per_agent_per_timestep/mean"] = float(mean_pa / max_steps)
        metrics[f"{model_name}/{key}_per_agent_per_timestep/std"] = float(std_pa / max_steps)
    return {"per_episode": per, "metrics": metrics, "frames": frames}
