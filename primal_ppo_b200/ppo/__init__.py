"""Learner side of the rollout loop (BASELINE.json configs[3]): the policy network and the PPO-Lagrangian update stay in
PyTorch (the only dense contraction of the reference, `net.py` / `model.py`); everything around them — the vector env,
action sampling, rollout buffer, GAE — runs on the hand-written kernels behind `include/mapf_b200.h`."""
from .lagrange import Lagrangian, PIDLagrangian, make_lagrangian  # noqa: F401
from .loss import PPOConfig, normalize_advantages, ppo_lagrange_loss  # noqa: F401
from .policy import REFERENCE_KEY_MAP, ScrimpPolicy  # noqa: F401
from .trainer import RolloutBuffer, VecPPOTrainer  # noqa: F401
from .evaluate import evaluate_fixed_episodes  # noqa: F401,E402
