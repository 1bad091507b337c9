#!/usr/bin/env python
"""Generates tests/golden/ppo_*.npz by running the UNMODIFIED reference learner (model.py / net.py / lagrange.py) on
CPU in this container.  Run from the repo root:  python tests/golden/make_ppo_golden.py

What is recorded (north_star: "the PPO loss and gradients must match within a stated fp32 tolerance"):
  * inputs of one `Model.train` minibatch (observations, vectors, returns, cost returns, old values, actions, old
    policies, trainValid, episode cost) and of one forward pass;
  * the reference network's outputs for those inputs (policy, value, blocking, policy_sig, features, logits, cost
    value), `net.py:101-155`, dropout disabled with `.eval()` (the reference never calls it; with dropout active the
    outputs are random and cannot be compared);
  * `Model.train`'s stats list (`model.py:186-199`): all_loss, policy_loss, entropy, critic_loss, valid_loss,
    cost_critic_loss, cost_loss, clipfrac, grad_norm, mean advantage, mean cost advantage, lagrangian;
  * the gradient of every parameter BEFORE clipping, summarised as (L2 norm, the first 24 and 24 strided elements), and
    the same elements of the parameter after the Adam step;
  * Lagrangian / PID-Lagrangian multiplier sequences for a cost series (`lagrange.py:27-88`).

The 8.2 M parameters themselves are not committed: both sides fill them with `fill_value(name, shape)` below (NumPy
PCG64 keyed by the parameter's reference name), so the fixture stays small.
"""
import os
import sys
import warnings
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

N_ELEMS = 24


def fill_value(name: str, shape, seed: int = 2024) -> np.ndarray:
    """Deterministic parameter values keyed by the reference parameter name (independent of construction order)."""
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    shape = tuple(int(s) for s in shape)
    if len(shape) >= 2:
        fan_in = int(np.prod(shape[1:])) if len(shape) != 3 else shape[1]
        return (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
    if name.endswith("norm.weight"):
        return (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
    return (0.02 * rng.standard_normal(shape)).astype(np.float32)


def sample_indices(numel: int) -> np.ndarray:
    first = np.arange(min(N_ELEMS, numel))
    strided = (np.arange(N_ELEMS) * max(1, numel // N_ELEMS)) % numel
    return np.concatenate([first, strided]).astype(np.int64)


def make_inputs(rng, B, N, C=6, F=9):
    obs = (rng.random((B, N, C, F, F)) < 0.15).astype(np.float32)
    vec = rng.standard_normal((B, N, 4)).astype(np.float32)
    vec[..., 3] = 0.0
    logits = rng.standard_normal((B, N, 5)).astype(np.float32)
    old_ps = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    return dict(obs=obs, vec=vec, returns=rng.standard_normal((B, N)).astype(np.float32),
                cost_returns=rng.random((B, N)).astype(np.float32),
                old_v=rng.standard_normal((B, N)).astype(np.float32),
                old_cv=rng.random((B, N)).astype(np.float32),
                actions=rng.integers(0, 5, size=(B, N)).astype(np.int64), old_ps=old_ps.astype(np.float32),
                train_valid=(rng.random((B, N, 5)) < 0.6).astype(np.float32),
                hidden=np.zeros((B, 2, N, 512), dtype=np.float32), episode_cost=np.float64(37.5))


def run_case(name, *, B, N, seed, cost_value_coef=None, cost_coef=None, minus_adv=None, lagrangian_type=None):
    import torch
    from ref_loader import load_reference
    _, _, AP = load_reference(N)
    if cost_value_coef is not None:
        AP.TrainingParameters.COST_VALUE_COEF = cost_value_coef
    if cost_coef is not None:
        AP.TrainingParameters.COST_COEF = cost_coef
    if minus_adv is not None:
        AP.TrainingParameters.MINUS_ADV_WITH_CADV = minus_adv
    if lagrangian_type is not None:
        AP.LagrangianParameters.LAGRANGIAN_TYPE = lagrangian_type
    import model as ref_model
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ref_model.Model(0, torch.device("cpu"), True)
    net = m.network
    names = []
    with torch.no_grad():
        for pname, p in net.named_parameters():
            p.copy_(torch.from_numpy(fill_value(pname, p.shape)))
            names.append(pname)
    net.eval()                                          # dropout off: see the module docstring
    rng = np.random.default_rng(seed)
    d = make_inputs(rng, B, N)

    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = net(torch.from_numpy(d["obs"]), torch.from_numpy(d["vec"]), None)
    fwd = {k: o.numpy().copy() for k, o in zip(("policy", "value", "blocking", "policy_sig", "features", "logits",
                                                 "cost_value"), out)}

    grads = {}
    orig_clip = torch.nn.utils.clip_grad_norm_

    def recording_clip(params, max_norm, *a, **k):
        params = list(params)
        for pname, p in net.named_parameters():
            grads[pname] = None if p.grad is None else p.grad.detach().clone()
        return orig_clip(params, max_norm, *a, **k)

    torch.nn.utils.clip_grad_norm_ = recording_clip
    lag_before = m.lagrange.get_lagrangian_param()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            stats = m.train(d["obs"], d["vec"], d["returns"], d["cost_returns"], d["old_v"], d["old_cv"], d["actions"],
                            d["old_ps"], d["hidden"], d["train_valid"], float(d["episode_cost"]))
    finally:
        torch.nn.utils.clip_grad_norm_ = orig_clip
    lag_after = m.lagrange.get_lagrangian_param()

    P = len(names)
    gnorm = np.zeros(P, dtype=np.float64)
    gelem = np.zeros((P, 2 * N_ELEMS), dtype=np.float32)
    pelem = np.zeros((P, 2 * N_ELEMS), dtype=np.float32)
    for k, (pname, p) in enumerate(net.named_parameters()):
        idx = sample_indices(p.numel())
        g = grads[pname]
        if g is not None:
            gnorm[k] = float(g.double().norm())
            gelem[k, :len(idx)] = g.flatten()[idx].numpy()
        pelem[k, :len(idx)] = p.detach().flatten()[idx].numpy()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), param_names=np.array(names), grad_norm=gnorm, grad_elems=gelem,
        param_elems_after_step=pelem, stats=np.array([float(s) for s in stats], dtype=np.float64),
        stat_names=np.array(AP.RecordingParameters.LOSS_NAME), lagrangian_before=np.float64(lag_before),
        lagrangian_after=np.float64(lag_after), n_agents=np.int32(N),
        cfg=np.array([AP.TrainingParameters.COST_VALUE_COEF, AP.TrainingParameters.COST_COEF,
                      float(AP.TrainingParameters.MINUS_ADV_WITH_CADV), float(AP.LagrangianParameters.LAGRANGIAN_TYPE),
                      AP.TrainingParameters.lr, AP.TrainingParameters.COST_LIMIT_PER_AGENT], dtype=np.float64),
        **{"in_" + k: v for k, v in d.items() if k != "hidden"}, **{"fwd_" + k: v for k, v in fwd.items()})
    print(name, "stats:", dict(zip(AP.RecordingParameters.LOSS_NAME, [round(float(s), 6) for s in stats])))


def make_lagrange_golden():
    from ref_loader import load_reference
    _, _, AP = load_reference(2)
    import lagrange
    costs = np.array([7.0, 9.5, 3.0, 1.0, 12.0, 6.0, 5.0, 4.5, 8.0, 0.5, 0.0, 15.0], dtype=np.float64)
    seqs = {}
    for typ, key in ((0, "vanilla"), (1, "pid")):
        lg = lagrange.get_lagrangian(lagrange.LagrangianType(typ), AP.TrainingParameters.COST_LIMIT_PER_AGENT)
        vals = [lg.get_lagrangian_param()]
        for c in costs:
            lg.update_lagrangian_multiplier(float(c))
            vals.append(lg.get_lagrangian_param())
        seqs[key] = np.array(vals, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ppo_lagrange.npz"), costs=costs,
                        cost_limit=np.float64(AP.TrainingParameters.COST_LIMIT_PER_AGENT), **seqs)
    print("ppo_lagrange:", {k: v[-1] for k, v in seqs.items()})


if __name__ == "__main__":
    import subprocess
    if len(sys.argv) > 1:                      # one case per process: the reference's config classes are process-global
        which = sys.argv[1]
        if which == "default":
            run_case("ppo_default_n3", B=8, N=3, seed=1)
        elif which == "cost":
            run_case("ppo_costterms_n2", B=6, N=2, seed=2, cost_value_coef=0.05, cost_coef=0.3, minus_adv=False)
        elif which == "pid":
            run_case("ppo_pid_n4", B=4, N=4, seed=3, cost_value_coef=0.02, cost_coef=0.1, lagrangian_type=1)
        elif which == "lagrange":
            make_lagrange_golden()
    else:
        for which in ("default", "cost", "pid", "lagrange"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), which])
