"""GPU tests of the policy-side glue and the rollout/update loop (BASELINE.json configs[3])."""
import numpy as np
import pytest
import torch

from oracle import OracleMapfGym, gae_oracle, sample_actions_oracle
from primal_ppo_b200 import random_scenario

pytestmark = pytest.mark.gpu


def test_gpu_sample_actions_bit_exact_vs_oracle_and_distribution():
    from primal_ppo_b200 import sample_actions
    rng = np.random.default_rng(3)
    logits = rng.standard_normal((70001, 5)).astype(np.float32) * 2
    ps = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    ps[::7, 2] = 0.0                                   # zero-probability actions must never be drawn
    ps[5::11] = np.array([0, 0, 0, 0, 1], dtype=np.float32)
    ps = ps.astype(np.float32)
    d = torch.from_numpy(ps).cuda()
    for draw in (0, 1, 77):
        cp = torch.empty(ps.shape[0], device="cuda")
        a = sample_actions(d, seed=99, draw=draw, chosen_p=cp)
        ra, rcp = sample_actions_oracle(ps, seed=99, draw=draw)
        np.testing.assert_array_equal(a.cpu().numpy(), ra)
        np.testing.assert_array_equal(cp.cpu().numpy().view(np.uint32), rcp.view(np.uint32))
        assert (ps[np.arange(len(ra)), ra] > 0).all()
    # distribution: one fixed row sampled 400 000 times
    p = np.array([0.05, 0.25, 0.1, 0.4, 0.2], dtype=np.float32)
    a = sample_actions(torch.from_numpy(np.tile(p, (400000, 1))).cuda(), seed=5, draw=3).cpu().numpy()
    freq = np.bincount(a, minlength=5) / a.size
    assert np.abs(freq - p).max() < 4e-3, freq
    # different draws / seeds decorrelate
    b = sample_actions(torch.from_numpy(np.tile(p, (400000, 1))).cuda(), seed=5, draw=4).cpu().numpy()
    assert 0.2 < (a == b).mean() < 0.35          # sum p^2 = 0.275


def test_gpu_rollout_collect_matches_oracle_replay_and_update_runs():
    """collect(): the buffer the kernels filled in place equals an oracle replay of the sampled actions (bit-exact env
    outputs, GAE returns), then a few PPO minibatches run and change the parameters."""
    from primal_ppo_b200 import BatchedMapfGym
    from primal_ppo_b200.ppo import PPOConfig, ScrimpPolicy, VecPPOTrainer
    W, N, T = 24, 4, 6
    sc = random_scenario(W, 12, 12, N, density=(0.05, 0.2), queue_len=4, seed=8)
    env = BatchedMapfGym(sc, use_tape=False, seed=1)
    torch.manual_seed(0)
    pol = ScrimpPolicy().cuda().eval().use_channels_last()
    cfg = PPOConfig(n_steps=T, n_epochs=2, minibatch_size=16)
    tr = VecPPOTrainer(env, pol, cfg, rows_per_minibatch=16, seed=11)
    perf = tr.collect()
    b = tr.buf
    orc = OracleMapfGym(sc, seed=1, threads=2, use_tape=False)
    o_obs, o_vec = orc.getAllObservations()
    assert np.array_equal(b.obs[0].cpu().numpy(), o_obs) and np.array_equal(b.vec[0].cpu().numpy(), o_vec)
    for t in range(T):
        a = b.actions[t].cpu().numpy()
        ra, _ = sample_actions_oracle(b.ps[t].cpu().numpy(), seed=11, draw=t)
        np.testing.assert_array_equal(a, ra)
        ref = orc.step(a)
        for key, mine in (("status", b.status), ("reward", b.rewards), ("cost", b.cost_rewards),
                          ("train_valid", b.train_valid), ("goals_reached", b.goals_reached), ("violated", b.violated)):
            assert mine[t].cpu().numpy().tobytes() == ref[key].tobytes(), (t, key)
        o_obs, o_vec = orc.getAllObservations()
        assert np.array_equal(b.obs[t + 1].cpu().numpy(), o_obs), t
        assert np.array_equal(b.vec[t + 1].cpu().numpy(), o_vec), t
    # returns: the GAE kernel on the buffer == runner.py:120-149 restated, bootstrapped with the policy's value of the
    # last observation (the policy is in eval mode here, so it is deterministic)
    with torch.no_grad():
        last = pol(b.obs[T], b.vec[T])
        v0 = pol(b.obs[0], b.vec[0])
    np.testing.assert_allclose(b.values[0].cpu().numpy(), v0.value.squeeze(-1).cpu().numpy(), rtol=1e-4, atol=1e-5)
    r, v = b.rewards.cpu().numpy(), b.values.cpu().numpy()
    ref_ret, _ = gae_oracle(r, v, last.value.squeeze(-1).cpu().numpy())
    np.testing.assert_allclose(b.returns.cpu().numpy(), ref_ret, rtol=1e-4, atol=1e-4)
    cr, cv = b.cost_rewards.cpu().numpy(), b.cost_values.cpu().numpy()
    ref_cret, _ = gae_oracle(cr, cv, last.cost_value.squeeze(-1).cpu().numpy())
    np.testing.assert_allclose(b.cost_returns.cpu().numpy(), ref_cret, rtol=1e-4, atol=1e-4)
    assert np.isfinite(perf["episodeReward"]) and perf["episodeReward"] < 0
    before = torch.cat([p.detach().flatten() for p in pol.parameters()]).clone()
    stats = tr.update(perf, max_minibatches=3)
    after = torch.cat([p.detach().flatten() for p in pol.parameters()])
    assert len(stats) == 3 and all(np.isfinite(s["all_loss"]) for s in stats)
    assert float((after - before).abs().max()) > 0


def test_gpu_policy_bf16_autocast_close_to_fp32():
    from primal_ppo_b200.ppo import ScrimpPolicy
    torch.manual_seed(0)
    pol = ScrimpPolicy().cuda().eval()
    obs = (torch.rand(64, 8, 6, 9, 9, device="cuda") < 0.15).float()
    vec = torch.randn(64, 8, 4, device="cuda")
    with torch.no_grad():
        ref = pol(obs, vec)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lo = pol(obs, vec)
    assert float((ref.policy - lo.policy.float()).abs().max()) < 3e-2
    assert float((ref.value - lo.value.float()).abs().max()) < 0.15


def test_gpu_evaluate_fixed_episodes_from_reference_fixture():
    """evaluate.py's loop, batched: the reference's fixture folder -> Scenario -> all episodes at once."""
    import os
    from primal_ppo_b200.episode_io import load_fixed_episode_infos, scenario_from_fixed_episode_infos
    from primal_ppo_b200.ppo import ScrimpPolicy, evaluate_fixed_episodes
    fix = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixed_episode_infos")
    sc = scenario_from_fixed_episode_infos(load_fixed_episode_infos(fix), max_steps=48, use_da=True, use_hp=True)
    torch.manual_seed(0)
    pol = ScrimpPolicy().cuda().eval()
    res = evaluate_fixed_episodes(pol, sc, max_steps=48, greedy=False, model_name="PPOL-HP+DA", frames_for_world=1)
    per, m = res["per_episode"], res["metrics"]
    assert per["episodeReward"].shape == (4,) and (per["episodeReward"] < 0).all() and not per["err"].any()
    assert set(k.split("/")[1] for k in m) == {f"{a}_per_agent{b}" for a in ("hc", "cv", "ecr", "goals") for b in ("", "_per_timestep")}
    assert abs(m["PPOL-HP+DA/goals_per_agent/mean"] - per["totalGoals"].mean() / 3) < 1e-12
    assert len(res["frames"]) == 49 and res["frames"][0].dtype == np.uint8
    res2 = evaluate_fixed_episodes(pol, sc, max_steps=48, greedy=True)
    res3 = evaluate_fixed_episodes(pol, sc, max_steps=48, greedy=True)
    assert np.array_equal(res2["per_episode"]["episodeReward"], res3["per_episode"]["episodeReward"])   # greedy is deterministic


def test_gpu_fresh_worlds_every_rollout():
    """reset(scenario): worlds replaced in place by a device-generated batch; the trainer does it before every rollout
    (runner.py:30).  The env after reset(new) behaves like an env created on the new scenario."""
    from primal_ppo_b200 import BatchedMapfGym, generate_scenario_device
    from primal_ppo_b200.ppo import PPOConfig, ScrimpPolicy, VecPPOTrainer
    gen = lambda k: generate_scenario_device(32, 40, 60, 4, kind="warehouse", size_range=(10, 40), queue_len=4, seed=100 + k)
    env = BatchedMapfGym(gen(0), use_tape=False, seed=3)
    a = torch.randint(0, 5, (5, 32, 4), dtype=torch.int8, device="cuda")
    for t in range(5):
        env.step_observe(a[t])
    env.reset(gen(1))
    ref = BatchedMapfGym(gen(1), use_tape=False, seed=3)
    assert torch.equal(env.getAllObservations()[0], ref.getAllObservations()[0])
    for t in range(5):
        o1, ob1, _ = env.step_observe(a[t]); o2, ob2, _ = ref.step_observe(a[t])
        assert torch.equal(ob1, ob2) and torch.equal(o1.reward, o2.reward) and torch.equal(o1.status, o2.status)
    assert torch.equal(env.counters(), ref.counters())
    with pytest.raises(ValueError):
        env.reset(generate_scenario_device(16, 40, 60, 4, kind="warehouse", queue_len=4, seed=1))
    torch.manual_seed(0)
    tr = VecPPOTrainer(BatchedMapfGym(gen(0), use_tape=False), ScrimpPolicy().cuda().eval(), PPOConfig(n_steps=4, n_epochs=1),
                       rows_per_minibatch=16, fresh_worlds=gen)
    tr.collect(); first = tr.buf.obs[0].clone()
    tr.collect()
    assert not torch.equal(first, tr.buf.obs[0])
    assert torch.equal(tr.buf.obs[0], BatchedMapfGym(gen(1), use_tape=False).getAllObservations()[0])


def test_gpu_policy_fused_conv_relu_path_equals_plain_path():
    """no_grad + channels_last takes cuDNN's fused conv+bias+ReLU; it must give the values of the autograd path."""
    from primal_ppo_b200.ppo import ScrimpPolicy
    torch.manual_seed(0)
    pol = ScrimpPolicy().cuda().eval().use_channels_last()
    obs = (torch.rand(96, 4, 6, 9, 9, device="cuda") < 0.15).float()
    vec = torch.randn(96, 4, 4, device="cuda")
    for dt in (None, torch.bfloat16):
        with torch.autocast("cuda", dtype=dt, enabled=dt is not None):
            with torch.no_grad():
                a = pol(obs, vec)
            b = pol(obs, vec)                       # grad enabled -> plain path
        tol = 1e-5 if dt is None else 2e-2
        assert float((a.policy.float() - b.policy.detach().float()).abs().max()) < tol
        assert float((a.value.float() - b.value.detach().float()).abs().max()) < tol * 10
