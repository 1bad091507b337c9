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


def test_gpu_gae2_both_streams_one_launch_bit_exact():
    """mapf_gae2 (both rollout streams in one launch) = two mapf_gae calls = the oracle, bit for bit."""
    import numpy as np
    from oracle import gae_oracle
    from primal_ppo_b200 import gae, gae2
    rng = np.random.default_rng(3)
    for T, cols in ((64, 4096), (7, 33), (256, 8 * 32)):
        a = [rng.normal(size=(T, cols)).astype(np.float32) for _ in range(4)]
        lv, lcv = (rng.normal(size=(cols,)).astype(np.float32) for _ in range(2))
        t = lambda x: torch.from_numpy(x).cuda()
        ret, cret = gae2(t(a[0]), t(a[1]), t(lv), t(a[2]), t(a[3]), t(lcv))
        assert torch.equal(ret, gae(t(a[0]), t(a[1]), t(lv))) and torch.equal(cret, gae(t(a[2]), t(a[3]), t(lcv)))
        assert ret.cpu().numpy().tobytes() == gae_oracle(a[0], a[1], lv)[0].tobytes()
        assert cret.cpu().numpy().tobytes() == gae_oracle(a[2], a[3], lcv)[0].tobytes()


@pytest.mark.parametrize("variant", ["default", "cost_terms", "no_mix"])
def test_gpu_fused_ppo_loss_matches_pytorch_loss_values_and_gradients(variant):
    """The fused elementwise loss kernel (csrc/ppo_loss.cu) against ppo/loss.py (the PyTorch restatement of model.py:104-164
    that tests/test_ppo_parity.py pins to the reference): loss, every statistic, and the gradients with respect to all four
    network outputs.  fp32; tolerance 2e-5 relative to the largest gradient element, 1e-5 + 1e-5*|x| on the statistics."""
    from primal_ppo_b200.ppo.fused_loss import fused_ppo_lagrange_loss
    from primal_ppo_b200.ppo.loss import PPOConfig, ppo_lagrange_loss
    from primal_ppo_b200.ppo.policy import PolicyOutput
    g = torch.Generator(device="cuda").manual_seed(11)
    B, N = 96, 32
    cfg = {"default": PPOConfig(), "cost_terms": PPOConfig(cost_value_coef=0.05, cost_coef=0.2),
           "no_mix": PPOConfig(minus_adv_with_cadv=False, cost_coef=0.3, cost_value_coef=0.1)}[variant]
    lag = 0.7

    def inputs():
        logits = torch.randn((B, N, 5), generator=g, device="cuda")
        policy = torch.softmax(logits, -1)
        policy[0, 0] = torch.tensor([1.0, 0, 0, 0, 0], device="cuda")          # exercises the clamps
        value = torch.randn((B, N, 1), generator=g, device="cuda")
        cost_value = torch.randn((B, N, 1), generator=g, device="cuda")
        sig = torch.sigmoid(3 * torch.randn((B, N, 5), generator=g, device="cuda"))
        sig[0, 1, 0] = 1.0; sig[0, 1, 1] = 0.0
        return [x.requires_grad_(True) for x in (policy, value, cost_value, sig)]
    data = dict(returns=torch.randn((B, N), generator=g, device="cuda"), cost_returns=torch.randn((B, N), generator=g, device="cuda"),
                old_v=torch.randn((B, N), generator=g, device="cuda"), old_cv=torch.randn((B, N), generator=g, device="cuda"),
                actions=torch.randint(0, 5, (B, N), generator=g, device="cuda", dtype=torch.int8),
                old_ps=torch.softmax(torch.randn((B, N, 5), generator=g, device="cuda"), -1),
                train_valid=(torch.rand((B, N, 5), generator=g, device="cuda") < 0.6).float())
    data["old_v"][1] = data["returns"][1] + 0.01                              # values inside and outside the clip range
    leaves = inputs()
    clones = [x.detach().clone().requires_grad_(True) for x in leaves]

    def run(fn, xs):
        kw = {f: None for f in PolicyOutput._fields}
        kw.update(policy=xs[0], value=xs[1], cost_value=xs[2], policy_sig=xs[3])
        out = PolicyOutput(**kw)
        loss, stats = fn(out, lagrangian=lag, cfg=cfg, group=None, **data)
        loss.backward()
        return loss, stats
    l1, s1 = run(ppo_lagrange_loss, leaves)
    l2, s2 = run(fused_ppo_lagrange_loss, clones)
    assert abs(float(l1) - float(l2)) <= 1e-5 + 1e-5 * abs(float(l1)), (float(l1), float(l2))
    for k in s1:
        a, b = float(s1[k]), float(s2[k])
        assert abs(a - b) <= 1e-5 + 1e-5 * abs(a), (k, a, b)
    for x, y, name in zip(leaves, clones, ("policy", "value", "cost_value", "policy_sig")):
        scale = float(x.grad.abs().max())
        if scale == 0.0:
            assert float(y.grad.abs().max()) == 0.0, name
            continue
        err = float((x.grad - y.grad).abs().max())
        assert err <= 2e-5 * scale, (name, err, scale)


def test_gpu_learner_with_fused_loss_equals_learner_with_pytorch_loss():
    """PPOLearner end to end (forward, loss, backward, flat gradient): fused kernel vs eager loss, same weights, fp32."""
    from primal_ppo_b200.ppo import PPOConfig, ScrimpPolicy
    from primal_ppo_b200.ppo.trainer import PPOLearner
    g = torch.Generator(device="cuda").manual_seed(5)
    B, N = 8, 16
    batch = dict(obs=(torch.rand((B, N, 6, 9, 9), generator=g, device="cuda") < 0.2).float(),
                 vec=torch.randn((B, N, 4), generator=g, device="cuda"), returns=torch.randn((B, N), generator=g, device="cuda"),
                 cost_returns=torch.randn((B, N), generator=g, device="cuda"), values=torch.randn((B, N), generator=g, device="cuda"),
                 cost_values=torch.randn((B, N), generator=g, device="cuda"),
                 actions=torch.randint(0, 5, (B, N), generator=g, device="cuda", dtype=torch.int8),
                 ps=torch.softmax(torch.randn((B, N, 5), generator=g, device="cuda"), -1),
                 train_valid=(torch.rand((B, N, 5), generator=g, device="cuda") < 0.6).float())
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    try:
        grads, stats = [], []
        for fused in (False, True):
            torch.manual_seed(3)
            pol = ScrimpPolicy().cuda().eval()
            lr = PPOLearner(pol, PPOConfig(cost_value_coef=0.05, cost_coef=0.2), fused_loss=fused)
            stats.append(lr.compute_gradients(batch))
            grads.append(lr.flat_grad.clone())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    scale = float(grads[0].abs().max())
    assert float((grads[0] - grads[1]).abs().max()) <= 5e-5 * scale
    for k in stats[0]:
        assert abs(stats[0][k] - stats[1][k]) <= 1e-5 + 1e-4 * abs(stats[0][k]), k
