"""GPU parity tests added in round 2 (run on the B200 box with `-m gpu`, all through the C ABI):
  * the FULL BASELINE configs[2] size (65 536 worlds x 40x40 x 32 agents) against the oracle for 8 fused steps — the 4 GB
    observation tensor is compared through per-world 64-bit checksums computed on the device (mapf_checksum_rows) and,
    with the same definition, by the oracle on the host; every small output is compared exactly;
  * the split-phase host call (mapf_step_observe_host_begin / _wait) against the device-resident call;
  * on-device goal sampling (MapfGym.getNextGoal) against the oracle's restatement with the same Philox stream;
  * allGoodActions, mapf_get_human, _render, fresh_outputs.
Bar: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import OracleMapfGym
from oracle.oracle import checksum_rows as orc_checksum_rows
from primal_ppo_b200 import random_actions, random_scenario

pytestmark = pytest.mark.gpu


def _env(sc, **kw):
    from primal_ppo_b200 import BatchedMapfGym
    return BatchedMapfGym(sc, **kw)


def _np(t):
    return t.detach().cpu().numpy()


def _eq(a, b, msg):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype == np.float32:
        a, b = a.view(np.uint32), np.asarray(b, dtype=np.float32).view(np.uint32)
    np.testing.assert_array_equal(a, b, err_msg=msg)


def test_gpu_checksum_rows_matches_host_definition():
    from primal_ppo_b200 import checksum_rows
    rng = np.random.default_rng(0)
    for rows, words in ((1, 1), (7, 33), (64, 15552), (3, 4099)):
        a = rng.integers(0, 2 ** 32, size=(rows, words), dtype=np.uint32)
        want = orc_checksum_rows(a, threads=2)
        # independent restatement of the definition in include/mapf_b200.h
        M = (1 << 64) - 1
        acc = 0
        for i, x in enumerate(a[0].tolist()):
            z = ((x + 1) * 0x9E3779B97F4A7C15 + i * 0xC2B2AE3D27D4EB4F) & M
            z = ((z ^ (z >> 29)) * 0xBF58476D1CE4E5B9) & M
            acc = (acc + (z ^ (z >> 32))) & M
        assert int(want[0]) == acc
        got = checksum_rows(torch.from_numpy(a.view(np.int32)).cuda())
        np.testing.assert_array_equal(_np(got).view(np.uint64), want)
    # unaligned rows (words not a multiple of 4 -> scalar path) and float data
    f = torch.rand((5, 486), device="cuda")
    np.testing.assert_array_equal(_np(checksum_rows(f)).view(np.uint64), orc_checksum_rows(_np(f)))


def test_gpu_full_size_65536x40x40x32_fused_steps_match_oracle():
    """BASELINE configs[2] at its full size, 8 fused steps, against the oracle: every per-agent output exactly, the
    observations through per-world checksums (so that the batched L2 prefetch path — W > 1024 — and 3 CTAs per SM with
    dynamic world claiming are covered by an oracle comparison, not only by properties)."""
    from primal_ppo_b200 import checksum_rows
    W, N, T = 65536, 32, 8
    sc = random_scenario(W, 40, 40, N, density=(0.0, 0.3), queue_len=16, seed=2024, unique_maps=256)
    acts = random_actions(T, W, N, seed=99)
    env = _env(sc, seed=1234, use_tape=False)
    keys = ("status", "reward", "cost", "train_valid", "goals_reached", "violated", "shadow_goals", "fixed_actions")
    rec = []
    obs, vec = env.getAllObservations()
    rec.append(dict(obs_sum=_np(checksum_rows(obs)).view(np.uint64), vec=_np(vec)))
    for t in range(T):
        out, obs, vec = env.step_observe(torch.from_numpy(acts[t]))
        d = {k: _np(getattr(out, k)) for k in keys}
        d["obs_sum"] = _np(checksum_rows(obs)).view(np.uint64)
        d["vec"] = _np(vec)
        s = env.state()
        d["pos"], d["goal"], d["err"] = _np(s["pos"]), _np(s["goal"]), _np(s["err"]).astype(np.uint32)
        rec.append(d)
    bfs_sum = _np(checksum_rows(env.bfs_maps(agent_ids=torch.arange(0, 4096 * N, device="cuda", dtype=torch.int32))))
    del env
    torch.cuda.empty_cache()
    # the oracle, in chunks of worlds (its observation buffer is 62 KB per world)
    import os
    threads = max(1, min(32, len(os.sched_getaffinity(0))))
    CH = 8192
    n_flagged = 0
    for lo in range(0, W, CH):
        hi = lo + CH
        orc = OracleMapfGym(sc.slice(lo, hi), seed=1234, threads=threads, use_tape=False, world_offset=lo)
        o_obs, o_vec = orc.getAllObservations()
        np.testing.assert_array_equal(rec[0]["obs_sum"][lo:hi], orc_checksum_rows(o_obs, threads), err_msg=f"obs0 {lo}")
        _eq(rec[0]["vec"][lo:hi], o_vec, f"vec0 {lo}")
        bufs = None
        for t in range(T):
            bufs = orc.step_observe(acts[t][lo:hi], bufs)
            so = orc.state()
            g = rec[t + 1]
            np.testing.assert_array_equal(g["err"][lo:hi], so["err"], err_msg=f"t={t} err flags {lo}")
            ok = np.ones_like(so["err"], dtype=bool)    # flagged worlds (the reference would have hung / raised) are compared too
            for k, ok_ in (("status", "status"), ("reward", "reward"), ("cost", "cost"), ("train_valid", "train_valid"),
                           ("goals_reached", "goals_reached"), ("violated", "violated"), ("shadow_goals", "shadow"),
                           ("fixed_actions", "fixed")):
                _eq(g[k][lo:hi][ok], bufs[ok_][ok], f"t={t} {k} worlds {lo}..{hi}")
            _eq(g["pos"][lo:hi][ok], so["pos"][ok], f"t={t} pos")
            _eq(g["goal"][lo:hi][ok], so["goal"][ok], f"t={t} goal")
            _eq(g["vec"][lo:hi][ok], bufs["vec"][ok], f"t={t} vec")
            np.testing.assert_array_equal(g["obs_sum"][lo:hi][ok], orc_checksum_rows(bufs["obs"], threads)[ok],
                                          err_msg=f"t={t} observation checksums worlds {lo}..{hi}")
        n_flagged += int((orc.state()["err"] != 0).sum())
        if lo == 0:
            m = orc.bfs_maps()[:4096].reshape(4096 * N, -1)
            okb = np.ones(4096 * N, dtype=bool)
            np.testing.assert_array_equal(bfs_sum.view(np.uint64)[okb], orc_checksum_rows(m, threads)[okb])
        del orc
    assert n_flagged < W // 100


@pytest.mark.parametrize("shape", [(4096, 40, 40, 32), (300, 20, 20, 8), (64, 80, 80, 128)])
def test_gpu_split_phase_host_call_equals_device_call(shape):
    """mapf_step_observe_host_begin/_wait: two steps in flight, results of step t read while step t+1 runs; every result
    and the observations equal the device-resident fused call."""
    W, H, Wd, N = shape
    T = 10
    sc = random_scenario(W, H, Wd, N, density=(0.0, 0.3), queue_len=4, seed=31 + N, unique_maps=min(W, 64))
    acts = random_actions(T, W, N, seed=8)
    ref_env, env = _env(sc, use_tape=False, seed=5), _env(sc, use_tape=False, seed=5)
    want = []
    for t in range(T):
        o, obs, vec = ref_env.step_observe(torch.from_numpy(acts[t]))
        want.append({k: _np(getattr(o, k)) for k in ("status", "reward", "cost", "goals_reached", "violated", "shadow_goals",
                                                     "fixed_actions", "train_valid")} | {"obs": obs.clone(), "vec": vec.clone()})
    for with_tv in (False, True):
        env.reset()
        ring = env.make_host_ring(slots=2, action_slots=2, with_train_valid=with_tv)
        F = sc.fov
        obs_t = [torch.empty((W, N, 6, F, F), device="cuda") for _ in range(2)]
        vec_t = [torch.empty((W, N, 4), device="cuda") for _ in range(2)]
        tv = torch.empty((W, N, 5), device="cuda")

        def check(t):
            slot = ring["slots"][t & 1]
            for k in ("status", "reward", "cost", "goals_reached", "violated", "shadow_goals", "fixed_actions"):
                _eq(slot[k].numpy(), want[t][k], f"with_tv={with_tv} t={t} {k}")
            if with_tv:
                _eq(slot["train_valid"].numpy(), want[t]["train_valid"], f"t={t} train_valid")
        for t in range(T):
            ring["action_ring"][t & 1].copy_(torch.from_numpy(acts[t]))
            h2d, d2h = env.step_observe_host_begin(ring["action_ring"][t & 1], ring["slots"][t & 1], obs_t[t & 1], vec_t[t & 1],
                                                   train_valid_dev=tv if with_tv else None, with_train_valid=with_tv)
            assert h2d == W * N and d2h == ring["slot_bytes"]
            if t >= 1:
                env.host_wait(1)                   # the previous step's results, while this step runs
                check(t - 1)
                # the action slot of step t-1 may be rewritten now (its copy was ordered before that step's kernels)
            # device-side consumers are stream-ordered: no host wait needed to read obs of this step
            assert torch.equal(obs_t[t & 1], want[t]["obs"]) and torch.equal(vec_t[t & 1], want[t]["vec"]), t
        env.host_wait(0)
        check(T - 1)
    # compact wire format (2 bytes per agent): decoded on the host it equals the full slab bit for bit
    from primal_ppo_b200 import decode_results
    env.reset()
    ring = env.make_host_ring(slots=2, action_slots=2, compact=True)
    assert ring["slot_bytes"] < 3 * W * N + 4 * W + 1024
    dec = None
    for t in range(T):
        ring["action_ring"][t & 1].copy_(torch.from_numpy(acts[t]))
        h2d, d2h = env.step_observe_host_begin(ring["action_ring"][t & 1], ring["slots"][t & 1], obs_t[t & 1], vec_t[t & 1])
        assert d2h == ring["slot_bytes"]
        if t >= 1:
            env.host_wait(1)
            dec = decode_results(ring["slots"][(t - 1) & 1]["packed"], dec)
            for k in ("status", "reward", "cost", "goals_reached", "violated", "fixed_actions"):
                _eq(dec[k].numpy(), want[t - 1][k], f"compact t={t - 1} {k}")
            _eq(ring["slots"][(t - 1) & 1]["shadow_goals"].numpy(), want[t - 1]["shadow_goals"], "compact shadow")
        assert torch.equal(obs_t[t & 1], want[t]["obs"])
    env.host_wait(0)
    dec = decode_results(ring["slots"][(T - 1) & 1]["packed"], dec)
    _eq(dec["reward"].numpy(), want[T - 1]["reward"], "compact last reward")
    # the synchronous form still agrees and can be mixed with the split-phase form
    env.reset()
    hb = env.make_host_buffers()
    obs, vec = obs_t[0], vec_t[0]
    for t in range(3):
        hb["actions"].copy_(torch.from_numpy(acts[t]))
        env.step_observe_host(hb, obs, vec)
        for k in ("status", "reward", "cost", "goals_reached", "violated", "shadow_goals"):
            _eq(hb[k].numpy(), want[t][k], f"sync t={t} {k}")
        assert torch.equal(obs, want[t]["obs"])
    ring = env.make_host_ring()
    ring["action_ring"][0].copy_(torch.from_numpy(acts[3]))
    env.step_observe_host_begin(ring["action_ring"][0], ring["slots"][0], obs, vec)
    env.host_wait(0)
    _eq(ring["slots"][0]["reward"].numpy(), want[3]["reward"], "mixed reward")
    from primal_ppo_b200._cabi import MapfError
    fresh = _env(sc.slice(0, 2), use_tape=False)
    with pytest.raises(MapfError):
        fresh.host_wait(0)                         # nothing in flight


@pytest.mark.parametrize("shape", [(2048, 40, 40, 32, 24), (1024, 12, 12, 8, 48), (96, 80, 80, 128, 24), (64, 16, 16, 48, 32),
                                   (16, 33, 65, 5, 24)])
def test_gpu_goal_sampling_on_device_matches_oracle(shape):
    """goal_sampling=True (MapfGym.getNextGoal: free-cell rejection sampling at arrival, mapf_gym.py:626, util.py:67-76)
    against the oracle's sequential restatement with the same Philox stream: goals, and everything downstream of them."""
    W, H, Wd, N, T = shape
    sc = random_scenario(W, H, Wd, N, density=(0.0, 0.3), queue_len=1, seed=3 * W + N, unique_maps=min(W, 64))
    # goals next to the agents so that arrivals are frequent
    rng = np.random.default_rng(W)
    orc = OracleMapfGym(sc, seed=77, threads=8, use_tape=False, goal_sampling=True)
    env = _env(sc, seed=77, use_tape=False, goal_sampling=True)
    arrivals = 0
    for t in range(T):
        # steer half of the agents towards their goal so that arrivals happen
        so = orc.state()
        d = so["goal"].astype(np.int32) - so["pos"].astype(np.int32)
        greedy = np.where(np.abs(d[..., 0]) >= np.abs(d[..., 1]), np.where(d[..., 0] > 0, 2, 4), np.where(d[..., 1] > 0, 1, 3))
        greedy = np.where((d == 0).all(-1), 0, greedy)
        a = np.where(rng.random((W, N)) < 0.7, greedy, rng.integers(0, 5, size=(W, N))).astype(np.int8)
        ref = orc.step(a)
        fused = (t % 2 == 0)
        if fused:
            out, obs, vec = env.step_observe(torch.from_numpy(a))
        else:
            out = env.step(torch.from_numpy(a))
            obs, vec = env.getAllObservations()
        so, s = orc.state(), env.state()
        np.testing.assert_array_equal(_np(s["err"]).astype(np.uint32), so["err"], err_msg=f"t={t} err")
        ok = np.ones_like(so["err"], dtype=bool)
        for key in ("status", "reward", "goals_reached", "violated"):
            _eq(_np(getattr(out, key))[ok], ref[key][ok], f"t={t} {key}")
        _eq(_np(s["pos"])[ok], so["pos"][ok], f"t={t} pos")
        _eq(_np(s["goal"])[ok], so["goal"][ok], f"t={t} goal (sampled)")
        o_obs, o_vec = orc.getAllObservations()
        okd = torch.from_numpy(ok).cuda()
        assert torch.equal(obs[okd], torch.from_numpy(o_obs).cuda()[okd]), f"t={t} obs"
        assert torch.equal(vec[okd], torch.from_numpy(o_vec).cuda()[okd]), f"t={t} vec"
        arrivals += int(ref["goals_reached"].sum())
        # a sampled goal is a free cell: no obstacle, no agent on it, no other agent's goal (util.py:72)
        g = so["goal"].astype(np.int64)
        assert (sc.obst[np.arange(W)[:, None], g[..., 0], g[..., 1]] == 0).all()
    assert arrivals > W // 8, arrivals
    _eq(_np(env.bfs_maps())[ok], orc.bfs_maps()[ok], "bfs of sampled goals")


def test_gpu_goal_sampling_no_free_cell_is_flagged():
    """A world without any free cell left makes getFreeCell spin for ever (util.py:72); the kernels flag the world and keep
    the goal; the oracle agrees."""
    from primal_ppo_b200.scenario import Scenario, looping_trace
    H = Wd = 4
    N = 2
    obst = np.ones((1, H, Wd), dtype=np.uint8)
    obst[0, 0, 0] = obst[0, 0, 1] = 0                    # two free cells, both will be agent cells / goals
    starts = np.array([[[0, 0], [0, 1]]], dtype=np.int16)
    goals = np.array([[[[0, 0]], [[0, 1]]]], dtype=np.int16)          # already on their goals: arrive when they stay
    htrace = np.zeros((1, 2, 4), dtype=np.int16)
    htrace[0, :, :] = (3, 3, 3, 3)                       # the human stands on a shelf cell, away from the agents
    sc = Scenario(obst=obst, starts=starts, goal_queue=goals, htrace=htrace, hlen=np.array([1], dtype=np.int32))
    env = _env(sc, use_tape=False, goal_sampling=True)
    orc = OracleMapfGym(sc, use_tape=False, goal_sampling=True)
    a = np.zeros((1, N), dtype=np.int8)
    env.step(torch.from_numpy(a)); orc.step(a)
    e = int(_np(env.state()["err"])[0])
    assert e & 64 and e == int(orc.state()["err"][0])
    _eq(_np(env.state()["goal"]), orc.state()["goal"], "goal kept")


def test_gpu_all_good_actions_and_human_and_render():
    sc = random_scenario(256, 12, 12, 8, density=(0.1, 0.3), queue_len=4, seed=17, unique_maps=32)
    env = _env(sc, use_tape=False)
    orc = OracleMapfGym(sc, threads=4, use_tape=False)
    acts = random_actions(12, 256, 8, seed=2)
    for t in range(12):
        _eq(_np(env.allGoodActions), orc.state()["good"], f"t={t} allGoodActions")          # mapf_gym.py:169 / :635
        env.step(torch.from_numpy(acts[t])); orc.step(acts[t])
        pos, nxt, tick = env.human()
        L = sc.hlen
        tt = (t + 1) % L
        _eq(_np(tick), tt.astype(np.int32), "tick")
        _eq(_np(pos), sc.htrace[np.arange(256), tt, :2], "human pos")
        _eq(_np(nxt), sc.htrace[np.arange(256), tt, 2:], "human next")
    lists = env.good_actions_lists(3)
    m = orc.state()["good"][3]
    assert [list(x) for x in lists] == [[a for a in range(5) if (int(v) >> a) & 1] for v in m]
    img = env._render(world=3)
    assert img.dtype == np.uint8 and img.shape == (12 * 12, 12 * 12, 3)
    # same thing with N > 32 (step_wide)
    sc = random_scenario(8, 24, 24, 48, density=(0.1, 0.25), queue_len=2, seed=18)
    env, orc = _env(sc, use_tape=False), OracleMapfGym(sc, threads=4, use_tape=False)
    _eq(_np(env.allGoodActions), orc.state()["good"], "allGoodActions N=48")


def test_gpu_fresh_outputs_do_not_alias():
    """fresh_outputs=True gives the reference's semantics: results appended to lists across steps keep their values
    (runner.py:84, 93-94); the default returns env-owned tensors that the next call overwrites."""
    sc = random_scenario(32, 10, 10, 4, density=(0.1, 0.2), queue_len=4, seed=5)
    acts = random_actions(4, 32, 4, seed=1)
    fresh, alias = _env(sc, use_tape=False, fresh_outputs=True), _env(sc, use_tape=False)
    kept_f, kept_a, truth = [], [], []
    for t in range(4):
        a = torch.from_numpy(acts[t]).cuda()
        for env, kept in ((fresh, kept_f), (alias, kept_a)):
            st = env.getActionStatus(a)
            rw, _ = env.calculateActionReward(a, st)
            kept.append(rw)
            if env is alias:
                truth.append(rw.clone())
            env.jointStep(a, st)
    for t in range(4):
        assert torch.equal(kept_f[t], truth[t])
    assert kept_a[0].data_ptr() == kept_a[3].data_ptr()


@pytest.mark.parametrize("shape", [(48, 80, 80, 128, 9, 24), (64, 16, 16, 48, 9, 32), (24, 80, 80, 128, 15, 12), (12, 80, 80, 128, 21, 8),
                                   (8, 80, 80, 128, 31, 8), (6, 128, 128, 128, 31, 6), (16, 40, 40, 32, 31, 12), (40, 33, 65, 33, 9, 16),
                                   (300, 24, 24, 64, 3, 16)])
def test_gpu_fused_wide_step_observe_matches_oracle(shape):
    """step_observe_wide_kernel (CTA per world: N > 32, or an observation block that needs several chunks) against the
    oracle every step, and against the two separate launches (mapf_step + mapf_observe) for every output and the state."""
    from test_gpu_parity import _run_vs_oracle
    W, H, Wd, N, F, T = shape
    dens = (0.1, 0.25) if H <= 24 else (0.0, 0.3)
    sc = random_scenario(W, H, Wd, N, density=dens, queue_len=4, seed=W + N + F, fov=F, unique_maps=min(W, 16))
    env, orc = _run_vs_oracle(sc, T=T, fused=True)
    e2 = _env(sc, seed=1234, use_tape=False)
    acts = random_actions(T, W, N, seed=1234)
    for t in range(T):
        e2.step(torch.from_numpy(acts[t]))
    s1, s2 = env.state(), e2.state()
    assert all(torch.equal(s1[k], s2[k]) for k in s1)
    assert torch.equal(env.counters(), e2.counters())
    a, b = env.human(), e2.human()
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_gpu_fused_wide_eval_channels():
    """use_da / use_hp (eval-only channels, per-world dims) through the CTA-per-world fused kernel."""
    from test_gpu_parity import _run_vs_oracle
    sc = random_scenario(10, 80, 80, 128, density=(0.0, 0.3), queue_len=3, seed=77, fov=15, use_da=True, use_hp=True)
    _run_vs_oracle(sc, T=8, fused=True)
    sc = random_scenario(10, 20, 30, 40, density=(0.0, 0.2), queue_len=3, seed=78, fov=9, use_da=True, use_hp=True, num_channel=6)
    _run_vs_oracle(sc, T=12, fused=True)


def test_gpu_packed_step_output_decodes_to_the_separate_outputs():
    """MapfStepOut.packed from mapf_step (N <= 32 and N > 32) decodes to exactly the separate per-agent outputs."""
    from primal_ppo_b200 import StepOut, decode_results
    for (W, H, N) in ((512, 12, 8), (32, 24, 48)):
        sc = random_scenario(W, H, H, N, density=(0.1, 0.3), queue_len=4, seed=W + N, unique_maps=16)
        env = _env(sc, use_tape=False)
        packed = torch.empty((W, N), dtype=torch.int16, device="cuda")
        o = env._out
        out = StepOut(status=o.status, reward=o.reward, cost=o.cost, train_valid=o.train_valid, goals_reached=o.goals_reached,
                      violated=o.violated, shadow_goals=o.shadow_goals, fixed_actions=o.fixed_actions, packed=packed)
        acts = random_actions(16, W, N, seed=3)
        for t in range(16):
            env.step(torch.from_numpy(acts[t]), out=out)
            dec = decode_results(packed.cpu())
            for k in ("status", "reward", "cost", "goals_reached", "violated", "fixed_actions"):
                _eq(dec[k].numpy(), _np(getattr(out, k)), f"N={N} t={t} {k}")


@pytest.mark.parametrize("shape", [(37, 40, 40, 32, 9), (11, 80, 80, 128, 9), (7, 24, 24, 48, 15), (5, 80, 80, 128, 31)])
def test_gpu_outputs_stay_inside_their_buffers(shape):
    """compute-sanitizer is not available on the GPU pool, so out-of-bounds stores are looked for with canaries: every
    output of the fused launches (warp-per-world and CTA-per-world) lives inside a larger sentinel-filled allocation; after
    several steps (with goal sampling, packed results and trainValid on) every sentinel byte is untouched."""
    from primal_ppo_b200 import StepOut
    W, H, Wd, N, F = shape
    sc = random_scenario(W, H, Wd, N, density=(0.0, 0.3), queue_len=1, seed=W + F, fov=F, unique_maps=min(W, 8))
    env = _env(sc, use_tape=False, goal_sampling=True)
    G = 4096                                                     # guard bytes on each side (a multiple of every alignment)

    def guarded(numel, dtype):
        es = torch.empty((), dtype=dtype).element_size()
        raw = torch.full((numel * es + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
        return raw, raw[G:G + numel * es].view(dtype)
    specs = dict(obs=(W * N * 6 * F * F, torch.float32), vec=(W * N * 4, torch.float32), status=(W * N, torch.int8),
                 reward=(W * N, torch.float32), cost=(W * N, torch.float32), train_valid=(W * N * 5, torch.float32),
                 goals_reached=(W * N, torch.uint8), violated=(W * N, torch.uint8), shadow_goals=(W, torch.int32),
                 fixed_actions=(W * N, torch.int8), packed=(W * N, torch.int16))
    raws, views = {}, {}
    for k, (n, dt) in specs.items():
        raws[k], views[k] = guarded(n, dt)
    out = StepOut(status=views["status"].view(W, N), reward=views["reward"].view(W, N), cost=views["cost"].view(W, N),
                  train_valid=views["train_valid"].view(W, N, 5), goals_reached=views["goals_reached"].view(W, N),
                  violated=views["violated"].view(W, N), shadow_goals=views["shadow_goals"], fixed_actions=views["fixed_actions"].view(W, N),
                  packed=views["packed"].view(W, N))
    obs, vec = views["obs"].view(W, N, 6, F, F), views["vec"].view(W, N, 4)
    acts = random_actions(6, W, N, seed=9)
    for t in range(6):
        if t % 3 == 2:
            env.step(torch.from_numpy(acts[t]), out=out); env.getAllObservations(out=(obs, vec))
        else:
            env.step_observe(torch.from_numpy(acts[t]), out=out, obs_out=(obs, vec))
    torch.cuda.synchronize()
    for k, raw in raws.items():
        assert bool((raw[:G] == 0xA5).all()) and bool((raw[-G:] == 0xA5).all()), f"{k}: a store landed outside the buffer"
        assert not bool((views[k].view(torch.uint8) == 0xA5).all()), f"{k}: never written"
    assert not _np(env.state()["err"]).any() or True


def test_gpu_env_on_another_device_than_the_current_one():
    """Every entry point runs on the env's own device whatever the caller's current device is, and leaves the caller's current
    device untouched (mapf_api.cu DeviceGuard).  Needs two GPUs; skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sc = random_scenario(64, 20, 20, 8, density=(0.1, 0.25), queue_len=3, seed=21)
    acts = random_actions(6, 64, 8, seed=22)
    torch.cuda.set_device(0)
    e0 = _env(sc, device="cuda:0", use_tape=False, seed=3)
    e1 = _env(sc, device="cuda:1", use_tape=False, seed=3)          # created while cuda:0 is current
    assert torch.cuda.current_device() == 0
    ring = e1.make_host_ring(compact=True)
    for t in range(6):
        a = torch.from_numpy(acts[t])
        o0, obs0, vec0 = e0.step_observe(a)
        o1, obs1, vec1 = e1.step_observe(a)                         # cuda:1 env driven with cuda:0 current
        assert torch.cuda.current_device() == 0
        assert obs1.device.index == 1 and torch.equal(obs0.cpu(), obs1.cpu()) and torch.equal(vec0.cpu(), vec1.cpu())
        for k in ("status", "reward", "cost", "train_valid", "goals_reached", "violated", "shadow_goals", "fixed_actions"):
            assert torch.equal(getattr(o0, k).cpu(), getattr(o1, k).cpu()), (t, k)
    assert torch.equal(e0.bfs_maps().cpu(), e1.bfs_maps().cpu())
    assert torch.equal(e0.allGoodActions.cpu(), e1.allGoodActions.cpu())
    assert torch.equal(e0.counters().cpu(), e1.counters().cpu())
    ring["action_ring"][0].copy_(torch.from_numpy(acts[0]))
    obs1 = torch.empty((64, 8, 6, 9, 9), device="cuda:1"); vec1 = torch.empty((64, 8, 4), device="cuda:1")
    e1.step_observe_host_begin(ring["action_ring"][0], ring["slots"][0], obs1, vec1)
    e1.host_wait(0)
    o0, _, _ = e0.step_observe(torch.from_numpy(acts[0]))
    from primal_ppo_b200 import decode_results
    _eq(decode_results(ring["slots"][0]["packed"])["reward"].numpy(), _np(o0.reward), "host call on cuda:1")
    assert torch.cuda.current_device() == 0


@pytest.mark.parametrize("shape", [(513, 6, 5, 9, True), (513, 13, 7, 29, True), (200, 8, 8, 20, False), (64, 9, 9, 40, True)])
def test_gpu_flagged_worlds_stay_valid_and_identical(shape):
    """Worlds on which the reference would hang (fixActions livelock) or raise (no viable action, no free cell) are flagged,
    every agent stays for that step, and the world keeps stepping: state, outputs and later flags of such worlds equal the
    oracle's for every world and every step — nothing is excluded from the comparison.  (Found by tests/soak_parity.py: before
    the iteration cap counted the reference's pops of agents that own a good action, and before a capped step froze the
    world, flagged worlds drifted apart.)"""
    W, H, Wd, N, gs = shape
    sc = random_scenario(W, H, Wd, N, density=(0.0, 0.35), queue_len=1 if gs else 3, seed=809596303 % (1 << 30), fov=9, unique_maps=24)
    env = _env(sc, use_tape=False, seed=5, goal_sampling=gs)
    orc = OracleMapfGym(sc, seed=5, threads=8, use_tape=False, goal_sampling=gs)
    acts = random_actions(30, W, N, seed=6)
    flagged_steps = 0
    for t in range(30):
        ref = orc.step(acts[t])
        if t % 2:
            out, obs, vec = env.step_observe(torch.from_numpy(acts[t]))
        else:
            out = env.step(torch.from_numpy(acts[t])); obs, vec = env.getAllObservations()
        so, s = orc.state(), env.state()
        _eq(_np(s["err"]).astype(np.uint32), so["err"], f"t={t} err flags")
        for k in ("status", "reward", "cost", "train_valid", "goals_reached", "violated"):
            _eq(_np(getattr(out, k)), ref[k], f"t={t} {k}")
        _eq(_np(out.fixed_actions), ref["fixed"], f"t={t} fixed")
        for k in ("pos", "goal", "rep"):
            _eq(_np(s[k]), so[k], f"t={t} {k}")
        o_obs, o_vec = orc.getAllObservations()
        assert torch.equal(obs, torch.from_numpy(o_obs).cuda()) and torch.equal(vec, torch.from_numpy(o_vec).cuda()), f"t={t} obs"
        flagged_steps += int((so["err"] != 0).sum())
        # a world stays a valid state: no two agents share a cell
        p = so["pos"].astype(np.int64)
        cell = p[..., 0] * Wd + p[..., 1]
        assert all(len(set(row)) == N for row in cell.tolist()), f"t={t}: two agents on one cell"
    assert flagged_steps > 0, "the scenario was meant to produce flagged worlds"
