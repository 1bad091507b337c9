"""Eval-fixture path: the reference's `fixed_episode_infos` folder -> batched Scenario.  Fixtures were
written by the reference itself (tests/golden/make_eval_golden.py).  CPU part: A* tie-breaking, JSON round trip, and the
converted scenario replayed through the oracle == the reference's FixedMapfGym driven from the same fixture."""
import filecmp
import os

import numpy as np
import pytest

from oracle import OracleMapfGym
from primal_ppo_b200.episode_io import (astar_path, load_fixed_episode_infos, render_world, save_fixed_episode_infos,
                                        scenario_from_fixed_episode_infos)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIX = os.path.join(GOLDEN, "fixed_episode_infos")


def test_astar_restatement_reproduces_reference_paths():
    d = np.load(os.path.join(GOLDEN, "astar_paths.npz"))
    n = int(d["count"])
    assert n >= 300
    for k in range(n):
        world, (s, g), ref = d[f"world{k}"].astype(int), d[f"sg{k}"], d[f"path{k}"]
        p = astar_path(world, tuple(s), tuple(g))
        if ref.shape[0] == 0:
            assert p is None, k
        else:
            assert p is not None and np.array_equal(np.asarray(p, dtype=np.int16), ref), k


def test_fixture_round_trip_is_byte_identical(tmp_path):
    infos = load_fixed_episode_infos(FIX)
    assert infos["numEpisodes"] == 4 and len(infos["agentsSequence"][0]) == 3
    save_fixed_episode_infos(infos, str(tmp_path))
    assert filecmp.cmp(os.path.join(FIX, "infos.json"), tmp_path / "infos.json", shallow=False)
    for i in range(4):
        assert np.array_equal(np.load(os.path.join(FIX, f"obstacleMap{i}.npy")), np.load(tmp_path / f"obstacleMap{i}.npy"))


def _replay(env_factory, mtype):
    g = np.load(os.path.join(GOLDEN, "eval_episodes.npz"))
    infos = load_fixed_episode_infos(FIX)
    sc = scenario_from_fixed_episode_infos(infos, human_movement_type=mtype, max_steps=48, use_da=True, use_hp=True)
    env = env_factory(sc)
    T = int(g[f"m{mtype}_T"])
    W, N = sc.num_worlds, sc.num_agents

    def check_obs(t):
        obs, vec = env.getAllObservations()
        obs, vec = np.asarray(obs.cpu() if hasattr(obs, "cpu") else obs), np.asarray(vec.cpu() if hasattr(vec, "cpu") else vec)
        for e in range(W):
            ref = np.unpackbits(g[f"m{mtype}_obs_e{e}"][t])[:N * 6 * 81].reshape(N, 6, 9, 9)
            assert np.array_equal(obs[e].astype(np.uint8), ref), (mtype, t, e)
            assert np.array_equal(vec[e].view(np.uint32), g[f"m{mtype}_vec_e{e}"][t].view(np.uint32)), (mtype, t, e)
    check_obs(0)
    for t in range(T):
        out = env.step(g["actions"][t])
        get = (lambda k: np.asarray(getattr(out, k).cpu())) if not isinstance(out, dict) else (lambda k: out[k])
        assert np.array_equal(get("status"), g[f"m{mtype}_status"][t]), (mtype, t)
        assert np.array_equal(get("reward").view(np.uint32), g[f"m{mtype}_reward"][t].view(np.uint32)), (mtype, t)
        st = env.state()
        pos = np.asarray(st["pos"].cpu() if hasattr(st["pos"], "cpu") else st["pos"])
        assert np.array_equal(pos, g[f"m{mtype}_pos"][t + 1]), (mtype, t)
        # the human the scenario carries is the human the reference walked
        tick = (t + 1) % sc.hlen
        assert np.array_equal(sc.htrace[np.arange(W), tick], g[f"m{mtype}_human"][t + 1]), (mtype, t)
        check_obs(t + 1)


@pytest.mark.parametrize("mtype", [0, 1])
def test_fixture_scenario_replays_reference_episodes_on_oracle(mtype):
    _replay(lambda sc: OracleMapfGym(sc, threads=2), mtype)


@pytest.mark.gpu
@pytest.mark.parametrize("mtype", [0, 1])
def test_gpu_fixture_scenario_replays_reference_episodes(mtype):
    import torch
    from primal_ppo_b200 import BatchedMapfGym

    class Env(BatchedMapfGym):
        def step(self, a):
            return super().step(torch.from_numpy(np.ascontiguousarray(a)))
    _replay(lambda sc: Env(sc), mtype)


def test_render_world_frame():
    infos = load_fixed_episode_infos(FIX)
    m = infos["obstacleMap"][0]
    seqs = infos["agentsSequence"][0]
    img = render_world(m != 0, np.array([s[0] for s in seqs]), np.array([s[1] for s in seqs]), infos["humanStart"][0], scale=10)
    assert img.shape == (m.shape[0] * 10, m.shape[1] * 10, 3) and img.dtype == np.uint8
    assert len(np.unique(img.reshape(-1, 3), axis=0)) >= 5
