"""The reference's actual TRAINING environment (`MapfGym()`: random warehouse, goal-re-drawing `Human`, `getFreeCell`
starts and goals) recorded by tests/golden/make_training_golden.py, replayed from the recorded exogenous draws.
CPU: through the oracle; GPU: through the C ABI (fused step+observe).  Everything must match bit for bit."""
import os

import numpy as np
import pytest

from oracle import OracleMapfGym
from primal_ppo_b200.scenario import Scenario

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "training_env.npz"))


def _scenario(ci):
    p = f"c{ci}_"
    ht = np.ascontiguousarray(G[p + "htrace"])
    W, L = ht.shape[:2]
    sc = Scenario(obst=G[p + "obst"], starts=G[p + "starts"], goal_queue=G[p + "goal_queue"], htrace=ht,
                  hlen=np.full((W,), L, dtype=np.int32), dims=G[p + "dims"])
    sc.validate()
    return sc


def _check(ci, make_env, to_np):
    p = f"c{ci}_"
    sc = _scenario(ci)
    env = make_env(sc)
    T, W, N = int(G["T"]), sc.num_worlds, sc.num_agents

    def check_obs(t, obs, vec):
        obs, vec = to_np(obs), to_np(vec)
        for w in range(W):
            ref = np.unpackbits(G[p + f"obs_w{w}"][t])[:N * 6 * 81].reshape(N, 6, 9, 9)
            assert np.array_equal(obs[w].astype(np.uint8), ref), (ci, t, w)
        assert np.array_equal(vec.view(np.uint32), G[p + "vec"][t].view(np.uint32)), (ci, t)
    check_obs(0, *env.getAllObservations())
    for t in range(T):
        out, obs, vec = env.step_observe(G[p + "actions"][t])
        for key in ("status", "reward", "cost", "train_valid", "goals_reached", "violated"):
            a, b = to_np(out[key] if isinstance(out, dict) else getattr(out, key)), G[p + key][t]
            assert a.tobytes() == b.astype(a.dtype).tobytes(), (ci, t, key)
        sh = to_np(out["shadow"] if isinstance(out, dict) else out.shadow_goals)
        assert np.array_equal(sh, G[p + "shadow"][t]), (ci, t)
        st = env.state()
        assert np.array_equal(to_np(st["pos"]), G[p + "pos"][t + 1]) and np.array_equal(to_np(st["goal"]), G[p + "goal"][t + 1]), (ci, t)
        assert not to_np(st["err"]).any()
        check_obs(t + 1, obs, vec)


class _OracleEnv(OracleMapfGym):
    def step_observe(self, a):
        out = self.step(np.ascontiguousarray(a))
        obs, vec = self.getAllObservations()
        return out, obs, vec


@pytest.mark.parametrize("ci", range(int(G["n_cases"])))
def test_training_env_replay_on_oracle(ci):
    _check(ci, lambda sc: _OracleEnv(sc, threads=2), np.asarray)


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(int(G["n_cases"])))
def test_gpu_training_env_replay(ci):
    import torch
    from primal_ppo_b200 import BatchedMapfGym

    class Env(BatchedMapfGym):
        def step_observe(self, a):
            return super().step_observe(torch.from_numpy(np.ascontiguousarray(a)))
    _check(ci, lambda sc: Env(sc), lambda t: t.cpu().numpy())
