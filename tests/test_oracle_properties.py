"""Property tests (hypothesis) of the oracle on random scenarios — the invariants the reference asserts or implies
(SURVEY.md §4): fixActions post-condition (`mapf_gym.py:600-610`): executed actions never collide; agents stay on free
cells, one per cell; BFS maps are 1-Lipschitz along free edges with the goal at 0; observations are 0/1 with the agent's
own cell marked; trainValid marks every unconditionally good action.  The same properties are checked for the CUDA path
at full size in tests/test_gpu_parity.py."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import OracleMapfGym
from primal_ppo_b200 import random_actions, random_scenario

DIRS = np.array([[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0]])


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(0, 10_000), st.sampled_from([(8, 8, 6), (10, 14, 8), (16, 16, 12), (20, 20, 8)]), st.floats(0.0, 0.3))
def test_step_invariants(seed, shape, dens):
    H, Wd, N = shape
    W = 12
    sc = random_scenario(W, H, Wd, N, density=(dens, dens), queue_len=4, seed=seed)
    env = OracleMapfGym(sc, seed=seed, threads=2, use_tape=False)
    acts = random_actions(10, W, N, seed=seed + 1)
    free = sc.obst == 0
    widx = np.arange(W)[:, None]
    for t in range(10):
        before = env.state()["pos"].astype(np.int64)
        out = env.step(acts[t])
        s = env.state()
        ok = s["err"] == 0
        pos = s["pos"].astype(np.int64)
        # executed action moves the agent by exactly its direction
        np.testing.assert_array_equal((pos - before)[ok], DIRS[out["fixed"].astype(np.int64)][ok])
        assert (pos >= 0).all() and (pos[..., 0] < H).all() and (pos[..., 1] < Wd).all()
        assert free[widx, pos[..., 0], pos[..., 1]][ok].all(), "agents stay on free cells"
        cell = np.sort(pos[..., 0] * Wd + pos[..., 1], axis=1)
        assert (cell[:, 1:] != cell[:, :-1])[ok].all(), "one agent per cell (fixActions post-condition)"
        # no swaps: an agent never ends on the previous cell of an agent that ended on its own previous cell
        for w in np.flatnonzero(ok):
            prev = {tuple(p): i for i, p in enumerate(before[w])}
            for i in range(N):
                j = prev.get(tuple(pos[w, i]))
                if j is not None and j != i:
                    assert tuple(pos[w, j]) != tuple(before[w, i]), "swap collision executed"
        assert set(np.unique(out["status"])) <= {-4, -3, -2, -1, 1}
        assert np.isin(out["reward"][ok].view(np.uint32), np.array([-2, -0.35, -0.3, -0.5, np.float32(-0.35) + np.float32(1.5),
                       np.float32(-0.3) + np.float32(1.5)], dtype=np.float32).view(np.uint32)).all()
        assert ((out["cost"] >= 0) & (out["cost"] <= 1)).all()
        tv = out["train_valid"]
        assert set(np.unique(tv)) <= {0.0, 1.0}
        obs, vec = env.getAllObservations()
        assert set(np.unique(obs)) <= {0.0, 1.0} and (obs[:, :, 0, 4, 4] == 1).all() and (vec[..., 3] == 0).all()


@settings(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(0, 10_000), st.sampled_from([(7, 11), (12, 12), (24, 17), (40, 40)]), st.floats(0.0, 0.35))
def test_bfs_maps_are_distance_fields(seed, shape, dens):
    H, Wd = shape
    sc = random_scenario(4, H, Wd, 3, density=(dens, dens), queue_len=2, seed=seed)
    env = OracleMapfGym(sc, threads=1, use_tape=False)
    maps = env.bfs_maps().astype(np.int64)
    goals = env.state()["goal"]
    for w in range(4):
        free = sc.obst[w] == 0
        for i in range(3):
            m = maps[w, i]
            gr, gc = goals[w, i]
            assert m[gr, gc] == 0
            assert (m[~free & ~((np.arange(H)[:, None] == gr) & (np.arange(Wd)[None, :] == gc))] == -1).all()
            reach = m >= 0
            # 1-Lipschitz along free edges, and every reached cell other than the goal has a neighbour one step closer
            for dr, dc in ((0, 1), (1, 0)):
                a, b = m[:H - dr, :Wd - dc], m[dr:, dc:]
                both = (a >= 0) & (b >= 0)
                assert (np.abs(a - b)[both] <= 1).all()
                assert not ((a >= 0) & (b == -2)).any() and not ((a == -2) & (b >= 0)).any(), "reached next to unreached free cell"
            pad = np.full((H + 2, Wd + 2), 10 ** 6)
            pad[1:-1, 1:-1] = np.where(reach, m, 10 ** 6)
            nmin = np.minimum.reduce([pad[:-2, 1:-1], pad[2:, 1:-1], pad[1:-1, :-2], pad[1:-1, 2:]])
            inner = reach & (m > 0)
            assert (nmin[inner] == m[inner] - 1).all()


def test_oracle_flagged_worlds_stay_valid():
    """Worlds on which the reference would hang (fixActions livelock, mapf_gym.py:563) or raise (IndexError, :588) are flagged and
    every agent stays for that step: whatever the step, no two agents ever share a cell and nobody stands on an obstacle."""
    import numpy as np
    from oracle import OracleMapfGym
    from primal_ppo_b200 import random_actions, random_scenario
    W, H, Wd, N = 300, 6, 5, 9
    sc = random_scenario(W, H, Wd, N, density=(0.0, 0.35), queue_len=2, seed=12, unique_maps=24)
    orc = OracleMapfGym(sc, seed=3, threads=4, use_tape=False)
    acts = random_actions(40, W, N, seed=4)
    for t in range(40):
        before = orc.state()["pos"].copy()
        e0 = orc.state()["err"].copy()
        out = orc.step(acts[t])
        s = orc.state()
        p = s["pos"].astype(np.int64)
        cell = p[..., 0] * Wd + p[..., 1]
        assert all(len(set(r)) == N for r in cell.tolist()), f"t={t}: two agents on one cell"
        assert (sc.obst[np.arange(W)[:, None], p[..., 0], p[..., 1]] == 0).all()
        newly = (s["err"] & 3) & ~(e0 & 3)                       # NO_VIABLE / FIX_ITER_CAP raised in THIS step
        if newly.any():
            np.testing.assert_array_equal(s["pos"][newly != 0], before[newly != 0])     # nobody moved in such a world
            assert (out["fixed"][newly != 0] == 0).all()
    assert (orc.state()["err"] & 3).any(), "the scenario was meant to produce flagged worlds"
