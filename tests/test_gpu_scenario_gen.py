"""On-device scenario generation: structural invariants, the warehouse layout against the
reference's generateWarehouse, distribution checks, and a generated batch stepped on the GPU vs the oracle."""
import os
from collections import deque

import numpy as np
import pytest
import torch

from oracle import OracleMapfGym
from primal_ppo_b200 import random_actions

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bfs(free, src):
    H, Wd = free.shape
    dist = np.full((H, Wd), -1, dtype=np.int32)
    dist[src] = 0
    q = deque([src])
    while q:
        r, c = q.popleft()
        for dr, dc in ((0, 1), (1, 0), (0, -1), (-1, 0)):
            rr, cc = r + dr, c + dc
            if 0 <= rr < H and 0 <= cc < Wd and free[rr, cc] and dist[rr, cc] < 0:
                dist[rr, cc] = dist[r, c] + 1
                q.append((rr, cc))
    return dist


def _check_invariants(sc, human_loops=1):
    W, H, Wd, N = sc.num_worlds, sc.height, sc.width, sc.num_agents
    err = sc.meta["gen_err"]
    for w in range(W):
        if err[w] & ~4:
            continue
        rows, cols = sc.dims[w]
        free = sc.obst[w] == 0
        assert not free[rows:, :].any() and not free[:, cols:].any()
        st = sc.starts[w]; g0 = sc.goal_queue[w, :, 0]
        assert free[st[:, 0], st[:, 1]].all() and free[g0[:, 0], g0[:, 1]].all()
        hstart = tuple(sc.htrace[w, 0, :2])
        cells = [tuple(x) for x in st] + [tuple(x) for x in g0] + [hstart]
        assert len(set(cells)) == 2 * N + 1, "starts, first goals and the human's cell are pairwise distinct"
        q = sc.goal_queue[w]
        assert free[q[..., 0], q[..., 1]].all()
        assert (np.abs(np.diff(q.astype(np.int32), axis=1)).sum(-1) > 0).all(), "consecutive goals differ"
        # human: starts on row 0 / column 0, walks 4-connected over free cells, next = following pos, out-and-back shortest
        L = int(sc.hlen[w])
        tr = sc.htrace[w, :L].astype(np.int32)
        assert hstart[0] == 0 or hstart[1] == 0
        assert free[tr[:, 0], tr[:, 1]].all()
        if err[w] & 4:
            assert L == 1
            continue
        assert (np.abs(np.diff(tr[:, :2], axis=0)).sum(-1) <= 1).all()
        assert (tr[:-1, 2:] == tr[1:, :2]).all() and (tr[-1, 2:] == tr[-1, :2]).all() or human_loops > 1
        if human_loops == 1:
            assert L % 2 == 1
            d = (L - 1) // 2
            assert tuple(tr[-1, :2]) == hstart and (tr[:d + 1, :2] == tr[::-1][:d + 1, :2]).all()
            assert _bfs(free, hstart)[tuple(tr[d, :2])] == d, "the walk to the goal is a shortest path"
            assert (np.abs(np.diff(tr[:, :2], axis=0)).sum(-1) == 1).all()
            k = min(5, L - 1)
            assert (sc.hp5[w, :k] == tr[1:1 + k, :2]).all() and (sc.hp5[w, k:] == -1).all()


def _host(dsc):
    sc = dsc.to_host()
    sc.meta["gen_err"] = dsc.gen_err.cpu().numpy()
    return sc


def test_gpu_generated_warehouse_matches_reference_layout():
    from primal_ppo_b200 import generate_scenario_device
    ref = np.load(os.path.join(GOLDEN, "warehouse_maps.npz"))
    d = generate_scenario_device(600, 40, 60, 8, kind="warehouse", size_range=(10, 40), queue_len=4, seed=3)
    sc = _host(d)
    assert not (sc.meta["gen_err"] & ~4).any()
    lengths = set()
    for w in range(sc.num_worlds):
        rows, cols = sc.dims[w]
        lengths.add(int(rows))
        m = ref[f"L{rows}"]
        assert m.shape == (rows, cols)
        np.testing.assert_array_equal(sc.obst[w, :rows, :cols], m)
    assert min(lengths) == 10 and max(lengths) == 40 and len(lengths) >= 28       # np.random.randint(10, 41)
    _check_invariants(sc)


@pytest.mark.parametrize("tri", [False, True])
def test_gpu_generated_density_maps(tri):
    from primal_ppo_b200 import generate_scenario_device
    d = generate_scenario_device(1024, 40, 40, 32, kind="density", density=(0.0, 0.3), triangular=tri, queue_len=6, seed=11)
    sc = _host(d)
    dens = sc.obst.reshape(1024, -1).mean(1)
    assert 0.0 <= dens.min() and dens.max() < 0.36
    # U[0, .3] has mean .15; triangular(0, .198, .3) has mean (0 + .198 + .3) / 3 = .166
    assert abs(dens.mean() - (0.166 if tri else 0.15)) < 0.012, dens.mean()
    ok = (sc.meta["gen_err"] & ~4) == 0
    assert ok.mean() > 0.97
    _check_invariants(sc)
    # per-cell obstacle frequency is uniform over the grid (no positional bias)
    cellfreq = sc.obst.mean(0)
    assert abs(cellfreq.mean() - dens.mean()) < 1e-6 and cellfreq.std() < 0.03
    # starts are spread over the whole grid
    occ = np.zeros((40, 40)); np.add.at(occ, (sc.starts[..., 0].ravel(), sc.starts[..., 1].ravel()), 1)
    assert (occ > 0).mean() > 0.99
    # same seed -> same scenario; world_offset shifts the stream (sharded jobs)
    d2 = generate_scenario_device(1024, 40, 40, 32, kind="density", density=(0.0, 0.3), triangular=tri, queue_len=6, seed=11)
    assert torch.equal(d.obst, d2.obst) and torch.equal(d.goal_queue, d2.goal_queue) and torch.equal(d.htrace, d2.htrace)
    d3 = generate_scenario_device(512, 40, 40, 32, kind="density", density=(0.0, 0.3), triangular=tri, queue_len=6, seed=11,
                                  world_offset=512)
    assert torch.equal(d.obst[512:], d3.obst) and torch.equal(d.starts[512:], d3.starts) and torch.equal(d.htrace[512:], d3.htrace)


def test_gpu_generated_random_sizes_and_human_reloops():
    from primal_ppo_b200 import generate_scenario_device
    d = generate_scenario_device(400, 40, 40, 6, kind="density", density=(0.0, 0.25), size_range=(10, 40), queue_len=3,
                                 human_loops=3, seed=5)
    sc = _host(d)
    sides = sc.dims[:, 0]
    assert set(np.unique(sides)) == {10, 25, 40} and (sc.dims[:, 0] == sc.dims[:, 1]).all()
    frac = [(sides == s).mean() for s in (10, 25, 40)]
    assert abs(frac[0] - 0.5) < 0.08 and abs(frac[1] - 0.25) < 0.07          # np.random.choice p=[.5,.25,.25]
    _check_invariants(sc, human_loops=3)
    # several walks: the trace returns to the entrance more than once
    w = int(np.argmax(sc.hlen))
    tr = sc.htrace[w, :sc.hlen[w], :2]
    assert ((tr == tr[0]).all(1)).sum() >= 3


def test_gpu_env_on_generated_scenario_matches_oracle():
    """A generated batch (arrays stay in HBM) stepped by the GPU env == the oracle on the downloaded arrays."""
    from primal_ppo_b200 import BatchedMapfGym, generate_scenario_device
    for kind, H, Wd, N in (("warehouse", 40, 60, 8), ("density", 20, 20, 8)):
        d = generate_scenario_device(256, H, Wd, N, kind=kind, density=(0.05, 0.25), queue_len=4, seed=21)
        env = BatchedMapfGym(d, use_tape=False, seed=9)
        sc = d.to_host()
        sc.validate()
        orc = OracleMapfGym(sc, seed=9, threads=4, use_tape=False)
        acts = random_actions(24, 256, N, seed=2)
        for t in range(24):
            out, obs, vec = env.step_observe(torch.from_numpy(acts[t]))
            ref = orc.step(acts[t])
            ok = orc.state()["err"] == 0
            for key in ("status", "reward", "cost", "train_valid", "goals_reached", "violated"):
                assert np.array_equal(getattr(out, key).cpu().numpy()[ok], ref[key][ok]), (kind, t, key)
            o_obs, o_vec = orc.getAllObservations()
            assert np.array_equal(obs.cpu().numpy()[ok], o_obs[ok]) and np.array_equal(vec.cpu().numpy()[ok], o_vec[ok])
        assert np.array_equal(env.bfs_maps().cpu().numpy()[ok], orc.bfs_maps()[ok])
