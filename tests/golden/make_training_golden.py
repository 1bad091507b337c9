#!/usr/bin/env python
"""tests/golden/training_env.npz — the reference's ACTUAL training environment, `MapfGym()` (`mapf_gym.py:163-173`:
random warehouse, `Human` that re-draws its goal after every out-and-back walk, agent starts / goals from
`getFreeCell`), driven for T steps with seeded random actions through the runner's call order (`runner.py:64-100`).
Everything the env draws from the global RNGs is recorded as it happens (goals handed out per agent, the human's
(pos, next) per tick), which is exactly what a `Scenario` carries; every per-step output is recorded for comparison.
Run from the repo root: python tests/golden/make_training_golden.py"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
T = 160
CASES = [(2, 6, 11), (5, 5, 12)]          # (N_AGENTS, worlds, seed)

if __name__ == "__main__":
    from ref_loader import load_reference
    rec = {"T": np.int32(T), "n_cases": np.int32(len(CASES))}
    for ci, (N, W, seed) in enumerate(CASES):
        mapf_gym, util, AP = load_reference(N)
        rng = np.random.default_rng(seed)
        worlds = []
        for w in range(W):
            while True:
                np.random.seed(seed * 1000 + w * 7 + len(worlds)); random.seed(seed * 1000 + w)
                try:
                    env = mapf_gym.MapfGym(num_agents=N, size=(10, 22))
                except Exception as ex:               # astar failure on a walled-off goal etc.: re-draw
                    seed += 1
                    continue
                break
            handed = [[] for _ in range(N)]
            orig = env.getNextGoal

            def logging_next_goal(worldMap, agentId=None, _orig=orig, _handed=handed):
                g = _orig(worldMap, agentId)
                _handed[agentId].append((int(g[0]), int(g[1])))
                return g
            env.getNextGoal = logging_next_goal
            d = dict(obst=(env.obstacleMap != 0).astype(np.uint8), starts=[tuple(int(x) for x in a.getPos()) for a in env.agentList],
                     goals0=[tuple(int(x) for x in a.getGoal()) for a in env.agentList], human=[], pos=[], goal=[], status=[],
                     reward=[], cost=[], tv=[], gr=[], cv=[], shadow=[], obs=[], vec=[])

            def snap():
                d["human"].append(list(env.human.getPos()) + list(env.human.getNextPos()))
                d["pos"].append([list(a.getPos()) for a in env.agentList])
                d["goal"].append([list(a.getGoal()) for a in env.agentList])
                o, v = env.getAllObservations()
                d["obs"].append(np.packbits(o.astype(np.uint8).ravel())); d["vec"].append(v[0].copy())
            snap()
            acts = rng.integers(0, 5, size=(T, N)).astype(np.int8)
            ok = True
            for t in range(T):
                a = acts[t].astype(np.float64)
                try:
                    st = env.getActionStatus(a)
                    r, sg = env.calculateActionReward(a, st)
                    c = env.calculateCostReward(a)
                    tv = env.getTrainValid(a)
                    g, cv = env.jointStep(a, st)
                except Exception as ex:
                    print("case", ci, "world", w, "reference raised", type(ex).__name__, "at", t)
                    ok = False
                    break
                r[0, g == 1] += 1.5
                d["status"].append(st.copy()); d["reward"].append(r[0].copy()); d["cost"].append(c[0].copy()); d["tv"].append(tv.copy())
                d["gr"].append(g.copy()); d["cv"].append(cv.copy()); d["shadow"].append(sg)
                snap()
            assert ok, "re-run with another seed"
            d["acts"] = acts
            d["handed"] = handed
            worlds.append(d)
        H = max(x["obst"].shape[0] for x in worlds); Wd = max(x["obst"].shape[1] for x in worlds)
        Q = max(1 + len(h) for x in worlds for h in x["handed"])
        obst = np.ones((W, H, Wd), dtype=np.uint8); dims = np.zeros((W, 2), dtype=np.int16)
        queue = np.zeros((W, N, Q, 2), dtype=np.int16)
        for w, x in enumerate(worlds):
            h, wd = x["obst"].shape
            obst[w, :h, :wd] = x["obst"]; dims[w] = (h, wd)
            for i in range(N):
                seq = [x["goals0"][i]] + x["handed"][i]
                queue[w, i, :len(seq)] = np.asarray(seq, dtype=np.int16)
                queue[w, i, len(seq):] = seq[-1]
        p = f"c{ci}_"
        rec[p + "obst"] = obst; rec[p + "dims"] = dims; rec[p + "goal_queue"] = queue
        rec[p + "starts"] = np.array([x["starts"] for x in worlds], dtype=np.int16)
        rec[p + "htrace"] = np.array([x["human"] for x in worlds], dtype=np.int16)            # [W, T+1, 4]
        rec[p + "actions"] = np.stack([x["acts"] for x in worlds], axis=1)                     # [T, W, N]
        for key, name, dt in (("pos", "pos", np.int16), ("goal", "goal", np.int16), ("status", "status", np.int8),
                              ("reward", "reward", np.float32), ("cost", "cost", np.float32), ("tv", "train_valid", np.float32),
                              ("gr", "goals_reached", np.uint8), ("cv", "violated", np.uint8), ("shadow", "shadow", np.int32),
                              ("vec", "vec", np.float32)):
            rec[p + name] = np.stack([np.asarray(x[key]) for x in worlds], axis=1).astype(dt)  # [T(+1), W, ...]
        for w, x in enumerate(worlds):
            rec[p + f"obs_w{w}"] = np.stack(x["obs"])
        print("case", ci, "N", N, "W", W, "dims", dims.tolist(), "goals handed", [sum(len(h) for h in x["handed"]) for x in worlds])
    np.savez_compressed(os.path.join(HERE, "training_env.npz"), **rec)
