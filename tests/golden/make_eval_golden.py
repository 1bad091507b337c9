#!/usr/bin/env python
"""Fixtures for the eval-fixture path, produced by the UNMODIFIED reference in this container:

  tests/golden/fixed_episode_infos/      written by the reference's own generateFixedEpisodeInfos + saveFixedEpisodeInfos
                                         (evaluate.py:48-121) for 4 episodes, 3 agents, MAX_STEPS = 48
  tests/golden/eval_episodes.npz         for both human movement types (evaluate.py:216-218): FixedMapfGym driven from the
                                         loaded fixture with seeded random actions; per step the human's (pos, next),
                                         agent cells, rewards, status, and packed observations (useDA = useHP = True)
  tests/golden/astar_paths.npz           astar_4 paths (astar_4.py:21-109) on random and warehouse maps
Run from the repo root: python tests/golden/make_eval_golden.py"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
N_AGENTS, EPISODES, MAX_STEPS = 3, 4, 48

if __name__ == "__main__":
    from ref_loader import load_reference
    mapf_gym, util, AP = load_reference(N_AGENTS)
    if "setproctitle" not in sys.modules:
        try:
            import setproctitle  # noqa: F401
        except Exception:
            sys.modules["setproctitle"] = types.ModuleType("setproctitle")
    AP.EvalParameters.N_AGENTS = N_AGENTS
    AP.EvalParameters.EPISODES = EPISODES
    AP.EvalParameters.MAX_STEPS = MAX_STEPS
    AP.EvalParameters.FIXED_EPISODE_INFOS_PATH = os.path.join(HERE, "fixed_episode_infos")
    AP.EnvParameters.WORLD_SIZE = (10, 16)
    import evaluate as ref_eval
    import astar_4 as ref_astar
    np.random.seed(7); random.seed(7)
    infos = ref_eval.generateFixedEpisodeInfos()
    ref_eval.saveFixedEpisodeInfos(infos)
    infos = ref_eval.loadFixedEpisodeInfos()

    rec = {}
    rng = np.random.default_rng(3)
    actions = rng.integers(0, 5, size=(MAX_STEPS, EPISODES, N_AGENTS)).astype(np.int8)
    rec["actions"] = actions
    for mtype in (0, 1):
        hum, pos, rew, stat, obs_bits, vecs = [], [], [], [], [], []
        for e in range(EPISODES):
            np.random.seed(100 + e); random.seed(100 + e)
            seqs = [util.Sequence(itemsIn=list(s.items)) for s in infos["agentsSequence"][e]]
            hs = None if mtype == 0 else infos["humanSequence"][e]
            env = mapf_gym.FixedMapfGym(infos["obstacleMap"][e], seqs, infos["humanStart"][e], infos["humanGoal"][e],
                                        numChannel=6, useDA=True, useHP=True, humanSequence=hs)
            h_e, p_e, r_e, s_e, o_e, v_e = [], [], [], [], [], []

            def snap():
                h_e.append(list(env.human.getPos()) + list(env.human.getNextPos()))
                p_e.append([list(a.getPos()) for a in env.agentList])
                o, v = env.getAllObservations()
                o_e.append(np.packbits(o.astype(np.uint8).ravel())); v_e.append(v[0].copy())
            snap()
            for t in range(MAX_STEPS):
                a = actions[t, e].astype(np.float64)
                try:
                    st = env.getActionStatus(a)
                    r, _ = env.calculateActionReward(a, st)
                    g, _ = env.jointStep(a, st)
                except Exception as ex:          # a crash of the reference ends the comparable part of this episode
                    print("episode", e, "mtype", mtype, "stopped at", t, type(ex).__name__)
                    break
                r[0, g == 1] += 1.5
                r_e.append(r[0].copy()); s_e.append(st.copy())
                snap()
            hum.append(h_e); pos.append(p_e); rew.append(r_e); stat.append(s_e); obs_bits.append(o_e); vecs.append(v_e)
        T = min(len(x) for x in rew)
        rec[f"m{mtype}_T"] = np.int32(T)
        rec[f"m{mtype}_human"] = np.array([[h[t] for h in hum] for t in range(T + 1)], dtype=np.int16)
        rec[f"m{mtype}_pos"] = np.array([[p[t] for p in pos] for t in range(T + 1)], dtype=np.int16)
        rec[f"m{mtype}_reward"] = np.array([[r[t] for r in rew] for t in range(T)], dtype=np.float32)
        rec[f"m{mtype}_status"] = np.array([[s[t] for s in stat] for t in range(T)], dtype=np.int8)
        for e in range(EPISODES):          # maps differ in size per episode -> observations are stored per episode
            rec[f"m{mtype}_obs_e{e}"] = np.stack([obs_bits[e][t] for t in range(T + 1)])
            rec[f"m{mtype}_vec_e{e}"] = np.stack([vecs[e][t] for t in range(T + 1)]).astype(np.float32)
        print("mtype", mtype, "T", T)
    np.savez_compressed(os.path.join(HERE, "eval_episodes.npz"), **rec)

    # astar paths
    ap = {}
    rng = np.random.default_rng(11)
    k = 0
    for trial in range(400):
        if trial % 2 == 0:
            H, Wd = int(rng.integers(5, 24)), int(rng.integers(5, 24))
            world = -(rng.random((H, Wd)) < rng.uniform(0.0, 0.35)).astype(int)
        else:
            import map_generator
            L = int(rng.integers(6, 30))
            world = map_generator.generateWarehouse(num_block=[L, L])
        free = np.argwhere(world == 0)
        if len(free) < 2:
            continue
        a, b = rng.choice(len(free), size=2, replace=False)
        s, g = tuple(int(x) for x in free[a]), tuple(int(x) for x in free[b])
        out = ref_astar.astar_4(world, s, g)
        if isinstance(out, ValueError):
            path = np.zeros((0, 2), dtype=np.int16)
        else:
            path = np.asarray(out[0][::-1], dtype=np.int16).reshape(-1, 2)        # start -> goal
        ap[f"world{k}"] = world.astype(np.int8); ap[f"sg{k}"] = np.array([s, g], dtype=np.int16); ap[f"path{k}"] = path
        k += 1
    ap["count"] = np.int32(k)
    np.savez_compressed(os.path.join(HERE, "astar_paths.npz"), **ap)
    print("astar paths:", k, "of which unreachable:", sum(ap[f"path{i}"].shape[0] == 0 for i in range(k)))
