#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE (authoring container only).

For every case below this script builds scenarios, drives the reference's ``FixedMapfGym`` with the
runner's call order (``runner.py:64-100``) and records, per step, every output of the hot path:
status, reward (+goal bonus), cost, trainValid, goalsReached, constraintsViolated, shadowGoals,
positions, goals, the actions after ``fixActions``, observations (bit-packed), vectors and BFS maps.
It also records the *tape* of the only non-deterministic-order branch of the env, ``fixActions`` branch 3
(``mapf_gym.py:587-598``): each ``random.choice`` result and the order in which conflicting agents are
evicted (Python-set iteration order), captured with ``sys.settrace`` on the live reference frame.

Worlds on which the reference itself raises (``IndexError`` from ``random.choice([])``,
``Exception('lets see')``) or livelocks in its ``while`` loop are dropped and re-drawn, and counted.

Usage:  python tests/golden/make_golden.py [case ...]
Output: tests/golden/<case>.npz   (+ tests/golden/gae_runner.npz from a real ``Runner.run()``)
"""
import os
import random
import signal
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_loader import load_reference  # noqa: E402
from primal_ppo_b200.scenario import Scenario, looping_trace, largest_component  # noqa: E402

GOAL_REWARD = 1.5   # alg_parameters.py:38, added by the runner (runner.py:89-91)


class _Timeout(Exception):
    pass


def _alarm(signum, frame):
    raise _Timeout()


class TapeRecorder:
    """Records fixActions branch-3 events from the live reference."""

    def __init__(self, mapf_gym):
        self.events = []
        self._mg = mapf_gym
        src = open(mapf_gym.__file__).read().split("\n")
        self.append_line = None
        for i, line in enumerate(src):
            if "problemAgents.append(conflict[0])" in line:
                self.append_line = i + 1
        assert self.append_line is not None
        self._orig_choice = random.choice

    def __enter__(self):
        def choice(seq):
            r = self._orig_choice(seq)
            self.events.append([int(r)])
            return r
        random.choice = choice
        code = self._mg.MapfGym.fixActions.__code__

        def tracer(frame, event, arg):
            if frame.f_code is not code:
                return None

            def local(frame, event, arg):
                if event == "line" and frame.f_lineno == self.append_line:
                    self.events[-1].append(int(frame.f_locals["conflict"][0]))
                return local
            return local
        sys.settrace(tracer)
        return self

    def __exit__(self, *a):
        sys.settrace(None)
        random.choice = self._orig_choice

    def flat(self):
        out = []
        for ev in self.events:
            out += [ev[0], len(ev) - 1] + ev[1:]
        return out


def _density_map(rng, H, Wd, dens):
    p = rng.uniform(dens[0], dens[1])
    return (rng.random((H, Wd)) < p)


def _draw_world(rng, case, mapf_gym):
    """Returns (obst_bool[H,Wd], starts[N], goal lists[N][Q], human spec)."""
    N, Q = case["N"], case["Q"]
    while True:
        if case["map"] == "warehouse":
            np.random.seed(int(rng.integers(1 << 31)))
            world = mapf_gym.generateWarehouse(num_block=case["size"])   # map_generator.py:127-138
            ob = world != 0
        else:
            H, Wd = case["size"]
            ob = _density_map(rng, H, Wd, case["density"])
        free = ~ob
        comp = largest_component(free)
        cells = np.argwhere(comp)
        if len(cells) < 4 or free.sum() < N + 3:
            continue
        allfree = np.argwhere(free)
        sel = rng.choice(len(cells), size=4, replace=False)
        hseq = [tuple(int(x) for x in cells[s]) for s in sel]
        clustered = rng.random() < case.get("cluster", 0.0)
        if clustered:
            # agents start packed around the human's start (nearest free cells first, ties shuffled): human collisions
            # (-2), the -2 -> -3 overwrite of getActionStatus (mapf_gym.py:460-472) and fixActions branch 3 need a crowd
            d = np.abs(allfree - np.asarray(hseq[0])).sum(1) + rng.random(len(allfree)) * 0.5
            perm = np.argsort(d, kind="stable")
        else:
            perm = rng.permutation(len(allfree))
        starts = []
        for k in perm:
            c = tuple(int(x) for x in allfree[k])
            if c != hseq[0]:
                starts.append(c)
            if len(starts) == N:
                break
        if len(starts) < N:
            continue
        goals = []
        for i in range(N):
            g, prev = [], starts[i]
            for q in range(Q):
                # mostly nearby goals so that arrivals happen within a short trace
                for _ in range(50):
                    if clustered and q == 0 and rng.random() < 0.8:
                        # first goal next to the human's goal: the crowd walks with the human
                        cand = (hseq[1][0] + int(rng.integers(-2, 3)), hseq[1][1] + int(rng.integers(-2, 3)))
                    elif rng.random() < case.get("near", 0.7):
                        cand = (prev[0] + int(rng.integers(-3, 4)), prev[1] + int(rng.integers(-3, 4)))
                    else:
                        cand = tuple(int(x) for x in allfree[rng.integers(len(allfree))])
                    if (0 <= cand[0] < ob.shape[0] and 0 <= cand[1] < ob.shape[1]
                            and free[cand] and cand != prev):
                        break
                else:
                    cand = tuple(int(x) for x in allfree[rng.integers(len(allfree))])
                g.append(cand)
                prev = cand
            goals.append(g)
        return ob, starts, goals, hseq


def run_case(name, case, seed, out_dir=None):
    N = case["N"]
    mapf_gym, util, AP = load_reference(N)
    rng = np.random.default_rng(seed)
    T, Wn, Q = case["T"], case["W"], case["Q"]
    C = case.get("C", 6)
    use_da, use_hp = case.get("use_da", False), case.get("use_hp", False)
    fixed_human = case.get("human", "loop") == "fixed"
    worlds = []
    dropped = {"IndexError": 0, "Exception": 0, "livelock": 0, "astar": 0}
    signal.signal(signal.SIGALRM, _alarm)
    t0 = time.time()
    while len(worlds) < Wn:
        ob, starts, goals, hseq = _draw_world(rng, case, mapf_gym)
        H, Wd = ob.shape
        obst_ref = -(ob.astype(np.int64))
        seqs = [util.Sequence(itemsIn=[starts[i]] + list(goals[i])) for i in range(N)]
        random.seed(int(rng.integers(1 << 31)))
        try:
            if fixed_human:
                env = mapf_gym.FixedMapfGym(obst_ref, seqs, hseq[0], hseq[1], numChannel=C,
                                            useDA=use_da, useHP=use_hp, humanSequence=list(hseq))
            else:
                env = mapf_gym.FixedMapfGym(obst_ref, seqs, hseq[0], hseq[1], numChannel=C,
                                            useDA=use_da, useHP=use_hp)
        except TypeError:
            dropped["astar"] += 1          # astar_4 *returns* ValueError when no path (astar_4.py:109)
            continue
        hp5 = np.full((5, 2), -1, dtype=np.int16)
        p5 = env.human.path[1:6]
        hp5[:len(p5)] = np.asarray(p5, dtype=np.int16).reshape(-1, 2)
        rec = dict(status=[], reward=[], cost=[], train_valid=[], goals_reached=[], violated=[],
                   shadow=[], pos=[], goal=[], fixed=[], obs=[], vec=[], actions=[], hticks=[], hp5t=[])

        def snap():
            rec["pos"].append(np.array([a.getPos() for a in env.agentList], dtype=np.int16))
            rec["goal"].append(np.array([a.getGoal() for a in env.agentList], dtype=np.int16))
            o, v = env.getAllObservations()
            assert o.dtype == np.float32 and v.dtype == np.float32
            assert np.all((o == 0) | (o == 1))
            rec["obs"].append(o[0].astype(np.uint8))
            rec["vec"].append(v[0].copy())
            rec["hticks"].append(list(env.human.getPos()) + list(env.human.getNextPos()))
            q5 = np.full((5, 2), -1, dtype=np.int16)
            pp = env.human.path[1:6]
            q5[:len(pp)] = np.asarray(pp, dtype=np.int16).reshape(-1, 2)
            rec["hp5t"].append(q5)
        snap()
        bfs0 = np.stack([a.bfsMap for a in env.agentList]).astype(np.int16)
        ok = True
        tape = TapeRecorder(mapf_gym)
        with tape:
            for t in range(T):
                acts = rng.integers(0, 5, size=N)
                # steer some agents down their own BFS field so that goals are actually reached
                for i, a in enumerate(env.agentList):
                    if rng.random() < case.get("greedy", 0.5):
                        r, c = a.getPos()
                        best, bd = None, a.bfsMap[r, c]
                        for k in range(1, 5):
                            dr, dc = a.dirDict[k]
                            rr, cc = r + dr, c + dc
                            if 0 <= rr < H and 0 <= cc < Wd and 0 <= a.bfsMap[rr, cc] < bd:
                                best, bd = k, a.bfsMap[rr, cc]
                        if best is not None:
                            acts[i] = best
                pre = np.array([a.getPos() for a in env.agentList])
                signal.alarm(5)
                try:
                    st = env.getActionStatus(acts)
                    rw, sg = env.calculateActionReward(acts, st)
                    cr = env.calculateCostReward(acts)
                    tv = env.getTrainValid(acts)
                    gr, cv = env.jointStep(acts, st)
                except _Timeout:
                    dropped["livelock"] += 1
                    ok = False
                except IndexError:
                    dropped["IndexError"] += 1
                    ok = False
                except Exception:
                    dropped["Exception"] += 1
                    ok = False
                finally:
                    signal.alarm(0)
                if not ok:
                    break
                for i, v in enumerate(gr):
                    if v == 1:
                        rw[0, i] += GOAL_REWARD          # runner.py:89-91
                post = np.array([a.getPos() for a in env.agentList])
                delta = post - pre
                fixed = np.zeros(N, dtype=np.int8)
                for i in range(N):
                    fixed[i] = [k for k in range(5) if tuple(a.dirDict[k]) == tuple(delta[i])][0]
                rec["actions"].append(acts.astype(np.int8))
                rec["status"].append(st.astype(np.int8))
                rec["reward"].append(rw[0].copy())
                rec["cost"].append(cr[0].copy())
                rec["train_valid"].append(tv.copy())
                rec["goals_reached"].append(gr.astype(np.uint8))
                rec["violated"].append(cv.astype(np.uint8))
                rec["shadow"].append(int(sg))
                rec["fixed"].append(fixed)
                snap()
        if not ok:
            continue
        bfsT = np.stack([a.bfsMap for a in env.agentList]).astype(np.int16)
        hticks = np.asarray(rec["hticks"], dtype=np.int16)          # [T+1,4]
        if fixed_human:
            htrace = hticks                                          # raw tick trace, never wraps
            hp5 = np.asarray(rec["hp5t"], dtype=np.int16)            # path[1:6] changes with every new path
        else:
            htrace = looping_trace(env.human.path)
            L = htrace.shape[0]
            for t in range(T + 1):                                   # the loop model must reproduce the ticks
                assert np.array_equal(htrace[t % L], hticks[t]), (t, htrace[t % L], hticks[t])
        worlds.append(dict(obst=ob.astype(np.uint8), starts=np.asarray(starts, dtype=np.int16),
                           goals=np.asarray(goals, dtype=np.int16), htrace=htrace, hp5=hp5,
                           tape=np.asarray(tape.flat(), dtype=np.int8), bfs0=bfs0, bfsT=bfsT,
                           **{k: np.asarray(v) for k, v in rec.items() if k not in ("hticks", "hp5t")}))
    # ---- stack worlds --------------------------------------------------------------------
    if case["map"] == "warehouse":
        # warehouse sizes differ per world: the batched env needs one (H,Wd) -> pad with obstacles
        Hm = max(w["obst"].shape[0] for w in worlds)
        Wm = max(w["obst"].shape[1] for w in worlds)
        # padded cells are obstacles in storage and out of bounds for that world (Scenario.dims);
        # BFS maps hold -1 there.
    else:
        Hm, Wm = case["size"]
    W = len(worlds)
    obst = np.ones((W, Hm, Wm), dtype=np.uint8)
    sizes = np.zeros((W, 2), dtype=np.int16)
    L = max(w["htrace"].shape[0] for w in worlds)
    TL = max(1, max(len(w["tape"]) for w in worlds))
    htrace = np.zeros((W, L, 4), dtype=np.int16)
    hlen = np.zeros((W,), dtype=np.int32)
    tape = np.zeros((W, TL), dtype=np.int8)
    tape_len = np.zeros((W,), dtype=np.int32)
    WB = min(W, case.get("bfs_worlds", W))     # BFS maps are the bulk of a fixture: large cases keep the first WB worlds'
    bfs0 = np.full((WB, N, Hm, Wm), -1, dtype=np.int16)
    bfsT = np.full((WB, N, Hm, Wm), -1, dtype=np.int16)
    for k, w in enumerate(worlds):
        h, wd = w["obst"].shape
        sizes[k] = (h, wd)
        obst[k, :h, :wd] = w["obst"]
        htrace[k, :w["htrace"].shape[0]] = w["htrace"]
        hlen[k] = w["htrace"].shape[0]
        tape[k, :len(w["tape"])] = w["tape"]
        tape_len[k] = len(w["tape"])
        if k < WB:
            bfs0[k, :, :h, :wd] = w["bfs0"]
            bfsT[k, :, :h, :wd] = w["bfsT"]
    sc = Scenario(obst=obst, starts=np.stack([w["starts"] for w in worlds]),
                  goal_queue=np.stack([w["goals"] for w in worlds]), htrace=htrace, hlen=hlen,
                  hp5=np.stack([w["hp5"] for w in worlds]) if worlds[0]["hp5"].ndim == 2 else
                  np.stack([np.concatenate([w["hp5"], np.repeat(w["hp5"][-1:], L - w["hp5"].shape[0], 0)])
                            for w in worlds]), tape=tape, tape_len=tape_len,
                  dims=sizes if case["map"] == "warehouse" else None,
                  fov=9, num_channel=C, use_da=use_da, use_hp=use_hp)
    sc.validate()

    def st(key, axis=1):
        return np.stack([w[key] for w in worlds], axis=axis)
    obs = st("obs")                                                 # [T+1,W,N,C,F,F] u8
    out = sc.to_npz_dict()
    out.update(sizes=sizes, actions=st("actions"), status=st("status"), reward=st("reward"),
               cost=st("cost"), train_valid=st("train_valid"), goals_reached=st("goals_reached"),
               violated=st("violated"), shadow=st("shadow"), pos=st("pos"), goal=st("goal"),
               fixed=st("fixed"), vec=st("vec"), obs_shape=np.asarray(obs.shape, dtype=np.int64),
               obs_bits=np.packbits(obs.reshape(-1)), bfs0=bfs0, bfsT=bfsT)
    path = os.path.join(out_dir or HERE, name + ".npz")
    np.savez_compressed(path, **out)
    nev = int(sum((w["tape"].size > 0) for w in worlds))
    print(f"{name}: {W} worlds, T={T}, N={N}, {Hm}x{Wm}; dropped {dropped}; "
          f"worlds with tape events {nev}; goals reached {int(out['goals_reached'].sum())}; "
          f"status counts { {int(s): int((out['status'] == s).sum()) for s in (-1, -2, -3, -4, 1)} }; "
          f"{os.path.getsize(path) / 1024:.0f} KiB; {time.time() - t0:.1f}s")


CASES = {
    # BASELINE.json config 1 shape
    "g_10x10_n8": dict(map="density", size=(10, 10), density=(0.2, 0.2), N=8, W=24, T=40, Q=12, seed=1),
    # BASELINE.json config 2 shape
    "g_20x20_n8": dict(map="density", size=(20, 20), density=(0.2, 0.2), N=8, W=12, T=40, Q=12, seed=2),
    # BASELINE.json config 3 shape
    # BASELINE.json config 3 shape (the headline): 64 worlds x 64 steps, half of them with the agents packed around the
    # human so that status -2, the -2 -> -3 overwrite and the fixActions tape all occur at this shape
    "g_40x40_n32": dict(map="density", size=(40, 40), density=(0.0, 0.3), N=32, W=64, T=64, Q=8, seed=3, cluster=0.5,
                        bfs_worlds=8),
    # crowded: exercises status -3 overwrite and the fixActions tape
    "g_8x8_n8_dense": dict(map="density", size=(8, 8), density=(0.25, 0.3), N=8, W=40, T=30, Q=10, seed=4,
                           greedy=0.6),
    # the reference's own default map family, non-square, fixed-path human, eval channels on
    "g_warehouse_n4_eval": dict(map="warehouse", size=(10, 14), N=4, W=8, T=60, Q=10, seed=5, human="fixed",
                                use_da=True, use_hp=True),
    # 5-channel nets (NUM_CHANNEL without trajectory prediction) + exhausted goal queues
    "g_12x9_n6_c5": dict(map="density", size=(12, 9), density=(0.1, 0.2), N=6, W=10, T=50, Q=2, seed=6, C=5,
                         greedy=0.8, near=1.0),
}


def make_gae_golden():
    """A real ``Runner.run()`` (``runner.py:26-151``) with the reference net on CPU: pins the GAE scan."""
    mapf_gym, util, AP = load_reference(2)
    import torch
    import runner as ref_runner
    import model as ref_model
    util.set_global_seeds(7)
    AP.TrainingParameters.N_STEPS = 64
    captured = {}
    orig_value = ref_model.Model.value

    def value(self, *a, **k):
        out = orig_value(self, *a, **k)
        captured["last"] = [np.array(x) for x in out]
        return out
    ref_model.Model.value = value
    try:
        r = ref_runner.Runner(1)
        weights = r.local_model.network.state_dict()
        mb, perf = r.run(weights)
    finally:
        ref_model.Model.value = orig_value
        AP.TrainingParameters.N_STEPS = 2 ** 8
    last_v, last_cv = np.squeeze(np.array(captured["last"]))
    np.savez_compressed(os.path.join(HERE, "gae_runner.npz"),
                        rewards=mb.rewards, values=mb.values, cost_rewards=mb.costRewards,
                        cost_values=mb.costValues, last_values=last_v.astype(np.float32),
                        last_cost_values=last_cv.astype(np.float32), returns=mb.returns,
                        cost_returns=mb.costReturns, gamma=np.float64(AP.TrainingParameters.GAMMA),
                        lam=np.float64(AP.TrainingParameters.LAM))
    print("gae_runner:", mb.rewards.shape, mb.rewards.dtype, mb.values.dtype, mb.returns.dtype)


if __name__ == "__main__":
    names = sys.argv[1:] or (list(CASES) + ["gae_runner"])
    for n in names:
        if n == "gae_runner":
            make_gae_golden()
        else:
            run_case(n, CASES[n], CASES[n]["seed"])
