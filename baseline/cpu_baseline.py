#!/usr/bin/env python
"""Time the UNMODIFIED Python reference env (baseline/_ref/mapf_gym.py) on this box's host cores.

P forked worker processes (P = cores this process may use), one `FixedMapfGym` each, driven through the rollout loop's own
call order (runner.py:64-100: getActionStatus, calculateActionReward, calculateCostReward, getTrainValid, jointStep,
getAllObservations) with uniform random actions on worlds drawn by the same generator bench.py uses, for a fixed wall
budget.  Worlds on which the reference raises (IndexError from random.choice([]), Exception('lets see')) or livelocks in
fixActions are replaced and counted.  Prints / returns aggregate and per-core agent-steps/s.

This is test / measurement infrastructure: nothing under primal_ppo_b200/ imports it."""
import json
import multiprocessing as mp
import os
import signal
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "mapf_gym.py"))


def _load(n_agents):
    """The stubs of tests/golden/ref_loader.py (skimage / imageio / matplotlib / ray are imported at module scope by the
    reference but never used on the env path), pointed at baseline/_ref."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for n in ["skimage", "skimage.measure", "skimage.morphology", "imageio", "matplotlib", "matplotlib.colors"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["skimage"].morphology = sys.modules["skimage.morphology"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    if not hasattr(sys.modules["matplotlib.colors"], "hsv_to_rgb"):
        sys.modules["matplotlib.colors"].hsv_to_rgb = lambda x: x
    for n in ("wandb", "cv2"):                       # imported by util.py for logging / rendering only
        try:
            __import__(n)
        except Exception:
            sys.modules[n] = types.ModuleType(n)
    os.environ.setdefault("WANDB_MODE", "disabled")
    import alg_parameters as AP
    AP.EnvParameters.N_AGENTS = int(n_agents)
    import mapf_gym
    import util
    return mapf_gym, util


class _Timeout(Exception):
    pass


def _alarm(signum, frame):
    raise _Timeout()


def _worker(rank, size, n_agents, density, budget_s, seed, q):
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[rank % len(os.sched_getaffinity(0))]})
    except Exception:
        pass
    sys.path.insert(0, ROOT)
    import io
    import contextlib
    from primal_ppo_b200.scenario import random_scenario
    mapf_gym, util = _load(n_agents)
    rng = np.random.default_rng(seed + rank)
    signal.signal(signal.SIGALRM, _alarm)
    steps = dropped = worlds = 0
    sink = io.StringIO()
    t_env = 0.0
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end:
        sc = random_scenario(1, size, size, n_agents, density=density, queue_len=16, seed=int(rng.integers(1 << 30)))
        ob = -(sc.obst[0].astype(np.int64))
        seqs = [util.Sequence(itemsIn=[tuple(int(x) for x in sc.starts[0, i])] + [tuple(int(x) for x in g) for g in sc.goal_queue[0, i]])
                for i in range(n_agents)]
        hs = tuple(int(x) for x in sc.htrace[0, 0, :2])
        hg = tuple(int(x) for x in sc.htrace[0, int(sc.hlen[0]) // 2, :2])
        try:
            with contextlib.redirect_stdout(sink):
                env = mapf_gym.FixedMapfGym(ob, seqs, hs, hg)
        except Exception:
            dropped += 1
            continue
        worlds += 1
        for t in range(64):                                     # 64 steps per world, then a fresh world
            if time.perf_counter() >= t_end:
                break
            acts = rng.integers(0, 5, size=n_agents)
            t0 = time.perf_counter()
            signal.alarm(5)
            try:
                with contextlib.redirect_stdout(sink):
                    st = env.getActionStatus(acts)
                    env.calculateActionReward(acts, st)
                    env.calculateCostReward(acts)
                    env.getTrainValid(acts)
                    env.jointStep(acts, st)
                    env.getAllObservations()
            except (_Timeout, IndexError, Exception):
                dropped += 1
                break
            finally:
                signal.alarm(0)
            t_env += time.perf_counter() - t0
            steps += 1
        sink.seek(0); sink.truncate(0)
    q.put((steps, t_env, worlds, dropped))


def run(size=40, n_agents=32, density=(0.0, 0.3), budget_s=6.0, procs=None, seed=0):
    """Returns a dict: aggregate agent-steps/s over `procs` worker processes (env time only: scenario construction between
    worlds is excluded), per-core figure, counts."""
    if not available():
        return {"unavailable": "baseline/_ref is missing (run `python baseline/install_ref.py` where /root/reference exists)"}
    P = procs or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, size, n_agents, density, budget_s, seed, q)) for r in range(P)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    res = [q.get(timeout=budget_s * 4 + 120) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    per_core = [r[0] * n_agents / r[1] for r in res if r[1] > 0]
    return {"config": f"{size}x{size}, {n_agents} agents, density U{list(density)}, FixedMapfGym + looping human, uniform random actions",
            "processes": P, "env_steps": steps, "agent_steps_per_s_per_core": float(np.mean(per_core)) if per_core else 0.0,
            "agent_steps_per_s": float(np.sum(per_core)) if per_core else 0.0, "wall_s": wall,
            "worlds": sum(r[2] for r in res), "dropped": sum(r[3] for r in res),
            "impl": "UNMODIFIED reference mapf_gym.py (baseline/_ref), 5-call step + getAllObservations (runner.py:64-100)"}


if __name__ == "__main__":
    b = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
    print(json.dumps({"40x40x32": run(40, 32, (0.0, 0.3), b), "10x10x8": run(10, 8, (0.2, 0.2), b)}, indent=1))
