#!/usr/bin/env python
"""Randomised differential soak: random shapes (grid, agents, FOV, channels, density, goal sampling on/off, eval channels), the
CUDA path (fused launch or the split-phase host call with the compact slab, two launches, five-call API in rotation; BFS maps
refreshed in place after every step) against the oracle on every output of every step, for EVERY
world — also those carrying error flags (where the reference would have hung or raised, every agent stays for that step).
SOAK_CROWDED=1 draws small dense worlds (fixActions queues, evictions, livelock caps, no-free-cell flags).
    python tests/soak_parity.py [seconds] [seed]        (a script, not collected by pytest; it lives under tests/ because it drives the oracle)
Prints one line per scenario and a summary; exits non-zero on the first mismatch."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
from oracle import OracleMapfGym  # noqa: E402
from primal_ppo_b200 import BatchedMapfGym, random_actions, random_scenario  # noqa: E402
from primal_ppo_b200.build import build  # noqa: E402

build()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
t_end = time.time() + budget
keys = ("status", "reward", "cost", "train_valid", "goals_reached", "violated")
n_scen = n_steps = n_agent_steps = 0
while time.time() < t_end:
    if os.environ.get("SOAK_CROWDED"):                        # small dense worlds: fixActions queue, evictions, livelock cap, no-viable flags
        H = int(rng.integers(5, 15)); Wd = int(rng.integers(5, 15))
        N = int(rng.integers(2, max(3, min(48, H * Wd // 3))))
    else:
        N = int(rng.choice([1, 2, 3, 5, 8, 9, 12, 16, 17, 24, 31, 32, 33, 40, 48, 64, 100, 128]))
        H = int(rng.integers(6, 97)); Wd = int(rng.integers(6, 97))
        while H * Wd < 4 * N + 8:
            H += 4; Wd += 4
    F = int(rng.choice([3, 5, 9, 9, 9, 11, 15, 21, 31]))
    C = int(rng.choice([5, 6, 6]))
    dens_hi = float(rng.choice([0.2, 0.3, 0.35])) if os.environ.get("SOAK_CROWDED") else float(rng.choice([0.1, 0.2, 0.3]))
    ev = bool(rng.random() < 0.3)
    gs = bool(rng.random() < 0.5)
    W = int(rng.choice([1, 3, 17, 64, 200, 513]))
    if N >= 64 or F >= 21:
        W = min(W, 64)
    T = int(rng.choice([8, 24, 48]))
    seed = int(rng.integers(1 << 30))
    try:
        sc = random_scenario(W, H, Wd, N, density=(0.0, dens_hi), queue_len=1 if gs else 3, seed=seed, fov=F, num_channel=C,
                             use_da=ev, use_hp=ev and C == 6, unique_maps=min(W, 24))
    except Exception as ex:                                   # a shape the host-side generator cannot place
        continue
    off = 0
    if W > 8 and rng.random() < 0.3:                            # a shard of a larger job: worlds [lo, hi) with world_offset = lo
        lo = int(rng.integers(1, W // 2)); hi = int(rng.integers(lo + 1, W + 1))
        sc = sc.slice(lo, hi); off = lo; W = hi - lo
    env = BatchedMapfGym(sc, use_tape=False, seed=seed, goal_sampling=gs, world_offset=off)
    orc = OracleMapfGym(sc, seed=seed, threads=8, use_tape=False, goal_sampling=gs, world_offset=off)
    ckpt_at = int(rng.integers(0, T)) if rng.random() < 0.25 else -1     # save_state / load_state round trip in the middle
    acts = random_actions(T, W, N, seed=seed + 1)
    maps = env.bfs_maps() if rng.random() < 0.5 else None          # kept current with mapf_bfs_refresh after every step
    use_host = rng.random() < 0.3                                   # the split-phase host call (compact slab) instead of mode 0
    ring = env.make_host_ring(slots=2, action_slots=2, compact=True, numa_bind=False) if use_host else None
    obs_h = torch.empty((W, N, C, F, F), device="cuda") if use_host else None
    vec_h = torch.empty((W, N, 4), device="cuda") if use_host else None
    for t in range(T):
        if t == ckpt_at:                                            # checkpoint, scramble, restore: the rollout must continue unchanged
            blob = env.save_state().clone()
            env.step(torch.from_numpy(acts[(t + 1) % T]))
            env.load_state(blob)
        a = torch.from_numpy(acts[t])
        mode = (t + n_scen) % 3
        ref = orc.step(acts[t])
        if mode == 0 and use_host:
            from primal_ppo_b200 import decode_results
            ring["action_ring"][t & 1].copy_(a)
            env.step_observe_host_begin(ring["action_ring"][t & 1], ring["slots"][t & 1], obs_h, vec_h, train_valid_dev=env._out.train_valid)
            env.host_wait(0)
            dec = decode_results(ring["slots"][t & 1]["packed"])
            out = type("O", (), dict(status=dec["status"], reward=dec["reward"], cost=dec["cost"], train_valid=env._out.train_valid,
                                     goals_reached=dec["goals_reached"], violated=dec["violated"]))
            obs, vec = obs_h, vec_h
        elif mode == 0:
            out, obs, vec = env.step_observe(a)
        elif mode == 1:
            out = env.step(a); obs, vec = env.getAllObservations()
        else:
            ad = a.cuda()
            st = env.getActionStatus(ad); rw, sg = env.calculateActionReward(ad, st); rw = rw.clone()
            cost = env.calculateCostReward(ad).clone(); tv = env.getTrainValid(ad).clone(); st = st.clone()
            g, cv = env.jointStep(ad, st)
            rw[g == 1] += 1.5
            out = type("O", (), dict(status=st, reward=rw, cost=cost, train_valid=tv, goals_reached=g, violated=cv))
            obs, vec = env.getAllObservations()
        so, s = orc.state(), env.state()
        e_gpu = s["err"].cpu().numpy().astype(np.uint32)
        if not np.array_equal(e_gpu, so["err"]):
            print("MISMATCH err flags", dict(W=W, H=H, Wd=Wd, N=N, F=F, C=C, gs=gs, ev=ev, seed=seed, t=t, mode=mode)); sys.exit(1)
        ok = np.ones_like(so["err"], dtype=bool)     # flagged worlds (livelock cap, no viable action) are compared too: they stay valid
        flagged = so["err"] != 0
        for k in keys:
            x = getattr(out, k).cpu().numpy()[ok]
            if x.tobytes() != ref[k][ok].tobytes():
                print("MISMATCH", k, dict(W=W, H=H, Wd=Wd, N=N, F=F, C=C, gs=gs, ev=ev, seed=seed, t=t, mode=mode)); sys.exit(1)
        for k in ("pos", "goal", "rep"):
            if s[k].cpu().numpy()[ok].tobytes() != so[k][ok].tobytes():
                print("MISMATCH state", k, dict(W=W, H=H, Wd=Wd, N=N, F=F, C=C, gs=gs, ev=ev, seed=seed, t=t, mode=mode)); sys.exit(1)
        o_obs, o_vec = orc.getAllObservations()
        if obs.cpu().numpy()[ok].tobytes() != o_obs[ok].tobytes() or vec.cpu().numpy()[ok].tobytes() != o_vec[ok].tobytes():
            print("MISMATCH obs/vec", dict(W=W, H=H, Wd=Wd, N=N, F=F, C=C, gs=gs, ev=ev, seed=seed, t=t, mode=mode)); sys.exit(1)
        if maps is not None:
            env.refresh_bfs(maps, torch.from_numpy(ref["goals_reached"]).cuda())
        n_steps += 1
        n_agent_steps += int(ok.sum()) * N
    ob = orc.bfs_maps()
    if env.bfs_maps().cpu().numpy().tobytes() != ob.tobytes():
        print("MISMATCH bfs", dict(W=W, H=H, Wd=Wd, N=N, F=F, seed=seed)); sys.exit(1)
    if maps is not None and maps.cpu().numpy().tobytes() != ob.tobytes():
        print("MISMATCH bfs maps refreshed in place", dict(W=W, H=H, Wd=Wd, N=N, F=F, seed=seed, gs=gs)); sys.exit(1)
    n_scen += 1
    print(f"ok  W={W:4d} {H:3d}x{Wd:<3d} N={N:3d} F={F:2d} C={C} goal_sampling={int(gs)} eval={int(ev)} T={T} flagged={int(flagged.sum())}", flush=True)
    del env, orc
print(f"soak: {n_scen} scenarios, {n_steps} steps, {n_agent_steps} agent-steps compared bit for bit, 0 mismatches")
