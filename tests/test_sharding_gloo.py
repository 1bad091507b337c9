"""N > 1 path on CPU: two gloo ranks each own a shard of the worlds (no env-path collective), results gathered on
rank 0 must equal the unsharded run bit for bit.  The env arithmetic is done by the oracle here (no GPU in this
container); the sharding / world_offset logic is the code under test and is what bench.py and the GPU env use."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import OracleMapfGym
from primal_ppo_b200 import random_actions, random_scenario
from primal_ppo_b200.shard import shard_range, shard_scenario

W, N, T = 96, 8, 24


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_rank(rank, world_size, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    sc = random_scenario(W, 8, 8, N, density=(0.2, 0.3), queue_len=6, seed=5)
    acts = random_actions(T, W, N, seed=6)
    mine, off = shard_scenario(sc, rank, world_size)
    lo, hi = shard_range(W, rank, world_size)
    env = OracleMapfGym(mine, seed=99, use_tape=False, world_offset=off)
    rewards = np.zeros((T, W, N), dtype=np.float32)
    pos = np.zeros((W, N, 2), dtype=np.int16)
    for t in range(T):
        out = env.step(acts[t, lo:hi])
        rewards[t, lo:hi] = out["reward"]
    pos[lo:hi] = env.state()["pos"]
    obs = np.zeros((W, N, 6, 9, 9), dtype=np.float32)
    obs[lo:hi] = env.getAllObservations()[0]
    # shards are disjoint, so a SUM reduce assembles the whole job (used only to check; the env path itself never communicates)
    tr, tp, to = torch.from_numpy(rewards), torch.from_numpy(pos.astype(np.int32)), torch.from_numpy(obs)
    for x in (tr, tp, to):
        dist.reduce(x, dst=0, op=dist.ReduceOp.SUM)
    # timing convention of bench.py: max over ranks
    tmax = torch.tensor([float(rank + 1)])
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        ref = OracleMapfGym(sc, seed=99, use_tape=False)
        rr = np.zeros_like(rewards)
        for t in range(T):
            rr[t] = ref.step(acts[t])["reward"]
        ok = (np.array_equal(tr.numpy().view(np.uint32), rr.view(np.uint32))
              and np.array_equal(tp.numpy(), ref.state()["pos"].astype(np.int32))
              and np.array_equal(to.numpy(), ref.getAllObservations()[0]) and float(tmax) == world_size)
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_worlds():
    for Wn in (1, 7, 64, 65536):
        for G in (1, 2, 3, 8):
            edges = [shard_range(Wn, r, G) for r in range(G)]
            assert edges[0][0] == 0 and edges[-1][1] == Wn
            assert all(edges[r][1] == edges[r + 1][0] for r in range(G - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shards_equal_unsharded():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_rank, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    ok = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)
