#!/usr/bin/env python
"""BASELINE.json configs[3]: full PPO rollout + update with the restated net.py policy, the GPU vector env, the GAE
kernel and the NCCL gradient all-reduce.  One process per GPU:

    python tools/bench_ppo.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_ppo.py

Prints one JSON line (rank 0): rollout agent-steps/s (policy forward + sampling + env step + observe), the env-only
share, update throughput (agent-rows/s through forward+backward+all-reduce+Adam), and the all-reduce time."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", type=int, default=2048, help="worlds per GPU")
    ap.add_argument("--agents", type=int, default=32)
    ap.add_argument("--size", type=int, default=40)
    ap.add_argument("--steps", type=int, default=16, help="rollout length T")
    ap.add_argument("--rows", type=int, default=256, help="(time, world) rows per minibatch per GPU")
    ap.add_argument("--minibatches", type=int, default=8)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--bf16-obs", action="store_true", help="rollout observations in the env's optional bf16 format")
    args = ap.parse_args()
    from primal_ppo_b200 import BatchedMapfGym, random_scenario
    from primal_ppo_b200.build import build
    from primal_ppo_b200.ppo import PPOConfig, ScrimpPolicy, VecPPOTrainer
    from primal_ppo_b200.shard import shard_range
    build()
    rank = int(os.environ.get("RANK", "0")); ws = int(os.environ.get("WORLD_SIZE", "1"))
    lr_ = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    group = None
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    W, N, T = args.worlds, args.agents, args.steps
    sc = random_scenario(W, args.size, args.size, N, density=(0.0, 0.3), queue_len=16, seed=500 + rank, unique_maps=min(W, 128))
    env = BatchedMapfGym(sc, device=dev, seed=1234, use_tape=False, world_offset=rank * W)
    torch.manual_seed(0)                                  # identical initial weights on every rank
    pol = ScrimpPolicy().to(dev).use_channels_last()
    cfg = PPOConfig(n_steps=T, n_epochs=1)
    amp = None if args.fp32 else torch.bfloat16
    tr = VecPPOTrainer(env, pol, cfg, group=group, amp_dtype=amp, rows_per_minibatch=args.rows, seed=1234 + rank,
                       obs_dtype=torch.bfloat16 if args.bf16_obs else torch.float32)

    def sync():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    perf = tr.collect()                                   # warm-up (cuDNN autotune, allocator)
    sync()
    t0 = time.perf_counter(); perf = tr.collect(); sync(); t_roll = time.perf_counter() - t0
    # env-only share of the rollout: replay the recorded actions through the fused env call
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for t in range(T):
        env.step_observe(tr.buf.actions[t], obs_out=(tr.buf.obs[t + 1], tr.buf.vec[t + 1]))
    b.record(); torch.cuda.synchronize(dev)
    env_ms = a.elapsed_time(b)
    tr.update(perf, max_minibatches=2)                    # warm-up
    sync()
    t0 = time.perf_counter(); stats = tr.update(perf, max_minibatches=args.minibatches); sync(); t_upd = time.perf_counter() - t0
    ar_ms = None
    if ws > 1:
        a.record()
        for _ in range(10):
            dist.all_reduce(tr.learner.flat_grad)
        b.record(); torch.cuda.synchronize(dev)
        ar_ms = a.elapsed_time(b) / 10
    tt = torch.tensor([t_roll, t_upd], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_roll, t_upd = float(tt[0]), float(tt[1])
    if rank == 0:
        nparam = tr.learner.flat_grad.numel()
        print(json.dumps({
            "workload": f"PPO rollout+update, {W} worlds/GPU {args.size}x{args.size}, {N} agents, T={T}, policy ScrimpPolicy "
                        f"({nparam} params, {'fp32' if args.fp32 else 'bf16 autocast'}), minibatch {args.rows} rows x {N} agents per GPU",
            "n_gpus": ws, "rollout_agent_steps_per_s": W * N * T * ws / t_roll, "rollout_ms_per_step": 1e3 * t_roll / T,
            "env_ms_per_step": env_ms / T, "env_share_of_rollout": (env_ms / T) / (1e3 * t_roll / T),
            "update_agent_rows_per_s": args.rows * N * len(stats) * ws / t_upd, "update_ms_per_minibatch": 1e3 * t_upd / len(stats),
            "grad_allreduce_ms": ar_ms, "grad_bytes": nparam * 4, "last_stats": stats[-1], "perf": perf}), flush=True)
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
