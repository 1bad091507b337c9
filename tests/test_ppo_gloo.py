"""Data-parallel learner on CPU (gloo, world_size 2): each rank holds half of a global minibatch; after
`PPOLearner.compute_gradients` (loss share -> backward -> ONE all-reduce of the flat gradient buffer) every rank must
hold the gradient a single process computes on the whole minibatch, incl. the globally normalised advantages
(SURVEY.md §8e), and identical parameters after the step."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

B, N = 8, 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(seed=0):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_ppo_golden import make_inputs
    d = make_inputs(np.random.default_rng(seed), B, N)
    t = lambda k: torch.from_numpy(d[k])
    return dict(obs=t("obs"), vec=t("vec"), returns=t("returns"), cost_returns=t("cost_returns"), values=t("old_v"),
                cost_values=t("old_cv"), actions=t("actions"), ps=t("old_ps"), train_valid=t("train_valid"))


def _learner(group):
    from primal_ppo_b200.ppo import PPOConfig, ScrimpPolicy
    from primal_ppo_b200.ppo.trainer import PPOLearner
    torch.manual_seed(7)
    pol = ScrimpPolicy().eval()
    return PPOLearner(pol, PPOConfig(cost_value_coef=0.05, cost_coef=0.2), group=group)


def _run_rank(rank, world_size, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    full = _batch()
    lo, hi = rank * B // world_size, (rank + 1) * B // world_size
    mine = {k: v[lo:hi].contiguous() for k, v in full.items()}
    lr = _learner(dist.group.WORLD)
    stats = lr.compute_gradients(mine)
    g_dp = lr.flat_grad.clone()
    st2 = lr.train_minibatch(mine, episode_cost=21.0, n_agents=N)
    flat_p = torch.cat([p.detach().flatten() for p in lr.params])
    # both ranks must agree exactly with each other
    other = [torch.zeros_like(flat_p) for _ in range(world_size)]
    dist.all_gather(other, flat_p)
    same_params = all(torch.equal(other[0], o) for o in other)
    if rank == 0:
        single = _learner(None)
        s1 = single.compute_gradients(full)
        g1 = single.flat_grad.clone()
        st1 = single.train_minibatch(full, episode_cost=21.0, n_agents=N)
        p1 = torch.cat([p.detach().flatten() for p in single.params])
        scale = float(g1.abs().max())
        ok = dict(grad=float((g_dp - g1).abs().max()) <= 2e-5 * scale,
                  stats=all(abs(stats[k] - s1[k]) <= 1e-5 + 1e-5 * abs(s1[k]) for k in s1),
                  norm=abs(st2["grad_norm"] - st1["grad_norm"]) <= 1e-4 * st1["grad_norm"],
                  # the first Adam step is lr*g/(|g|+1e-8): elements with |g| ~ 1e-8 turn fp32 summation-order noise into a
                  # different step, bounded by 2*lr; everything else must agree closely
                  params=float((flat_p - p1).abs().max()) <= 2.5e-5 and float((flat_p - p1).abs().mean()) <= 2e-7,
                  same=same_params)
        ret.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_rank, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    ok = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert all(ok.values()), ok
    assert all(p.exitcode == 0 for p in procs)


class _StubEnv:
    """Just enough of BatchedMapfGym for VecPPOTrainer.update(): shapes and a device (the rollout buffer is filled by hand)."""
    def __init__(self, W, N):
        self.W, self.N, self.C, self.F, self.device = W, N, 6, 9, torch.device("cpu")

    def getAllObservations(self, out=None):
        out[0].zero_(); out[1].zero_()
        return out


def _run_unequal(rank, world_size, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from primal_ppo_b200.ppo import PPOConfig, ScrimpPolicy
    from primal_ppo_b200.ppo.trainer import VecPPOTrainer
    from primal_ppo_b200.shard import shard_range
    lo, hi = shard_range(5, rank, world_size)              # 5 worlds over 2 ranks: 3 + 2
    W, T = hi - lo, 2
    torch.manual_seed(7)
    pol = ScrimpPolicy().eval()
    tr = VecPPOTrainer(_StubEnv(W, N), pol, PPOConfig(n_steps=T, n_epochs=2), group=dist.group.WORLD, rows_per_minibatch=2,
                       seed=3 + rank)
    g = torch.Generator().manual_seed(100 + rank)
    b = tr.buf
    b.obs.copy_((torch.rand(b.obs.shape, generator=g) < 0.2).float()); b.vec.copy_(torch.randn(b.vec.shape, generator=g))
    b.actions.copy_(torch.randint(0, 5, b.actions.shape, generator=g).to(torch.int8))
    b.ps.copy_(torch.softmax(torch.randn(b.ps.shape, generator=g), -1))
    for name in ("values", "cost_values", "rewards", "cost_rewards"):
        getattr(b, name).copy_(torch.randn(getattr(b, name).shape, generator=g))
    b.train_valid.copy_((torch.rand(b.train_valid.shape, generator=g) < 0.7).float())
    b.returns = torch.randn(b.values.shape, generator=g); b.cost_returns = torch.randn(b.values.shape, generator=g)
    # rank 0 has 6 (time, world) rows = 3 minibatches per epoch, rank 1 has 4 rows = 2: both must run 2 per epoch
    stats = tr.update(dict(episodeCostReward=1.0))
    flat_p = torch.cat([p.detach().flatten() for p in tr.learner.params])
    other = [torch.zeros_like(flat_p) for _ in range(world_size)]
    dist.all_gather(other, flat_p)
    if rank == 0:
        ret.put(dict(n=len(stats), same=all(torch.equal(other[0], o) for o in other)))
    dist.barrier()
    dist.destroy_process_group()


def test_unequal_shards_run_the_same_number_of_minibatches():
    """shard_range gives the first W % G ranks one world more; every minibatch issues collectives, so all ranks must loop the
    same number of times (the all-reduced minimum) instead of hanging on the first unmatched all_reduce."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_unequal, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    ok = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert ok == dict(n=4, same=True), ok
    assert all(p.exitcode == 0 for p in procs)
