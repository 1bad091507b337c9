"""Scenario bundle: every exogenous input of a batch of MAPF worlds, as plain numpy arrays.

The reference env draws all of these from process-global RNGs at construction / on goal arrival
(``mapf_gym.py:164-190``, ``util.py:67-76``) or takes them through ``FixedMapfGym``
(``mapf_gym.py:648-669``).  The batched env takes them as arrays:

* ``obst``        u8  [W,H,Wd]     1 = obstacle (reference: -1), 0 = free
* ``starts``      i16 [W,N,2]      agent start cells (row, col)  — ``Sequence.items[0]``
* ``goal_queue``  i16 [W,N,Q,2]    goals in the order ``Sequence.getNext`` hands them out
                                   (``util.py:33-39``); when exhausted the last one repeats
* ``htrace``      i16 [W,L,4]      human (pos_r, pos_c, next_r, next_c) per tick; the tick advances
                                   once per ``jointStep`` (``mapf_gym.py:629``) and wraps at ``hlen[w]``
* ``hlen``        i32 [W]
* ``hp5``         i16 [W,5,2]      ``human.path[1:6]`` for the eval-only channel 5 (``mapf_gym.py:293-297``),
                                   rows padded with -1; or [W,L,5,2] = one entry per human tick (same index as
                                   ``htrace``) for walkers whose path changes (``FixedPathHuman``/``Human``)
* ``tape``        i8  [W,TL]       optional fixActions tape (``mapf_gym.py:587-598``): per random-branch
                                   event ``[chosen_action, n_evicted, evicted ids...]``
* ``tape_len``    i32 [W]
* ``dims``        i16 [W,2]        optional per-world (rows, cols) <= (H, Wd): the reference draws a different
                                   warehouse size per env (``map_generator.py:127-138``); cells outside a world's
                                   dims are out of bounds for that world and must be marked 1 in ``obst``

This file is host-side input plumbing only; no env arithmetic lives here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

DIRS = np.array([[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0]], dtype=np.int64)  # mapf_gym.py:97


@dataclass
class Scenario:
    obst: np.ndarray
    starts: np.ndarray
    goal_queue: np.ndarray
    htrace: np.ndarray
    hlen: np.ndarray
    hp5: Optional[np.ndarray] = None
    tape: Optional[np.ndarray] = None
    tape_len: Optional[np.ndarray] = None
    dims: Optional[np.ndarray] = None
    fov: int = 9
    num_channel: int = 6
    use_da: bool = False
    use_hp: bool = False
    meta: dict = field(default_factory=dict)

    @property
    def num_worlds(self) -> int:
        return int(self.obst.shape[0])

    @property
    def height(self) -> int:
        return int(self.obst.shape[1])

    @property
    def width(self) -> int:
        return int(self.obst.shape[2])

    @property
    def num_agents(self) -> int:
        return int(self.starts.shape[1])

    def validate(self) -> None:
        """Shape / dtype / range checks; raises ValueError (the kernels index shared memory with these values, so an
        out-of-range cell would be a silent out-of-bounds access, not an err[w] flag)."""
        def need(cond, msg):
            if not cond:
                raise ValueError("Scenario: " + msg)
        need(self.obst.ndim == 3 and self.obst.dtype == np.uint8, "obst must be uint8 [W,H,Wd]")
        W, H, Wd = self.obst.shape
        N = self.num_agents
        need(self.starts.shape == (W, N, 2) and self.starts.dtype == np.int16, "starts must be int16 [W,N,2]")
        need(self.goal_queue.ndim == 4 and self.goal_queue.shape[:2] == (W, N) and self.goal_queue.shape[3] == 2
             and self.goal_queue.dtype == np.int16 and self.goal_queue.shape[2] >= 1, "goal_queue must be int16 [W,N,Q>=1,2]")
        need(self.htrace.ndim == 3 and self.htrace.shape[0] == W and self.htrace.shape[2] == 4
             and self.htrace.dtype == np.int16, "htrace must be int16 [W,L,4]")
        need(self.hlen.shape == (W,) and self.hlen.dtype == np.int32, "hlen must be int32 [W]")
        need(bool(np.all(self.hlen >= 1) and np.all(self.hlen <= self.htrace.shape[1])), "1 <= hlen <= L")
        need(self.fov % 2 == 1 and self.fov >= 3, "fov must be odd and >= 3")
        need(self.num_channel in (5, 6), "num_channel must be 5 or 6")
        if self.hp5 is not None:
            need(self.hp5.dtype == np.int16 and (self.hp5.shape == (W, 5, 2) or self.hp5.shape == (W, self.htrace.shape[1], 5, 2)),
                 "hp5 must be int16 [W,5,2] or [W,L,5,2]")
        if self.tape is not None:
            need(self.tape.dtype == np.int8 and self.tape.shape[0] == W, "tape must be int8 [W,TL]")
            need(self.tape_len is not None and self.tape_len.shape == (W,), "tape needs tape_len [W]")
        rows = np.full((W,), H, dtype=np.int64)
        cols = np.full((W,), Wd, dtype=np.int64)
        if self.dims is not None:
            need(self.dims.shape == (W, 2) and self.dims.dtype == np.int16, "dims must be int16 [W,2]")
            need(bool(np.all(self.dims[:, 0] <= H) and np.all(self.dims[:, 1] <= Wd) and np.all(self.dims >= 1)), "1 <= dims <= (H, Wd)")
            rr = np.arange(H)[None, :, None] >= self.dims[:, 0][:, None, None]
            cc = np.arange(Wd)[None, None, :] >= self.dims[:, 1][:, None, None]
            need(bool(np.all(self.obst[rr | cc] == 1)), "cells outside dims must be obstacles")
            rows, cols = self.dims[:, 0].astype(np.int64), self.dims[:, 1].astype(np.int64)
        # starts and goals: inside the world's dims and on free cells
        widx = np.arange(W)[:, None]
        st = self.starts.astype(np.int64)
        need(bool(np.all(st >= 0) and np.all(st[..., 0] < rows[:, None]) and np.all(st[..., 1] < cols[:, None])),
             "starts outside the world")
        need(bool(np.all(self.obst[widx, st[..., 0], st[..., 1]] == 0)), "starts on obstacle cells")
        gq = self.goal_queue.astype(np.int64)
        need(bool(np.all(gq >= 0) and np.all(gq[..., 0] < rows[:, None, None]) and np.all(gq[..., 1] < cols[:, None, None])),
             "goals outside the world")
        need(bool(np.all(self.obst[widx[:, :, None], gq[..., 0], gq[..., 1]] == 0)), "goals on obstacle cells")
        ht = self.htrace.astype(np.int64)
        need(bool(np.all(ht >= -1) and np.all(ht[..., 0::2] < H) and np.all(ht[..., 1::2] < Wd)), "human trace outside the grid")

    def slice(self, lo: int, hi: int) -> "Scenario":
        """Worlds [lo, hi) — how ranks shard a job (no communication on the env path)."""
        def s(a):
            return None if a is None else np.ascontiguousarray(a[lo:hi])
        return Scenario(obst=s(self.obst), starts=s(self.starts), goal_queue=s(self.goal_queue),
                        htrace=s(self.htrace), hlen=s(self.hlen), hp5=s(self.hp5), tape=s(self.tape),
                        tape_len=s(self.tape_len), dims=s(self.dims), fov=self.fov, num_channel=self.num_channel,
                        use_da=self.use_da, use_hp=self.use_hp, meta=dict(self.meta))

    def to_npz_dict(self) -> dict:
        d = dict(obst=self.obst, starts=self.starts, goal_queue=self.goal_queue, htrace=self.htrace,
                 hlen=self.hlen, fov=np.int32(self.fov), num_channel=np.int32(self.num_channel),
                 use_da=np.bool_(self.use_da), use_hp=np.bool_(self.use_hp))
        if self.hp5 is not None:
            d["hp5"] = self.hp5
        if self.tape is not None:
            d["tape"] = self.tape
            d["tape_len"] = self.tape_len
        if self.dims is not None:
            d["dims"] = self.dims
        return d

    @staticmethod
    def from_npz_dict(d) -> "Scenario":
        return Scenario(obst=d["obst"], starts=d["starts"], goal_queue=d["goal_queue"],
                        htrace=d["htrace"], hlen=d["hlen"],
                        hp5=d["hp5"] if "hp5" in d else None,
                        tape=d["tape"] if "tape" in d else None,
                        tape_len=d["tape_len"] if "tape_len" in d else None,
                        dims=d["dims"] if "dims" in d else None,
                        fov=int(d["fov"]), num_channel=int(d["num_channel"]),
                        use_da=bool(d["use_da"]), use_hp=bool(d["use_hp"]))


def looping_trace(path) -> np.ndarray:
    """(pos, next) per tick of a ``LoopingHuman`` walking ``path`` (``mapf_gym.py:25-31,46-50``).

    tick s: pos = path[s]; next = path[s+1], or path[-1] on the last tick; after the last tick the
    walker restarts at s = 0 (``getNextGoal`` is a no-op for the looping human, ``mapf_gym.py:65-70``).
    """
    p = np.asarray(path, dtype=np.int16).reshape(-1, 2)
    L = p.shape[0]
    nxt = np.concatenate([p[1:], p[-1:]], axis=0)
    return np.concatenate([p, nxt], axis=1).reshape(L, 4)


def _bfs_dist(free: np.ndarray, src) -> np.ndarray:
    H, Wd = free.shape
    dist = np.full((H, Wd), -1, dtype=np.int32)
    dist[src] = 0
    frontier = [src]
    while frontier:
        nf = []
        for (r, c) in frontier:
            for dr, dc in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                rr, cc = r + dr, c + dc
                if 0 <= rr < H and 0 <= cc < Wd and free[rr, cc] and dist[rr, cc] < 0:
                    dist[rr, cc] = dist[r, c] + 1
                    nf.append((rr, cc))
        frontier = nf
    return dist


def _shortest_path(free: np.ndarray, src, dst):
    """A shortest 4-connected path src -> dst (inclusive). Tie-breaking is NOT astar_4's; the human is
    an exogenous input, so any valid walk is admissible for synthetic scenarios."""
    dist = _bfs_dist(free, dst)
    if dist[src] < 0:
        return None
    path = [src]
    r, c = src
    H, Wd = free.shape
    while (r, c) != dst:
        for dr, dc in ((0, 1), (1, 0), (0, -1), (-1, 0)):
            rr, cc = r + dr, c + dc
            if 0 <= rr < H and 0 <= cc < Wd and dist[rr, cc] == dist[r, c] - 1 and dist[rr, cc] >= 0:
                r, c = rr, cc
                break
        path.append((r, c))
    return path


def largest_component(free: np.ndarray) -> np.ndarray:
    """Boolean mask of the largest 4-connected free component."""
    H, Wd = free.shape
    seen = np.zeros_like(free, dtype=bool)
    best = None
    for r in range(H):
        for c in range(Wd):
            if free[r, c] and not seen[r, c]:
                d = _bfs_dist(free, (r, c)) >= 0
                seen |= d
                if best is None or d.sum() > best.sum():
                    best = d
    return best if best is not None else np.zeros_like(free, dtype=bool)


def random_scenario(num_worlds: int, height: int, width: int, num_agents: int, *, density=(0.0, 0.3),
                    queue_len: int = 16, seed: int = 0, fov: int = 9, num_channel: int = 6,
                    use_da: bool = False, use_hp: bool = False, max_human_len: int = 0,
                    unique_maps: int = 0) -> Scenario:
    """Synthetic density-map scenarios in the style of ``random_generator`` (``map_generator.py:23``):
    ``obst = rand(H, Wd) < p`` with ``p`` drawn per world from U[density]; distinct agent starts on free
    cells; goal queues of free cells (consecutive goals differ); a looping human whose start and goal lie
    in the largest free component.

    ``unique_maps`` > 0 builds only that many distinct worlds with the (slow, pure-Python) generator and
    tiles them to ``num_worlds`` with per-world re-drawn starts/goals — the benchmark's way of getting
    65 536 worlds without minutes of host-side BFS.
    """
    rng = np.random.default_rng(seed)
    W, H, Wd, N, Q = num_worlds, height, width, num_agents, queue_len
    U = W if unique_maps <= 0 else min(unique_maps, W)
    obst_u = np.zeros((U, H, Wd), dtype=np.uint8)
    traces = []
    hp5_u = np.full((U, 5, 2), -1, dtype=np.int16)
    for u in range(U):
        while True:
            p = rng.uniform(density[0], density[1])
            ob = rng.random((H, Wd)) < p
            free = ~ob
            comp = largest_component(free)
            cells = np.argwhere(comp)
            if len(cells) < max(2, 2):
                continue
            if free.sum() < N + 2:
                continue
            a, b = rng.choice(len(cells), size=2, replace=False)
            hs, hg = tuple(int(x) for x in cells[a]), tuple(int(x) for x in cells[b])
            out = _shortest_path(free, hs, hg)
            if out is None or len(out) < 2:
                continue
            loop = out + out[::-1][1:]          # start -> goal -> start (mapf_gym.py:33-37)
            if max_human_len and len(loop) > max_human_len:
                continue
            obst_u[u] = ob
            traces.append(looping_trace(loop))
            k = min(5, len(loop) - 1)
            hp5_u[u, :k] = np.asarray(loop[1:1 + k], dtype=np.int16)
            break
    L = max(t.shape[0] for t in traces)
    htrace_u = np.zeros((U, L, 4), dtype=np.int16)
    hlen_u = np.zeros((U,), dtype=np.int32)
    for u, t in enumerate(traces):
        htrace_u[u, :t.shape[0]] = t
        htrace_u[u, t.shape[0]:] = t[-1]
        hlen_u[u] = t.shape[0]

    idx = np.arange(W) % U
    obst = obst_u[idx]
    htrace = htrace_u[idx]
    hlen = hlen_u[idx]
    hp5 = hp5_u[idx]

    # starts and goal queues, vectorised over worlds: rank free cells by random keys.
    free = obst == 0                                        # [W,H,Wd]
    flat_free = free.reshape(W, H * Wd)
    hstart = htrace[:, 0, 0].astype(np.int64) * Wd + htrace[:, 0, 1].astype(np.int64)
    keys = rng.random((W, H * Wd))
    keys[~flat_free] = 2.0
    keys[np.arange(W), hstart] = 2.0                        # starts avoid the human's start (populateMap)
    order = np.argsort(keys, axis=1)[:, :N]                 # N distinct free cells per world
    starts = np.stack([order // Wd, order % Wd], axis=-1).astype(np.int16)

    goal_queue = np.zeros((W, N, Q, 2), dtype=np.int16)
    nfree = flat_free.sum(axis=1)
    free_sorted = np.argsort(~flat_free, axis=1, kind="stable")   # free cell ids first
    prev = order.copy()
    for q in range(Q):
        pick = (rng.random((W, N)) * nfree[:, None]).astype(np.int64)
        cell = np.take_along_axis(free_sorted, pick, axis=1)
        same = cell == prev
        if same.any():                                       # consecutive goals differ
            pick2 = (pick + 1) % nfree[:, None]
            cell = np.where(same, np.take_along_axis(free_sorted, pick2, axis=1), cell)
        goal_queue[:, :, q, 0] = cell // Wd
        goal_queue[:, :, q, 1] = cell % Wd
        prev = cell
    sc = Scenario(obst=np.ascontiguousarray(obst), starts=starts, goal_queue=goal_queue,
                  htrace=np.ascontiguousarray(htrace), hlen=np.ascontiguousarray(hlen),
                  hp5=np.ascontiguousarray(hp5), fov=fov, num_channel=num_channel,
                  use_da=use_da, use_hp=use_hp,
                  meta=dict(kind="random_density", density=list(density), seed=seed, unique_maps=U))
    sc.validate()
    return sc


def random_actions(num_steps: int, num_worlds: int, num_agents: int, seed: int = 1234) -> np.ndarray:
    """Uniform random joint actions in {0..4}, int8 [T,W,N] (SURVEY.md §8d synthetic inputs)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 5, size=(num_steps, num_worlds, num_agents), dtype=np.int8)
