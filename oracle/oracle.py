"""ctypes front-end of the C oracle (``oracle/mapf_oracle.c``) — TEST INFRASTRUCTURE ONLY.

``OracleMapfGym`` mirrors the reference's method names (``mapf_gym.py:327-637``) over a batch of W
independent worlds, all arrays on the host with a leading world dimension.  PARITY PINNED against the
reference-generated fixtures in ``tests/golden`` (see ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "mapf_oracle.c")
_BUILD = os.path.join(_HERE, "_build")
_LIB = os.path.join(_BUILD, "libmapf_oracle.so")

GOAL_REWARD = np.float32(1.5)    # alg_parameters.py:38; added by the rollout loop (runner.py:89-91)


def oracle_lib_path() -> str:
    return _LIB


def build_oracle(force: bool = False) -> str:
    """gcc -O2 -fopenmp -ffp-contract=off: no FMA contraction so that GAE rounds like NumPy."""
    os.makedirs(_BUILD, exist_ok=True)

    def stale():
        return force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC)
    if stale():
        import fcntl
        with open(os.path.join(_BUILD, ".lock"), "w") as lock:       # concurrent test / bench processes: one builds
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if stale():
                    cmd = ["gcc", "-O2", "-std=c11", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-fvisibility=hidden", "-o", _LIB + ".tmp", _SRC, "-lm"]
                    subprocess.run(cmd, check=True)
                    os.replace(_LIB + ".tmp", _LIB)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build_oracle()
        lib = C.CDLL(_LIB)
        p = C.c_void_p
        lib.orc_create.restype = p
        lib.orc_create.argtypes = [C.c_int] * 10 + [p] * 6 + [C.c_int, p, p, C.c_int, p, C.c_uint64, C.c_int, C.c_int]
        lib.orc_destroy.argtypes = [p]
        lib.orc_reset.argtypes = [p]
        lib.orc_get_action_status.argtypes = [p, p, p]
        lib.orc_calculate_action_reward.argtypes = [p, p, p, p, p]
        lib.orc_calculate_cost_reward.argtypes = [p, p, p]
        lib.orc_get_train_valid.argtypes = [p, p, p]
        lib.orc_joint_step.argtypes = [p, p, p, p, p, p]
        lib.orc_get_all_observations.argtypes = [p, p, p]
        lib.orc_bfs.argtypes = [p, p]
        lib.orc_step_observe.argtypes = [p] * 12
        lib.orc_state.argtypes = [p, p, p, p, p, p]
        lib.orc_gae.argtypes = [p, p, p, C.c_double, C.c_double, C.c_int, C.c_int, p, p]
        lib.orc_sample_actions.argtypes = [p, C.c_longlong, C.c_uint64, C.c_uint32, p, p]
        lib.orc_max_threads.restype = C.c_int
        lib.orc_set_goal_sampling.argtypes = [p, C.c_int]
        lib.orc_checksum_rows.argtypes = [p, C.c_longlong, C.c_longlong, p, C.c_int]
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleMapfGym:
    """Batched CPU oracle with the reference's method surface (leading world dim W)."""

    def __init__(self, scenario, seed: int = 1234, threads: int = 1, use_tape: bool = True, world_offset: int = 0,
                 goal_sampling: bool = False):
        lib = _load()
        sc = scenario
        sc.validate()
        self.W, self.H, self.Wd, self.N = sc.num_worlds, sc.height, sc.width, sc.num_agents
        self.F, self.Cn = sc.fov, sc.num_channel
        self._keep = [np.ascontiguousarray(x) if x is not None else None for x in
                      (sc.obst, sc.starts, sc.goal_queue, sc.htrace, sc.hlen, sc.hp5,
                       sc.tape if use_tape else None, sc.tape_len if use_tape else None)]
        TL = 0 if self._keep[6] is None else int(self._keep[6].shape[1])
        self._dims = None if sc.dims is None else np.ascontiguousarray(sc.dims)
        self._h = lib.orc_create(self.W, self.H, self.Wd, self.N, int(sc.goal_queue.shape[2]),
                                 int(sc.htrace.shape[1]), sc.fov, sc.num_channel, int(sc.use_da), int(sc.use_hp),
                                 *[_ptr(x) for x in self._keep[:6]],
                                 int(sc.hp5 is not None and sc.hp5.ndim == 4), _ptr(self._keep[6]), _ptr(self._keep[7]),
                                 TL, _ptr(self._dims), C.c_uint64(seed), int(threads), int(world_offset))
        self._lib = lib
        self.threads = int(threads)
        lib.orc_set_goal_sampling(self._h, int(bool(goal_sampling)))
        lib.orc_reset(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_destroy(self._h)
            self._h = None

    def reset(self):
        self._lib.orc_reset(self._h)

    def _acts(self, actions):
        a = np.ascontiguousarray(np.asarray(actions), dtype=np.int8)
        assert a.shape == (self.W, self.N), a.shape
        return a

    def getActionStatus(self, actions):
        a = self._acts(actions)
        st = np.zeros((self.W, self.N), dtype=np.int8)
        self._lib.orc_get_action_status(self._h, _ptr(a), _ptr(st))
        return st

    def calculateActionReward(self, actions, status):
        a = self._acts(actions)
        st = np.ascontiguousarray(status, dtype=np.int8)
        rw = np.zeros((self.W, self.N), dtype=np.float32)
        sg = np.zeros((self.W,), dtype=np.int32)
        self._lib.orc_calculate_action_reward(self._h, _ptr(a), _ptr(st), _ptr(rw), _ptr(sg))
        return rw, sg

    def calculateCostReward(self, actions):
        a = self._acts(actions)
        c = np.zeros((self.W, self.N), dtype=np.float32)
        self._lib.orc_calculate_cost_reward(self._h, _ptr(a), _ptr(c))
        return c

    def getTrainValid(self, actions):
        a = self._acts(actions)
        tv = np.zeros((self.W, self.N, 5), dtype=np.float32)
        self._lib.orc_get_train_valid(self._h, _ptr(a), _ptr(tv))
        return tv

    def jointStep(self, actions, status):
        a = self._acts(actions)
        st = np.ascontiguousarray(status, dtype=np.int8)
        g = np.zeros((self.W, self.N), dtype=np.uint8)
        v = np.zeros((self.W, self.N), dtype=np.uint8)
        self.fixed_actions = np.zeros((self.W, self.N), dtype=np.int8)
        self._lib.orc_joint_step(self._h, _ptr(a), _ptr(st), _ptr(g), _ptr(v), _ptr(self.fixed_actions))
        return g, v

    def getAllObservations(self, out=None):
        if out is None:
            obs = np.empty((self.W, self.N, self.Cn, self.F, self.F), dtype=np.float32)
            vec = np.empty((self.W, self.N, 4), dtype=np.float32)
        else:
            obs, vec = out
        self._lib.orc_get_all_observations(self._h, _ptr(obs), _ptr(vec))
        return obs, vec

    def bfs_maps(self):
        out = np.empty((self.W, self.N, self.H, self.Wd), dtype=np.int16)
        self._lib.orc_bfs(self._h, _ptr(out))
        return out

    def state(self):
        pos = np.empty((self.W, self.N, 2), dtype=np.int16)
        goal = np.empty((self.W, self.N, 2), dtype=np.int16)
        rep = np.empty((self.W, self.N), dtype=np.int8)
        err = np.empty((self.W,), dtype=np.uint32)
        good = np.empty((self.W, self.N), dtype=np.uint8)
        self._lib.orc_state(self._h, _ptr(pos), _ptr(goal), _ptr(rep), _ptr(err), _ptr(good))
        return dict(pos=pos, goal=goal, rep=rep, err=err, good=good)

    def step(self, actions):
        """The rollout loop's five calls in order plus the goal bonus (runner.py:64-91)."""
        st = self.getActionStatus(actions)
        rw, sg = self.calculateActionReward(actions, st)
        cost = self.calculateCostReward(actions)
        tv = self.getTrainValid(actions)
        g, v = self.jointStep(actions, st)
        rw[g == 1] += GOAL_REWARD
        return dict(status=st, reward=rw, cost=cost, train_valid=tv, goals_reached=g, violated=v, shadow=sg,
                    fixed=self.fixed_actions)


    def step_observe(self, actions, bufs=None):
        """One fused CPU call per env step (five calls + goal bonus + getAllObservations), one OpenMP region."""
        a = self._acts(actions)
        if bufs is None:
            W, N = self.W, self.N
            bufs = dict(status=np.empty((W, N), np.int8), reward=np.empty((W, N), np.float32),
                        cost=np.empty((W, N), np.float32), train_valid=np.empty((W, N, 5), np.float32),
                        goals_reached=np.empty((W, N), np.uint8), violated=np.empty((W, N), np.uint8),
                        shadow=np.empty((W,), np.int32), fixed=np.empty((W, N), np.int8),
                        obs=np.empty((W, N, self.Cn, self.F, self.F), np.float32), vec=np.empty((W, N, 4), np.float32))
        self._lib.orc_step_observe(self._h, _ptr(a), *[_ptr(bufs[k]) for k in
                                   ("status", "reward", "cost", "train_valid", "goals_reached", "violated", "shadow",
                                    "fixed", "obs", "vec")])
        return bufs


def gae_oracle(rewards, values, last_values, gamma=0.95, lam=0.95):
    """runner.py:120-149 for one stream. rewards, values: f32 [T, ...]; last_values: f32 [...]."""
    lib = _load()
    r = np.ascontiguousarray(rewards, dtype=np.float32)
    v = np.ascontiguousarray(values, dtype=np.float32)
    lv = np.ascontiguousarray(last_values, dtype=np.float32)
    T = r.shape[0]
    cols = int(np.prod(r.shape[1:])) if r.ndim > 1 else 1
    assert v.shape == r.shape and lv.size == cols
    ret = np.empty_like(r)
    adv = np.empty_like(r)
    lib.orc_gae(_ptr(r), _ptr(v), _ptr(lv), float(gamma), float(lam), T, cols, _ptr(ret), _ptr(adv))
    return ret, adv


def sample_actions_oracle(ps, seed=1234, draw=0):
    """model.py:38-40 with Philox draws (see orc_sample_actions). ps: f32 [..., 5] -> (int8 [...], f32 [...])."""
    lib = _load()
    p = np.ascontiguousarray(ps, dtype=np.float32)
    rows = p.size // 5
    a = np.empty(p.shape[:-1], dtype=np.int8)
    cp = np.empty(p.shape[:-1], dtype=np.float32)
    lib.orc_sample_actions(_ptr(p), rows, int(seed) & (2 ** 64 - 1), int(draw) & 0xffffffff, _ptr(a), _ptr(cp))
    return a, cp


def checksum_rows(a: np.ndarray, threads: int = 1) -> np.ndarray:
    """Row checksums with the definition of ``mapf_checksum_rows`` (include/mapf_b200.h): a [rows, ...] array of 4-byte
    items -> uint64 [rows]."""
    lib = _load()
    a = np.ascontiguousarray(a)
    rows = a.shape[0]
    words = (a.size // max(rows, 1)) * a.itemsize // 4 if rows else 0
    assert a.itemsize * (a.size // max(rows, 1)) % 4 == 0
    out = np.empty((rows,), dtype=np.uint64)
    if rows:
        lib.orc_checksum_rows(_ptr(a), rows, words, _ptr(out), int(threads))
    return out


def max_threads() -> int:
    return int(_load().orc_max_threads())
