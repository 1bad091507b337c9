"""CPU model of `bfs_gray_kernel` (primal_ppo_b200/csrc/bfs.cu), checked against the oracle's makeBfsMap restatement.

The kernel never writes a level per cell: the map is one bit string (bit i = cell r*Wd + c), levels are kept as
Gray-coded bit planes (one plane flips per level for the cells that are still unvisited), and at the end a bit-matrix
transpose per 32-bit word turns planes x cells into int16.  This file restates exactly those steps with NumPy uint32
words — the wavefront with masked +-1 shifts and +-Wd shifts, the first-touch / xor plane update, Gray -> binary,
the -1 / -2 patch in plane space, both transposes (8x8 on bytes with sign-extending widening for levels < 128,
16x16 on half-words otherwise) — so the algorithm is pinned on the CPU suite; the CUDA kernel itself is compared with
the same oracle in tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from oracle import OracleMapfGym
from primal_ppo_b200 import random_scenario

M32 = (1 << 32) - 1


def _wavefront_planes(free_bits, goal, H, Wd):
    """Level loop on Python big ints.  Returns (planes, unreached, n_levels)."""
    cells = H * Wd
    full = (1 << cells) - 1
    nc0 = sum(1 << i for i in range(cells) if i % Wd != 0)           # cells not in column 0
    ncl = sum(1 << i for i in range(cells) if i % Wd != Wd - 1)      # cells not in the last column
    g = 1 << goal
    fr, fm = g, free_bits & ~g
    planes = []
    level = 0
    while True:
        nw = (((fr << 1) & nc0) | ((fr >> 1) & ncl) | (fr << Wd) | (fr >> Wd)) & fm & full
        if nw == 0:
            break
        nl = level + 1
        b = (nl & -nl).bit_length() - 1                                # ctz(level + 1)
        if nl == 1 << b:
            assert len(planes) == b
            planes.append(fm)                                          # first touch: the plane was all zero
        else:
            planes[b] ^= fm                                            # every cell that was unvisited, the new frontier included
        fr, fm = nw, fm & ~nw
        level = nl
    return planes, fm, level


def _words(x, n):
    return np.array([(x >> (32 * j)) & M32 for j in range(n)], dtype=np.uint32)


def _transpose_stage(P, s, mask):
    for q in range(len(P)):
        if q & s == 0:
            t = ((P[q] >> np.uint32(s)) ^ P[q + s]) & np.uint32(mask)
            P[q + s] = P[q + s] ^ t
            P[q] = P[q] ^ (t << np.uint32(s))


def _convert(planes, free_with_goal, unreached, H, Wd, nb):
    """Plane words -> int16 map, word by word, the way a lane does it."""
    cells = H * Wd
    nwords = (cells + 31) // 32
    ob = ~_words(free_with_goal, nwords)
    un = _words(unreached, nwords)
    out = np.empty(nwords * 32, dtype=np.int16)
    if nb <= 7:
        Q = [(_words(planes[q], nwords) if q < nb else np.zeros(nwords, np.uint32)) for q in range(8)]
        for q in range(6, -1, -1):
            Q[q] = Q[q] ^ Q[q + 1]                                     # Gray -> binary
        Q[0] = (Q[0] | ob) & ~un                                       # -1 = 0xff, -2 = 0xfe
        for q in range(1, 8):
            Q[q] = Q[q] | ob | un
        _transpose_stage(Q, 4, 0x0F0F0F0F); _transpose_stage(Q, 2, 0x33333333); _transpose_stage(Q, 1, 0x55555555)
        for q in range(8):                                             # byte h of Q[q] = level of cell 8h + q
            for h in range(4):
                byte = ((Q[q] >> np.uint32(8 * h)) & np.uint32(0xFF)).astype(np.uint8)
                out[8 * h + q::32] = byte.view(np.int8).astype(np.int16)          # prmt with sign replication
    else:
        P = [(_words(planes[q], nwords) if q < nb else np.zeros(nwords, np.uint32)) for q in range(16)]
        for q in range(14, -1, -1):
            P[q] = P[q] ^ P[q + 1]
        P[0] = (P[0] | ob) & ~un
        for q in range(1, 16):
            P[q] = P[q] | ob | un
        _transpose_stage(P, 8, 0x00FF00FF); _transpose_stage(P, 4, 0x0F0F0F0F)
        _transpose_stage(P, 2, 0x33333333); _transpose_stage(P, 1, 0x55555555)
        for q in range(16):                                            # P[q] = level(cell q) | level(cell 16 + q) << 16
            out[q::32] = (P[q] & np.uint32(0xFFFF)).astype(np.uint16).view(np.int16)
            out[16 + q::32] = (P[q] >> np.uint32(16)).astype(np.uint16).view(np.int16)
    return out[:cells].reshape(H, Wd)


def gray_plane_bfs(obst, goal_rc):
    H, Wd = obst.shape
    flat = obst.reshape(-1)
    free_bits = sum(1 << i for i in range(H * Wd) if flat[i] == 0)
    goal = int(goal_rc[0]) * Wd + int(goal_rc[1])
    planes, unreached, level = _wavefront_planes(free_bits, goal, H, Wd)
    return _convert(planes, free_bits | (1 << goal), unreached, H, Wd, level.bit_length())


@pytest.mark.parametrize("shape,dens,seed", [((8, 5), 0.2, 1), ((12, 31), 0.3, 2), ((9, 32), 0.25, 3), ((20, 33), 0.3, 4),
                                             ((40, 40), 0.3, 5), ((16, 64), 0.35, 6), ((7, 9), 0.1, 7), ((10, 100), 0.3, 8)])
def test_gray_plane_model_matches_oracle(shape, dens, seed):
    H, Wd = shape
    sc = random_scenario(3, H, Wd, 4, density=(0.0, dens), queue_len=2, seed=seed, fov=3)
    orc = OracleMapfGym(sc, threads=2, use_tape=False)
    ref = orc.bfs_maps()
    goals = orc.state()["goal"]
    for w in range(sc.num_worlds):
        for i in range(sc.num_agents):
            got = gray_plane_bfs(sc.obst[w], goals[w, i])
            np.testing.assert_array_equal(got, ref[w, i], err_msg=f"world {w} agent {i} shape {shape}")


def test_gray_plane_model_deep_maze_uses_the_16_plane_path():
    """A serpentine corridor: distances of several hundred, i.e. more than 7 planes -> the 16x16 transpose path."""
    H = 24
    m = np.zeros((H, H), dtype=np.uint8)
    for r in range(1, H, 2):
        m[r, :] = 1
        m[r, H - 1 if (r // 2) % 2 == 0 else 0] = 0
    got = gray_plane_bfs(m, (0, 0))
    # plain BFS for the expectation
    exp = np.where(m != 0, -1, -2).astype(np.int16)
    exp[0, 0] = 0
    frontier, d = [(0, 0)], 0
    while frontier:
        d += 1
        nxt = []
        for r, c in frontier:
            for dr, dc in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                rr, cc = r + dr, c + dc
                if 0 <= rr < H and 0 <= cc < H and exp[rr, cc] == -2:
                    exp[rr, cc] = d
                    nxt.append((rr, cc))
        frontier = nxt
    assert exp.max() > 255
    np.testing.assert_array_equal(got, exp)


def test_gray_plane_model_goal_on_an_obstacle_and_unreachable_cells():
    """mapf_gym.py:216: the goal cell gets 0 even when it is not free; walled-off free cells stay -2."""
    m = np.zeros((8, 8), dtype=np.uint8)
    m[3, :] = 1                       # a wall splits the map
    m[0, 0] = 1                       # the goal sits on an obstacle
    got = gray_plane_bfs(m, (0, 0))
    assert got[0, 0] == 0 and got[0, 1] == 1 and got[2, 7] == 9
    assert (got[3, :] == -1).all() and (got[4:, :] == -2).all()
