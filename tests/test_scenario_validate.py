"""Scenario.validate: the kernels index shared memory with starts / goals / human cells, so out-of-range or blocked cells
must be refused on the host with a ValueError (not an assert that `python -O` removes, not a silent out-of-bounds access)."""
import numpy as np
import pytest

from primal_ppo_b200 import random_scenario
from primal_ppo_b200.scenario import Scenario


def _copy(sc, **kw):
    d = dict(obst=sc.obst.copy(), starts=sc.starts.copy(), goal_queue=sc.goal_queue.copy(), htrace=sc.htrace.copy(),
             hlen=sc.hlen.copy(), hp5=None if sc.hp5 is None else sc.hp5.copy(), dims=None if sc.dims is None else sc.dims.copy(),
             fov=sc.fov, num_channel=sc.num_channel)
    d.update(kw)
    return Scenario(**d)


def test_valid_scenarios_pass():
    random_scenario(5, 12, 9, 4, seed=1).validate()
    sc = random_scenario(3, 10, 10, 2, seed=2)
    dims = np.array([[10, 10], [8, 10], [10, 7]], dtype=np.int16)
    ob = sc.obst.copy()
    ob[1, 8:, :] = 1; ob[2, :, 7:] = 1
    st = sc.starts.copy(); gq = sc.goal_queue.copy(); ht = sc.htrace.copy()
    # move everything of worlds 1, 2 into their (smaller) dims on free cells
    for w in (1, 2):
        free = np.argwhere(ob[w] == 0)
        st[w] = free[:2]; gq[w] = free[2:3][None].repeat(2, 0).reshape(2, 1, 2).repeat(gq.shape[2], 1)
        ht[w, :, :2] = free[4]; ht[w, :, 2:] = free[4]
    _copy(sc, obst=ob, starts=st.astype(np.int16), goal_queue=gq.astype(np.int16), htrace=ht.astype(np.int16), dims=dims).validate()


@pytest.mark.parametrize("what", ["start_oob", "start_negative", "start_on_obstacle", "goal_oob", "goal_on_obstacle", "dtype", "hlen",
                                  "dims_not_blocked", "start_outside_dims", "fov_even", "human_oob"])
def test_bad_scenarios_raise_value_error(what):
    sc = random_scenario(3, 10, 10, 2, density=(0.2, 0.2), seed=3)
    ob_cell = tuple(np.argwhere(sc.obst[0] == 1)[0])
    kw = {}
    if what == "start_oob":
        s = sc.starts.copy(); s[0, 0] = (10, 3); kw["starts"] = s
    elif what == "start_negative":
        s = sc.starts.copy(); s[1, 1] = (-1, 0); kw["starts"] = s
    elif what == "start_on_obstacle":
        s = sc.starts.copy(); s[0, 1] = ob_cell; kw["starts"] = s
    elif what == "goal_oob":
        g = sc.goal_queue.copy(); g[2, 0, -1] = (3, 10); kw["goal_queue"] = g
    elif what == "goal_on_obstacle":
        g = sc.goal_queue.copy(); g[0, 1, 0] = ob_cell; kw["goal_queue"] = g
    elif what == "dtype":
        kw["starts"] = sc.starts.astype(np.int32)
    elif what == "hlen":
        kw["hlen"] = np.zeros_like(sc.hlen)
    elif what == "dims_not_blocked":
        kw["dims"] = np.array([[10, 10], [9, 10], [10, 10]], dtype=np.int16)          # row 9 of world 1 is not all obstacles
    elif what == "start_outside_dims":
        ob = sc.obst.copy(); ob[1, 9, :] = 1
        s = sc.starts.copy(); s[1, 0] = (9, 0)
        kw.update(obst=ob, starts=s, dims=np.array([[10, 10], [9, 10], [10, 10]], dtype=np.int16))
    elif what == "fov_even":
        kw["fov"] = 8
    elif what == "human_oob":
        h = sc.htrace.copy(); h[0, 0, 0] = 10; kw["htrace"] = h
    with pytest.raises(ValueError):
        _copy(sc, **kw).validate()
