"""Pins the C oracle (oracle/mapf_oracle.c) bit-for-bit against traces of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py running /root/reference/mapf_gym.py).
Integer / byte outputs and f32 outputs are all compared with exact equality."""
import numpy as np
import pytest

from golden_util import ENV_CASES, Golden
from oracle import OracleMapfGym, gae_oracle


def _check_oracle_against(g, case):
    env = OracleMapfGym(g.scenario, threads=2)
    s0 = env.state()
    np.testing.assert_array_equal(s0["pos"], g["pos"][0])
    np.testing.assert_array_equal(s0["goal"], g["goal"][0])
    obs, vec = env.getAllObservations()
    np.testing.assert_array_equal(obs, g.obs[0].astype(np.float32))
    np.testing.assert_array_equal(vec.view(np.uint32), g["vec"][0].view(np.uint32))
    np.testing.assert_array_equal(env.bfs_maps()[:len(g["bfs0"])], g["bfs0"])
    for t in range(g.T):
        out = env.step(g["actions"][t])
        for key, ref in (("status", "status"), ("goals_reached", "goals_reached"), ("violated", "violated"),
                         ("shadow", "shadow"), ("fixed", "fixed")):
            np.testing.assert_array_equal(out[key], g[ref][t], err_msg=f"{case} t={t} {key}")
        for key in ("reward", "cost", "train_valid"):
            np.testing.assert_array_equal(out[key].view(np.uint32), g[key][t].view(np.uint32),
                                          err_msg=f"{case} t={t} {key}")
        s = env.state()
        assert not s["err"].any(), (case, t, s["err"])
        np.testing.assert_array_equal(s["pos"], g["pos"][t + 1], err_msg=f"{case} t={t} pos")
        np.testing.assert_array_equal(s["goal"], g["goal"][t + 1], err_msg=f"{case} t={t} goal")
        obs, vec = env.getAllObservations()
        np.testing.assert_array_equal(obs, g.obs[t + 1].astype(np.float32), err_msg=f"{case} t={t} obs")
        np.testing.assert_array_equal(vec.view(np.uint32), g["vec"][t + 1].view(np.uint32), err_msg=f"{case} t={t} vec")
    np.testing.assert_array_equal(env.bfs_maps()[:len(g["bfsT"])], g["bfsT"])


@pytest.mark.parametrize("case", ENV_CASES)
def test_oracle_matches_reference_trace(case):
    _check_oracle_against(Golden(case), case)


LIVE_CASES = {
    "live_10x10_n8": (dict(map="density", size=(10, 10), density=(0.15, 0.25), N=8, W=10, T=30, Q=8), 101),
    "live_8x8_n8_dense": (dict(map="density", size=(8, 8), density=(0.25, 0.3), N=8, W=16, T=24, Q=8, greedy=0.6), 102),
    "live_40x40_n32": (dict(map="density", size=(40, 40), density=(0.0, 0.3), N=32, W=2, T=12, Q=6), 103),
    "live_warehouse_n3_eval": (dict(map="warehouse", size=(10, 14), N=3, W=4, T=40, Q=8, human="fixed", use_da=True, use_hp=True), 104),
}


@pytest.mark.reference
@pytest.mark.parametrize("name", sorted(LIVE_CASES))
def test_oracle_matches_live_reference_on_fresh_scenarios(name, tmp_path):
    """Beyond the committed fixtures: run the UNMODIFIED reference right now (only where /root/reference exists) on
    scenarios drawn with seeds that are not in tests/golden, and hold the oracle to the same bit-exact standard."""
    import sys
    from golden_util import GOLDEN_DIR
    sys.path.insert(0, GOLDEN_DIR)
    from ref_loader import reference_available
    if not reference_available():
        pytest.skip("live reference not present (GPU box)")
    import make_golden
    case, seed = LIVE_CASES[name]
    make_golden.run_case(name, case, seed, out_dir=str(tmp_path))
    d = dict(np.load(tmp_path / (name + ".npz")))

    class _G(Golden):
        def __init__(self, d):
            from primal_ppo_b200.scenario import Scenario
            self.name, self.d = name, d
            self.scenario = Scenario.from_npz_dict(d)
            shape = tuple(int(x) for x in d["obs_shape"])
            self.obs = np.unpackbits(d["obs_bits"])[:int(np.prod(shape))].reshape(shape)
            self.T = int(d["actions"].shape[0])
    _check_oracle_against(_G(d), name)


def test_oracle_gae_matches_reference_runner():
    import os
    from golden_util import GOLDEN_DIR
    d = np.load(os.path.join(GOLDEN_DIR, "gae_runner.npz"))
    ret, _ = gae_oracle(d["rewards"], d["values"], d["last_values"], float(d["gamma"]), float(d["lam"]))
    np.testing.assert_array_equal(ret.view(np.uint32), d["returns"].view(np.uint32))
    cret, _ = gae_oracle(d["cost_rewards"], d["cost_values"], d["last_cost_values"], float(d["gamma"]), float(d["lam"]))
    np.testing.assert_array_equal(cret.view(np.uint32), d["cost_returns"].view(np.uint32))


def test_oracle_fused_step_observe_equals_separate_calls():
    g = Golden("g_8x8_n8_dense")
    e1, e2 = OracleMapfGym(g.scenario, threads=2), OracleMapfGym(g.scenario, threads=3)
    for t in range(g.T):
        a = e1.step(g["actions"][t])
        o1, v1 = e1.getAllObservations()
        b = e2.step_observe(g["actions"][t])
        for key in ("status", "reward", "cost", "train_valid", "goals_reached", "violated", "shadow", "fixed"):
            np.testing.assert_array_equal(a[key], b[key], err_msg=f"t={t} {key}")
        np.testing.assert_array_equal(o1, b["obs"])
        np.testing.assert_array_equal(v1.view(np.uint32), b["vec"].view(np.uint32))
