"""Loading of the reference-generated golden fixtures (tests/golden/*.npz; see make_golden.py)."""
import glob
import os

import numpy as np

from primal_ppo_b200.scenario import Scenario

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENV_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "g_*.npz")))


class Golden:
    def __init__(self, name):
        d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
        self.name = name
        self.scenario = Scenario.from_npz_dict(d)
        self.d = d
        shape = tuple(int(x) for x in d["obs_shape"])
        n = int(np.prod(shape))
        self.obs = np.unpackbits(d["obs_bits"])[:n].reshape(shape)      # u8 [T+1,W,N,C,F,F]
        self.T = int(d["actions"].shape[0])

    def __getitem__(self, k):
        return self.d[k]
