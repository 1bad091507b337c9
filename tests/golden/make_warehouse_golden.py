#!/usr/bin/env python
"""tests/golden/warehouse_maps.npz: the reference's generateWarehouse (map_generator.py:127-138) for every length the
training env can draw (EnvParameters.WORLD_SIZE = (10, 40), alg_parameters.py:34), called with num_block=[L, L] so that
the only random draw is forced.  Used to check that the on-device generator lays out the same shelves."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

if __name__ == "__main__":
    from ref_loader import load_reference
    load_reference(2)
    import map_generator
    out = {}
    for L in range(4, 65):
        w = map_generator.generateWarehouse(num_block=[L, L])
        assert w.shape[0] == L
        out[f"L{L}"] = (w != 0).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "warehouse_maps.npz"), **out)
    print("warehouse_maps:", len(out), "maps; L=40 ->", out["L40"].shape, int(out["L40"].sum()), "shelf cells")
