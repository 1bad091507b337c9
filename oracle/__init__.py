"""CPU oracle for the MAPF hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product package ``primal_ppo_b200`` never imports it.
"""
from .oracle import OracleMapfGym, build_oracle, gae_oracle, oracle_lib_path, sample_actions_oracle  # noqa: F401
