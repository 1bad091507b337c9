"""Import the UNMODIFIED reference (Nielsencu/primal-ppo) in the authoring container.

This module only works where ``/root/reference`` exists (the authoring container).  It is
used by ``make_golden.py`` to generate the committed golden fixtures and by a few optional
CPU tests that cross-check the oracle against the live reference.  Nothing here is imported
by the product path, by ``-m gpu`` tests, by ``smoke()`` or by ``bench.py``.

The reference imports ``skimage``, ``imageio`` and ``matplotlib`` at module scope
(``map_generator.py:5,7``, ``util.py:3,8``) and ``ray`` in ``runner.py:2``; none of them is used on
the env hot path, so empty stub modules are injected before importing.
"""
import os
import sys
import types

REFERENCE_PATH = os.environ.get("MAPF_REFERENCE_PATH", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_PATH, "mapf_gym.py"))


def load_reference(n_agents: int):
    """Returns the reference modules (mapf_gym, util, alg_parameters) with N_AGENTS patched.

    ``EnvParameters.N_AGENTS`` is a process-global read inside the env methods
    (``mapf_gym.py:328-331,371-372,437,439,485,...``), so it must be set before every
    construction of an env with a different agent count.
    """
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_PATH}")
    if REFERENCE_PATH not in sys.path:
        sys.path.insert(0, REFERENCE_PATH)
    for n in ["skimage", "skimage.measure", "skimage.morphology", "imageio",
              "matplotlib", "matplotlib.colors"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["skimage"].morphology = sys.modules["skimage.morphology"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    if not hasattr(sys.modules["matplotlib.colors"], "hsv_to_rgb"):
        sys.modules["matplotlib.colors"].hsv_to_rgb = lambda x: x
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")
        ray.remote = lambda *a, **k: (lambda c: c)
        sys.modules["ray"] = ray
    os.environ.setdefault("WANDB_MODE", "disabled")
    import alg_parameters as AP
    AP.EnvParameters.N_AGENTS = int(n_agents)
    import mapf_gym
    import util
    return mapf_gym, util, AP
