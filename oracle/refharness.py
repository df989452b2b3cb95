"""TEST / BASELINE INFRASTRUCTURE (never imported by the product package).

Run the UNMODIFIED reference env (SPaRC_Gym/SPaRC_Gym.py of tobiTKM/SPaRC-Gym) to pin the oracle, to generate the
golden vectors under tests/golden/, and as the CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind =
"reference"`).  The reference tree is looked up in this order: $SPARC_REFERENCE_ROOT, /root/reference (the build
container), baseline/_ref (a git-ignored install made by oracle/install_reference.py; it travels to the GPU box).

The reference needs `gymnasium` and `pygame` (not installed, no network) and downloads
its puzzles from the HF hub (SPaRC_Gym.py:77).  Recipe:

* inert stub modules for `gymnasium` (Env.reset seeds `self.np_random` with
  `np.random.default_rng(seed)` -- the same PCG64(SeedSequence(seed)) stream gymnasium
  uses), `gymnasium.spaces`, `gymnasium.envs.registration` and `pygame`;
* `datasets.load_dataset` monkey-patched to hand back our rows as a DataFrame.

On the GPU box `/root/reference` does not exist: only baseline/_ref can serve there, and only bench.py's CPU
arm uses it (the `-m gpu` tests and smoke() never read the reference).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALL_ROOT = os.path.join(REPO_ROOT, "baseline", "_ref")


def _has_engine(root):
    return bool(root) and os.path.isfile(os.path.join(root, "SPaRC_Gym", "SPaRC_Gym.py"))


def _find_root():
    for root in (os.environ.get("SPARC_REFERENCE_ROOT"), "/root/reference", INSTALL_ROOT):
        if _has_engine(root):
            return root
    return os.environ.get("SPARC_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "SPaRC_Gym", "SPaRC_Gym.py"))


def _install_stubs():
    if "gymnasium" in sys.modules and not getattr(sys.modules["gymnasium"], "_SPARC_STUB", False):
        return  # the real gymnasium is importable: use it
    gym = types.ModuleType("gymnasium")
    gym._SPARC_STUB = True

    class Env:
        np_random = None
        metadata = {}

        def reset(self, *, seed=None, options=None):
            if seed is not None or self.np_random is None:
                self.np_random = np.random.default_rng(seed)

        def close(self):
            pass

    class _Space:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k

    spaces = types.ModuleType("gymnasium.spaces")
    for name in ("Dict", "Box", "Text", "Discrete"):
        setattr(spaces, name, type(name, (_Space,), {}))
    envs = types.ModuleType("gymnasium.envs")
    registration = types.ModuleType("gymnasium.envs.registration")
    registry = {}

    def register(id, entry_point=None, **kw):
        registry[id] = entry_point

    registration.register = register
    registration.registry = registry
    envs.registration = registration
    gym.Env, gym.spaces, gym.envs = Env, spaces, envs
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.envs": envs,
                        "gymnasium.envs.registration": registration})
    if "pygame" not in sys.modules:
        sys.modules["pygame"] = types.ModuleType("pygame")


class _FakeDataset:
    def __init__(self, rows):
        self.rows = rows

    def to_pandas(self):
        import pandas as pd
        return pd.DataFrame(self.rows)


def load_reference_class():
    """Import the reference package and return its SPaRC_Gym class."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    try:
        import gymnasium  # noqa: F401  (real one, if it exists)
    except ImportError:
        _install_stubs()
    try:
        import pygame  # noqa: F401
    except ImportError:
        sys.modules["pygame"] = types.ModuleType("pygame")
    # Load the reference's engine module under a private package name, straight from its files: the
    # repo ships its own `SPaRC_Gym` import shim, which must not shadow (or be shadowed by) the reference.
    import importlib
    name = "_sparc_reference"
    if name + ".SPaRC_Gym" not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "SPaRC_Gym")]
        sys.modules[name] = pkg
        importlib.import_module(name + ".SPaRC_Gym")       # its `from .render import ...` resolves inside pkg.__path__
    return sys.modules[name + ".SPaRC_Gym"]


def make_reference_env(rows, **kwargs):
    """`SPaRC_Gym(**kwargs)` of the reference over `rows` (dataset-schema dicts)."""
    ref_mod = load_reference_class()
    ref_mod.load_dataset = lambda *a, **k: _FakeDataset(rows)
    return ref_mod.SPaRC_Gym(**kwargs)


def zero_leak(env, index=None):
    """Undo the reference's state leak across episodes (SPaRC_Gym.py:149-151,
    185-187): the cached `visited` / `agent_location` planes are never cleared."""
    idxs = range(len(env.puzzles)) if index is None else [index]
    for i in idxs:
        for k in ("visited", "agent_location"):
            env.puzzles[i]["obs_array"][k][...] = 0
