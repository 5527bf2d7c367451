"""Import shim that lets the UNMODIFIED reference (cinemere/evacuation, /root/reference) run
in a container without gymnasium / matplotlib.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  This file is used by tests/golden/gen_golden.py (run in the
authoring container, where /root/reference is mounted) to produce the committed golden
vectors, and by bench.py's CPU arm (`--impl reference`, `cpu_baseline`) to time the unmodified
reference where it is reachable (/root/reference here; baseline/_ref on the GPU box).  Nothing in
the product path imports it.

It provides the tiny gymnasium surface the reference's env package touches
(SURVEY.md appendix A): gymnasium.Env, gymnasium.ObservationWrapper, spaces.Box / spaces.Dict,
and empty matplotlib modules (the reference imports them at module top; rendering is never hit).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference_root():
    """First of $EVAC_REFERENCE_ROOT, /root/reference (authoring container), <repo>/baseline/_ref (the pip --target install
    made by oracle/install_reference.py; it travels to the GPU box) that holds the reference's `src/env` package."""
    for cand in (os.environ.get("EVAC_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO_ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "src", "env", "env", "area.py")):
            return cand
    return None


REFERENCE_ROOT = find_reference_root() or "/root/reference"


def reference_available() -> bool:
    return find_reference_root() is not None


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class _Dict(dict):
    def __init__(self, spaces=None, **kw):
        super().__init__()
        if spaces is not None:
            self.update(spaces)
        self.update(kw)


class _Env:
    metadata: dict = {}

    def reset(self, seed=None, options=None):
        return None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class _ObservationWrapper:
    def __init__(self, env):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, obs):
        raise NotImplementedError


def install() -> None:
    """Insert stub modules into sys.modules and put the reference on sys.path."""
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box, spaces.Dict = _Box, _Dict
        gym.spaces = spaces
        gym.Env = _Env
        gym.ObservationWrapper = _ObservationWrapper
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    m = sys.modules["matplotlib"]
    for sub in ("pyplot", "patches", "animation"):
        setattr(m, sub, sys.modules[f"matplotlib.{sub}"])
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference():
    """Return the reference's `src.env` module (setup_env, EnvConfig, EnvWrappersConfig, ...)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install()
    import warnings

    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import importlib

    return importlib.import_module("src.env")
