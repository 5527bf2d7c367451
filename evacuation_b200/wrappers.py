"""Observation wrappers with the reference's names and constructor signatures
(src/env/wrappers/wrappers.py:8-96, src/env/wrappers/gravity_encoding.py:41-81).

In the reference each wrapper post-processes the observation in Python every step.  Here the
encodings are fused into the CUDA step kernel: a wrapper's constructor switches its encoding
on in the wrapped environment (`EvacuationEnv._configure_obs`) and `reset`/`step` only
re-shape the flat observation row the kernel wrote.  Stack them in the same order as
`EnvWrappersConfig.wrap_env` does; the outermost wrapper defines the returned structure."""
from __future__ import annotations

import warnings

import numpy as np

from .spaces import Box, Dict
from .statuses import Status


class ObservationWrapper:
    def __init__(self, env):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __getattr__(self, name):  # num_envs, rollout, get_state, ... are forwarded
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, seed=None, options=None):
        return self.unwrapped.reset(seed=seed, options=options)

    def step(self, action, noise=None):
        return self.unwrapped.step(action, noise=noise)

    def close(self):
        self.unwrapped.close()


class RelativePosition(ObservationWrapper):
    """wrappers.py:8-27: pedestrians and exit relative to the agent, scaled by sqrt(2)."""

    def __init__(self, env):
        super().__init__(env)
        self.unwrapped._configure_obs(positions="rel")


class PedestriansStatuses(ObservationWrapper):
    """wrappers.py:29-57: add `pedestrians_statuses` (one-hot [N,4] or categorical [N])."""

    def __init__(self, env, type: str = "ohe"):
        super().__init__(env)
        self.type = type
        n = env.unwrapped.pedestrians.num
        if self.type == "ohe":
            self.observation_space = Dict(self.observation_space)
            self.observation_space["pedestrians_statuses"] = Box(low=0, high=1, shape=(n, len(Status)), dtype=np.float32)
        elif self.type == "cat":
            self.observation_space = Dict(self.observation_space)
            self.observation_space["pedestrians_statuses"] = Box(low=0, high=1, shape=(n,), dtype=np.float32)
        elif self.type == "no":
            warnings.warn(f"No statuses will be added to observation as `type`='{self.type}'.")
        else:
            raise ValueError(f"Invalid value of `type`='{self.type}'. Must be 'ohe' or 'cat'.")
        self.unwrapped._configure_obs(statuses=self.type)


class MatrixObs(PedestriansStatuses):
    """wrappers.py:59-96: Box observation [N+2, 2|3|6], rows = agent, exit, pedestrians."""

    def __init__(self, env, type: str = "no"):
        if type not in ("no", "ohe", "cat"):
            raise ValueError(f"Invalid value of `type`='{type}'. Must be 'no', 'ohe' or 'cat'.")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            super().__init__(env, type=type)
        n = env.unwrapped.pedestrians.num
        cols = {"ohe": 2 + len(Status), "cat": 3, "no": 2}[self.type]
        self.observation_space = Box(low=-1, high=1, shape=(n + 2, cols), dtype=np.float32)
        self.unwrapped._configure_obs(obs_type="Box")


class GravityEncoding(ObservationWrapper):
    """gravity_encoding.py:41-81: agent position + gradients of the pedestrians' and the exit's
    gravity-like potentials."""

    def __init__(self, env, alpha: float, eps: float = None):
        super().__init__(env)
        self.alpha = alpha
        self.eps = self.env.unwrapped.area.eps if eps is None else eps
        if self.eps != self.env.unwrapped.area.eps:
            raise NotImplementedError("a GravityEncoding eps different from EnvConfig.eps is not supported")
        self.observation_space = Dict({
            "agent_position": self.observation_space["agent_position"],
            "grad_potential_pedestrians": Box(low=-1, high=1, shape=(2,), dtype=np.float32),
            "grad_potential_exit": Box(low=-1, high=1, shape=(2,), dtype=np.float32),
        })
        self.unwrapped._configure_obs(positions="grav", alpha=float(alpha))
