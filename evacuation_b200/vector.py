"""`EvacuationVectorEnv` -- the drop-in for the vector environment the reference's trainer builds
(`src/agents/rpo_agent.py:121-124`):

    gym.vector.SyncVectorEnv([make_env(env_config, env_wrappers_config, gamma) for _ in range(num_envs)])

where `make_env` / `wrapping()` (`rpo_agent.py:24-39`) stack, per sub-environment,
FlattenObservation -> RecordEpisodeStatistics -> ClipAction -> NormalizeObservation -> clip(-1, 1) ->
NormalizeReward(gamma) -> clip(-100, 100).  Here the `num_envs` environments are ONE batched handle (one fused kernel
launch per `step`), and the wrapper chain runs on the device with one independent running estimate per environment,
exactly as the per-env wrappers of the reference keep them.

Surface (gymnasium 0.27-0.29 vector API, as consumed by `RPOAgent.__init__/learn`, rpo_agent.py:121-203):
`num_envs`, `single_observation_space`, `single_action_space`, `observation_space`, `action_space`,
`reset(seed=None, options=None) -> (obs, {})`, `step(actions) -> (obs, rewards, terminations, truncations, infos)` with
same-step auto-reset and `infos["final_info"]` / `infos["_final_info"]` carrying `{"episode": {"r", "l", "t"}}`
(RecordEpisodeStatistics: raw, un-normalised return and length) for the environments that finished, `close()`.

`output="numpy"` (default) returns host arrays like SyncVectorEnv; `output="torch"` keeps everything on the device.
Not provided: `infos["final_observation"]` (the terminal observation is overwritten by the same-step reset inside the
kernel; the reference's trainer never reads it) -- and, for the same reason, the observation normaliser is not updated
with terminal observations (once per episode and environment).
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

from .config import EnvConfig, EnvWrappersConfig
from .rollout import VectorNormalizer
from .spaces import Box


class EvacuationVectorEnv:
    def __init__(self, env_config: Optional[EnvConfig] = None, env_wrappers_config: Optional[EnvWrappersConfig] = None, num_envs: int = 1,
                 gamma: float = 0.99, device="cuda", seed: int = 0, output: str = "numpy", env_index_offset: int = 0):
        from . import setup_env  # late: evacuation_b200/__init__ imports this module's siblings

        if output not in ("numpy", "torch"):
            raise ValueError("output must be 'numpy' or 'torch'")
        self.output = output
        # batched=True also for num_envs == 1 (the class default): torch tensors on the device, Philox streams, auto-reset
        self.env = setup_env(env_config, env_wrappers_config, num_envs=num_envs, device=device, seed=seed, auto_reset=True,
                             env_index_offset=env_index_offset, batched=True)
        u = self.env.unwrapped
        self.num_envs, self.device, self.obs_dim = u.num_envs, u.device, u.obs_dim
        self.gamma = gamma
        # FlattenObservation of the wrapped space (Dict keys in sorted order == the kernel's row layout)
        self.single_observation_space = Box(low=-np.inf, high=np.inf, shape=(self.obs_dim,), dtype=np.float32)
        self.single_action_space = Box(low=-1.0, high=1.0, shape=(2,), dtype=np.float32)
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(self.num_envs, self.obs_dim), dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(self.num_envs, 2), dtype=np.float32)
        # float64 running statistics like gymnasium's RunningMeanStd (the float32 variant is the fused rollout path's)
        self.norm = VectorNormalizer(self.num_envs, self.obs_dim, gamma=gamma, device=self.device, dtype=torch.float64)
        self._t0 = time.perf_counter()
        self.closed = False

    # ------------------------------------------------------------------
    @property
    def unwrapped(self):
        return self.env.unwrapped

    def _out(self, t: torch.Tensor):
        return t if self.output == "torch" else t.cpu().numpy()

    def reset(self, seed=None, options=None):
        """Every sub-env's `reset()` through the wrapper chain: NormalizeObservation.reset also updates its estimate."""
        self.env.reset(seed=seed)
        return self._out(self.norm.observation(self.unwrapped.flat_observation).float()), {}

    def step(self, actions):
        act = torch.as_tensor(np.asarray(actions, dtype=np.float32) if not torch.is_tensor(actions) else actions)
        act = act.to(device=self.device, dtype=torch.float32).reshape(self.num_envs, 2).clamp(-1.0, 1.0).contiguous()  # ClipAction
        _, reward, term, trunc, _ = self.env.step(act)
        # FlattenObservation: the kernel's flat row IS the flattened observation for every wrapper configuration
        # (Dict keys in sorted order, Box rows row-major) -- the structured Dict / Box view is not re-flattened here
        obs_n = self.norm.observation(self.unwrapped.flat_observation)
        rew_n = self.norm.reward(reward, term)
        infos = {}
        done = term | trunc
        if bool(done.any()):  # one scalar D2H per step, like SyncVectorEnv's per-env Python check
            stats, finished, _ = self.unwrapped.episode_statistics()
            fin = finished.cpu().numpy()
            st = stats.cpu().numpy()
            final = np.full(self.num_envs, None, dtype=object)
            elapsed = time.perf_counter() - self._t0
            for e in np.nonzero(fin)[0]:  # RecordEpisodeStatistics: return of the RAW rewards, length, wall time
                final[e] = {"episode": {"r": np.array([st[e, 2]], dtype=np.float32), "l": np.array([int(st[e, 3])], dtype=np.int32),
                                        "t": np.array([elapsed], dtype=np.float32)},
                            "escaped_pedestrians": int(st[e, 4]), "exiting_pedestrians": int(st[e, 5]),
                            "following_pedestrians": int(st[e, 6]), "viscek_pedestrians": int(st[e, 7])}
            infos["final_info"] = final
            infos["_final_info"] = fin.astype(bool)
        return self._out(obs_n.float()), self._out(rew_n.float()), self._out(term), self._out(trunc), infos

    def close(self):
        if not self.closed:
            self.unwrapped.close()
            self.closed = True


__all__ = ["EvacuationVectorEnv"]
