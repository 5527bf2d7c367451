"""EnvConfig / EnvWrappersConfig -- same field names and defaults as the reference
(src/env/env/config.py:3-100, src/env/wrappers/config.py:8-93)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional


@dataclass
class EnvConfig:
    experiment_name: str = "test"
    # ---- geometry
    number_of_pedestrians: int = 10
    width: float = 1.0
    height: float = 1.0
    step_size: float = 0.01
    noise_coef: float = 0.2
    eps: float = 1e-8
    # ---- leader
    enslaving_degree: float = 1.0
    # ---- reward
    is_new_exiting_reward: bool = False
    is_new_followers_reward: bool = True
    intrinsic_reward_coef: float = 0.0
    is_termination_agent_wall_collision: bool = False
    init_reward_each_step: float = -1.0
    # ---- timing
    max_timesteps: int = 2_000
    n_episodes: int = 0
    n_timesteps: int = 0
    # ---- logging / drawing (env.py:23-31,110-127,153-165): host-side, for the tracked environment (env.unwrapped.tracked_env)
    render_mode: Optional[str] = None
    draw: bool = False
    verbose: bool = False
    giff_freq: int = 500
    wandb_enabled: bool = True
    path_giff: str = "saved_data/giff"
    path_png: str = "saved_data/png"
    path_logs: str = "saved_data/logs"

    def __post_init__(self):
        # config.py:97-100
        assert self.n_episodes == 0, NotImplementedError
        assert self.n_timesteps == 0, NotImplementedError


@dataclass
class EnvWrappersConfig:
    """Observation wrappers params (wrappers/config.py:8-44)."""

    num_obs_stacks: int = 1
    positions: Literal["abs", "rel", "grav"] = "abs"
    statuses: Literal["no", "ohe", "cat"] = "no"
    type: Literal["Dict", "Box"] = "Dict"
    alpha: float = 3

    def __post_init__(self):
        assert self.num_obs_stacks == 1, NotImplementedError  # wrappers/config.py:42-44

    def wrap_env(self, env):
        """Same dispatch table as wrappers/config.py:76-93.  The wrappers do not run Python
        per step here: each one switches on the corresponding encoding inside the fused CUDA
        step kernel and only reshapes the flat observation row it gets back."""
        from .wrappers import GravityEncoding, MatrixObs, PedestriansStatuses, RelativePosition

        if self.positions == "grav":
            if self.type == "Dict":
                return GravityEncoding(env, alpha=self.alpha)
            elif self.type == "Box":
                raise NotImplementedError
            else:
                raise ValueError
        if self.positions == "rel":
            env = RelativePosition(env)
        if self.type == "Box":
            return MatrixObs(env, type=self.statuses)
        if self.statuses != "no":
            env = PedestriansStatuses(env, type=self.statuses)
        return env
