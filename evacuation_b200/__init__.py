"""evacuation_b200 -- B200-native batched implementation of the cinemere/evacuation
environment step, behind the reference's own Python API (src/env/__init__.py:3-21)."""
from .agents import BaseAgent, RandomAgent, RotatingAgent, WacuumCleaner
from .config import EnvConfig, EnvWrappersConfig
from .env import EvacuationEnv
from .statuses import Status, SwitchDistances
from .wrappers import GravityEncoding, MatrixObs, PedestriansStatuses, RelativePosition

__all__ = [
    "setup_env", "EvacuationEnv", "Status", "SwitchDistances", "EnvConfig", "EnvWrappersConfig",
    "GravityEncoding", "PedestriansStatuses", "RelativePosition", "MatrixObs",
    "BaseAgent", "RandomAgent", "RotatingAgent", "WacuumCleaner", "EvacuationVectorEnv",
]


def __getattr__(name):  # lazy: vector.py pulls in torch-side glue (rollout.py)
    if name == "EvacuationVectorEnv":
        from .vector import EvacuationVectorEnv

        return EvacuationVectorEnv
    raise AttributeError(name)


def setup_env(env_config: EnvConfig = None, wrap_config: EnvWrappersConfig = None, **batch_kwargs):
    """src/env/__init__.py:18-21.  Extra keyword arguments (num_envs, device, seed, auto_reset,
    precision, rng, env_index_offset) select the batched face; without them the result behaves
    like the reference's single environment.  Classes instead of instances are accepted
    (README.md:72 passes `EnvConfig, EnvWrappersConfig`)."""
    env_config = EnvConfig() if env_config is None else (env_config() if isinstance(env_config, type) else env_config)
    wrap_config = EnvWrappersConfig() if wrap_config is None else (wrap_config() if isinstance(wrap_config, type) else wrap_config)
    env = EvacuationEnv(env_config, **batch_kwargs)
    env = wrap_config.wrap_env(env)
    return env
