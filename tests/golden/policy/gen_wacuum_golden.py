"""Golden action trace of the UNMODIFIED reference WacuumCleaner (src/agents/baseline_wacuum_cleaner.py) driving a
free agent point (pos += step_size * action / (|action| + eps), the agent_step of area.py:182-198 without walls).
The class is loaded as a bare module with stub `env` / parent packages (its real parents import gymnasium).
Run in the authoring container only:  python tests/golden/policy/gen_wacuum_golden.py"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/src"
agents = types.ModuleType("refagents")
agents.__path__ = [os.path.join(REF, "agents")]


class BaseAgent:
    def __init__(self, action_space):
        self.action_space = action_space


agents.BaseAgent = BaseAgent
sys.modules["refagents"] = agents
env_pkg = types.ModuleType("env")
env_pkg.EvacuationEnv = object
consts = types.ModuleType("env.constants")
spec = importlib.util.spec_from_file_location("env.constants", os.path.join(REF, "env", "constants.py"))
consts = importlib.util.module_from_spec(spec)
spec.loader.exec_module(consts)
sys.modules["env"], sys.modules["env.constants"] = env_pkg, consts
spec = importlib.util.spec_from_file_location("refagents.baseline_wacuum_cleaner", os.path.join(REF, "agents", "baseline_wacuum_cleaner.py"))
mod = importlib.util.module_from_spec(spec)
sys.modules[spec.name] = mod
spec.loader.exec_module(mod)


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


out = {}
for tag, (w, h, step) in {"unit": (1.0, 1.0, 0.01), "wide": (1.5, 0.8, 0.02)}.items():
    area = _NS(width=w, height=h, step_size=step, exit=_NS(position=np.array([0, -1], dtype=np.float32)))
    agent = mod.WacuumCleaner(_NS(area=area))
    pos = np.zeros(2, dtype=np.float32)
    P, A = [], []
    for t in range(4000):
        a = np.asarray(agent.act({"agent_position": pos}), dtype=np.float32)
        P.append(pos.copy()); A.append(a.copy())
        pos = (pos + np.float32(step) * a / (np.linalg.norm(a) + np.float32(1e-8))).astype(np.float32)
    out[tag + "_pos"], out[tag + "_act"] = np.array(P), np.array(A)
    out[tag + "_cfg"] = np.array([w, h, step])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "wacuum_actions.npz"), **out)
print({k: v.shape for k, v in out.items()})
