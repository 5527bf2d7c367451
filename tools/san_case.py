"""Tiny workload for compute-sanitizer: create, reset, a few steps (per-step launches + one rollout launch), get_state.
Usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/san_case.py E N [steps] [wrap: rel|grav]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import evacuation_b200 as eb

E, n = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wrap = dict(positions="grav", alpha=3) if (len(sys.argv) > 4 and sys.argv[4] == "grav") else dict(positions="rel", statuses="ohe", type="Box")
env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=n, is_new_exiting_reward=True, max_timesteps=3), eb.EnvWrappersConfig(**wrap),
                   num_envs=E, seed=0, auto_reset=True, batched=True)
env.reset()
acts = torch.rand((steps, E, 2), device="cuda") * 2 - 1
for s in range(steps):
    env.step(acts[s])
env.rollout(steps, agent="random")
img = env.unwrapped.save_state()
env.unwrapped.load_state(img)
st = env.unwrapped.get_state()
torch.cuda.synchronize()
print(f"san_case E={E} N={n} steps={steps} cells={env.unwrapped.num_cells} ok, |pos|max={float(st['positions'].abs().max()):.3f}")
