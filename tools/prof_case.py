"""Run a few steps of one configuration (for ncu captures).  Usage: python tools/prof_case.py E N [steps] [rollout|step] [search] [warm-up rollout steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import evacuation_b200 as eb

E, n = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
mode = sys.argv[4] if len(sys.argv) > 4 else "step"
search = sys.argv[5] if len(sys.argv) > 5 else "auto"
warm = int(sys.argv[6]) if len(sys.argv) > 6 else 0
env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=n, is_new_exiting_reward=True), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                   num_envs=E, seed=0, auto_reset=True, neighbor_search=search)
env.reset()
env.rollout(8, agent="random")  # module load, attribute set-up
if warm:
    env.rollout(warm, agent="random")
acts = torch.rand((steps, E, 2), device="cuda") * 2 - 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
if mode == "rollout":
    env.rollout(steps, agent="random")
else:
    for s in range(steps):
        env.step(acts[s])
e1.record()
torch.cuda.synchronize()
print(f"E={E} N={n} {mode} x{steps} warm={warm} search={search} cells={env.unwrapped.num_cells}: {1e3 * e0.elapsed_time(e1) / steps:.1f} us/step")
