"""The bench's timed regime for ncu: R independent C2 batches (inputs larger than L2) stepped round-robin WITHOUT a CUDA
graph; the cudaProfilerStart/Stop range covers one round (R launches, every batch once) after warm-up rounds, so that
`ncu --replay-mode range --cache-control none` replays the round as a whole and every launch finds its state in HBM,
exactly as in the timed region of bench.py.  Usage: python tools/prof_timed.py [R] [rounds-in-range]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import evacuation_b200 as eb

R = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
E = 4096
kw = dict(number_of_pedestrians=60, enslaving_degree=1.0, noise_coef=0.2, is_new_exiting_reward=True, is_new_followers_reward=True)
wrap = dict(positions="rel", statuses="ohe", type="Box")
dev = torch.device("cuda", 0)
acts = torch.rand((16, E, 2), device=dev) * 2 - 1
sets = []
for r in range(R):
    env = eb.setup_env(eb.EnvConfig(**kw), eb.EnvWrappersConfig(**wrap), num_envs=E, device=dev, seed=r, auto_reset=True)
    env.reset()
    sets.append(env)
for s in range(4):  # warm-up rounds
    for env in sets:
        env.unwrapped.step(acts[s])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for s in range(rounds):
    for env in sets:
        env.unwrapped.step(acts[4 + s])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"profiled {rounds} round(s) of {R} launches, {E} envs each")
