"""Experiment: the step kernel writing its observation block straight into page-locked HOST memory (UVA zero-copy) vs the
device block + one D2H copy of evac_step_host.  Prints per-step times (CUDA events and wall clock)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import evacuation_b200 as eb
from evacuation_b200 import _native as nat

E, K = 4096, 200
kw = dict(number_of_pedestrians=60, enslaving_degree=1.0, noise_coef=0.2, is_new_exiting_reward=True)
env = eb.setup_env(eb.EnvConfig(**kw), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"), num_envs=E, seed=0, auto_reset=True)
env.reset()
u = env.unwrapped
lib = nat.load()
D = u.obs_dim
acts = torch.rand((K, E, 2), device="cuda") * 2 - 1
obs_host = torch.empty((E, D), dtype=torch.float32, pin_memory=True)
obs_dev = torch.empty((E, D), dtype=torch.float32, device="cuda")
rew, term, trunc = torch.empty(E, device="cuda"), torch.empty(E, dtype=torch.uint8, device="cuda"), torch.empty(E, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
out = {"lib": os.path.basename(os.environ.get("EVAC_B200_LIB", "libevac_b200.so"))}
for label, obs in (("device_obs", obs_dev), ("zero_copy_host_obs", obs_host)):
    for s in range(5):
        nat.check(lib.evac_step(u._h, P(acts[s]), None, P(obs), P(rew), P(term), P(trunc), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(K):
        nat.check(lib.evac_step(u._h, P(acts[s]), None, P(obs), P(rew), P(term), P(trunc), st))
    e1.record()
    torch.cuda.synchronize()
    out[label + "_us"] = round(1e3 * e0.elapsed_time(e1) / K, 2)
    # per step with a host sync (what a gym-style caller sees)
    t0 = time.perf_counter()
    for s in range(K):
        nat.check(lib.evac_step(u._h, P(acts[s]), None, P(obs), P(rew), P(term), P(trunc), st))
        if obs is obs_dev:
            obs_host.copy_(obs_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    out[label + "_synced_wall_us"] = round(1e6 * (time.perf_counter() - t0) / K, 2)
chk = obs_host.clone()
nat.check(lib.evac_step(u._h, P(acts[0]), None, P(obs_dev), P(rew), P(term), P(trunc), st))
torch.cuda.synchronize()
out["finite"] = bool(torch.isfinite(chk).all())
print(json.dumps(out))
