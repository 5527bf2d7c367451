"""Where one config-5 rollout iteration goes: kernel durations by name (CUPTI through torch.profiler; the loop runs eagerly so
every launch is visible; durations are the device's own, not ncu's cold-cache replays).
Usage: python tools/c5_breakdown.py [E] [iterations]"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import evacuation_b200 as eb
from evacuation_b200.rollout import FusedRPOTransformerPolicy, PolicyRollout, RPOTransformerPolicy

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=60, is_new_exiting_reward=True), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                   num_envs=E, seed=1, auto_reset=True)
torch.manual_seed(1)
policy = FusedRPOTransformerPolicy(RPOTransformerPolicy(372, 60).cuda(), 60, device="cuda", seed=1)
ro = PolicyRollout(env, policy, use_graph=False, store=False)
ro.reset()
ro.run(8)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ro.run(iters)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name[:90]]
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
rows = sorted(({"kernel": k, "launches_per_iteration": v[0] / iters, "us_per_iteration": v[1] / iters} for k, v in agg.items()), key=lambda r: -r["us_per_iteration"])
print(json.dumps({"E": E, "iterations": iters, "sum_us_per_iteration": sum(r["us_per_iteration"] for r in rows), "kernels": rows}, indent=1))
