"""BASELINE config 5 on one rank: E envs x 60 pedestrians with the RPO transformer-embedding policy in the rollout
loop (policy forward -> ClipAction -> fused env step -> Normalize{Observation,Reward}), CUDA-graph replayed.
Usage: python tools/c5_rollout.py [E] [steps] [graph|eager] [fused|torch]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import evacuation_b200 as eb
from evacuation_b200.rollout import FusedRPOTransformerPolicy, PolicyRollout, RPOTransformerPolicy


def run(E, steps, use_graph=True, warmup=8, fused=True):
    env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=60, is_new_exiting_reward=True), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                       num_envs=E, seed=1, auto_reset=True)
    torch.manual_seed(1)
    policy = RPOTransformerPolicy(env.unwrapped.obs_dim if hasattr(env.unwrapped, "obs_dim") else 372, 60).cuda()
    if fused:
        policy = FusedRPOTransformerPolicy(policy, 60, device="cuda", seed=1)
    ro = PolicyRollout(env, policy, use_graph=use_graph, store=False)
    ro.reset()
    ro.run(warmup)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    ro.run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # env-only share of the same loop
    acts = torch.rand((E, 2), device="cuda") * 2 - 1
    e0.record()
    for _ in range(steps):
        env.step(acts)
    e1.record()
    torch.cuda.synchronize()
    ms_env = e0.elapsed_time(e1) / steps
    return dict(E=E, steps=steps, graph=use_graph, fused=fused, ms_per_step=ms, env_steps_per_s=E / ms * 1e3, ped_steps_per_s=E * 60 / ms * 1e3,
                env_only_ms_per_step=ms_env, policy_share=1 - ms_env / ms, finite=bool(torch.isfinite(ro.next_obs).all()),
                mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)


if __name__ == "__main__":
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    mode = sys.argv[3] if len(sys.argv) > 3 else "graph"
    impl = sys.argv[4] if len(sys.argv) > 4 else "fused"
    print(json.dumps(run(E, steps, use_graph=(mode == "graph"), fused=(impl == "fused"))))
