"""Small GPU timing harness used while optimising (not the contract bench): rollout and per-step timings at a few batch sizes."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evacuation_b200 as eb

def run(E, K, wrap=dict(positions="rel", statuses="ohe", type="Box"), n=60, **envkw):
    env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=n, is_new_exiting_reward=True, **envkw), eb.EnvWrappersConfig(**wrap), num_envs=E, seed=0, auto_reset=True)
    env.reset()
    env.rollout(200, agent="random")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); env.rollout(K, agent="random"); e1.record(); torch.cuda.synchronize()
    t_roll = e0.elapsed_time(e1) / K
    acts = torch.rand((K, E, 2), device="cuda") * 2 - 1
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for k in range(3): env.step(acts[k])
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for k in range(K): env.step(acts[k])
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t_graph = e0.elapsed_time(e1) / K
    return dict(E=E, n=n, us_per_step_rollout=1e3 * t_roll, us_per_step_graph=1e3 * t_graph,
                gped_rollout=E * n / t_roll / 1e6, gped_graph=E * n / t_graph / 1e6,
                fp32_frac_rollout=E * (8 * n * n + 120 * n) / (t_roll * 1e-3) / 74.45e12)

if __name__ == "__main__":
    for E in (4096, 16384, 65536):
        print(json.dumps(run(E, 200)))
    print(json.dumps(run(4096, 200, wrap=dict(positions="grav", alpha=3), enslaving_degree=0.5, noise_coef=0.5)))
    print(json.dumps(run(256, 20, n=4096)))
