"""PCIe floor of the e2e leg: D2H of the 6.1 MB result block, H2D of the 32 KB action block (page-locked), and the pieces
of one evac_step_host call.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evacuation_b200 as eb


def timed(fn, n=200):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    dev = torch.device("cuda", 0)
    out = {}
    E, D = 4096, 372
    nbytes = E * D * 4 + E * 6
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    a_h = torch.empty(E * 8, dtype=torch.uint8, pin_memory=True)
    a_d = torch.empty(E * 8, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10
    out["d2h_6p1MB_us_device_timed"] = us
    out["d2h_GBps"] = nbytes / us / 1e3
    out["d2h_6p1MB_us_host_sync_each"] = timed(lambda: (h.copy_(d, non_blocking=True), torch.cuda.current_stream().synchronize()))
    out["h2d_32KB_us_host_sync_each"] = timed(lambda: (a_d.copy_(a_h, non_blocking=True), torch.cuda.current_stream().synchronize()))
    out["h2d_then_d2h_us"] = timed(lambda: (a_d.copy_(a_h, non_blocking=True), h.copy_(d, non_blocking=True), torch.cuda.current_stream().synchronize()))
    env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=60, is_new_exiting_reward=True), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                       num_envs=E, device=dev, seed=0, auto_reset=True, batched=False, rng="philox")
    env.reset()
    act = (np.random.rand(E, 2).astype(np.float32) * 2 - 1)
    for _ in range(5):
        env.step(act)
    out["env_step_host_us"] = timed(lambda: env.step(act), 300)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
