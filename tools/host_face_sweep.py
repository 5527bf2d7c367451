"""End-to-end step time of the host face (NumPy in / NumPy out, evac_step_host) over batch sizes, for the zero-copy threshold of
evac_step_host (EVAC_HOST_ZEROCOPY_BYTES).  Usage: python tools/host_face_sweep.py [E ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import evacuation_b200 as eb

out = {"zerocopy_bytes": os.environ.get("EVAC_HOST_ZEROCOPY_BYTES", "32768")}
for E in [int(v) for v in sys.argv[1:]] or [1, 8, 32, 128, 512, 2048]:
    env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=60, wandb_enabled=False), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                       num_envs=E, device="cuda:0", seed=1, auto_reset=True, batched=False, rng="philox")
    acts = np.random.RandomState(1).uniform(-1, 1, size=(64, E, 2)).astype(np.float32)
    env.reset()
    for t in range(30):
        env.step(acts[t % 64])
    n = 400
    t0 = time.perf_counter()
    for t in range(n):
        env.step(acts[t % 64])
    out[f"E{E}_us"] = round(1e6 * (time.perf_counter() - t0) / n, 1)
    env.unwrapped.close()
print(json.dumps(out))
