"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.evac_oracle import OracleConfig, OracleEnv, compute_statuses  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN_DIR = os.path.join(HERE, "golden")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = json.loads(str(z["case_json"]))
    return case, z


def oracle_config_from_case(case) -> OracleConfig:
    kw = dict(case["env"])
    kw.update(case["wrap"])
    return OracleConfig(**kw)


def case_noise(case):
    """Dense noise table [T,N] of an 'injected' golden case (same construction as
    tests/golden/gen_golden.py::case_inputs)."""
    T = case["steps"]
    n = case["env"].get("number_of_pedestrians", 10)
    c = case["env"].get("noise_coef", 0.2)
    noise = np.random.RandomState(5000 + case["seed"]).uniform(-c / 2, c / 2, size=(T, n))
    return noise.astype(np.float32).astype(np.float64)


def make_oracle_at_golden_start(case, z) -> OracleEnv:
    """Oracle env whose state equals the reference's state right after reset()."""
    env = OracleEnv(oracle_config_from_case(case))
    env.set_state(z["init_positions"], z["init_directions"], z["init_statuses"], np.zeros(2, np.float32))
    env.n_episodes = 1
    return env


# ---------------------------------------------------------------------------------------------
# NumPy restatement of the library's counter-based random streams (csrc/philox.cuh) so the
# in-kernel RNG can be fed to the oracle bit for bit.
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
STREAM_NOISE, STREAM_RESET, STREAM_AGENT = 0, 1, 2


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0 & 0xFFFFFFFF), np.uint64(k1 & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_M0) * c[0]
        p1 = np.uint64(_M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(_W0)) & mask
        k1 = (k1 + np.uint64(_W1)) & mask
    return [x.astype(np.uint32) for x in c]


def evac_random(seed, stream, env, episode, now, ped):
    k0 = (seed & 0xFFFFFFFF) ^ ((stream * _W0) & 0xFFFFFFFF)
    k1 = (seed >> 32) & 0xFFFFFFFF
    return philox4x32_10(ped, now, episode, env, k0, k1)


def u01(r):
    return (np.asarray(r, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


_M2 = 0xD256D193


def philox2x32_10(c0, c1, k):
    """Vectorised Philox2x32-10 (csrc/philox.cuh::philox2x32_10); returns 2 uint32 arrays."""
    c0, c1, k = (np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for x in np.broadcast_arrays(c0, c1, k))
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p = np.uint64(_M2) * c0
        hi, lo = p >> np.uint64(32), p & mask
        c0, c1 = hi ^ k ^ c1, lo
        k = (k + np.uint64(_W0)) & mask
    return c0.astype(np.uint32), c1.astype(np.uint32)


def evac_key32(seed, stream, episode):
    return ((seed & 0xFFFFFFFF) ^ ((((seed >> 32) & 0xFFFFFFFF) * _W0) & 0xFFFFFFFF) ^ ((episode * _W1) & 0xFFFFFFFF)
            ^ ((stream * 0x85EBCA6B) & 0xFFFFFFFF))


def philox_noise(seed, env, episode, now, n, noise_coef):
    """Dense [n] float32 noise the kernels draw for (env, episode, step-in-episode `now`): pedestrian i reads word
    (i >> 5) & 1 of Philox2x32 block (i & 31) | (i >> 6) << 5, counter ((block & 2047) | now << 11, env), the upper block
    bits (pedestrians >= 4096) folded into the key."""
    i = np.arange(n, dtype=np.int64)
    block = (i & 31) | ((i >> 6) << 5)
    key = np.uint64(evac_key32(seed, STREAM_NOISE, episode)) ^ ((block >> 11).astype(np.uint64) * np.uint64(0xC2B2AE35) & np.uint64(0xFFFFFFFF))
    r = philox2x32_10((block & 2047) | (now << 11), env, key)
    words = np.where(((i >> 5) & 1) == 1, r[1], r[0])
    return (u01(words) - np.float32(0.5)) * np.float32(noise_coef)


def philox_action(seed, env, episode, now):
    r = philox2x32_10(now << 11, env, evac_key32(seed, STREAM_AGENT, episode))
    return np.array([np.float32(2) * u01(r[0]) - np.float32(1), np.float32(2) * u01(r[1]) - np.float32(1)], dtype=np.float32).reshape(2)


def philox_layout(seed, env, episode, n, dtype=np.float32):
    """(positions [n,2], directions [n,2]) of a kernel-side reset (csrc/evac_kernels.cuh::random_layout)."""
    r = evac_random(seed, STREAM_RESET, env, episode, 0, np.arange(n))
    two, one = np.float32(2), np.float32(1)
    px, py, vx, vy = (two * u01(r[i]) - one for i in range(4))
    pos = np.stack([px, py], axis=1).astype(dtype)
    v = np.stack([vx, vy], axis=1).astype(dtype)
    v[(v[:, 0] == 0) & (v[:, 1] == 0), 0] = 1
    nrm = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1])
    return pos, (v / nrm[:, None]).astype(dtype)
