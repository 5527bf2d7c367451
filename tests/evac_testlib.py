"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.evac_oracle import OracleConfig, OracleEnv

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
