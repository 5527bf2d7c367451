"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/gen_golden.py

It imports the reference through oracle/ref_shim.py (stub gymnasium / matplotlib modules; no
reference code is changed or copied), drives `setup_env(...).reset()/step()` and stores
compact per-step records.  tests/test_oracle_golden.py replays every case with
oracle/evac_oracle.py and demands bit-identical results; the GPU parity tests then use the
oracle (which travels to the GPU box, while /root/reference does not).

Two RNG protocols:
  * "global":   np.random.seed(seed) before reset; the reference draws reset layout and the
                per-step noise from the global MT19937 stream (area.py:124, pedestrians.py:17-18).
  * "injected": reset as above, then np.random.uniform is patched for the duration of each
                step to return noise[t][fv_mask] from a dense, float32-representable
                noise[T,N] table -- the protocol the CUDA kernel's `noise` argument follows.
Actions and noise tables are regenerated from seeds by `case_inputs` (RandomState is frozen
by NumPy's compatibility policy), so only seeds are stored.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SNAP_EVERY = 100

BASE_ENV = dict(number_of_pedestrians=60)

CASES = [
    # name, env kwargs, wrapper kwargs, steps, seed, rng protocol, action kind
    dict(name="kat0_rel_ohe_box", env=dict(number_of_pedestrians=60, is_new_exiting_reward=True, intrinsic_reward_coef=1.0),
         wrap=dict(positions="rel", statuses="ohe", type="Box"), steps=100, seed=0, rng="global", actions="rotating"),
    dict(name="c1_default_seed0", env=dict(number_of_pedestrians=60), wrap=dict(), steps=2000, seed=0, rng="global", actions="random"),
    dict(name="c1_default_seed1", env=dict(number_of_pedestrians=60), wrap=dict(), steps=2000, seed=1, rng="global", actions="random"),
    dict(name="c2_rel_ohe_box_seed0", env=dict(number_of_pedestrians=60, enslaving_degree=1.0, noise_coef=0.2, is_new_exiting_reward=True),
         wrap=dict(positions="rel", statuses="ohe", type="Box"), steps=2000, seed=0, rng="injected", actions="random"),
    dict(name="c2_rel_ohe_box_seed3", env=dict(number_of_pedestrians=60, enslaving_degree=1.0, noise_coef=0.2, is_new_exiting_reward=True,
                                                 intrinsic_reward_coef=0.5),
         wrap=dict(positions="rel", statuses="ohe", type="Box"), steps=2000, seed=3, rng="injected", actions="sweep"),
    dict(name="c3_grav_a2", env=dict(number_of_pedestrians=60, enslaving_degree=0.5, noise_coef=0.5),
         wrap=dict(positions="grav", alpha=2), steps=500, seed=2, rng="injected", actions="random"),
    dict(name="c3_grav_a3", env=dict(number_of_pedestrians=60, enslaving_degree=0.5, noise_coef=0.5),
         wrap=dict(positions="grav", alpha=3), steps=2000, seed=4, rng="injected", actions="sweep"),
    dict(name="c3_grav_a4", env=dict(number_of_pedestrians=60, enslaving_degree=0.5, noise_coef=0.5),
         wrap=dict(positions="grav", alpha=4), steps=500, seed=5, rng="global", actions="random"),
    dict(name="c3_grav_a5", env=dict(number_of_pedestrians=60, enslaving_degree=0.5, noise_coef=0.5),
         wrap=dict(positions="grav", alpha=5), steps=500, seed=6, rng="injected", actions="rotating"),
    dict(name="n10_abs_cat_dict", env=dict(), wrap=dict(statuses="cat"), steps=600, seed=7, rng="global", actions="random"),
    dict(name="n10_abs_ohe_dict", env=dict(enslaving_degree=0.1, noise_coef=0.8), wrap=dict(statuses="ohe"), steps=600, seed=8,
         rng="injected", actions="sweep"),
    dict(name="n7_rel_cat_box", env=dict(number_of_pedestrians=7, noise_coef=0.05), wrap=dict(positions="rel", statuses="cat", type="Box"),
         steps=400, seed=9, rng="injected", actions="rotating"),
    dict(name="n1_abs_no_box", env=dict(number_of_pedestrians=1), wrap=dict(type="Box"), steps=300, seed=10, rng="global", actions="random"),
    dict(name="n33_rel_no_dict", env=dict(number_of_pedestrians=33, step_size=0.05), wrap=dict(positions="rel"), steps=400, seed=11,
         rng="injected", actions="sweep"),
    dict(name="n64_rel_no_box", env=dict(number_of_pedestrians=64, step_size=0.1, is_new_exiting_reward=True), wrap=dict(positions="rel", type="Box"),
         steps=300, seed=12, rng="injected", actions="random"),
    dict(name="n60_wallterm", env=dict(number_of_pedestrians=60, is_termination_agent_wall_collision=True, step_size=0.05),
         wrap=dict(positions="rel", statuses="ohe", type="Box"), steps=120, seed=13, rng="injected", actions="straight"),
    dict(name="n60_truncate50", env=dict(number_of_pedestrians=60, max_timesteps=50, init_reward_each_step=0.0, intrinsic_reward_coef=1.0),
         wrap=dict(statuses="ohe", type="Box"), steps=60, seed=14, rng="global", actions="random"),
    dict(name="n200_rel_ohe_box", env=dict(number_of_pedestrians=200, is_new_exiting_reward=True), wrap=dict(positions="rel", statuses="ohe", type="Box"),
         steps=300, seed=15, rng="injected", actions="sweep"),
    dict(name="n60_box_1p5x0p8", env=dict(number_of_pedestrians=60, width=1.5, height=0.8, step_size=0.02),
         wrap=dict(positions="rel", statuses="cat", type="Box"), steps=400, seed=16, rng="injected", actions="sweep"),
    dict(name="n3_escape_all", env=dict(number_of_pedestrians=3, step_size=0.05, noise_coef=0.05, is_new_exiting_reward=True),
         wrap=dict(statuses="ohe", type="Box"), steps=2000, seed=17, rng="injected", actions="herd"),
]


def case_inputs(case):
    """Deterministic actions [T,2] float32 and dense noise [T,N] (float32-representable float64)."""
    T = case["steps"]
    n = case["env"].get("number_of_pedestrians", 10)
    noise_coef = case["env"].get("noise_coef", 0.2)
    rs = np.random.RandomState(1000 + case["seed"])
    kind = case["actions"]
    t = np.arange(1, T + 1)
    if kind == "random":  # RandomAgent: action_space.sample() ~ U[-1,1]^2 float32 (random_agent.py:8-9)
        actions = rs.uniform(-1, 1, size=(T, 2)).astype(np.float32)
    elif kind == "rotating":  # RotatingAgent (rotating_agent.py:12-16)
        actions = np.stack([np.sin(0.05 * t), np.cos(0.05 * t)], axis=1).astype(np.float32)
    elif kind == "sweep":  # slow sweep across the arena, ending near the exit
        actions = np.stack([np.cos(0.011 * t) + 0.1 * rs.uniform(-1, 1, T), -0.35 + np.sin(0.023 * t)], axis=1).astype(np.float32)
    elif kind == "straight":  # run into the right wall
        actions = np.tile(np.array([[1.0, 0.05]], dtype=np.float32), (T, 1))
    elif kind == "herd":  # placeholder, overwritten by the closed-loop herding policy in run_case
        actions = np.zeros((T, 2), dtype=np.float32)
    else:
        raise ValueError(kind)
    noise = np.random.RandomState(5000 + case["seed"]).uniform(-noise_coef / 2, noise_coef / 2, size=(T, n))
    noise = noise.astype(np.float32).astype(np.float64)
    return actions, noise


def herd_action(unwrapped):
    """Closed-loop scripted leader for the all-escaped termination case: walk to the nearest
    non-exiting pedestrian, then drag it to the exit."""
    from oracle import evac_oracle as O  # only for the status constants

    u = unwrapped
    st = np.array([s.value for s in u.pedestrians.statuses])
    free = np.where(st == O.VISCEK)[0]
    fol = np.where(st == O.FOLLOWER)[0]
    a = u.agent.position.astype(np.float64)
    if len(fol) > 0 or len(free) == 0:
        target = np.array([0.0, -0.75])
    else:
        d = np.linalg.norm(u.pedestrians.positions[free] - a, axis=1)
        target = u.pedestrians.positions[free[np.argmin(d)]]
    v = target - a
    if np.linalg.norm(v) < 1e-3:
        v = np.array([0.0, -1.0])
    return (v / np.linalg.norm(v)).astype(np.float32)


def run_case(ref, case, record=True):
    warnings.filterwarnings("ignore")
    cfg = ref.EnvConfig(wandb_enabled=False, giff_freq=10 ** 9, path_logs=tempfile.mkdtemp(), **case["env"])
    wrap = ref.EnvWrappersConfig(**case["wrap"])
    env = ref.setup_env(cfg, wrap)
    u = env.unwrapped
    actions, noise = case_inputs(case)
    T, n = case["steps"], cfg.number_of_pedestrians
    np.random.seed(case["seed"])
    obs, _ = env.reset()

    from oracle.evac_oracle import flatten_observation

    def st_u8():
        return np.array([s.value for s in u.pedestrians.statuses], dtype=np.uint8)

    rec = dict(
        init_positions=u.pedestrians.positions.copy(), init_directions=u.pedestrians.directions.copy(), init_statuses=st_u8(),
        init_obs=flatten_observation(obs),
        statuses=np.zeros((T, n), np.uint8), rewards=np.zeros(T), terminated=np.zeros(T, bool), truncated=np.zeros(T, bool),
        agent_position=np.zeros((T, 2), np.float32), pos_sum=np.zeros(T), dir_sum=np.zeros(T), obs_sum=np.zeros(T),
        actions_used=np.zeros((T, 2), np.float32),
    )
    snaps = {}
    orig_uniform = np.random.uniform
    n_done = T
    for t in range(T):
        a = herd_action(u) if case["actions"] == "herd" else actions[t]
        rec["actions_used"][t] = a
        if case["rng"] == "injected":
            fv = np.isin(st_u8(), (1, 2))

            def fake_uniform(low=0.0, high=1.0, size=None, _fv=fv, _t=t):
                assert size == int(_fv.sum())
                return noise[_t][_fv].copy()

            np.random.uniform = fake_uniform
        try:
            obs, r, term, trunc, _ = env.step(a.copy())
        finally:
            np.random.uniform = orig_uniform
        rec["statuses"][t] = st_u8()
        rec["rewards"][t] = r
        rec["terminated"][t] = term
        rec["truncated"][t] = trunc
        rec["agent_position"][t] = u.agent.position
        with np.errstate(all="ignore"):
            rec["pos_sum"][t] = u.pedestrians.positions.sum()
            rec["dir_sum"][t] = u.pedestrians.directions.sum()
            fo = flatten_observation(obs)
            rec["obs_sum"][t] = fo.sum()
        if (t + 1) % SNAP_EVERY == 0 or t == T - 1 or term:
            snaps[t] = (u.pedestrians.positions.copy(), u.pedestrians.directions.copy(), fo.copy())
        if term:  # the reference env keeps stepping after termination; goldens stop here
            n_done = t + 1
            break
    for k in ("statuses", "rewards", "terminated", "truncated", "agent_position", "pos_sum", "dir_sum", "obs_sum", "actions_used"):
        rec[k] = rec[k][:n_done]
    ks = sorted(snaps)
    rec["snap_steps"] = np.array(ks, dtype=np.int64)
    rec["snap_positions"] = np.stack([snaps[k][0] for k in ks])
    rec["snap_directions"] = np.stack([snaps[k][1] for k in ks])
    rec["snap_obs"] = np.stack([snaps[k][2] for k in ks])
    rec["episode_reward"] = np.float64(u.episode_reward)
    rec["episode_intrinsic_reward"] = np.float64(u.episode_intrinsic_reward)
    rec["episode_status_reward"] = np.float64(u.episode_status_reward)
    rec["case_json"] = np.array(json.dumps(case))
    return rec


def main():
    from oracle import ref_shim

    ref = ref_shim.load_reference()
    for case in CASES:
        rec = run_case(ref, case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **rec)
        st = rec["statuses"][-1]
        print(f"{case['name']:28s} steps={len(rec['rewards']):5d} sum_r={rec['rewards'].sum():14.6f} "
              f"final(esc,exi,fol,vis)=({(st==4).sum()},{(st==3).sum()},{(st==2).sum()},{(st==1).sum()}) "
              f"term={bool(rec['terminated'].any())} size={os.path.getsize(path)//1024}KB")


if __name__ == "__main__":
    main()
