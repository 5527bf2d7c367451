"""The oracle (oracle/evac_oracle.py) must be BIT-IDENTICAL to the unmodified reference on
every committed golden trajectory (tests/golden/*.npz, made by tests/golden/gen_golden.py).
CPU only."""
import numpy as np
import pytest

import evac_testlib  # noqa: F401  (puts the repo root on sys.path)
from oracle.evac_oracle import flatten_observation

from evac_testlib import OracleEnv, case_noise, golden_names, load_golden, oracle_config_from_case


def _eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_bit_identical_to_reference(name):
    case, z = load_golden(name)
    env = OracleEnv(oracle_config_from_case(case))
    np.random.seed(case["seed"])
    obs = env.reset()  # consumes the global stream like pedestrians.py:17-18
    assert _eq(env.positions, z["init_positions"])
    assert _eq(env.directions, z["init_directions"])
    assert _eq(env.statuses, z["init_statuses"])
    assert _eq(flatten_observation(obs), z["init_obs"])

    noise = case_noise(case) if case["rng"] == "injected" else None
    T = len(z["rewards"])
    snap = {int(s): i for i, s in enumerate(z["snap_steps"])}
    for t in range(T):
        obs, r, term, trunc, info = env.step(z["actions_used"][t].copy(), None if noise is None else noise[t])
        assert _eq(env.statuses, z["statuses"][t]), f"statuses differ at step {t}"
        assert r == z["rewards"][t] or (np.isnan(r) and np.isnan(z["rewards"][t])), f"reward differs at step {t}"
        assert bool(term) == bool(z["terminated"][t]) and bool(trunc) == bool(z["truncated"][t])
        assert _eq(env.agent_position, z["agent_position"][t])
        assert env.positions.sum() == z["pos_sum"][t], f"positions differ at step {t}"
        assert env.directions.sum() == z["dir_sum"][t], f"directions differ at step {t}"
        fo = flatten_observation(obs)
        assert fo.sum() == z["obs_sum"][t], f"observation differs at step {t}"
        if t in snap:
            i = snap[t]
            assert _eq(env.positions, z["snap_positions"][i])
            assert _eq(env.directions, z["snap_directions"][i])
            assert _eq(fo, z["snap_obs"][i])
    assert env.episode_reward == float(z["episode_reward"])
    assert env.episode_intrinsic_reward == float(z["episode_intrinsic_reward"])
    assert env.episode_status_reward == float(z["episode_status_reward"])


def test_kat0_values():
    """SURVEY.md 8(c) KAT-0, computed from the reference + numpy MT19937 + scipy."""
    case, z = load_golden("kat0_rel_ohe_box")
    assert z["init_positions"][0].tolist() == [0.0976270078546495, 0.43037873274483895]
    # (SURVEY.md prints the second component with 15 significant digits)
    assert np.allclose(z["init_directions"][0], [0.9999827162039409, 0.00587939566527009], rtol=1e-14, atol=0)
    assert z["init_positions"].sum() == -0.7459543057424016
    assert z["rewards"][:3].tolist() == [-2.1355140133627657, -2.1371264398539815, -2.138829876829163]
    assert z["pos_sum"][-1] == 8.936291463364315 and z["dir_sum"][-1] == 0.1050649092566393
    st = z["statuses"][-1]
    assert [(st == k).sum() for k in (4, 3, 2, 1)] == [6, 0, 5, 49]


@pytest.mark.parametrize("n,chunk", [(60, 7), (600, 64), (700, 1), (1500, 333)])
def test_row_chunked_alignment_is_bit_identical(n, chunk):
    """The > 8192-pedestrian parity cases evaluate area.py:105-119 on blocks of rows of the |fv| x |efv| distance matrix
    (32768^2 float64 entries do not fit in memory).  Every row's neighbour sums are reductions along the contiguous axis,
    so the blocked evaluation must give the same BITS as the one-expression evaluation the goldens pin."""
    from evac_testlib import OracleConfig

    cfg = OracleConfig(number_of_pedestrians=n, is_new_exiting_reward=True, intrinsic_reward_coef=0.5, enslaving_degree=0.6, noise_coef=0.4)
    envs = []
    for rc in (None, chunk):
        env = OracleEnv(cfg)
        env.row_chunk = rc
        np.random.seed(n)
        env.reset()
        env.positions[: n // 3] = np.array([0.4, 0.3]) + np.random.RandomState(1).normal(0, 0.03, (n // 3, 2))  # a dense blob
        envs.append(env)
    rs = np.random.RandomState(2)
    for t in range(4):
        a = rs.uniform(-1, 1, 2).astype(np.float32)
        nz = rs.uniform(-0.2, 0.2, n)
        outs = [e.step(a.copy(), nz.copy()) for e in envs]
        assert outs[0][1] == outs[1][1] and outs[0][4].margin == outs[1][4].margin
        assert _eq(envs[0].positions, envs[1].positions) and _eq(envs[0].directions, envs[1].directions) and _eq(envs[0].statuses, envs[1].statuses)
