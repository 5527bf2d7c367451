"""GPU tests of `EvacuationVectorEnv`, the drop-in for the reference trainer's
`gym.vector.SyncVectorEnv([make_env(...)] * num_envs)` (src/agents/rpo_agent.py:24-39,121-124): the wrapper chain
(FlattenObservation, RecordEpisodeStatistics, ClipAction, NormalizeObservation + clip, NormalizeReward + clip) against a
float64 restatement of gymnasium's per-env RunningMeanStd bookkeeping driven by the raw batched env."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _RMS:  # gymnasium.wrappers.normalize.RunningMeanStd, batch of one sample
    def __init__(self, shape):
        self.mean, self.var, self.count = np.zeros(shape), np.ones(shape), 1e-4

    def update(self, x):
        delta = x - self.mean
        tot = self.count + 1
        self.mean = self.mean + delta / tot
        self.var = (self.var * self.count + delta ** 2 * self.count / tot) / tot
        self.count = tot


def _cfgs(**kw):
    import evacuation_b200 as eb
    return (eb.EnvConfig(number_of_pedestrians=20, is_new_exiting_reward=True, **kw),
            eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"))


def test_vector_env_surface_matches_sync_vector_env():
    import evacuation_b200 as eb
    E = 6
    envs = eb.EvacuationVectorEnv(*_cfgs(), num_envs=E, gamma=0.99, seed=3)
    assert envs.num_envs == E and envs.single_observation_space.shape == (22 * 6,) and envs.single_action_space.shape == (2,)
    assert envs.observation_space.shape == (E, 132) and envs.action_space.shape == (E, 2)
    obs, info = envs.reset(seed=1)
    assert isinstance(obs, np.ndarray) and obs.shape == (E, 132) and obs.dtype == np.float32 and info == {}
    out = envs.step(np.stack([envs.single_action_space.sample() for _ in range(E)]) * 3.0)  # out-of-range actions are clipped
    obs, rew, term, trunc, infos = out
    assert obs.shape == (E, 132) and rew.shape == (E,) and term.shape == (E,) and trunc.shape == (E,)
    assert term.dtype == bool and trunc.dtype == bool and np.abs(obs).max() <= 1.0 and np.abs(rew).max() <= 100.0
    envs.close()
    dev = eb.EvacuationVectorEnv(*_cfgs(), num_envs=E, output="torch")
    o, _ = dev.reset()
    assert torch.is_tensor(o) and o.is_cuda


def test_vector_env_wrapper_chain_against_gymnasium_bookkeeping():
    import evacuation_b200 as eb
    E, T, gamma, max_t = 5, 130, 0.97, 40
    cfgs = _cfgs(max_timesteps=max_t)
    envs = eb.EvacuationVectorEnv(*cfgs, num_envs=E, gamma=gamma, seed=9)
    raw = eb.setup_env(*cfgs, num_envs=E, seed=9, auto_reset=True)   # the same batch without the chain
    D = envs.obs_dim
    obs_rms, ret_rms = [_RMS((D,)) for _ in range(E)], [_RMS(()) for _ in range(E)]
    returns, ep_r, ep_l = np.zeros(E), np.zeros(E), np.zeros(E, dtype=int)

    def norm_obs(o):
        out = np.empty_like(o, dtype=np.float64)
        for e in range(E):
            obs_rms[e].update(o[e].astype(np.float64))
            out[e] = np.clip((o[e] - obs_rms[e].mean) / np.sqrt(obs_rms[e].var + 1e-8), -1, 1)
        return out

    o_v, _ = envs.reset()
    o_r, _ = raw.reset()
    np.testing.assert_allclose(o_v, norm_obs(o_r.reshape(E, D).cpu().numpy()), rtol=0, atol=1e-6)
    rs = np.random.RandomState(0)
    finished_total = 0
    for t in range(T):
        act = rs.uniform(-1.5, 1.5, (E, 2)).astype(np.float32)
        o_v, r_v, term_v, trunc_v, infos = envs.step(act)
        o_r, r_r, term_r, trunc_r, _ = raw.step(torch.as_tensor(np.clip(act, -1, 1)).cuda())
        o_r, r_r = o_r.reshape(E, D).cpu().numpy(), r_r.cpu().numpy().astype(np.float64)
        term_r, trunc_r = term_r.cpu().numpy(), trunc_r.cpu().numpy()
        np.testing.assert_array_equal(term_v, term_r); np.testing.assert_array_equal(trunc_v, trunc_r)
        np.testing.assert_allclose(o_v, norm_obs(o_r), rtol=0, atol=1e-6)
        want_r = np.empty(E)
        for e in range(E):
            returns[e] = returns[e] * gamma * (1 - term_r[e]) + r_r[e]
            ret_rms[e].update(returns[e])
            want_r[e] = np.clip(r_r[e] / np.sqrt(ret_rms[e].var + 1e-8), -100, 100)
        np.testing.assert_allclose(r_v, want_r, rtol=1e-6, atol=1e-6)
        ep_r += r_r; ep_l += 1
        done = term_r | trunc_r
        if done.any():
            assert "final_info" in infos and np.array_equal(infos["_final_info"], done)
            for e in np.nonzero(done)[0]:
                ep = infos["final_info"][e]["episode"]
                assert int(ep["l"][0]) == ep_l[e] == max_t or term_r[e]
                np.testing.assert_allclose(float(ep["r"][0]), ep_r[e], rtol=1e-5)
                ep_r[e], ep_l[e] = 0.0, 0
                finished_total += 1
            assert all(infos["final_info"][e] is None for e in np.nonzero(~done)[0])
        else:
            assert "final_info" not in infos
    assert finished_total >= 3 * E   # truncation every 40 steps


@pytest.mark.parametrize("wrap_kw,obs_dim", [(dict(), 4 + 2 * 20), (dict(positions="grav", alpha=3), 6), (dict(positions="rel", statuses="cat"), 4 + 3 * 20),
                                             (dict(positions="rel", statuses="ohe", type="Box"), 22 * 6)])
@pytest.mark.parametrize("E", [1, 4])
def test_vector_env_every_wrapper_configuration_and_one_env(wrap_kw, obs_dim, E):
    """Dict observations (the default EnvWrappersConfig, and positions='grav' as in run_scripts/learn_grav_emb.sh)
    and the class default num_envs=1 go through the same flat row as the Box configuration: FlattenObservation order =
    sorted Dict keys, checked against the structured observation of the raw batched env."""
    import evacuation_b200 as eb

    cfg = eb.EnvConfig(number_of_pedestrians=20, max_timesteps=7)
    envs = eb.EvacuationVectorEnv(cfg, eb.EnvWrappersConfig(**wrap_kw), num_envs=E, seed=5, output="torch")
    raw = eb.setup_env(cfg, eb.EnvWrappersConfig(**wrap_kw), num_envs=E, seed=5, auto_reset=True, batched=True)
    assert envs.single_observation_space.shape == (obs_dim,)
    o_v, _ = envs.reset()
    o_r, _ = raw.reset()

    def flat(o):
        if isinstance(o, dict):
            return torch.cat([o[k].reshape(E, -1) for k in sorted(o)], dim=1)
        return o.reshape(E, -1)

    assert o_v.shape == (E, obs_dim) and torch.equal(flat(o_r), raw.unwrapped.flat_observation)
    for t in range(10):  # crosses a truncation + same-step auto-reset
        act = torch.full((E, 2), 0.3, device="cuda")
        o_v, r_v, term, trunc, infos = envs.step(act)
        o_r, r_r, term_r, trunc_r, _ = raw.step(act)
        assert o_v.shape == (E, obs_dim) and bool(torch.isfinite(o_v).all()) and float(o_v.abs().max()) <= 1.0
        assert torch.equal(flat(o_r), raw.unwrapped.flat_observation)
        assert torch.equal(term, term_r) and torch.equal(trunc, trunc_r)
        if t == 6:
            assert bool(trunc.all()) and infos["_final_info"].all()
    envs.close()
    with_numpy = eb.EvacuationVectorEnv(cfg, eb.EnvWrappersConfig(**wrap_kw), num_envs=E, seed=5)  # SyncVectorEnv-like host arrays
    o, _ = with_numpy.reset()
    o2, r, tm, tr, _ = with_numpy.step(np.zeros((E, 2), np.float32) + 0.3)
    assert isinstance(o2, np.ndarray) and o2.shape == (E, obs_dim) and r.shape == (E,) and tm.dtype == bool
    with_numpy.close()
