"""GPU tests of the scripted on-device agents and the per-step status statistics of `evac_rollout` (SURVEY section 8
row f2): the in-kernel WacuumCleaner state machine against the batched torch state machine (itself pinned to the action
trace of the unmodified reference class by tests/test_rollout_cpu.py), and the status-count trace / efficiency curve
against step-by-step bookkeeping.  Integer outputs must match exactly; states bit for bit (same kernel arithmetic)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(n, E=24, seed=11, **kw):
    import evacuation_b200 as eb
    cfg = dict(number_of_pedestrians=n, is_new_exiting_reward=True)
    cfg.update(kw.pop("env_kw", {}))
    return eb.setup_env(eb.EnvConfig(**cfg), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                        num_envs=E, seed=seed, **kw)


@pytest.mark.parametrize("n,kw", [(60, {}), (60, dict(env_kw=dict(max_timesteps=300))), (200, {}),
                                   (33, dict(env_kw=dict(width=1.5, height=0.8, max_timesteps=700)))])
def test_wacuum_rollout_matches_the_batched_state_machine(n, kw):
    """K steps inside one launch with agent="wacuum" == K single steps driven by WacuumCleaner.batched (re-armed on every
    finished episode), including same-step auto-resets."""
    from evacuation_b200.agents import WacuumCleaner
    K = 900
    a, b = _env(n, auto_reset=True, **dict(kw)), _env(n, auto_reset=True, **dict(kw))
    a.reset(); b.reset()
    ua, ub = a.unwrapped, b.unwrapped
    done_a = 0
    while done_a < K:  # several launches: the state machine persists in the handle between them
        a.rollout(300, agent="wacuum")
        done_a += 300
    agent = WacuumCleaner.batched(b)
    rewards = torch.zeros(ub.num_envs, device="cuda")
    for t in range(K):
        pos = ub.get_state()["agent_position"]
        act = agent.act(pos).to(torch.float32).contiguous()
        _, r, term, trunc, _ = b.step(act)
        agent.reset(term | trunc)
    sa, sb = ua.get_state(), ub.get_state()
    for k in ("positions", "directions", "statuses", "agent_position", "agent_direction", "now"):
        assert torch.equal(sa[k], sb[k]), k
    ta, tb = ua.episode_statistics()[2], ub.episode_statistics()[2]
    assert torch.equal(ta, tb)  # same finished episodes, same statistics
    if kw.get("env_kw", {}).get("max_timesteps", 2000) < K:
        assert float(ta[0]) > 0  # truncated episodes were re-armed inside the launch


def test_wacuum_sweep_evacuates_better_than_a_random_leader():
    """Sanity of the baseline's purpose (plot.py 'Quantification of evacuation efficiency'): after 2000 steps the sweep has
    evacuated more pedestrians than the random leader."""
    env = _env(60, E=256, auto_reset=False)
    m_w, _ = env.unwrapped.efficiency_curve(2000, agent="wacuum")
    m_r, _ = env.unwrapped.efficiency_curve(2000, agent="random")
    assert float(m_w[-1]) > float(m_r[-1]) + 5.0, (float(m_w[-1]), float(m_r[-1]))


@pytest.mark.parametrize("n", [60, 200])
def test_status_count_trace_matches_stepwise_bookkeeping(n):
    E, K = 16, 120
    a, b = _env(n, E=E, auto_reset=False), _env(n, E=E, auto_reset=False)
    a.reset(); b.reset()
    torch.manual_seed(0)
    acts = (torch.rand((K, E, 2), device="cuda") * 2 - 1)
    a.unwrapped.rollout(K, agent="table", actions=acts, status_counts=True)
    trace = a.unwrapped.last_status_counts.cpu().numpy()
    assert trace.shape == (K, E, 4) and (trace.sum(axis=2) == n).all()
    for t in range(K):
        b.step(acts[t])
        st = b.unwrapped.get_state()["statuses"].cpu().numpy()
        want = np.stack([(st == 4).sum(1), (st == 3).sum(1), (st == 2).sum(1), (st == 1).sum(1)], axis=1)
        np.testing.assert_array_equal(trace[t], want)


def test_efficiency_curve_is_monotone_and_bounded():
    n = 60
    env = _env(n, E=128, auto_reset=False)
    mean, std = env.unwrapped.efficiency_curve(600, agent="rotating", chunk=250)
    mean, std = mean.cpu().numpy(), std.cpu().numpy()
    assert mean.shape == (600,) and std.shape == (600,)
    assert (np.diff(mean) >= -1e-12).all() and mean[0] >= 0 and mean[-1] <= n   # ESCAPED is absorbing
    assert (std >= 0).all()


def test_rotating_agent_phase_survives_auto_reset():
    """the reference's RotatingAgent counts its act() calls (rotating_agent.py:8-16) and never restarts with an
    episode; the on-device agent therefore takes its phase from overall_timesteps, not from the in-episode step."""
    import evacuation_b200 as eb

    n, E, T, max_t = 5, 3, 20, 7
    env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=n, max_timesteps=max_t), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                       num_envs=E, seed=2, auto_reset=True)
    env.reset()
    obs, _, _, trunc = env.rollout(T, agent="rotating", obs_every_step=True)
    got = obs[:, :, 0, :2].cpu().numpy()  # agent row of the Box observation: absolute agent position after every step
    pos = np.zeros(2, dtype=np.float32)
    for t in range(1, T + 1):
        a = np.array([np.sin(0.05 * t), np.cos(0.05 * t)], dtype=np.float32)
        a = a / (np.sqrt(a[0] * a[0] + a[1] * a[1]) + np.float32(1e-8))
        pos = pos + np.float32(0.01) * a
        if t % max_t == 0:  # same-step auto-reset: the observation shows the fresh episode, the agent back at the origin
            pos = np.zeros(2, dtype=np.float32)
        assert np.allclose(got[t - 1], pos[None, :], atol=1e-6), (t, got[t - 1], pos)
    assert bool(trunc.all())
