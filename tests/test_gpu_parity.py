"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.

Run on the B200 box with `pytest -m gpu`.  Bars (BASELINE.json north_star): statuses, escaped
counts, termination / truncation flags BIT-EXACT; positions, directions, rewards, observations
within 1e-5 relative.  Because the reference state is float64 and the product kernel float32, a
thresholded decision (pair inside the vision radius, status switch, wall hit) may legitimately
differ when the float64 quantity sits within ~1e-7 of its threshold; the oracle reports that
distance (`StepInfo.margin`) and such steps are excluded from the bit-exact assertion and
COUNTED (near-threshold protocol).
"""
import os

import numpy as np
import pytest

import evac_testlib as T
from oracle.evac_oracle import OracleConfig, OracleEnv, flatten_observation

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

MARGIN_TOL = 1e-6     # teacher-forced: inputs differ from the oracle's only by float32 rounding
POS_RTOL = 1e-5


def _make_env(case_env, case_wrap, num_envs, **kw):
    import evacuation_b200 as eb

    return eb.setup_env(eb.EnvConfig(wandb_enabled=False, **case_env), eb.EnvWrappersConfig(**case_wrap),
                        num_envs=num_envs, batched=True, auto_reset=False, **kw)


def _flat(obs, E):
    if isinstance(obs, dict):
        return torch.cat([obs[k].reshape(E, -1) for k in sorted(obs)], dim=1).double().cpu().numpy()
    return obs.reshape(E, -1).double().cpu().numpy()


def _run_oracle_episode(case, z, T_max=None):
    """Run the oracle over the golden case with INJECTED noise and record pre/post-step states."""
    env = T.make_oracle_at_golden_start(case, z)
    n = env.n
    steps = len(z["rewards"]) if T_max is None else min(T_max, len(z["rewards"]))
    c = case["env"].get("noise_coef", 0.2)
    noise = np.random.RandomState(777 + case["seed"]).uniform(-c / 2, c / 2, size=(steps, n)).astype(np.float32)
    rec = dict(pre_pos=[], pre_dir=[], pre_st=[], pre_apos=[], pre_adir=[], post_pos=[], post_dir=[], post_st=[],
               post_apos=[], reward=[], term=[], trunc=[], obs=[], margin=[], resultant=[])
    for t in range(steps):
        rec["pre_pos"].append(env.positions.copy()); rec["pre_dir"].append(env.directions.copy())
        rec["pre_st"].append(env.statuses.copy()); rec["pre_apos"].append(env.agent_position.copy())
        rec["pre_adir"].append(env.agent_direction.copy())
        obs, r, term, trunc, info = env.step(z["actions_used"][t].copy(), noise[t].astype(np.float64))
        rec["post_pos"].append(env.positions.copy()); rec["post_dir"].append(env.directions.copy())
        rec["post_st"].append(env.statuses.copy()); rec["post_apos"].append(env.agent_position.copy())
        rec["reward"].append(r); rec["term"].append(term); rec["trunc"].append(trunc)
        rec["obs"].append(flatten_observation(obs)); rec["margin"].append(info.margin); rec["resultant"].append(info.min_resultant)
        if term:
            break
    out = {k: np.array(v) for k, v in rec.items()}
    out["noise"] = noise[: len(out["reward"])]
    out["actions"] = z["actions_used"][: len(out["reward"])]
    return out


def _assert_close(name, got, want, rtol=POS_RTOL, scale=1.0, mask=None, ill_conditioned=0.0, cap=10.0):
    """max |got - want| / max(|want|, scale) <= rtol.  `ill_conditioned` > 0 (directions only): the new
    heading is the ARGUMENT of a sum of unit vectors, whose float32 rounding error (~1e-7) is amplified
    by 1/|sum| when neighbours nearly cancel; that fraction of the entries may reach cap x rtol.  Directions
    are recomputed from scratch every step, so this error does not accumulate; positions (which integrate the
    directions) are always held to the strict bound."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if mask is not None:
        got, want = got[mask], want[mask]
    assert np.array_equal(np.isnan(got), np.isnan(want)), f"{name}: NaN pattern differs"
    err = np.abs(got - want) / np.maximum(np.abs(want), scale)
    err = np.nan_to_num(err, nan=0.0)
    if err.size == 0:
        return 0.0
    if ill_conditioned > 0.0:
        assert err.max() <= cap * rtol, f"{name}: max rel err {err.max():.3e} > {cap * rtol:.1e}"
        frac = float((err > rtol).mean())
        assert frac <= ill_conditioned, f"{name}: {frac:.2e} of the entries exceed {rtol:.1e}"
    else:
        assert err.max() <= rtol, f"{name}: max rel err {err.max():.3e} > {rtol:.1e}"
    return float(err.max())


def _assert_obs_close(wrap, got, want, mask=None, rtol=2e-5):
    """Observation rows [E,D].  Gravity encoding: each gradient is a SUM of terms ~1/r^(alpha+1) of both
    signs, so the float32 error is relative to the largest term, not to the (possibly cancelled) sum:
    compare relative to the row's largest magnitude."""
    got, want = np.atleast_2d(got), np.atleast_2d(want)
    if mask is not None:
        got, want = got[mask], want[mask]
    if wrap.get("positions") == "grav":
        with np.errstate(invalid="ignore"):
            scale = np.maximum(np.nanmax(np.abs(want), axis=1, keepdims=True), 1.0)
        assert np.array_equal(np.isfinite(got), np.isfinite(want))
        fin = np.isfinite(want)
        err = np.where(fin, np.abs(got - np.where(fin, want, 0)) / scale, 0.0)
        assert err.max(initial=0.0) <= rtol, f"grav observation: max scaled err {err.max():.3e}"
    else:
        _assert_close("observation", got, want, rtol=rtol)


@pytest.mark.parametrize("name", T.golden_names())
def test_teacher_forced_one_step_parity(name):
    """Every step of every golden episode as an independent one-step problem: env t of a batch of T
    environments starts from the oracle's float64 state before step t (rounded to float32) and must
    reproduce the oracle's state, reward, flags and observation after step t.  One kernel launch."""
    case, z = T.load_golden(name)
    rec = _run_oracle_episode(case, z, T_max=600)
    steps, n = len(rec["reward"]), rec["pre_pos"].shape[1]
    env = _make_env(case["env"], case["wrap"], steps)
    u = env.unwrapped
    u.reset()
    u.set_state(positions=rec["pre_pos"], directions=rec["pre_dir"], statuses=rec["pre_st"], agent_position=rec["pre_apos"],
                agent_direction=rec["pre_adir"], now=np.arange(steps, dtype=np.int32))
    obs, reward, term, trunc, _ = env.step(torch.as_tensor(rec["actions"]), noise=torch.as_tensor(rec["noise"]))
    st = u.get_state()
    ok = rec["margin"] > MARGIN_TOL  # steps whose thresholded decisions are numerically unambiguous
    n_skip = int((~ok).sum())
    assert n_skip <= max(2, steps // 50), f"too many near-threshold steps: {n_skip}/{steps}"
    got_st = st["statuses"].cpu().numpy()
    assert np.array_equal(got_st[ok], rec["post_st"][ok]), "statuses differ on an unambiguous step"
    assert np.array_equal(term.cpu().numpy()[ok], rec["term"][ok])
    assert np.array_equal(trunc.cpu().numpy(), rec["trunc"])
    assert np.array_equal(st["now"].cpu().numpy(), np.arange(1, steps + 1))
    step_size = case["env"].get("step_size", 0.01)
    same = ok & (got_st == rec["post_st"]).all(axis=1)
    _assert_close("agent_position", st["agent_position"].cpu().numpy(), rec["post_apos"], mask=same)
    _assert_close("positions", st["positions"].cpu().numpy(), rec["post_pos"], mask=same)
    # (an ESCAPED pedestrian's direction is dead state -- zeroed at the start of the next step, area.py:79-80 --
    #  and may differ in sign when the pedestrian lands exactly on the exit, which sits ON the wall y = -1)
    alive = (rec["post_st"] != 4)[..., None]
    _assert_close("directions", st["directions"].cpu().numpy() * alive, rec["post_dir"] * alive, scale=step_size, mask=same, ill_conditioned=1e-3)
    _assert_close("reward", reward.cpu().numpy(), rec["reward"], mask=same)
    _assert_obs_close(case["wrap"], _flat(obs, steps), rec["obs"], same)


def _record_summary(kind, name, summary):
    """Per-case parity numbers -> gpurun_out/parity/<kind>__<name>.json (merged back by gpurun; the round's copy is
    committed under profiles/): what the free-running comparison actually observed, not only that it stayed under a bound."""
    import json

    out = os.path.join(T.ROOT, "gpurun_out", "parity")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"{kind}__{name}.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


# free-running bars (north_star: 1e-5 relative over a 2000-step episode in fp32).  POSITIONS / REWARDS: strict 1e-5 on
# every step.  DIRECTIONS (relative to step_size): 1e-5 except on the entries whose heading is ill-conditioned on that step
# (|mean of the neighbour unit vectors| small: a float32 rounding of ~6e-8 in the sum turns into 6e-8 / |mean| of angle) --
# those may reach FREE_DIR_CAP and must stay a small fraction of all entries.  Values set from the recorded summaries
# (profiles/r02*/parity/): see DESIGN.md section 4.
FREE_POS_RTOL = 1e-5
FREE_DIR_CAP = 2e-4   # observed (profiles/r02a/parity): worst 6.7e-5 of a step, on an ill-conditioned heading
FREE_DIR_FRAC = 2e-3  # observed: at most 7.5e-4 of the entries above 1e-5


@pytest.mark.parametrize("name", ["c2_rel_ohe_box_seed0", "c2_rel_ohe_box_seed3", "c3_grav_a3", "c1_default_seed0", "c1_default_seed1",
                                  "n10_abs_cat_dict", "n33_rel_no_dict", "n60_box_1p5x0p8", "n3_escape_all"])
def test_free_running_episode_parity(name):
    """Free-running episode (up to 2000 steps): the kernel carries its own float32 state, the oracle its float64 state;
    same actions and injected noise; the kernel state is NEVER re-synchronised because of numerical error -- only when a
    thresholded decision flips at a step whose oracle margin is below the accumulated float32 drift (a status / neighbour
    set that legitimately differs; at most 1 per episode, asserted to be a near-threshold step; none observed so far).  Everything observed
    is recorded (resyncs, ill-conditioned steps, worst errors and when) in gpurun_out/parity/."""
    case, z = T.load_golden(name)
    rec = _run_oracle_episode(case, z)
    steps = len(rec["reward"])
    env = _make_env(case["env"], case["wrap"], 1)
    u = env.unwrapped
    u.reset()
    u.set_state(positions=rec["pre_pos"][0], directions=rec["pre_dir"][0], statuses=rec["pre_st"][0],
                agent_position=rec["pre_apos"][0], agent_direction=rec["pre_adir"][0], now=np.zeros(1, np.int32))
    step_size = case["env"].get("step_size", 0.01)
    actions = torch.as_tensor(rec["actions"]).cuda()
    noise = torch.as_tensor(rec["noise"]).cuda()
    pos_t, dir_t, st_t, rew_t, term_t, trunc_t = [], [], [], [], [], []
    resync_steps = []
    for t in range(steps):
        obs, reward, term, trunc, _ = env.step(actions[t:t + 1], noise=noise[t:t + 1])
        st = u.get_state()
        got_st = st["statuses"][0].cpu().numpy()
        pos_t.append(st["positions"][0].cpu().numpy()); dir_t.append(st["directions"][0].cpu().numpy()); st_t.append(got_st)
        rew_t.append(float(reward[0])); term_t.append(bool(term[0])); trunc_t.append(bool(trunc[0]))
        if not np.array_equal(got_st, rec["post_st"][t]) or bool(term[0]) != bool(rec["term"][t]):
            assert rec["margin"][t] < 2e-5, f"step {t}: status mismatch with oracle margin {rec['margin'][t]:.3e}"
            resync_steps.append(t)
            u.set_state(positions=rec["post_pos"][t], directions=rec["post_dir"][t], statuses=rec["post_st"][t],
                        agent_position=rec["post_apos"][t])
    pos_t, dir_t, st_t = np.array(pos_t, dtype=np.float64), np.array(dir_t, dtype=np.float64), np.array(st_t)
    ok = np.ones(steps, dtype=bool)
    ok[resync_steps] = False
    assert np.array_equal(np.array(trunc_t), rec["trunc"].astype(bool))
    assert np.array_equal(st_t[ok], rec["post_st"][ok]) and np.array_equal(np.array(term_t)[ok], rec["term"][ok].astype(bool))
    # errors per step (NaN patterns must agree; none of these goldens goes NaN)
    pos_err = np.abs(pos_t - rec["post_pos"]) / np.maximum(np.abs(rec["post_pos"]), 1.0)
    alive = (rec["post_st"] != 4)[..., None]
    dir_err = np.abs(dir_t * alive - rec["post_dir"] * alive) / np.maximum(np.abs(rec["post_dir"] * alive), step_size)
    rew_err = np.abs(np.array(rew_t) - rec["reward"]) / np.maximum(np.abs(rec["reward"]), 1.0)
    assert np.isfinite(pos_err).all() and np.isfinite(dir_err).all()
    pos_step, dir_step = pos_err.reshape(steps, -1).max(axis=1), dir_err.reshape(steps, -1).max(axis=1)
    pos_step[~ok] = 0.0; dir_step[~ok] = 0.0; rew_err[~ok] = 0.0
    illcond = rec["resultant"] < 0.02
    marks = [m for m in (100, 500, 1000, 1500, 2000) if m <= steps]
    summary = dict(
        case=name, steps=int(steps), status_resyncs=len(resync_steps), resync_steps=[int(x) for x in resync_steps],
        resync_margins=[float(rec["margin"][x]) for x in resync_steps], illcond_steps=int(illcond.sum()),
        worst_pos=float(pos_step.max()), worst_pos_step=int(pos_step.argmax()), worst_dir=float(dir_step.max()), worst_dir_step=int(dir_step.argmax()),
        worst_dir_well_conditioned=float(dir_step[~illcond].max(initial=0.0)), worst_reward=float(rew_err.max()),
        dir_entries_above_1e5=float((dir_err[ok] > 1e-5).mean()) if ok.any() else 0.0,
        pos_steps_above_1e5=int((pos_step > 1e-5).sum()),
        worst_pos_up_to={str(m): float(pos_step[:m].max()) for m in marks}, worst_dir_up_to={str(m): float(dir_step[:m].max()) for m in marks},
        bars=dict(pos=FREE_POS_RTOL, dir_cap=FREE_DIR_CAP, dir_frac=FREE_DIR_FRAC, max_status_resyncs=1))
    _record_summary("free_running", name, summary)
    print(f"[{name}] {summary}")
    assert len(resync_steps) <= 1, f"{len(resync_steps)} near-threshold re-synchronisations in {steps} steps"  # observed: 0 on every golden
    assert pos_step.max() <= FREE_POS_RTOL, f"positions: max rel err {pos_step.max():.3e} at step {pos_step.argmax()}"
    assert rew_err.max() <= FREE_POS_RTOL, f"reward: max rel err {rew_err.max():.3e}"
    assert dir_step.max() <= FREE_DIR_CAP, f"directions: max err {dir_step.max():.3e} of a step at step {dir_step.argmax()}"
    assert summary["dir_entries_above_1e5"] <= FREE_DIR_FRAC, f"directions: {summary['dir_entries_above_1e5']:.2e} of the entries exceed 1e-5"


@pytest.mark.parametrize("n,num_envs,steps", [(1, 7, 3), (2, 5, 3), (31, 3, 3), (32, 3, 3), (64, 4, 3), (65, 3, 2), (128, 2, 2), (129, 2, 2),
                                              (257, 2, 2), (600, 2, 2), (1500, 1, 1), (4096, 1, 1)])
@pytest.mark.parametrize("wrap", [dict(positions="rel", statuses="ohe", type="Box"), dict(positions="grav", alpha=2)])
@pytest.mark.parametrize("search", ["auto", "other"])
def test_kernel_shapes_random_states(n, num_envs, steps, wrap, search):
    """All CTA shapes (THREADS x PPT), ragged N, random layouts, both neighbour searches ("auto" = all-pairs warp tile
    for N <= 64 / cell list above; "other" = strip-culled warp tile for N <= 64 / all-pairs tiles above):
    teacher-forced against the oracle."""
    if search == "other":
        search = "cells" if n <= 64 else "brute"
    rs = np.random.RandomState(n)
    env_kw = dict(number_of_pedestrians=n, is_new_exiting_reward=True, intrinsic_reward_coef=0.3, enslaving_degree=0.7)
    cfg = OracleConfig(**env_kw, **wrap)
    env = _make_env(env_kw, wrap, num_envs, neighbor_search=search)
    u = env.unwrapped
    u.reset()
    assert (u.num_cells > 0) == ((n > 64) == (search == "auto"))
    oracles = []
    for e in range(num_envs):
        o = OracleEnv(cfg)
        np.random.seed(1000 * n + e)
        o.reset()
        o.agent_position = rs.uniform(-0.5, 0.5, 2).astype(np.float32)
        o.statuses = T.compute_statuses(o.positions, o.agent_position, o.exit_position)[0]
        oracles.append(o)
    for s in range(steps):
        u.set_state(positions=np.stack([o.positions for o in oracles]), directions=np.stack([o.directions for o in oracles]),
                    statuses=np.stack([o.statuses for o in oracles]), agent_position=np.stack([o.agent_position for o in oracles]),
                    agent_direction=np.stack([o.agent_direction for o in oracles]), now=np.array([o.now for o in oracles], np.int32))
        actions = rs.uniform(-1, 1, (num_envs, 2)).astype(np.float32)
        noise = rs.uniform(-0.1, 0.1, (num_envs, n)).astype(np.float32)
        obs, reward, term, trunc, _ = env.step(torch.as_tensor(actions), noise=torch.as_tensor(noise))
        st = u.get_state()
        flat = _flat(obs, num_envs)
        for e, o in enumerate(oracles):
            oobs, r, tm, tr, info = o.step(actions[e].copy(), noise[e].astype(np.float64))
            if info.margin < MARGIN_TOL:
                o.set_state(st["positions"][e].cpu().numpy(), st["directions"][e].cpu().numpy(), st["statuses"][e].cpu().numpy(),
                            st["agent_position"][e].cpu().numpy(), st["agent_direction"][e].cpu().numpy(), now=o.now)
                continue
            assert np.array_equal(st["statuses"][e].cpu().numpy(), o.statuses)
            assert bool(term[e]) == bool(tm) and bool(trunc[e]) == bool(tr)
            _assert_close("positions", st["positions"][e].cpu().numpy(), o.positions)
            alive = (o.statuses != 4)[:, None]
            _assert_close("directions", st["directions"][e].cpu().numpy() * alive, o.directions * alive, scale=0.01, ill_conditioned=5e-2)
            _assert_close("reward", reward[e].item(), r)
            _assert_obs_close(wrap, flat[e], flatten_observation(oobs), rtol=5e-5)


@pytest.mark.parametrize("n,width,height,vision", [(60, 1.0, 1.0, 0.1), (64, 1.5, 0.8, 0.1), (33, 1.0, 1.0, 0.3), (4096, 1.0, 1.0, 0.1), (1000, 1.5, 0.8, 0.1), (300, 1.0, 1.0, 0.35), (2048, 1.0, 1.0, 0.011), (200, 1.0, 1.0, 0.9),
                                                   (500, 0.3, 0.3, 0.1), (8192, 1.0, 1.0, 0.1), (6000, 1.0, 1.0, 0.02)])
def test_cell_list_equals_all_pairs_search(n, width, height, vision, monkeypatch):
    """Large crowds (BASELINE config 4): the cell-list neighbour search evaluates the same predicate on a superset
    of the contributing pairs, so a free-running rollout must retrace the all-pairs kernel (only the float32
    summation order differs).  Includes a dense cluster (hundreds of pedestrians in one cell), non-square arenas,
    a grid that hits the 64 x 64 cap, a 2 x 2 grid, and run-to-run determinism of the shared-memory-atomic build."""
    import evacuation_b200 as eb

    monkeypatch.setattr(eb.SwitchDistances, "to_pedestrian", vision)
    E, steps = 3, 12
    env_kw = dict(number_of_pedestrians=n, width=width, height=height, is_new_exiting_reward=True, intrinsic_reward_coef=0.3,
                  enslaving_degree=0.6, noise_coef=0.4)
    wrap = dict(positions="rel", statuses="ohe", type="Box")
    rs = np.random.RandomState(n)
    pos = rs.uniform(-1, 1, (E, n, 2)) * np.array([width, height])
    pos[1, : n // 2] = np.array([0.3 * width, -0.2 * height]) + rs.normal(0, 0.03, (n // 2, 2))  # dense cluster
    pos = np.clip(pos, [-width, -height], [width, height])
    ang = rs.uniform(0, 2 * np.pi, (E, n))
    dirs = np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    actions = rs.uniform(-1, 1, (steps, E, 2)).astype(np.float32)
    noise = rs.uniform(-0.2, 0.2, (steps, E, n)).astype(np.float32)
    out = {}
    for key, search in (("cells", "cells"), ("cells2", "cells"), ("brute", "brute")):
        env = _make_env(env_kw, wrap, E, neighbor_search=search)
        u = env.unwrapped
        u.reset()
        assert (u.num_cells > 0) == (search == "cells")
        u.set_state(positions=pos, directions=dirs, agent_position=np.zeros((E, 2), np.float32), agent_direction=np.zeros((E, 2), np.float32),
                    now=np.zeros(E, np.int32))
        rew = []
        for s in range(steps):
            obs, r, _, _, _ = env.step(torch.as_tensor(actions[s]), noise=torch.as_tensor(noise[s]))
            rew.append(r.cpu().numpy().copy())
        st = u.get_state()
        out[key] = dict(pos=st["positions"].cpu().numpy(), dir=st["directions"].cpu().numpy(), st=st["statuses"].cpu().numpy(),
                        rew=np.array(rew), obs=obs.cpu().numpy().copy())
        u.close()
    for k in ("pos", "dir", "st", "rew", "obs"):  # deterministic: identical bits run to run
        assert np.array_equal(out["cells"][k], out["cells2"][k], equal_nan=True), k
    assert (out["cells"]["st"] != out["brute"]["st"]).sum() <= 2
    # (an ill-conditioned sum -- neighbours nearly cancelling -- turns a 1-ulp difference of the sum into a visible angle)
    assert np.abs(out["cells"]["pos"] - out["brute"]["pos"]).max() <= 5e-4
    close = np.abs(out["cells"]["pos"] - out["brute"]["pos"]) <= 1e-6
    assert close.mean() >= 0.999
    np.testing.assert_allclose(out["cells"]["rew"], out["brute"]["rew"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("n,width,height,vision", [(4096, 1.0, 1.0, 0.1), (1000, 1.5, 0.8, 0.1), (300, 1.0, 1.0, 0.35), (2048, 1.0, 1.0, 0.011), (129, 1.0, 1.0, 0.1),
                                                   (8192, 1.0, 1.0, 0.1)])
def test_cell_list_paired_walk_is_bit_identical_to_the_single_slot_walk(n, width, height, vision, monkeypatch):
    """cell_list_pass step 5 has two walks of the sorted slots: one slot per thread, or two adjacent slots of one cell row
    sharing the union of their windows (the default from half a pedestrian per cell).  A slot outside a pedestrian's own
    window fails the distance test and adds an exact zero, so both walks must give the same bits -- states, rewards and
    observations -- over a free-running rollout, dense cluster and odd row lengths included."""
    import evacuation_b200 as eb

    monkeypatch.setattr(eb.SwitchDistances, "to_pedestrian", vision)
    E, steps = 3, 10
    env_kw = dict(number_of_pedestrians=n, width=width, height=height, is_new_exiting_reward=True, intrinsic_reward_coef=0.3,
                  enslaving_degree=0.6, noise_coef=0.4)
    wrap = dict(positions="rel", statuses="ohe", type="Box")
    rs = np.random.RandomState(7 + n)
    pos = rs.uniform(-1, 1, (E, n, 2)) * np.array([width, height])
    pos[1, : n // 2] = np.array([-0.4 * width, 0.5 * height]) + rs.normal(0, 0.03, (n // 2, 2))
    pos = np.clip(pos, [-width, -height], [width, height])
    ang = rs.uniform(0, 2 * np.pi, (E, n))
    dirs = np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    actions = rs.uniform(-1, 1, (steps, E, 2)).astype(np.float32)
    out = {}
    for walk in ("0", "1", "g"):  # one slot per thread / two adjacent slots per thread / four lanes sharing the window of 8 slots
        monkeypatch.setenv("EVAC_CELL_PAIR_WALK", "0" if walk == "0" else "1")
        monkeypatch.setenv("EVAC_CELL_GROUP", "1" if walk == "g" else "0")
        env = _make_env(env_kw, wrap, E, neighbor_search="cells")
        u = env.unwrapped
        u.reset()
        u.set_state(positions=pos, directions=dirs, agent_position=np.zeros((E, 2), np.float32), agent_direction=np.zeros((E, 2), np.float32),
                    now=np.zeros(E, np.int32))
        rew = []
        for s in range(steps):
            obs, r, _, _, _ = env.step(torch.as_tensor(actions[s]))
            rew.append(r.cpu().numpy().copy())
        st = u.get_state()
        out[walk] = dict(pos=st["positions"].cpu().numpy(), dir=st["directions"].cpu().numpy(), st=st["statuses"].cpu().numpy(),
                         rew=np.array(rew), obs=obs.cpu().numpy().copy())
        u.close()
    for k in ("pos", "dir", "st", "rew", "obs"):
        assert np.array_equal(out["0"][k], out["1"][k], equal_nan=True), k
        assert np.array_equal(out["0"][k], out["g"][k], equal_nan=True), ("grouped walk", k)


def test_pedestrian_count_limits():
    """1 .. 32768 pedestrians per environment in float32 with the cell list (one CTA up to 4096, a cluster of 2 / 4 / 8 CTAs
    above), 1 .. 8192 with the all-pairs search, 1 .. 4096 in the fp64 parity mode; anything else fails loudly at creation."""
    import evacuation_b200 as eb

    wrap = eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box")
    with pytest.raises(ValueError, match="32768"):
        eb.setup_env(eb.EnvConfig(number_of_pedestrians=32769), wrap, num_envs=2, batched=True).reset()
    with pytest.raises(ValueError, match="all-pairs"):
        eb.setup_env(eb.EnvConfig(number_of_pedestrians=9000), wrap, num_envs=2, batched=True, neighbor_search="brute").reset()
    with pytest.raises(ValueError, match="fp64"):
        eb.setup_env(eb.EnvConfig(number_of_pedestrians=5000), wrap, num_envs=2, batched=True, precision="fp64").reset()
    for n in (8192, 32768):
        env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=n), wrap, num_envs=2, batched=True, auto_reset=True)
        obs, _ = env.reset()
        assert tuple(obs.shape) == (2, n + 2, 6)
        for _ in range(3):
            obs, r, term, trunc, _ = env.step(torch.zeros((2, 2), device="cuda") + 0.5)
        assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(r).all())
        st = env.unwrapped.get_state()
        assert float(st["positions"].abs().max()) <= 1.0
        # statuses are a pure function of (position, agent position) [statuses.py:29-48]
        exp = T.compute_statuses(st["positions"][0].double().cpu().numpy(), st["agent_position"][0].cpu().numpy(), np.array([0.0, -1.0]))[0]
        assert (exp != st["statuses"][0].cpu().numpy()).mean() < 1e-3  # float32 vs float64 compares right at a threshold
        env.unwrapped.close()


@pytest.mark.parametrize("n,vision", [(8192, 0.1), (5000, 0.1), (8192, 0.02), (700, 0.1)])
def test_cluster_pass_is_bit_identical_across_cluster_sizes(n, vision, monkeypatch):
    """Crowds above 4096 pedestrians run one environment per thread-block cluster (evac_cluster.cuh): the sorted source tile
    is distributed over the CTAs' shared memory and sorted by (cell, pedestrian index) exactly like the one-CTA pass, so the
    one-CTA kernel (1024 x 8, EVAC_CLUSTER=1; itself checked against the all-pairs search) and clusters of 2, 4 and 8 CTAs
    must produce the same state bit for bit over a free-running rollout with auto-reset, a dense cluster of pedestrians
    included (rewards / gravity observations: sums over all pedestrians, identical between cluster sizes)."""
    import evacuation_b200 as eb

    monkeypatch.setattr(eb.SwitchDistances, "to_pedestrian", vision)
    E, steps = 3, 8
    env_kw = dict(number_of_pedestrians=n, is_new_exiting_reward=True, intrinsic_reward_coef=0.3, enslaving_degree=0.6, noise_coef=0.4,
                  max_timesteps=5)
    rs = np.random.RandomState(3 + n)
    pos = rs.uniform(-1, 1, (E, n, 2))
    pos[1, : n // 2] = np.array([0.55, 0.6]) + rs.normal(0, 0.04, (n // 2, 2))
    pos = np.clip(pos, -1, 1)
    ang = rs.uniform(0, 2 * np.pi, (E, n))
    dirs = np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    actions = rs.uniform(-1, 1, (steps, E, 2)).astype(np.float32)
    out = {}
    for wrap in (dict(positions="rel", statuses="ohe", type="Box"), dict(positions="grav", alpha=3)):
        for cl in ("1", "2", "4", "8"):
            monkeypatch.setenv("EVAC_CLUSTER", cl)
            env = eb.setup_env(eb.EnvConfig(wandb_enabled=False, **env_kw), eb.EnvWrappersConfig(**wrap), num_envs=E, batched=True, auto_reset=True)
            u = env.unwrapped
            u.reset()
            u.set_state(positions=pos, directions=dirs, agent_position=np.zeros((E, 2), np.float32), agent_direction=np.zeros((E, 2), np.float32),
                        now=np.zeros(E, np.int32))
            rew, obs_all = [], []
            for s in range(steps):  # the episodes are truncated at step 5 -> same-step auto-reset inside the rollout
                obs, r, _, trunc, _ = env.step(torch.as_tensor(actions[s]))
                rew.append(r.cpu().numpy().copy())
                obs_all.append(_flat(obs, E).copy())
            st = u.get_state()
            out[cl] = dict(pos=st["positions"].cpu().numpy(), dir=st["directions"].cpu().numpy(), st=st["statuses"].cpu().numpy(),
                           rew=np.array(rew), obs=np.array(obs_all))
            u.close()
        for cl in ("4", "8"):  # cluster sizes among themselves: every bit (the per-CTA partial sums are the same ones)
            for k in ("pos", "dir", "st", "rew", "obs"):
                assert np.array_equal(out["2"][k], out[cl][k], equal_nan=True), (wrap, cl, k)
        if vision == 0.02:  # the clustered pass caps the grid at 2048 cells: other cells, other float32 summation order
            assert np.abs(out["1"]["pos"] - out["2"]["pos"]).max() <= 5e-4 and (np.abs(out["1"]["pos"] - out["2"]["pos"]) <= 1e-6).mean() >= 0.999
            assert (out["1"]["st"] != out["2"]["st"]).sum() <= 2
            continue
        for k in ("pos", "dir", "st"):  # against the one-CTA kernel: the state bit for bit ...
            assert np.array_equal(out["1"][k], out["2"][k], equal_nan=True), (wrap, k)
        for k in ("rew", "obs"):  # ... sums over ALL pedestrians (intrinsic reward, gravity observation) group the float32 partials differently
            np.testing.assert_allclose(out["2"][k], out["1"][k], rtol=1e-5, atol=1e-6, err_msg=str((wrap, k)))


@pytest.mark.parametrize("n,cluster", [(16384, "4"), (16384, "8"), (32768, "8"), (32768, "4")])
def test_cluster_path_against_the_oracle(n, cluster, monkeypatch):
    """Crowds above one SM's shared memory (evac_cluster.cuh, area.py:105-119 semantics) against the ORACLE itself, not only
    against the one-CTA kernel: teacher-forced steps at 16 384 and 32 768 pedestrians, a uniform layout (env 0) and one
    with half the crowd in a dense blob (env 1), cluster sizes 4 and 8.  The oracle evaluates the float64 distance matrix in
    row blocks (same numbers, tests/test_oracle_golden.py::test_row_chunked_alignment_is_bit_identical).  With ~1e9 pairs
    SOME pair always sits within 1e-8 of the vision radius, so the near-threshold protocol is applied per pedestrian.  Both
    sides start every step from the SAME float32-representable state, so a pair test can only flip within ~1e-8 of the
    radius (float32 rounding of dx, dy, d^2): a pedestrian whose own closest pair is within 1e-7 of the radius, or whose
    status / wall distance is within 1e-6 of its threshold, is excluded (counted, < 3 % even inside the blob, where
    every pedestrian has ~1000 neighbours); every other pedestrian must match -- statuses bit-exactly."""
    import evacuation_b200 as eb

    monkeypatch.setenv("EVAC_CLUSTER", cluster)
    E, steps = 2, 2 if n <= 16384 else 1
    env_kw = dict(number_of_pedestrians=n, is_new_exiting_reward=True, intrinsic_reward_coef=0.3, enslaving_degree=0.6, noise_coef=0.4)
    wrap = dict(positions="rel", statuses="ohe", type="Box")
    rs = np.random.RandomState(n + int(cluster))
    env = _make_env(env_kw, wrap, E)
    u = env.unwrapped
    u.reset()
    assert u.num_cells > 0
    oracles = []
    for e in range(E):
        o = OracleEnv(OracleConfig(**env_kw, **wrap))
        o.row_chunk, o.chunk_threads = 256, min(16, os.cpu_count() or 1)
        np.random.seed(77 + e)
        o.reset()
        if e == 1:
            o.positions[: n // 4] = np.clip(np.array([0.5, 0.45]) + rs.normal(0, 0.06, (n // 4, 2)), -1, 1)
        o.agent_position = rs.uniform(-0.5, 0.5, 2).astype(np.float32)
        o.statuses = T.compute_statuses(o.positions, o.agent_position, o.exit_position)[0]
        oracles.append(o)
    excluded = []
    for s_ in range(steps):
        for o in oracles:  # teacher forcing from a float32-representable state (what set_state hands the kernel)
            o.positions = o.positions.astype(np.float32).astype(np.float64)
            o.directions = o.directions.astype(np.float32).astype(np.float64)
        u.set_state(positions=np.stack([o.positions for o in oracles]), directions=np.stack([o.directions for o in oracles]),
                    statuses=np.stack([o.statuses for o in oracles]), agent_position=np.stack([o.agent_position for o in oracles]),
                    agent_direction=np.stack([o.agent_direction for o in oracles]), now=np.array([o.now for o in oracles], np.int32))
        actions = rs.uniform(-1, 1, (E, 2)).astype(np.float32)
        noise = rs.uniform(-0.2, 0.2, (E, n)).astype(np.float32)
        obs, reward, term, trunc, _ = env.step(torch.as_tensor(actions), noise=torch.as_tensor(noise))
        st = u.get_state()
        flat = _flat(obs, E).reshape(E, n + 2, 6)
        for e, o in enumerate(oracles):
            pre_st = o.statuses.copy()
            oobs, r, tm, tr, info = o.step(actions[e].copy(), noise[e].astype(np.float64))
            # per-pedestrian ambiguity: own closest pair vs the vision radius, status distances, wall distance
            amb = np.zeros(n, dtype=bool)
            amb[np.nonzero(info.fv_mask)[0][info.pair_margin_rows < 1e-7]] = True
            d_ag = np.linalg.norm(o.positions - o.agent_position.astype(np.float64), axis=1)
            d_ex = np.linalg.norm(o.positions - np.array([0.0, -1.0]), axis=1)
            amb |= (np.abs(d_ag - 0.2) < MARGIN_TOL) | (np.abs(d_ex - 0.4) < MARGIN_TOL) | (np.abs(d_ex - 0.01) < MARGIN_TOL)
            amb |= (np.abs(np.abs(o.positions) - 1.0) < MARGIN_TOL).any(axis=1) & (pre_st <= 2)
            excluded.append(float(amb.mean()))
            ok = ~amb
            got_st = st["statuses"][e].cpu().numpy()
            assert np.array_equal(got_st[ok], o.statuses[ok]), f"env {e}: statuses differ on an unambiguous pedestrian"
            assert bool(trunc[e]) == bool(tr) and (bool(term[e]) == bool(tm) or amb.any())
            _assert_close("positions", st["positions"][e].cpu().numpy()[ok], o.positions[ok])
            alive = (o.statuses != 4)[:, None]
            _assert_close("directions", (st["directions"][e].cpu().numpy() * alive)[ok], (o.directions * alive)[ok], scale=0.01, ill_conditioned=5e-2)
            want = np.asarray(oobs, dtype=np.float64).reshape(n + 2, 6)
            _assert_close("observation rows", flat[e][2:][ok], want[2:][ok], rtol=2e-5)
            _assert_close("observation head", flat[e][:2], want[:2], rtol=2e-5)
            if np.array_equal(got_st, o.statuses):  # the reward counts EVERY pedestrian's status transition
                _assert_close("reward", reward[e].item(), r)
    assert max(excluded) < 0.03, excluded
    _record_summary("cluster_oracle", f"n{n}_cl{cluster}", dict(n=n, cluster=int(cluster), steps=steps, envs=E, excluded_fraction=excluded))


def _ambiguous(o, info, pre_st, pair_tol=1e-7, tol=MARGIN_TOL, w=1.0, h=1.0):
    """Per-pedestrian near-threshold mask of ONE oracle step from a float32-representable state: own closest pair within
    `pair_tol` of the vision radius (a float32 pair test can only flip within ~1e-8), status / wall distance within `tol`."""
    n = o.n
    amb = np.zeros(n, dtype=bool)
    amb[np.nonzero(info.fv_mask)[0][info.pair_margin_rows < pair_tol]] = True
    d_ag = np.linalg.norm(o.positions - o.agent_position.astype(np.float64), axis=1)
    d_ex = np.linalg.norm(o.positions - np.array([0.0, -1.0]), axis=1)
    amb |= (np.abs(d_ag - 0.2) < tol) | (np.abs(d_ex - 0.4) < tol) | (np.abs(d_ex - 0.01) < tol)
    amb |= (np.abs(np.abs(o.positions) - np.array([w, h])) < tol).any(axis=1) & (pre_st <= 2)
    return amb


def test_c4_flocked_trajectory_against_the_oracle():
    """BASELINE config 4 at FULL size (4096 pedestrians, cell list, paired walk) against the oracle along a flocked
    trajectory, not only on a fresh uniform layout: two environments (uniform start / a quarter of the crowd in a dense
    blob) are pre-rolled 120 steps by the kernel, then for 30 more steps
      (A) every step is checked exactly: the oracle is put into the kernel's own (float32) pre-step state and must
          reproduce the kernel's post-step state -- statuses bit-exact, positions 1e-5 -- for every pedestrian that is
          not within float32 reach of a threshold (per-pedestrian near-threshold protocol; the excluded fraction is
          asserted small and recorded);
      (B) a second oracle runs FREE from the step-0 state with the same actions and noise; at this size ~5 pair tests
          per step and environment sit within 1e-6 of the vision radius, each flip turns one heading by ~1/neighbours
          rad and spreads from there, so single pedestrians may leave the 1e-5 band: the distribution is recorded
          (gpurun_out/parity/) and bounded (statuses >= 99.9 % equal, 99th percentile of the position error <= 1e-5)."""
    n, E, pre, steps = 4096, 2, 120, 30
    env_kw = dict(number_of_pedestrians=n, is_new_exiting_reward=True, intrinsic_reward_coef=0.3, enslaving_degree=0.6, noise_coef=0.4)
    wrap = dict(positions="rel", statuses="ohe", type="Box")
    rs = np.random.RandomState(4096)
    env = _make_env(env_kw, wrap, E)
    u = env.unwrapped
    u.reset()
    pos = rs.uniform(-1, 1, (E, n, 2))
    pos[1, : n // 4] = np.clip(np.array([-0.45, 0.4]) + rs.normal(0, 0.06, (n // 4, 2)), -1, 1)
    ang = rs.uniform(0, 2 * np.pi, (E, n))
    u.set_state(positions=pos, directions=np.stack([np.cos(ang), np.sin(ang)], axis=-1), agent_position=np.zeros((E, 2), np.float32),
                agent_direction=np.zeros((E, 2), np.float32), now=np.zeros(E, np.int32))
    env.rollout(pre, agent="random")  # the kernel flocks the crowds on its own
    mk = lambda: OracleEnv(OracleConfig(**env_kw, **wrap))
    step_o, free_o = [mk() for _ in range(E)], [mk() for _ in range(E)]
    for o in step_o + free_o:
        o.row_chunk, o.chunk_threads = 256, min(16, os.cpu_count() or 1)
    st = u.get_state()
    for e, o in enumerate(free_o):
        o.set_state(st["positions"][e].cpu().numpy(), st["directions"][e].cpu().numpy(), st["statuses"][e].cpu().numpy(),
                    st["agent_position"][e].cpu().numpy(), st["agent_direction"][e].cpu().numpy(), now=int(st["now"][e]))
    excluded, free_stats = [], []
    for t in range(steps):
        pre_state = u.get_state()
        actions = rs.uniform(-1, 1, (E, 2)).astype(np.float32)
        noise = rs.uniform(-0.2, 0.2, (E, n)).astype(np.float32)
        obs, reward, term, trunc, _ = env.step(torch.as_tensor(actions), noise=torch.as_tensor(noise))
        post = u.get_state()
        for e in range(E):
            o = step_o[e]
            o.set_state(pre_state["positions"][e].cpu().numpy(), pre_state["directions"][e].cpu().numpy(), pre_state["statuses"][e].cpu().numpy(),
                        pre_state["agent_position"][e].cpu().numpy(), pre_state["agent_direction"][e].cpu().numpy(), now=int(pre_state["now"][e]))
            pre_st = o.statuses.copy()
            oobs, r, tm, tr, info = o.step(actions[e].copy(), noise[e].astype(np.float64))
            amb = _ambiguous(o, info, pre_st)
            excluded.append(float(amb.mean()))
            ok = ~amb
            got_st = post["statuses"][e].cpu().numpy()
            assert np.array_equal(got_st[ok], o.statuses[ok]), f"step {t} env {e}: statuses differ on an unambiguous pedestrian"
            _assert_close("positions", post["positions"][e].cpu().numpy()[ok], o.positions[ok])
            alive = (o.statuses != 4)[:, None]
            _assert_close("directions", (post["directions"][e].cpu().numpy() * alive)[ok], (o.directions * alive)[ok], scale=0.01, ill_conditioned=5e-2)
            if np.array_equal(got_st, o.statuses):
                _assert_close("reward", reward[e].item(), r)
            # (B) the free oracle
            f = free_o[e]
            f.step(actions[e].copy(), noise[e].astype(np.float64))
            perr = np.abs(post["positions"][e].cpu().numpy() - f.positions).max(axis=1)
            free_stats.append(dict(step=t, env=e, status_agreement=float((got_st == f.statuses).mean()), pos_err_median=float(np.median(perr)),
                                   pos_err_p99=float(np.quantile(perr, 0.99)), pos_err_max=float(perr.max()), within_1e5=float((perr <= 1e-5).mean())))
    assert max(excluded) < 0.03, max(excluded)
    last = [s_ for s_ in free_stats if s_["step"] == steps - 1]
    _record_summary("c4_flocked", "n4096", dict(n=n, envs=E, pre_roll=pre, steps=steps, excluded_fraction_max=max(excluded),
                                                excluded_fraction_mean=float(np.mean(excluded)), free_running=free_stats))
    # observed (profiles/r02h/parity/c4_flocked__n4096.json): statuses agree on every pedestrian through step 30, median position
    # error 7e-8, 99.93 % of the pedestrians within 1e-5, worst 1.8e-5 (downstream of a flipped near-threshold pair test)
    assert all(s_["status_agreement"] >= 0.999 for s_ in free_stats)
    assert all(s_["pos_err_p99"] <= 1e-5 and s_["within_1e5"] >= 0.995 for s_ in free_stats), last


def test_cell_list_nan_poisoning_matches_all_pairs():
    """A zero direction (0/0 = NaN unit vector, area.py:101) poisons EVERY neighbour sum in the reference; the cell
    list must reproduce that, not only for the pedestrians whose cells contain the NaN source."""
    n, E = 300, 2
    env_kw = dict(number_of_pedestrians=n)
    wrap = dict(positions="abs", statuses="no", type="Dict")
    rs = np.random.RandomState(5)
    pos = rs.uniform(-1, 1, (E, n, 2))
    pos[:, :, 1] = np.abs(pos[:, :, 1]) * 0.9 + 0.05  # nobody near the exit
    ang = rs.uniform(0, 2 * np.pi, (E, n))
    dirs = np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    dirs[1, 7] = 0.0
    res = {}
    for search in ("cells", "brute"):
        env = _make_env(env_kw, wrap, E, neighbor_search=search)
        u = env.unwrapped
        u.reset()
        u.set_state(positions=pos, directions=dirs, agent_position=np.full((E, 2), 0.9, np.float32))
        env.step(torch.zeros((E, 2)) + 0.1, noise=torch.zeros((E, n)))
        res[search] = u.get_state()["positions"].cpu().numpy()
    assert np.array_equal(np.isnan(res["cells"]), np.isnan(res["brute"]))
    assert not np.isnan(res["cells"][0]).any() and np.isnan(res["cells"][1]).sum() >= 2 * (n - 1)


def test_kat0_through_the_drop_in_api():
    """SURVEY.md 8(c) KAT-0 through the reference-shaped single-env API: np.random.seed(0), setup_env, reset,
    100 steps -- the CUDA path consumes the global NumPy stream exactly like the reference and must
    retrace its trajectory (golden produced by the unmodified reference)."""
    import evacuation_b200 as eb

    case, z = T.load_golden("kat0_rel_ohe_box")
    np.random.seed(0)
    env = eb.setup_env(eb.EnvConfig(wandb_enabled=False, **case["env"]), eb.EnvWrappersConfig(**case["wrap"]))
    obs, info = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (62, 6) and obs.dtype == np.float32 and info == {}
    assert np.allclose(obs.ravel(), z["init_obs"], rtol=1e-6, atol=1e-7)
    u = env.unwrapped
    assert u.pedestrians.status_stats == {"escaped": 0, "exiting": 2, "following": 3, "viscek": 55}
    rewards = []
    for t in range(100):
        a = np.array([np.sin(0.05 * (t + 1)), np.cos(0.05 * (t + 1))], dtype=np.float32)
        obs, r, term, trunc, info = env.step(a)
        assert isinstance(r, float) and isinstance(term, bool) and isinstance(trunc, bool)
        assert np.array_equal(u.pedestrians.statuses, z["statuses"][t]), f"statuses differ at step {t}"
        rewards.append(r)
    assert np.allclose(rewards, z["rewards"], rtol=1e-5)
    assert abs(sum(rewards) - (-90.59406600697285)) < 1e-3
    assert u.pedestrians.status_stats == {"escaped": 6, "exiting": 0, "following": 5, "viscek": 49}
    assert np.allclose(u.agent.position, [0.13844308, -0.1953266], atol=1e-6)
    assert abs(float(u.pedestrians.positions.sum()) - 8.936291463364315) < 1e-4
    assert abs(float(obs.sum()) - 69.00850792787969) < 1e-3


def test_fp64_parity_mode_free_running():
    """precision='fp64': the same kernel template instantiated in double retraces a whole 2000-step
    reference episode with bit-exact statuses and ~1e-12 state error, without any re-synchronisation."""
    case, z = T.load_golden("c2_rel_ohe_box_seed0")
    rec = _run_oracle_episode(case, z)
    steps = len(rec["reward"])
    env = _make_env(case["env"], case["wrap"], 1, precision="fp64")
    u = env.unwrapped
    u.reset()
    u.set_state(positions=rec["pre_pos"][0], directions=rec["pre_dir"][0], statuses=rec["pre_st"][0],
                agent_position=rec["pre_apos"][0], agent_direction=rec["pre_adir"][0], now=np.zeros(1, np.int32))
    actions, noise = torch.as_tensor(rec["actions"]).cuda(), torch.as_tensor(rec["noise"]).cuda()
    statuses, rewards = [], []
    for t in range(steps):
        obs, reward, term, trunc, _ = env.step(actions[t:t + 1], noise=noise[t:t + 1])
        statuses.append(u.get_state()["statuses"][0].clone())
        rewards.append(reward.clone())
    st = u.get_state()
    got = torch.stack(statuses).cpu().numpy()
    assert np.array_equal(got, rec["post_st"]), "fp64 mode: statuses must be bit-exact over the whole episode"
    assert np.abs(st["positions"][0].cpu().numpy() - rec["post_pos"][-1]).max() < 1e-9
    assert np.abs(st["directions"][0].cpu().numpy() - rec["post_dir"][-1]).max() < 1e-11
    assert np.allclose(torch.cat(rewards).cpu().numpy(), rec["reward"], rtol=1e-6)


def test_rollout_equals_single_steps_and_sharding_invariance():
    """K steps in one launch == K launches of one step (bit-exact), and a batch split over two handles
    with env_index_offset draws the same Philox streams as one big batch."""
    import evacuation_b200 as eb

    cfgs = (eb.EnvConfig(number_of_pedestrians=60, max_timesteps=37, is_new_exiting_reward=True), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"))
    K, E = 90, 6
    a = eb.setup_env(*cfgs, num_envs=E, seed=11, auto_reset=True)
    b = eb.setup_env(*cfgs, num_envs=E, seed=11, auto_reset=True)
    c0 = eb.setup_env(*cfgs, num_envs=E // 2, seed=11, auto_reset=True)
    c1 = eb.setup_env(*cfgs, num_envs=E // 2, seed=11, auto_reset=True, env_index_offset=E // 2)
    for env in (a, b, c0, c1):
        env.reset()
    acts = torch.rand((K, E, 2), device="cuda") * 2 - 1
    oa, ra, ta, tra = a.rollout(K, agent="table", actions=acts)
    rsum = torch.zeros(E, device="cuda")
    tb = torch.zeros(E, dtype=torch.bool, device="cuda")
    for k in range(K):
        ob, r, t, tr, _ = b.step(acts[k])
        rsum += r
        tb |= tr
    sa, sb = a.unwrapped.get_state(), b.unwrapped.get_state()
    for key in sa:
        assert torch.equal(sa[key], sb[key]), key
    assert torch.equal(oa, ob) and torch.equal(tra, tb)
    assert torch.allclose(ra, rsum, rtol=1e-5, atol=1e-4)
    # sharding invariance (random agent + Philox noise + auto-reset inside the kernel)
    a.rollout(50, agent="random"); c0.rollout(50, agent="random"); c1.rollout(50, agent="random")
    a2 = a.unwrapped.get_state()
    # bring c0/c1 to the same point: they have not done the table-driven steps, so compare fresh envs instead
    d = eb.setup_env(*cfgs, num_envs=E, seed=11, auto_reset=True)
    d.reset(); d.rollout(50, agent="random")
    sd, s0, s1 = d.unwrapped.get_state(), c0.unwrapped.get_state(), c1.unwrapped.get_state()
    for key in sd:
        assert torch.equal(sd[key], torch.cat([s0[key], s1[key]])), key


def test_philox_autoreset_against_oracle():
    """In-kernel Philox noise, RandomAgent actions and same-step auto-reset against the oracle fed with the
    NumPy restatement of the same streams (tests/evac_testlib.py)."""
    import evacuation_b200 as eb

    n, E, seed, steps, max_t = 24, 5, 2024, 70, 16
    env_kw = dict(number_of_pedestrians=n, max_timesteps=max_t, noise_coef=0.4, enslaving_degree=0.5, is_new_exiting_reward=True)
    env = eb.setup_env(eb.EnvConfig(**env_kw), eb.EnvWrappersConfig(positions="rel", statuses="cat", type="Box"),
                       num_envs=E, seed=seed, auto_reset=True, env_index_offset=100)
    obs0, _ = env.reset()
    oracles, episodes = [], [1] * E
    for e in range(E):
        o = OracleEnv(OracleConfig(**env_kw, positions="rel", statuses="cat", type="Box"))
        pos, dirs = T.philox_layout(seed, 100 + e, 1, n)
        o.set_state(pos, dirs, T.compute_statuses(pos.astype(np.float64), np.zeros(2, np.float32), o.exit_position)[0], np.zeros(2, np.float32))
        oracles.append(o)
        assert np.allclose(obs0[e].cpu().numpy(), o.observation(), atol=1e-6)
    finished_total = 0
    for s in range(steps):
        obs, r, term, trunc = env.rollout(1, agent="random")
        st = env.unwrapped.get_state()
        for e, o in enumerate(oracles):
            act = T.philox_action(seed, 100 + e, episodes[e], o.now)
            nz = T.philox_noise(seed, 100 + e, episodes[e], o.now, n, 0.4).astype(np.float64)
            oobs, orr, otm, otr, info = o.step(act, nz)
            assert bool(trunc[e]) == bool(otr) and bool(term[e]) == bool(otm)
            assert abs(float(r[e]) - orr) <= 1e-4 * max(1.0, abs(orr))
            if otm or otr:  # same-step auto-reset: the oracle is re-seeded from the next episode's layout
                episodes[e] += 1
                finished_total += 1
                pos, dirs = T.philox_layout(seed, 100 + e, episodes[e], n)
                o.set_state(pos, dirs, T.compute_statuses(pos.astype(np.float64), np.zeros(2, np.float32), o.exit_position)[0], np.zeros(2, np.float32))
                o.now = 0
                oobs = o.observation()
            if info.margin < MARGIN_TOL:
                o.set_state(st["positions"][e].cpu().numpy(), st["directions"][e].cpu().numpy(), st["statuses"][e].cpu().numpy(),
                            st["agent_position"][e].cpu().numpy(), st["agent_direction"][e].cpu().numpy(), now=o.now)
                continue
            assert np.array_equal(st["statuses"][e].cpu().numpy(), o.statuses), (s, e)
            assert int(st["now"][e]) == o.now
            assert np.allclose(st["positions"][e].cpu().numpy(), o.positions, atol=2e-5)
            assert np.allclose(obs[e].cpu().numpy(), oobs, atol=2e-5)
            o.set_state(st["positions"][e].cpu().numpy().astype(np.float64), st["directions"][e].cpu().numpy().astype(np.float64),
                        o.statuses, o.agent_position, o.agent_direction, now=o.now)
    stats, fin, totals = env.episode_statistics()
    assert int(totals[0]) == finished_total and finished_total >= E * (steps // max_t)
    assert torch.all(stats[:, 3] == max_t)  # episode_length of truncated episodes


def test_nan_semantics_match_reference():
    """A zero action with full enslaving zeroes the followers' direction; the next step's un-guarded
    normalisation (area.py:101) makes the reference go NaN.  The kernel reproduces that (documented in
    DESIGN.md) rather than silently diverging."""
    env_kw = dict(number_of_pedestrians=12)
    o = OracleEnv(OracleConfig(**env_kw))
    np.random.seed(5)
    o.reset()
    o.positions[:4] = np.array([[0.05, 0.0], [0.0, 0.05], [-0.05, 0.0], [0.02, 0.02]])
    o.statuses = T.compute_statuses(o.positions, o.agent_position, o.exit_position)[0]
    env = _make_env(env_kw, {}, 1)
    u = env.unwrapped
    u.reset()
    u.set_state(positions=o.positions, directions=o.directions, statuses=o.statuses, agent_position=o.agent_position)
    zero = np.zeros(2, dtype=np.float32)
    noise = np.zeros((1, 12), dtype=np.float32)
    for _ in range(3):
        with np.errstate(all="ignore"):
            oobs, r, tm, tr, info = o.step(zero.copy(), noise[0].astype(np.float64))
        obs, reward, term, trunc, _ = env.step(torch.zeros(1, 2), noise=torch.as_tensor(noise))
        st = u.get_state()
        assert np.array_equal(np.isnan(st["positions"][0].cpu().numpy()), np.isnan(o.positions))
        assert np.array_equal(st["statuses"][0].cpu().numpy(), o.statuses)
    assert np.isnan(o.positions).any()


def test_error_behaviour_matches_reference():
    import evacuation_b200 as eb

    with pytest.raises(NotImplementedError):  # wrappers/config.py:80-81
        eb.setup_env(eb.EnvConfig(), eb.EnvWrappersConfig(positions="grav", type="Box"))
    with pytest.raises(ValueError):  # wrappers.py:75
        eb.MatrixObs(eb.EvacuationEnv(eb.EnvConfig()), type="bad")
    with pytest.raises(AssertionError):  # wrappers/config.py:42-44
        eb.EnvWrappersConfig(num_obs_stacks=2)
    with pytest.raises(ValueError):
        eb.EvacuationEnv(eb.EnvConfig(number_of_pedestrians=40000)).reset()
    env = eb.setup_env(eb.EnvConfig, eb.EnvWrappersConfig)  # README.md:72 passes the classes
    obs, _ = env.reset()
    assert set(obs) == {"agent_position", "pedestrians_positions", "exit_position"}
    assert obs["pedestrians_positions"].shape == (10, 2)


def test_logging_and_trajectory_capture(tmp_path):
    """`draw` records the tracked environment's trajectory like Pedestrians.save / Agent.save, reset() logs
    the finished episode with the reference's keys (env.py:115-125), `save_animation` / `render` write the reference's GIF / PNG
    files (rasterised with Pillow, evacuation_b200/render.py)."""
    import evacuation_b200 as eb

    n = 10
    cfg = eb.EnvConfig(number_of_pedestrians=n, draw=True, wandb_enabled=False, path_logs=str(tmp_path / "logs"), path_giff=str(tmp_path / "giff"),
                       path_png=str(tmp_path / "png"), is_new_exiting_reward=True, intrinsic_reward_coef=1.0, experiment_name="t")
    np.random.seed(3)
    env = eb.setup_env(cfg, eb.EnvWrappersConfig())
    u = env.unwrapped
    env.reset()
    oracle = OracleEnv(OracleConfig(number_of_pedestrians=n, is_new_exiting_reward=True, intrinsic_reward_coef=1.0))
    st = u.get_state()
    oracle.set_state(st["positions"][0].cpu().numpy(), st["directions"][0].cpu().numpy(), st["statuses"][0].cpu().numpy(), np.zeros(2, np.float32))
    total = 0.0
    for t in range(6):
        a = np.array([np.sin(0.3 * t), np.cos(0.3 * t)], dtype=np.float32)
        noise = np.random.RandomState(t).uniform(-0.1, 0.1, n).astype(np.float32)
        _, r, _, _, _ = env.step(a, noise=noise)
        _, ro, _, _, _ = oracle.step(a.copy(), noise.astype(np.float64))
        total += ro
        assert abs(r - ro) <= 1e-5 * max(1.0, abs(ro))
    assert len(u.pedestrians.memory["positions"]) == 7 and len(u.pedestrians.memory["statuses"]) == 7 and len(u.agent.memory["position"]) == 6
    np.testing.assert_allclose(u.pedestrians.memory["positions"][-1], oracle.positions, atol=2e-6)
    d = u.episode_log()
    assert list(d) == list(eb._native.EPISODE_STAT_KEYS)
    assert d["episode_length"] == 6 and d["overall_timesteps"] == 6 and abs(d["episode_reward"] - total) <= 1e-4 * abs(total)
    assert d["escaped_pedestrians"] + d["exiting_pedestrians"] + d["following_pedestrians"] + d["viscek_pedestrians"] == n
    from PIL import Image

    gif = u.save_animation()   # env.py:241-324, file name like the reference's: <experiment_name>_ep-<n_episodes>.gif
    assert gif.endswith(os.path.join("giff", "t_ep-1.gif")) and Image.open(gif).n_frames == 6
    png = u.render()           # env.py:173-240: <experiment_name>_<now>.png
    assert png.endswith(os.path.join("png", "t_6.png")) and Image.open(png).size == (500, 500)
    env.reset()  # logs the finished episode like env.py:114-127
    assert u.last_episode_log["episode_length"] == 6 and u.time.n_episodes == 2
    logfile = tmp_path / "logs" / "logs_t.log"
    assert logfile.exists() and "episode_reward=" in logfile.read_text()


def _status_from_positions(pos, agent_pos, thr):
    """statuses.py:29-48 in float64 torch on the device: FOLLOWER < 0.2 of the agent, EXITING < 0.4 / ESCAPED < 0.01 of the exit."""
    p = pos.double()
    da = (p - agent_pos.double()[:, None, :]).norm(dim=-1)
    de = (p - torch.tensor([0.0, -1.0], dtype=torch.float64, device=pos.device)).norm(dim=-1)
    st = torch.ones_like(da, dtype=torch.uint8)
    st[da < thr["leader"]] = 2
    st[de < thr["exit"]] = 3
    st[de < thr["escape"]] = 4
    margin = torch.minimum(torch.minimum((da - thr["leader"]).abs(), (de - thr["exit"]).abs()), (de - thr["escape"]).abs())
    return st, margin


@pytest.mark.parametrize("label,E,n,wrap,steps", [
    ("C2", 4096, 60, dict(positions="rel", statuses="ohe", type="Box"), 300),
    ("C3", 4096, 60, dict(positions="grav", alpha=4), 300),
    ("C4", 256, 4096, dict(positions="rel", statuses="ohe", type="Box"), 25),
    ("C5", 65536, 60, dict(positions="rel", statuses="ohe", type="Box"), 60),
])
def test_full_size_invariants(label, E, n, wrap, steps):
    """BASELINE.json's full sizes through size-independent properties: positions stay in the arena, the
    status is a pure function of (position, agent position), escaped pedestrians are pinned to the exit with a zero
    direction from the following step on, ESCAPED is absorbing, rewards / observations are finite, the observation
    is consistent with the state, and a batch split over two handles (two 'ranks') equals the whole batch bit for bit."""
    import evacuation_b200 as eb

    env_kw = dict(number_of_pedestrians=n, is_new_exiting_reward=True, intrinsic_reward_coef=0.1, enslaving_degree=0.5, noise_coef=0.5)
    mk = lambda num, off: eb.setup_env(eb.EnvConfig(**env_kw), eb.EnvWrappersConfig(**wrap), num_envs=num, seed=11, auto_reset=False,
                                       env_index_offset=off)
    env, lo, hi = mk(E, 0), mk(E // 2, 0), mk(E - E // 2, E // 2)
    for e_ in (env, lo, hi):
        e_.reset()
    u = env.unwrapped
    thr = dict(leader=0.2, exit=0.4, escape=0.01)
    g = torch.Generator(device="cuda").manual_seed(5)
    prev_st = u.get_state()["statuses"].clone()
    for t in range(steps):
        actions = torch.rand((E, 2), device="cuda", generator=g) * 2 - 1
        obs, reward, term, trunc, _ = env.step(actions)
        lo.step(actions[: E // 2].contiguous())
        hi.step(actions[E // 2:].contiguous())
        if t % 10 != 9 and t != steps - 1:
            continue
        st = u.get_state()
        pos, dirs, sts, ap = st["positions"], st["directions"], st["statuses"], st["agent_position"]
        assert torch.isfinite(pos).all() and torch.isfinite(reward).all()
        assert (pos[..., 0].abs() <= 1.0).all() and (pos[..., 1].abs() <= 1.0).all() and (ap.abs() <= 1.0).all()
        want, margin = _status_from_positions(pos, ap, thr)
        clear = margin > 1e-6
        assert torch.equal(sts[clear], want[clear]), f"{label}: status is not f(position, agent) at step {t}"
        was_escaped = prev_st == 4
        assert (sts[was_escaped] == 4).all(), "ESCAPED must be absorbing"
        assert (pos[was_escaped] == torch.tensor([0.0, -1.0], device="cuda")).all() and (dirs[was_escaped] == 0).all()
        assert (dirs.norm(dim=-1) <= 0.01 * (1 + 1e-5)).all(), "|direction| <= step_size"
        assert bool(((sts == 4).sum(dim=1) == n).eq(term).all())
        if wrap["positions"] == "rel":  # observation rows = [agent; exit; pedestrians], relative / sqrt(2), one-hot of 4 - status
            o = obs.reshape(E, n + 2, 6)
            torch.testing.assert_close(o[:, 0, :2], ap)
            torch.testing.assert_close(o[:, 2:, :2], (pos - ap[:, None, :]) / 1.41421353816986083984375, rtol=1e-6, atol=1e-7)
            assert torch.equal(o[:, 2:, 2:].argmax(dim=-1), (4 - sts.long())) and (o[:, 2:, 2:].sum(dim=-1) == 1).all()
        prev_st = sts.clone()
    whole = u.get_state()
    for key in ("positions", "directions", "statuses", "agent_position"):
        parts = torch.cat([lo.unwrapped.get_state()[key], hi.unwrapped.get_state()[key]])
        assert torch.equal(whole[key], parts), f"{label}: sharded batch differs from the whole batch in {key}"


@pytest.mark.parametrize("n,precision", [(60, "fp32"), (60, "fp64"), (300, "fp32")])
def test_checkpoint_resume_is_bit_identical(n, precision):
    """evac_save_state / evac_load_state: the complete device state of a handle as one byte image (both layouts: the packed
    EnvBlock of the one-warp kernel and the array layout of the generic kernels).  Resuming from the image -- in the same
    handle or in a fresh one with the same configuration and seed -- must retrace the original run bit for bit, auto-resets,
    episode statistics and on-device agent state included (the random streams are counter-based)."""
    import evacuation_b200 as eb

    cfgs = (eb.EnvConfig(number_of_pedestrians=n, max_timesteps=23, is_new_exiting_reward=True, intrinsic_reward_coef=0.2),
            eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"))
    mk = lambda: eb.setup_env(*cfgs, num_envs=5, seed=77, auto_reset=True, precision=precision, batched=True)
    a = mk()
    a.reset()
    a.rollout(17, agent="wacuum")
    image = a.unwrapped.save_state()
    assert image.dtype == torch.uint8 and image.is_cuda and image.numel() > 5 * n * 17
    def cont(env):
        obs, r, term, trunc = env.rollout(40, agent="wacuum")
        st = env.unwrapped.get_state()
        acc, overall = env.unwrapped.accumulators()
        stats, fin, tot = env.unwrapped.episode_statistics()
        return [obs.clone(), r.clone(), trunc.clone(), acc, overall, stats, tot] + [st[k] for k in sorted(st)]
    want = cont(a)
    a.unwrapped.load_state(image)          # rewind the same handle
    got_same = cont(a)
    b = mk()                               # a fresh handle (never reset) resumed from the image
    b.unwrapped.load_state(image.cpu())    # (a host copy, e.g. read back from a file)
    got_fresh = cont(b)
    for w, g1, g2 in zip(want, got_same, got_fresh):
        assert torch.equal(w, g1) and torch.equal(w, g2)
    with pytest.raises(ValueError):
        b.unwrapped.load_state(image[:-16])


@pytest.mark.parametrize("E", [1, 3, 300, 1500])
def test_host_face_equals_the_device_face(E):
    """`evac_step_host` (NumPy in / NumPy out) against `evac_step` (device tensors) on the same seeded batch and actions:
    bit-identical observations, rewards and flags on every step.  E <= ~680 takes the zero-copy path (the kernel reads and
    writes the page-locked host block), larger batches the copy engine with the actions read from host memory."""
    import evacuation_b200 as eb

    env_kw = dict(number_of_pedestrians=60, is_new_exiting_reward=True, is_new_followers_reward=True, max_timesteps=12, wandb_enabled=False)
    wrap = dict(positions="rel", statuses="ohe", type="Box")
    dev_env = eb.setup_env(eb.EnvConfig(**env_kw), eb.EnvWrappersConfig(**wrap), num_envs=E, batched=True, auto_reset=True, seed=11)
    host_env = eb.setup_env(eb.EnvConfig(**env_kw), eb.EnvWrappersConfig(**wrap), num_envs=E, batched=False, auto_reset=True, seed=11,
                            rng="philox")
    o_d, _ = dev_env.reset()
    o_h, _ = host_env.reset()
    assert np.array_equal(np.asarray(o_h).reshape(E, -1), o_d.reshape(E, -1).cpu().numpy())
    acts = np.random.RandomState(5).uniform(-1, 1, size=(30, E, 2)).astype(np.float32)
    for t in range(30):  # crosses two truncations (same-step auto-reset) at max_timesteps = 12
        od, rd, td, ud, _ = dev_env.step(torch.as_tensor(acts[t]))
        oh, rh, th, uh, _ = host_env.step(acts[t] if E > 1 else acts[t, 0])
        assert np.array_equal(np.asarray(oh).reshape(E, -1), od.reshape(E, -1).cpu().numpy()), t
        assert np.array_equal(np.asarray(rh, dtype=np.float32).reshape(E), rd.cpu().numpy()), t
        assert np.array_equal(np.asarray(th).reshape(E), td.cpu().numpy().astype(bool)) and np.array_equal(np.asarray(uh).reshape(E), ud.cpu().numpy().astype(bool)), t
    assert np.asarray(uh).reshape(E).any() or t > 0
