"""CPU oracle for the evacuation environment's per-step pedestrian dynamics.

TEST INFRASTRUCTURE ONLY -- not part of the product path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
this module, and there only as the checker / reported CPU baseline.  The shipped path
(`evacuation_b200`) never imports it and fails loudly when the CUDA library is missing.

What it is: a from-scratch float64 NumPy restatement of the reference algorithm
(cinemere/evacuation, paths relative to /root/reference):

    src/env/env/area.py:76-210      Area.pedestrians_step / agent_step / _if_wall_collision
    src/env/env/statuses.py:29-48   update_statuses
    src/env/env/distances.py:24-56  is_distance_low / sum_distance
    src/env/env/reward.py:19-46     Reward
    src/env/env/pedestrians.py:16-31 Pedestrians.reset
    src/env/env/env.py:98-171       EvacuationEnv reset / step / _get_observation
    src/env/wrappers/wrappers.py:8-96, gravity_encoding.py:8-81  observation wrappers
    src/env/constants.py:35-38      switch distances

Third-party arithmetic the reference leans on and that is NOT under /root/reference:
`scipy.spatial.distance_matrix` (requirements.txt:4, unpinned; image has scipy 1.18.1) whose
published algorithm is  d = (sum_k |y_k - x_k| ** 2) ** 0.5  in float64, restated in
`_pairwise_distance` below; NumPy ufuncs / pairwise summation / legacy MT19937 `np.random`
(requirements.txt:2, unpinned; image has numpy 2.3.5) which the oracle calls directly.

Parity pin: the reference ships no tests or golden vectors, so the pin
is made here: `tests/golden/*.npz` hold trajectories produced by the UNMODIFIED reference run
in the authoring container through `oracle/ref_shim.py` (generator:
`tests/golden/gen_golden.py`), and `tests/test_oracle_golden.py` asserts this oracle is
BIT-IDENTICAL to them (statuses, positions, directions, rewards, observations), including the
reference's consumption of the global NumPy random stream.

Statuses are stored as uint8 with the reference's enum values
(VISCEK=1, FOLLOWER=2, EXITING=3, ESCAPED=4; statuses.py:16-27).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

VISCEK, FOLLOWER, EXITING, ESCAPED = 1, 2, 3, 4
NUM_STATUSES = 4

# src/env/constants.py:35-38
SWITCH_DISTANCE_TO_LEADER = 0.2
SWITCH_DISTANCE_TO_OTHER_PEDESTRIAN = 0.1
SWITCH_DISTANCE_TO_EXIT = 0.4
SWITCH_DISTANCE_TO_ESCAPE = 0.01


@dataclass
class OracleConfig:
    """Numerical fields of the reference's EnvConfig (config.py:3-100) and
    EnvWrappersConfig (wrappers/config.py:8-44), same names and defaults."""

    number_of_pedestrians: int = 10
    width: float = 1.0
    height: float = 1.0
    step_size: float = 0.01
    noise_coef: float = 0.2
    eps: float = 1e-8
    enslaving_degree: float = 1.0
    is_new_exiting_reward: bool = False
    is_new_followers_reward: bool = True
    intrinsic_reward_coef: float = 0.0
    is_termination_agent_wall_collision: bool = False
    init_reward_each_step: float = -1.0
    max_timesteps: int = 2000
    # wrappers
    positions: str = "abs"  # abs | rel | grav
    statuses: str = "no"  # no | ohe | cat
    type: str = "Dict"  # Dict | Box
    alpha: float = 3


def _pairwise_distance(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """scipy.spatial.distance_matrix(x, y, 2) restated: [m,2] x [n,2] -> [m,n] float64.

    scipy: minkowski_distance_p = np.sum(np.abs(y-x)**p, axis=-1); then ** (1/p).
    With p = 2 NumPy evaluates **2 as a square and **0.5 as a square root; the sum over the
    two coordinates is a single addition, so the expression below is bit-identical.
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    dx = y[np.newaxis, :, 0] - x[:, np.newaxis, 0]
    dy = y[np.newaxis, :, 1] - x[:, np.newaxis, 1]
    return np.sqrt(dx * dx + dy * dy)


def _distance_to_point(positions: np.ndarray, point: np.ndarray) -> np.ndarray:
    """Column of `_pairwise_distance(positions, point[None])` -> [n] float64."""
    p = np.asarray(point, dtype=np.float64)
    dx = p[0] - positions[:, 0]
    dy = p[1] - positions[:, 1]
    return np.sqrt(dx * dx + dy * dy)


def compute_statuses(positions, agent_position, exit_position, margin_out: Optional[list] = None):
    """statuses.py:29-48.  Every element is overwritten, so the status is a pure function of
    (pedestrian position, agent position): FOLLOWER if d(agent) < 0.2, overridden by EXITING
    if d(exit) < 0.4, overridden by ESCAPED if d(exit) < 0.01, else VISCEK."""
    d_agent = _distance_to_point(positions, agent_position)
    d_exit = _distance_to_point(positions, exit_position)
    st = np.full(positions.shape[0], VISCEK, dtype=np.uint8)
    st[d_agent < SWITCH_DISTANCE_TO_LEADER] = FOLLOWER
    st[d_exit < SWITCH_DISTANCE_TO_EXIT] = EXITING
    st[d_exit < SWITCH_DISTANCE_TO_ESCAPE] = ESCAPED
    if margin_out is not None and positions.shape[0] > 0:
        with np.errstate(invalid="ignore"):
            margin_out.append(np.nanmin(np.abs(d_agent - SWITCH_DISTANCE_TO_LEADER)))
            margin_out.append(np.nanmin(np.abs(d_exit - SWITCH_DISTANCE_TO_EXIT)))
            margin_out.append(np.nanmin(np.abs(d_exit - SWITCH_DISTANCE_TO_ESCAPE)))
    return st, d_exit


@dataclass
class StepInfo:
    """Side information the parity tests need (not part of the reference API)."""

    reward_agent: float = 0.0
    reward_pedestrians: float = 0.0
    intrinsic_reward: float = 0.0
    terminated_agent: bool = False
    terminated_pedestrians: bool = False
    # smallest distance of any thresholded quantity to its threshold during this step
    # (pair test vs 0.1, status tests vs 0.2 / 0.4 / 0.01, wall tests): a float32
    # implementation may legitimately decide differently when this is ~1e-7.
    margin: float = np.inf
    # smallest mean resultant length |mean of neighbour unit vectors| over the VISCEK/FOLLOWER pedestrians:
    # the new heading is atan2 of that mean, so a float32 heading error ~1e-7 / min_resultant is expected
    min_resultant: float = np.inf
    n_noise_used: int = 0
    fv_mask: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=bool))
    pair_margin_rows: np.ndarray = field(default_factory=lambda: np.zeros(0))


class OracleEnv:
    """One environment, reference semantics, float64 pedestrian state, float32 agent state."""

    def __init__(self, cfg: OracleConfig):
        self.cfg = cfg
        self.n = int(cfg.number_of_pedestrians)
        self.exit_position = np.array([0, -1], dtype=np.float32)  # area.py:36-39
        self.now = 0
        self.n_episodes = 0
        self.overall_timesteps = 0
        self.episode_reward = 0.0
        self.episode_intrinsic_reward = 0.0
        self.episode_status_reward = 0.0
        self.positions = np.zeros((self.n, 2))
        self.directions = np.zeros((self.n, 2))
        self.statuses = np.full(self.n, VISCEK, dtype=np.uint8)
        self.agent_position = np.zeros(2, dtype=np.float32)
        self.agent_direction = np.zeros(2, dtype=np.float32)
        if cfg.positions == "grav" and cfg.type == "Box":
            raise NotImplementedError  # wrappers/config.py:80-81
        # RelativePosition.__init__ (wrappers.py:12-18): sqrt(low**2 + high**2) in float32
        self._hyp = np.sqrt(np.float32(-1) ** 2 + np.float32(1) ** 2).astype(np.float32)
        # rows of the |fv| x |efv| distance matrix evaluated at once (None = all, like the reference); see _pedestrians_step
        self.row_chunk = None
        self.chunk_threads = 0  # > 0: row blocks evaluated by that many threads (same numbers; only for the big parity cases)

    # ------------------------------------------------------------------ state
    def get_state(self) -> dict:
        return dict(
            positions=self.positions.copy(),
            directions=self.directions.copy(),
            statuses=self.statuses.copy(),
            agent_position=self.agent_position.copy(),
            agent_direction=self.agent_direction.copy(),
            now=self.now,
        )

    def set_state(self, positions, directions, statuses, agent_position, agent_direction=None, now=0):
        self.positions = np.array(positions, dtype=np.float64).reshape(self.n, 2)
        self.directions = np.array(directions, dtype=np.float64).reshape(self.n, 2)
        self.statuses = np.array(statuses, dtype=np.uint8).reshape(self.n)
        self.agent_position = np.array(agent_position, dtype=np.float32).reshape(2)
        if agent_direction is None:
            agent_direction = np.zeros(2, dtype=np.float32)
        self.agent_direction = np.array(agent_direction, dtype=np.float32).reshape(2)
        self.now = int(now)

    # ------------------------------------------------------------------ reset
    def reset(self, rng=None):
        """env.py:106-139 -> Time.reset, Agent.reset (area.py:27-30), Pedestrians.reset
        (pedestrians.py:16-27).  `rng` defaults to the GLOBAL numpy stream like the reference
        (positions (N,2) first, then directions (N,2))."""
        rng = np.random if rng is None else rng
        self.episode_reward = 0.0
        self.episode_intrinsic_reward = 0.0
        self.episode_status_reward = 0.0
        self.now = 0
        self.n_episodes += 1
        self.agent_position = np.zeros(2, dtype=np.float32)
        self.agent_direction = np.zeros(2, dtype=np.float32)
        self.positions = rng.uniform(-1.0, 1.0, size=(self.n, 2))
        d = rng.uniform(-1.0, 1.0, size=(self.n, 2))
        self.directions = (d.T / np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])).T
        self.statuses, _ = compute_statuses(self.positions, self.agent_position, self.exit_position)
        return self.observation()

    # ------------------------------------------------------------------ step
    def _agent_step(self, action, info: StepInfo, margins: list):
        """area.py:182-210, float32 arithmetic when the action is float32 (NumPy-2 weak
        scalars keep `+ eps` and `step_size *` in float32)."""
        cfg = self.cfg
        action = np.array(action)
        action /= np.linalg.norm(action) + cfg.eps
        self.agent_direction = cfg.step_size * action
        pt = self.agent_position + self.agent_direction
        margins.append(float(min(abs(pt[0] + cfg.width), abs(pt[0] - cfg.width),
                                 abs(pt[1] + cfg.height), abs(pt[1] - cfg.height))))
        collide = (pt[0] < -cfg.width) or (pt[0] > cfg.width) or (pt[1] < -cfg.height) or (pt[1] > cfg.height)
        if not collide:
            self.agent_position += self.agent_direction
            info.terminated_agent, info.reward_agent = False, 0.0
        else:
            info.terminated_agent = bool(cfg.is_termination_agent_wall_collision)
            info.reward_agent = -5.0

    def _pedestrians_step(self, noise, info: StepInfo, margins: list):
        """area.py:76-180."""
        cfg = self.cfg
        pos, dirs, st = self.positions, self.directions, self.statuses
        exit_pos = self.exit_position

        escaped = st == ESCAPED
        dirs[escaped] = 0
        pos[escaped] = exit_pos

        exiting = st == EXITING
        if exiting.any():
            v = exit_pos - pos[exiting]
            length = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1])
            size = np.minimum(length, cfg.step_size)
            dirs[exiting] = (v.T / length * size).T

        following = st == FOLLOWER
        viscek = st == VISCEK
        efv = exiting | following | viscek
        fv = following | viscek
        info.fv_mask = fv.copy()

        e_dirs = dirs[efv]
        with np.errstate(invalid="ignore", divide="ignore"):
            e_unit = (e_dirs.T / np.sqrt(e_dirs[:, 0] * e_dirs[:, 0] + e_dirs[:, 1] * e_dirs[:, 1])).T

        # area.py:105-119.  One broadcast expression like the reference for ordinary sizes; for crowds whose |fv| x |efv|
        # float64 matrix would not fit in memory (the > 8192-pedestrian parity cases) the SAME expressions are evaluated on
        # blocks of rows -- every row's sums (a reduction along the contiguous axis) are the same numbers either way
        # (tests/test_oracle_golden.py::test_row_chunked_alignment_is_bit_identical).
        pos_fv, pos_efv = pos[fv], pos[efv]
        chunk = self.row_chunk if self.row_chunk else max(1, pos_fv.shape[0])

        def rows(r0):
            dm = _pairwise_distance(pos_fv[r0:r0 + chunk], pos_efv)
            gap = None
            if dm.size:
                with np.errstate(invalid="ignore"):
                    gap = np.abs(dm - SWITCH_DISTANCE_TO_OTHER_PEDESTRIAN).min(axis=1)  # NaN rows stay NaN
            inter = np.where(dm < SWITCH_DISTANCE_TO_OTHER_PEDESTRIAN, 1, 0)
            n_inter = np.maximum(1, inter.sum(axis=1))
            with np.errstate(invalid="ignore"):
                return ((inter * e_unit[:, 0]).sum(axis=1) / n_inter, (inter * e_unit[:, 1]).sum(axis=1) / n_inter, gap)

        starts = list(range(0, pos_fv.shape[0], chunk)) if pos_fv.shape[0] else [0]
        if self.chunk_threads and len(starts) > 1:  # NumPy releases the GIL inside its loops; the row blocks are independent
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(self.chunk_threads) as pool:
                parts = list(pool.map(rows, starts))
        else:
            parts = [rows(r0) for r0 in starts]
        mx, my = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
        gaps = [p[2] for p in parts if p[2] is not None]
        if gaps:
            row_gap = np.concatenate(gaps)
            info.pair_margin_rows = row_gap  # per VISCEK/FOLLOWER pedestrian: distance of its closest pair to the vision radius
            with np.errstate(invalid="ignore"):
                if np.isfinite(row_gap).any():
                    margins.append(float(np.nanmin(row_gap)))
        theta = np.arctan2(my, mx)
        if mx.size:
            with np.errstate(invalid="ignore"):
                res = np.sqrt(mx * mx + my * my)
                info.min_resultant = float(np.nanmin(res)) if np.isfinite(res).any() else np.inf

        n_fv = int(fv.sum())
        info.n_noise_used = n_fv
        if noise is None:
            nz = np.random.uniform(low=-cfg.noise_coef / 2, high=cfg.noise_coef / 2, size=n_fv)
        else:
            nz = np.asarray(noise, dtype=np.float64).reshape(self.n)[fv]
        theta = theta + nz
        dirs[fv] = np.vstack((np.cos(theta), np.sin(theta))).T * cfg.step_size

        # leader enslaving (area.py:138-142); e * agent.direction stays float32
        f_dirs = dirs[following]
        dirs[following] = cfg.enslaving_degree * self.agent_direction + (1.0 - cfg.enslaving_degree) * f_dirs

        pos[efv] += dirs[efv]

        # wall reflection (area.py:147-152)
        clipped = np.clip(pos, [-cfg.width, -cfg.height], [cfg.width, cfg.height])
        miss = pos - clipped
        # wall margin: only pedestrians whose direction sign matters afterwards (an EXITING pedestrian's
        # direction is rebuilt from its position next step, and the reflection of the position is continuous)
        moved = pos[fv]
        if moved.size:
            with np.errstate(invalid="ignore"):
                margins.append(float(np.nanmin(np.abs(np.abs(moved) - np.array([cfg.width, cfg.height])))))
        pos -= 2 * miss
        dirs *= np.where(miss != 0, -1, 1)

        old = st.copy()
        new, d_exit = compute_statuses(pos, self.agent_position, exit_pos, margins)

        # Reward.estimate_status_reward (reward.py:23-46); note code (15+10tf / 10+5tf), not docstring
        reward = cfg.init_reward_each_step
        tf = 1 - self.now / (200 * self.n)
        if cfg.is_new_exiting_reward:
            k = int((((old == VISCEK) | (old == FOLLOWER)) & (new == EXITING)).sum())
            reward += (15 + 10 * tf) * k
        if cfg.is_new_followers_reward:
            k = int(((old == VISCEK) & (new == FOLLOWER)).sum())
            reward += (10 + 5 * tf) * k
        info.reward_pedestrians = reward
        # Reward.estimate_intrinsic_reward (reward.py:19-21) + sum_distance (distances.py:51-56)
        info.intrinsic_reward = 0 - d_exit.sum() / self.n
        self.statuses = new
        info.terminated_pedestrians = bool((new == ESCAPED).sum() == self.n)

    def step(self, action, noise=None):
        """env.py:141-171.  `noise`: None -> draw |fv| values from the global numpy stream
        exactly like area.py:124; or a dense [N] array whose entries at the VISCEK/FOLLOWER
        slots are used (the injection protocol of SURVEY.md 8c)."""
        cfg = self.cfg
        info = StepInfo()
        margins: list = []
        self.now += 1
        self.overall_timesteps += 1
        truncated = self.now >= cfg.max_timesteps
        self._agent_step(action, info, margins)
        self._pedestrians_step(noise, info, margins)
        reward = info.reward_agent + info.reward_pedestrians + cfg.intrinsic_reward_coef * info.intrinsic_reward
        self.episode_reward += reward
        self.episode_intrinsic_reward += info.intrinsic_reward
        self.episode_status_reward += info.reward_agent + info.reward_pedestrians
        m = [x for x in margins if x == x]
        info.margin = min(m) if m else np.inf
        terminated = info.terminated_agent or info.terminated_pedestrians
        return self.observation(), reward, terminated, truncated, info

    # ------------------------------------------------------------------ observations
    def status_stats(self):
        s = self.statuses
        return dict(escaped=int((s == ESCAPED).sum()), exiting=int((s == EXITING).sum()),
                    following=int((s == FOLLOWER).sum()), viscek=int((s == VISCEK).sum()))

    def _status_encoding(self):
        """PedestriansStatuses.observation (wrappers.py:47-57): s = 4 - status.value."""
        s = NUM_STATUSES - self.statuses.astype(np.int64)
        if self.cfg.statuses == "ohe":
            enc = np.zeros((self.n, NUM_STATUSES))
            enc[np.arange(self.n), s] = 1
            return enc
        if self.cfg.statuses == "cat":
            return s / NUM_STATUSES
        return None

    def observation(self):
        cfg = self.cfg
        agent = self.agent_position
        if cfg.positions == "grav":
            return self._gravity_observation()
        ped = self.positions
        ext = self.exit_position
        if cfg.positions == "rel":  # wrappers.py:20-27
            ped = (ped - agent) / self._hyp
            ext = (ext - agent) / self._hyp
        if cfg.type == "Box":  # MatrixObs, wrappers.py:77-96
            pos = np.vstack((agent, ext, ped))
            if cfg.statuses == "ohe":
                stat = np.vstack((np.array([0, 0, 0, 0], dtype=np.float32),
                                  np.array([1, 0, 0, 0], dtype=np.float32), self._status_encoding()))
                return np.hstack((pos, stat)).astype(np.float32)
            if cfg.statuses == "cat":
                stat = np.hstack(([0, 1], self._status_encoding()))
                return np.hstack((pos, stat[:, np.newaxis])).astype(np.float32)
            return pos
        obs = {"agent_position": agent, "pedestrians_positions": ped, "exit_position": ext}
        if cfg.statuses != "no":
            obs["pedestrians_statuses"] = self._status_encoding()
        return obs

    def _gravity_observation(self):
        """GravityEncoding.observation (gravity_encoding.py:59-81)."""
        cfg = self.cfg
        alpha, eps = cfg.alpha, cfg.eps
        agent = self.agent_position
        viscek = self.statuses == VISCEK
        n_followers = np.int64((self.statuses == FOLLOWER).sum())
        R = agent[np.newaxis, :] - self.positions[viscek, :]
        if len(R) != 0:
            with np.errstate(all="ignore"):
                norm = np.sqrt(R[:, 0] * R[:, 0] + R[:, 1] * R[:, 1])[:, np.newaxis] + eps
                g_ped = (-alpha / norm ** (alpha + 2) * R).sum(axis=0)
        else:
            g_ped = np.zeros(2)
        Re = agent - self.exit_position
        with np.errstate(all="ignore"):
            norm_e = np.linalg.norm(Re) + eps
            g_exit = (-alpha / norm_e ** (alpha + 2) * Re) * n_followers
        return {"agent_position": agent, "grad_potential_pedestrians": g_ped, "grad_potential_exit": g_exit}


def flatten_observation(obs) -> np.ndarray:
    """Flatten any observation to one float64 vector (dict keys in sorted order, the order
    gymnasium's FlattenObservation uses) -- the layout of the CUDA path's flat obs rows is
    compared against this in the parity tests."""
    if isinstance(obs, dict):
        return np.concatenate([np.asarray(obs[k], dtype=np.float64).ravel() for k in sorted(obs)])
    return np.asarray(obs, dtype=np.float64).ravel()
