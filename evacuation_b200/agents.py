"""Scripted agents with the reference's interface (src/agents/base_agent.py,
random_agent.py:8-9, rotating_agent.py:12-16, baseline_wacuum_cleaner.py:7-80).  The batched env also
offers Random / Rotating ON DEVICE through `EvacuationEnv.rollout(agent="random"|"rotating")`, and
`WacuumCleaner.batched(env)` is a vectorised torch state machine for the sweep baseline."""
import numpy as np

from .statuses import SwitchDistances


class BaseAgent:
    def __init__(self, action_space):
        self.action_space = action_space

    def act(self, obs):
        raise NotImplementedError()

    def update(self, **kwargs):
        raise NotImplementedError


class RandomAgent(BaseAgent):
    def act(self, obs):
        return self.action_space.sample()


class RotatingAgent(BaseAgent):
    def __init__(self, action_space, parameter=0.05):
        super().__init__(action_space)
        self.i = 0
        self.parameter = parameter

    def act(self, obs):
        self.i += 1
        step = self.i * self.parameter
        return [np.sin(step), np.cos(step)]


class WacuumCleaner(BaseAgent):
    """Scripted sweep baseline (src/agents/baseline_wacuum_cleaner.py:7-80): climb to the top wall, sweep the
    arena in horizontal lanes (25 steps down between lanes), then walk to the exit.  `act(obs)` reads
    obs['agent_position'] like the reference ([2] NumPy for the single-env face).

    `WacuumCleaner.batched(env)` returns the same policy as a vectorised torch state machine over all
    environments of a batched env (positions [E,2] on the env's device), so scripted rollouts stay on the GPU."""

    LANE_STEPS = 25

    def __init__(self, env):
        u = env.unwrapped
        half_reach = SwitchDistances.to_leader / 2  # SWITCH_DISTANCE_TO_LEADER / 2
        self.step_size = u.area.step_size
        self.exit_position = np.asarray(u.area.exit.position, dtype=np.float32)
        self.top = u.area.height - half_reach + self.step_size
        self.right_edge = u.area.width - half_reach + self.step_size
        self.left_edge = -u.area.width + half_reach - self.step_size
        self.bottom = -u.area.height + half_reach - self.step_size
        self.phase = 0          # 0: climb, 1: sweep, 2: go to the exit   (task_done of the reference)
        self.heading = 1.0      # +1: sweeping right, -1: sweeping left   (task_1_direction)
        self.down_left = 0      # steps still to go down                   (task_1_time_to_go_down)

    def _sweep(self, pos):
        if self.down_left > 0:
            self.down_left -= 1
            if pos[1] > self.bottom:
                return np.array([0.0, -1.0], dtype=np.float32)
            self.phase = 2
            return self.exit_position - pos
        lane_open = pos[0] < self.right_edge if self.heading > 0 else pos[0] > self.left_edge
        if lane_open:
            return np.array([self.heading, 0.0], dtype=np.float32)
        self.heading = -self.heading
        self.down_left = self.LANE_STEPS
        return np.array([0.0, -1.0], dtype=np.float32)

    def act(self, obs):
        pos = np.asarray(obs["agent_position"], dtype=np.float32).reshape(2)
        if self.phase == 0:
            if pos[1] < self.top:
                return np.array([0.0, 1.0], dtype=np.float32)
            self.phase = 1
            return np.array([1.0, 0.0], dtype=np.float32)
        if self.phase == 1:
            return self._sweep(pos)
        return self.exit_position - pos

    @classmethod
    def batched(cls, env):
        return _WacuumCleanerBatched(cls(env), env.unwrapped.num_envs, env.unwrapped.device)


class _WacuumCleanerBatched:
    """The WacuumCleaner state machine for E environments at once (torch, on device).  `act(agent_position[E,2])`;
    `reset(mask)` re-arms the environments that started a new episode."""

    def __init__(self, proto: WacuumCleaner, num_envs: int, device):
        import torch

        self.t, self.p = torch, proto
        self.phase = torch.zeros(num_envs, dtype=torch.int32, device=device)
        self.heading = torch.ones(num_envs, dtype=torch.float32, device=device)
        self.down_left = torch.zeros(num_envs, dtype=torch.int32, device=device)
        self.exit_position = torch.as_tensor(proto.exit_position, device=device)

    def reset(self, mask=None):
        if mask is None:
            self.phase.zero_(); self.heading.fill_(1.0); self.down_left.zero_()
        else:
            self.phase.masked_fill_(mask, 0); self.heading.masked_fill_(mask, 1.0); self.down_left.masked_fill_(mask, 0)

    def act(self, agent_position):
        t, p = self.t, self.p
        pos = agent_position.reshape(-1, 2)
        x, y = pos[:, 0], pos[:, 1]
        zeros, ones = t.zeros_like(x), t.ones_like(x)
        to_exit = self.exit_position - pos
        # phase 0: climb
        climbing = self.phase == 0
        reached_top = climbing & ~(y < p.top)
        act = t.stack([zeros, ones], dim=1)                                              # up
        act = t.where(reached_top[:, None], t.stack([ones, zeros], dim=1), act)          # first step to the right
        # phase 1: sweep
        sweeping = self.phase == 1
        going_down = sweeping & (self.down_left > 0)
        hit_bottom = going_down & ~(y > p.bottom)
        lane_open = t.where(self.heading > 0, x < p.right_edge, x > p.left_edge)
        lane = sweeping & ~going_down & lane_open
        turn = sweeping & ~going_down & ~lane_open
        down = t.stack([zeros, -ones], dim=1)
        act = t.where((going_down & ~hit_bottom)[:, None] | turn[:, None], down, act)
        act = t.where(lane[:, None], t.stack([self.heading, zeros], dim=1), act)
        act = t.where((hit_bottom | (self.phase == 2))[:, None], to_exit, act)
        # state updates (after the decisions, like the reference's in-place mutations)
        self.down_left = t.where(going_down, self.down_left - 1, self.down_left)
        self.down_left = t.where(turn, t.full_like(self.down_left, p.LANE_STEPS), self.down_left)
        self.heading = t.where(turn, -self.heading, self.heading)
        self.phase = t.where(reached_top, t.ones_like(self.phase), self.phase)
        self.phase = t.where(hit_bottom, t.full_like(self.phase, 2), self.phase)
        return act
