"""Scripted agents with the reference's interface (src/agents/base_agent.py,
random_agent.py:8-9, rotating_agent.py:12-16).  The batched env also offers the same two
policies ON DEVICE through `EvacuationEnv.rollout(agent="random"|"rotating")`."""
import numpy as np


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
