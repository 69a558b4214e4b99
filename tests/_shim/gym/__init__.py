"""TEST-ONLY minimal stand-in for OpenAI gym (old API) -- see tests/_shim/README.md."""
from . import spaces  # noqa: F401
from .envs import registration as _registration
from .envs.registration import make, register  # noqa: F401

__version__ = "0.0-shim"


class Env:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self


class GoalEnv(Env):
    def compute_reward(self, achieved_goal, desired_goal, info):
        raise NotImplementedError
