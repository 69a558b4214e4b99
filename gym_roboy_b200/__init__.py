"""gym_roboy_b200 -- B200-native batched implementation of gym-roboy's environment hot path.

`RoboyEnv.step / reset / compute_reward` for the MSJ robot, fused into hand-written sm_100a
CUDA kernels behind the reference's `SimulationClient` / `RoboyRobot` plug-in API.

    import gym_roboy_b200
    env = gym_roboy_b200.make("msj-control-v1", num_envs=1 << 20)
    obs = env.reset()
    obs, reward, done, info = env.step(actions)      # CUDA tensors, zero-copy
"""
from .registration import make, register, spec  # noqa: F401

__version__ = "0.1.0"

# README.md:24 of the reference documents this id; gym_roboy/__init__.py:3-6 never registered it.
register(id="msj-control-v1", entry_point="gym_roboy_b200.envs:RoboyEnv")
