"""Environment layer of gym_roboy_b200: the batched, GPU-fused goal-reaching env.

`RoboyEnv` keeps the constructor and attribute surface of the reference env so existing callers
keep working; everything it computes runs in the CUDA kernels behind include/roboy_b200.h.
"""
from . import robots, simulations
from .roboy_env import RoboyEnv

__all__ = ["RoboyEnv", "robots", "simulations"]
