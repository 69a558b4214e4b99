"""Small, ragged-size exercise of every kernel for compute-sanitizer (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector, gae
rng = np.random.default_rng(0)
for n in (1, 31, 33, 257, 1000):
    for penalty in (False, True):
        c = CudaSimulationClient(num_envs=n, seed=3, device="cuda:0")
        e = RoboyEnv(c, joint_vel_penalty=penalty, strict=False, auto_reset=True)
        if n == 1:
            e._single = False
        c.enable_terminal_obs(True)
        e.reset(); e.reset(mask=torch.ones(n, dtype=torch.uint8))
        c.set_step_num(np.full(n, 399, np.int32))
        c.set_goal(np.zeros((n, 3), np.float32)); c.set_state(np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.ones(n, np.uint8))
        for t in range(3):
            a = rng.uniform(-1, 1, (n, 8)).astype(np.float32); a[::3] = 0
            e.step(torch.as_tensor(a, device="cuda:0"))
        obs, rew, done = np.empty((n, 9), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8)
        c.step_host(rng.uniform(-1, 1, (n, 8)).astype(np.float32), obs, rew, done)
        c.read_state(); c.forward_step_command(torch.zeros((n, 8))); c.forward_reset_command(); c.get_new_goal_joint_angles()
        q = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
        e.step_from_states(q, q * 0.1, np.ones(n, np.uint8)); e.reset_from_states(q, q * 0.1)
        c.compute_reward(q, q * 0.1, np.ones(n, np.uint8), q, None)
        gae(torch.zeros((4, n), device="cuda:0"), torch.zeros((4, n), device="cuda:0"), torch.zeros((4, n), dtype=torch.uint8, device="cuda:0"), torch.zeros(n, device="cuda:0"))
        torch.cuda.synchronize(); c.stats(); c.errors(); c.close()
# the fused policy rollout kernels (float32 with 1 and 2 envs per thread; tensor cores), ragged sizes
for n in (2, 33, 257, 1000):
    for mode, ept in (("fp32", 1), ("fp32", 2), ("tc", 1), ("tc", 2), ("tc_exact", 0)):
        c = CudaSimulationClient(num_envs=n, seed=3, device="cuda:0")
        col = RolloutCollector(RoboyEnv(c), MlpPolicy().to("cuda:0"), n_steps=3, fused=mode, envs_per_thread=ept)
        c.set_step_num(np.full(n, 399, np.int32))
        col.noise = torch.zeros((3, n, 8), device="cuda:0")
        col.collect(); col.collect()
        torch.cuda.synchronize(); c.errors(); c.close()
print("sanitize run ok")
