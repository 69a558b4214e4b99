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
# robots other than MSJ (the generic kernels: every path of the fused step, injection, external feed, un-fused calls, host
# buffers, done-index list) and MSJ-shaped robots with other limits (the range-checked instantiation of the tuned kernels)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cuda_adaptor import robot_from_bounds
from oracle import oracle as orc
ROBOTS = [dict(dim_joint=6, dim_action=14, angle_low=-2.5, angle_high=2.5, vel_low=-0.6, vel_high=0.6, act_low=-0.25, act_high=0.25),
          dict(dim_joint=15, dim_action=64, angle_low=-np.linspace(1.0, 3.0, 15), angle_high=np.linspace(0.5, 3.1, 15), vel_low=-0.5,
               vel_high=0.5, act_low=-0.5, act_high=0.5),
          dict(dim_joint=1, dim_action=3, angle_low=0.0, angle_high=2.0, vel_low=0.0, vel_high=0.5, act_low=0.0, act_high=0.4),
          dict(dim_joint=11, dim_action=17, angle_low=-1.0, angle_high=2.0, vel_low=-0.5, vel_high=0.25, act_low=-0.125, act_high=0.125),
          dict(angle_low=0.0, angle_high=2.5, vel_low=-0.3, vel_high=0.9, act_low=0.0, act_high=0.4)]   # MSJ-shaped, other limits
for b in ROBOTS:
    J, A, _, bb = orc.robot_bounds(b)
    zero_action, _ = orc.hold_action(b)
    for n in (1, 31, 33, 257, 1000):
        for penalty in (False, True):
            c = CudaSimulationClient(robot=robot_from_bounds(b), num_envs=n, seed=3, device="cuda:0")
            e = RoboyEnv(c, joint_vel_penalty=penalty, strict=False, auto_reset=True)
            if n == 1:
                e._single = False
            c.enable_terminal_obs(True); c.enable_done_index(True)
            e.reset(); e.reset(mask=torch.ones(n, dtype=torch.uint8))
            c.set_step_num(np.full(n, 399, np.int32))
            mid = ((np.broadcast_to(bb["angle_low"], (J,)) + np.broadcast_to(bb["angle_high"], (J,))) / 2).astype(np.float32)
            c.set_goal(np.tile(mid, (n, 1)))
            c.set_state(np.tile(mid, (n, 1)), np.zeros((n, J), np.float32), np.ones(n, np.uint8))
            for t in range(3):
                a = rng.uniform(-1, 1, (n, A)).astype(np.float32); a[::3] = zero_action; a[1::7, 0] = np.nan
                e.step(torch.as_tensor(a, device="cuda:0"))
                c.done_indices(with_terminal_obs=True)
            acts = rng.uniform(-1, 1, (2, n, A)).astype(np.float32)
            c.step_many(torch.as_tensor(acts, device="cuda:0"))
            obs, rew, done = np.empty((n, 3 * J), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8)
            c.step_host(rng.uniform(-1, 1, (n, A)).astype(np.float32), obs, rew, done)
            c.read_state(); c.forward_step_command(torch.zeros((n, A))); c.forward_reset_command(); c.get_new_goal_joint_angles()
            q = rng.uniform(-3, 3, (n, J)).astype(np.float32)
            e.step_from_states(q, q * 0.1, np.ones(n, np.uint8)); e.reset_from_states(q, q * 0.1)
            c.compute_reward(q, q * 0.1, np.ones(n, np.uint8), q, None)
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
