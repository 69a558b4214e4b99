"""INTEGRATION.md level 1, for real: the REFERENCE's own `RoboyEnv` (unmodified, from /root/reference or the copy
under baseline/_ref that travels to the GPU box) driven over `CudaSimulationClient(num_envs=1)` through the
four-call `SimulationClient` plug-in API (simulation_client.py:6-23), replaying the golden fixture that was recorded
from the reference over its own `StubSimulationClient` -- every observation, done flag, goal and step counter
bit-exact, rewards to 1e-6.  The arithmetic here is the reference's numpy code; the CUDA side supplies the states."""
import contextlib
import io

import numpy as np
import pytest

import golden_replay as gr
from oracle import reference_harness as rh

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.available(), reason="reference (baseline/_ref) not present")]


def test_reference_roboy_env_over_cuda_client_replays_the_golden_fixture():
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    RefRoboyEnv, RefMsjRobot, _, _ = rh._import_reference()
    fx = gr.load("rollout_manual_reset")
    N, T, seed = fx["N"], fx["T"], fx["seed"]
    envs, clients = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for e in range(N):
            # the reference's own robot plug-in: its typeguard-checked helpers want gym's Box, which MsjRobot carries
            client = CudaSimulationClient(robot=RefMsjRobot(), num_envs=1, seed=seed, env_id_base=e, device="cuda:0")
            env = RefRoboyEnv(client, joint_vel_penalty=fx["joint_vel_penalty"],
                              is_agent_getting_bonus_for_reaching_goal=fx["bonus"])          # roboy_env.py:12-38
            assert type(env).__module__.startswith("gym_roboy.")
            envs.append(env); clients.append(client)
    goals = lambda: np.stack([np.asarray(e._goal_state.joint_angles, np.float32) for e in envs])  # noqa: E731
    assert np.array_equal(goals(), fx["init_goal"])
    assert envs[0].reward_range == pytest.approx(tuple(fx["reward_range"]), rel=1e-12)

    # the fixture's call counter is global to the vec env (it advances for every env on every step / reset call)
    t_global = 0

    def sync_counter():
        for c in clients:
            c.counter = t_global

    with contextlib.redirect_stdout(io.StringIO()):
        sync_counter()
        obs0 = np.stack([e.reset() for e in envs])                                           # roboy_env.py:82-87
        t_global += 1
        assert obs0.dtype == np.float64 and np.array_equal(obs0.astype(np.float32), fx["reset_obs"])
        by_step = {}
        for row in fx["events"]:
            by_step.setdefault(int(row[0]), []).append(row)
        worst = 0.0
        for t in range(T):
            for row in by_step.get(t, []):
                e, kind, p = int(row[1]), int(row[2]), row[3:].astype(np.float32)
                if kind == gr.EV_SET_GOAL:
                    envs[e]._set_new_goal(goal_joint_angle=p[0:3])                           # test_roboy_env.py:62-63
                elif kind == gr.EV_SET_STATE:
                    clients[e].set_state(p[0:3], p[3:6], feasible=[bool(p[6])])              # test_roboy_env.py:76
                else:
                    envs[e].step_num = int(p[6])                                             # test_roboy_env.py:173
            sync_counter()
            out = [env.step(fx["actions"][t][e]) for e, env in enumerate(envs)]              # roboy_env.py:51-70
            t_global += 1
            obs = np.stack([o[0] for o in out]).astype(np.float32)
            rew = np.array([o[1] for o in out]); done = np.array([o[2] for o in out])
            assert all(isinstance(o[1], float) and isinstance(o[2], bool) for o in out)
            assert np.array_equal(done, fx["done"][t]), "done mask, step %d" % t
            assert np.array_equal(obs, fx["obs"][t]), "observations, step %d" % t
            rel = np.abs(rew - fx["reward"][t]) / np.maximum(np.abs(fx["reward"][t]), 1e-30)
            worst = max(worst, float(rel.max()))
            assert rel.max() <= gr.REWARD_RTOL, "reward, step %d: %g" % (t, rel.max())
            if done.any():
                sync_counter()
                for e in np.flatnonzero(done):
                    ro = envs[e].reset()
                    assert np.array_equal(ro.astype(np.float32), fx["reset_obs_after"][t][e]), "reset obs, step %d" % t
                t_global += 1
            assert np.array_equal(goals(), fx["goal_after"][t]), "goals, step %d" % t
            assert np.array_equal(np.array([e.step_num for e in envs]), fx["step_num_after"][t]), "step_num, step %d" % t
    assert fx["done"].sum() > 30 and worst < gr.REWARD_RTOL
    for c in clients:
        c.close()
