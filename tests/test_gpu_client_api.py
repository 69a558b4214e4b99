"""The four SimulationClient calls of CudaSimulationClient (un-fused kernels) against the oracle's
draws, and the reference-style single-env protocol driven through them."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_plugin_calls_match_the_stub_semantics():
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    n, seed = 513, 31
    c = CudaSimulationClient(num_envs=n, seed=seed, env_id_base=7, device="cuda:0")
    gids = np.arange(7, 7 + n)
    # __init__: random held state (simulation_client.py:31)
    s = c.read_state()
    q0, qd0 = orc.draw_state(seed, gids, 0)
    assert np.array_equal(s.joint_angles.cpu().numpy(), q0)
    assert np.array_equal(s.joint_vels.cpu().numpy(), qd0)
    assert s.is_feasible.all()
    # zero action -> the held state, unchanged and not advanced (:38-39)
    h = c.forward_step_command(torch.zeros((n, 8)))
    assert torch.equal(h.joint_angles, s.joint_angles) and torch.equal(h.joint_vels, s.joint_vels)
    # non-zero action -> fresh sample that is NOT stored (:40)
    a = torch.full((n, 8), 0.1)
    a[::2] = 0.0
    a[::2, 3] = 2e-8   # just above numpy's atol
    a[1::4] = 0.0      # these hold
    r = c.forward_step_command(a)
    t = c.counter
    fresh_q, _ = orc.draw_state(seed, gids, t)
    hold = np.zeros(n, bool); hold[1::4] = True
    got = r.joint_angles.cpu().numpy()
    assert np.array_equal(got[~hold], fresh_q[~hold]) and np.array_equal(got[hold], s.joint_angles.cpu().numpy()[hold])
    again = c.read_state()
    assert torch.equal(again.joint_angles, s.joint_angles)
    # reset -> zero state (:42-44); goals differ call to call and stay inside the angle space (:46-47)
    z = c.forward_reset_command()
    assert not z.joint_angles.any() and not z.joint_vels.any()
    g1, g2 = c.get_new_goal_joint_angles(), c.get_new_goal_joint_angles()
    assert not torch.equal(g1, g2)
    assert (g1.abs() <= float(orc.PI32)).all() and (g2.abs() <= float(orc.PI32)).all()
    assert np.array_equal(g1.cpu().numpy(), orc.draw_goal(seed, gids, c.counter, sub=0))
    assert np.array_equal(g2.cpu().numpy(), orc.draw_goal(seed, gids, c.counter, sub=1))


def test_single_env_view_returns_reference_types():
    from gym_roboy_b200.envs.robots import RobotState
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    c = CudaSimulationClient(num_envs=1, seed=5, device="cuda:0")
    s = c.forward_step_command([0.1] * 8)
    assert isinstance(s, RobotState) and isinstance(s.joint_angles, np.ndarray) and s.joint_angles.shape == (3,)
    assert isinstance(s.is_feasible, bool)
    g = c.get_new_goal_joint_angles()
    assert isinstance(g, np.ndarray) and g.shape == (3,) and g.dtype == np.float32
    with pytest.raises(AssertionError):
        c.forward_step_command([0.0] * 7)            # simulation_client.py:37
