"""The oracle's stand-alone compute_reward / _did_reach_goal pinned against the UNMODIFIED reference on SALTED values:
NaN, +-inf, 3e38 (squares overflow float32), denormals, the bounds of the spaces, zeros -- in the state, the goal and the
goal velocities -- for MSJ and two other robots, all four flag combinations.  tools/soak_api.py compares the CUDA path
with the oracle on such values; this test is what makes the oracle's answers there the reference's answers
(roboy_env.py:92-112 incl. the assert at :109, :125-134, :137-140; roboy_robot.py:80-95)."""
import contextlib
import io
import re
import warnings

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import reference_harness as rh
from test_oracle_vs_reference import GENERIC_ROBOTS

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference checkout not present")

ROBOTS = {"msj": {}, "six_joints_14_tendons": GENERIC_ROBOTS["six_joints_14_tendons"],
          "five_joints_11_tendons_per_component": GENERIC_ROBOTS["five_joints_11_tendons_per_component"]}
SALT = np.array([np.nan, np.inf, -np.inf, 3e38, -3e38, 1e-42, -1e-42, 0.0, 1e-20, 2.5e19, -2.5e19], np.float32)


def salted(rng, lo, hi, n, p_salt):
    """n rows inside [lo, hi] (float32), each component replaced by a salt value / a bound with probability p_salt"""
    J = lo.size
    x = rng.uniform(lo, hi, (n, J)).astype(np.float32)
    pick = rng.random((n, J))
    s = SALT[rng.integers(0, SALT.size, (n, J))]
    x = np.where(pick < p_salt, s, x)
    x = np.where((pick >= p_salt) & (pick < p_salt + 0.05), lo, x)
    x = np.where((pick >= p_salt + 0.05) & (pick < p_salt + 0.10), hi, x)
    return x.astype(np.float32)


@pytest.mark.parametrize("bonus", [True, False])
@pytest.mark.parametrize("penalty", [False, True])
@pytest.mark.parametrize("name", sorted(ROBOTS))
def test_compute_reward_on_salted_values_matches_reference(name, penalty, bonus):
    b = ROBOTS[name]
    J, A, _, bb = orc.robot_bounds(b)
    RoboyEnv, MsjRobot, RobotState, Stub = rh._import_reference()
    robot = (rh._custom_reference_robot(b) if b else MsjRobot)()
    with contextlib.redirect_stdout(io.StringIO()):
        env = RoboyEnv(Stub(robot=robot), joint_vel_penalty=penalty, is_agent_getting_bonus_for_reaching_goal=bonus)
    lo_r, hi_r = orc.reward_range(joint_vel_penalty=penalty, bonus=bonus, **b)
    assert np.allclose(env.reward_range, (lo_r, hi_r), rtol=1e-6, atol=0)
    cfg = orc.make_cfg(1, joint_vel_penalty=penalty, bonus=bonus, reward_range=env.reward_range, **b)
    rng = np.random.default_rng(77 + 2 * penalty + bonus)
    n = 700
    q = salted(rng, bb["angle_low"], bb["angle_high"], n, 0.12)
    qd = salted(rng, bb["vel_low"], bb["vel_high"], n, 0.12)
    g = salted(rng, bb["angle_low"], bb["angle_high"], n, 0.08)
    # a third of the goals next to the state: the reached test and the bonus fire, also with salted components
    near = rng.random(n) < 0.33
    g[near] = (q[near].astype(np.float64) + rng.normal(size=(near.sum(), J)) * 0.01).astype(np.float32)
    qd[near & (rng.random(n) < 0.7)] *= np.float32(0.05)
    gqd = salted(rng, bb["vel_low"], bb["vel_high"], n, 0.10)
    use_gqd = rng.random(n) < 0.4          # else the float64 zeros of roboy_env.py:23
    feasible = rng.random(n) < 0.7
    o_rew = np.empty(n); o_reached = np.empty(n, bool); o_viol = np.empty(n, bool)
    for sel, with_gqd in ((use_gqd, True), (~use_gqd, False)):
        r, re_, v = orc.compute_reward(cfg, q[sel], qd[sel], feasible[sel], g[sel], gqd[sel] if with_gqd else None)
        o_rew[sel], o_reached[sel], o_viol[sel] = r, re_, v
    kinds = {"nan": 0, "inf": 0, "finite": 0, "raised": 0, "reached": 0}
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        for i in range(n):
            cur = RobotState(joint_angles=q[i].copy(), joint_vels=qd[i].copy(), is_feasible=bool(feasible[i]))
            goal = RobotState(joint_angles=g[i].copy(), joint_vels=gqd[i].copy() if use_gqd[i] else env._GOAL_JOINT_VEL,
                              is_feasible=True)
            reached = env._did_reach_goal(current_state=cur, goal_state=goal)
            try:
                reward, raised = env.compute_reward(current_state=cur, goal_state=goal), False
            except AssertionError as exc:   # roboy_env.py:109: "'<reward>' not between ..."
                reward, raised = float(re.match(r"'([^']*)'", str(exc)).group(1)), True
            assert reached == o_reached[i], (i, q[i], qd[i], g[i])
            assert raised == o_viol[i], (i, reward, o_rew[i], env.reward_range)
            if np.isnan(reward):
                assert np.isnan(o_rew[i]), (i, reward, o_rew[i])
                kinds["nan"] += 1
            elif np.isinf(reward):
                assert o_rew[i] == reward, (i, reward, o_rew[i])
                kinds["inf"] += 1
            else:
                assert abs(o_rew[i] - reward) <= 1e-6 * abs(reward), (i, reward, o_rew[i])
                kinds["finite"] += 1
            kinds["raised"] += raised
            kinds["reached"] += reached
    assert kinds["inf"] > 20 and kinds["finite"] > 100 and kinds["raised"] > 50 and kinds["reached"] > 10, kinds
    assert kinds["nan"] > 5 or not penalty, kinds   # NaN survives only through :99 (no NaN guard there)
