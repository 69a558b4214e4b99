"""The oracle (C restatement) against the golden vectors generated from the unmodified
reference (oracle/gen_golden.py).  Runs anywhere: needs neither the reference nor a GPU."""
import pytest

import golden_replay
from oracle_adaptor import OracleAdaptor

FIXTURES = ["rollout_default", "rollout_penalty", "rollout_nobonus", "rollout_manual_reset",
            # robots other than MSJ: 6 joints / 14 tendons, and 5 joints / 11 tendons with one bound per component
            "rollout_six_joints_14_tendons", "rollout_per_component_5_joints_11_tendons"]


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_reference_rollout(name):
    fx = golden_replay.load(name)
    worst = golden_replay.replay(fx, OracleAdaptor(fx))
    assert worst <= golden_replay.REWARD_RTOL


def test_fixtures_exercise_the_branches():
    """The fixtures must actually contain what they claim to pin."""
    fx = golden_replay.load("rollout_default")
    valid = fx["valid"]
    success = fx["done"] & (fx["reward"] > 500) & valid            # bonus paid -> goal reached
    timeout = fx["done"] & (fx["reward"] < 500) & valid
    hold = (abs(fx["actions"]).max(axis=2) <= 2.0 ** -24)
    assert success.sum() >= 10 and timeout.sum() >= 30 and hold.sum() >= 300
    assert (fx["step_num_after"] == 1)[fx["done"]].all()           # auto-reset puts step_num back to 1
    pen = golden_replay.load("rollout_penalty")
    assert (pen["raised_at"] >= 0).sum() >= 1                      # roboy_env.py:109 fired in the reference


def test_robot_fixtures_exercise_the_branches():
    for name, J, A in (("rollout_six_joints_14_tendons", 6, 14), ("rollout_per_component_5_joints_11_tendons", 5, 11)):
        fx = golden_replay.load(name)
        assert (fx["J"], fx["A"]) == (J, A) and fx["obs"].shape[2] == 3 * J and fx["actions"].shape[2] == A
        success = fx["done"] & (fx["reward"] > 500) & fx["valid"]
        assert success.sum() >= 5 and fx["done"].sum() > success.sum()          # goal reached, and episodes ending otherwise
        assert (fx["reward"] < -1.5).any() and fx["bounds"] is not None
    per = golden_replay.load("rollout_per_component_5_joints_11_tendons")["bounds"]
    assert len(set(per["angle_high"].tolist())) > 1 and len(set(per["act_low"].tolist())) > 1
