"""CUDA path against the golden vectors recorded from the unmodified reference."""
import pytest

import golden_replay

pytestmark = pytest.mark.gpu

FIXTURES = ["rollout_default", "rollout_penalty", "rollout_nobonus", "rollout_manual_reset",
            # robots other than MSJ: 6 joints / 14 tendons, and 5 joints / 11 tendons with one bound per component
            "rollout_six_joints_14_tendons", "rollout_per_component_5_joints_11_tendons"]


@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_reproduces_reference_rollout(name):
    from cuda_adaptor import CudaAdaptor
    fx = golden_replay.load(name)
    worst = golden_replay.replay(fx, CudaAdaptor(fx))
    assert worst <= golden_replay.REWARD_RTOL
