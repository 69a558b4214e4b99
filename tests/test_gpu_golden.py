"""CUDA path against the golden vectors recorded from the unmodified reference."""
import pytest

import golden_replay

pytestmark = pytest.mark.gpu

FIXTURES = ["rollout_default", "rollout_penalty", "rollout_nobonus", "rollout_manual_reset"]


@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_reproduces_reference_rollout(name):
    from cuda_adaptor import CudaAdaptor
    fx = golden_replay.load(name)
    worst = golden_replay.replay(fx, CudaAdaptor(fx))
    assert worst <= golden_replay.REWARD_RTOL
