"""Host-side pieces that need no GPU: spaces, robots (reference test_robot_state.py:7-60
restated), registration, sharding arithmetic."""
import numpy as np
import pytest

import gym_roboy_b200
from gym_roboy_b200.envs.robots import MsjRobot, RobotState, RoboyRobot
from gym_roboy_b200.sharding import shard_range, summarize
from gym_roboy_b200.spaces import Box

MSJ = MsjRobot()


def test_box_old_gym_semantics():
    b = Box(-1, 1, (8,), "float32")
    assert b.contains(np.zeros(8)) and b.contains(np.ones(8, np.float32)) and b.contains(-np.ones(8))
    assert not b.contains(np.full(8, 1.0000001)) and not b.contains(np.full(8, np.nan)) and not b.contains(np.zeros(7))
    assert b.low.dtype == np.float32 and b.sample().dtype == np.float32 and b.contains(b.sample())
    c = Box(low=-np.random.uniform(size=3), high=np.random.uniform(size=3), dtype="float32")
    assert c.shape == (3,)


def test_msj_spaces_are_the_reference_float32_constants():       # msj_robot.py:8-16
    assert float(MSJ.get_joint_angles_space().high[0]).hex() == "0x1.921fb60000000p+1"
    assert float(MSJ.get_joint_vels_space().high[0]).hex() == "0x1.0c15240000000p-1"
    assert float(MSJ.get_action_space().high[0]).hex() == "0x1.3333340000000p-2"
    assert MSJ.get_action_space().shape == (8,) and MSJ.get_joint_angles_space().shape == (3,)
    with pytest.raises(NotImplementedError):
        RoboyRobot.get_action_space()


def test_robot_state_interpolate():                              # test_robot_state.py:7-13
    s1, s2 = MsjRobot.new_random_state(), MsjRobot.new_random_state()
    mid = RobotState.interpolate(s1, s2)
    assert np.allclose(mid.joint_angles, (s1.joint_angles + s2.joint_angles) / 2)
    assert np.allclose(mid.joint_vels, (s1.joint_vels + s2.joint_vels) / 2)
    assert mid.is_feasible is True


def test_state_factories():                                      # test_robot_state.py:16-38
    za = MsjRobot.new_random_zero_angles_state()
    assert np.allclose(za.joint_angles, 0) and not np.allclose(za.joint_vels, 0)
    zv = MsjRobot.new_random_zero_vels_state()
    assert np.allclose(zv.joint_vels, 0) and not np.allclose(zv.joint_angles, 0)
    assert MsjRobot.new_zero_state().joint_angles.dtype == np.float64          # roboy_robot.py:43-44
    mx, mn = MSJ.new_max_state(), MSJ.new_min_state()
    assert np.allclose(mx.joint_angles, MSJ.get_joint_angles_space().high) and mx.is_feasible is False
    assert np.allclose(mn.joint_vels, MSJ.get_joint_vels_space().low)
    rs = MsjRobot.new_random_state()
    assert np.abs(rs.joint_vels).max() <= np.pi    # quirk: velocities come from the ANGLE space (roboy_robot.py:38)


def test_normalize():                                            # test_robot_state.py:41-60
    one = np.ones(3)
    n = MSJ.normalize_state(MSJ.new_max_state())
    assert np.allclose(n.joint_angles, one) and np.allclose(n.joint_vels, one)
    n = MSJ.normalize_state(MSJ.new_min_state())
    assert np.allclose(n.joint_angles, -one) and np.allclose(n.joint_vels, -one)
    hi = np.random.random(); lo = hi - abs(np.random.random())
    assert np.isclose(1, MSJ._normalize_between_minus1_and1(hi, max_val=hi, min_val=lo))
    assert np.isclose(-1, MsjRobot._normalize_between_minus1_and1(lo, max_val=hi, min_val=lo))


def test_new_state_checks_angles_and_feasible_type():            # roboy_robot.py:71-78
    MsjRobot.new_state([0, 0, 0], [9, 9, 9], True)               # velocities are NOT checked (:77)
    with pytest.raises(AssertionError):
        MsjRobot.new_state([4, 0, 0], [0, 0, 0], True)
    with pytest.raises(TypeError):
        MsjRobot.new_state([0, 0, 0], [0, 0, 0], 1)
    with pytest.raises(TypeError):
        RobotState([0, 0, 0], [0, 0, 0], "yes")


def test_registration():
    entry, kwargs = gym_roboy_b200.spec("msj-control-v1")
    assert entry == "gym_roboy_b200.envs:RoboyEnv" and kwargs == {}
    with pytest.raises(KeyError):
        gym_roboy_b200.make("msj-control-v0")                    # the reference's id never worked; not provided
    gym_roboy_b200.register("unit-test-env-v0", entry_point=lambda **kw: ("made", kw), kwargs=dict(a=1))
    assert gym_roboy_b200.make("unit-test-env-v0", b=2) == ("made", dict(a=1, b=2))


@pytest.mark.parametrize("total,world", [(1 << 20, 8), (1000, 3), (7, 8), (4096, 1), (134_217_728, 8)])
def test_shard_range_partitions_the_population(total, world):
    blocks = [shard_range(total, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == total
    assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
    sizes = [e - b for b, e in blocks]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(total, world, world)


def test_summarize():
    import torch
    s = summarize(torch.tensor([1000.0, 10, 4, 6, -2500.0, 900, 3, 0], dtype=torch.float64))
    assert s["success_rate"] == 0.4 and s["mean_step_reward"] == -2.5 and s["mean_episode_len"] == 90


def test_rescale_from_one_space_to_other():                      # test_roboy_env.py:195-207
    from gym_roboy_b200.envs.roboy_env import _l2_distance, _rescale_from_one_space_to_other
    np.random.seed(0)
    mk = lambda: Box(low=-np.random.uniform(size=8), high=np.random.uniform(size=8), dtype="float32")  # noqa: E731
    src, dst = mk(), mk()
    assert np.allclose(_rescale_from_one_space_to_other(input_val=src.high, input_space=src, output_space=dst), dst.high)
    assert np.allclose(_rescale_from_one_space_to_other(input_val=src.low, input_space=src, output_space=dst), dst.low)
    with pytest.raises(AssertionError):
        _rescale_from_one_space_to_other(input_val=src.high * 2, input_space=src, output_space=dst)
    # the MSJ action rescale is the map whose zero set the kernel's hold interval encodes
    unit, msj = Box(-1, 1, (8,), "float32"), MsjRobot.get_action_space()
    out = _rescale_from_one_space_to_other(np.zeros(8, np.float32), unit, msj)
    assert out.dtype == np.float32 and np.allclose(out, 0, atol=1e-8)
    assert _l2_distance(np.array([np.inf, 1.0, 0.0]), np.array([np.inf, 0.0, 0.0])) == 1.0   # inf - inf counts as 0


def test_policy_image_layouts_match_the_header():
    """The packed-policy layouts are ABI: _native's constants must equal the #defines of include/roboy_b200.h, and the
    packers must put every weight where the kernels' index formulas (header comment) say."""
    import os
    import re
    import torch
    from gym_roboy_b200 import _native as N
    from gym_roboy_b200.rollout import MlpPolicy, pack_policy_image, pack_policy_image_tc
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "roboy_b200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define (ROBOY_(?:POLICY|TC)_\w+) (\d+)", hdr)}
    for name, val in defs.items():
        assert getattr(N, name[len("ROBOY_"):]) == val, name
    torch.manual_seed(3)
    pol = MlpPolicy()
    img = pack_policy_image(pol)
    assert img.numel() == N.POLICY_IMAGE_FLOATS
    for base, net in ((N.POLICY_OFF_VF, pol.vf), (N.POLICY_OFF_PI, pol.pi)):
        w1, w2, w3 = (net[i].weight.detach() for i in (0, 2, 4))
        assert img[base + N.POLICY_OFF_W1 + 5 * 64 + 17] == w1[17, 5]                  # W1^T [9][64]
        assert img[base + N.POLICY_OFF_W2 + 40 * 64 + 3] == w2[3, 40]                  # W2^T [64][64]
        assert img[base + N.POLICY_OFF_W3 + 63 * 8 + 0] == w3[0, 63]                   # W3^T [64][8]
        assert img[base + N.POLICY_OFF_B3] == net[4].bias[0]
    assert torch.equal(img[N.POLICY_OFF_STD:N.POLICY_OFF_STD + 8], pol.log_std.detach().exp())
    tc = pack_policy_image_tc(pol)
    assert tc.numel() == N.TC_IMAGE_BYTES and tc.dtype == torch.uint8
    halves = tc[:N.TC_OFF_LO_BYTES].view(torch.float16)
    lows = tc[N.TC_OFF_LO_BYTES:N.TC_OFF_STD_BYTES].view(torch.float16)

    def at(base, off, K, n, k):
        return halves[base + off + (n // 8) * (K // 8) * 64 + (k // 8) * 64 + (n % 8) * 8 + (k % 8)]
    for base, net in ((N.TC_OFF_VF, pol.vf), (N.TC_OFF_PI, pol.pi)):
        w1, w2, w3 = (net[i].weight.detach() for i in (0, 2, 4))
        assert at(base, N.TC_OFF_W1, 16, 17, 5) == w1[17, 5].half() and at(base, N.TC_OFF_W1, 16, 17, 9) == net[0].bias[17].half()
        assert at(base, N.TC_OFF_W1, 16, 17, 10) == 0
        assert at(base, N.TC_OFF_W2, 80, 3, 40) == w2[3, 40].half() and at(base, N.TC_OFF_W2, 80, 3, 64) == net[2].bias[3].half()
        assert at(base, N.TC_OFF_W3, 80, 0, 63) == w3[0, 63].half() and at(base, N.TC_OFF_W3, 80, 0, 64) == net[4].bias[0].half()
        assert at(base, N.TC_OFF_W3, 80, 15, 63) == 0                                   # output rows padded to 16
        i = base + N.TC_OFF_W2 + (3 // 8) * (80 // 8) * 64 + (40 // 8) * 64 + (3 % 8) * 8 + (40 % 8)
        assert lows[i] == (w2[3, 40] - w2[3, 40].half().float()).half()                 # low half of the split
        # (a small weight's low half is a float16 subnormal: absolute spacing 2^-24)
        assert abs(float(halves[i].float() + lows[i].float() - w2[3, 40])) <= 2.0 ** -21 * abs(float(w2[3, 40])) + 2.0 ** -25
    tail = tc[N.TC_OFF_STD_BYTES:].view(torch.float32)
    assert torch.equal(tail[:8], pol.log_std.detach().exp())


def test_packed_policy_images_reproduce_the_policy_on_the_cpu():
    """Decode both packed images with nothing but the header's layout formulas and run the forward pass from them:
    the float32 image must reproduce the torch policy to rounding, the tensor-core image (hi + lo halves, bias as an
    extra input column) to float32 level with both halves and to ~1e-3 with the high halves alone."""
    import torch
    from gym_roboy_b200 import _native as N
    from gym_roboy_b200.rollout import MlpPolicy, pack_policy_image, pack_policy_image_tc
    torch.manual_seed(11)
    pol = MlpPolicy()
    with torch.no_grad():
        pol.log_std.copy_(torch.linspace(-0.5, 0.3, 8))
    obs = torch.rand(257, 9) * 6.0 - 3.0
    with torch.no_grad():
        want_mean, want_value = pol(obs)

    img = pack_policy_image(pol).double()

    def fwd32(base, n_out):
        w1 = img[base + N.POLICY_OFF_W1: base + N.POLICY_OFF_B1].view(9, 64)
        w2 = img[base + N.POLICY_OFF_W2: base + N.POLICY_OFF_B2].view(64, 64)
        w3 = img[base + N.POLICY_OFF_W3: base + N.POLICY_OFF_B3].view(64, 8)
        b1, b2 = img[base + N.POLICY_OFF_B1: base + N.POLICY_OFF_W2], img[base + N.POLICY_OFF_B2: base + N.POLICY_OFF_W3]
        b3 = img[base + N.POLICY_OFF_B3: base + N.POLICY_NET_FLOATS]
        h = torch.tanh(torch.tanh(obs.double() @ w1 + b1) @ w2 + b2) @ w3 + b3
        return h[:, :n_out]
    assert torch.allclose(fwd32(N.POLICY_OFF_PI, 8).float(), want_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(fwd32(N.POLICY_OFF_VF, 1)[:, 0].float(), want_value, rtol=1e-5, atol=1e-6)

    tc = pack_policy_image_tc(pol)
    halves = (tc[:N.TC_OFF_LO_BYTES].view(torch.float16).double(), tc[N.TC_OFF_LO_BYTES:N.TC_OFF_STD_BYTES].view(torch.float16).double())

    def dense(part, base, off, n_pad, k_pad):      # undo the K-major core-matrix layout: [n/8, k/8, 8, 8] -> [n, k]
        flat = part[base + off: base + off + n_pad * k_pad]
        return flat.view(n_pad // 8, k_pad // 8, 8, 8).permute(0, 2, 1, 3).reshape(n_pad, k_pad)

    def fwd_tc(base, n_out, use_lo):
        def lin(x, off, n_pad, k_pad):             # x: [batch, k_pad] with the constant 1 already in place
            w = dense(halves[0], base, off, n_pad, k_pad) + (dense(halves[1], base, off, n_pad, k_pad) if use_lo else 0)
            return x @ w.t()
        x = torch.zeros(obs.shape[0], 16, dtype=torch.float64); x[:, :9] = obs; x[:, 9] = 1.0
        h = torch.tanh(lin(x, N.TC_OFF_W1, 64, 16))
        for off, n_pad in ((N.TC_OFF_W2, 64), (N.TC_OFF_W3, 16)):
            x = torch.zeros(obs.shape[0], 80, dtype=torch.float64); x[:, :64] = h; x[:, 64] = 1.0
            h = lin(x, off, n_pad, 80)
            if n_pad == 64:
                h = torch.tanh(h)
        return h[:, :n_out]
    assert torch.allclose(fwd_tc(N.TC_OFF_PI, 8, True).float(), want_mean, rtol=1e-5, atol=2e-6)
    assert torch.allclose(fwd_tc(N.TC_OFF_VF, 1, True)[:, 0].float(), want_value, rtol=1e-5, atol=2e-6)
    assert float((fwd_tc(N.TC_OFF_PI, 8, False).float() - want_mean).abs().max()) < 2e-3    # high halves alone: float16 weights
    tail = tc[N.TC_OFF_STD_BYTES:].view(torch.float32)
    assert torch.equal(tail[12:76], pol.vf[2].bias.detach()) and torch.equal(tail[12 + 80:12 + 80 + 64], pol.pi[2].bias.detach())
    assert tail[12 + 64] == pol.vf[4].bias.detach()[0] and torch.equal(tail[12 + 80 + 64:12 + 80 + 72], pol.pi[4].bias.detach())
