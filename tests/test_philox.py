"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) for both oracle implementations,
and the draw -> float32 mapping."""
import numpy as np

from oracle import oracle as orc

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_known_answers_c_and_numpy():
    L = orc.lib()
    for ctr, key, want in KATS:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        L.orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(x) for x in o) == want
        got = orc.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_draws_stay_inside_the_angle_space_and_match_between_implementations():
    n = 100000
    gids = np.arange(5, 5 + n)
    q, qd = orc.draw_state(987654321, gids, 123)
    for d in (q, qd, orc.draw_goal(987654321, gids, 123)):
        assert d.dtype == np.float32 and (np.abs(d) <= orc.PI32).all()
        assert abs(d.mean()) < 0.02 and abs(d.std() - 2 * np.pi / np.sqrt(12)) < 0.02
    six = np.concatenate([q, qd], axis=1)           # the six 21-bit fields of one block are independent
    assert np.abs(np.corrcoef(six.T) - np.eye(6)).max() < 0.02
    env = orc.OracleEnv(n, seed=987654321, env_id_base=5)
    q0, qd0 = orc.draw_state(987654321, gids, 0)
    assert np.array_equal(env.goal.T, orc.draw_goal(987654321, gids, 0))
    assert np.array_equal(env.held[0:3].T, q0) and np.array_equal(env.held[3:6].T, qd0)
    # 64-bit env ids and call counters reach the upper counter words
    big = orc.draw_goal(1, np.array([2 ** 40 + 3], np.uint64), 2 ** 33 + 1)
    small = orc.draw_goal(1, np.array([3], np.uint64), 1)
    assert not np.array_equal(big, small)


def test_draw_fields_are_uniform_and_uncorrelated():
    """Statistical sanity of the draw mapping: each of the six 21-bit state fields and the three
    goal words is uniform (chi-square over 64 bins), and draws are uncorrelated across neighbouring
    env ids and consecutive call counters."""
    n = 400_000
    gids = np.arange(n)
    q, qd = orc.draw_state(42, gids, 7)
    g = orc.draw_goal(42, gids, 7)
    cols = [q[:, k] for k in range(3)] + [qd[:, k] for k in range(3)] + [g[:, k] for k in range(3)]
    for c in cols:
        hist, _ = np.histogram(c, bins=64, range=(-float(orc.PI32), float(orc.PI32)))
        chi2 = ((hist - n / 64) ** 2 / (n / 64)).sum()
        assert chi2 < 130, chi2            # 63 dof: mean 63, p(>130) ~ 1e-6
    q_next_env = orc.draw_state(42, gids + 1, 7)[0]
    q_next_call = orc.draw_state(42, gids, 8)[0]
    for other in (q_next_env, q_next_call, qd, g):
        r = np.corrcoef(q[:, 0], other[:, 0])[0, 1]
        assert abs(r) < 0.01, r
    # the 21-bit grid: values are low + k * span * 2^-21
    span21 = np.float32(np.float32(2 * orc.PI32) * np.float32(2.0 ** -21))
    k = np.round((q[:1000, 0].astype(np.float64) + float(orc.PI32)) / float(span21))
    assert np.abs(k * float(span21) - float(orc.PI32) - q[:1000, 0]).max() < 1e-6 and k.max() < 2 ** 21
