"""Known-answer pins the reference's own tests and constants place on the hot path
(SURVEY.md section 4 and the KAT table of section 8a), checked on the oracle."""
import numpy as np

from oracle import oracle as orc

CFG = orc.make_cfg(1)


def test_thresholds_and_reward_ranges():
    ta, tv = orc.thresholds(CFG)
    assert float(ta).hex() == "0x1.bdc2640000000p-5" and float(tv).hex() == "0x1.7377540000000p-2"
    assert orc.reward_range(False, True) == (-32.947744369506836, 999.0)
    assert orc.reward_range(False, False) == (-32.947744369506836, -1.0)
    assert orc.reward_range(True, True) == (-143.61798095703125, 998.6321411132812)
    worst = -np.exp(np.linalg.norm(2 * np.ones(3))) - 1            # test_roboy_env.py:82-86
    assert np.isclose(orc.reward_range(False, True)[0], worst)


def test_kat_table():
    rows = [  # q, qd, goal, feasible, reward(pen=False), reward(pen=True), reached
        ((0.5, -1.0, 2.0), (0.1, -0.2, 0.3), (0.25, -0.75, 1.5), True, -1.2152187824249268, -2.5922478480921027, False),
        ((0.26, -0.74, 1.51), (0.05, 0, -0.05), (0.25, -0.75, 1.5), True, 998.9944458007812, 998.4434189188644, True),
        ((0.26, -0.74, 1.51), (0.3, 0.3, 0.3), (0.25, -0.75, 1.5), True, -1.0055285692214966, -2.7323260694901483, False),
        ((3, -3, 3), (0.5, -0.5, 0.5), (-3, 3, -3), False, -28.329681396484375, -73.53261015329933, False),
        ((1, 1, 1), (0, 0, 0), (1, 1, 1), True, 999.0, 998.6321206092834, True),
    ]
    for q, qd, g, feas, r0, r1, reached in rows:
        for pen, want in ((False, r0), (True, r1)):
            cfg = orc.make_cfg(1, joint_vel_penalty=pen)
            r, got_reached, _ = orc.compute_reward(cfg, [q], [qd], [feas], [g])
            assert abs(r[0] - want) <= 1e-6 * abs(want), (q, pen, r[0], want)
            assert bool(got_reached[0]) is reached


def test_env_protocol_pins():
    env = orc.OracleEnv(64, seed=3, auto_reset=False)
    g0 = env.goal.copy()
    obs = env.reset()
    assert not obs[:, :6].any() and (env.step_num == 1).all()            # test_roboy_env.py:36-46,183-188
    assert not np.array_equal(g0, env.goal) and (np.abs(env.goal) <= orc.PI32).all()   # :49-57
    env.goal[:] = 0                                                      # goal := current (zero) state
    obs, rew, done = env.step(np.zeros((64, 8), np.float32))
    assert np.allclose(rew, 999.0) and done.all()                        # :60-68
    # done does not reset step_num without reset(); the terminal obs carries the OLD goal
    assert (env.step_num == 2).all() and not obs[:, 6:].any() and env.goal.any()
    env.reset()
    env.step_flags[:] = (env.step_flags & ~np.uint32(orc.STEP_MASK)) | np.uint32(399)
    far = np.full((64, 8), 0.5, np.float32)
    _, _, done = env.step(far)
    assert not done.any()                                                # :170-180
    _, _, done = env.step(far)
    assert done.all()
    _, _, done = env.step(far)
    assert done.all()                                                    # stays done until reset


def test_hold_interval_boundaries():
    """np.allclose(rescaled, 0) <=> every component in [-2^-24, 2^-25]  (SURVEY 8a a4)."""
    up, dn = np.float32(2.0 ** -25), np.float32(-(2.0 ** -24))
    rows = [(np.full(8, up), True), (np.full(8, np.nextafter(up, np.float32(1))), False),
            (np.full(8, dn), True), (np.full(8, np.nextafter(dn, np.float32(-1))), False),
            (np.zeros(8), True), (np.array([0, 0, 0, 0, 0, 0, 0, 1e-7]), False)]
    env = orc.OracleEnv(len(rows), seed=1)
    env.reset()
    a = np.stack([r for r, _ in rows]).astype(np.float32)
    obs, _, _ = env.step(a)
    for i, (_, hold) in enumerate(rows):
        assert (not obs[i, :6].any()) == hold, i     # hold returns the (zero) held state
    assert env.stats()["holds"] == sum(h for _, h in rows)
