"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

    python -m oracle.gen_golden            (build container only: needs /root/reference)

Each fixture is a lock-step rollout of N reference `RoboyEnv(StubSimulationClient(ReplayRobot))`
instances under the vec-env worker's reset-on-done (oracle/reference_harness.py), driven by a
committed action tensor and a list of injection events, with every output of every step
recorded.  The actions contain exact-zero rows and rows at the boundaries of numpy's
`allclose(action, 0)` (hold branch); the events poke goals / held states / step counters the way
the reference's own tests do (test_roboy_env.py:62-63,76,173) so that goal-reached, the
infeasible penalty, timeouts and near-threshold distances all occur.

Envs whose reference instance raised an AssertionError (roboy_env.py:109 fires with
joint_vel_penalty=True because the Stub draws velocities from the angle space) are marked
invalid from that step on; the step at which they raised is recorded in `raised_at`.
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import oracle as orc  # noqa: E402
from oracle.reference_harness import ReferenceVecEnv  # noqa: E402

GOLDEN_DIR = os.path.join(_ROOT, "tests", "golden")

EV_SET_GOAL, EV_SET_STATE, EV_SET_STEP = 0, 1, 2  # event kinds


def make_actions(rng, T, N, bounds=None):
    if bounds is not None:   # another robot: hold rows sit at the per-tendon zero crossing of the rescale
        _, A, _, _ = orc.robot_bounds(bounds)
        zero, _ = orc.hold_action(bounds)
        a = rng.uniform(-1, 1, (T, N, A)).astype(np.float32)
        a[rng.random((T, N)) < 0.05] = zero
        a[rng.random((T, N)) < 0.01] = np.nextafter(zero, np.float32(1))      # right next to the hold interval
        a[rng.random((T, N)) < 0.01] = np.nextafter(zero, np.float32(-1))
        one_off = zero.copy(); one_off[-1] = np.float32(0.5)
        a[rng.random((T, N)) < 0.01] = one_off                                # one tendon off: no hold
        a[3, 1] = np.float32(1.0)                                             # the ends of the action space
        a[5, 2] = np.float32(-1.0)
        return a, zero
    a = rng.uniform(-1, 1, (T, N, 8)).astype(np.float32)
    a[rng.random((T, N)) < 0.03] = 0.0                      # hold branch
    up, dn = np.float32(2.0 ** -25), np.float32(-(2.0 ** -24))
    specials = [
        np.full(8, up, np.float32),                          # largest positive value that still holds
        np.full(8, np.nextafter(up, np.float32(1)), np.float32),
        np.full(8, dn, np.float32),                          # most negative value that still holds
        np.full(8, np.nextafter(dn, np.float32(-1)), np.float32),
        np.array([0, 0, 0, 0, 0, 0, 0, 1e-7], np.float32),   # one component off
        np.array([up, dn, 0, 0, up, dn, 0, 0], np.float32),
        np.full(8, 1.0, np.float32), np.full(8, -1.0, np.float32),
    ]
    for k, row in enumerate(specials):
        for rep in range(3):
            a[(7 * k + 31 * rep + 5) % T, (5 * k + 11 * rep + 2) % N] = row
    return a


def make_events(rng, T, N, seed, thr_a, thr_v, J=3, lo=None, hi=None):
    """List of (before_step, env, kind, payload[2J+1]) -- payload: q[J], qd[J], flag."""
    ev = []
    zero_action_at = []  # (step, env) whose action row must hold (exactly zero for MSJ)
    lo = np.full(J, -3.0) if lo is None else np.asarray(lo, np.float64) * 0.95
    hi = np.full(J, 3.0) if hi is None else np.asarray(hi, np.float64) * 0.95
    rootJ = np.float32(np.sqrt(J))
    zeros = (0.0,) * J

    def add(t, e, kind, q=zeros, qd=zeros, flag=0.0):
        ev.append((t, e, kind, list(q) + list(qd) + [flag]))

    def first(d):       # (d, 0, 0, ...)
        return (d,) + (0.0,) * (J - 1)

    def alternating(d):  # (d, -d, d, ...)
        return tuple(d if k % 2 == 0 else -d for k in range(J))

    # (a) goal := current (zero) state right after the initial reset -> reached on a zero action
    #     (test_roboy_env.py:60-68), exactly at / just inside / just outside the angle threshold.
    for k, scale in enumerate([0.0, 0.5, 1.0 - 3e-7, 1.0, 1.0 + 3e-7, 2.0]):
        e, t = k, 3 + k
        d = np.float32(thr_a) * np.float32(scale)
        add(t, e, EV_SET_STATE, zeros, zeros, 1.0)                           # float32 held state
        add(t, e, EV_SET_GOAL, first(d))
        zero_action_at.append((t, e))
    # (b) float32 held state at the goal but moving: velocity norm around its threshold
    #     (test_roboy_env.py:71-79)
    for k, scale in enumerate([0.5, 1.0 - 3e-7, 1.0, 1.0 + 3e-7, 3.0]):
        e, t = 8 + k, 10 + k
        q = rng.uniform(lo, hi).astype(np.float32)
        v = np.float32(thr_v) * np.float32(scale) / rootJ
        add(t, e, EV_SET_STATE, q, (v,) * J, 1.0)
        add(t, e, EV_SET_GOAL, q)
        zero_action_at.append((t, e))
    # (c) infeasible held state -> boundary penalty (roboy_env.py:102-103), near and far from goal
    for k in range(3):
        e, t = 14 + k, 20 + k
        q = rng.uniform(lo, hi).astype(np.float32)
        add(t, e, EV_SET_STATE, q, rng.uniform(-0.2, 0.2, J).astype(np.float32), 0.0)
        add(t, e, EV_SET_GOAL, q if k == 0 else rng.uniform(lo, hi).astype(np.float32))
        zero_action_at.append((t, e))
    # (d) float64 zero held state (after a reset) with the goal near zero -> reached in float64
    for k, scale in enumerate([0.3, 1.0 - 1e-7, 1.0 + 1e-7]):
        e, t = 18 + k, 30 + k
        d = np.float32(thr_a) * np.float32(scale) / rootJ
        add(t, e, EV_SET_GOAL, alternating(d))
        zero_action_at.append((t, e))
    # (e) episode counter pokes (test_roboy_env.py:170-180): timeouts at scattered steps
    for k in range(8):
        add(5 + 13 * k, (22 + k) % N, EV_SET_STEP, flag=float(395 + k % 6))
    return ev, zero_action_at


def sampled_branch_goals(seed, N, T, thr_v, counter_of_step, taken, low=orc.MSJ["angle_low"], high=orc.MSJ["angle_high"], J=3):
    """Events that make _did_reach_goal fire on freshly SAMPLED states: pick (env, step) whose
    Philox velocity draw is slow enough and put the goal next to the drawn angles just before."""
    ev = []
    for t in range(40, T):
        c = counter_of_step(t)
        q_all, qd = orc.draw_state(seed, np.arange(N), c, low, high, J=J)
        qd = qd.astype(np.float64)
        slow = np.flatnonzero(np.sqrt((qd * qd).sum(1)) < 0.9 * float(thr_v))
        for e in slow:
            if (t, int(e)) in taken:
                continue
            q = q_all[int(e)]
            g = np.clip(q + np.float32(0.01), np.asarray(low, np.float32), np.asarray(high, np.float32))
            ev.append((t, int(e), EV_SET_GOAL, list(g) + [0] * (J + 1)))
    return ev


def run_fixture(name, N, T, seed, joint_vel_penalty, bonus, auto_reset=True, bounds=None):
    rng = np.random.default_rng(seed)
    J, A, D, extra = 3, 8, 9, {}
    if bounds is None:
        cfg = orc.make_cfg(N, seed=seed)
        thr_a, thr_v = orc.thresholds(cfg)
        actions = make_actions(rng, T, N)
        events, zero_rows = make_events(rng, T, N, seed, thr_a, thr_v)
        for (t, e) in zero_rows:
            actions[t, e] = 0.0
        # call counter: 0 = construction, 1 = initial reset, step index t runs at counter t + 2
        events += sampled_branch_goals(seed, N, T, thr_v, lambda t: t + 2, set(zero_rows))
    else:   # a robot with other dims / per-component bounds (a RoboyRobot subclass of the reference)
        J, A, _, bb = orc.robot_bounds(bounds)
        D = 3 * J
        cfg = orc.make_cfg(N, seed=seed, **bounds)
        thr_a, thr_v = orc.thresholds(cfg)
        actions, zero = make_actions(rng, T, N, bounds)
        events, zero_rows = make_events(rng, T, N, seed, thr_a, thr_v, J, bb["angle_low"], bb["angle_high"])
        for (t, e) in zero_rows:
            actions[t, e] = zero
        events += sampled_branch_goals(seed, N, T, thr_v, lambda t: t + 2, set(zero_rows), bb["angle_low"], bb["angle_high"], J)
        extra = {"robot_" + k: v for k, v in bb.items()}
    events.sort(key=lambda x: (x[0], x[1], x[2]))

    ref = ReferenceVecEnv(N, seed=seed, joint_vel_penalty=joint_vel_penalty, bonus=bonus, auto_reset=auto_reset, bounds=bounds)
    out = dict(
        obs=np.zeros((T, N, D), np.float32), reward=np.zeros((T, N), np.float64), done=np.zeros((T, N), bool),
        terminal_obs=np.zeros((T, N, D), np.float32), goal_after=np.zeros((T, N, J), np.float32),
        step_num_after=np.zeros((T, N), np.int32), valid=np.ones((T, N), bool),
    )
    init_goal = ref.goals()
    reset_obs = ref.reset().astype(np.float32)
    alive = np.ones(N, bool)
    raised_at = np.full(N, -1, np.int64)
    by_step = {}
    for ev in events:
        by_step.setdefault(ev[0], []).append(ev)
    for t in range(T):
        for (_, e, kind, p) in by_step.get(t, []):
            if kind == EV_SET_GOAL:
                ref.set_goal(e, p[0:J])
            elif kind == EV_SET_STATE:
                ref.set_state(e, p[0:J], p[J:2 * J], feasible=bool(p[2 * J]))
            else:
                ref.set_step_num(e, int(p[2 * J]))
        o, r, d, term, raised = ref.step(actions[t])
        for i, msg in enumerate(raised):
            if msg and alive[i]:
                alive[i] = False
                raised_at[i] = t
        out["obs"][t], out["reward"][t], out["done"][t] = o.astype(np.float32), r, d
        out["terminal_obs"][t] = term.astype(np.float32)
        out["valid"][t] = alive
        if not auto_reset:  # plain gym loop: the caller resets finished envs
            if d.any():
                ro = ref.reset(mask=d)
                out.setdefault("reset_obs_after", np.zeros((T, N, D), np.float32))[t] = ro.astype(np.float32)
        out["goal_after"][t] = ref.goals()
        out["step_num_after"][t] = ref.step_nums()
    ev_arr = np.array([[e[0], e[1], e[2]] + [float(x) for x in e[3]] for e in events], np.float64)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(
        path, actions=actions, events=ev_arr, init_goal=init_goal, reset_obs=reset_obs, raised_at=raised_at,
        reward_range=np.array(ref.reward_range, np.float64),
        meta=np.array([N, T, seed, int(joint_vel_penalty), int(bonus), int(auto_reset)] + ([J, A] if bounds is not None else []),
                      np.int64), **extra, **out)
    n_succ = int((out["done"] & (out["reward"] > 500) & out["valid"]).sum())
    print("{}: N={} T={} done={} bonus-rewards={} raised={} events={} -> {} ({} KiB)".format(
        name, N, T, int((out["done"] & out["valid"]).sum()), n_succ, int((raised_at >= 0).sum()), len(events),
        os.path.relpath(path, _ROOT), os.path.getsize(path) // 1024))


FIXTURES = [
    # name, N, T, seed, joint_vel_penalty, bonus, auto_reset
    ("rollout_default", 48, 450, 20240901, False, True, True),
    ("rollout_penalty", 48, 450, 20240912, True, True, True),
    ("rollout_nobonus", 32, 420, 20240903, False, False, True),
    ("rollout_manual_reset", 32, 420, 20240904, False, True, False),
]

# robots other than MSJ: synthetic RoboyRobot subclasses of the reference
ROBOT_FIXTURES = [
    ("rollout_six_joints_14_tendons", 32, 300, 20241001, False, True, True,
     dict(dim_joint=6, dim_action=14, angle_low=-2.5, angle_high=2.5, vel_low=-0.6, vel_high=0.6, act_low=-0.2, act_high=0.2)),
    ("rollout_per_component_5_joints_11_tendons", 32, 300, 20241002, False, True, True,
     dict(angle_low=[-3.1, -1.0, -2.0, -0.7, -1.3], angle_high=[3.1, 1.0, 2.5, 0.9, 1.3],
          vel_low=[-0.5, -0.4, -0.3, -0.2, -0.1], vel_high=[0.5, 0.4, 0.6, 0.2, 0.3],
          act_low=[-0.3, -0.1, -0.2, -0.3, -0.25, -0.4, -0.3, -0.25, -0.5, -0.125, -0.3],
          act_high=[0.3, 0.4, 0.2, 0.1, 0.25, 0.4, 0.6, 0.25, 0.5, 0.375, 0.1])),
]

if __name__ == "__main__":
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    which = sys.argv[1:]
    for fx in FIXTURES + ROBOT_FIXTURES:
        if not which or fx[0] in which:
            run_fixture(*fx)
