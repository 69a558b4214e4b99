"""Replays a tests/golden/*.npz fixture (generated from the unmodified reference by
oracle/gen_golden.py) against any implementation exposing the small adaptor below, and checks
every recorded output: done / terminal masks, step counters, goals and observations bit-exactly,
rewards to 1e-6 relative."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EV_SET_GOAL, EV_SET_STATE, EV_SET_STEP = 0, 1, 2
REWARD_RTOL = 1e-6   # north_star: within 1e-6 relative for fp32 observations and rewards


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    fx = {k: z[k] for k in z.files}
    meta = fx["meta"].tolist()
    N, T, seed, pen, bonus, auto = meta[:6]
    J, A = (meta[6], meta[7]) if len(meta) >= 8 else (3, 8)
    fx.update(N=N, T=T, seed=seed, joint_vel_penalty=bool(pen), bonus=bool(bonus), auto_reset=bool(auto), J=J, A=A)
    # robots other than MSJ carry their spaces (per component) in the fixture
    fx["bounds"] = {k[len("robot_"):]: fx[k] for k in fx if isinstance(k, str) and k.startswith("robot_")} or None
    return fx


class Adaptor:
    """What replay() needs from an implementation (numpy in / numpy out)."""

    def goals(self): raise NotImplementedError          # [N,3] float32
    def step_nums(self): raise NotImplementedError      # [N] int
    def set_goal(self, e, g): raise NotImplementedError
    def set_state(self, e, q, qd, feasible): raise NotImplementedError
    def set_step_num(self, e, k): raise NotImplementedError
    def reset(self, mask=None): raise NotImplementedError   # -> obs [N,9]
    def step(self, actions): raise NotImplementedError      # -> obs, reward, done, terminal_obs
    def violations(self): return None                       # -> (err word, first bad env) or None


def replay(fx, impl: Adaptor):
    N, T = fx["N"], fx["T"]
    assert np.array_equal(impl.goals(), fx["init_goal"]), "goal drawn at construction"
    assert np.array_equal(impl.reset(), fx["reset_obs"]), "initial reset observation"
    by_step = {}
    for row in fx["events"]:
        by_step.setdefault(int(row[0]), []).append(row)
    worst = 0.0
    for t in range(T):
        for row in by_step.get(t, []):
            e, kind, p = int(row[1]), int(row[2]), row[3:].astype(np.float32)
            J = fx["J"]
            if kind == EV_SET_GOAL:
                impl.set_goal(e, p[0:J])
            elif kind == EV_SET_STATE:
                impl.set_state(e, p[0:J], p[J:2 * J], bool(p[2 * J]))
            else:
                impl.set_step_num(e, int(p[2 * J]))
        obs, reward, done, term = impl.step(fx["actions"][t])
        ok = fx["valid"][t]   # envs whose reference instance has not raised so far
        assert np.array_equal(done[ok], fx["done"][t][ok]), "done mask, step %d" % t
        assert np.array_equal(obs[ok], fx["obs"][t][ok]), "observations, step %d" % t
        d = ok & fx["done"][t]
        if fx["auto_reset"] and term is not None and d.any():
            assert np.array_equal(term[d], fx["terminal_obs"][t][d]), "terminal observations, step %d" % t
        want = fx["reward"][t][ok]
        rel = np.abs(reward[ok].astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)
        worst = max(worst, float(rel.max()) if rel.size else 0.0)
        assert rel.size == 0 or rel.max() <= REWARD_RTOL, "reward, step %d: rel err %g" % (t, rel.max())
        if not fx["auto_reset"] and fx["done"][t].any():
            ro = impl.reset(mask=fx["done"][t])
            m = fx["done"][t] & ok
            assert np.array_equal(ro[m], fx["reset_obs_after"][t][m]), "reset obs after done, step %d" % t
        assert np.array_equal(impl.goals()[ok], fx["goal_after"][t][ok]), "goal state, step %d" % t
        assert np.array_equal(impl.step_nums()[ok], fx["step_num_after"][t][ok]), "step_num, step %d" % t
        # an env's reference instance raised AssertionError at this step -> we must have flagged it
        newly = np.flatnonzero(fx["raised_at"] == t)
        if newly.size and impl.violations() is not None:
            word, first = impl.violations()
            assert word & 2, "reward-range violation (roboy_env.py:109) not flagged at step %d" % t
    return worst
