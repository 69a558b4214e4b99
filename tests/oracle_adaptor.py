import numpy as np

from golden_replay import Adaptor
from oracle import oracle as orc


class OracleAdaptor(Adaptor):
    def __init__(self, fx, **kw):
        kw = dict(kw, **(fx.get("bounds") or {}))
        self.env = orc.OracleEnv(fx["N"], seed=fx["seed"], joint_vel_penalty=fx["joint_vel_penalty"],
                                 bonus=fx["bonus"], auto_reset=fx["auto_reset"], **kw)
        self.J = self.env.J

    def goals(self):
        return self.env.goal.T.copy()

    def step_nums(self):
        return self.env.step_num

    def set_goal(self, e, g):
        self.env.goal[:, e] = g

    def set_state(self, e, q, qd, feasible):
        self.env.held[0:self.J, e] = q
        self.env.held[self.J:2 * self.J, e] = qd
        sf = int(self.env.step_flags[e]) & orc.STEP_MASK
        self.env.step_flags[e] = sf | (0 if feasible else orc.F_HELD_INFEASIBLE)

    def set_step_num(self, e, k):
        self.env.step_flags[e] = (int(self.env.step_flags[e]) & ~orc.STEP_MASK) | int(k)

    def reset(self, mask=None):
        return self.env.reset(mask)

    def step(self, actions):
        return self.env.step(actions, want_terminal_obs=True)

    def violations(self):
        return self.env.errors()
