"""Runs the UNMODIFIED reference (gym-roboy) on injected draws -- TEST INFRASTRUCTURE.

Needs the reference -- /root/reference (read-only, build container) or the unmodified copy under
baseline/_ref/ that __graft_entry__.build() installs with pip (git-ignored; it travels to the GPU
box) -- plus the test-only import shim under tests/_shim (gym / rclpy stand-ins; see
tests/_shim/README.md).  It is how the oracle is pinned and how tests/golden/*.npz are generated
(oracle/gen_golden.py).  Never imported by the product package.

The reference's random draws (gym Box.sample) are unpinned, so parity is established by
INJECTION: a `ReplayRobot(MsjRobot)` overrides only `new_random_state()` to return the
Philox draws the CUDA path / oracle would make for that (env id, call counter, stream).  All
other code that runs -- RoboyEnv.step/reset/compute_reward/_did_reach_goal,
StubSimulationClient, RoboyRobot.normalize_state, MsjRobot's spaces -- is the reference's own.
The vec-env worker's reset-on-done (stable-baselines SubprocVecEnv, external to the reference)
is restated in `ReferenceVecEnv.step`.
"""
import contextlib
import io
import os
import sys

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference():
    """The read-only checkout in the build container, else the unmodified copy __graft_entry__.build() installs
    under baseline/_ref/ (git-ignored; it travels to the GPU box with the tree, /root/reference does not)."""
    for cand in (os.environ.get("ROBOY_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "gym_roboy")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference()
_SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "_shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_roboy"))


def _import_reference():
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    with contextlib.redirect_stdout(io.StringIO()):
        from gym_roboy.envs import RoboyEnv
        from gym_roboy.envs.robots import MsjRobot, RobotState
        from gym_roboy.envs.simulations import StubSimulationClient
    return RoboyEnv, MsjRobot, RobotState, StubSimulationClient


def _custom_reference_robot(bounds):
    """A robot plug-in written against the REFERENCE's RoboyRobot base: any joint / tendon count, scalar or
    per-component bounds (roboy_robot.py:21-33 -- what "Upper Body, etc." of README.md:6-7 would subclass)."""
    from gym import spaces
    from gym_roboy.envs.robots import RoboyRobot
    from oracle import oracle as orc
    J, A, _, b = orc.robot_bounds(bounds)

    class CustomRobot(RoboyRobot):
        _A = spaces.Box(low=b["angle_low"], high=b["angle_high"], dtype="float32")
        _V = spaces.Box(low=b["vel_low"], high=b["vel_high"], dtype="float32")
        _T = spaces.Box(low=b["act_low"], high=b["act_high"], dtype="float32")

        @classmethod
        def get_action_space(cls):
            return cls._T

        @classmethod
        def get_joint_angles_space(cls):
            return cls._A

        @classmethod
        def get_joint_vels_space(cls):
            return cls._V

    assert CustomRobot._A.shape == (J,) and CustomRobot._T.shape == (A,)
    return CustomRobot


def _make_replay_robot(MsjRobot, RobotState, state_fn, goal_fn):
    """A fresh MsjRobot subclass whose only override is the source of random samples."""

    class ReplayRobot(MsjRobot):
        ctx = dict(gid=0, t=0, goal_real=True)
        log = []

        @classmethod
        def new_random_state(cls):
            caller = sys._getframe(1).f_code.co_name
            gid, t = cls.ctx["gid"], cls.ctx["t"]
            if caller == "get_new_goal_joint_angles":
                cls.log.append("goal")
                J = cls.get_joint_angles_space().shape[0]
                else_goal = np.zeros(J, np.float32)   # goal that the worker's reset() overwrites at once: never observable
                g = goal_fn(gid, t) if cls.ctx["goal_real"] else else_goal
                return RobotState(joint_angles=g, joint_vels=np.zeros(J, np.float32), is_feasible=True)
            assert caller in ("forward_step_command", "__init__"), caller
            cls.log.append("state")
            q, qd = state_fn(gid, t)
            return RobotState(joint_angles=q, joint_vels=qd, is_feasible=True)

    return ReplayRobot()


class ReferenceVecEnv:
    """N reference RoboyEnv(StubSimulationClient(ReplayRobot)) instances in lock-step."""

    def __init__(self, n_envs, seed=1234, env_id_base=0, joint_vel_penalty=False, bonus=True, auto_reset=True,
                 bounds=None):
        from oracle import oracle as orc

        RoboyEnv, MsjRobot, RobotState, Stub = _import_reference()
        low, high, J = orc.MSJ["angle_low"], orc.MSJ["angle_high"], 3
        if bounds is not None:   # another robot (a RoboyRobot subclass of the reference): other dims and / or spaces
            MsjRobot = _custom_reference_robot(bounds)
            J, _, _, b = orc.robot_bounds(bounds)
            low, high = b["angle_low"], b["angle_high"]
        self.J, self.obs_dim = J, 3 * J
        self.RobotState = RobotState
        self.n = n_envs
        self.seed = seed
        self.auto_reset = auto_reset
        self.t = 0
        self.envs, self.robots = [], []

        def state_fn(gid, t):
            q, qd = orc.draw_state(seed, [gid], t, low, high, J=J)
            return q[0], qd[0]

        def goal_fn(gid, t):
            return orc.draw_goal(seed, [gid], t, low, high, J=J)[0]

        with contextlib.redirect_stdout(io.StringIO()):
            for i in range(n_envs):
                robot = _make_replay_robot(MsjRobot, RobotState, state_fn, goal_fn)
                type(robot).ctx = dict(gid=env_id_base + i, t=0, goal_real=True)
                type(robot).log = []
                env = RoboyEnv(Stub(robot=robot), joint_vel_penalty=joint_vel_penalty,
                               is_agent_getting_bonus_for_reaching_goal=bonus)
                assert type(robot).log == ["state", "goal"], type(robot).log  # SURVEY 8a a11 draw order
                self.envs.append(env)
                self.robots.append(robot)
        self.gid0 = env_id_base
        self.reward_range = self.envs[0].reward_range

    # ---- injection (the batched API's set_state / set_goal / set_step_num) ----
    def set_goal(self, i, goal_q):
        self.envs[i]._set_new_goal(goal_joint_angle=np.asarray(goal_q, np.float32))

    def set_state(self, i, q, qd, feasible=True):
        self.envs[i]._simulation_client._state = self.RobotState(
            joint_angles=np.asarray(q, np.float32), joint_vels=np.asarray(qd, np.float32), is_feasible=bool(feasible))

    def set_step_num(self, i, k):
        self.envs[i].step_num = int(k)

    # ---- observers ----
    def goals(self):
        return np.stack([np.asarray(e._goal_state.joint_angles, np.float32) for e in self.envs])

    def step_nums(self):
        return np.array([e.step_num for e in self.envs], np.int64)

    def _ctx(self, i, goal_real):
        cls = type(self.robots[i])
        cls.ctx = dict(gid=self.gid0 + i, t=self.t, goal_real=goal_real)
        cls.log = []
        return cls

    def reset(self, mask=None):
        self.t += 1
        obs = np.zeros((self.n, self.obs_dim), np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            for i, env in enumerate(self.envs):
                if mask is not None and not mask[i]:
                    continue
                cls = self._ctx(i, True)
                obs[i] = env.reset()
                assert cls.log == ["goal"], cls.log
        return obs

    def step(self, actions):
        """actions float32 [n,8].  Returns obs f64 [n,9], reward f64 [n], done bool [n],
        terminal_obs f64 [n,9], raised [n] (AssertionError text or '')."""
        self.t += 1
        n = self.n
        obs = np.zeros((n, self.obs_dim), np.float64)
        term = np.zeros((n, self.obs_dim), np.float64)
        rew = np.zeros(n, np.float64)
        done = np.zeros(n, bool)
        raised = [""] * n
        with contextlib.redirect_stdout(io.StringIO()):
            for i, env in enumerate(self.envs):
                cls = self._ctx(i, not self.auto_reset)
                try:
                    o, r, d, _ = env.step(np.asarray(actions[i], np.float32))
                except AssertionError as exc:  # roboy_env.py:52 / :109
                    raised[i] = "AssertionError: %s" % (exc,)
                    continue
                assert isinstance(r, float) and isinstance(d, bool)
                rew[i], done[i] = r, d
                if d and self.auto_reset:  # SubprocVecEnv worker: reset obs replaces terminal obs
                    term[i] = o
                    cls.ctx["goal_real"] = True
                    o = env.reset()
                    assert cls.log in (["state", "goal", "goal"], ["goal", "goal"]), cls.log
                obs[i] = o
        return obs, rew, done, term, raised


class ReferenceFeedEnv:
    """N reference RoboyEnv instances over a test-double SimulationClient that is FED its states,
    the way RosSimulationClient receives them from the external simulator
    (ros_simulation_client.py:40-60): python-float lists -> robot.new_state(...) -> float64 arrays.
    Pins the oracle's / kernel's external-simulator path (roboy_step_external)."""

    def __init__(self, n_envs, seed=1234, env_id_base=0, joint_vel_penalty=False, bonus=True):
        from oracle import oracle as orc

        RoboyEnv, MsjRobot, RobotState, _ = _import_reference()
        from gym_roboy.envs.simulations import SimulationClient
        outer = self
        self.n, self.seed, self.gid0, self.t = n_envs, seed, env_id_base, 0

        class FeedClient(SimulationClient):
            def __init__(self, robot, gid):
                self.robot, self.gid, self.next_state = robot, gid, None

            def read_state(self):
                return self.next_state

            def forward_step_command(self, action):
                return self.next_state

            def forward_reset_command(self):
                return self.next_state

            def get_new_goal_joint_angles(self):
                return orc.draw_goal(outer.seed, [self.gid], outer.t)[0]

        self.robot = MsjRobot()
        with contextlib.redirect_stdout(io.StringIO()):
            self.clients = [FeedClient(self.robot, env_id_base + i) for i in range(n_envs)]
            self.envs = [RoboyEnv(c, joint_vel_penalty=joint_vel_penalty,
                                  is_agent_getting_bonus_for_reaching_goal=bonus) for c in self.clients]

    def _feed(self, i, q, qd, feasible):
        # exactly RosSimulationClient._make_robot_state: lists of python floats
        self.clients[i].next_state = self.robot.new_state(
            joint_angle=[float(x) for x in q], joint_vel=[float(x) for x in qd], is_feasible=bool(feasible))

    def goals(self):
        return np.stack([np.asarray(e._goal_state.joint_angles, np.float32) for e in self.envs])

    def set_step_num(self, i, k):
        self.envs[i].step_num = int(k)

    def reset(self, q, qd, mask=None):
        self.t += 1
        obs = np.zeros((self.n, 9))
        with contextlib.redirect_stdout(io.StringIO()):
            for i, env in enumerate(self.envs):
                if mask is not None and not mask[i]:
                    continue
                self._feed(i, q[i], qd[i], True)
                obs[i] = env.reset()
        return obs

    def step(self, q, qd, feasible):
        self.t += 1
        obs, rew, done = np.zeros((self.n, 9)), np.zeros(self.n), np.zeros(self.n, bool)
        with contextlib.redirect_stdout(io.StringIO()):
            for i, env in enumerate(self.envs):
                self._feed(i, q[i], qd[i], feasible[i])
                obs[i], rew[i], done[i], _ = env.step(np.zeros(8, np.float32))
        return obs, rew, done
