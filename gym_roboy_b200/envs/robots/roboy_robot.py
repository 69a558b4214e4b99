"""`RobotState` and the `RoboyRobot` plug-in base class.

Mirrors the reference's interface (envs/robots/roboy_robot.py:6-95) -- same names, arguments
and error behaviour -- so a robot written for gym-roboy plugs in here and vice versa.  A robot
is described by three bounded spaces; subclasses only declare them (`MsjRobot`).  The values
held by a `RobotState` may be numpy arrays (one env) or torch tensors `[N,3]` (batched).
"""
import numpy as np

from ...spaces import Box


def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch"


class RobotState:
    """Joint angles, joint velocities and the simulator's feasibility verdict (roboy_robot.py:6-11)."""

    __slots__ = ("joint_angles", "joint_vels", "is_feasible")

    def __init__(self, joint_angles, joint_vels, is_feasible):
        if not (isinstance(is_feasible, (bool, np.bool_)) or _is_torch(is_feasible)):
            raise TypeError("is_feasible must be a bool (or a bool tensor for batched states)")
        self.joint_angles = joint_angles if _is_torch(joint_angles) else np.array(joint_angles)
        self.joint_vels = joint_vels if _is_torch(joint_vels) else np.array(joint_vels)
        self.is_feasible = bool(is_feasible) if isinstance(is_feasible, (bool, np.bool_)) else is_feasible

    @classmethod
    def interpolate(cls, state1, state2):
        """Midpoint of two states (roboy_robot.py:13-18)."""
        assert isinstance(state1, cls) and isinstance(state2, cls)
        both = state1.is_feasible & state2.is_feasible if _is_torch(state1.is_feasible) \
            else (state1.is_feasible and state2.is_feasible)
        return cls((state1.joint_angles + state2.joint_angles) / 2, (state1.joint_vels + state2.joint_vels) / 2, both)

    def __repr__(self):
        return "RobotState(q={}, qd={}, feasible={})".format(self.joint_angles, self.joint_vels, self.is_feasible)


class RoboyRobot:
    """Base of the robot plug-ins (roboy_robot.py:21-95)."""

    @classmethod
    def get_action_space(cls) -> Box:
        raise NotImplementedError

    @classmethod
    def get_joint_angles_space(cls) -> Box:
        raise NotImplementedError

    @classmethod
    def get_joint_vels_space(cls) -> Box:
        raise NotImplementedError

    # ---- host-side state factories (test helpers in the reference; not on the device path) ----
    @classmethod
    def _zeros(cls, space):
        return np.zeros(space.shape)  # float64, as roboy_robot.py:43-44

    @classmethod
    def new_random_state(cls) -> RobotState:
        # reference quirk kept: the velocities are sampled from the ANGLE space (roboy_robot.py:38)
        angles = cls.get_joint_angles_space()
        return RobotState(angles.sample(), angles.sample(), True)

    @classmethod
    def new_zero_state(cls) -> RobotState:
        return RobotState(cls._zeros(cls.get_joint_angles_space()), cls._zeros(cls.get_joint_vels_space()), True)

    @classmethod
    def new_random_zero_vels_state(cls) -> RobotState:
        return RobotState(cls.get_joint_angles_space().sample(), cls._zeros(cls.get_joint_vels_space()), True)

    @classmethod
    def new_random_zero_angles_state(cls) -> RobotState:
        return RobotState(cls._zeros(cls.get_joint_angles_space()), cls.get_joint_vels_space().sample(), True)

    @classmethod
    def new_max_state(cls) -> RobotState:
        return RobotState(cls.get_joint_angles_space().high, cls.get_joint_vels_space().high, False)

    @classmethod
    def new_min_state(cls) -> RobotState:
        return RobotState(cls.get_joint_angles_space().low, cls.get_joint_vels_space().low, False)

    @classmethod
    def new_state(cls, joint_angle, joint_vel, is_feasible) -> RobotState:
        """roboy_robot.py:71-78: angles must lie in the (closed) angle space; velocities are not checked."""
        if not isinstance(is_feasible, (bool, np.bool_)):
            raise TypeError("is_feasible must be a bool")
        if not isinstance(joint_angle, np.ndarray):
            joint_angle = np.array(joint_angle)
        if not isinstance(joint_vel, np.ndarray):
            joint_vel = np.array(joint_vel)
        assert cls.get_joint_angles_space().contains(joint_angle), joint_angle
        return RobotState(joint_angle, joint_vel, bool(is_feasible))

    # ---- normalisation to [-1, 1] (roboy_robot.py:80-95); host version of the device formula ----
    @staticmethod
    def _normalize_between_minus1_and1(val, max_val, min_val):
        return (2 * val - max_val - min_val) / (max_val - min_val)

    def normalize_state(self, state: RobotState) -> RobotState:
        angles, vels = self.get_joint_angles_space(), self.get_joint_vels_space()
        return RobotState(
            self._normalize_between_minus1_and1(state.joint_angles, angles.high, angles.low),
            self._normalize_between_minus1_and1(state.joint_vels, vels.high, vels.low),
            state.is_feasible,
        )
