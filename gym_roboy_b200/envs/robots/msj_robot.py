"""The MSJ platform: 3 joint angles, 8 tendons (reference: envs/robots/msj_robot.py:6-28)."""
import numpy as np

from ...spaces import Box
from .roboy_robot import RoboyRobot


class MsjRobot(RoboyRobot):
    DIM_JOINT_ANGLE = 3        # msj_robot.py:8
    DIM_ACTION = 8             # msj_robot.py:12
    MAX_TENDON_LENGTH = 0.3    # msj_robot.py:15 (cm)
    MAX_JOINT_ANGLE = np.pi    # msj_robot.py:9
    MAX_JOINT_VEL = np.pi / 6  # msj_robot.py:10

    _spaces = {
        "angle": Box(-MAX_JOINT_ANGLE, MAX_JOINT_ANGLE, (DIM_JOINT_ANGLE,), "float32"),
        "vel": Box(-MAX_JOINT_VEL, MAX_JOINT_VEL, (DIM_JOINT_ANGLE,), "float32"),
        "action": Box(-MAX_TENDON_LENGTH, MAX_TENDON_LENGTH, (DIM_ACTION,), "float32"),
    }

    @classmethod
    def get_action_space(cls) -> Box:
        return cls._spaces["action"]

    @classmethod
    def get_joint_angles_space(cls) -> Box:
        return cls._spaces["angle"]

    @classmethod
    def get_joint_vels_space(cls) -> Box:
        return cls._spaces["vel"]
