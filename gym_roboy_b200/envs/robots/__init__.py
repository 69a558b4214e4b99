"""Robot plug-ins: the value type shared with simulation clients, the plug-in base class and
the MSJ platform (3 joint angles, 8 tendons).  A robot is just three bounded spaces."""
from .msj_robot import MsjRobot
from .roboy_robot import RoboyRobot, RobotState

__all__ = ["MsjRobot", "RoboyRobot", "RobotState"]
