from .roboy_robot import RobotState, RoboyRobot  # noqa: F401
from .msj_robot import MsjRobot  # noqa: F401
