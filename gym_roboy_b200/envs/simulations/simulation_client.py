"""`SimulationClient`: how a RoboyEnv talks to whatever simulates the robot.

Same four calls and `robot` attribute as the reference's plug-in interface
(envs/simulations/simulation_client.py:6-23).  Implementations here: `CudaSimulationClient`.
"""
from ..robots import RobotState, RoboyRobot


class SimulationClient:
    robot = RoboyRobot()

    def read_state(self) -> RobotState:
        raise NotImplementedError

    def forward_step_command(self, action) -> RobotState:
        raise NotImplementedError

    def forward_reset_command(self) -> RobotState:
        raise NotImplementedError

    def get_new_goal_joint_angles(self):
        raise NotImplementedError
