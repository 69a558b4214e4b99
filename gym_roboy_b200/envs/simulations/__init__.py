"""Simulation clients: the four-call plug-in interface an env talks to, and the in-process CUDA
implementation (`CudaSimulationClient`) that holds every env's state in HBM.  No ROS client
here: the external-simulator case is served by `RoboyEnv.step_from_states`."""
from .cuda_simulation_client import CudaSimulationClient
from .simulation_client import SimulationClient

__all__ = ["CudaSimulationClient", "SimulationClient"]
