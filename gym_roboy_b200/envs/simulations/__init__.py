from .simulation_client import SimulationClient  # noqa: F401
from .cuda_simulation_client import CudaSimulationClient  # noqa: F401
