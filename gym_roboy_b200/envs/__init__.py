from .roboy_env import RoboyEnv  # noqa: F401
