"""A stable-baselines-shaped `VecEnv` over the batched env, for trainers that keep numpy on the host.

`train_parallel.py:29` wraps one reference env per process in `SubprocVecEnv`.  This adapter offers
the same calling convention -- `reset()`, `step_async(actions)` / `step_wait()`, `step(actions)`,
`num_envs`, spaces, `infos[i]['terminal_observation']` on done -- on top of ONE fused kernel launch
per step.  Data crosses PCIe once per step through pinned buffers (`roboy_step_host`).
stable-baselines is external to the reference; the contract followed is its
documented worker behaviour: on done the env is reset and the reset observation is returned.
"""
import numpy as np
import torch

from .envs import RoboyEnv
from .envs.simulations import CudaSimulationClient


class RoboyVecEnv:
    def __init__(self, num_envs, seed=0, device=None, env_id_base=0, terminal_observation=True, **env_kwargs):
        client = CudaSimulationClient(num_envs=num_envs, seed=seed, device=device, env_id_base=env_id_base)
        self.env = RoboyEnv(client, auto_reset=True, strict=False, **env_kwargs)
        if num_envs == 1:
            self.env._single = False   # a VecEnv is batched even with one env
        self.client = client
        self.num_envs = num_envs
        self.observation_space, self.action_space = self.env.observation_space, self.env.action_space
        self.reward_range = self.env.reward_range
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
        self._actions = pin((num_envs, 8), torch.float32)
        self._obs = pin((num_envs, 9), torch.float32)
        self._rew = pin((num_envs,), torch.float32)
        self._done = pin((num_envs,), torch.uint8)
        self._want_terminal = terminal_observation
        if terminal_observation:
            client.enable_terminal_obs(True)
        self._pending = False

    def seed(self, seed=None):
        self.env.seed(seed)

    def reset(self):
        return self.env.reset().cpu().numpy().reshape(self.num_envs, 9)

    def step_async(self, actions):
        np.clip(np.asarray(actions, np.float32).reshape(self.num_envs, 8), -1.0, 1.0, out=self._actions)
        self._pending = True

    def step_wait(self):
        assert self._pending, "step_async() first"
        self._pending = False
        self.client.step_host(self._actions, self._obs, self._rew, self._done)
        done = self._done.astype(bool)
        infos = [{} for _ in range(self.num_envs)]
        if self._want_terminal and done.any():
            idx = np.flatnonzero(done)
            term = self.client.terminal_obs[torch.as_tensor(idx, device=self.client.device)].cpu().numpy()
            for i, row in zip(idx, term):
                infos[i]["terminal_observation"] = row
        return self._obs.copy(), self._rew.copy(), done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_attr(self, name, indices=None):
        return [getattr(self.env, name)] * (self.num_envs if indices is None else len(indices))

    def env_method(self, name, *args, indices=None, **kwargs):
        return [getattr(self.env, name)(*args, **kwargs)]

    def close(self):
        self.env.close()
