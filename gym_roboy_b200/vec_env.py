"""A stable-baselines-shaped `VecEnv` over the batched env, for trainers that keep numpy on the host.

`train_parallel.py:29` wraps one reference env per process in `SubprocVecEnv`.  This adapter offers
the same calling convention -- `reset()`, `step_async(actions)` / `step_wait()`, `step(actions)`,
`num_envs`, spaces, `infos[i]['terminal_observation']` on done, per-env `get_attr` / `set_attr` /
`env_method` -- on top of ONE fused kernel launch per step.  Data crosses PCIe once per step through
page-locked buffers (`roboy_step_host`); which envs finished, and their pre-reset observations, are
compacted on the device (`roboy_done_indices`), so the host never scans the done mask.
stable-baselines is external to the reference; the contract followed is its
documented worker behaviour: on done the env is reset and the reset observation is returned.

Actions are NOT clipped here: stable-baselines' runner clips before `env.step`, and the reference env
asserts (`roboy_env.py:52`).  An out-of-range or NaN action raises the same `AssertionError` from
`step_wait()` (the device error word), unless the adapter is built with `clip_actions=True`.
"""
import numpy as np

from .envs import RoboyEnv
from .envs.simulations import CudaSimulationClient

# attributes of RoboyEnv that differ per env: get_attr returns one value per selected env
_PER_ENV_ATTRS = ("step_num",)


class RoboyVecEnv:
    def __init__(self, num_envs, seed=0, device=None, env_id_base=0, terminal_observation=True, clip_actions=False,
                 robot=None, **env_kwargs):
        client = CudaSimulationClient(robot=robot, num_envs=num_envs, seed=seed, device=device, env_id_base=env_id_base)
        self.env = RoboyEnv(client, auto_reset=True, strict=False, **env_kwargs)
        if num_envs == 1:
            self.env._single = False   # a VecEnv is batched even with one env
        self.client = client
        self.num_envs = num_envs
        self.observation_space, self.action_space = self.env.observation_space, self.env.action_space
        self.reward_range = self.env.reward_range
        self._actions, self._obs, self._rew, self._done = client.host_buffers()
        self._clip = bool(clip_actions)
        self._want_terminal = terminal_observation
        if terminal_observation:
            client.enable_terminal_obs(True)
            client.enable_done_index(True)
        self._pending = False

    def seed(self, seed=None):
        self.env.seed(seed)

    def reset(self):
        return self.env.reset().cpu().numpy().reshape(self.num_envs, self.client.dim_obs)

    def step_async(self, actions):
        a = np.asarray(actions, np.float32).reshape(self.num_envs, self.client.dim_action)
        if self._clip:
            np.clip(a, -1.0, 1.0, out=self._actions)
        else:
            self._actions[...] = a
        self._pending = True

    def step_wait(self):
        assert self._pending, "step_async() first"
        self._pending = False
        self.client.step_host(self._actions, self._obs, self._rew, self._done)
        self.env.check_errors()   # roboy_env.py:52 / :109 -> AssertionError, as the worker's env.step would raise
        infos = [{} for _ in range(self.num_envs)]
        if self._want_terminal:
            idx, rows = self.client.done_indices(with_terminal_obs=True)   # device-side compaction + gather
            if idx.numel():
                for i, row in zip(idx.cpu().numpy(), rows.cpu().numpy()):
                    infos[int(i)]["terminal_observation"] = row
        return self._obs.copy(), self._rew.copy(), self._done.astype(bool), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # ---- per-env accessors (stable-baselines VecEnv.get_attr / set_attr / env_method) ----
    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, (int, np.integer)):
            return [int(indices)]
        return [int(i) for i in indices]

    def get_attr(self, name, indices=None):
        """One value PER selected env, as SubprocVecEnv returns them."""
        idx = self._indices(indices)
        value = getattr(self.env, name)
        if name in _PER_ENV_ATTRS:
            per_env = value.cpu().numpy()
            return [int(per_env[i]) for i in idx]
        return [value for _ in idx]

    def set_attr(self, name, value, indices=None):
        idx = self._indices(indices)
        if name == "step_num":
            self.client.set_step_num(np.full(len(idx), int(value), np.int32), idx=idx)
        else:
            setattr(self.env, name, value)

    def env_method(self, name, *args, indices=None, **kwargs):
        """Call `name` once for the batch and hand every selected env its own result (batched results are split)."""
        idx = self._indices(indices)
        out = getattr(self.env, name)(*args, **kwargs)
        if hasattr(out, "shape") and len(getattr(out, "shape", ())) >= 1 and out.shape[0] == self.num_envs:
            out = out.cpu().numpy() if hasattr(out, "cpu") else np.asarray(out)
            return [out[i] for i in idx]
        return [out for _ in idx]

    def close(self):
        self.env.close()
