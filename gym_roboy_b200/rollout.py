"""PPO-style rollout collection on the GPU -- the caller of the env hot path.

The reference trains with stable-baselines' PPO2 over a `SubprocVecEnv`
(gym_roboy/train_parallel.py:28-35): per step the policy runs in the trainer process, the clipped
action crosses a pipe to each env process and the observation comes back the same way.  Here the
whole loop stays on one device: a torch policy reads the observation tensor the step kernel wrote,
actions are clipped on the device, and the step kernel writes observation / reward / done
straight into the `[T, N]` rollout buffers (its output pointers ARE the buffer slots).
Advantages come from the `roboy_gae` kernel.  Because the env's call counter is device state, the
entire T-step loop can be captured in one CUDA graph and replayed, which is what makes small
populations (e.g. 4,096 envs), otherwise launch-bound, run at speed.

stable-baselines itself is not part of the reference repo (parity unpinned); the
policy shape and the GAE recurrence follow its PPO2 `MlpPolicy` / runner as documented upstream:
separate 64-64 tanh networks for policy and value, state-independent log-std, actions clipped to
the action space before `env.step`.
"""
import ctypes
import math

import torch

from . import _native


class MlpPolicy(torch.nn.Module):
    def __init__(self, obs_dim=9, act_dim=8, hidden=64):
        super().__init__()
        def net(out):
            return torch.nn.Sequential(torch.nn.Linear(obs_dim, hidden), torch.nn.Tanh(),
                                       torch.nn.Linear(hidden, hidden), torch.nn.Tanh(), torch.nn.Linear(hidden, out))
        self.pi, self.vf = net(act_dim), net(1)
        self.log_std = torch.nn.Parameter(torch.zeros(act_dim))

    def forward(self, obs):
        return self.pi(obs), self.vf(obs).squeeze(-1)


def gae(rewards, values, dones, last_value, gamma=0.99, lam=0.95, adv=None, ret=None):
    """GAE(lambda) on the device via the `roboy_gae` kernel.  rewards/values float32 [T,N], dones
    uint8/bool [T,N] (done[t]: the episode ended at step t), last_value float32 [N]."""
    T, n = rewards.shape
    adv = torch.empty_like(rewards) if adv is None else adv
    ret = torch.empty_like(rewards) if ret is None else ret
    d = dones.view(torch.uint8) if dones.dtype == torch.bool else dones
    for t in (rewards, values, d, last_value, adv, ret):
        assert t.is_cuda and t.is_contiguous()
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    stream = ctypes.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    _native.check(_native.load().roboy_gae(T, n, p(rewards), p(values), p(d), p(last_value), float(gamma), float(lam),
                                           p(adv), p(ret), stream))
    return adv, ret


class RolloutCollector:
    """Collects `[T, N]` rollouts from a batched `RoboyEnv` with a torch policy, all on the device."""

    def __init__(self, env, policy, n_steps=128, gamma=0.99, lam=0.95):
        self.env, self.client, self.policy = env, env._simulation_client, policy
        self.T, self.N = int(n_steps), env.num_envs
        self.gamma, self.lam = gamma, lam
        dev, T, N = self.client.device, self.T, self.N
        f32 = dict(dtype=torch.float32, device=dev)
        self.obs = torch.zeros((T + 1, N, 9), **f32)
        self.actions = torch.zeros((T, N, 8), **f32)
        self.logp = torch.zeros((T, N), **f32)
        self.values = torch.zeros((T + 1, N), **f32)
        self.rewards = torch.zeros((T, N), **f32)
        self.dones = torch.zeros((T, N), dtype=torch.uint8, device=dev)
        self.adv = torch.zeros((T, N), **f32)
        self.ret = torch.zeros((T, N), **f32)
        self._graph = None
        self.obs[0].copy_(env.reset())

    def _loop(self):
        policy, client = self.policy, self.client
        std = policy.log_std.exp()
        log_norm = -0.5 * math.log(2 * math.pi) * std.numel() - policy.log_std.sum()
        for t in range(self.T):
            mean, value = policy(self.obs[t])
            noise = torch.randn_like(mean)
            self.logp[t] = log_norm - 0.5 * (noise * noise).sum(-1)
            torch.clamp(mean + std * noise, -1.0, 1.0, out=self.actions[t])   # the runner's np.clip, on the device
            self.values[t] = value
            # zero-copy: the step kernel's output pointers are the rollout buffer slots
            client.step_fused(self.actions[t], obs=self.obs[t + 1], reward=self.rewards[t], done=self.dones[t])
        self.values[self.T] = policy(self.obs[self.T])[1]
        gae(self.rewards, self.values[: self.T], self.dones, self.values[self.T], self.gamma, self.lam, self.adv, self.ret)
        self.obs[0].copy_(self.obs[self.T])   # next rollout continues where this one stopped

    @torch.no_grad()
    def collect(self):
        """One rollout of T steps (replays the captured CUDA graph if `capture()` was called)."""
        if self._graph is not None:
            self._graph.replay()
        else:
            self._loop()
        return self

    @torch.no_grad()
    def capture(self, warmup=1):
        """Capture the whole T-step loop (policy, clip, env step, GAE) in one CUDA graph."""
        side = torch.cuda.Stream(device=self.client.device)
        side.wait_stream(torch.cuda.current_stream(self.client.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._loop()
        torch.cuda.current_stream(self.client.device).wait_stream(side)
        torch.cuda.synchronize(self.client.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._loop()
        self._graph = graph
        return self
