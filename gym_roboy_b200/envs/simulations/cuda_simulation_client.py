"""`CudaSimulationClient`: the in-process simulation client, batched on one B200.

It plays the role `StubSimulationClient` plays in the reference
(envs/simulations/simulation_client.py:26-47) -- the ROS-free, in-process client the unit tests
use -- for `num_envs` environments at once, with all state resident in HBM:

* the four `SimulationClient` calls (`read_state`, `forward_step_command`,
  `forward_reset_command`, `get_new_goal_joint_angles`) run as (un-fused) kernels, so the
  *reference's own* `RoboyEnv` can sit on top of this client unchanged (`num_envs=1`);
* `step_fused` / `reset_fused` run the whole `RoboyEnv.step` / `reset` in one kernel; the
  batched `RoboyEnv` of this package drives those.

Everything goes through the C-ABI in include/roboy_b200.h; there is no CPU path.
"""
import ctypes
import os

import numpy as np
import torch

from ... import _native

_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)   # (device index) -> cudaStream_t as int
from ..robots import MsjRobot, RobotState, RoboyRobot
from .simulation_client import SimulationClient


def _fill_bounds(cfg, robot):
    """The robot's three spaces (roboy_robot.py:23-33) into struct roboy_cfg: dims, and one bound per component when a
    space is not uniform.  Returns (dim_joint, dim_action)."""
    angles, vels, acts = robot.get_joint_angles_space(), robot.get_joint_vels_space(), robot.get_action_space()
    if len(angles.shape) != 1 or angles.shape != vels.shape or len(acts.shape) != 1:
        raise ValueError("joint angle / velocity spaces must be 1-D and of equal length, the action space 1-D")
    J, A = int(angles.shape[0]), int(acts.shape[0])
    if not (1 <= J <= _native.MAX_JOINT and 1 <= A <= _native.MAX_ACTION):
        raise ValueError("the CUDA path takes robots with 1..{} joints and 1..{} tendons (numpy sums longer vectors in a "
                         "CPU-dependent order, so parity with the reference is undefined beyond)".format(
                             _native.MAX_JOINT, _native.MAX_ACTION))
    sides = dict(angle_low=angles.low, angle_high=angles.high, vel_low=vels.low, vel_high=vels.high,
                 act_low=acts.low, act_high=acts.high)
    sides = {k: np.asarray(v, np.float32).reshape(-1) for k, v in sides.items()}
    cfg.dim_joint, cfg.dim_action = J, A
    cfg.per_component_bounds = int(any(np.unique(v).size != 1 for v in sides.values()))
    for name, v in sides.items():
        setattr(cfg, name, float(v[0]))
        arr = getattr(cfg, name + "_v")
        for k, x in enumerate(v):
            arr[k] = float(x)
    return J, A


class CudaSimulationClient(SimulationClient):

    def __init__(self, robot: RoboyRobot = None, num_envs: int = 1, device=None, seed: int = None,
                 env_id_base: int = 0):
        self.robot = robot if robot is not None else MsjRobot()
        self.num_envs = int(num_envs)
        if self.num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        self._lib = _native.load()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cuda:0"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _native.RoboyNativeError("CudaSimulationClient needs a CUDA device; there is no CPU fallback")
        self._device_index = self.device.index if self.device.index is not None else (
            torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.seed_value = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & (2 ** 64 - 1)
        self.env_id_base = int(env_id_base)

        cfg = _native.RoboyCfg()
        _native.check(self._lib.roboy_cfg_msj(ctypes.byref(cfg)))
        cfg.n_envs, cfg.env_id_base, cfg.seed = self.num_envs, self.env_id_base, self.seed_value
        self.dim_joint, self.dim_action = _fill_bounds(cfg, self.robot)
        self.dim_obs = 3 * self.dim_joint
        self._cfg = cfg
        handle = ctypes.c_void_p()
        _native.check(self._lib.roboy_create(ctypes.byref(cfg), self.device.index or 0, ctypes.byref(handle)))
        self._h = handle
        self._roboy_step = self._lib.roboy_step   # bound once: the eager hot path

        # zero-copy torch views of the HBM buffers the handle owns (DLPack)
        self.goal = self._view(_native.BUF_GOAL)                # float32 [J, N]
        self.step_flags = self._view(_native.BUF_STEP_FLAGS)    # int32   [N]
        self.held = self._view(_native.BUF_HELD)                # float32 [2J, N]
        self.obs = self._view(_native.BUF_OBS)                  # float32 [N, 3J]
        self.reward = self._view(_native.BUF_REWARD)            # float32 [N]
        self.done_u8 = self._view(_native.BUF_DONE)             # uint8   [N]
        self.done = self.done_u8.view(torch.bool)
        self.stats_tensor = self._view(_native.BUF_STATS)       # float64 [8]
        self.terminal_obs = None
        self.info_sink = None
        self._all_idx = None
        self._host_allocs = []
        self._done_index, self._done_idx, self._done_count, self._done_rows = False, None, None, None

    # ------------------------------------------------------------------ plumbing
    def _view(self, which):
        ptr = self._lib.roboy_export_dlpack(self._h, which)
        return torch.from_dlpack(_native.dlpack_capsule(ptr))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        destroy = getattr(getattr(self, "_lib", None), "roboy_destroy", None)
        if h and destroy is not None:   # (None during interpreter shutdown)
            destroy(h)
            for ptr in getattr(self, "_host_allocs", []):
                self._lib.roboy_host_free(ctypes.c_void_p(ptr))
            self._host_allocs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        # torch's current stream on this device as a raw handle.  The private accessor skips building a Stream object
        # (1.5 us of a 9 us eager step at launch-bound sizes); the public one is the fall-back.
        if _RAW_STREAM is not None:
            return _RAW_STREAM(self._device_index)
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, x, dtype, shape=None):
        t = torch.as_tensor(x, dtype=dtype, device=self.device).contiguous()
        if shape is not None:
            t = t.reshape(shape)
        return t

    def _idx(self, idx):
        if idx is None:
            if self._all_idx is None:
                self._all_idx = torch.arange(self.num_envs, dtype=torch.int64, device=self.device)
            return self._all_idx
        return self._dev(idx, torch.int64).reshape(-1)

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()   # argtypes are c_void_p: ctypes converts the int

    def _state_out(self, q, qd, feasible):
        # feasible byte: bit 0 is_feasible, bit 1 "this is the float64 zero state" (roboy_robot.py:41-45)
        if self.num_envs == 1:
            f = int(feasible[0].item())
            dt = np.float64 if f & 2 else np.float32   # numpy then promotes as it does over the reference's Stub
            return RobotState(q[0].cpu().numpy().astype(dt), qd[0].cpu().numpy().astype(dt), bool(f & 1))
        return RobotState(q, qd, (feasible & 1).view(torch.bool))

    # ------------------------------------------------------------------ SimulationClient API
    @property
    def msj_kernels(self):
        """True when the tuned MSJ kernels run this robot's fused step (MSJ's dims, uniform bounds), else the generic ones."""
        m = ctypes.c_int()
        _native.check(self._lib.roboy_robot_dims(self._h, None, None, None, ctypes.byref(m)))
        return bool(m.value)

    @property
    def fast_division(self):
        """True when the fused step divides by the robot's spans with the proved three-instruction core (same results as
        IEEE division: offline proof for MSJ, exhaustive check on this device at construction for any other robot)."""
        m = ctypes.c_int()
        _native.check(self._lib.roboy_fast_division(self._h, ctypes.byref(m)))
        return bool(m.value)

    @property
    def penalty_float32(self):
        """True when the fused step evaluates the velocity penalty (roboy_env.py:98-100) of sampled states in float32 with
        a float64 re-run next to the bounds of reward_range (same error word, rewards within 1e-6 of the reference)."""
        m = ctypes.c_int()
        _native.check(self._lib.roboy_penalty_float32(self._h, ctypes.byref(m)))
        return bool(m.value)

    def read_state(self) -> RobotState:
        """simulation_client.py:33-34"""
        n = self.num_envs
        q = torch.empty((n, self.dim_joint), dtype=torch.float32, device=self.device)
        qd = torch.empty_like(q)
        f = torch.empty(n, dtype=torch.uint8, device=self.device)
        _native.check(self._lib.roboy_read_state(self._h, n, self._p(self._idx(None)), self._p(q), self._p(qd),
                                                 self._p(f), self._stream()))
        return self._state_out(q, qd, f)

    def forward_step_command(self, action) -> RobotState:
        """simulation_client.py:36-40.  `action`: 8 floats in robot units (or a `[N,8]` tensor)."""
        if not torch.is_tensor(action):
            assert self.robot.get_action_space().shape[0] == len(action)
        a = self._dev(action, torch.float32, (self.num_envs, self.dim_action))
        n = self.num_envs
        q = torch.empty((n, self.dim_joint), dtype=torch.float32, device=self.device)
        qd = torch.empty_like(q)
        f = torch.empty(n, dtype=torch.uint8, device=self.device)
        _native.check(self._lib.roboy_sim_step(self._h, self._p(a), self._p(q), self._p(qd), self._p(f), self._stream()))
        return self._state_out(q, qd, f)

    def forward_reset_command(self, mask=None) -> RobotState:
        """simulation_client.py:42-44 (optionally only the envs selected by a `[N]` mask)."""
        m = None if mask is None else self._dev(mask, torch.uint8, (self.num_envs,))
        _native.check(self._lib.roboy_sim_reset(self._h, self._p(m), self._stream()))
        return self.read_state()

    def get_new_goal_joint_angles(self):
        """simulation_client.py:46-47: `(3,)` numpy array for one env, `[N,3]` tensor otherwise."""
        g = torch.empty((self.num_envs, self.dim_joint), dtype=torch.float32, device=self.device)
        _native.check(self._lib.roboy_new_goal(self._h, self._p(g), self._stream()))
        return g[0].cpu().numpy() if self.num_envs == 1 else g

    # ------------------------------------------------------------------ fused path (RoboyEnv drives these)
    def configure_env(self, joint_vel_penalty, bonus_for_goal, auto_reset):
        _native.check(self._lib.roboy_set_flags(self._h, int(joint_vel_penalty), int(bonus_for_goal), int(auto_reset)))

    def set_reward_range(self, lo, hi):
        _native.check(self._lib.roboy_set_reward_range(self._h, float(lo), float(hi)))

    def set_seed(self, seed):
        self.seed_value = int(seed) & (2 ** 64 - 1)
        _native.check(self._lib.roboy_set_seed(self._h, self.seed_value))

    def reseed(self, seed):
        """Re-key and redo the construction draws (held state, first goal, counter 0): `RoboyEnv(client, seed=s)`."""
        self.seed_value = int(seed) & (2 ** 64 - 1)
        _native.check(self._lib.roboy_reseed(self._h, self.seed_value))

    def enable_terminal_obs(self, enable=True):
        if enable and self.terminal_obs is None:
            self.terminal_obs = torch.zeros((self.num_envs, self.dim_obs), dtype=torch.float32, device=self.device)
        if not enable:
            self.terminal_obs = None
        if self.info_sink is not None:   # RoboyEnv's info dict: carries the side buffer when enabled
            self.info_sink.pop("terminal_observation", None)
            if self.terminal_obs is not None:
                self.info_sink["terminal_observation"] = self.terminal_obs
        _native.check(self._lib.roboy_set_terminal_obs(self._h, self._p(self.terminal_obs)))

    def step_fused(self, actions, obs=None, reward=None, done=None):
        """One launch: RoboyEnv.step for all envs.  `actions` float32 CUDA `[N,8]`, contiguous.
        Results land in the handle's obs/reward/done buffers unless output tensors are given."""
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous() \
                or actions.numel() != self.num_envs * self.dim_action:
            raise ValueError("actions must be a contiguous float32 CUDA tensor of shape [N, {}]".format(self.dim_action))
        code = self._roboy_step(self._h, actions.data_ptr(), self._p(obs), self._p(reward), self._p(done), self._stream())
        if code:
            _native.check(code)

    def step_many(self, actions, obs=None, reward=None, done=None):
        """Open-loop: T fused steps on pre-recorded `actions` float32 CUDA `[T,N,8]` in ONE launch (the env
        state stays in registers between steps).  Returns `(obs [T,N,9], reward [T,N], done uint8 [T,N])`,
        bit-identical to T `step_fused` calls."""
        T, n = actions.shape[0], self.num_envs
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous() \
                or actions.numel() != T * n * self.dim_action:
            raise ValueError("actions must be a contiguous float32 CUDA tensor of shape [T, N, {}]".format(self.dim_action))
        obs = torch.empty((T, n, self.dim_obs), dtype=torch.float32, device=self.device) if obs is None else obs
        reward = torch.empty((T, n), dtype=torch.float32, device=self.device) if reward is None else reward
        done = torch.empty((T, n), dtype=torch.uint8, device=self.device) if done is None else done
        _native.check(self._lib.roboy_step_many(self._h, T, self._p(actions), self._p(obs), self._p(reward),
                                                self._p(done), self._stream()))
        return obs, reward, done

    def reset_fused(self, mask=None, obs=None):
        m = None if mask is None else self._dev(mask, torch.uint8, (self.num_envs,))
        _native.check(self._lib.roboy_reset(self._h, self._p(m), self._p(obs), self._stream()))

    def step_external(self, q, qd, feasible=None, obs=None, reward=None, done=None):
        """RoboyEnv.step on states produced by an external simulator (float32 CUDA `[N,3]`, uint8 `[N]`)."""
        n = self.num_envs
        q, qd = self._dev(q, torch.float32, (n, self.dim_joint)), self._dev(qd, torch.float32, (n, self.dim_joint))
        f = None if feasible is None else self._dev(feasible, torch.uint8, (n,))
        _native.check(self._lib.roboy_step_external(self._h, self._p(q), self._p(qd), self._p(f), self._p(obs),
                                                    self._p(reward), self._p(done), self._stream()))

    def reset_external(self, q, qd, mask=None, obs=None):
        n = self.num_envs
        q, qd = self._dev(q, torch.float32, (n, self.dim_joint)), self._dev(qd, torch.float32, (n, self.dim_joint))
        m = None if mask is None else self._dev(mask, torch.uint8, (n,))
        _native.check(self._lib.roboy_reset_external(self._h, self._p(m), self._p(q), self._p(qd), self._p(obs),
                                                     self._stream()))

    def set_host_pipeline(self, stage_envs=1 << 19, n_streams=2, ramp=True, ring=True):
        _native.check(self._lib.roboy_set_host_pipeline(self._h, int(stage_envs), int(n_streams)))
        _native.check(self._lib.roboy_set_host_ramp(self._h, int(bool(ramp))))
        _native.check(self._lib.roboy_set_host_pattern(self._h, int(bool(ring))))

    def set_host_autotune(self, enable=True):
        """Default pipeline, stream count tuned by measurement on the next calls (`roboy_set_host_autotune`)."""
        _native.check(self._lib.roboy_set_host_autotune(self._h, int(bool(enable))))

    def host_pipeline(self):
        stage, streams, pattern = ctypes.c_uint64(), ctypes.c_int(), ctypes.c_int()
        _native.check(self._lib.roboy_get_host_pipeline(self._h, ctypes.byref(stage), ctypes.byref(streams), ctypes.byref(pattern)))
        return dict(stage_envs=stage.value, n_streams=streams.value, pattern="ring" if pattern.value else "split")

    def set_host_mode(self, mode):
        """`_native.HOST_STAGED` (copy engines both ways), `HOST_MAPPED_OUT` (the kernel stores its outputs straight into
        the page-locked host buffers) or `HOST_MAPPED_ALL` (it also reads the actions from them)."""
        _native.check(self._lib.roboy_set_host_mode(self._h, int(mode)))

    def host_buffers(self, write_combined_actions=False):
        """Page-locked, device-mapped numpy buffers `(actions [N,8], obs [N,9], reward [N], done [N])` for `step_host`
        (`roboy_host_alloc`); freed when the client closes."""
        n, out = self.num_envs, []
        for shape, dt, wc in (((n, self.dim_action), np.float32, write_combined_actions),
                              ((n, self.dim_obs), np.float32, False), ((n,), np.float32, False), ((n,), np.uint8, False)):
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            ptr = ctypes.c_void_p()
            _native.check(self._lib.roboy_host_alloc(nbytes, int(wc), ctypes.byref(ptr)))
            self._host_allocs.append(ptr.value)
            buf = (ctypes.c_char * nbytes).from_address(ptr.value)
            out.append(np.frombuffer(buf, dtype=dt).reshape(shape))
        return tuple(out)

    def copy_probe(self, actions, obs, reward, done, directions=3, monolithic=False, iters=3):
        """Milliseconds per pass of `step_host`'s copies WITHOUT the kernel (bench.py's e2e.copy_ceiling)."""
        ms = ctypes.c_double()
        _native.check(self._lib.roboy_host_copy_probe(self._h, actions.ctypes.data, obs.ctypes.data, reward.ctypes.data,
                                                      done.ctypes.data, int(directions), int(monolithic), int(iters),
                                                      ctypes.byref(ms)))
        return ms.value

    # ------------------------------------------------------------------ done-index list
    def enable_done_index(self, enable=True):
        _native.check(self._lib.roboy_enable_done_index(self._h, int(bool(enable))))
        self._done_index = bool(enable)
        if enable and self._done_idx is None:
            self._done_idx = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
            self._done_count = torch.zeros(1, dtype=torch.int32, device=self.device)

    def done_indices(self, with_terminal_obs=False, capacity=None):
        """Ascending local ids of the envs the last step finished (roboy_env.py:65-68), computed on the device from the
        step kernel's done bits.  Returns `(idx int32 [k], terminal_rows float32 [k, 9] or None)`; one small D2H read
        (the count) synchronises the stream."""
        if not getattr(self, "_done_index", False):
            self.enable_done_index(True)
        cap = self.num_envs if capacity is None else int(capacity)
        rows = None
        if with_terminal_obs:
            if self._done_rows is None or self._done_rows.shape[0] < cap:
                self._done_rows = torch.empty((cap, self.dim_obs), dtype=torch.float32, device=self.device)
            rows = self._done_rows
        _native.check(self._lib.roboy_done_indices(self._h, self._p(self._done_idx), cap, self._p(self._done_count),
                                                   self._p(rows), self._stream()))
        k = min(int(self._done_count.item()), cap)
        return self._done_idx[:k], (rows[:k] if rows is not None else None)

    def step_host(self, actions, obs, reward, done):
        """The fused step through HOST numpy buffers (pinned for full speed); synchronous."""
        for a, dt, k in ((actions, np.float32, self.dim_action), (obs, np.float32, self.dim_obs), (reward, np.float32, 1),
                         (done, np.uint8, 1)):
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"] or a.size != self.num_envs * k:
                raise ValueError("host buffers must be C-contiguous float32 [N,A], float32 [N,3J], float32 [N], uint8 [N]")
        _native.check(self._lib.roboy_step_host(self._h, actions.ctypes.data, obs.ctypes.data, reward.ctypes.data,
                                                done.ctypes.data))

    def compute_reward(self, q, qd, feasible, goal_q, goal_qd=None, check_range=False):
        """Batched compute_reward + _did_reach_goal: returns (reward float64 [k], reached bool [k])."""
        J = self.dim_joint
        q = self._dev(q, torch.float32).reshape(-1, J)
        k = q.shape[0]
        qd = self._dev(qd, torch.float32, (k, J))
        goal_q = self._dev(goal_q, torch.float32, (k, J))
        gqd = None if goal_qd is None else self._dev(goal_qd, torch.float32, (k, J))
        f = None if feasible is None else self._dev(feasible, torch.uint8, (k,))
        reward = torch.empty(k, dtype=torch.float64, device=self.device)
        reached = torch.empty(k, dtype=torch.uint8, device=self.device)
        _native.check(self._lib.roboy_compute_reward(self._h, k, self._p(q), self._p(qd), self._p(f), self._p(goal_q),
                                                     self._p(gqd), self._p(reward), self._p(reached), int(check_range),
                                                     self._stream()))
        return reward, reached.view(torch.bool)

    # ------------------------------------------------------------------ injection / introspection
    def set_goal(self, goal_q, idx=None):
        i = self._idx(idx)
        g = self._dev(goal_q, torch.float32, (i.numel(), self.dim_joint))
        _native.check(self._lib.roboy_set_goal(self._h, i.numel(), self._p(i), self._p(g), self._stream()))

    def set_state(self, q, qd, feasible=None, idx=None):
        i = self._idx(idx)
        q = self._dev(q, torch.float32, (i.numel(), self.dim_joint))
        qd = self._dev(qd, torch.float32, (i.numel(), self.dim_joint))
        f = None if feasible is None else self._dev(feasible, torch.uint8, (i.numel(),))
        _native.check(self._lib.roboy_set_state(self._h, i.numel(), self._p(i), self._p(q), self._p(qd), self._p(f),
                                                self._stream()))

    def set_step_num(self, step_num, idx=None):
        i = self._idx(idx)
        s = self._dev(step_num, torch.int32).reshape(-1).expand(i.numel()).contiguous()
        _native.check(self._lib.roboy_set_step_num(self._h, i.numel(), self._p(i), self._p(s), self._stream()))

    @property
    def step_num(self):
        return self.step_flags & _native.STEP_MASK

    @property
    def counter(self):
        t = ctypes.c_uint64()
        _native.check(self._lib.roboy_get_counter(self._h, ctypes.byref(t)))
        return t.value

    @counter.setter
    def counter(self, t):
        _native.check(self._lib.roboy_set_counter(self._h, int(t)))

    def stats(self):
        out = (ctypes.c_double * len(_native.STAT_NAMES))()
        _native.check(self._lib.roboy_stats(self._h, out, self._stream()))
        return dict(zip(_native.STAT_NAMES, list(out)))

    def clear_stats(self):
        _native.check(self._lib.roboy_clear_stats(self._h, self._stream()))

    def clear_errors(self):
        _native.check(self._lib.roboy_clear_errors(self._h, self._stream()))

    def errors(self):
        """(error word, first offending global env id or None); synchronises the stream."""
        flags, first = ctypes.c_uint32(), ctypes.c_uint64()
        _native.check(self._lib.roboy_errors(self._h, ctypes.byref(flags), ctypes.byref(first), self._stream()))
        return flags.value, (None if first.value == 2 ** 64 - 1 else first.value)

    def launch_count(self):
        n = ctypes.c_uint64()
        _native.check(self._lib.roboy_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def null_step(self):
        """An empty kernel launched like the step kernel (launch-floor measurements)."""
        _native.check(self._lib.roboy_null_step(self._h, self._stream()))

    def step_geometry(self):
        g, b, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _native.check(self._lib.roboy_step_geometry(self._h, ctypes.byref(g), ctypes.byref(b), ctypes.byref(s)))
        return dict(grid=g.value, block=b.value, smem_bytes=s.value)

    # ------------------------------------------------------------------ checkpoint
    def state_dict(self):
        """The complete checkpoint of the shard: SoA state, call counter, seed, statistics, the error word and the
        CURRENT observation (the Stub never stores a sampled state, so the last observation exists only in `obs`)."""
        flags, first = self.errors()
        return dict(goal=self.goal.clone(), step_flags=self.step_flags.clone(), held=self.held.clone(),
                    obs=self.obs.clone(), counter=self.counter, seed=self.seed_value, env_id_base=self.env_id_base,
                    num_envs=self.num_envs, stats=self.stats_tensor.clone(), err_flags=flags, first_bad=first)

    def load_state_dict(self, sd):
        if int(sd.get("num_envs", self.num_envs)) != self.num_envs or int(sd["env_id_base"]) != self.env_id_base:
            raise ValueError("checkpoint is for {} envs at env_id_base {}, this client holds {} at {}".format(
                sd.get("num_envs"), sd["env_id_base"], self.num_envs, self.env_id_base))
        self.goal.copy_(sd["goal"])
        self.step_flags.copy_(sd["step_flags"])
        self.held.copy_(sd["held"])
        if "obs" in sd:
            self.obs.copy_(sd["obs"])
        self.stats_tensor.copy_(sd["stats"])
        self.clear_errors()   # the error word is not restorable through the C-ABI; a checkpoint taken with errors says so
        if sd.get("err_flags"):
            raise ValueError("checkpoint was taken with a non-zero error word ({}, first env {})".format(
                sd["err_flags"], sd.get("first_bad")))
        self.set_seed(sd["seed"])
        self.counter = sd["counter"]
