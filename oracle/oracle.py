"""ctypes wrapper around oracle/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see the header of roboy_oracle.c).  It also carries an independent numpy
Philox4x32-10 used to cross-check the C one and to feed the reference replay harness.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

STEP_MASK = 0x00FFFFFF
F_HELD_ZERO64 = 1 << 24
F_HELD_INFEASIBLE = 1 << 25
ERR_ACTION, ERR_REWARD_RANGE, ERR_GOAL_BOUNDS = 1, 2, 4
STAT_NAMES = ("steps", "episodes", "successes", "timeouts", "sum_reward", "sum_episode_len", "holds", "violations")
STREAM_STATE, STREAM_GOAL = 0, 2

# float32 MSJ bounds, msj_robot.py:9,10,16
PI32 = np.float32(np.pi)
MSJ = dict(
    angle_low=float(-PI32), angle_high=float(PI32),
    vel_low=float(np.float32(-np.pi / 6)), vel_high=float(np.float32(np.pi / 6)),
    act_low=float(np.float32(-0.3)), act_high=float(np.float32(0.3)),
)


class OrcCfg(ctypes.Structure):
    _fields_ = [
        ("n_envs", ctypes.c_uint64), ("env_id_base", ctypes.c_uint64), ("seed", ctypes.c_uint64),
        ("angle_low", ctypes.c_float), ("angle_high", ctypes.c_float),
        ("vel_low", ctypes.c_float), ("vel_high", ctypes.c_float),
        ("act_low", ctypes.c_float), ("act_high", ctypes.c_float),
        ("max_episode_len", ctypes.c_int32), ("joint_vel_penalty", ctypes.c_int32),
        ("bonus_for_goal", ctypes.c_int32), ("auto_reset", ctypes.c_int32),
        ("penalty_boundary", ctypes.c_float), ("bonus_goal", ctypes.c_float),
        ("reward_lo", ctypes.c_double), ("reward_hi", ctypes.c_double),
    ]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "roboy_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, u64, f32p = ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_float)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [ctypes.POINTER(OrcCfg)]
        L.orc_destroy.argtypes = [vp]
        for name, rt in (("orc_goal", f32p), ("orc_step_flags", ctypes.POINTER(ctypes.c_uint32)),
                         ("orc_held", f32p), ("orc_stats", ctypes.POINTER(ctypes.c_double))):
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [vp]
        L.orc_counter.restype = u64
        L.orc_counter.argtypes = [vp]
        L.orc_set_counter.argtypes = [vp, u64]
        L.orc_err_flags.restype = ctypes.c_uint32
        L.orc_err_flags.argtypes = [vp]
        L.orc_first_bad_env.restype = u64
        L.orc_first_bad_env.argtypes = [vp]
        L.orc_set_reward_range.argtypes = [vp, ctypes.c_double, ctypes.c_double]
        L.orc_reset.argtypes = [vp, vp, vp]
        L.orc_step.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int]
        L.orc_external.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.orc_compute_reward.argtypes = [ctypes.POINTER(OrcCfg), u64, vp, vp, vp, vp, vp, vp, vp, vp]
        L.orc_thresholds.argtypes = [ctypes.POINTER(OrcCfg), f32p, f32p]
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_draw_state.argtypes = [ctypes.POINTER(OrcCfg), u64, u64, vp, vp]
        L.orc_draw_goal.argtypes = [ctypes.POINTER(OrcCfg), u64, u64, vp]
        _lib = L
    return _lib


def make_cfg(n_envs, seed=1234, env_id_base=0, joint_vel_penalty=False, bonus=True, auto_reset=True,
             reward_range=None, max_episode_len=400, **bounds):
    b = dict(MSJ)
    b.update(bounds)
    lo, hi = (-math.inf, math.inf) if reward_range is None else reward_range
    return OrcCfg(n_envs=n_envs, env_id_base=env_id_base, seed=seed,
                  angle_low=b["angle_low"], angle_high=b["angle_high"],
                  vel_low=b["vel_low"], vel_high=b["vel_high"],
                  act_low=b["act_low"], act_high=b["act_high"],
                  max_episode_len=max_episode_len, joint_vel_penalty=int(joint_vel_penalty),
                  bonus_for_goal=int(bonus), auto_reset=int(auto_reset),
                  penalty_boundary=1.0, bonus_goal=1000.0, reward_lo=lo, reward_hi=hi)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def thresholds(cfg):
    a, v = ctypes.c_float(), ctypes.c_float()
    lib().orc_thresholds(ctypes.byref(cfg), ctypes.byref(a), ctypes.byref(v))
    return np.float32(a.value), np.float32(v.value)


def compute_reward(cfg, q, qd, feasible, goal_q, goal_qd=None):
    """Stand-alone compute_reward + _did_reach_goal over float32 [n,3] arrays."""
    q = np.ascontiguousarray(q, np.float32)
    qd = np.ascontiguousarray(qd, np.float32)
    goal_q = np.ascontiguousarray(goal_q, np.float32)
    n = q.shape[0]
    feasible = np.ascontiguousarray(feasible, np.uint8)
    gqd = None if goal_qd is None else np.ascontiguousarray(goal_qd, np.float32)
    reward = np.empty(n, np.float64)
    reached = np.empty(n, np.uint8)
    viol = np.empty(n, np.uint8)
    lib().orc_compute_reward(ctypes.byref(cfg), n, _ptr(q), _ptr(qd), _ptr(feasible), _ptr(goal_q), _ptr(gqd),
                             _ptr(reward), _ptr(reached), _ptr(viol))
    return reward, reached.astype(bool), viol.astype(bool)


def reward_range(joint_vel_penalty=False, bonus=True, **bounds):
    """roboy_env.py:40-49 _create_reward_range -> (min_reward, max_reward) python floats."""
    cfg = make_cfg(1, joint_vel_penalty=joint_vel_penalty, bonus=bonus, **bounds)
    hi_q = np.full((1, 3), cfg.angle_high, np.float32)
    hi_v = np.full((1, 3), cfg.vel_high, np.float32)
    lo_q = np.full((1, 3), cfg.angle_low, np.float32)
    lo_v = np.full((1, 3), cfg.vel_low, np.float32)
    rmax, _, _ = compute_reward(cfg, hi_q, hi_v, [1], hi_q, hi_v)
    rmin, _, _ = compute_reward(cfg, lo_q, lo_v, [0], hi_q, hi_v)
    return float(rmin[0]), float(rmax[0])


class OracleEnv:
    """N batched envs stepped by the C restatement (reference semantics + vec-env auto-reset)."""

    def __init__(self, n_envs, seed=1234, env_id_base=0, joint_vel_penalty=False, bonus=True,
                 auto_reset=True, check_reward_range=True, threads=1, **bounds):
        self.n = int(n_envs)
        rr = reward_range(joint_vel_penalty, bonus, **bounds) if check_reward_range else None
        self.reward_range = rr
        self.cfg = make_cfg(self.n, seed, env_id_base, joint_vel_penalty, bonus, auto_reset, rr, **bounds)
        self._h = lib().orc_create(ctypes.byref(self.cfg))
        self.threads = threads
        n = self.n
        L = lib()
        self.goal = np.ctypeslib.as_array(L.orc_goal(self._h), shape=(3, n))
        self.step_flags = np.ctypeslib.as_array(L.orc_step_flags(self._h), shape=(n,))
        self.held = np.ctypeslib.as_array(L.orc_held(self._h), shape=(6, n))
        self._stats = np.ctypeslib.as_array(L.orc_stats(self._h), shape=(len(STAT_NAMES),))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None and lib is not None:   # (module globals are None at interpreter shutdown)
            _lib.orc_destroy(h)

    @property
    def counter(self):
        return int(lib().orc_counter(self._h))

    @counter.setter
    def counter(self, t):
        lib().orc_set_counter(self._h, int(t))

    @property
    def step_num(self):
        return (self.step_flags & STEP_MASK).astype(np.int64)

    def stats(self):
        return dict(zip(STAT_NAMES, self._stats.tolist()))

    def errors(self):
        return int(lib().orc_err_flags(self._h)), int(lib().orc_first_bad_env(self._h))

    def reset(self, mask=None):
        obs = np.zeros((self.n, 9), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orc_reset(self._h, _ptr(m), _ptr(obs))
        return obs

    def step_external(self, q, qd, feasible=None):
        q, qd = np.ascontiguousarray(q, np.float32), np.ascontiguousarray(qd, np.float32)
        f = None if feasible is None else np.ascontiguousarray(feasible, np.uint8)
        obs, rew, done = np.zeros((self.n, 9), np.float32), np.zeros(self.n, np.float32), np.zeros(self.n, np.uint8)
        lib().orc_external(self._h, 0, None, _ptr(q), _ptr(qd), _ptr(f), _ptr(obs), _ptr(rew), _ptr(done))
        return obs, rew, done.astype(bool)

    def reset_external(self, q, qd, mask=None):
        q, qd = np.ascontiguousarray(q, np.float32), np.ascontiguousarray(qd, np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        obs = np.zeros((self.n, 9), np.float32)
        lib().orc_external(self._h, 1, _ptr(m), _ptr(q), _ptr(qd), None, _ptr(obs), None, None)
        return obs

    def step(self, actions, want_terminal_obs=False):
        a = np.ascontiguousarray(actions, np.float32)
        assert a.shape == (self.n, 8)
        obs = np.empty((self.n, 9), np.float32)
        rew = np.empty(self.n, np.float32)
        done = np.empty(self.n, np.uint8)
        term = np.zeros((self.n, 9), np.float32) if want_terminal_obs else None
        lib().orc_step(self._h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(done), _ptr(term), int(self.threads))
        if want_terminal_obs:
            return obs, rew, done.astype(bool), term
        return obs, rew, done.astype(bool)


# ---------------------------------------------------------------------------------------------
# Independent numpy Philox4x32-10 (vectorised) + the draw mapping, for cross-checks and replay.
# ---------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, np.uint64) & np.uint64(0xFFFFFFFF) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    sh = np.uint64(32)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> sh) ^ c1 ^ np.uint64(k0)
        n1 = p1 & mask
        n2 = (p0 >> sh) ^ c3 ^ np.uint64(k1)
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def _uniform(x, low, high):
    u = (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    low, high = np.float32(low), np.float32(high)
    span = np.float32(high - low)
    return (low + (span * u).astype(np.float32)).astype(np.float32)


def _block(seed, gids, t, stream, sub=0):
    gids = np.asarray(gids, np.uint64)
    c3 = (int(stream) << 28) | ((int(sub) & 0xFF) << 20) | ((int(t) >> 32) & 0x000FFFFF)
    return philox4x32_10(gids & np.uint64(0xFFFFFFFF), gids >> np.uint64(32), int(t) & 0xFFFFFFFF, c3,
                         int(seed) & 0xFFFFFFFF, int(seed) >> 32)


def draw_goal(seed, gids, t, low=MSJ["angle_low"], high=MSJ["angle_high"], sub=0):
    """float32 [n,3] goal draws: v = low + span * ((x >> 8) * 2^-24) from words x, y, z of the block."""
    x = _block(seed, gids, t, STREAM_GOAL, sub)
    return np.stack([_uniform(x[k], low, high) for k in range(3)], axis=-1)


def draw_state(seed, gids, t, low=MSJ["angle_low"], high=MSJ["angle_high"]):
    """(q, qd) float32 [n,3] each: six 21-bit integers from ONE block, v = low + (span*2^-21)*k."""
    r = [w.astype(np.uint64) for w in _block(seed, gids, t, STREAM_STATE)]
    m21 = np.uint64(0x1FFFFF)
    k = [r[0] >> np.uint64(11), (((r[0] & np.uint64(0x7FF)) << np.uint64(10)) | (r[1] >> np.uint64(22))) & m21,
         (r[1] >> np.uint64(1)) & m21,
         r[2] >> np.uint64(11), (((r[2] & np.uint64(0x7FF)) << np.uint64(10)) | (r[3] >> np.uint64(22))) & m21,
         (r[3] >> np.uint64(1)) & m21]
    low = np.float32(low)
    span21 = np.float32(np.float32(np.float32(high) - low) * np.float32(2.0 ** -21))
    v = [(low + (span21 * ki.astype(np.float32)).astype(np.float32)).astype(np.float32) for ki in k]
    return np.stack(v[0:3], axis=-1), np.stack(v[3:6], axis=-1)
