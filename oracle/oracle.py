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
ERR_ACTION, ERR_REWARD_RANGE, ERR_GOAL_BOUNDS, ERR_STATE_BOUNDS = 1, 2, 4, 8
STAT_NAMES = ("steps", "episodes", "successes", "timeouts", "sum_reward", "sum_episode_len", "holds", "violations")
STREAM_STATE, STREAM_GOAL = 0, 2

# float32 MSJ bounds, msj_robot.py:9,10,16
PI32 = np.float32(np.pi)
MSJ = dict(
    angle_low=float(-PI32), angle_high=float(PI32),
    vel_low=float(np.float32(-np.pi / 6)), vel_high=float(np.float32(np.pi / 6)),
    act_low=float(np.float32(-0.3)), act_high=float(np.float32(0.3)),
)


MAX_JOINT, JOINT_PAD, MAX_ACTION = 15, 16, 64


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
        ("dim_joint", ctypes.c_int32), ("dim_action", ctypes.c_int32),
        ("per_component_bounds", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("angle_low_v", ctypes.c_float * JOINT_PAD), ("angle_high_v", ctypes.c_float * JOINT_PAD),
        ("vel_low_v", ctypes.c_float * JOINT_PAD), ("vel_high_v", ctypes.c_float * JOINT_PAD),
        ("act_low_v", ctypes.c_float * MAX_ACTION), ("act_high_v", ctypes.c_float * MAX_ACTION),
    ]

    @property
    def J(self):
        return self.dim_joint or 3

    @property
    def A(self):
        return self.dim_action or 8


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


def robot_bounds(bounds):
    """Normalise a bounds dict (scalars or per-component sequences; optional dim_joint / dim_action) to float32 arrays:
    returns (J, A, per_component, {name: float32 array of length J or A})."""
    b = dict(MSJ)
    b.update(bounds)
    J = int(b.pop("dim_joint", 0)) or None
    A = int(b.pop("dim_action", 0)) or None
    is_vec = {k: np.ndim(v) > 0 for k, v in b.items()}       # a sequence (even of length 1) gives the dimension
    arrs = {k: np.atleast_1d(np.asarray(v, np.float32)) for k, v in b.items()}
    for k in ("angle_low", "angle_high", "vel_low", "vel_high"):
        if is_vec[k]:
            J = J or arrs[k].size
    for k in ("act_low", "act_high"):
        if is_vec[k]:
            A = A or arrs[k].size
    J, A = J or 3, A or 8
    per_component = any(is_vec.values())
    out = {}
    for k, a in arrs.items():
        n = A if k.startswith("act") else J
        out[k] = np.broadcast_to(a, (n,)).astype(np.float32).copy()
    return J, A, per_component, out


def make_cfg(n_envs, seed=1234, env_id_base=0, joint_vel_penalty=False, bonus=True, auto_reset=True,
             reward_range=None, max_episode_len=400, **bounds):
    J, A, per_component, b = robot_bounds(bounds)
    lo, hi = (-math.inf, math.inf) if reward_range is None else reward_range
    cfg = OrcCfg(n_envs=n_envs, env_id_base=env_id_base, seed=seed,
                 angle_low=float(b["angle_low"][0]), angle_high=float(b["angle_high"][0]),
                 vel_low=float(b["vel_low"][0]), vel_high=float(b["vel_high"][0]),
                 act_low=float(b["act_low"][0]), act_high=float(b["act_high"][0]),
                 max_episode_len=max_episode_len, joint_vel_penalty=int(joint_vel_penalty),
                 bonus_for_goal=int(bonus), auto_reset=int(auto_reset),
                 penalty_boundary=1.0, bonus_goal=1000.0, reward_lo=lo, reward_hi=hi,
                 dim_joint=J, dim_action=A, per_component_bounds=int(per_component))
    for name in ("angle_low", "angle_high", "vel_low", "vel_high", "act_low", "act_high"):
        dst = getattr(cfg, name + "_v")
        for k, v in enumerate(b[name]):
            dst[k] = float(v)
    return cfg


def hold_action(bounds):
    """Per tendon, a [-1, 1] action whose float32 rescale (roboy_env.py:157-158) passes numpy's allclose(., 0) -- where
    the Stub holds (simulation_client.py:38) -- searched among the float32 neighbours of the real-valued zero crossing.
    Returns (action float32 [A], can_hold): for some bounds no float32 action rescales to within 1e-8 of zero and the
    robot never holds."""
    _, A, _, bb = robot_bounds(bounds)
    lo, hi = bb["act_low"], bb["act_high"]
    slope = ((hi - lo) / np.float32(2.0)).astype(np.float32)
    guess = (1 - 2 * hi.astype(np.float64) / (hi.astype(np.float64) - lo.astype(np.float64))).astype(np.float32)
    out, ok = guess.copy(), np.zeros(A, bool)
    for k in range(A):
        up = dn = guess[k]
        cand = [guess[k]]
        for _ in range(8):
            up = np.nextafter(up, np.float32(2)); dn = np.nextafter(dn, np.float32(-2))
            cand += [up, dn]
        for c in cand:
            r = np.float32(np.float32(slope[k] * np.float32(c - np.float32(1.0))) + hi[k])
            if abs(float(r)) <= 1e-8:
                out[k], ok[k] = c, True
                break
    return out, bool(ok.all())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def thresholds(cfg):
    a, v = ctypes.c_float(), ctypes.c_float()
    lib().orc_thresholds(ctypes.byref(cfg), ctypes.byref(a), ctypes.byref(v))
    return np.float32(a.value), np.float32(v.value)


def compute_reward(cfg, q, qd, feasible, goal_q, goal_qd=None):
    """Stand-alone compute_reward + _did_reach_goal over float32 [n,J] arrays."""
    q = np.ascontiguousarray(q, np.float32).reshape(-1, cfg.J)
    qd = np.ascontiguousarray(qd, np.float32).reshape(-1, cfg.J)
    goal_q = np.ascontiguousarray(goal_q, np.float32).reshape(-1, cfg.J)
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
    _, _, _, b = robot_bounds(bounds)
    hi_q, hi_v = b["angle_high"][None, :], b["vel_high"][None, :]
    lo_q, lo_v = b["angle_low"][None, :], b["vel_low"][None, :]
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
        assert self._h, "robot exceeds the oracle's caps (%d joints, %d tendons)" % (MAX_JOINT, MAX_ACTION)
        self.threads = threads
        n = self.n
        L = lib()
        self.J, self.A = self.cfg.J, self.cfg.A
        self.obs_dim = 3 * self.J
        self.bounds = robot_bounds(bounds)[3]
        self.goal = np.ctypeslib.as_array(L.orc_goal(self._h), shape=(self.J, n))
        self.step_flags = np.ctypeslib.as_array(L.orc_step_flags(self._h), shape=(n,))
        self.held = np.ctypeslib.as_array(L.orc_held(self._h), shape=(2 * self.J, n))
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
        obs = np.zeros((self.n, self.obs_dim), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orc_reset(self._h, _ptr(m), _ptr(obs))
        return obs

    def step_external(self, q, qd, feasible=None):
        q, qd = np.ascontiguousarray(q, np.float32), np.ascontiguousarray(qd, np.float32)
        f = None if feasible is None else np.ascontiguousarray(feasible, np.uint8)
        obs, rew, done = np.zeros((self.n, self.obs_dim), np.float32), np.zeros(self.n, np.float32), np.zeros(self.n, np.uint8)
        lib().orc_external(self._h, 0, None, _ptr(q), _ptr(qd), _ptr(f), _ptr(obs), _ptr(rew), _ptr(done))
        return obs, rew, done.astype(bool)

    def reset_external(self, q, qd, mask=None):
        q, qd = np.ascontiguousarray(q, np.float32), np.ascontiguousarray(qd, np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        obs = np.zeros((self.n, self.obs_dim), np.float32)
        lib().orc_external(self._h, 1, _ptr(m), _ptr(q), _ptr(qd), None, _ptr(obs), None, None)
        return obs

    def step(self, actions, want_terminal_obs=False):
        a = np.ascontiguousarray(actions, np.float32)
        assert a.shape == (self.n, self.A)
        obs = np.empty((self.n, self.obs_dim), np.float32)
        rew = np.empty(self.n, np.float32)
        done = np.empty(self.n, np.uint8)
        term = np.zeros((self.n, self.obs_dim), np.float32) if want_terminal_obs else None
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


def _block(seed, gids, t, stream, sub=0, block=0):
    gids = np.asarray(gids, np.uint64)
    c3 = (int(stream) << 28) | ((int(sub) & 0xFF) << 20) | ((int(t) >> 32) & 0x000FFFFF)
    c1 = (gids >> np.uint64(32)) | np.uint64(int(block) << 24)    # block: multi-block draws of robots with > 3 joints
    return philox4x32_10(gids & np.uint64(0xFFFFFFFF), c1, int(t) & 0xFFFFFFFF, c3,
                         int(seed) & 0xFFFFFFFF, int(seed) >> 32)


def _per_joint(low, high, J=None):
    low, high = np.atleast_1d(np.asarray(low, np.float32)), np.atleast_1d(np.asarray(high, np.float32))
    J = J or max(low.size, high.size, 3 if max(low.size, high.size) == 1 else 1)
    return np.broadcast_to(low, (J,)).astype(np.float32), np.broadcast_to(high, (J,)).astype(np.float32)


def draw_goal(seed, gids, t, low=MSJ["angle_low"], high=MSJ["angle_high"], sub=0, J=None):
    """float32 [n,J] goal draws: component k = low_k + span_k * ((x >> 8) * 2^-24) from word k % 3 of goal-stream block k // 3."""
    low, high = _per_joint(low, high, J)
    cols, x = [], None
    for k in range(low.size):
        if k % 3 == 0:
            x = _block(seed, gids, t, STREAM_GOAL, sub, k // 3)
        cols.append(_uniform(x[k % 3], low[k], high[k]))
    return np.stack(cols, axis=-1)


def draw_state(seed, gids, t, low=MSJ["angle_low"], high=MSJ["angle_high"], J=None):
    """(q, qd) float32 [n,J] each: value c of (q_0..q_{J-1}, qd_0..qd_{J-1}) is the (c % 6)-th 21-bit integer of
    state-stream block c // 6, v = low + (span*2^-21)*k with the ANGLE bounds of its component (roboy_robot.py:38)."""
    low, high = _per_joint(low, high, J)
    J = low.size
    m21 = np.uint64(0x1FFFFF)
    vals, k = [], None
    for c in range(2 * J):
        if c % 6 == 0:
            r = [w.astype(np.uint64) for w in _block(seed, gids, t, STREAM_STATE, 0, c // 6)]
            k = [r[0] >> np.uint64(11), (((r[0] & np.uint64(0x7FF)) << np.uint64(10)) | (r[1] >> np.uint64(22))) & m21,
                 (r[1] >> np.uint64(1)) & m21,
                 r[2] >> np.uint64(11), (((r[2] & np.uint64(0x7FF)) << np.uint64(10)) | (r[3] >> np.uint64(22))) & m21,
                 (r[3] >> np.uint64(1)) & m21]
        j = c % J
        span21 = np.float32(np.float32(high[j] - low[j]) * np.float32(2.0 ** -21))
        vals.append((low[j] + (span21 * k[c % 6].astype(np.float32)).astype(np.float32)).astype(np.float32))
    return np.stack(vals[0:J], axis=-1), np.stack(vals[J:2 * J], axis=-1)
