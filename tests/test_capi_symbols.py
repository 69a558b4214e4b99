"""The C-ABI shared library loads and exports every symbol include/roboy_b200.h declares;
no compute calls (this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_cuda
from gym_roboy_b200 import _native

HEADER = os.path.join(ROOT, "include", "roboy_b200.h")


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(roboy_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_functions()
    assert len(names) >= 30
    assert sorted(_native.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.lib_path())
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert _native.load().roboy_abi_version() == _native.ABI_VERSION


def test_cfg_struct_and_msj_constants():
    lib = _native.load()
    cfg = _native.RoboyCfg()
    assert ctypes.sizeof(cfg) == 88 + 16 + 4 * 4 * _native.JOINT_PAD + 2 * 4 * _native.MAX_ACTION
    _native.check(lib.roboy_cfg_msj(ctypes.byref(cfg)))
    # float32 bounds of msj_robot.py:9,10,16 (hex from SURVEY.md 8a a13)
    assert float(cfg.angle_high).hex() == "0x1.921fb60000000p+1" and cfg.angle_low == -cfg.angle_high
    assert float(cfg.vel_high).hex() == "0x1.0c15240000000p-1" and cfg.vel_low == -cfg.vel_high
    assert float(cfg.act_high).hex() == "0x1.3333340000000p-2" and cfg.act_low == -cfg.act_high
    assert (cfg.max_episode_len, cfg.joint_vel_penalty, cfg.bonus_for_goal, cfg.auto_reset) == (400, 0, 1, 1)
    assert (cfg.penalty_boundary, cfg.bonus_goal) == (1.0, 1000.0)
    assert (cfg.dim_joint, cfg.dim_action, cfg.per_component_bounds) == (3, 8, 0)
    assert list(cfg.angle_high_v)[:3] == [cfg.angle_high] * 3 and list(cfg.act_low_v)[:8] == [cfg.act_low] * 8


def test_hold_intervals_per_tendon_and_robot_caps():
    lib = _native.load()
    cfg = _native.RoboyCfg()
    lib.roboy_cfg_msj(ctypes.byref(cfg))
    lo, hi = (ctypes.c_float * 64)(), (ctypes.c_float * 64)()
    _native.check(lib.roboy_hold_intervals(ctypes.byref(cfg), lo, hi))
    assert all(lo[k] == -(2.0 ** -24) and hi[k] == 2.0 ** -25 for k in range(8))     # SURVEY.md 8a a4 [probe]
    cfg.per_component_bounds, cfg.dim_action = 1, 3
    for k, (a, b) in enumerate(((-0.3, 0.3), (-0.1, 0.4), (-0.5, 0.25))):
        cfg.act_low_v[k], cfg.act_high_v[k] = a, b
    _native.check(lib.roboy_hold_intervals(ctypes.byref(cfg), lo, hi))
    assert lo[0] == -(2.0 ** -24) and hi[0] == 2.0 ** -25
    for k in (1, 2):   # asymmetric tendon ranges hold around the action that rescales to zero, if any float32 does
        zero = 1 - 2 * cfg.act_high_v[k] / (cfg.act_high_v[k] - cfg.act_low_v[k])
        assert lo[k] > hi[k] or (lo[k] <= hi[k] and abs(lo[k] - zero) < 1e-6 and abs(hi[k] - zero) < 1e-6)
    cfg.dim_joint = 16
    h = ctypes.c_void_p()
    assert lib.roboy_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1 and b"dim_joint" in lib.roboy_last_error()
    cfg.dim_joint, cfg.dim_action = 3, 65
    assert lib.roboy_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1 and b"dim_action" in lib.roboy_last_error()
    # joint limits of 2^100 and beyond: outside what the normalisation's fused numerator is proved for (test_fastdiv_proof)
    lib.roboy_cfg_msj(ctypes.byref(cfg))
    cfg.n_envs, cfg.angle_high = 8, 2.0 ** 100
    assert lib.roboy_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1 and b"2^100" in lib.roboy_last_error()


def test_argument_errors_are_reported_not_crashed():
    lib = _native.load()
    assert lib.roboy_cfg_msj(None) == -1
    assert lib.roboy_step(None, None, None, None, None, None) == -1
    assert b"NULL" in lib.roboy_last_error()
    cfg = _native.RoboyCfg()
    lib.roboy_cfg_msj(ctypes.byref(cfg))
    cfg.n_envs = 0
    h = ctypes.c_void_p()
    assert lib.roboy_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1 and not h.value


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cuda_device_fails_loudly_instead_of_falling_back():
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    with pytest.raises(_native.RoboyNativeError, match="no CPU"):
        CudaSimulationClient(num_envs=4)
    lib = _native.load()
    cfg = _native.RoboyCfg()
    lib.roboy_cfg_msj(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    assert lib.roboy_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -2


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under gym_roboy_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gym_roboy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "tests._shim" not in text, f
