"""Proof by exhaustion (CPU) that the step kernel's 3-instruction division by the MSJ spans is
bit-identical to the IEEE float32 division the reference performs (roboy_robot.py:93-95), and
that the host-side hold interval is the exact pre-image of numpy's allclose(rescaled, 0)."""
import ctypes
import os
import subprocess

import numpy as np

from conftest import ROOT
from gym_roboy_b200 import _native


def test_fast_division_is_exact_for_the_msj_spans(tmp_path):
    exe = str(tmp_path / "verify_fastdiv")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-o", exe,
                           os.path.join(ROOT, "oracle", "verify_fastdiv.c"), "-lm"])
    a_span = np.float32(np.float32(np.pi) - np.float32(-np.pi))
    v_span = np.float32(np.float32(np.pi / 6) - np.float32(-np.pi / 6))
    assert hex(a_span.view(np.uint32)) == "0x40c90fdb" and hex(v_span.view(np.uint32)) == "0x3f860a92"
    for span in (a_span, v_span):   # every float32 numerator with 2^-30 <= |x| <= 8, and 0
        out = subprocess.run([exe, "%08x" % span.view(np.uint32), repr(2.0 ** -30), "8"], capture_output=True, text=True)
        assert out.returncode == 0 and " 0 mismatches" in out.stdout, out.stdout + out.stderr


def test_fused_numerator_is_exact(tmp_path):
    """fma(2, v, -max) == fl(fl(2*v) - max) for EVERY float32 v (all 2^32 bit patterns): the FFMA ptxas emits for the
    numerator of roboy_robot.py:95 (a value-preserving contraction it performs even under -fmad=false) is the reference's
    multiply-then-subtract.  MSJ's two `max` values, a negative and a large one; and the check is sensitive -- it fails
    for |max| >= 2^103, bounds roboy_create refuses."""
    exe = str(tmp_path / "verify_fused_numerator")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-o", exe,
                           os.path.join(ROOT, "oracle", "verify_fused_numerator.c"), "-lm"])
    for mx in (np.float32(np.pi), np.float32(np.pi / 6), np.float32(-2.5), np.float32(1e30)):
        out = subprocess.run([exe, "%08x" % mx.view(np.uint32)], capture_output=True, text=True)
        assert out.returncode == 0 and " 0 mismatches" in out.stdout, out.stdout + out.stderr
    out = subprocess.run([exe, "7f000000"], capture_output=True, text=True)   # max = 2^127
    assert out.returncode == 1 and " 0 mismatches" not in out.stdout


def test_numerators_are_zero_or_not_tiny():
    """The premise of the proof's range: t = (2v - max) - min is 0 or >= 2^-25 in magnitude."""
    rng = np.random.default_rng(0)
    for hi in (np.float32(np.pi), np.float32(np.pi / 6)):
        v = np.concatenate([rng.uniform(-np.pi, np.pi, 2_000_000), 10.0 ** rng.uniform(-40, 0, 1_000_000),
                            -(10.0 ** rng.uniform(-40, 0, 1_000_000)), [0.0, hi / 2, -hi / 2, hi, -hi]]).astype(np.float32)
        t = (np.float32(2) * v - hi) - (-hi)
        nz = t[t != 0]
        assert np.abs(nz).min() >= 2.0 ** -25 and np.abs(t).max() <= 8.0
        assert not np.signbit(t[t == 0]).any()        # never -0


def test_hold_interval_is_the_allclose_preimage():
    lib = _native.load()
    cfg = _native.RoboyCfg()
    lib.roboy_cfg_msj(ctypes.byref(cfg))
    lo, hi = ctypes.c_float(), ctypes.c_float()
    assert lib.roboy_hold_interval(ctypes.byref(cfg), ctypes.byref(lo), ctypes.byref(hi)) == 0
    assert lo.value == -(2.0 ** -24) and hi.value == 2.0 ** -25        # SURVEY.md 8a a4 ([probe])

    def holds(a):   # roboy_env.py:157-158 in float32, then numpy's allclose(., 0) on python floats
        a = np.float32(a)
        slope = np.float32((np.float32(0.3) - np.float32(-0.3)) / (np.float32(1) - np.float32(-1)))
        r = slope * (a - np.float32(1)) + np.float32(0.3)
        return bool(np.allclose([float(r)], 0))

    lo32, hi32 = np.float32(lo.value), np.float32(hi.value)
    assert holds(lo32) and holds(hi32) and holds(0.0)
    assert not holds(np.nextafter(lo32, np.float32(-1))) and not holds(np.nextafter(hi32, np.float32(1)))
    inside = np.random.default_rng(1).uniform(lo.value, hi.value, 2000).astype(np.float32)
    assert all(holds(a) for a in inside)
