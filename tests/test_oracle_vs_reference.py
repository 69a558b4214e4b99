"""Oracle vs the UNMODIFIED reference, live (build container only: skipped where
/root/reference does not exist).  Also re-runs the reference's own unit tests through the
test-only gym shim, so the shim itself is pinned."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as orc
from oracle import reference_harness as rh

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not rh.available(), reason="reference checkout not present")]


def test_reference_own_suite_passes_through_the_shim(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "_shim"), rh.REFERENCE_ROOT]),
               PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(rh.REFERENCE_ROOT, "gym_roboy", "envs", "tests"),
                          "-q", "-p", "no:cacheprovider"], cwd=tmp_path, env=env, capture_output=True, text=True)
    assert "21 passed, 12 skipped" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("penalty,bonus,auto", [(False, True, True), (True, True, True), (False, False, False)])
def test_random_rollout_matches_reference(penalty, bonus, auto):
    N, T = 12, 150
    ref = rh.ReferenceVecEnv(N, seed=5, joint_vel_penalty=penalty, bonus=bonus, auto_reset=auto)
    o = orc.OracleEnv(N, seed=5, joint_vel_penalty=penalty, bonus=bonus, auto_reset=auto)
    assert ref.reward_range == o.reward_range
    assert np.array_equal(ref.goals().T, o.goal)
    assert np.array_equal(ref.reset().astype(np.float32), o.reset())
    steps = np.array([1 + (53 * i) % 400 for i in range(N)])
    for i in range(N):
        ref.set_step_num(i, steps[i])
    o.step_flags[:] = (o.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
    rng = np.random.default_rng(0)
    alive = np.ones(N, bool)
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)
        a[rng.random(N) < 0.1] = 0
        ro, rr, rd, rt, raised = ref.step(a)
        oo, orw, od, ot = o.step(a, want_terminal_obs=True)
        alive &= np.array([m == "" for m in raised])
        assert np.array_equal(ro.astype(np.float32)[alive], oo[alive])
        assert np.array_equal(rd[alive], od[alive])
        assert np.allclose(orw[alive], rr[alive], rtol=1e-6, atol=0)
        if not auto and rd.any():
            ref.reset(mask=rd); o.reset(rd.astype(np.uint8))
        assert np.array_equal(ref.goals()[alive], o.goal.T[alive])
        assert np.array_equal(ref.step_nums()[alive], o.step_num[alive])
    if not alive.all():   # the reference raised roboy_env.py:109 -> the oracle flagged the same env first
        assert o.errors()[0] & orc.ERR_REWARD_RANGE


def test_external_simulator_feed_matches_reference():
    """states fed in from outside (float64 wire values, as RosSimulationClient holds
    them) through the reference's own RoboyEnv vs the oracle's external path."""
    N, T = 10, 60
    rng = np.random.default_rng(3)
    ref = rh.ReferenceFeedEnv(N, seed=9)
    o = orc.OracleEnv(N, seed=9, auto_reset=False)
    assert np.array_equal(ref.goals().T, o.goal)
    q0 = rng.uniform(-3, 3, (N, 3)).astype(np.float32)
    qd0 = rng.uniform(-0.5, 0.5, (N, 3)).astype(np.float32)
    assert np.array_equal(ref.reset(q0, qd0).astype(np.float32), o.reset_external(q0, qd0))
    for i in range(N):
        ref.set_step_num(i, 380 + 3 * i)
    o.step_flags[:] = (o.step_flags & ~np.uint32(orc.STEP_MASK)) | (380 + 3 * np.arange(N)).astype(np.uint32)
    n_done = n_reached = 0
    for t in range(T):
        q = rng.uniform(-3, 3, (N, 3)).astype(np.float32)
        qd = rng.uniform(-0.5, 0.5, (N, 3)).astype(np.float32)
        near = rng.random(N) < 0.3                       # some states land next to the goal, slowly
        q[near] = np.clip(o.goal.T[near] + rng.uniform(-0.04, 0.04, (near.sum(), 3)), -3.1, 3.1).astype(np.float32)
        qd[near] = rng.uniform(-0.25, 0.25, (near.sum(), 3)).astype(np.float32)
        feas = rng.random(N) > 0.2
        ro, rr, rd = ref.step(q, qd, feas)
        oo, orw, od = o.step_external(q, qd, feas)
        assert np.array_equal(rd, od) and np.array_equal(ro.astype(np.float32), oo)
        assert np.allclose(orw, rr, rtol=1e-6, atol=0)
        n_done += rd.sum(); n_reached += (rd & (rr > 500)).sum()
        if rd.any():
            assert np.array_equal(ref.reset(q, qd, mask=rd).astype(np.float32)[rd], o.reset_external(q, qd, rd)[rd])
        assert np.array_equal(ref.goals().T, o.goal)
    assert n_reached >= 5 and n_done > n_reached


CUSTOM_ROBOTS = {
    "symmetric": dict(angle_low=-2.0, angle_high=2.0, vel_low=-0.7, vel_high=0.7, act_low=-0.5, act_high=0.5),
    # asymmetric spaces: normalisation no longer cancels, the hold interval is not centred
    "asymmetric": dict(angle_low=-1.0, angle_high=2.5, vel_low=-0.25, vel_high=0.75, act_low=-0.1, act_high=0.4),
}


@pytest.mark.parametrize("name", sorted(CUSTOM_ROBOTS))
def test_other_robot_bounds_match_reference(name):
    """a RoboyRobot plug-in with other spaces (same dims) through the reference."""
    b = CUSTOM_ROBOTS[name]
    N, T = 8, 120
    ref = rh.ReferenceVecEnv(N, seed=2, bounds=b)
    o = orc.OracleEnv(N, seed=2, **b)
    assert np.allclose(ref.reward_range, o.reward_range, rtol=1e-6, atol=0)
    assert np.array_equal(ref.goals().T, o.goal)
    assert np.array_equal(ref.reset().astype(np.float32), o.reset())
    for i in range(N):
        ref.set_step_num(i, 330 + 9 * i)
    o.step_flags[:] = (o.step_flags & ~np.uint32(orc.STEP_MASK)) | (330 + 9 * np.arange(N)).astype(np.uint32)
    rng = np.random.default_rng(1)
    alive = np.ones(N, bool)
    # the action that rescales to exactly 0 in the robot's space is where the Stub holds
    zero_action = np.float32(1 - 2 * b["act_high"] / (b["act_high"] - b["act_low"]))
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)
        a[rng.random(N) < 0.15] = zero_action
        ro, rr, rd, rt, raised = ref.step(a)
        oo, orw, od, ot = o.step(a, want_terminal_obs=True)
        alive &= np.array([m == "" for m in raised])
        assert np.array_equal(ro.astype(np.float32)[alive], oo[alive]), t
        assert np.array_equal(rd[alive], od[alive]), t
        assert np.allclose(orw[alive], rr[alive], rtol=1e-6, atol=0)
        assert np.array_equal(ref.goals()[alive], o.goal.T[alive])
    assert o.stats()["episodes"] >= N and o.stats()["holds"] > 0


# Robots with other joint / tendon counts and per-component bounds (roboy_robot.py:21-33,
# README.md:6-7 "Upper Body, etc."): synthetic RoboyRobot subclasses of the REFERENCE drive the unmodified RoboyEnv.
GENERIC_ROBOTS = {
    "six_joints_14_tendons": dict(dim_joint=6, dim_action=14, angle_low=-2.5, angle_high=2.5, vel_low=-0.6, vel_high=0.6,
                                  act_low=-0.2, act_high=0.2),
    "per_component_msj_dims": dict(angle_low=[-3.0, -1.5, -0.5], angle_high=[3.0, 2.0, 2.5],
                                   vel_low=[-0.5, -0.25, -1.0], vel_high=[0.5, 0.75, 1.0],
                                   act_low=[-0.3, -0.1, -0.2, -0.3, -0.05, -0.4, -0.3, -0.25],
                                   act_high=[0.3, 0.4, 0.2, 0.1, 0.05, 0.4, 0.6, 0.25]),
    "five_joints_11_tendons_per_component": dict(
        angle_low=[-3.1, -1.0, -2.0, -0.7, -1.3], angle_high=[3.1, 1.0, 2.5, 0.9, 1.3],
        vel_low=[-0.5, -0.4, -0.3, -0.2, -0.1], vel_high=[0.5, 0.4, 0.6, 0.2, 0.3],
        act_low=-np.linspace(0.1, 0.6, 11), act_high=np.linspace(0.15, 0.5, 11)),
    "fifteen_joints_64_tendons": dict(dim_joint=15, dim_action=64, angle_low=-np.linspace(1.0, 3.0, 15),
                                      angle_high=np.linspace(0.5, 3.1, 15), vel_low=-0.5, vel_high=0.5,
                                      act_low=-0.3, act_high=np.linspace(0.1, 0.9, 64)),
}


generic_zero_action = orc.hold_action


@pytest.mark.parametrize("penalty", [False, True])
@pytest.mark.parametrize("name", sorted(GENERIC_ROBOTS))
def test_generic_robots_match_reference(name, penalty):
    b = GENERIC_ROBOTS[name]
    J, A, _, bb = orc.robot_bounds(b)
    N, T = 6, 140
    ref = rh.ReferenceVecEnv(N, seed=5, bounds=b, joint_vel_penalty=penalty)
    o = orc.OracleEnv(N, seed=5, joint_vel_penalty=penalty, **b)
    assert (o.J, o.A) == (J, A) and ref.obs_dim == 3 * J
    assert np.allclose(ref.reward_range, o.reward_range, rtol=1e-6, atol=0)
    assert np.array_equal(ref.goals().T, o.goal)
    assert np.array_equal(ref.reset().astype(np.float32), o.reset())
    for i in range(N):
        ref.set_step_num(i, 300 + 17 * i)
    o.step_flags[:] = (o.step_flags & ~np.uint32(orc.STEP_MASK)) | (300 + 17 * np.arange(N)).astype(np.uint32)
    rng = np.random.default_rng(3)
    alive = np.ones(N, bool)
    zero_action, can_hold = generic_zero_action(b)
    thr_a, thr_v = orc.thresholds(o.cfg)
    reached = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
        hold = rng.random(N) < 0.25
        a[hold] = zero_action
        a[rng.random(N) < 0.05] = np.nextafter(zero_action, np.float32(1))     # right next to the hold interval
        if t % 9 == 4:
            # near-threshold goals: around the held float64 zero state (hold rows) and around the next sampled state
            q, _ = orc.draw_state(5, np.arange(N), o.counter + 1, bb["angle_low"], bb["angle_high"], J=J)
            d = rng.normal(size=(N, J)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            r = float(thr_a) * (1 + rng.choice([-1e-7, 1e-7, -0.3, 0.2], N))
            base = np.where(hold[:, None], 0.0, q.astype(np.float64))
            g = np.clip(base + d * r[:, None], bb["angle_low"], bb["angle_high"]).astype(np.float32)
            for i in range(N):
                ref.set_goal(i, g[i])
            o.goal[:] = g.T
        if t % 20 == 7:   # an injected float32 held state, some infeasible
            for i in range(N):
                q = rng.uniform(bb["angle_low"], bb["angle_high"]).astype(np.float32)
                qd = rng.uniform(bb["vel_low"], bb["vel_high"]).astype(np.float32) * np.float32(0.2)
                feas = bool(rng.random() < 0.6)
                ref.set_state(i, q, qd, feas)
                o.held[0:J, i] = q; o.held[J:2 * J, i] = qd
                o.step_flags[i] = (int(o.step_flags[i]) & orc.STEP_MASK) | (0 if feas else orc.F_HELD_INFEASIBLE)
        ro, rr, rd, rt, raised = ref.step(a)
        oo, orw, od, ot = o.step(a, want_terminal_obs=True)
        alive &= np.array([m == "" for m in raised])
        assert np.array_equal(ro.astype(np.float32)[alive], oo[alive]), t
        assert np.array_equal(rd[alive], od[alive]), t
        assert np.allclose(orw[alive], rr[alive], rtol=1e-6, atol=0), t
        assert np.array_equal(rt.astype(np.float32)[alive & rd], ot[alive & rd]), t
        assert np.array_equal(ref.goals()[alive], o.goal.T[alive])
        assert np.array_equal(ref.step_nums()[alive], o.step_num[alive])
        reached += int((rd & (rr > 500) & alive).sum())
    assert o.stats()["episodes"] >= N and (o.stats()["holds"] > 0) == can_hold and alive.sum() >= (1 if penalty else N)
    assert penalty or reached > 0 or not can_hold   # (a robot that can never hold reaches goals only by sampling)


def random_robot(rng, J, A=None):
    """A random RoboyRobot: J joints, A tendons, one (asymmetric) bound per component."""
    A = int(rng.integers(1, 65)) if A is None else A
    a_hi = rng.uniform(0.3, 3.1, J); a_lo = -rng.uniform(0.3, 3.1, J)
    v_hi = rng.uniform(0.1, 1.0, J); v_lo = -rng.uniform(0.1, 1.0, J)
    t_hi = rng.uniform(0.05, 0.9, A); t_lo = -rng.uniform(0.05, 0.9, A)
    if rng.random() < 0.5:   # some tendon ranges symmetric and dyadic: those robots CAN hold
        t_hi = np.full(A, rng.choice([0.25, 0.5, 0.125])); t_lo = -t_hi
    return dict(angle_low=a_lo, angle_high=a_hi, vel_low=v_lo, vel_high=v_hi, act_low=t_lo, act_high=t_hi)


@pytest.mark.parametrize("J", list(range(1, 16)))
def test_every_joint_count_matches_reference(J):
    """One random robot per joint count 1..15 (every instantiation of the CUDA generic step kernel has an oracle that is
    itself pinned against the unmodified reference)."""
    rng = np.random.default_rng(1000 + J)
    b = random_robot(rng, J)
    _, A, _, bb = orc.robot_bounds(b)
    N, T = 3, 45
    penalty = bool(J % 2)
    ref = rh.ReferenceVecEnv(N, seed=J, bounds=b, joint_vel_penalty=penalty)
    o = orc.OracleEnv(N, seed=J, joint_vel_penalty=penalty, **b)
    assert np.allclose(ref.reward_range, o.reward_range, rtol=1e-6, atol=0)
    assert np.array_equal(ref.goals().T, o.goal)
    assert np.array_equal(ref.reset().astype(np.float32), o.reset())
    zero_action, can_hold = orc.hold_action(b)
    thr_a, _ = orc.thresholds(o.cfg)
    alive = np.ones(N, bool)
    for i in range(N):
        ref.set_step_num(i, 380 + 10 * i)
    o.step_flags[:] = (o.step_flags & ~np.uint32(orc.STEP_MASK)) | (380 + 10 * np.arange(N)).astype(np.uint32)
    for t in range(T):
        a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
        hold = rng.random(N) < 0.3
        a[hold] = zero_action
        if t % 6 == 2:
            d = rng.normal(size=(N, J)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            g = np.clip(d * float(thr_a) * (1 + rng.choice([-1e-7, 1e-7, -0.4], N))[:, None], bb["angle_low"], bb["angle_high"]).astype(np.float32)
            for i in range(N):
                ref.set_goal(i, g[i])
            o.goal[:] = g.T
        ro, rr, rd, rt, raised = ref.step(a)
        oo, orw, od, ot = o.step(a, want_terminal_obs=True)
        alive &= np.array([m == "" for m in raised])
        assert np.array_equal(ro.astype(np.float32)[alive], oo[alive]), t
        assert np.array_equal(rd[alive], od[alive]), t
        assert np.allclose(orw[alive], rr[alive], rtol=1e-6, atol=0), t
        assert np.array_equal(ref.goals()[alive], o.goal.T[alive])
    assert o.stats()["episodes"] >= 1 and (o.stats()["holds"] > 0) == can_hold


def test_oracle_vs_reference_soak_slice():
    """An 8-second slice of tools/soak_oracle_vs_reference.py (random robots, flags, seeds; a master seed of its own; the
    long records are in profiles/r2_soak_oracle_vs_reference.json)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("soak_ovr", os.path.join(ROOT, "tools", "soak_oracle_vs_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    seed = 424242
    s = mod.soak(8.0, master_seed=seed)
    assert not s["mismatches"], (seed, s["mismatches"][:2])
    assert s["configs"] >= 3 and s["env_steps"] > 500


@pytest.mark.parametrize("name", ["msj", "five_joints_11_tendons_per_component"])
def test_external_state_bounds_match_the_reference_new_state(name):
    """roboy_robot.py:76 through the reference's client (ros_simulation_client.py:40-46 builds every received state with
    robot.new_state(python floats)): the envs whose new_state() raises are exactly the envs the oracle's external step
    flags with ERR_STATE_BOUNDS -- bounds themselves inside, one ulp outside out, NaN and inf out, velocities unchecked."""
    b = {} if name == "msj" else GENERIC_ROBOTS[name]
    J, _, _, bb = orc.robot_bounds(b)
    _, MsjRobot, _, _ = rh._import_reference()
    robot = (rh._custom_reference_robot(b) if b else MsjRobot)()
    n = 400
    rng = np.random.default_rng(21)
    lo, hi = bb["angle_low"], bb["angle_high"]
    q = rng.uniform(lo, hi, (n, J)).astype(np.float32)
    qd = (rng.uniform(bb["vel_low"], bb["vel_high"], (n, J)) * 5).astype(np.float32)
    pick = rng.random((n, J))
    q = np.where(pick < 0.03, np.nextafter(hi, np.float32(10)), q)
    q = np.where((pick >= 0.03) & (pick < 0.06), np.nextafter(lo, np.float32(-10)), q)
    q = np.where((pick >= 0.06) & (pick < 0.09), hi, q)
    q = np.where((pick >= 0.09) & (pick < 0.12), lo, q)
    q = np.where((pick >= 0.12) & (pick < 0.13), np.float32(np.nan), q)
    q = np.where((pick >= 0.13) & (pick < 0.14), np.float32(np.inf), q).astype(np.float32)
    raised = np.zeros(n, bool)
    for i in range(n):
        try:   # exactly RosSimulationClient._make_robot_state: lists of python floats
            robot.new_state(joint_angle=[float(x) for x in q[i]], joint_vel=[float(x) for x in qd[i]], is_feasible=True)
        except AssertionError:
            raised[i] = True
    assert 20 < raised.sum() < n - 20
    for reset in (True, False):
        o = orc.OracleEnv(n, seed=1, env_id_base=50, auto_reset=False, **b)
        if reset:
            o.reset_external(q, qd)
        else:
            o.step_external(q, qd)
        flags, first = o.errors()
        assert flags & orc.ERR_STATE_BOUNDS and first == 50 + int(np.flatnonzero(raised)[0])
        # without the offenders nothing is flagged
        ok = ~raised
        o2 = orc.OracleEnv(int(ok.sum()), seed=1, auto_reset=False, **b)
        if reset:
            o2.reset_external(q[ok], qd[ok])
        else:
            o2.step_external(q[ok], qd[ok])
        assert not (o2.errors()[0] & orc.ERR_STATE_BOUNDS)
        # and every single offender is flagged on its own
        bad = np.flatnonzero(raised)
        o3 = orc.OracleEnv(int(bad.size), seed=1, auto_reset=False, **b)
        if not reset:
            o3.step_external(q[bad], qd[bad])
            assert o3.stats()["violations"] == bad.size
