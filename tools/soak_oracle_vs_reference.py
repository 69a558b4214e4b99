"""Randomised differential soak of the CPU ORACLE against the UNMODIFIED reference (CPU only; needs /root/reference or
baseline/_ref): random robots (1..15 joints, 1..64 tendons, symmetric / asymmetric / one-sided spaces, MSJ itself), random
flags and seeds, N reference RoboyEnv(StubSimulationClient(ReplayRobot)) instances in lock-step with one OracleEnv on the
same injected Philox draws -- observations, done flags, goals, step counters and terminal observations bit-exact, rewards
within 1e-6, the reference's AssertionErrors (roboy_env.py:52 / :109) against the oracle's error word.  Holds, near-hold
actions, out-of-range and NaN actions, goals planted around the reached threshold (both sides, on sampled and on held
states), injected float32 held states (some infeasible), manual and automatic resets.
This is what the GPU soaks (tools/soak_parity.py, soak_generic.py, ...) lean on: they compare CUDA with the oracle.
usage: python tools/soak_oracle_vs_reference.py [seconds] [out.json]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import oracle as orc
from oracle import reference_harness as rh


def random_robot(rng):
    kind = int(rng.integers(0, 5))
    if kind == 0:
        return {}, "msj"
    J = int(rng.integers(1, 16))
    A = int(rng.integers(1, 65))
    a_hi = rng.uniform(0.3, 3.1, J); a_lo = -rng.uniform(0.3, 3.1, J)
    v_hi = rng.uniform(0.1, 1.0, J); v_lo = -rng.uniform(0.1, 1.0, J)
    t_hi = rng.uniform(0.05, 0.9, A); t_lo = -rng.uniform(0.05, 0.9, A)
    name = "asymmetric"
    if kind == 1:     # symmetric dyadic tendon ranges (a robot that CAN hold), symmetric spaces
        t_hi = np.full(A, rng.choice([0.25, 0.5, 0.125])); t_lo = -t_hi
        a_lo = -a_hi; v_lo = -v_hi
        name = "symmetric"
    elif kind == 2:   # one-sided spaces; a one-sided tendon range holds at a = -1, the edge of the action space
        a_lo = np.zeros(J); v_lo = np.zeros(J); t_lo = np.zeros(A)
        name = "one_sided"
    elif kind == 3:   # MSJ's dims with other scalar limits
        return dict(angle_low=-float(rng.uniform(0.3, 3.1)), angle_high=float(rng.uniform(0.3, 3.1)),
                    vel_low=-float(rng.uniform(0.1, 1.0)), vel_high=float(rng.uniform(0.1, 1.0)),
                    act_low=-float(rng.choice([0.25, 0.5, 0.1])), act_high=float(rng.choice([0.25, 0.5, 0.3]))), "msj_shaped"
    return dict(angle_low=a_lo, angle_high=a_hi, vel_low=v_lo, vel_high=v_hi, act_low=t_lo, act_high=t_hi), name


def soak(budget=120.0, master_seed=20261018):
    master = np.random.default_rng(master_seed)
    t0 = time.time()
    summary = {"configs": 0, "env_steps": 0, "robots": {}, "joint_counts": {}, "episodes": 0, "reached": 0, "holds": 0,
               "raised": 0, "worst_reward_rel": 0.0, "mismatches": []}
    while time.time() - t0 < budget:
        b, kind = random_robot(master)
        J, A, _, bb = orc.robot_bounds(b)
        N = int(master.choice([3, 6, 9]))
        T = int(master.integers(40, 160))
        seed = int(master.integers(0, 2 ** 31))
        base = int(master.choice([0, 7, 2 ** 33 + 5]))
        flags = dict(joint_vel_penalty=bool(master.integers(0, 2)), bonus=bool(master.integers(0, 2)), auto_reset=bool(master.integers(0, 2)))
        tag = dict(kind=kind, J=J, A=A, N=N, T=T, seed=seed, base=base, **flags)
        try:
            ref = rh.ReferenceVecEnv(N, seed=seed, env_id_base=base, bounds=b if b else None, **flags)
            o = orc.OracleEnv(N, seed=seed, env_id_base=base, **flags, **b)
            assert np.allclose(ref.reward_range, o.reward_range, rtol=1e-6, atol=0), "reward_range"
            assert np.array_equal(ref.goals().T, o.goal), "construction goals"
            assert np.array_equal(ref.reset().astype(np.float32), o.reset()), "reset obs"
            rng = np.random.default_rng(seed)
            steps = rng.integers(300, 400, N)
            for i in range(N):
                ref.set_step_num(i, int(steps[i]))
            o.step_flags[:] = (o.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
            zero_action, can_hold = orc.hold_action(b)
            thr_a, _ = orc.thresholds(o.cfg)
            alive = np.ones(N, bool)
            ref_bits = 0   # what the reference raised, as the oracle's error bits: 1 = :52 (no message), 2 = :109 ("not between")
            for t in range(T):
                a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
                hold = rng.random(N) < 0.2
                a[hold] = zero_action
                a[rng.random(N) < 0.05] = np.nextafter(zero_action, np.float32(1))
                if t % 11 == 5:
                    a[rng.integers(0, N), rng.integers(0, A)] = [np.nan, 1.5, -1.0000001][t % 3]   # roboy_env.py:52
                if t % 9 == 4:   # goals around the reached threshold: of the next sampled state, and of the held zero state
                    q, _ = orc.draw_state(seed, np.arange(base, base + N, dtype=np.uint64), o.counter + 1, bb["angle_low"], bb["angle_high"], J=J)
                    d = rng.normal(size=(N, J)); d /= np.linalg.norm(d, axis=1, keepdims=True)
                    r = float(thr_a) * (1 + rng.choice([-1e-7, 1e-7, -1e-5, 1e-5, -0.3, 0.2], N))
                    origin = np.where(hold[:, None], 0.0, q.astype(np.float64))
                    g = np.clip(origin + d * r[:, None], bb["angle_low"], bb["angle_high"]).astype(np.float32)
                    for i in range(N):
                        if alive[i]:
                            ref.set_goal(i, g[i])
                    o.goal[:, alive] = g[alive].T
                if t % 20 == 7:  # an injected float32 held state, some infeasible
                    for i in range(N):
                        if not alive[i]:
                            continue
                        q = rng.uniform(bb["angle_low"], bb["angle_high"]).astype(np.float32)
                        qd = (rng.uniform(bb["vel_low"], bb["vel_high"]) * 0.2).astype(np.float32)
                        feas = bool(rng.random() < 0.6)
                        ref.set_state(i, q, qd, feas)
                        o.held[0:J, i] = q; o.held[J:2 * J, i] = qd
                        o.step_flags[i] = (int(o.step_flags[i]) & orc.STEP_MASK) | (0 if feas else orc.F_HELD_INFEASIBLE)
                ro, rr, rd, rt, raised = ref.step(a)
                oo, orw, od, ot = o.step(a, want_terminal_obs=True)
                newly = np.array([m != "" for m in raised])
                for i in np.flatnonzero(newly & alive):
                    ref_bits |= 2 if "not between" in raised[i] else 1
                # every assert the reference raised this step is in the oracle's (sticky) error word by now
                assert (o.errors()[0] & ref_bits) == ref_bits, "error word %d lacks %d, step %d" % (o.errors()[0], ref_bits, t)
                summary["raised"] += int((newly & alive).sum())
                alive &= ~newly    # a reference env that raised has not advanced: it leaves the comparison
                if not alive.any():
                    break
                assert np.array_equal(ro.astype(np.float32)[alive], oo[alive]), "obs, step %d" % t
                assert np.array_equal(rd[alive], od[alive]), "done, step %d" % t
                rel = np.abs(orw[alive].astype(np.float64) - rr[alive]) / np.maximum(np.abs(rr[alive]), 1e-30)
                summary["worst_reward_rel"] = max(summary["worst_reward_rel"], float(rel.max()))
                assert rel.max() <= 1e-6, "reward rel %g, step %d" % (rel.max(), t)
                if flags["auto_reset"]:
                    assert np.array_equal(rt.astype(np.float32)[alive & rd], ot[alive & rd]), "terminal obs, step %d" % t
                else:
                    m = (rd & alive)
                    if m.any():
                        ref.reset(m); o.reset(m.astype(np.uint8))
                assert np.array_equal(ref.goals()[alive], o.goal.T[alive]), "goals, step %d" % t
                assert np.array_equal(ref.step_nums()[alive], o.step_num[alive]), "step counters, step %d" % t
                summary["env_steps"] += int(alive.sum())
                summary["reached"] += int((rd & (rr > 500) & alive).sum())
            # the oracle's error word against what the reference raised
            if alive.all():   # nobody raised: the oracle must not have flagged anything either
                assert o.errors()[0] == 0, "oracle error word %d although the reference never raised" % o.errors()[0]
            st = o.stats()
            summary["episodes"] += int(st["episodes"]); summary["holds"] += int(st["holds"])
        except AssertionError as exc:
            summary["mismatches"].append(dict(tag, error=str(exc)))
        summary["configs"] += 1
        summary["robots"][kind] = summary["robots"].get(kind, 0) + 1
        summary["joint_counts"][str(J)] = summary["joint_counts"].get(str(J), 0) + 1
    summary["seconds"] = time.time() - t0
    return summary


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    s = soak(budget, master_seed=int(os.environ.get("ROBOY_SOAK_SEED", "20261018")))
    print(json.dumps(s))
    if len(sys.argv) > 2:
        json.dump(s, open(sys.argv[2], "w"), indent=1)
    sys.exit(1 if s["mismatches"] else 0)
