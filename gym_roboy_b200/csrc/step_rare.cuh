// Rare, divergent paths of one env-step, shared by the step kernels (roboy_kernels.cu) and the fused
// policy rollout kernel (roboy_policy.cu): the Stub's hold branch and the end of an episode.  Both are
// kept out of line so they stay out of the hot paths' register budgets.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/roboy_b200.h"
#include "roboy_kernels.cuh"

namespace roboy {
namespace {

struct HoldOut {
    float q0, q1, q2, qd0, qd1, qd2, reward;
    uint32_t flags;  // bit 0 reached, bit 1 violation
};

// Hold branch of the Stub (simulation_client.py:38-39): the stored state is returned.  Rare and
// divergent, so it is kept out of line and out of the hot path's register budget.
static __device__ __noinline__ HoldOut hold_branch(const StepParams &p, uint32_t e, uint32_t sf, float g0, float g1, float g2,
                                            bool penalty, bool bonus) {
    HeldState s;
    if (sf & ROBOY_F_HELD_ZERO64) {
#pragma unroll
        for (int k = 0; k < 3; ++k) s.q[k] = s.qd[k] = 0.0;
        s.is64 = true;
        s.feasible = true;
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s.q[k] = (double)p.held[(size_t)k * p.n + e];
            s.qd[k] = (double)p.held[(size_t)(3 + k) * p.n + e];
        }
        s.is64 = false;
        s.feasible = !(sf & ROBOY_F_HELD_INFEASIBLE);
    }
    double r;
    bool reached, violation;
    const float g[3] = {g0, g1, g2};
    const float no_gqd[3] = {0.f, 0.f, 0.f};
    reward_reached_general(s, g, false, no_gqd, penalty, bonus, p.c, r, reached, violation);
    HoldOut o;
    o.q0 = (float)s.q[0]; o.q1 = (float)s.q[1]; o.q2 = (float)s.q[2];
    o.qd0 = (float)s.qd[0]; o.qd1 = (float)s.qd[1]; o.qd2 = (float)s.qd[2];
    o.reward = (float)r;
    o.flags = (reached ? 1u : 0u) | (violation ? 2u : 0u);
    return o;
}

// Rare: an env finished its episode (:65-68) -- new goal, and under auto-reset the worker's
// reset() (:82-87).  Out of line: ~1/400 of env-steps.
static __device__ __noinline__ uint32_t finish_episode(const StepParams &p, uint64_t t, uint32_t e, uint32_t step,
                                                bool reached, bool auto_reset, float *row, unsigned int *s_cnt) {
    const uint64_t gid = p.gid_base + e;
    // Under auto-reset the reference draws twice (:68 then :86) and only the second goal is ever
    // observable, so a single draw is materialised.
    const uint4 rg = philox_draw(gid, t, kStreamGoal, p.keys);
    const float ng0 = uniform_in24(rg.x, p.c.a_lo, p.f.a_span24);
    const float ng1 = uniform_in24(rg.y, p.c.a_lo, p.f.a_span24);
    const float ng2 = uniform_in24(rg.z, p.c.a_lo, p.f.a_span24);
    p.goal[e] = ng0;
    p.goal1[e] = ng1;
    p.goal2[e] = ng2;
    atomicAdd(&s_cnt[0], 1u);
    if (reached) atomicAdd(&s_cnt[1], 1u);
    atomicAdd(&s_cnt[4], step - 1);              // episode length (steps taken since reset(): step_num starts at 1)
    if (!auto_reset) return step | 0x80000000u;  // top bit: keep the flag bits
    if (p.terminal_obs) {
#pragma unroll
        for (int k = 0; k < kObsDim; ++k) p.terminal_obs[(size_t)e * kObsDim + k] = row[k];
    }
    row[0] = row[1] = row[2] = row[3] = row[4] = row[5] = 0.0f;  // reset(): zero state, :83-84,87
    row[6] = ng0; row[7] = ng1; row[8] = ng2;
    return 1u;                                                   // :85, with the flags replaced
}

// The same episode end, worked off after the CTA's hot loop from its shared-memory queue (closed-loop step kernel): the
// observation row of the env is already in global memory (`obs`), so the terminal observation is copied from there and
// the reset observation is written over it.
static __device__ __noinline__ void finish_episode_queued(const StepParams &p, uint64_t t, uint32_t e, uint32_t step,
                                                          bool reached, bool auto_reset, float *obs, unsigned int *s_cnt) {
    const uint4 rg = philox_draw(p.gid_base + e, t, kStreamGoal, p.keys);
    const float ng0 = uniform_in24(rg.x, p.c.a_lo, p.f.a_span24);
    const float ng1 = uniform_in24(rg.y, p.c.a_lo, p.f.a_span24);
    const float ng2 = uniform_in24(rg.z, p.c.a_lo, p.f.a_span24);
    p.goal[e] = ng0;
    p.goal1[e] = ng1;
    p.goal2[e] = ng2;
    atomicAdd(&s_cnt[0], 1u);
    if (reached) atomicAdd(&s_cnt[1], 1u);
    atomicAdd(&s_cnt[4], step - 1);
    if (!auto_reset) return;
    float *row = obs + (size_t)e * kObsDim;
    if (p.terminal_obs) {
#pragma unroll
        for (int k = 0; k < kObsDim; ++k) p.terminal_obs[(size_t)e * kObsDim + k] = row[k];
    }
    row[0] = row[1] = row[2] = row[3] = row[4] = row[5] = 0.0f;  // reset(): zero state, :83-84,87
    row[6] = ng0; row[7] = ng1; row[8] = ng2;
}

}  // namespace
}  // namespace roboy
