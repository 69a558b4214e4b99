// sm_100a kernels of the batched MSJ environment step.
//
// K1 step_kernel   RoboyEnv.step (envs/roboy_env.py:51-70) fused with the Stub simulation update
//                  (envs/simulations/simulation_client.py:36-40), normalisation
//                  (envs/robots/roboy_robot.py:80-95), compute_reward (:92-112), the done test
//                  (:65-66,:125-134), goal resampling (:117-123) and the vec-env reset-on-done.
// K2 init_or_reset RoboyEnv.__init__ / reset (:12-38, :82-87) over the Stub (:29-31, :42-44).
// K3 episode statistics are folded into K1's tail (warp reduce -> one atomic set per CTA).
//
// The path is elementwise and HBM-bound (93 algorithmic bytes, ~0.6 flop/B): no tensor cores.
// What matters is that every global access is a full-sector coalesced access:
//   * actions [n][8]: each warp reads its 32 envs' 1 KiB as 2 x LDG.128 per lane.  The Stub only
//     ever REDUCES the action (range assert + "all close to zero"), so no transpose is needed:
//     every lane tests the float4 it loaded and two warp ballots hand each env its verdict.
//   * goal / step_flags: structure of arrays, one 128 B line per warp access.
//   * obs [n][9] row-major (36 B rows): staged through shared memory so a warp emits its
//     1152 contiguous bytes as 72 x STG.128.
// The grid is persistent (a multiple of the SM count); statistics live in registers across the
// grid-stride loop, so the atomics are per CTA, not per env.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/roboy_b200.h"
#include "roboy_kernels.cuh"

namespace roboy {

namespace {

constexpr uint32_t kFull = 0xffffffffu;

__device__ __forceinline__ float4 ld_stream4(const float4 *p) { return __ldcs(p); }

// |fl(fl(slope*fl(a - in_hi)) + act_hi)| <= hold_tol  <=> np.allclose(rescaled, 0) for this
// component (roboy_env.py:157-158 then simulation_client.py:38; hold_tol is the largest float32
// not above numpy's atol 1e-8, so the float32 compare equals numpy's float64 one).
struct ActionTest {
    float in_hi, in_lo, slope, act_hi, hold_tol;
    __device__ __forceinline__ bool ok(float a) const { return a >= in_lo && a <= in_hi; }  // NaN -> false
    __device__ __forceinline__ bool hold(float a) const {
        const float r = __fadd_rn(__fmul_rn(slope, __fsub_rn(a, in_hi)), act_hi);
        return fabsf(r) <= hold_tol;  // NaN -> false
    }
    __device__ __forceinline__ bool ok4(const float4 &v) const { return ok(v.x) && ok(v.y) && ok(v.z) && ok(v.w); }
    __device__ __forceinline__ bool hold4(const float4 &v) const {
        return hold(v.x) && hold(v.y) && hold(v.z) && hold(v.w);
    }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Hold branch of the Stub (simulation_client.py:38-39): the stored state is returned.  Rare and
// divergent, so it is kept out of line and out of the hot path's register budget.
__device__ __noinline__ void hold_branch(const StepParams &p, uint64_t e, uint32_t sf, const float g[3],
                                         bool penalty, bool bonus, float q[3], float qd[3], float &reward,
                                         bool &reached, bool &violation) {
    HeldState s;
    if (sf & ROBOY_F_HELD_ZERO64) {
#pragma unroll
        for (int k = 0; k < 3; ++k) s.q[k] = s.qd[k] = 0.0;
        s.is64 = true;
        s.feasible = true;
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s.q[k] = (double)p.held[(uint64_t)k * p.n + e];
            s.qd[k] = (double)p.held[(uint64_t)(3 + k) * p.n + e];
        }
        s.is64 = false;
        s.feasible = !(sf & ROBOY_F_HELD_INFEASIBLE);
    }
    double r;
    const float no_gqd[3] = {0.f, 0.f, 0.f};
    reward_reached_general(s, g, false, no_gqd, penalty, bonus, p.c, r, reached, violation);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        q[k] = (float)s.q[k];
        qd[k] = (float)s.qd[k];
    }
    reward = (float)r;
}

}  // namespace

template <bool PENALTY, bool BONUS, bool AUTO_RESET>
__global__ void __launch_bounds__(kStepBlock) step_kernel(const __grid_constant__ StepParams p) {
    __shared__ __align__(16) float s_obs[kWarpsPerBlock][32 * kObsDim];
    __shared__ double s_red[kWarpsPerBlock][ROBOY_STAT_COUNT];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint64_t n = p.n;
    const uint64_t n_end = p.e_end;
    const uint64_t chunk0 = p.e_begin >> 5;
    const uint64_t n_chunks = (n_end + 31) >> 5;
    const uint64_t warp_stride = (uint64_t)gridDim.x * kWarpsPerBlock;
    const float hold_tol = (double)1e-8f > 1e-8 ? __uint_as_float(__float_as_uint(1e-8f) - 1u) : 1e-8f;
    const ActionTest at{p.act_in_hi, p.act_in_lo, p.act_slope, p.act_hi, hold_tol};

    uint32_t n_steps = 0, n_done = 0, n_succ = 0, n_hold = 0, n_viol = 0, sum_eplen = 0;
    double sum_reward = 0.0;

    for (uint64_t chunk = chunk0 + (uint64_t)blockIdx.x * kWarpsPerBlock + warp; chunk < n_chunks; chunk += warp_stride) {
        const uint64_t base = chunk << 5;
        const uint64_t e = base + lane;
        const bool full = base + 32 <= n_end;
        const bool live = e < n_end;

        // ---- loads: 2 x 16 B of actions, 3 x 4 B goal, 4 B step word, all issued up front ----
        const float4 *a4 = reinterpret_cast<const float4 *>(p.actions) + base * 2;
        float4 A0, A1;
        bool v0 = true, v1 = true;
        if (full) {
            A0 = ld_stream4(a4 + lane);
            A1 = ld_stream4(a4 + 32 + lane);
        } else {
            v0 = base * 2 + lane < n_end * 2;
            v1 = base * 2 + 32 + lane < n_end * 2;
            A0 = v0 ? ld_stream4(a4 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            A1 = v1 ? ld_stream4(a4 + 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float g[3] = {0.f, 0.f, 0.f};
        uint32_t sf = 1u;
        if (live) {
            g[0] = p.goal[e];
            g[1] = p.goal[n + e];
            g[2] = p.goal[2 * n + e];
            sf = p.step_flags[e];
        }

        // ---- roboy_env.py:52 assert + simulation_client.py:38 allclose, as warp ballots ----
        const uint32_t okm0 = __ballot_sync(kFull, at.ok4(A0) || !v0);
        const uint32_t okm1 = __ballot_sync(kFull, at.ok4(A1) || !v1);
        const uint32_t hdm0 = __ballot_sync(kFull, at.hold4(A0) && v0);
        const uint32_t hdm1 = __ballot_sync(kFull, at.hold4(A1) && v1);
        const uint32_t sh = (lane & 15) << 1;  // env `lane` owns float4 2*lane and 2*lane+1 of the chunk
        const bool act_ok = (((lane < 16 ? okm0 : okm1) >> sh) & 3u) == 3u;
        const bool hold = live && (((lane < 16 ? hdm0 : hdm1) >> sh) & 3u) == 3u;

        const uint64_t gid = p.gid_base + e;
        float q[3], qd[3], reward;
        bool reached, violation;
        if (!hold) {
            // simulation_client.py:40 -> roboy_robot.py:35-39: fresh sample; velocities are drawn
            // from the ANGLE space too (reference quirk, :38)
            const uint4 ra = philox_draw(gid, p.t, kStreamStateQ, p.keys);
            const uint4 rb = philox_draw(gid, p.t, kStreamStateQd, p.keys);
            q[0] = uniform_in(ra.x, p.c.a_lo, p.c.a_span);
            q[1] = uniform_in(ra.y, p.c.a_lo, p.c.a_span);
            q[2] = uniform_in(ra.z, p.c.a_lo, p.c.a_span);
            qd[0] = uniform_in(rb.x, p.c.a_lo, p.c.a_span);
            qd[1] = uniform_in(rb.y, p.c.a_lo, p.c.a_span);
            qd[2] = uniform_in(rb.z, p.c.a_lo, p.c.a_span);
            reward_reached_sampled<PENALTY, BONUS>(q, qd, g, p.c, reward, reached, violation);
        } else {
            hold_branch(p, e, sf, g, PENALTY, BONUS, q, qd, reward, reached, violation);
            ++n_hold;
        }

        uint32_t step = sf & ROBOY_STEP_MASK;
        step += step < ROBOY_STEP_MASK;                       // roboy_env.py:60
        const bool timeout = (int32_t)step > p.max_len;       // :72-73
        const bool done = reached || timeout;                 // :65-66
        uint32_t flags = sf & ~ROBOY_STEP_MASK;

        float o[kObsDim] = {q[0], q[1], q[2], qd[0], qd[1], qd[2], g[0], g[1], g[2]};  // :75-80

        if (done && live) {
            // :67-68 new goal.  Under auto-reset the worker's reset() (:82-87) draws once more and
            // only that goal is ever observable, so a single draw is materialised.
            const uint4 rg = philox_draw(gid, p.t, kStreamGoal, p.keys);
            const float ng0 = uniform_in(rg.x, p.c.a_lo, p.c.a_span);
            const float ng1 = uniform_in(rg.y, p.c.a_lo, p.c.a_span);
            const float ng2 = uniform_in(rg.z, p.c.a_lo, p.c.a_span);
            p.goal[e] = ng0;
            p.goal[n + e] = ng1;
            p.goal[2 * n + e] = ng2;
            if (AUTO_RESET) {
                if (p.terminal_obs) {
#pragma unroll
                    for (int k = 0; k < kObsDim; ++k) p.terminal_obs[e * kObsDim + k] = o[k];
                }
                o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.0f;  // reset(): zero state, :83-84,87
                o[6] = ng0;
                o[7] = ng1;
                o[8] = ng2;
                sum_eplen += step - 1;
                step = 1;                                        // :85
                flags = ROBOY_F_HELD_ZERO64;
            }
            ++n_done;
            n_succ += reached;
        }

        bool bad = live && (violation || !act_ok);
        if (bad) {
            atomicOr(p.err_flags, (violation ? ROBOY_ERR_REWARD_RANGE : 0u) | (!act_ok ? ROBOY_ERR_ACTION : 0u));
            atomicMin(p.first_bad, (unsigned long long)gid);
            ++n_viol;
        }

        // ---- stores ----
        float *so = s_obs[warp];
#pragma unroll
        for (int k = 0; k < kObsDim; ++k) so[lane * kObsDim + k] = o[k];  // stride 9: conflict-free
        if (live) {
            p.step_flags[e] = step | flags;
            __stcs(p.reward + e, reward);
            p.done[e] = (uint8_t)done;
            ++n_steps;
            sum_reward += (double)reward;
        }
        __syncwarp();
        if (full) {
            float4 *dst = reinterpret_cast<float4 *>(p.obs + base * kObsDim);  // 1152 B per chunk: 16 B aligned
            const float4 *src = reinterpret_cast<const float4 *>(so);
            __stcs(dst + lane, src[lane]);
            __stcs(dst + 32 + lane, src[32 + lane]);
            if (lane < 8) __stcs(dst + 64 + lane, src[64 + lane]);
        } else {
            const uint32_t n_valid = (uint32_t)(n_end - base) * kObsDim;
            for (uint32_t i = lane; i < n_valid; i += 32) p.obs[base * kObsDim + i] = so[i];
        }
        __syncwarp();
    }

    // ---- K3: episode statistics, one set of atomics per CTA ----
    const double vals[ROBOY_STAT_COUNT] = {
        (double)__reduce_add_sync(kFull, n_steps), (double)__reduce_add_sync(kFull, n_done),
        (double)__reduce_add_sync(kFull, n_succ),  (double)__reduce_add_sync(kFull, n_done - n_succ),
        warp_sum(sum_reward),                      (double)__reduce_add_sync(kFull, sum_eplen),
        (double)__reduce_add_sync(kFull, n_hold),  (double)__reduce_add_sync(kFull, n_viol)};
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < ROBOY_STAT_COUNT; ++k) s_red[warp][k] = vals[k];
    }
    __syncthreads();
    if (threadIdx.x < ROBOY_STAT_COUNT) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) acc += s_red[w][threadIdx.x];
        if (acc != 0.0) atomicAdd(p.stats + threadIdx.x, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
namespace {

template <bool P, bool B, bool A>
int blocks_per_sm() {
    static int cached = 0;
    if (!cached) {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, step_kernel<P, B, A>, kStepBlock, 0) != cudaSuccess || b < 1)
            b = 1;
        cached = b;
    }
    return cached;
}

int grid_for(uint64_t n_range, int per_sm, int sm_count) {
    const uint64_t n_chunks = (n_range + 31) / 32;
    const uint64_t want = (n_chunks + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const uint64_t cap = (uint64_t)sm_count * per_sm;  // persistent: a multiple of the SM count
    return (int)(want < cap ? want : cap);
}

template <bool P, bool B, bool A>
cudaError_t launch_step_t(const StepParams &p, int sm_count, cudaStream_t stream) {
    const int grid = grid_for(p.e_end - p.e_begin, blocks_per_sm<P, B, A>(), sm_count);
    step_kernel<P, B, A><<<grid, kStepBlock, 0, stream>>>(p);
    return cudaGetLastError();
}

int blocks_per_sm_sel(int sel) {
    switch (sel) {
        case 0: return blocks_per_sm<false, false, false>();
        case 1: return blocks_per_sm<false, false, true>();
        case 2: return blocks_per_sm<false, true, false>();
        case 3: return blocks_per_sm<false, true, true>();
        case 4: return blocks_per_sm<true, false, false>();
        case 5: return blocks_per_sm<true, false, true>();
        case 6: return blocks_per_sm<true, true, false>();
        default: return blocks_per_sm<true, true, true>();
    }
}

}  // namespace

LaunchGeom step_geometry(uint64_t n_range, bool penalty, bool bonus, bool auto_reset, int sm_count) {
    const int sel = (penalty ? 4 : 0) | (bonus ? 2 : 0) | (auto_reset ? 1 : 0);
    return LaunchGeom{grid_for(n_range, blocks_per_sm_sel(sel), sm_count), kStepBlock,
                      (int)(sizeof(float) * kWarpsPerBlock * 32 * kObsDim + sizeof(double) * kWarpsPerBlock * ROBOY_STAT_COUNT)};
}

cudaError_t launch_step(const StepParams &p, bool penalty, bool bonus, bool auto_reset, int sm_count,
                        cudaStream_t stream) {
    if (p.e_end <= p.e_begin) return cudaSuccess;
    const int sel = (penalty ? 4 : 0) | (bonus ? 2 : 0) | (auto_reset ? 1 : 0);
    switch (sel) {
        case 0: return launch_step_t<false, false, false>(p, sm_count, stream);
        case 1: return launch_step_t<false, false, true>(p, sm_count, stream);
        case 2: return launch_step_t<false, true, false>(p, sm_count, stream);
        case 3: return launch_step_t<false, true, true>(p, sm_count, stream);
        case 4: return launch_step_t<true, false, false>(p, sm_count, stream);
        case 5: return launch_step_t<true, false, true>(p, sm_count, stream);
        case 6: return launch_step_t<true, true, false>(p, sm_count, stream);
        default: return launch_step_t<true, true, true>(p, sm_count, stream);
    }
}

// ---------------------------------------------------------------------------------------------
// K2: construction and reset
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) init_or_reset_kernel(const __grid_constant__ InitParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.n; e += stride) {
        if (p.mask && !p.mask[e]) continue;
        const uint64_t gid = p.gid_base + e;
        const uint4 rg = philox_draw(gid, p.t, kStreamGoal, p.keys);
        const float g0 = uniform_in(rg.x, p.a_lo, p.a_span);
        const float g1 = uniform_in(rg.y, p.a_lo, p.a_span);
        const float g2 = uniform_in(rg.z, p.a_lo, p.a_span);
        p.goal[e] = g0;
        p.goal[p.n + e] = g1;
        p.goal[2 * p.n + e] = g2;
        if (p.held) {
            // StubSimulationClient.__init__ (simulation_client.py:31): _state = new_random_state()
            const uint4 ra = philox_draw(gid, p.t, kStreamStateQ, p.keys);
            const uint4 rb = philox_draw(gid, p.t, kStreamStateQd, p.keys);
            p.held[e] = uniform_in(ra.x, p.a_lo, p.a_span);
            p.held[p.n + e] = uniform_in(ra.y, p.a_lo, p.a_span);
            p.held[2 * p.n + e] = uniform_in(ra.z, p.a_lo, p.a_span);
            p.held[3 * p.n + e] = uniform_in(rb.x, p.a_lo, p.a_span);
            p.held[4 * p.n + e] = uniform_in(rb.y, p.a_lo, p.a_span);
            p.held[5 * p.n + e] = uniform_in(rb.z, p.a_lo, p.a_span);
            p.step_flags[e] = 1u;  // roboy_env.py:38
        } else {
            // forward_reset_command (simulation_client.py:42-44): _state = float64 zero state
            p.step_flags[e] = 1u | ROBOY_F_HELD_ZERO64;  // roboy_env.py:85
        }
        if (p.obs) {
            float *o = p.obs + e * kObsDim;
            o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.0f;
            o[6] = g0;
            o[7] = g1;
            o[8] = g2;
        }
    }
}

cudaError_t launch_init_or_reset(const InitParams &p, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const uint64_t want = (p.n + 255) / 256;
    const int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    init_or_reset_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Stand-alone compute_reward / _did_reach_goal
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compute_reward_kernel(const __grid_constant__ RewardParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.k; i += stride) {
        HeldState s;
        float g[3], gqd[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s.q[k] = (double)p.q[i * 3 + k];
            s.qd[k] = (double)p.qd[i * 3 + k];
            g[k] = p.goal_q[i * 3 + k];
            if (p.goal_qd) gqd[k] = p.goal_qd[i * 3 + k];
        }
        s.is64 = false;
        s.feasible = p.feasible ? p.feasible[i] != 0 : true;
        RobotConsts c = p.c;
        if (!p.check_range) {
            c.reward_lo = -INFINITY;
            c.reward_hi = INFINITY;
        }
        double r;
        bool reached, violation;
        reward_reached_general(s, g, p.goal_qd != nullptr, gqd, p.penalty, p.bonus, c, r, reached, violation);
        p.reward[i] = r;
        if (p.reached) p.reached[i] = (uint8_t)reached;
        if (violation) {
            atomicOr(p.err_flags, ROBOY_ERR_REWARD_RANGE);
            atomicMin(p.first_bad, (unsigned long long)(p.gid_base + i));
            atomicAdd(p.stats + ROBOY_STAT_VIOLATIONS, 1.0);
        }
    }
}

cudaError_t launch_compute_reward(const RewardParams &p, cudaStream_t stream) {
    if (p.k == 0) return cudaSuccess;
    const uint64_t want = (p.k + 255) / 256;
    const int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    compute_reward_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Indexed state injection / read-back
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_kernel(const __grid_constant__ ScatterParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.k; i += stride) {
        const int64_t e64 = p.idx ? p.idx[i] : (int64_t)i;
        if (e64 < 0 || (uint64_t)e64 >= p.n) continue;
        const uint64_t e = (uint64_t)e64;
        if (p.goal_q) {
            bool inside = true;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float v = p.goal_q[i * 3 + k];
                inside = inside && (v >= p.a_lo && v <= p.a_hi);  // roboy_robot.py:76
                p.goal[(uint64_t)k * p.n + e] = v;
            }
            if (!inside) {
                atomicOr(p.err_flags, ROBOY_ERR_GOAL_BOUNDS);
                atomicMin(p.first_bad, (unsigned long long)(p.gid_base + e));
            }
        }
        if (p.q) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                p.held[(uint64_t)k * p.n + e] = p.q[i * 3 + k];
                p.held[(uint64_t)(3 + k) * p.n + e] = p.qd[i * 3 + k];
            }
            uint32_t sf = p.step_flags[e] & ROBOY_STEP_MASK;
            if (p.feasible && !p.feasible[i]) sf |= ROBOY_F_HELD_INFEASIBLE;
            p.step_flags[e] = sf;
        }
        if (p.step) {
            const uint32_t s = (uint32_t)p.step[i] & ROBOY_STEP_MASK;
            p.step_flags[e] = (p.step_flags[e] & ~ROBOY_STEP_MASK) | s;
        }
        if (p.out_q) {  // SimulationClient.read_state
            const uint32_t sf = p.step_flags[e];
            const bool z = sf & ROBOY_F_HELD_ZERO64;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                p.out_q[i * 3 + k] = z ? 0.0f : p.held[(uint64_t)k * p.n + e];
                p.out_qd[i * 3 + k] = z ? 0.0f : p.held[(uint64_t)(3 + k) * p.n + e];
            }
            if (p.out_feasible) p.out_feasible[i] = z ? 1 : !(sf & ROBOY_F_HELD_INFEASIBLE);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Un-fused SimulationClient calls (the plug-in API the reference's own RoboyEnv drives)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sim_kernel(const __grid_constant__ SimParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.n; e += stride) {
        const uint64_t gid = p.gid_base + e;
        if (p.mode == 2) {
            const uint4 rg = philox_draw(gid, p.t, kStreamGoal, p.keys, p.sub);
            p.out_q[e * 3 + 0] = uniform_in(rg.x, p.a_lo, p.a_span);
            p.out_q[e * 3 + 1] = uniform_in(rg.y, p.a_lo, p.a_span);
            p.out_q[e * 3 + 2] = uniform_in(rg.z, p.a_lo, p.a_span);
            continue;
        }
        uint32_t sf = p.step_flags[e];
        bool hold = true;
        if (p.mode == 1) {
            if (p.mask && !p.mask[e]) continue;
            sf = (sf & ROBOY_STEP_MASK) | ROBOY_F_HELD_ZERO64;  // _state = new_zero_state()
            p.step_flags[e] = sf;
        } else {
            // simulation_client.py:38 np.allclose(action, 0): |a| <= 1e-8 in float64, NaN fails
            const float4 a0 = reinterpret_cast<const float4 *>(p.actions)[e * 2];
            const float4 a1 = reinterpret_cast<const float4 *>(p.actions)[e * 2 + 1];
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) hold = hold && (fabs((double)a[k]) <= 1e-8);
        }
        float q[3], qd[3];
        bool feasible = true;
        if (hold) {
            const bool z = sf & ROBOY_F_HELD_ZERO64;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                q[k] = z ? 0.0f : p.held[(uint64_t)k * p.n + e];
                qd[k] = z ? 0.0f : p.held[(uint64_t)(3 + k) * p.n + e];
            }
            feasible = z || !(sf & ROBOY_F_HELD_INFEASIBLE);
            if (p.mode == 0) atomicAdd(p.stats + ROBOY_STAT_HOLDS, 1.0);
        } else {
            const uint4 ra = philox_draw(gid, p.t, kStreamStateQ, p.keys);
            const uint4 rb = philox_draw(gid, p.t, kStreamStateQd, p.keys);
            q[0] = uniform_in(ra.x, p.a_lo, p.a_span);
            q[1] = uniform_in(ra.y, p.a_lo, p.a_span);
            q[2] = uniform_in(ra.z, p.a_lo, p.a_span);
            qd[0] = uniform_in(rb.x, p.a_lo, p.a_span);
            qd[1] = uniform_in(rb.y, p.a_lo, p.a_span);
            qd[2] = uniform_in(rb.z, p.a_lo, p.a_span);
        }
        if (p.out_q) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                p.out_q[e * 3 + k] = q[k];
                p.out_qd[e * 3 + k] = qd[k];
            }
            if (p.out_feasible) p.out_feasible[e] = (uint8_t)feasible;
        }
    }
}

cudaError_t launch_sim(const SimParams &p, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const uint64_t want = (p.n + 255) / 256;
    const int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    sim_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_scatter(const ScatterParams &p, cudaStream_t stream) {
    if (p.k == 0) return cudaSuccess;
    const uint64_t want = (p.k + 255) / 256;
    const int grid = (int)(want < 148ull * 8 ? want : 148ull * 8);
    scatter_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace roboy
