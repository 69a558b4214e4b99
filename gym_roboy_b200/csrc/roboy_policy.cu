// sm_100a kernel of the closed-loop rollout: the policy INSIDE the env-step kernel.
//
// The reference collects rollouts with stable-baselines' PPO2 runner over a SubprocVecEnv
// (gym_roboy/train_parallel.py:28-35): per step the MlpPolicy (two 64-64 tanh networks, Gaussian
// with a state-independent log-std) runs in the trainer process, the action is clipped to the action
// space and crosses a pipe to every env process, and obs / reward / done come back the same way.
// Here ONE launch runs T such steps for every env: a thread owns E envs, keeps their observation,
// goal and step word in registers across the T steps, evaluates both networks in float32
// (weights in shared memory, FFMA2 = fma.rn.f32x2), draws the Gaussian noise from Philox, clips, and
// runs the same env step as step_kernel (RoboyEnv.step, roboy_env.py:51-70, with the Stub update,
// reward, done test, goal resampling and reset-on-done) on the action it just produced.  Nothing
// is read from HBM per step; what is written is what a PPO update consumes:
// action 32 + logp 4 + value 4 + obs 36 + reward 4 + done 1 = 81 B per env-step.
//
// Unlike the env step this kernel is bound by float32 FMA throughput (10,368 FMA per env-step for
// the two networks, i.e. ~21 kflop against 81 B), not by HBM.  The matrix products are done per
// thread ("one env = one row"), with the weight row broadcast from shared memory as LDS.128 and
// the activations of the previous layer read back from a thread-private shared-memory column, so
// that the 64 accumulators per env stay in registers; E = 2 envs per thread halves the
// shared-memory traffic per FMA (17 LDS per 64 FFMA2), which is what keeps the FMA pipe fed.
//
// The env arithmetic (done mask, observations, rewards) is the step kernel's, instruction for
// instruction: given the actions this kernel stores, T roboy_step calls reproduce its env outputs
// bit for bit (tests/test_gpu_policy_rollout.py).  The policy arithmetic is float32 with FMA and a
// tanh evaluated as 1 - 2/(exp2(2x*log2 e) + 1) on the MUFU unit (absolute error ~3e-7); it agrees
// with the torch float32 MlpPolicy to ~1e-6.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/roboy_b200.h"
#include "roboy_kernels.cuh"
#include "roboy_policy.cuh"
#include "step_rare.cuh"

namespace roboy {

namespace {

constexpr uint32_t kFullMask = 0xffffffffu;
constexpr int kHid = ROBOY_POLICY_HIDDEN;  // 64

__device__ __forceinline__ float tanh_mufu(float x) {
    // tanh x = 1 - 2 / (e^{2x} + 1); ex2.approx and rcp.approx are accurate to ~1 ulp, so the
    // absolute error is a few 1e-7 everywhere (the cancellation near 0 costs relative, not absolute,
    // accuracy).  e = inf gives 1, e = 0 gives -1, NaN propagates.
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(x, 2.885390081777927f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(e, 1.0f)));
    return fmaf(-2.0f, r, 1.0f);
}

// acc[e][j2] (outputs 2*j2, 2*j2+1 of env e) += sum_k x_e[k] * W[k][.]; W is [K][64] in shared memory
// (every lane reads the same 16 bytes: a broadcast), x_e[k] sits at x[k * ks + e * es] (thread-private).
template <int E>
__device__ __forceinline__ void dense64(const float *__restrict__ W, const float *__restrict__ x, int K, int ks, int es,
                                        float2 (&acc)[E][32]) {
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        float2 hh[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float h = x[k * ks + e * es];
            hh[e] = make_float2(h, h);
        }
        const float4 *w4 = reinterpret_cast<const float4 *>(W + k * kHid);
#pragma unroll
        for (int jb = 0; jb < 16; ++jb) {
            const float4 w = w4[jb];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                acc[e][2 * jb] = __ffma2_rn(make_float2(w.x, w.y), hh[e], acc[e][2 * jb]);
                acc[e][2 * jb + 1] = __ffma2_rn(make_float2(w.z, w.w), hh[e], acc[e][2 * jb + 1]);
            }
        }
    }
}

template <int E>
__device__ __forceinline__ void load_bias64(const float *__restrict__ b, float2 (&acc)[E][32]) {
#pragma unroll
    for (int jb = 0; jb < 16; ++jb) {
        const float4 v = reinterpret_cast<const float4 *>(b)[jb];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            acc[e][2 * jb] = make_float2(v.x, v.y);
            acc[e][2 * jb + 1] = make_float2(v.z, v.w);
        }
    }
}

// One 9 -> 64 -> 64 -> 8 tanh network (the value network's output layer is zero-padded from 1 to 8
// columns so both networks run the same code).  `row` is the envs' observation rows in the obs stage
// (stride 9), `hs` the thread-private column of the hidden-activation scratch [64][32*E].
template <int E>
__device__ __forceinline__ void mlp_forward(const float *__restrict__ net, const float *__restrict__ row, float *__restrict__ hs,
                                            float2 (&out)[E][4]) {
    float2 acc[E][32];
    load_bias64<E>(net + ROBOY_POLICY_OFF_B1, acc);
    dense64<E>(net + ROBOY_POLICY_OFF_W1, row, kObsDim, 1, 32 * kObsDim, acc);
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            hs[(2 * j2) * (32 * E) + 32 * e] = tanh_mufu(acc[e][j2].x);
            hs[(2 * j2 + 1) * (32 * E) + 32 * e] = tanh_mufu(acc[e][j2].y);
        }
    }
    load_bias64<E>(net + ROBOY_POLICY_OFF_B2, acc);
    dense64<E>(net + ROBOY_POLICY_OFF_W2, hs, kHid, 32 * E, 32, acc);
    // output layer fused into the activation of the second: out[e][0..7] += tanh(acc[e][j]) * W3[j][0..7]
    {
        const float4 b0 = reinterpret_cast<const float4 *>(net + ROBOY_POLICY_OFF_B3)[0];
        const float4 b1 = reinterpret_cast<const float4 *>(net + ROBOY_POLICY_OFF_B3)[1];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            out[e][0] = make_float2(b0.x, b0.y);
            out[e][1] = make_float2(b0.z, b0.w);
            out[e][2] = make_float2(b1.x, b1.y);
            out[e][3] = make_float2(b1.z, b1.w);
        }
    }
    const float4 *w3 = reinterpret_cast<const float4 *>(net + ROBOY_POLICY_OFF_W3);
#pragma unroll
    for (int j = 0; j < kHid; ++j) {
        const float4 wa = w3[2 * j], wb = w3[2 * j + 1];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float t = tanh_mufu((j & 1) ? acc[e][j >> 1].y : acc[e][j >> 1].x);
            const float2 tt = make_float2(t, t);
            out[e][0] = __ffma2_rn(make_float2(wa.x, wa.y), tt, out[e][0]);
            out[e][1] = __ffma2_rn(make_float2(wa.z, wa.w), tt, out[e][1]);
            out[e][2] = __ffma2_rn(make_float2(wb.x, wb.y), tt, out[e][2]);
            out[e][3] = __ffma2_rn(make_float2(wb.z, wb.w), tt, out[e][3]);
        }
    }
}

// Two standard normals from two 32-bit words (Box-Muller on the MUFU unit).
__device__ __forceinline__ float2 box_muller(uint32_t x, uint32_t y) {
    const float u1 = fmaf(__uint2float_rn(x >> 8), 0x1p-24f, 0x1p-25f);  // (0, 1)
    const float u2 = __fmul_rn(__uint2float_rn(y >> 8), 0x1p-24f);        // [0, 1)
    const float r = __fsqrt_rn(__fmul_rn(-2.0f, __logf(u1)));
    float s, c;
    __sincosf(__fmul_rn(6.283185307179586f, u2), &s, &c);
    return make_float2(__fmul_rn(r, c), __fmul_rn(r, s));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

struct EnvRegs {
    float g0, g1, g2, ng0, ng1, ng2;
    uint32_t sf;
};

}  // namespace

// The env flags are run-time here (the kernel is FMA-bound in the networks; 2 instantiations instead of 32).
template <bool FASTDIV>
__device__ __forceinline__ void reward_reached_rt(bool penalty, bool bonus, float q0, float q1, float q2, float qd0,
                                                  float qd1, float qd2, const EnvRegs &s, const RobotConsts &c,
                                                  const FastConsts &f, float &reward, bool &reached, bool &violation) {
    if (penalty) {
        if (bonus) reward_reached_sampled_ng<true, true, FASTDIV>(q0, q1, q2, qd0, qd1, qd2, s.g0, s.g1, s.g2, s.ng0, s.ng1, s.ng2, c, f, reward, reached, violation);
        else reward_reached_sampled_ng<true, false, FASTDIV>(q0, q1, q2, qd0, qd1, qd2, s.g0, s.g1, s.g2, s.ng0, s.ng1, s.ng2, c, f, reward, reached, violation);
    } else {
        if (bonus) reward_reached_sampled_ng<false, true, FASTDIV>(q0, q1, q2, qd0, qd1, qd2, s.g0, s.g1, s.g2, s.ng0, s.ng1, s.ng2, c, f, reward, reached, violation);
        else reward_reached_sampled_ng<false, false, FASTDIV>(q0, q1, q2, qd0, qd1, qd2, s.g0, s.g1, s.g2, s.ng0, s.ng1, s.ng2, c, f, reward, reached, violation);
    }
}

template <bool FASTDIV>
__device__ __forceinline__ void normalize_goal(EnvRegs &s, const RobotConsts &c, const FastConsts &f) {
    s.ng0 = normalize32_hot<FASTDIV>(s.g0, c.a_hi, c.a_lo, c.a_span, f.a_rc);
    s.ng1 = normalize32_hot<FASTDIV>(s.g1, c.a_hi, c.a_lo, c.a_span, f.a_rc);
    s.ng2 = normalize32_hot<FASTDIV>(s.g2, c.a_hi, c.a_lo, c.a_span, f.a_rc);
}

template <int E>
__global__ void __launch_bounds__(kPolicyMaxBlock, 1) policy_rollout_kernel(const __grid_constant__ StepParams p,
                                                                             const __grid_constant__ PolicyParams q) {
    const bool PENALTY = q.penalty, BONUS = q.bonus, AUTO_RESET = q.auto_reset, FASTDIV = q.fastdiv;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    // shared memory: policy image | per warp: hidden scratch [64][32E], obs stage [32E][9] | counters
    float *img = smem;
    float *hs = smem + ROBOY_POLICY_IMAGE_FLOATS + warp * (kHid * 32 * E) + lane;
    float *stage = smem + ROBOY_POLICY_IMAGE_FLOATS + n_warps * (kHid * 32 * E) + warp * (32 * E * kObsDim);
    double *s_red = reinterpret_cast<double *>(smem + ROBOY_POLICY_IMAGE_FLOATS + n_warps * (kHid * 32 * E + 32 * E * kObsDim));
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(s_red + n_warps);
    for (int i = threadIdx.x; i < ROBOY_POLICY_IMAGE_FLOATS / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(img)[i] = reinterpret_cast<const float4 *>(q.image)[i];
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0;
    __syncthreads();

    const uint64_t t_first = counter_begin(p.cc);
    const uint32_t n_end = (uint32_t)p.e_end;
    const size_t n = (size_t)p.n;
    const uint32_t n_chunks = (n_end + 32 * E - 1) / (32 * E);
    const uint32_t warp_stride = gridDim.x * n_warps;
    const float *vf_net = img + ROBOY_POLICY_OFF_VF, *pi_net = img + ROBOY_POLICY_OFF_PI;
    float sum_reward = 0.0f;

    for (uint32_t chunk = blockIdx.x * n_warps + warp; chunk < n_chunks; chunk += warp_stride) {
        const uint32_t base = chunk * (32 * E);
        const bool full = base + 32 * E <= n_end;
        EnvRegs s[E];
        bool live[E];
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const uint32_t env = base + 32 * e + lane;
            live[e] = env < n_end;
            s[e].g0 = live[e] ? p.goal[env] : 0.f;
            s[e].g1 = live[e] ? p.goal1[env] : 0.f;
            s[e].g2 = live[e] ? p.goal2[env] : 0.f;
            s[e].sf = live[e] ? p.step_flags[env] : 1u;
            if (FASTDIV) normalize_goal<true>(s[e], p.c, p.f);
            else normalize_goal<false>(s[e], p.c, p.f);
            // the observation the rollout starts from: slot 0 of the obs buffer
            float *row = stage + (32 * e + lane) * kObsDim;
#pragma unroll
            for (int k = 0; k < kObsDim; ++k) row[k] = live[e] ? q.obs[(size_t)env * kObsDim + k] : 0.f;
        }

        for (uint32_t tt = 0;; ++tt) {
            // ---- value network on obs[tt] (also the bootstrap value at tt == T) ----
            float2 out[E][4];
            mlp_forward<E>(vf_net, stage + lane * kObsDim, hs, out);
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (live[e]) q.values[(size_t)tt * n + base + 32 * e + lane] = out[e][0].x;
            if (tt == q.T) break;
            // ---- policy network: mean of the Gaussian ----
            mlp_forward<E>(pi_net, stage + lane * kObsDim, hs, out);
            const uint64_t t = t_first + tt;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const uint32_t env = base + 32 * e + lane;
                const uint64_t gid = p.gid_base + env;
                // ---- a = mean + std * N(0, 1); the runner clips it to the action space for env.step ----
                const uint4 r0 = philox_draw(gid, t, kStreamNoise, q.noise_keys, 0);
                const uint4 r1 = philox_draw(gid, t, kStreamNoise, q.noise_keys, 1);
                const float2 z01 = box_muller(r0.x, r0.y), z23 = box_muller(r0.z, r0.w);
                const float2 z45 = box_muller(r1.x, r1.y), z67 = box_muller(r1.z, r1.w);
                const float z[8] = {z01.x, z01.y, z23.x, z23.y, z45.x, z45.y, z67.x, z67.y};
                const float mean[8] = {out[e][0].x, out[e][0].y, out[e][1].x, out[e][1].y,
                                       out[e][2].x, out[e][2].y, out[e][3].x, out[e][3].y};
                const float *sd = img + ROBOY_POLICY_OFF_STD;
                float u[8], a[8], zz = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    u[k] = fmaf(sd[k], z[k], mean[k]);
                    const float c = fminf(fmaxf(u[k], p.act_in_lo), p.act_in_hi);
                    a[k] = (u[k] != u[k]) ? u[k] : c;  // a NaN stays a NaN (and trips the :52 assert below)
                    zz = fmaf(z[k], z[k], zz);
                }
                const float logp = fmaf(-0.5f, zz, img[ROBOY_POLICY_OFF_LOGNORM]);
                if (live[e]) {
                    float4 *ap = reinterpret_cast<float4 *>(q.actions + ((size_t)tt * n + env) * kActDim);
                    ap[0] = make_float4(u[0], u[1], u[2], u[3]);
                    ap[1] = make_float4(u[4], u[5], u[6], u[7]);
                    q.logp[(size_t)tt * n + env] = logp;
                    if (q.noise) {
                        float4 *np = reinterpret_cast<float4 *>(q.noise + ((size_t)tt * n + env) * kActDim);
                        np[0] = make_float4(z[0], z[1], z[2], z[3]);
                        np[1] = make_float4(z[4], z[5], z[6], z[7]);
                    }
                }

                // ---- RoboyEnv.step on that action: same arithmetic as step_kernel ----
                bool act_ok = true, hold = live[e];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    act_ok = act_ok && fabsf(a[k]) <= p.act_in_hi;        // roboy_env.py:52
                    hold = hold && a[k] >= p.hold_lo && a[k] <= p.hold_hi;  // simulation_client.py:38
                }
                const float g0 = s[e].g0, g1 = s[e].g1, g2 = s[e].g2;
                const uint32_t sf = s[e].sf;
                float q0, q1, q2, qd0, qd1, qd2, reward;
                bool reached, violation;
                if (!hold) {
                    const Draw6 d = split6x21(philox_draw(gid, t, kStreamState, p.keys));
                    q0 = uniform_in21(d.k[0], p.c.a_lo, p.f.a_span21);
                    q1 = uniform_in21(d.k[1], p.c.a_lo, p.f.a_span21);
                    q2 = uniform_in21(d.k[2], p.c.a_lo, p.f.a_span21);
                    qd0 = uniform_in21(d.k[3], p.c.a_lo, p.f.a_span21);
                    qd1 = uniform_in21(d.k[4], p.c.a_lo, p.f.a_span21);
                    qd2 = uniform_in21(d.k[5], p.c.a_lo, p.f.a_span21);
                    if (FASTDIV) reward_reached_rt<true>(PENALTY, BONUS, q0, q1, q2, qd0, qd1, qd2, s[e], p.c, p.f, reward, reached, violation);
                    else reward_reached_rt<false>(PENALTY, BONUS, q0, q1, q2, qd0, qd1, qd2, s[e], p.c, p.f, reward, reached, violation);
                } else {
                    const HoldOut h = hold_branch(p, env, sf, g0, g1, g2, PENALTY, BONUS);
                    q0 = h.q0; q1 = h.q1; q2 = h.q2; qd0 = h.qd0; qd1 = h.qd1; qd2 = h.qd2;
                    reward = h.reward;
                    reached = h.flags & 1u;
                    violation = h.flags & 2u;
                    atomicAdd(&s_cnt[2], 1u);
                }
                uint32_t step = sf & ROBOY_STEP_MASK;
                step += step < ROBOY_STEP_MASK;                          // roboy_env.py:60
                const bool done = reached || (int32_t)step > p.max_len;  // :65-66, :72-73
                uint32_t word = step | (sf & ~ROBOY_STEP_MASK);
                float *row = stage + (32 * e + lane) * kObsDim;          // obs = [q, qd, goal], :75-80
                row[0] = q0; row[1] = q1; row[2] = q2; row[3] = qd0; row[4] = qd1; row[5] = qd2;
                row[6] = g0; row[7] = g1; row[8] = g2;
                if (done && live[e]) {
                    const uint32_t r = finish_episode(p, t, env, step, reached, AUTO_RESET, row, s_cnt);
                    word = (r & 0x80000000u) ? word : (r | ROBOY_F_HELD_ZERO64);
                    s[e].g0 = p.goal[env];   // the new goal was stored by finish_episode (same thread)
                    s[e].g1 = p.goal1[env];
                    s[e].g2 = p.goal2[env];
                    if (FASTDIV) normalize_goal<true>(s[e], p.c, p.f);
                    else normalize_goal<false>(s[e], p.c, p.f);
                }
                if (live[e] && (violation || !act_ok)) {
                    atomicOr(p.err_flags, (violation ? ROBOY_ERR_REWARD_RANGE : 0u) | (!act_ok ? ROBOY_ERR_ACTION : 0u));
                    atomicMin(p.first_bad, (unsigned long long)gid);
                    atomicAdd(&s_cnt[3], 1u);
                }
                s[e].sf = word;
                if (live[e]) {
                    p.reward[(size_t)tt * n + env] = reward;
                    p.done[(size_t)tt * n + env] = (uint8_t)done;
                    sum_reward += reward;
                }
            }
            // ---- obs[tt + 1]: the stage now holds the chunk's 32E rows; store them coalesced ----
            __syncwarp();
            float *dst = q.obs + ((size_t)(tt + 1) * n + base) * kObsDim;
            if (full && q.obs_aligned) {
                const float4 *src = reinterpret_cast<const float4 *>(stage);
#pragma unroll
                for (int i = 0; i < (32 * E * kObsDim) / 4 / 32; ++i)
                    reinterpret_cast<float4 *>(dst)[i * 32 + lane] = src[i * 32 + lane];
                if (lane < ((32 * E * kObsDim) / 4) % 32)
                    reinterpret_cast<float4 *>(dst)[((32 * E * kObsDim) / 4 / 32) * 32 + lane] =
                        src[((32 * E * kObsDim) / 4 / 32) * 32 + lane];
            } else {
                const uint32_t rows = full ? 32u * E : n_end - base;
                for (uint32_t i = lane; i < rows * kObsDim; i += 32) dst[i] = stage[i];
            }
            __syncwarp();
        }
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (live[e]) p.step_flags[base + 32 * e + lane] = s[e].sf;
    }

    // ---- episode statistics, one set of atomics per CTA (as in step_kernel) ----
    const double w_reward = warp_sum_d((double)sum_reward);
    if (lane == 0) s_red[warp] = w_reward;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int w = 0; w < n_warps; ++w) r += s_red[w];
        const double done = (double)s_cnt[0], succ = (double)s_cnt[1];
        const double steps = blockIdx.x == 0 ? (double)(p.e_end - p.e_begin) * (double)q.T : 0.0;
        const double v[ROBOY_STAT_COUNT] = {steps, done, succ, done - succ, r, (double)s_cnt[4],
                                            (double)s_cnt[2], (double)s_cnt[3]};
#pragma unroll
        for (int k = 0; k < ROBOY_STAT_COUNT; ++k)
            if (v[k] != 0.0) atomicAdd(p.stats + k, v[k]);
        counter_end(p.cc, t_first);
    }
}

PolicyGeom policy_geometry(uint64_t n_envs, int sm_count, int envs_per_thread) {
    PolicyGeom g;
    const uint64_t resident = (uint64_t)sm_count * (kPolicyMaxBlock / 32) * 32;  // envs with one env per thread
    g.envs_per_thread = envs_per_thread ? envs_per_thread : (n_envs <= resident ? 1 : 2);
    const uint64_t n_chunks = (n_envs + 32 * g.envs_per_thread - 1) / (32 * g.envs_per_thread);
    // spread the warps over the SMs first, then fill the CTAs (up to kPolicyMaxBlock threads, 1 CTA per SM)
    g.grid = (int)(n_chunks < (uint64_t)sm_count ? n_chunks : (uint64_t)sm_count);
    uint64_t wpb = (n_chunks + g.grid - 1) / g.grid;
    if (wpb > kPolicyMaxBlock / 32) wpb = kPolicyMaxBlock / 32;
    g.block = (int)wpb * 32;
    g.smem = (int)(sizeof(float) * (ROBOY_POLICY_IMAGE_FLOATS + wpb * (kHid * 32 * g.envs_per_thread + 32 * g.envs_per_thread * kObsDim)) +
                   sizeof(double) * wpb + sizeof(unsigned int) * 8);
    return g;
}

cudaError_t launch_policy_rollout(const StepParams &p, const PolicyParams &q, bool penalty, bool bonus, bool auto_reset,
                                  bool fastdiv, int sm_count, int envs_per_thread, cudaStream_t stream) {
    if (p.e_end <= p.e_begin) return cudaSuccess;
    PolicyParams qq = q;
    qq.penalty = penalty;
    qq.bonus = bonus;
    qq.auto_reset = auto_reset;
    qq.fastdiv = fastdiv;
    const PolicyGeom g = policy_geometry(p.e_end, sm_count, envs_per_thread);
    void (*fn)(const StepParams, const PolicyParams) =
        g.envs_per_thread == 1 ? policy_rollout_kernel<1> : policy_rollout_kernel<2>;
    cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem);
    if (err != cudaSuccess) return err;
    fn<<<g.grid, g.block, g.smem, stream>>>(p, qq);
    return cudaGetLastError();
}

}  // namespace roboy
