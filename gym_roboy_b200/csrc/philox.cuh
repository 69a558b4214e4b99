// Philox4x32-10 counter-based RNG (Salmon, Moraes, Dror, Shaw -- SC'11) for sm_100a.
//
// Replaces gym's Box.sample() behind RoboyRobot.new_random_state (roboy_robot.py:35-39):
// that stream is third-party and unpinned, so draws are defined here instead and parity
// with the reference is shown by injecting these draws into it (oracle/reference_harness.py).
//
// counter = (global env id lo, global env id hi, call counter lo,
//            stream<<28 | sub<<20 | call counter hi (20 bits))
//   sub numbers repeated goal draws inside one call counter value (SimulationClient.
//   get_new_goal_joint_angles called several times between steps); the fused step uses sub 0.
// key     = (seed lo, seed hi)
// Nothing is stored per env: a draw is a pure function of (seed, env id, call, stream), so the
// result is independent of how envs are sharded over GPUs.
#pragma once
#include <stdint.h>

namespace roboy {

enum : uint32_t { kStreamState = 0, kStreamGoal = 2, kStreamNoise = 4 };  // kStreamNoise: the policy's Gaussian noise (own key)

struct PhiloxKeys {
    // the ten round keys, k + r*W, precomputed on the host (uniform across the grid)
    uint32_t k0[10];
    uint32_t k1[10];
};

__host__ inline PhiloxKeys make_philox_keys(uint64_t seed) {
    PhiloxKeys ks;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        ks.k0[r] = a;
        ks.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return ks;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKeys &ks) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // one IMAD.WIDE.U32 per product gives both halves
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ ks.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ ks.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 philox_draw(uint64_t gid, uint64_t t, uint32_t stream, const PhiloxKeys &ks,
                                             uint32_t sub = 0) {
    return philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t,
                         (stream << 28) | ((sub & 0xffu) << 20) | ((uint32_t)(t >> 32) & 0x000fffffu), ks);
}

// u = (x >> 8) * 2^-24 in [0,1);  v = low + span*u as an unfused float32 multiply and add, so a
// numpy float32 restatement reproduces it bit for bit.  span = high - low (float32).
__device__ __forceinline__ float uniform_in(uint32_t x, float low, float span) {
    const float u = __fmul_rn(__uint2float_rn(x >> 8), 0x1p-24f);
    return __fadd_rn(low, __fmul_rn(span, u));
}

// Same value as uniform_in: (x>>8) * 2^-24 is exact and so is span * 2^-24, hence
// RN(span * RN((x>>8) * 2^-24)) == RN(span24 * float(x>>8)) with span24 = span * 2^-24.
__device__ __forceinline__ float uniform_in24(uint32_t x, float low, float span24) {
    return __fadd_rn(low, __fmul_rn(span24, __uint2float_rn(x >> 8)));
}

// A state draw (RoboyRobot.new_random_state: 3 angles + 3 velocities, roboy_robot.py:35-39) takes
// ONE Philox block: its 128 bits are cut into six 21-bit integers k0..k5 (bits 127..2, the last
// two unused), v_i = low + (span * 2^-21) * float(k_i).  21 bits put the samples on a grid of
// 3e-6 rad; a second block per draw would cost a quarter of the step kernel's instructions.
struct Draw6 {
    uint32_t k[6];
};

__device__ __forceinline__ Draw6 split6x21(const uint4 r) {
    Draw6 d;
    d.k[0] = r.x >> 11;
    d.k[1] = __funnelshift_r(r.y, r.x, 22) & 0x1fffffu;  // low 11 bits of x, high 10 bits of y
    d.k[2] = (r.y >> 1) & 0x1fffffu;
    d.k[3] = r.z >> 11;
    d.k[4] = __funnelshift_r(r.w, r.z, 22) & 0x1fffffu;
    d.k[5] = (r.w >> 1) & 0x1fffffu;
    return d;
}

__device__ __forceinline__ float uniform_in21(uint32_t k, float low, float span21) {
    return __fadd_rn(low, __fmul_rn(span21, __uint2float_rn(k)));
}

}  // namespace roboy
