// Parameter block and launcher of the fused policy rollout kernel (roboy_policy.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/roboy_b200.h"
#include "roboy_kernels.cuh"

namespace roboy {

constexpr int kPolicyMaxBlock = 256;
constexpr int kPolicyTcMaxBlock = 512;      // tensor-core variant: up to four 128-env tiles per CTA
constexpr int kPolicyTcImagePad = 15040;     // ROBOY_TC_IMAGE_BYTES rounded up to 128 bytes, in floats  // threads per CTA at most; one CTA per SM (shared memory bound)

struct PolicyParams {
    const float *__restrict__ image;  // [ROBOY_POLICY_IMAGE_FLOATS] packed policy (include/roboy_b200.h)
    uint32_t T;
    int32_t obs_aligned;              // every obs[t] slice is 16-byte aligned
    float *__restrict__ obs;          // [T+1][n][9]; slot 0 is the input, slots 1..T are written
    float *__restrict__ actions;      // [T][n][8]  un-clipped samples (what PPO2's runner stores)
    float *__restrict__ logp;         // [T][n]
    float *__restrict__ values;       // [T+1][n]
    float *__restrict__ noise;        // [T][n][8] or nullptr
    PhiloxKeys noise_keys;
    bool penalty, bonus, auto_reset, fastdiv;  // env flags, run-time in this kernel (filled by the launcher)
};

struct PolicyGeom {
    int grid, block, smem, envs_per_thread;
};

// envs_per_thread: 0 = choose (1 while every env fits on the chip at once, else 2)
PolicyGeom policy_geometry(uint64_t n_envs, int sm_count, int envs_per_thread);
// p: as for launch_step_many (reward / done are [T][n]; p.actions and p.obs are not used)
cudaError_t launch_policy_rollout(const StepParams &p, const PolicyParams &q, bool penalty, bool bonus, bool auto_reset,
                                  bool fastdiv, int sm_count, int envs_per_thread, cudaStream_t stream);

// tensor-core (tcgen05, TF32) variant: q.image is the ROBOY_TC_* image
// variant: 0 = choose; 1 = one 128-env tile per group of 128 threads; 2 = two tiles per group, ping-pong; 3 = merged
// (both networks of a tile advance together: small populations).  PolicyGeom::envs_per_thread returns the choice.
// exact: split-float16 operands (float32-level accuracy) instead of single float16 operands
PolicyGeom policy_tc_geometry(uint64_t n_envs, int sm_count, int tiles_per_group, bool exact);
cudaError_t launch_policy_rollout_tc(const StepParams &p, const PolicyParams &q, bool penalty, bool bonus, bool auto_reset,
                                     bool fastdiv, int sm_count, int tiles_per_group, bool exact, cudaStream_t stream);

}  // namespace roboy
