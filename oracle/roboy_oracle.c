/*
 * oracle/roboy_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of gym-roboy's environment hot path for any RoboyRobot plug-in (MsjRobot and robots with
 * other joint / tendon counts and per-component bounds, roboy_robot.py:21-33), used
 * ONLY as the checker for the CUDA path (tests/, __graft_entry__.smoke(), and bench.py's
 * cpu_baseline / --impl reference legs).  Nothing under gym_roboy_b200/ may import, link or
 * call it.  Every function cites the reference lines it follows (paths relative to
 * /root/reference/gym_roboy/).
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_vs_reference.py,
 * oracle/gen_golden.py) against the unmodified reference run in the build container through
 * a test-only gym shim and a replay robot, and against golden vectors generated that way
 * (tests/golden/); its stand-alone compute_reward / _did_reach_goal additionally on SALTED inputs -- NaN, +-inf,
 * 3e38, denormals, in the state, the goal and the goal velocities, 3 robots x 4 flag combinations, reward / reached /
 * the :109 assert all compared (tests/test_oracle_salted_vs_reference.py).  Third-party arithmetic the reference leans on was pinned empirically in
 * that container (numpy 2.3.5 + scipy-openblas 0.3.30, x86-64):
 *   - np.linalg.norm(float32[3], ord=2) == sqrtf((float)(sum_k (double)(float)(x_k*x_k)))
 *     (OpenBLAS sdot: float products accumulated sequentially in a double, cast to float);
 *   - np.linalg.norm(float64[n]), n < 16 == sqrt(fma(x_{n-1}, x_{n-1}, ... fma(x1, x1, x0*x0)))  (OpenBLAS ddot's
 *     scalar tail loop, compiled with FMA contraction: pinned on 1,500 random vectors per length, 0 mismatches for
 *     n = 3..15, while the unfused sum mismatches ~12 % of them; with float32-valued inputs -- all the MSJ stub path
 *     produces -- the products are exact and the two agree).  From n = 16 (float64) / n = 32 (float32) on OpenBLAS
 *     switches to SIMD-blocked partial sums whose order depends on the host CPU's kernel: that is why the CUDA path
 *     and this oracle cap a robot at ORC_MAX_JOINT = 15 joints;
 *   - np.exp(float32) is NOT correctly rounded (about 40 % of inputs differ from libm expf by
 *     one ulp), so rewards are compared with a relative tolerance (1e-6), never bit-exactly.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off; never -ffast-math).
 *
 * The random draws (goal / stub state) are NOT the reference's: gym's Box.sample stream is
 * unpinned.  Draws come from Philox4x32-10 keyed by (seed) and countered by
 * (global env id, call counter, stream); parity against the reference is established by
 * replaying these same draws INTO the reference (oracle/reference_harness.py).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Configuration and state (mirrors include/roboy_b200.h field by field, written independently)
 * ---------------------------------------------------------------------------------------- */
#define ORC_MAX_JOINT 15   /* see the header: numpy's norm changes its summation order from 16 float64 elements on */
#define ORC_JOINT_PAD 16
#define ORC_MAX_ACTION 64

typedef struct {
    uint64_t n_envs;
    uint64_t env_id_base;      /* global id of local env 0 (multi-GPU sharding) */
    uint64_t seed;
    float angle_low, angle_high;   /* msj_robot.py:9   +-pi  as float32 (scalar bounds: every component) */
    float vel_low, vel_high;       /* msj_robot.py:10  +-pi/6 as float32 */
    float act_low, act_high;       /* msj_robot.py:16  +-0.3 as float32 */
    int32_t max_episode_len;       /* roboy_env.py:28  400 */
    int32_t joint_vel_penalty;     /* roboy_env.py:13 */
    int32_t bonus_for_goal;        /* roboy_env.py:14 */
    int32_t auto_reset;            /* vec-env worker semantics */
    float penalty_boundary;        /* roboy_env.py:26  1 */
    float bonus_goal;              /* roboy_env.py:27  1000 */
    double reward_lo, reward_hi;   /* roboy_env.py:30,109 (use -inf/+inf to disable) */
    /* other robots (roboy_robot.py:21-33): dims (0 = MSJ's 3 / 8) and, if per_component_bounds, one bound per component */
    int32_t dim_joint, dim_action;
    int32_t per_component_bounds;
    int32_t reserved;
    float angle_low_v[ORC_JOINT_PAD], angle_high_v[ORC_JOINT_PAD];
    float vel_low_v[ORC_JOINT_PAD], vel_high_v[ORC_JOINT_PAD];
    float act_low_v[ORC_MAX_ACTION], act_high_v[ORC_MAX_ACTION];
} orc_cfg;

static int cfg_J(const orc_cfg *c) { return c->dim_joint > 0 ? c->dim_joint : 3; }
static int cfg_A(const orc_cfg *c) { return c->dim_action > 0 ? c->dim_action : 8; }
static float a_lo(const orc_cfg *c, int k) { return c->per_component_bounds ? c->angle_low_v[k] : c->angle_low; }
static float a_hi(const orc_cfg *c, int k) { return c->per_component_bounds ? c->angle_high_v[k] : c->angle_high; }
static float v_lo(const orc_cfg *c, int k) { return c->per_component_bounds ? c->vel_low_v[k] : c->vel_low; }
static float v_hi(const orc_cfg *c, int k) { return c->per_component_bounds ? c->vel_high_v[k] : c->vel_high; }
static float t_lo(const orc_cfg *c, int k) { return c->per_component_bounds ? c->act_low_v[k] : c->act_low; }
static float t_hi(const orc_cfg *c, int k) { return c->per_component_bounds ? c->act_high_v[k] : c->act_high; }

/* step_flags word: low 24 bits step_num, then flag bits */
#define ORC_STEP_MASK 0x00ffffffu
#define ORC_F_HELD_ZERO64 (1u << 24) /* held state is the reference's float64 zero state */
#define ORC_F_HELD_INFEASIBLE (1u << 25)

/* error bits (reference AssertionErrors restated as flags) */
#define ORC_ERR_ACTION 1u       /* roboy_env.py:52 */
#define ORC_ERR_REWARD_RANGE 2u /* roboy_env.py:109 */
#define ORC_ERR_GOAL_BOUNDS 4u  /* roboy_robot.py:76 */
#define ORC_ERR_STATE_BOUNDS 8u /* roboy_robot.py:76 through ros_simulation_client.py:40-46: an external state's angles */

enum { ST_STEPS = 0, ST_EPISODES, ST_SUCCESSES, ST_TIMEOUTS, ST_SUM_REWARD, ST_SUM_EPLEN, ST_HOLDS, ST_VIOLATIONS, ST_N };

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11, Random123) -- published algorithm restated.
 * ---------------------------------------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { STREAM_STATE = 0, STREAM_GOAL = 2 };

/* counter = (env id lo, env id hi | block << 24, call counter lo, stream<<28 | sub<<20 | call counter hi);
 * sub numbers repeated goal draws at one call counter (un-fused plug-in API); 0 in the step.  block numbers the
 * Philox blocks of one draw for robots with more than 3 joints (0 for MSJ; global env ids stay below 2^56). */
static void draw4b(const orc_cfg *cfg, uint64_t gid, uint64_t t, uint32_t stream, uint32_t sub, uint32_t block, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32) | (block << 24), (uint32_t)t,
                       (stream << 28) | ((sub & 0xffu) << 20) | ((uint32_t)(t >> 32) & 0x000fffffu)};
    uint32_t key[2] = {(uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}

/* u = (x>>8)*2^-24 in [0,1); v = low + (high-low)*u, float32 mul then add (no FMA).
 * Stands in for gym Box.sample() (roboy_robot.py:37-38), whose own stream is unpinned. */
static float uniform_in(uint32_t x, float low, float high) {
    float u = (float)(x >> 8) * 0x1p-24f;
    float span = high - low;
    float m = span * u;
    return low + m;
}

/* roboy_robot.py:35-39 new_random_state(): q ~ U[angle space]^J, and -- reference quirk --
 * the velocities are ALSO drawn from the ANGLE space (:38). is_feasible = True.
 * The 2J values (q_0..q_{J-1}, qd_0..qd_{J-1}) are numbered c = 0..2J-1; value c comes from Philox block c / 6, whose
 * 128 bits are cut into six 21-bit integers k0..k5 (bits 127..2): v = low + (span * 2^-21) * (float)k  (float32
 * multiply then add), with low / span of the ANGLE component the value belongs to.  MSJ (J = 3): one block. */
void orc_draw_state(const orc_cfg *cfg, uint64_t gid, uint64_t t, float *q, float *qd) {
    const int J = cfg_J(cfg);
    uint32_t r[4] = {0, 0, 0, 0}, k[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 2 * J; ++c) {
        if (c % 6 == 0) {
            draw4b(cfg, gid, t, STREAM_STATE, 0, (uint32_t)(c / 6), r);
            k[0] = r[0] >> 11;
            k[1] = ((r[0] & 0x7ffu) << 10) | (r[1] >> 22);
            k[2] = (r[1] >> 1) & 0x1fffffu;
            k[3] = r[2] >> 11;
            k[4] = ((r[2] & 0x7ffu) << 10) | (r[3] >> 22);
            k[5] = (r[3] >> 1) & 0x1fffffu;
        }
        const int j = c < J ? c : c - J;
        const float span21 = (a_hi(cfg, j) - a_lo(cfg, j)) * 0x1p-21f;
        float m = span21 * (float)k[c % 6];
        float v = a_lo(cfg, j) + m;
        if (c < J) q[j] = v; else qd[j] = v;
    }
}

/* simulation_client.py:46-47 get_new_goal_joint_angles(): new_random_state().joint_angles.  Goal component k comes from
 * word k % 3 of Philox block k / 3 on the goal stream (24-bit draws). */
void orc_draw_goal_sub(const orc_cfg *cfg, uint64_t gid, uint64_t t, uint32_t sub, float *g) {
    const int J = cfg_J(cfg);
    uint32_t a[4] = {0, 0, 0, 0};
    for (int k = 0; k < J; ++k) {
        if (k % 3 == 0) draw4b(cfg, gid, t, STREAM_GOAL, sub, (uint32_t)(k / 3), a);
        g[k] = uniform_in(a[k % 3], a_lo(cfg, k), a_hi(cfg, k));
    }
}
void orc_draw_goal(const orc_cfg *cfg, uint64_t gid, uint64_t t, float *g) { orc_draw_goal_sub(cfg, gid, t, 0, g); }

/* ------------------------------------------------------------------------------------------
 * numpy arithmetic restated (see header for how each was pinned)
 * ---------------------------------------------------------------------------------------- */
#define JP ORC_JOINT_PAD

/* roboy_env.py:137-140 _l2_distance on float32 operands */
static float l2_f32(int J, const float *a, const float *b) {
    double s = 0.0;
    for (int k = 0; k < J; ++k) {
        float d = a[k] - b[k];      /* :138 np.subtract in float32 */
        if (isnan(d)) d = 0.0f;     /* :139 */
        float p = d * d;            /* OpenBLAS sdot: float product ... */
        s += (double)p;             /* ... accumulated in double */
    }
    return sqrtf((float)s);         /* :140 np.linalg.norm -> sqrt(float32) */
}

/* roboy_env.py:137-140 _l2_distance once either operand is float64 (OpenBLAS ddot tail: sequential FMA) */
static double l2_f64(int J, const double *a, const double *b) {
    double s = 0.0;
    for (int k = 0; k < J; ++k) {
        double d = a[k] - b[k];
        if (isnan(d)) d = 0.0;
        s = fma(d, d, s);
    }
    return sqrt(s);
}

/* roboy_robot.py:93-95  (2*val - max_val - min_val) / (max_val - min_val), in val's dtype */
static float normalize_f32(float v, float hi, float lo) {
    float t = 2.0f * v;
    t = t - hi;
    t = t - lo;
    return t / (hi - lo);
}
static double normalize_f64(double v, float hi, float lo) {
    double t = 2.0 * v;
    t = t - (double)hi;
    t = t - (double)lo;
    return t / (double)(hi - lo); /* max-min is float32-float32 -> float32, then promoted */
}

/* One env's state as the reference holds it: values + the dtype numpy would carry. */
typedef struct {
    double q[JP], qd[JP]; /* float32-valued when is64 == 0 */
    int is64;             /* 1: float64 arrays (zero state roboy_robot.py:41-45, wire values), 0: float32 sample */
    int feasible;
} orc_state;

/* roboy_env.py:125-134 _did_reach_goal.  goal_q is float32; goal velocities are either the
 * float64 zeros of roboy_env.py:23 (goal_qd == NULL) or float32 values (stand-alone call). */
static int did_reach_goal(const orc_cfg *cfg, const orc_state *s, const float *goal_q,
                          const float *goal_qd, float thr_angle, float thr_vel) {
    const int J = cfg_J(cfg);
    int angles_close, vels_close;
    if (!s->is64) {
        float q[JP];
        for (int k = 0; k < J; ++k) q[k] = (float)s->q[k];
        angles_close = l2_f32(J, q, goal_q) < thr_angle;           /* :126-127 float32 */
    } else {
        double g[JP];
        for (int k = 0; k < J; ++k) g[k] = goal_q[k];
        angles_close = l2_f64(J, s->q, g) < (double)thr_angle;     /* float64 state promotes */
    }
    if (!s->is64 && goal_qd) {
        float v[JP];
        for (int k = 0; k < J; ++k) v[k] = (float)s->qd[k];
        vels_close = l2_f32(J, v, goal_qd) < thr_vel;
    } else {
        double gv[JP];
        for (int k = 0; k < J; ++k) gv[k] = goal_qd ? (double)goal_qd[k] : 0.0;
        vels_close = l2_f64(J, s->qd, gv) < (double)thr_vel;       /* :129-130 float64 */
    }
    return angles_close && vels_close;                             /* :134 */
}

/* roboy_env.py:92-112 compute_reward.  Returns the python float; *violation set when the
 * assert at :109 would fire.  `reached` is the value of the _did_reach_goal call at :105. */
static double compute_reward(const orc_cfg *cfg, const orc_state *s, const float *goal_q,
                             const float *goal_qd, int reached, int *violation) {
    const int J = cfg_J(cfg);
    double reward; /* carries float32 values exactly while the reference is in float32 */
    int r64;       /* dtype numpy would carry */
    float r32 = 0.0f;
    float ng[JP];
    for (int k = 0; k < J; ++k) ng[k] = normalize_f32(goal_q[k], a_hi(cfg, k), a_lo(cfg, k));           /* :95 */
    if (!s->is64) {
        float nq[JP];
        for (int k = 0; k < J; ++k) nq[k] = normalize_f32((float)s->q[k], a_hi(cfg, k), a_lo(cfg, k));  /* :94 */
        r32 = -expf(l2_f32(J, nq, ng));                                                                 /* :96 */
        r64 = 0;
        reward = r32;
    } else {
        double nq[JP], g64[JP];
        for (int k = 0; k < J; ++k) { nq[k] = normalize_f64(s->q[k], a_hi(cfg, k), a_lo(cfg, k)); g64[k] = ng[k]; }
        reward = -exp(l2_f64(J, nq, g64));
        r64 = 1;
    }
    if (cfg->joint_vel_penalty) {                                                     /* :98-100 */
        if (!s->is64 && goal_qd) { /* all-float32 stand-alone call (roboy_env.py:40-49) */
            float nv[JP], ngv[JP];
            for (int k = 0; k < J; ++k) {
                nv[k] = normalize_f32((float)s->qd[k], v_hi(cfg, k), v_lo(cfg, k));
                ngv[k] = normalize_f32(goal_qd[k], v_hi(cfg, k), v_lo(cfg, k));
            }
            double acc = 0.0;   /* np.linalg.norm(float32 difference): no NaN guard at :99 */
            for (int k = 0; k < J; ++k) { float d = nv[k] - ngv[k]; float p = d * d; acc += (double)p; }
            float v = sqrtf((float)acc);
            float e = expf(r32);
            float diff = r32 - e;
            r32 = (v + 1.0f) * diff;
            reward = r32;
        } else {
            double acc = 0.0;
            for (int k = 0; k < J; ++k) {
                double nv = s->is64 ? normalize_f64(s->qd[k], v_hi(cfg, k), v_lo(cfg, k))
                                    : (double)normalize_f32((float)s->qd[k], v_hi(cfg, k), v_lo(cfg, k));
                double ngv = goal_qd ? (double)normalize_f32(goal_qd[k], v_hi(cfg, k), v_lo(cfg, k))
                                     : normalize_f64(0.0, v_hi(cfg, k), v_lo(cfg, k));
                double d = nv - ngv;
                acc = fma(d, d, acc);                   /* np.linalg.norm, float64: sequential FMA */
            }
            double v = sqrt(acc);
            double diff = r64 ? (reward - exp(reward)) : (double)(r32 - expf(r32));
            reward = (v + 1.0) * diff;
            r64 = 1;
        }
    }
    if (!s->feasible) {                                                               /* :102-103 */
        /* np.abs(int) is an int64 scalar: float32 - int64 -> float64 (NEP 50) */
        reward = reward - (double)cfg->penalty_boundary;
        r64 = 1;
    }
    if (reached && cfg->bonus_for_goal) {                                             /* :105-107 */
        if (r64) reward = reward + (double)cfg->bonus_goal;
        else { r32 = (float)reward + cfg->bonus_goal; reward = r32; }
    }
    if (!(cfg->reward_lo <= reward && reward <= cfg->reward_hi)) *violation = 1;      /* :109 */
    return reward;                                                                    /* :112 */
}

/* roboy_env.py:24-25 + :127,:130 thresholds:  _l2_distance(low, high)/200 and /5, float32 */
void orc_thresholds(const orc_cfg *cfg, float *thr_angle, float *thr_vel) {
    const int J = cfg_J(cfg);
    float lo[JP], hi[JP];
    for (int k = 0; k < J; ++k) { lo[k] = a_lo(cfg, k); hi[k] = a_hi(cfg, k); }
    *thr_angle = l2_f32(J, lo, hi) / 200.0f;
    for (int k = 0; k < J; ++k) { lo[k] = v_lo(cfg, k); hi[k] = v_hi(cfg, k); }
    *thr_vel = l2_f32(J, lo, hi) / 5.0f;
}

/* Stand-alone GoalEnv.compute_reward(current_state, goal_state) over float32 arrays
 * (roboy_env.py:92-112 as called from :40-49 and by the reference's tests).
 * q,qd,goal_q,goal_qd: [n][J] row-major; feasible: [n] bytes.  goal_qd may be NULL (= the
 * float64 zeros of roboy_env.py:23).  Outputs reward (double) and reached (byte). */
void orc_compute_reward(const orc_cfg *cfg, uint64_t n, const float *q, const float *qd,
                        const uint8_t *feasible, const float *goal_q, const float *goal_qd,
                        double *reward, uint8_t *reached, uint8_t *violation) {
    const int J = cfg_J(cfg);
    float ta, tv;
    orc_thresholds(cfg, &ta, &tv);
    for (uint64_t i = 0; i < n; ++i) {
        orc_state s;
        for (int k = 0; k < J; ++k) { s.q[k] = q[J * i + k]; s.qd[k] = qd[J * i + k]; }
        s.is64 = 0;
        s.feasible = feasible ? feasible[i] : 1;
        const float *gv = goal_qd ? goal_qd + J * i : NULL;
        int r = did_reach_goal(cfg, &s, goal_q + J * i, gv, ta, tv);
        int viol = 0;
        reward[i] = compute_reward(cfg, &s, goal_q + J * i, gv, r, &viol);
        if (reached) reached[i] = (uint8_t)r;
        if (violation) violation[i] = (uint8_t)viol;
    }
}

/* ------------------------------------------------------------------------------------------
 * Batched env state.  Layout mirrors the CUDA path's structure-of-arrays:
 *   goal[J][n] f32, step_flags[n] u32, held[2J][n] f32 (q_0..q_{J-1}, qd_0..qd_{J-1}).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    orc_cfg cfg;
    float thr_angle, thr_vel;
    uint64_t t; /* call counter: 0 = construction, +1 per reset()/step() call */
    float *goal;
    uint32_t *step_flags;
    float *held;
    double stats[ST_N];
    uint32_t err_flags;
    uint64_t first_bad_env;
} orc_env;

orc_env *orc_create(const orc_cfg *cfg) {
    const int J = cfg_J(cfg);
    if (J > ORC_MAX_JOINT || cfg_A(cfg) > ORC_MAX_ACTION) return NULL;
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->cfg = *cfg;
    orc_thresholds(cfg, &e->thr_angle, &e->thr_vel);
    uint64_t n = cfg->n_envs;
    e->goal = (float *)calloc((size_t)J * n + 1, sizeof(float));
    e->step_flags = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
    e->held = (float *)calloc((size_t)2 * J * n + 1, sizeof(float));
    e->first_bad_env = UINT64_MAX;
    e->t = 0;
    /* RoboyEnv.__init__ (roboy_env.py:12-38) over StubSimulationClient.__init__
     * (simulation_client.py:29-31): held state := random sample, goal := random, step_num = 1 */
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t gid = cfg->env_id_base + i;
        float q[JP], qd[JP], g[JP];
        orc_draw_state(cfg, gid, 0, q, qd);
        orc_draw_goal(cfg, gid, 0, g);
        for (int k = 0; k < J; ++k) {
            e->held[(size_t)k * n + i] = q[k];
            e->held[(size_t)(J + k) * n + i] = qd[k];
            e->goal[(size_t)k * n + i] = g[k];
        }
        e->step_flags[i] = 1u;
    }
    return e;
}

void orc_destroy(orc_env *e) {
    if (!e) return;
    free(e->goal); free(e->step_flags); free(e->held); free(e);
}

float *orc_goal(orc_env *e) { return e->goal; }
uint32_t *orc_step_flags(orc_env *e) { return e->step_flags; }
float *orc_held(orc_env *e) { return e->held; }
double *orc_stats(orc_env *e) { return e->stats; }
uint64_t orc_counter(orc_env *e) { return e->t; }
void orc_set_counter(orc_env *e, uint64_t t) { e->t = t; }
uint32_t orc_err_flags(orc_env *e) { return e->err_flags; }
uint64_t orc_first_bad_env(orc_env *e) { return e->first_bad_env; }
void orc_set_reward_range(orc_env *e, double lo, double hi) { e->cfg.reward_lo = lo; e->cfg.reward_hi = hi; }

static void note_error(orc_env *e, uint32_t bits, uint64_t gid) {
    e->err_flags |= bits;
    if (gid < e->first_bad_env) e->first_bad_env = gid;
}

/* RoboyEnv.reset() (roboy_env.py:82-87) over the Stub (simulation_client.py:42-44,33-34),
 * for envs with mask[i] != 0 (mask == NULL: all).  obs [n][3J] rows are written for reset envs
 * only (obs may be NULL). */
void orc_reset(orc_env *e, const uint8_t *mask, float *obs) {
    const uint64_t n = e->cfg.n_envs;
    const int J = cfg_J(&e->cfg), D = 3 * J;
    e->t += 1;
    for (uint64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        float g[JP];
        orc_draw_goal(&e->cfg, e->cfg.env_id_base + i, e->t, g);       /* :86 */
        for (int k = 0; k < J; ++k) e->goal[(size_t)k * n + i] = g[k];
        e->step_flags[i] = 1u | ORC_F_HELD_ZERO64;                      /* :83-85 */
        if (obs) {
            for (int k = 0; k < 2 * J; ++k) obs[D * i + k] = 0.0f;      /* :87 */
            for (int k = 0; k < J; ++k) obs[D * i + 2 * J + k] = g[k];
        }
    }
}

/* RoboyEnv.step (roboy_env.py:51-70) / reset (:82-87) on states handed in by an external simulator
 * (RosSimulationClient's role, ros_simulation_client.py:40-60; wire values are float64 arrays).
 * q, qd [n][J] float32 values; feasible [n] or NULL; reset != 0: masked reset, obs carries the NEW goal. */
void orc_external(orc_env *e, int reset, const uint8_t *mask, const float *q, const float *qd,
                  const uint8_t *feasible, float *obs, float *reward, uint8_t *done) {
    const orc_cfg *cfg = &e->cfg;
    const uint64_t n = cfg->n_envs;
    const int J = cfg_J(cfg), D = 3 * J;
    e->t += 1;
    for (uint64_t i = 0; i < n; ++i) {
        if (reset && mask && !mask[i]) continue;
        const uint64_t gid = cfg->env_id_base + i;
        orc_state s;
        for (int k = 0; k < J; ++k) { s.q[k] = q[J * i + k]; s.qd[k] = qd[J * i + k]; }
        s.is64 = 1;
        s.feasible = feasible ? feasible[i] != 0 : 1;
        /* robot.new_state() in the client (ros_simulation_client.py:40-46) asserts the angles inside the angle space
         * (roboy_robot.py:76: Box.contains, closed interval, NaN fails; the velocity assert is commented out, :77) */
        int bad_state = 0;
        for (int k = 0; k < J; ++k) {
            if (!(q[J * i + k] >= a_lo(cfg, k) && q[J * i + k] <= a_hi(cfg, k))) bad_state = 1;
        }
        if (bad_state && reset) note_error(e, ORC_ERR_STATE_BOUNDS, gid);
        float g[JP];
        for (int k = 0; k < J; ++k) g[k] = e->goal[(size_t)k * n + i];
        uint32_t sf = e->step_flags[i];
        int new_goal = reset;
        if (!reset) {
            int reached = did_reach_goal(cfg, &s, g, NULL, e->thr_angle, e->thr_vel);
            int viol = 0;
            double r = compute_reward(cfg, &s, g, NULL, reached, &viol);
            uint32_t step = sf & ORC_STEP_MASK;
            if (step < ORC_STEP_MASK) step += 1;
            int dn = reached || (int32_t)step > cfg->max_episode_len;
            sf = step | (sf & ~ORC_STEP_MASK);
            reward[i] = (float)r;
            done[i] = (uint8_t)dn;
            new_goal = dn;
            e->stats[ST_STEPS] += 1.0;
            e->stats[ST_SUM_REWARD] += (double)(float)r;
            if (dn) { e->stats[ST_EPISODES] += 1.0; e->stats[reached ? ST_SUCCESSES : ST_TIMEOUTS] += 1.0; e->stats[ST_SUM_EPLEN] += (double)step - 1.0; }
            if (viol || bad_state) {
                e->stats[ST_VIOLATIONS] += 1.0;
                note_error(e, (viol ? ORC_ERR_REWARD_RANGE : 0u) | (bad_state ? ORC_ERR_STATE_BOUNDS : 0u), gid);
            }
            for (int k = 0; k < J; ++k) { obs[D * i + k] = (float)s.q[k]; obs[D * i + J + k] = (float)s.qd[k]; obs[D * i + 2 * J + k] = g[k]; }
        } else {
            sf = 1u | (sf & ~ORC_STEP_MASK);
        }
        if (new_goal) {
            orc_draw_goal(cfg, gid, e->t, g);
            for (int k = 0; k < J; ++k) e->goal[(size_t)k * n + i] = g[k];
        }
        if (reset)
            for (int k = 0; k < J; ++k) { obs[D * i + k] = (float)s.q[k]; obs[D * i + J + k] = (float)s.qd[k]; obs[D * i + 2 * J + k] = g[k]; }
        e->step_flags[i] = sf;
    }
}

typedef struct {
    orc_env *e;
    const float *actions;
    float *obs, *reward, *terminal_obs;
    uint8_t *done;
    uint64_t lo, hi;
    double stats[ST_N];
    uint32_t err;
    uint64_t first_bad;
} step_job;

/* RoboyEnv.step (roboy_env.py:51-70) over StubSimulationClient.forward_step_command
 * (simulation_client.py:36-40), then -- if cfg.auto_reset -- the vec-env worker's
 * reset-on-done, for envs [lo, hi). */
static void step_range(step_job *j) {
    orc_env *e = j->e;
    const orc_cfg *cfg = &e->cfg;
    const uint64_t n = cfg->n_envs;
    const uint64_t t = e->t;
    const int J = cfg_J(cfg), A = cfg_A(cfg), D = 3 * J;
    const float in_hi = 1.0f, in_lo = -1.0f;                    /* roboy_env.py:31 */
    float slope[ORC_MAX_ACTION];
    for (int k = 0; k < A; ++k) slope[k] = (t_hi(cfg, k) - t_lo(cfg, k)) / (in_hi - in_lo); /* :157 float32, per component */
    for (uint64_t i = j->lo; i < j->hi; ++i) {
        const uint64_t gid = cfg->env_id_base + i;
        const float *a = j->actions + (size_t)A * i;
        uint32_t err = 0;
        /* :52 assert action_space.contains(action) -- closed interval, NaN fails */
        int ok = 1, hold = 1;
        for (int k = 0; k < A; ++k) {
            if (!(a[k] >= in_lo && a[k] <= in_hi)) ok = 0;
            float r = a[k] - in_hi;                             /* :158 float32, unfused */
            r = slope[k] * r;
            r = r + t_hi(cfg, k);
            /* simulation_client.py:38 np.allclose(list_of_py_floats, 0): |x| <= 1e-8, NaN/inf fail */
            if (!(fabs((double)r) <= 1e-8)) hold = 0;
        }
        if (!ok) err |= ORC_ERR_ACTION;

        uint32_t sf = e->step_flags[i];
        float g[JP];
        for (int k = 0; k < J; ++k) g[k] = e->goal[(size_t)k * n + i];
        orc_state s;
        if (hold) {                                             /* simulation_client.py:38-39 */
            if (sf & ORC_F_HELD_ZERO64) {
                for (int k = 0; k < J; ++k) { s.q[k] = 0.0; s.qd[k] = 0.0; }
                s.is64 = 1; s.feasible = 1;
            } else {
                for (int k = 0; k < J; ++k) {
                    s.q[k] = e->held[(size_t)k * n + i];
                    s.qd[k] = e->held[(size_t)(J + k) * n + i];
                }
                s.is64 = 0; s.feasible = !(sf & ORC_F_HELD_INFEASIBLE);
            }
        } else {                                                /* :40 fresh sample, not stored */
            float q[JP], qd[JP];
            orc_draw_state(cfg, gid, t, q, qd);
            for (int k = 0; k < J; ++k) { s.q[k] = q[k]; s.qd[k] = qd[k]; }
            s.is64 = 0; s.feasible = 1;
        }
        uint32_t step = (sf & ORC_STEP_MASK);
        if (step < ORC_STEP_MASK) step += 1;                    /* roboy_env.py:60 (saturating) */

        float o[3 * JP];                                        /* :62 -> :75-80 */
        for (int k = 0; k < J; ++k) { o[k] = (float)s.q[k]; o[J + k] = (float)s.qd[k]; o[2 * J + k] = g[k]; }

        int reached = did_reach_goal(cfg, &s, g, NULL, e->thr_angle, e->thr_vel); /* :65 / :105 */
        int viol = 0;
        double rew = compute_reward(cfg, &s, g, NULL, reached, &viol);           /* :64 */
        if (viol) err |= ORC_ERR_REWARD_RANGE;
        int timeout = (int32_t)step > cfg->max_episode_len;                       /* :72-73 */
        int done = reached || timeout;                                           /* :65-66 */

        uint32_t flags = sf & ~ORC_STEP_MASK;
        if (done) {
            float ng[JP];
            /* :67-68 _set_new_goal(); under auto-reset the worker's reset() (:82-87) draws
             * again and only that second goal is observable, so one draw is materialised. */
            orc_draw_goal(cfg, gid, t, ng);
            for (int k = 0; k < J; ++k) e->goal[(size_t)k * n + i] = ng[k];
            j->stats[ST_SUM_EPLEN] += (double)step - 1.0;   /* steps taken since reset(): step_num starts at 1 */
            if (cfg->auto_reset) {
                if (j->terminal_obs) memcpy(j->terminal_obs + (size_t)D * i, o, sizeof(float) * D);
                for (int k = 0; k < 2 * J; ++k) o[k] = 0.0f;
                for (int k = 0; k < J; ++k) o[2 * J + k] = ng[k];
                step = 1;
                flags = ORC_F_HELD_ZERO64;
            }
            j->stats[ST_EPISODES] += 1.0;
            if (reached) j->stats[ST_SUCCESSES] += 1.0;
            else j->stats[ST_TIMEOUTS] += 1.0;
        }
        e->step_flags[i] = step | flags;
        memcpy(j->obs + (size_t)D * i, o, sizeof(float) * D);
        j->reward[i] = (float)rew;
        j->done[i] = (uint8_t)done;
        j->stats[ST_STEPS] += 1.0;
        j->stats[ST_SUM_REWARD] += (double)(float)rew;
        if (hold) j->stats[ST_HOLDS] += 1.0;
        if (err) {
            j->stats[ST_VIOLATIONS] += 1.0;
            j->err |= err;
            if (gid < j->first_bad) j->first_bad = gid;
        }
    }
}

static void *step_thread(void *p) { step_range((step_job *)p); return NULL; }

/* One batched step over all envs with `threads` host threads (>=1).
 * actions [n][A] f32; obs [n][3J] f32; reward [n] f32; done [n] u8; terminal_obs [n][3J] or NULL */
void orc_step(orc_env *e, const float *actions, float *obs, float *reward, uint8_t *done,
              float *terminal_obs, int threads) {
    const uint64_t n = e->cfg.n_envs;
    e->t += 1;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = (int)(n ? n : 1);
    step_job *jobs = (step_job *)calloc((size_t)threads, sizeof(step_job));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int k = 0; k < threads; ++k) {
        jobs[k].e = e; jobs[k].actions = actions; jobs[k].obs = obs; jobs[k].reward = reward;
        jobs[k].done = done; jobs[k].terminal_obs = terminal_obs;
        jobs[k].lo = n * (uint64_t)k / (uint64_t)threads;
        jobs[k].hi = n * (uint64_t)(k + 1) / (uint64_t)threads;
        jobs[k].first_bad = UINT64_MAX;
        if (threads > 1) pthread_create(&tid[k], NULL, step_thread, &jobs[k]);
        else step_range(&jobs[k]);
    }
    for (int k = 0; k < threads; ++k) {
        if (threads > 1) pthread_join(tid[k], NULL);
        for (int s = 0; s < ST_N; ++s) e->stats[s] += jobs[k].stats[s];
        if (jobs[k].err) note_error(e, jobs[k].err, jobs[k].first_bad);
    }
    free(jobs); free(tid);
}
