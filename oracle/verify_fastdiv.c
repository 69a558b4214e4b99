/*
 * oracle/verify_fastdiv.c -- TEST INFRASTRUCTURE: proof by exhaustion for the CUDA step kernel's
 * division-by-constant.
 *
 * The reference normalises with an IEEE float32 division by (max - min)
 * (gym_roboy/envs/robots/roboy_robot.py:93-95).  The kernel's hot path replaces  x / c  by
 *     q0 = RN(x * rc);  r = fma(-q0, c, x);  q = fma(r, rc, q0)        with rc = RN(1 / c)
 * which is 3 instructions instead of ~10.  That sequence is NOT correctly rounded for every
 * (x, c): it fails when the residual r goes subnormal (|x| below about 2^-100) and it loses
 * the sign of -0.  This program checks, for the given divisor c, x = +0 and EVERY float32 x
 * with lo <= |x| <= hi, and reports the number of mismatches against the hardware IEEE
 * division.  Exit code 0 iff there are none.
 *
 * The kernel uses the fast form only on its sampled-state path with the MSJ constants, where
 * the numerators are  t = (2*v - max) - min  with v, max, min of magnitude <= pi: every such
 * t is 0 or a multiple of 2^-25, and |t| <= 2*pi + ulp.  [2^-30, 8] covers that with margin.
 *
 * usage: verify_fastdiv <c as float hex bits> <lo> <hi>      (gcc -O2 -mfma -fopenmp)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const float c = bits2f((uint32_t)strtoul(argv[1], NULL, 16));
    const float lo = (float)atof(argv[2]), limit = (float)atof(argv[3]);
    const float rc = (float)(1.0 / (double)c);
    const uint32_t bottom = f2bits(lo), top = f2bits(limit);
    unsigned long long bad = 0, n = 0;
    {
        const float q0 = 0.0f * rc, r = fmaf(-q0, c, 0.0f), q = fmaf(r, rc, q0);
        if (f2bits(q) != f2bits(0.0f / c)) ++bad;
    }
#pragma omp parallel for reduction(+ : bad, n) schedule(static)
    for (uint32_t u = bottom; u <= top; ++u) {
        for (int sign = 0; sign < 2; ++sign) {
            const float x = bits2f(u | ((uint32_t)sign << 31));
            volatile float want = x / c;
            const float q0 = x * rc;
            const float r = fmaf(-q0, c, x);
            const float q = fmaf(r, rc, q0);
            if (f2bits(q) != f2bits(want)) {
                if (bad < 5) fprintf(stderr, "mismatch x=%a c=%a: fast=%a ieee=%a\n", x, c, q, want);
                ++bad;
            }
            ++n;
        }
    }
    printf("c=%a rc=%a %g<=|x|<=%g and x=0: %llu values checked, %llu mismatches\n", c, rc, lo, limit, n + 1, bad);
    return bad ? 1 : 0;
}
