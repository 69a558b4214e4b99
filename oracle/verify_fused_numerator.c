/*
 * oracle/verify_fused_numerator.c -- TEST INFRASTRUCTURE: proof by exhaustion for a contraction ptxas performs in the
 * CUDA kernels' normalisation.
 *
 * The reference normalises with  (2*v - max - min) / (max - min)  evaluated left to right in float32
 * (gym_roboy/envs/robots/roboy_robot.py:93-95): t1 = fl(2*v); t2 = fl(t1 - max); t = fl(t2 - min).  The kernels spell
 * exactly that with round-to-nearest intrinsics and are compiled with -fmad=false -- and ptxas still emits
 * FFMA t2, v, 2, -max for it (cuobjdump -sass: "FFMA R28, R12, 2, -R26"), because the contraction preserves the value:
 * 2*v is exact in binary floating point (the exponent moves, the significand does not; true for subnormals as well), so
 * both forms round the same real number once.  The only exception is an overflowing 2*v that max pulls back under
 * FLT_MAX's rounding boundary, which needs |max| >= 2^103; roboy_create refuses robots with a bound of 2^100 or more.
 *
 * This program checks the identity for EVERY float32 v (all 2^32 bit patterns; NaN results compared as NaN) and the
 * given max.  Exit code 0 iff there is no mismatch.
 *
 * usage: verify_fused_numerator <max as float hex bits>      (gcc -O2 -mfma -ffp-contract=off -fopenmp)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const float hi = bits2f((uint32_t)strtoul(argv[1], NULL, 16));
    unsigned long long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long i = 0; i < (1ll << 32); ++i) {
        const float v = bits2f((uint32_t)i);
        volatile float t1 = 2.0f * v;          /* volatile: no contraction, no reassociation */
        volatile float want = t1 - hi;
        const float got = fmaf(2.0f, v, -hi);
        const int both_nan = (want != want) && (got != got);
        if (!both_nan && f2bits(got) != f2bits(want)) {
            if (bad < 5) fprintf(stderr, "mismatch v=%a max=%a: fused=%a unfused=%a\n", v, hi, got, (float)want);
            ++bad;
        }
    }
    printf("max=%a: 4294967296 values of v checked, %llu mismatches\n", hi, bad);
    return bad ? 1 : 0;
}
