/*
 * oracle/wn_math_ref.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Pinned transcendental functions for the CPU restatement of the reference's
 * WaveNet sample loop.  The reference evaluates tanh / sigmoid / exp / log with
 * whatever TensorFlow-Eigen / numpy / libm build the user has (unpinned:
 * wavenet/model.py:86, wavenet/mixture.py:103-111, generate.py:219-222).  To make
 * "bit-exact integer samples" a testable statement we pin each function to a
 * fixed sequence of IEEE-754 operations (+, -, *, /, fma, integer bit ops) that
 * any conforming CPU or GPU evaluates identically.  DESIGN.md section "Pinned
 * arithmetic" is the specification; this file and the product's
 * csrc/wn_math.cuh are two independent renderings of that specification.
 *
 * Compile with -ffp-contract=off (the Makefile does) so the compiler never
 * fuses or splits an operation that the specification spells out.
 */
#ifndef WN_MATH_REF_H
#define WN_MATH_REF_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t orc_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float orc_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint64_t orc_d2u(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double orc_u2d(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

/* exp32: e^x, fp32.  Spec:
 *   x < -87      -> +0
 *   x  > 88      -> treated as 88
 *   n  = RN(x*log2e) via the 1.5*2^23 magic-add;  r = x - n*ln2 (two-term, fma)
 *   e^r by the degree-5 Cephes polynomial in r, evaluated with fma (Horner)
 *   result = e^r * 2^n by adding n to the exponent field.                       */
static inline float orc_exp32(float x)
{
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    float t = fmaf(x, 1.44269504088896341f, 12582912.0f);
    float n = t - 12582912.0f;
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = fmaf(p, r2, r) + 1.0f;
    int32_t ni = (int32_t)n;
    return orc_u2f(orc_f2u(e) + ((uint32_t)ni << 23));
}

/* sigmoid32: 1/(1+e^-x) with an IEEE division. */
static inline float orc_sigmoid32(float x)
{
    return 1.0f / (1.0f + orc_exp32(-x));
}

/* tanh32: sign(x) * (1 - 2/(e^{2|x|}+1)); |x| > 44 -> +-1. */
static inline float orc_tanh32(float x)
{
    float ax = fabsf(x);
    float r;
    if (ax > 44.0f) {
        r = 1.0f;
    } else {
        float e = orc_exp32(ax + ax);
        r = 1.0f - 2.0f / (e + 1.0f);
    }
    return copysignf(r, x);
}

/* log32: natural log, fp32, x >= 0.  log(0) = -inf.  Cephes logf structure:
 * x = m * 2^e with m in [sqrt(1/2), sqrt(2)); polynomial in (m-1).           */
static inline float orc_log32(float x)
{
    if (x == 0.0f) return -INFINITY;
    int32_t eadj = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; eadj = -23; }
    uint32_t b = orc_f2u(x);
    int32_t e = (int32_t)((b >> 23) & 0xffu) - 126 + eadj;
    float m = orc_u2f((b & 0x007fffffu) | 0x3f000000u);      /* [0.5, 1) */
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; }
    else { m = m - 1.0f; }
    float z = m * m;
    float y = 7.0376836292e-2f;
    y = fmaf(y, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = (y * m) * z;
    float fe = (float)e;
    y = fmaf(-2.12194440e-4f, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(0.693359375f, fe, r);
    return r;
}

/* log1p32 for v in [0, 1]: u = 1+v; u==1 -> v; else log(u) * (v/(u-1)). */
static inline float orc_log1p32(float v)
{
    float u = 1.0f + v;
    if (u == 1.0f) return v;
    return orc_log32(u) * (v / (u - 1.0f));
}

/* logaddexp32 after numpy's npy_logaddexpf (generate.py:221 np.logaddexp). */
static inline float orc_logaddexp32(float a, float b)
{
    if (a == b) return a + 0.693147180559945309f;
    float d = a - b;
    if (d > 0.0f) return a + orc_log1p32(orc_exp32(-d));
    if (d <= 0.0f) return b + orc_log1p32(orc_exp32(d));
    return d; /* NaN */
}

/* exp64: e^x in fp64 for the float64 softmax (wavenet/model.py:243).
 *   x < -708 -> +0 (the fp32 cast of the quotient is 0 there anyway);  x > 709 -> 709
 *   n = RN(x*log2e) via 1.5*2^52 magic; r = x - n*ln2 (two-term fma);
 *   degree-13 Taylor polynomial, Horner with fma; scale by exponent add.     */
static inline double orc_exp64(double x)
{
    if (x < -708.0) return 0.0;
    if (x > 709.0) x = 709.0;
    double t = fma(x, 1.4426950408889634074, 6755399441055744.0);
    double n = t - 6755399441055744.0;
    double r = fma(n, -6.93147180369123816490e-01, x);
    r = fma(n, -1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int64_t ni = (int64_t)n;
    return orc_u2d(orc_d2u(p) + ((uint64_t)ni << 52));
}

#endif
