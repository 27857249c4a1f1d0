// wn_math.cuh -- pinned fp32/fp64 transcendentals for the sm_100a WaveNet sample-loop kernel.
//
// The reference evaluates tanh/sigmoid (wavenet/model.py:86), exp/log (wavenet/mixture.py:103-111,
// generate.py:219-222) and a float64 softmax (wavenet/model.py:243) with whatever TensorFlow/numpy
// build is installed.  DESIGN.md "Pinned arithmetic" fixes each of them to an explicit sequence of
// IEEE-754 round-to-nearest operations.  Every operation below is an explicit _rn intrinsic, so nvcc
// can neither contract (a*b+c -> fma) nor reassociate; the result is a pure function of the input bits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wn {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// e^x: x < -87 -> 0; x > 88 -> 88; n = RN(x*log2e) by magic add; Cody-Waite r; degree-5 polynomial.
__device__ __forceinline__ float exp32(float x)
{
    // branch-free: evaluate on the clamped argument, select +0 for x < -87 at the end (same bits as the
    // early-return form of the specification; no divergence, so two independent calls interleave)
    const bool tiny = x < -87.0f;
    float xc = tiny ? -87.0f : x;
    xc = (xc > 88.0f) ? 88.0f : xc;
    const float magic = 12582912.0f;                    // 1.5 * 2^23
    float n = fsub(ffma(xc, 1.44269504088896341f, magic), magic);
    float r = ffma(n, -0.693359375f, xc);
    r = ffma(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = ffma(p, r, 1.3981999507e-3f);
    p = ffma(p, r, 8.3334519073e-3f);
    p = ffma(p, r, 4.1665795894e-2f);
    p = ffma(p, r, 1.6666665459e-1f);
    p = ffma(p, r, 5.0000001201e-1f);
    float e = fadd(ffma(p, fmul(r, r), r), 1.0f);
    int ni = __float2int_rz(n);
    float res = __uint_as_float(__float_as_uint(e) + ((unsigned)ni << 23));
    return tiny ? 0.0f : res;
}

__device__ __forceinline__ float sigmoid32(float x)
{
    return fdiv(1.0f, fadd(1.0f, exp32(-x)));
}

__device__ __forceinline__ float tanh32(float x)
{
    float ax = fabsf(x);
    float r = 1.0f;
    if (!(ax > 44.0f)) {
        float e = exp32(fadd(ax, ax));
        r = fsub(1.0f, fdiv(2.0f, fadd(e, 1.0f)));
    }
    return copysignf(r, x);
}

// natural log, x >= 0; log(0) = -inf; subnormals pre-scaled by 2^23.
__device__ __forceinline__ float log32(float x)
{
    if (x == 0.0f) return __int_as_float(0xff800000);
    int eadj = 0;
    if (x < 1.17549435e-38f) { x = fmul(x, 8388608.0f); eadj = -23; }
    unsigned b = __float_as_uint(x);
    int e = (int)((b >> 23) & 0xffu) - 126 + eadj;
    float m = __uint_as_float((b & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = fsub(fadd(m, m), 1.0f); }
    else { m = fsub(m, 1.0f); }
    float z = fmul(m, m);
    float y = 7.0376836292e-2f;
    y = ffma(y, m, -1.1514610310e-1f);
    y = ffma(y, m, 1.1676998740e-1f);
    y = ffma(y, m, -1.2420140846e-1f);
    y = ffma(y, m, 1.4249322787e-1f);
    y = ffma(y, m, -1.6668057665e-1f);
    y = ffma(y, m, 2.0000714765e-1f);
    y = ffma(y, m, -2.4999993993e-1f);
    y = ffma(y, m, 3.3333331174e-1f);
    y = fmul(fmul(y, m), z);
    float fe = (float)e;
    y = ffma(-2.12194440e-4f, fe, y);
    y = ffma(-0.5f, z, y);
    float r = fadd(m, y);
    return ffma(0.693359375f, fe, r);
}

__device__ __forceinline__ float log1p32(float v)
{
    float u = fadd(1.0f, v);
    if (u == 1.0f) return v;
    return fmul(log32(u), fdiv(v, fsub(u, 1.0f)));
}

// numpy's npy_logaddexpf
__device__ __forceinline__ float logaddexp32(float a, float b)
{
    if (a == b) return fadd(a, 0.693147180559945309f);
    float d = fsub(a, b);
    if (d > 0.0f) return fadd(a, log1p32(exp32(-d)));
    if (d <= 0.0f) return fadd(b, log1p32(exp32(d)));
    return d;
}

// e^x in fp64: x < -708 -> 0; x > 709 -> 709; degree-13 Taylor in r, Horner with fma.
__device__ __forceinline__ double exp64(double x)
{
    if (x < -708.0) return 0.0;
    x = (x > 709.0) ? 709.0 : x;
    const double magic = 6755399441055744.0;            // 1.5 * 2^52
    double n = __dsub_rn(__fma_rn(x, 1.4426950408889634074, magic), magic);
    double r = __fma_rn(n, -6.93147180369123816490e-01, x);
    r = __fma_rn(n, -1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = __fma_rn(p, r, 1.0 / 479001600.0);
    p = __fma_rn(p, r, 1.0 / 39916800.0);
    p = __fma_rn(p, r, 1.0 / 3628800.0);
    p = __fma_rn(p, r, 1.0 / 362880.0);
    p = __fma_rn(p, r, 1.0 / 40320.0);
    p = __fma_rn(p, r, 1.0 / 5040.0);
    p = __fma_rn(p, r, 1.0 / 720.0);
    p = __fma_rn(p, r, 1.0 / 120.0);
    p = __fma_rn(p, r, 1.0 / 24.0);
    p = __fma_rn(p, r, 1.0 / 6.0);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    long long ni = __double2ll_rz(n);
    return __longlong_as_double((long long)((unsigned long long)__double_as_longlong(p) + ((unsigned long long)ni << 52)));
}

}  // namespace wn
