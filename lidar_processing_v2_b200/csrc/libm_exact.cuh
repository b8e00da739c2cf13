// Device (and host-testable) restatements of the three glibc 2.39 libm routines whose results
// are observable in the reference hot path:
//
//   atan2f  -> Dataloader::addRingInfo quadrant test      (src/dataloader/src/dataloader.cpp:96)
//              Clusterer azimuth                           (lidar_processing_lib/src/clusterer.cpp:74)
//   atanf   -> Clusterer elevation                         (clusterer.cpp:86)
//              Segmenter ring-less height index            (segmenter.cpp:157,182)
//   expf    -> JCP neighbour weights                       (segmenter.cpp:588)
//
// CUDA's own atan2f/atanf/expf differ from glibc's by 1-2 ulp on a large fraction of inputs, which
// would move points across voxel / quadrant boundaries and flip JCP votes. These functions follow
// the published algorithms glibc 2.39 uses on x86-64 (fdlibm's single-precision atan/atan2 from
// Sun Microsystems; the Arm Optimized Routines double-precision-core expf with a 32-entry table)
// operation by operation, so the results are bit-identical to the host library as long as the
// compiler does not contract float multiplies/adds (build with -fmad=false; the explicit fma()
// calls in expf mirror glibc's FMA multiarch variant, which is what ifunc selects on any host
// CPU with FMA). tests/test_libm_exact.py compiles this header with g++ and compares it with the
// running libm (random + exhaustive-range sweeps).
#pragma once

#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define LPL_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define LPL_HD inline
#endif

namespace lpl
{
LPL_HD std::uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    std::uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
#endif
}

LPL_HD float u2f(std::uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    std::memcpy(&f, &u, 4);
    return f;
#endif
}

LPL_HD std::uint64_t d2u(double d)
{
#if defined(__CUDA_ARCH__)
    return static_cast<std::uint64_t>(__double_as_longlong(d));
#else
    std::uint64_t u;
    std::memcpy(&u, &d, 8);
    return u;
#endif
}

LPL_HD double u2d(std::uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(static_cast<long long>(u));
#else
    double d;
    std::memcpy(&d, &u, 8);
    return d;
#endif
}

// ---------------------------------------------------------------- atanf (fdlibm s_atanf.c)
LPL_HD float atanf_glibc(float x)
{
    const float atanhi0 = 4.6364760399e-01f, atanhi1 = 7.8539812565e-01f;
    const float atanhi2 = 9.8279368877e-01f, atanhi3 = 1.5707962513e+00f;
    const float atanlo0 = 5.0121582440e-09f, atanlo1 = 3.7748947079e-08f;
    const float atanlo2 = 3.4473217170e-08f, atanlo3 = 7.5497894159e-08f;
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f;
    const float aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f;
    const float aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f;
    const float aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const float one = 1.0f;

    const std::int32_t hx = static_cast<std::int32_t>(f2u(x));
    const std::int32_t ix = hx & 0x7fffffff;
    int id;
    float hi = 0.f, lo = 0.f;
    if (ix >= 0x4c000000) // |x| >= 2^25
    {
        if (ix > 0x7f800000)
        {
            return x + x; // NaN
        }
        return (hx > 0) ? (atanhi3 + atanlo3) : (-atanhi3 - atanlo3);
    }
    if (ix < 0x3ee00000) // |x| < 0.4375
    {
        if (ix < 0x31000000) // |x| < 2^-29
        {
            return x;
        }
        id = -1;
    }
    else
    {
        x = u2f(static_cast<std::uint32_t>(ix)); // fabsf
        if (ix < 0x3f980000) // |x| < 1.1875
        {
            if (ix < 0x3f300000) // 7/16 <= |x| < 11/16
            {
                id = 0;
                hi = atanhi0;
                lo = atanlo0;
                x = (2.0f * x - one) / (2.0f + x);
            }
            else // 11/16 <= |x| < 19/16
            {
                id = 1;
                hi = atanhi1;
                lo = atanlo1;
                x = (x - one) / (x + one);
            }
        }
        else
        {
            if (ix < 0x401c0000) // |x| < 2.4375
            {
                id = 2;
                hi = atanhi2;
                lo = atanlo2;
                x = (x - 1.5f) / (one + 1.5f * x);
            }
            else // 2.4375 <= |x| < 2^25
            {
                id = 3;
                hi = atanhi3;
                lo = atanlo3;
                x = -1.0f / x;
            }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0)
    {
        return x - x * (s1 + s2);
    }
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return (hx < 0) ? -r : r;
}

// ---------------------------------------------------------------- atan2f (fdlibm e_atan2f.c)
LPL_HD float atan2f_glibc(float y, float x)
{
    const float tiny = 1.0e-30f;
    const float pi_o_4 = 7.8539818525e-01f;
    const float pi_o_2 = 1.5707963705e+00f;
    const float pi = 3.1415927410e+00f;
    const float pi_lo = -8.7422776573e-08f;

    const std::int32_t hx = static_cast<std::int32_t>(f2u(x));
    const std::int32_t hy = static_cast<std::int32_t>(f2u(y));
    const std::int32_t ix = hx & 0x7fffffff;
    const std::int32_t iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000)
    {
        return x + y; // NaN
    }
    if (hx == 0x3f800000)
    {
        return atanf_glibc(y); // x == 1.0
    }
    const std::int32_t m = ((hy >> 31) & 1) | ((hx >> 30) & 2); // 2*sign(x) + sign(y)
    if (iy == 0)
    {
        switch (m)
        {
        case 0:
        case 1:
            return y;
        case 2:
            return pi + tiny;
        default:
            return -pi - tiny;
        }
    }
    if (ix == 0)
    {
        return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    }
    if (ix == 0x7f800000)
    {
        if (iy == 0x7f800000)
        {
            switch (m)
            {
            case 0:
                return pi_o_4 + tiny;
            case 1:
                return -pi_o_4 - tiny;
            case 2:
                return 3.0f * pi_o_4 + tiny;
            default:
                return -3.0f * pi_o_4 - tiny;
            }
        }
        switch (m)
        {
        case 0:
            return 0.0f;
        case 1:
            return -0.0f;
        case 2:
            return pi + tiny;
        default:
            return -pi - tiny;
        }
    }
    if (iy == 0x7f800000)
    {
        return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    }
    const std::int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60)
    {
        z = pi_o_2 + 0.5f * pi_lo;
    }
    else if (hx < 0 && k < -60)
    {
        z = 0.0f;
    }
    else
    {
        const float q = y / x;
        z = atanf_glibc(u2f(f2u(q) & 0x7fffffffU));
    }
    switch (m)
    {
    case 0:
        return z;
    case 1:
        return u2f(f2u(z) ^ 0x80000000U);
    case 2:
        return pi - (z - pi_lo);
    default:
        return (z - pi_lo) - pi;
    }
}

// ---------------------------------------------------------------- expf (Arm Optimized Routines)
// Valid for the range the hot path uses (|x| < 88, no overflow / underflow handling): JCP calls
// expf(-amplification * distance) with distance <= kernel_threshold (segmenter.cpp:578-588).
#if defined(__CUDACC__)
__device__ __constant__ std::uint64_t kExp2fTab[32] = {
#else
static const std::uint64_t kExp2fTab[32] = {
#endif
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL,
};

#if defined(__CUDA_ARCH__)
#define LPL_EXP2F_TAB(i) kExp2fTab[i]
#elif defined(__CUDACC__)
// host pass of nvcc: the table above is a device symbol; host code never calls expf_glibc
#define LPL_EXP2F_TAB(i) 0ULL
#else
#define LPL_EXP2F_TAB(i) kExp2fTab[i]
#endif

LPL_HD float expf_glibc(float x)
{
    const double N = 32.0;
    const double InvLn2N = 0x1.71547652b82fep+0 * N;
    const double Shift = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / N / N / N;
    const double C1 = 0x1.ebfce50fac4f3p-3 / N / N;
    const double C2 = 0x1.62e42ff0c52d6p-1 / N;
    const double xd = static_cast<double>(x);
    double z = InvLn2N * xd;
    double kd = z + Shift;
    const std::uint64_t ki = d2u(kd);
    kd -= Shift;
    const double r = z - kd;
    std::uint64_t t = LPL_EXP2F_TAB(ki % 32);
    t += ki << (52 - 5);
    const double s = u2d(t);
    z = fma(C0, r, C1);
    const double r2 = r * r;
    double y = fma(C2, r, 1.0);
    y = fma(z, r2, y);
    y = y * s;
    return static_cast<float>(y);
}

// ---------------------------------------------------------------- atan2Approx
// lidar_processing_lib/include/lidar_processing_lib/common.hpp:33-62 — a minimax polynomial in
// plain float operations; restated with the same operation order (no contraction).
LPL_HD float atan2_approx(float y, float x)
{
    const float ax = fabsf(x);
    const float ay = fabsf(y);
    const float mx = fmaxf(ay, ax);
    const float mn = fminf(ay, ax);
    const float a = mn / mx;
    const float s = a * a;
    const float c = s * a;
    const float q = s * s;
    float r = 0.024840285F * q + 0.18681418F;
    const float t = -0.094097948F * q - 0.33213072F;
    r = r * s + t;
    r = r * c + a;
    if (ay > ax)
    {
        r = 1.57079637F - r;
    }
    if (x < 0)
    {
        r = 3.14159274F - r;
    }
    if (y < 0)
    {
        r = -r;
    }
    return r;
}
} // namespace lpl
