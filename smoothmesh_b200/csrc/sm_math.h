// sm_math.h -- scalar/vector FP64 primitives shared by the CUDA kernels and by
// host code that must reproduce their results bit for bit.
//
// Why this exists: the reference evaluates its angle constraints with
// std::acos (src/smoothMesh.C:783, :993, :995).  glibc's acos and CUDA's acos
// differ in the last ulp for some arguments, which would make freeze decisions
// on the GPU and on the CPU restatement diverge on razor-edge inputs.  Every
// other operation the hot path uses (+ - * / sqrt) is IEEE-754 correctly
// rounded on both sides as long as no FMA contraction happens
// (nvcc --fmad=false, gcc -ffp-contract=off), so one shared acos closes the gap.
//
// sm_acos follows the classical argument-reduction scheme for arccosine
// (rational minimax R(z) ~ (asin(x)-x)/x^3 on |x|<0.5, half-angle identities
// with a split square root elsewhere); its error is < 1 ulp and
// tests/test_math.py pins it against libm.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define SM_HD __host__ __device__ __forceinline__
#else
#define SM_HD inline
#endif

// OpenFOAM tolerances the reference relies on (src/smoothMeshCommon.H:14-17 use
// GREAT; src/smoothMesh.C:259,266 use VSMALL).  Values recalled from OpenFOAM's
// scalar.H: they are not defined inside the reference tree.
#define SM_VSMALL 1.0e-300
#define SM_SMALL 1.0e-15
#define SM_GREAT 1.0e+15
#define SM_VGREAT 1.0e+300
#define SM_ROOTVSMALL 1.0e-150
#define SM_PI 3.14159265358979323846 /* M_PI */
#define SM_COS_CLAMP 0.99999         /* src/smoothMesh.C:781, :991 */

SM_HD uint64_t sm_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, sizeof(u));
    return u;
#endif
}

SM_HD double sm_from_bits(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, sizeof(x));
    return x;
#endif
}

SM_HD double sm_sqrt(double x)
{
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return sqrt(x);
#endif
}

// Rational part shared by the three branches: R(z) = z*P(z)/Q(z).
SM_HD double sm_acos_R(double z)
{
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return p / q;
}

SM_HD double sm_acos(double x)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17,
                 pi = 3.14159265358979311600e+00;
    const uint64_t b = sm_bits(x);
    const uint32_t hx = (uint32_t)(b >> 32), lx = (uint32_t)b;
    const uint32_t ix = hx & 0x7fffffffu;
    if (ix >= 0x3ff00000u)
    { // |x| >= 1
        if (((ix - 0x3ff00000u) | lx) == 0u)
            return (hx >> 31) ? pi + 2.0 * pio2_lo : 0.0;
        return (x - x) / (x - x); // NaN
    }
    if (ix < 0x3fe00000u)
    { // |x| < 0.5
        if (ix <= 0x3c600000u)
            return pio2_hi + pio2_lo;
        const double r = sm_acos_R(x * x);
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (hx >> 31)
    { // x < -0.5
        const double z = (1.0 + x) * 0.5;
        const double r = sm_acos_R(z);
        const double s = sm_sqrt(z);
        const double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    { // x > 0.5
        const double z = (1.0 - x) * 0.5;
        const double s = sm_sqrt(z);
        const double df = sm_from_bits(sm_bits(s) & 0xffffffff00000000ull);
        const double c = (z - df * df) / (s + df);
        const double r = sm_acos_R(z);
        const double w = r * s + c;
        return 2.0 * (df + w);
    }
}

// std::max(-MAX, std::min(MAX, c)) exactly as libstdc++ evaluates it
// (src/smoothMesh.C:782, :992, :994): std::min(a,b) = (b<a)?b:a and
// std::max(a,b) = (a<b)?b:a, so a NaN cosine becomes +MAX.
SM_HD double sm_clamp_cos(double c)
{
    const double MAXC = SM_COS_CLAMP;
    const double t = (c < MAXC) ? c : MAXC;
    return (-MAXC < t) ? t : -MAXC;
}

// OpenFOAM's scalar equality, used by vector operator== (mag(a-b) <= VSMALL).
SM_HD bool sm_equal(double a, double b) { return fabs(a - b) <= SM_VSMALL; }
