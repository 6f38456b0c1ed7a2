// ciede_fast.h -- ColorUtilities::lab_ciede00 (/root/reference/src/color_utilities.cpp:190-294) with the
// FP64 libm calls (atan2, sin, cos, exp, pow) replaced by branch-free evaluations that one GPU thread can
// schedule as a single basic block: the merge loop's critical path carries ONE such evaluation per merge,
// and the library routines (slow-path branches, Horner chains behind calls) cost ~17k cycles there.
//
// Expression order and the float/double mix of every non-transcendental step follow the reference line by
// line; each elementary function below is accurate to ~1 ulp(double) on the argument range the formula can
// produce, so the float result differs from a glibc evaluation only when the double value falls within
// ~2 ulp(double) of a float rounding boundary (about one evaluation in 10^8).
// Host-compilable: tests/test_ciede_fast.py builds this header with g++ and checks it against the oracle.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define F3PS_HD __host__ __device__ __forceinline__
#else
#define F3PS_HD inline
#endif

namespace f3ps_fastmath {

F3PS_HD double pow2i(int k) {                    // 2^k, -1022 < k < 1024
    const uint64_t bits = (uint64_t)(k + 1023) << 52;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)bits);
#else
    double d; memcpy(&d, &bits, 8); return d;
#endif
}

// Division and square root.  On the device: reciprocal / reciprocal-square-root seed (20+ bits) and Newton steps with
// a final residual correction, no special-case branch -- the compiler's IEEE sequences end in a slow-path call that
// splits the formula into ~20 basic blocks and leaves one warp nothing to overlap.  Results are within 1 ulp of the
// correctly rounded ones the host build produces (operands here are normal, finite and, for sqrt, non-negative).
#if defined(__CUDA_ARCH__)
F3PS_HD double ddiv(double n, double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0); y = fma(y, e, y);
    e = fma(-x, y, 1.0); y = fma(y, e, y);
    const double q = n * y;
    return fma(fma(-x, q, n), y, q);
}
F3PS_HD double dsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    g = fma(fma(-g, g, x), h, g);
    return x > 0.0 ? g : (x == 0.0 ? 0.0 : x * __longlong_as_double(0x7ff8000000000000ll));   // sqrt(+-0) = 0; negative or NaN -> NaN, as sqrt() does
}
#else
F3PS_HD double ddiv(double n, double x) { return n / x; }
F3PS_HD double dsqrt(double x) { return sqrt(x); }
#endif

// sin and cos of |x| < 2^20: Cody-Waite reduction by pi/2 (33-bit head), Taylor kernels on [-pi/4, pi/4]
F3PS_HD void sincos_fast(double x, double& s, double& c) {
    const double two_over_pi = 0x1.45f306dc9c883p-1;
    const double pio2_hi = 0x1.921fb54400000p+0;    // first 33 bits of pi/2
    const double pio2_lo = 0x1.0b4611a626331p-34;    // pi/2 - pio2_hi
    const double kd = rint(x * two_over_pi);
    double r = fma(-kd, pio2_hi, x);
    r = fma(-kd, pio2_lo, r);
    const int k = (int)kd;
    const double r2 = r * r;
    // sin r = r + r^3 (S3 + r^2 (S5 + ... S17))
    double ps = 1.0 / 355687428096000.0;                  // 1/17!
    ps = fma(ps, r2, -1.0 / 1307674368000.0);             // -1/15!
    ps = fma(ps, r2, 1.0 / 6227020800.0);                 // 1/13!
    ps = fma(ps, r2, -1.0 / 39916800.0);                  // -1/11!
    ps = fma(ps, r2, 1.0 / 362880.0);                     // 1/9!
    ps = fma(ps, r2, -1.0 / 5040.0);                      // -1/7!
    ps = fma(ps, r2, 1.0 / 120.0);                        // 1/5!
    ps = fma(ps, r2, -1.0 / 6.0);                         // -1/3!
    const double sr = fma(r * r2, ps, r);
    // cos r = 1 - r^2/2 + r^4 (C4 + r^2 (C6 + ... C18))
    double pcs = -1.0 / 6402373705728000.0;               // -1/18!
    pcs = fma(pcs, r2, 1.0 / 20922789888000.0);           // 1/16!
    pcs = fma(pcs, r2, -1.0 / 87178291200.0);             // -1/14!
    pcs = fma(pcs, r2, 1.0 / 479001600.0);                // 1/12!
    pcs = fma(pcs, r2, -1.0 / 3628800.0);                 // -1/10!
    pcs = fma(pcs, r2, 1.0 / 40320.0);                    // 1/8!
    pcs = fma(pcs, r2, -1.0 / 720.0);                     // -1/6!
    pcs = fma(pcs, r2, 1.0 / 24.0);                       // 1/4!
    const double cr = fma(r2 * r2, pcs, fma(-0.5, r2, 1.0));
    const double s0 = (k & 1) ? cr : sr, c0 = (k & 1) ? sr : cr;
    s = (k & 2) ? -s0 : s0;
    c = ((k + 1) & 2) ? -c0 : c0;
}
F3PS_HD double sin_fast(double x) { double s, c; sincos_fast(x, s, c); return s; }
F3PS_HD double cos_fast(double x) { double s, c; sincos_fast(x, s, c); return c; }

// atan(i/8), i = 0..8, as double-double (device copy in constant memory)
#define F3PS_ATAN_HI {0x0.0p+0, 0x1.fd5ba9aac2f6ep-4, 0x1.f5b75f92c80ddp-3, 0x1.6f61941e4def1p-2, 0x1.dac670561bb4fp-2, 0x1.1e00babdefeb4p-1, 0x1.4978fa3269ee1p-1, 0x1.700a7c5784634p-1, 0x1.921fb54442d18p-1}
#define F3PS_ATAN_LO {0x0.0p+0, -0x1.cd37686760c17p-59, 0x1.8ab6e3cf7afbdp-57, -0x1.c63aae6f6e918p-56, 0x1.a2b7f222f65e2p-56, -0x1.928df287a668fp-58, 0x1.2419a87f2a458p-56, -0x1.8c34d25aadef6p-56, 0x1.1a62633145c07p-55}
static const double kAtanHiHost[9] = F3PS_ATAN_HI;
static const double kAtanLoHost[9] = F3PS_ATAN_LO;
#if defined(__CUDACC__)
__device__ __constant__ double kAtanHiDev[9] = F3PS_ATAN_HI;
__device__ __constant__ double kAtanLoDev[9] = F3PS_ATAN_LO;
#endif

// atan2 for finite arguments that are not both zero; the ratio is moved to |t| <= 1/16 around one of nine
// breakpoints i/8 with a single division: atan(m/M) = atan(i/8) + atan((m - cM)/(M + cm)), c = i/8.
F3PS_HD double atan2_fast(double y, double x) {
    const double pio2 = 0x1.921fb54442d18p+0, pio2_tail = 0x1.1a62633145c07p-54;
    const double pi = 0x1.921fb54442d18p+1, pi_tail = 0x1.1a62633145c07p-53;
    const double ax = fabs(x), ay = fabs(y);
    const double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
#if defined(__CUDA_ARCH__)
    const float tf = __fdividef((float)mn, (float)mx);      // only picks the breakpoint; any nearby one works
#else
    const float tf = (float)mn / (float)mx;
#endif
    const int i = (int)rintf(tf * 8.0f);
    const double cc = (double)i * 0.125;
    const double num = fma(-cc, mx, mn), den = fma(cc, mn, mx);
    const double t = ddiv(num, den);
    const double t2 = t * t;
    double p = 1.0 / 17.0;
    p = fma(p, t2, -1.0 / 15.0);
    p = fma(p, t2, 1.0 / 13.0);
    p = fma(p, t2, -1.0 / 11.0);
    p = fma(p, t2, 1.0 / 9.0);
    p = fma(p, t2, -1.0 / 7.0);
    p = fma(p, t2, 1.0 / 5.0);
    p = fma(p, t2, -1.0 / 3.0);
    const double at = fma(t * t2, p, t);
#if defined(__CUDA_ARCH__)
    double a = kAtanHiDev[i] + (at + kAtanLoDev[i]);
#else
    double a = kAtanHiHost[i] + (at + kAtanLoHost[i]);
#endif
    a = ay > ax ? (pio2 - a) + pio2_tail : a;
    a = x < 0.0 ? (pi - a) + pi_tail : a;
    return y < 0.0 ? -a : a;
}

// exp for -700 < x <= 0
F3PS_HD double exp_fast(double x) {
    const double log2e = 0x1.71547652b82fep+0;
    const double ln2_hi = 0x1.62e42fef00000p-1;     // 33-bit head of ln 2
    const double ln2_lo = 0x1.473de6af278edp-34;
    const double kd = rint(x * log2e);
    double r = fma(-kd, ln2_hi, x);
    r = fma(-kd, ln2_lo, r);
    double p = 1.0 / 6227020800.0;                        // 1/13!
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
    return p * pow2i((int)kd);
}

F3PS_HD double pow7(double x) { const double x2 = x * x, x4 = x2 * x2; return (x4 * x2) * x; }

// lab_ciede00, kL = kC = kH = 1 (src/color_utilities.cpp:190-294)
F3PS_HD float ciede00(const float lab1[3], const float lab2[3]) {
    const double PI = 3.14159265358979323846;             // M_PI
    const float L1 = lab1[0], a1 = lab1[1], b1 = lab1[2];
    const float L2 = lab2[0], a2 = lab2[1], b2 = lab2[2];
    const double Cab1 = (double)sqrtf(a1 * a1 + b1 * b1);                                  // :200 (float sqrt)
    const double Cab2 = (double)sqrtf(a2 * a2 + b2 * b2);
    const double Cab = (Cab1 + Cab2) / 2.0;
    const double p25_7 = 6103515625.0;                                                     // pow(25.0, 7.0), exact
    const double Cab7 = pow7(Cab);
    const double G = 0.5 * (1.0 - dsqrt(ddiv(Cab7, Cab7 + p25_7)));
    const double ap1 = (1.0 + G) * (double)a1;
    const double ap2 = (1.0 + G) * (double)a2;
    const double Cp1 = dsqrt(ap1 * ap1 + (double)(b1 * b1));                                // :211 (b*b is a float product)
    const double Cp2 = dsqrt(ap2 * ap2 + (double)(b2 * b2));
    const double Cp_prod = (Cp2 * Cp1);
    const bool z1 = (fabs(ap1) + (double)fabsf(b1)) == 0.0, z2 = (fabs(ap2) + (double)fabsf(b2)) == 0.0;
    double hp1 = atan2_fast((double)b1, z1 ? 1.0 : ap1);
    hp1 = hp1 < 0 ? hp1 + 2.0 * PI : hp1;
    hp1 = z1 ? 0.0 : hp1;
    double hp2 = atan2_fast((double)b2, z2 ? 1.0 : ap2);
    hp2 = hp2 < 0 ? hp2 + 2.0 * PI : hp2;
    hp2 = z2 ? 0.0 : hp2;
    const double dL = (double)(L2 - L1);                                                   // :233 (float subtraction)
    const double dC = (Cp2 - Cp1);
    double dhp = (hp2 - hp1);
    dhp = dhp > PI ? dhp - 2.0 * PI : (dhp < -PI ? dhp + 2.0 * PI : dhp);
    dhp = Cp_prod == 0.0 ? 0.0 : dhp;
    const double dH = 2.0 * dsqrt(Cp_prod) * sin_fast(dhp / 2.0);
    const double Lp = (double)(L2 + L1) / 2.0;                                             // :254 (float addition)
    const double Cp = (Cp1 + Cp2) / 2.0;
    double hp = (hp1 + hp2) / 2.0;
    hp = fabs(hp1 - hp2) > PI ? hp - PI : hp;
    hp = hp < 0 ? hp + 2.0 * PI : hp;
    hp = Cp_prod == 0.0 ? hp1 + hp2 : hp;
    const double Lpm502 = (Lp - 50.0) * (Lp - 50.0);
    const double T = 1.0 - 0.17 * cos_fast(hp - PI / 6.0) + 0.24 * cos_fast(2.0 * hp)
                   + 0.32 * cos_fast(3.0 * hp + PI / 30.0) - 0.20 * cos_fast(4.0 * hp - 63.0 * PI / 180.0);
    const double hq = ddiv(180.0 / PI * hp - 275.0, 25.0);
    const double dheta_rad = (30.0 * PI / 180.0) * exp_fast(-(hq * hq));                   // pow(x, 2.0) == x*x
    const double Cp7 = pow7(Cp);
    const double Rc = 2.0 * dsqrt(ddiv(Cp7, Cp7 + p25_7));
    const double kLSL = (1.0 + ddiv(0.015 * Lpm502, dsqrt(20.0 + Lpm502)));
    const double kLSC = (1.0 + 0.045 * Cp);
    const double kHSH = (1.0 + 0.015 * Cp * T);
    const double RT = -sin_fast(2.0 * dheta_rad) * Rc;
    const double tL = ddiv(dL, kLSL), tC = ddiv(dC, kLSC), tH = ddiv(dH, kHSH);
    return (float)dsqrt(tL * tL + tC * tC + tH * tH + RT * tC * tH);
}

} // namespace f3ps_fastmath
