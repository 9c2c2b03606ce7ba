// Device arithmetic shared by the permutation kernels (conn.cu, conn_tile64.cu): the exact
// quotient, log1p and the float32 node value of /root/reference/scTDA/main.py:982-987.
#pragma once
#include <math.h>
#include <stdint.h>

namespace {

// a / b for finite a, b > 0 given rb = RN(1/b): one Markstein correction step.  With a
// correctly rounded reciprocal and fused multiply-adds this returns the correctly rounded
// quotient RN(a/b), i.e. the same bits as the IEEE division the reference performs, without the
// slow path the generic double division takes for a zero numerator (checked exhaustively for
// the divisors used here by sctda_selftest_div).
__device__ __forceinline__ double div_rn(double a, double b, double rb) {
    double q = a * rb;
    double r = fma(-q, b, a);
    return fma(r, rb, q);
}

// log1p(x) for finite x > 0, error below 1 ulp (the accuracy class of the libm numexpr / numpy call at
// ref:983), without the special-case and slow paths of the generic routine: the node loop calls it
// V times per gene and permutation, and the generic log1p made the node epilogue ~250 instructions.
// Method (the classical one): 1 + x = 2^k * (1 + f) with 1 + f in [sqrt(2)/2, sqrt(2)), the rounding
// error of 1 + x carried as c; s = f / (2 + f), log(1 + f) = f - f^2/2 + s * (f^2/2 + R(s^2)) with a
// degree-7 minimax polynomial R; the quotient by a Newton-refined approximate reciprocal.
// (Measured: -25 % executed instructions, no change in throughput — the kernel is bound by the
// gather, not by issue.)
__device__ __forceinline__ double log1p_pos(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01, Lp3 = 2.857142874366239149e-01,
                 Lp4 = 2.222219843214978396e-01, Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
                 Lp7 = 1.479819860511658591e-01;
    double f = x, c = 0.0;
    int k = 0;
    if (!(x < 0.41421356237309503)) {
        const double u = 1.0 + x;
        int hu = __double2hiint(u);
        k = (hu >> 20) - 1023;
        c = (k > 0) ? 1.0 - (u - x) : x - (u - 1.0);            // rounding error of 1 + x
        double ru;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ru) : "d"(u));   // c / u only needs ~20 bits
        c *= ru;
        hu &= 0x000fffff;
        double un;
        if (hu < 0x6a09e) un = __hiloint2double(hu | 0x3ff00000, __double2loint(u));
        else { k += 1; un = __hiloint2double(hu | 0x3fe00000, __double2loint(u)); }
        f = un - 1.0;
    }
    const double hfsq = 0.5 * f * f;
    const double d = 2.0 + f;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    double s = f * r;
    s = fma(fma(-d, s, f), r, s);                                // s = f / (2 + f)
    const double z = s * s;
    const double R = z * fma(z, fma(z, fma(z, fma(z, fma(z, fma(z, Lp7, Lp6), Lp5), Lp4), Lp3), Lp2), Lp1);
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    const double dk = (double)k;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + (dk * ln2_lo + c))) - f);
}

// float32(log1p(t1) / 0.693147) for t1 >= 0 (ref:983 and the float32 array of ref:972).
__device__ __forceinline__ float pm_log2_value(double t1) {
    const double c = 0.693147;
    const double rc = 1.0 / 0.693147;                 // RN(1/c), folded at compile time
    const bool ok = t1 > 0.0 && t1 < 1e300;
    const float r = (float)div_rn(log1p_pos(ok ? t1 : 1.0), c, rc);
    // t1 == 0 -> 0, NaN -> NaN, +inf -> +inf: all equal to float(t1); nothing else reaches here
    // (a finite sum of 2^x - 1 above 1e300 would need x > 996)
    return ok ? r : (float)t1;
}


}  // namespace
