// dvg_crmath.cuh -- correctly rounded double cos / acos / pow(x, 1./3.) for the few predicate evaluations that sit
// exactly on a decision boundary.
//
// Why: the reference decides "does the +x ray from the sample cross this cubic" from the roots of a cubic in t
// computed with the trigonometric / Cardano closed forms of solve.h:29-59, i.e. through glibc's double acos, cos
// and pow, and then tests `t >= 0 && t <= 1`.  When a sample's y coordinate EQUALS a segment end point's y as a
// float (it happens: ~1 sample per million on the bundled SVG assets, whose coordinates carry three decimals) the
// constant term of the cubic is exactly 0, one root is mathematically 0, and the closed form returns 0.0 or
// +-1e-16 depending on the last bit of acos and cos: the sample flips between "crossing counted" and "not
// counted" with the maths library.  glibc 2.39's acos / cos / pow are correctly rounded for 99.9 % of their
// arguments (measured against 200-bit arithmetic: 0.07 % / 0.14 % / 0.08 % misrounded); CUDA's are within 1-2
// ulp, i.e. often a bit off.  Reproducing the reference's verdict at those samples therefore needs correctly
// rounded functions on the device, but only there: dvg_geom.cuh calls the fast functions first and re-solves with
// these when a root lands within 1e-9 of a decision boundary.
//
// Method: double-double arithmetic (error-free two_sum / two_prod with explicit fma; ~106 bits).
//   cos   : reduction by multiples of pi/2 with a three-term pi/2, Taylor series of sin / cos on |r| <= pi/4
//   acos  : one Newton step on cos(y) = x from the library value, residual evaluated in double-double
//   pow13 : x^e with e = (double)(1./3.): cube root by one Newton step in double-double, times
//           1 + (e - 1/3) ln x   (e - 1/3 = -2^-54 / 3 ... the first-order term is ~1e-17, its square is far
//           below the 106-bit working precision)
// All three return the double nearest to the ~1e-31-accurate result.
#pragma once
#include "dvg_common.cuh"

namespace dvg {

struct DD { double hi, lo; };

#if defined(__CUDA_ARCH__)
#define DVG_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define DVG_FMA(a, b, c) fma((a), (b), (c))
#endif

DVG_HD DD dd_mk(double h, double l) { DD r; r.hi = h; r.lo = l; return r; }
DVG_HD DD dd_two_sum(double a, double b) {
    const double s = a + b, bb = s - a;
    return dd_mk(s, (a - (s - bb)) + (b - bb));
}
DVG_HD DD dd_fast_two_sum(double a, double b) {   // |a| >= |b|
    const double s = a + b;
    return dd_mk(s, b - (s - a));
}
DVG_HD DD dd_two_prod(double a, double b) {
    const double p = a * b;
    return dd_mk(p, DVG_FMA(a, b, -p));
}
DVG_HD DD dd_add(DD a, DD b) {
    DD s = dd_two_sum(a.hi, b.hi);
    const DD t = dd_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = dd_fast_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return dd_fast_two_sum(s.hi, s.lo);
}
DVG_HD DD dd_add_d(DD a, double b) {
    DD s = dd_two_sum(a.hi, b);
    s.lo += a.lo;
    return dd_fast_two_sum(s.hi, s.lo);
}
DVG_HD DD dd_neg(DD a) { return dd_mk(-a.hi, -a.lo); }
DVG_HD DD dd_mul(DD a, DD b) {
    DD p = dd_two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return dd_fast_two_sum(p.hi, p.lo);
}
DVG_HD DD dd_mul_d(DD a, double b) {
    DD p = dd_two_prod(a.hi, b);
    p.lo += a.lo * b;
    return dd_fast_two_sum(p.hi, p.lo);
}
DVG_HD DD dd_div_d(DD a, double b) {   // a / b, b a double
    const double q1 = a.hi / b;
    const DD p = dd_two_prod(q1, b);
    const DD r = dd_add(a, dd_neg(p));
    const double q2 = r.hi / b;
    return dd_fast_two_sum(q1, q2);
}
DVG_HD DD dd_div(DD a, DD b) {
    const double q1 = a.hi / b.hi;
    DD r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
    const double q2 = r.hi / b.hi;
    r = dd_add(r, dd_neg(dd_mul_d(b, q2)));
    const double q3 = r.hi / b.hi;
    return dd_add_d(dd_fast_two_sum(q1, q2), q3);
}

// sin and cos of a double-double r, |r| <= ~0.8, by their Taylor series (terms below 1e-33 dropped)
DVG_HD_NOINLINE void dd_sincos_small(DD r, DD *s, DD *c) {
    const DD r2 = dd_mul(r, r);
    DD ts = r, tc = dd_mk(1.0, 0.0);   // current terms r^(2n+1)/(2n+1)!, r^(2n)/(2n)!
    DD ss = r, cs = tc;
    for (int n = 1; n <= 15; n++) {
        tc = dd_div_d(dd_mul(tc, r2), (double)((2 * n - 1) * (2 * n)));
        ts = dd_div_d(dd_mul(ts, r2), (double)((2 * n) * (2 * n + 1)));
        if (n & 1) { cs = dd_add(cs, dd_neg(tc)); ss = dd_add(ss, dd_neg(ts)); }
        else { cs = dd_add(cs, tc); ss = dd_add(ss, ts); }
    }
    *s = ss; *c = cs;
}

// sin / cos of a double-double x, |x| < ~1e4
DVG_HD_NOINLINE void dd_sincos(DD x, DD *s, DD *c) {
    // pi/2 = P1 + P2 + P3 (three doubles, ~160 bits)
    const double P1 = 1.5707963267948966, P2 = 6.123233995736766e-17, P3 = -1.4973849048591698e-33;
    const double k = rint(x.hi * 0.6366197723675814);
    DD r = x;
    if (k != 0.0) {
        r = dd_add(r, dd_neg(dd_two_prod(k, P1)));
        r = dd_add(r, dd_neg(dd_two_prod(k, P2)));
        r = dd_add_d(r, -k * P3);
    }
    DD sr, cr;
    dd_sincos_small(r, &sr, &cr);
    const int q = (int)(((long long)k) & 3);
    if (q == 0) { *s = sr; *c = cr; }
    else if (q == 1) { *s = cr; *c = dd_neg(sr); }
    else if (q == 2) { *s = dd_neg(sr); *c = dd_neg(cr); }
    else { *s = dd_neg(cr); *c = sr; }
}

// correctly rounded cos(x), x a double of moderate size
DVG_HD_NOINLINE double cr_cos(double x) {
    if (!(fabs(x) < 1e4)) return cos(x);
    DD s, c;
    dd_sincos(dd_mk(x, 0.0), &s, &c);
    return c.hi + c.lo;
}

// correctly rounded acos(x), |x| < 1 (otherwise the library's exact / NaN cases)
DVG_HD_NOINLINE double cr_acos(double x) {
    if (!(fabs(x) < 1.0)) return acos(x);
    const double y0 = acos(x);
    DD s, c;
    dd_sincos(dd_mk(y0, 0.0), &s, &c);
    // cos(y) = x  =>  y1 = y0 + (cos(y0) - x) / sin(y0)
    const DD num = dd_add_d(c, -x);
    const DD corr = dd_div(num, s);
    const DD y = dd_add_d(corr, y0);
    return y.hi + y.lo;
}

// correctly rounded pow(x, 1./3.) for finite x > 0 (note: the exponent is the DOUBLE nearest to one third)
DVG_HD_NOINLINE double cr_pow13(double x) {
    if (!(x > 0.0) || !(x < 1e300) || x < 1e-300) return pow(x, 1. / 3.);
    const double c0 = cbrt(x);
    // Newton on c^3 = x in double-double: c1 = c0 - (c0^3 - x) / (3 c0^2)
    const DD c0d = dd_mk(c0, 0.0);
    const DD cube = dd_mul(dd_mul(c0d, c0d), c0d);
    const DD res = dd_add_d(cube, -x);
    const DD corr = dd_div_d(res, 3.0 * c0 * c0);
    DD c1 = dd_add(c0d, dd_neg(corr));
    // x^e = x^(1/3) * exp((e - 1/3) ln x),  e - 1/3 = -2^-54 / 3
    const double delta = -1.8503717077085941e-17;
    c1 = dd_add(c1, dd_mul_d(c1, delta * log(x)));
    return c1.hi + c1.lo;
}

}  // namespace dvg
