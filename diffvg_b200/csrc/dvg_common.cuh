// dvg_common.cuh -- scalar/vector math, RNG, polynomial solvers, pixel filters.
//
// Everything here is `__host__ __device__` so that the arithmetic can also be exercised by
// a host-compiled test harness (tests/host_emul/), but the product only ever runs it on
// the GPU.  ARITHMETIC CONTRACT: geometric predicates must reproduce the reference's
// float/double mix and operation order exactly; this translation unit is
// therefore compiled with -fmad=false (nvcc) / -ffp-contract=off (g++), and every place
// where the reference silently promotes to double (unqualified ::sqrt/::cos/::acos/::pow
// on float arguments resolve to the double overloads in the reference build) is written
// out explicitly here.
#pragma once
#include <string.h>
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define DVG_HD __host__ __device__ __forceinline__
#define DVG_HD_NOINLINE static __host__ __device__ __noinline__
#define DVG_D __device__ __forceinline__
#define DVG_D_NOINLINE __device__ __noinline__
#else
#define DVG_HD inline
#define DVG_HD_NOINLINE static inline
#define DVG_D inline
#define DVG_D_NOINLINE inline
#endif

namespace dvg {

struct F2 { float x, y; };
struct alignas(16) U4 { unsigned x, y, z, w; };   // 16-byte record word (one 128-bit load / store on the device)
struct F4 { float x, y, z, w; };

DVG_HD unsigned dvg_float_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    unsigned u; memcpy(&u, &f, sizeof u); return u;
#endif
}
DVG_HD float dvg_bits_float(unsigned u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, sizeof f); return f;
#endif
}
DVG_HD F2 mk2(float x, float y) { F2 r; r.x = x; r.y = y; return r; }
DVG_HD F4 mk4(float x, float y, float z, float w) { F4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
DVG_HD F2 operator+(F2 a, F2 b) { return mk2(a.x + b.x, a.y + b.y); }
DVG_HD F2 operator-(F2 a, F2 b) { return mk2(a.x - b.x, a.y - b.y); }
DVG_HD F2 operator-(F2 a) { return mk2(-a.x, -a.y); }
DVG_HD F2 operator*(float s, F2 a) { return mk2(s * a.x, s * a.y); }
DVG_HD F2 operator*(F2 a, float s) { return mk2(a.x * s, a.y * s); }
DVG_HD F2 operator*(F2 a, F2 b) { return mk2(a.x * b.x, a.y * b.y); }
// vector.h:379-385: vector / scalar multiplies by the reciprocal 1.f / s (NOT a true division)
DVG_HD F2 operator/(F2 a, float s) { float inv_s = 1.f / s; return mk2(a.x * inv_s, a.y * inv_s); }
DVG_HD float sum2(F2 a) { return a.x + a.y; }
DVG_HD float dot2(F2 a, F2 b) { return a.x * b.x + a.y * b.y; }
DVG_HD F4 operator+(F4 a, F4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
DVG_HD F4 operator-(F4 a, F4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
DVG_HD F4 operator*(F4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
DVG_HD F4 operator*(float s, F4 a) { return mk4(s * a.x, s * a.y, s * a.z, s * a.w); }
DVG_HD F4 operator*(F4 a, F4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
DVG_HD float sum4(F4 a) { return a.x + a.y + a.z + a.w; }  // vector.h:690-693 (left to right)

// reference diffvg.h:62-72: plain comparisons, NOT fmin/fmax (NaN behaviour differs)
DVG_HD float rmaxf(float a, float b) { return a > b ? a : b; }
DVG_HD float rminf(float a, float b) { return a < b ? a : b; }
DVG_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
DVG_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// vector.h:478-481, 535-538: length_squared(v1 - v0)
DVG_HD float dist_sq(F2 v0, F2 v1) {
    float dx = v1.x - v0.x, dy = v1.y - v0.y;
    return dx * dx + dy * dy;
}
// vector.h:491-494: sqrt resolves to ::sqrt(double) and the result is rounded back to
// float; for sqrt this double rounding is innocuous, so IEEE sqrtf is bit-identical.
DVG_HD float length2(F2 v) { return sqrtf(v.x * v.x + v.y * v.y); }
DVG_HD float distance2(F2 v0, F2 v1) { return sqrtf(dist_sq(v0, v1)); }
DVG_HD F2 normalize2(F2 v) { return v / length2(v); }

// ---------------------------------------------------------------- PCG32 (pcg.h:11-40)
struct Pcg32 { uint64_t state, inc; };

DVG_HD uint32_t pcg32_next(Pcg32 &rng) {
    uint64_t oldstate = rng.state;
    rng.state = oldstate * 6364136223846793005ULL + (rng.inc | 1);
    uint32_t xorshifted = (uint32_t)(((oldstate >> 18u) ^ oldstate) >> 27u);
    uint32_t rot = (uint32_t)(oldstate >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31));
}
DVG_HD float pcg32_next_float(Pcg32 &rng) {
    uint32_t u = (pcg32_next(rng) >> 9) | 0x3f800000u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u) - 1.0f;
#else
    union { uint32_t u; float f; } x; x.u = u; return x.f - 1.0f;
#endif
}
DVG_HD Pcg32 pcg32_init(int idx, uint64_t seed) {
    Pcg32 s;
    s.state = 0U;
    s.inc = (((uint64_t)idx + 1) << 1u) | 1u;
    pcg32_next(s);
    s.state += (0x853c49e6748fea9bULL + seed);
    pcg32_next(s);
    return s;
}

// ---------------------------------------------------------------- 3x3 transforms (matrix.h)
struct M3 { float m[9]; };  // row-major

// matrix.h:505-512
DVG_HD F2 xform_pt(const float *m, F2 p) {
    float t0 = m[0] * p.x + m[1] * p.y + m[2];
    float t1 = m[3] * p.x + m[4] * p.y + m[5];
    float t2 = m[6] * p.x + m[7] * p.y + m[8];
    return mk2(t0 / t2, t1 / t2);
}
// matrix.h:538-543
DVG_HD F2 xform_normal(const float *minv, F2 n) {
    return normalize2(mk2(minv[0] * n.x + minv[3] * n.y, minv[1] * n.x + minv[4] * n.y));
}
// matrix.h:209-231
DVG_HD void inverse3(const float *m, float *o) {
    float det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                m[2] * (m[3] * m[7] - m[4] * m[6]);
    float invdet = 1 / det;
    o[0] = (m[4] * m[8] - m[7] * m[5]) * invdet;
    o[1] = (m[2] * m[7] - m[1] * m[8]) * invdet;
    o[2] = (m[1] * m[5] - m[2] * m[4]) * invdet;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * invdet;
    o[4] = (m[0] * m[8] - m[2] * m[6]) * invdet;
    o[5] = (m[3] * m[2] - m[0] * m[5]) * invdet;
    o[6] = (m[3] * m[7] - m[6] * m[4]) * invdet;
    o[7] = (m[6] * m[1] - m[0] * m[7]) * invdet;
    o[8] = (m[0] * m[4] - m[3] * m[1]) * invdet;
}
// matrix.h:514-536: adjoint of xform_pt; d_m (9 floats) and d_pt are accumulated into.
DVG_HD void d_xform_pt(const float *m, F2 pt, F2 d_out, float *d_m, F2 &d_pt) {
    float t0 = m[0] * pt.x + m[1] * pt.y + m[2];
    float t1 = m[3] * pt.x + m[4] * pt.y + m[5];
    float t2 = m[6] * pt.x + m[7] * pt.y + m[8];
    float o0 = t0 / t2, o1 = t1 / t2;
    float dt0 = d_out.x / t2, dt1 = d_out.y / t2;
    float dt2 = -(d_out.x * o0 + d_out.y * o1) / t2;
    d_m[0] += dt0 * pt.x; d_m[1] += dt0 * pt.y; d_m[2] += dt0;
    d_m[3] += dt1 * pt.x; d_m[4] += dt1 * pt.y; d_m[5] += dt1;
    d_m[6] += dt2 * pt.x; d_m[7] += dt2 * pt.y; d_m[8] += dt2;
    d_pt.x += dt0 * m[0] + dt1 * m[3] + dt2 * m[6];
    d_pt.y += dt0 * m[1] + dt1 * m[4] + dt2 * m[7];
}

// ---------------------------------------------------------------- boxes (aabb.h)
struct Box { float x0, y0, x1, y1; };
DVG_HD bool box_inside(Box b, F2 p) {  // aabb.h:31-35
    return p.x >= b.x0 && p.x <= b.x1 && p.y >= b.y0 && p.y <= b.y1;
}
DVG_HD bool box_inside_r(Box b, F2 p, float r) {  // aabb.h:38-42 and 63-67 (identical predicates)
    return p.x >= b.x0 - r && p.x <= b.x1 + r && p.y >= b.y0 - r && p.y <= b.y1 + r;
}
DVG_HD bool box_ray_intersect(Box b, F2 p) {  // winding_number.h:33-42
    if (p.y < b.y0 || p.y > b.y1) return false;
    if (p.x > b.x1) return false;
    return true;
}

// ---------------------------------------------------------------- solvers (solve.h)
// solve.h:5-27, T = float.  sqrt(float) -> correctly rounded either way.
DVG_HD bool solve_quadratic_f(float a, float b, float c, float *t0, float *t1) {
    float discrim = b * b - 4 * a * c;
    if (discrim < 0) return false;
    float root_discrim = sqrtf(discrim);
    float q;
    if (b < 0) q = -0.5f * (b - root_discrim);
    else q = -0.5f * (b + root_discrim);
    *t0 = q / a;
    *t1 = c / q;
    if (*t0 > *t1) { float tmp = *t0; *t0 = *t1; *t1 = tmp; }
    return true;
}
DVG_HD_NOINLINE bool solve_quadratic_d(double a, double b, double c, double *t0, double *t1) {
    double discrim = b * b - 4 * a * c;
    if (discrim < 0) return false;
    double root_discrim = sqrt(discrim);
    double q;
    if (b < 0) q = -0.5f * (b - root_discrim);
    else q = -0.5f * (b + root_discrim);
    *t0 = q / a;
    *t1 = c / q;
    if (*t0 > *t1) { double tmp = *t0; *t0 = *t1; *t1 = tmp; }
    return true;
}

#define DVG_PI_D 3.14159265358979323846

// solve.h:29-59 with T = float.  The reference calls the *double* ::sqrt/::acos/::cos/::pow
// on float arguments and only rounds when storing to a float; the promotions are explicit here.
DVG_HD_NOINLINE int solve_cubic_f(float a, float b, float c, float d, float t[3]) {
    if (fabsf(a) < 1e-6f) {
        if (solve_quadratic_f(b, c, d, &t[0], &t[1])) return 2;
        return 0;
    }
    b /= a; c /= a; d /= a;
    float Q = (b * b - 3 * c) / 9.f;
    float R = (2 * b * b * b - 9 * b * c + 27 * d) / 54.f;
    if (R * R < Q * Q * Q) {
        float theta = (float)acos((double)R / sqrt((double)(Q * Q * Q)));
        double m2sq = (double)(-2.f) * sqrt((double)Q);
        float pi_f = (float)DVG_PI_D;
        t[0] = (float)(m2sq * cos((double)(theta / 3.f)) - (double)(b / 3.f));
        t[1] = (float)(m2sq * cos((double)((theta + 2.f * pi_f) / 3.f)) - (double)(b / 3.f));
        t[2] = (float)(m2sq * cos((double)((theta - 2.f * pi_f) / 3.f)) - (double)(b / 3.f));
        return 3;
    } else {
        double third = (double)(float)(1. / 3.);
        double sq = sqrt((double)(R * R - Q * Q * Q));
        float A = R > 0 ? (float)(-pow((double)R + sq, third)) : (float)pow((double)(-R) + sq, third);
        float B = fabsf(A) > 1e-6f ? Q / A : 0.f;
        t[0] = (A + B) - b / 3.f;
        return 1;
    }
}
// solve.h:29-59 with T = double.  FAST = one division + three multiplies for the normalisation:
// only for the isolator polynomial of the closest-point quintic, whose roots are rounded to float
// bracket ends (dvg_geom.cuh quintic_eval note); the winding test compares the double roots with
// 0 and 1 directly and must keep the reference's exact operation sequence.
template <bool FAST>
DVG_HD_NOINLINE int solve_cubic_dt(double a, double b, double c, double d, double t[3]) {
    if (fabs(a) < 1e-6f) {
        if (solve_quadratic_d(b, c, d, &t[0], &t[1])) return 2;
        return 0;
    }
    if (FAST) { const double inv_a = 1.0 / a; b *= inv_a; c *= inv_a; d *= inv_a; }
    else { b /= a; c /= a; d /= a; }
    double Q = (b * b - 3 * c) / 9.f;
    double R = (2 * b * b * b - 9 * b * c + 27 * d) / 54.f;
    if (R * R < Q * Q * Q) {
        double theta = acos(R / sqrt(Q * Q * Q));
        t[0] = -2.f * sqrt(Q) * cos(theta / 3.f) - b / 3.f;
        t[1] = -2.f * sqrt(Q) * cos((theta + 2.f * DVG_PI_D) / 3.f) - b / 3.f;
        t[2] = -2.f * sqrt(Q) * cos((theta - 2.f * DVG_PI_D) / 3.f) - b / 3.f;
        return 3;
    } else {
        double A = R > 0 ? -pow(R + sqrt(R * R - Q * Q * Q), 1. / 3.) : pow(-R + sqrt(R * R - Q * Q * Q), 1. / 3.);
        double B = fabs(A) > 1e-6f ? Q / A : 0.0;
        t[0] = (A + B) - b / 3.0;
        return 1;
    }
}
DVG_HD int solve_cubic_d(double a, double b, double c, double d, double t[3]) { return solve_cubic_dt<false>(a, b, c, d, t); }

// ---------------------------------------------------------------- pixel filters (filter.h)
struct Filter { int type; float radius; };

// filter.h:22-48.  cos(float) is the double ::cos in the reference build.
DVG_HD float filter_weight(Filter f, float dx, float dy) {
    if (fabsf(dx) > f.radius || fabsf(dy) > f.radius) return 0;
    if (f.type == 0) {
        float w = 2 * f.radius;
        return 1.f / (w * w);
    } else if (f.type == 1) {
        float r2 = f.radius * f.radius;
        return (f.radius - fabsf(dx)) * (f.radius - fabsf(dy)) / (r2 * r2);
    } else if (f.type == 2) {
        float sx = dx / f.radius, sy = dy / f.radius;
        return (4.f / 3.f) * (1 - sx * sx) * (4.f / 3.f) * (1 - sy * sy);
    } else {
        float ndx = (dx / (2 * f.radius)) + 0.5f;
        float ndy = (dy / (2 * f.radius)) + 0.5f;
        float two_pi = (float)(2 * DVG_PI_D);
        double a = (double)0.5f * ((double)1.f - cos((double)(two_pi * ndx)));
        double b = a * (double)0.5f * ((double)1.f - cos((double)(two_pi * ndy)));
        return (float)(b / (double)(f.radius * f.radius));
    }
}

// filter.h:50-106: returns the value the reference atomically adds to d_filter.radius.
DVG_HD_NOINLINE float d_filter_weight_radius(Filter f, float dx, float dy, float d_return) {
    float r = f.radius;
    if (f.type == 0) {
        float w = 2 * r;
        return d_return * (-2) * 2 * r / (w * w * w);
    } else if (f.type == 1) {
        float fx = r - fabsf(dx), fy = r - fabsf(dy);
        float norm = 1 / (r * r);
        float d_fx = d_return * fy * norm, d_fy = d_return * fx * norm;
        float d_norm = d_return * fx * fy;
        return (float)((double)(d_fx + d_fy) + (double)((-4) * d_norm) / pow((double)r, 5.0));
    } else if (f.type == 2) {
        float r3 = r * r * r;
        return -(2 * dx * dx + 2 * dy * dy) / r3;
    } else {
        float two_pi = (float)(2 * DVG_PI_D);
        float ndx = (dx / (2 * r)) + 0.5f, ndy = (dy / (2 * r)) + 0.5f;
        float fx = (float)((double)0.5f * ((double)1.f - cos((double)(two_pi * ndx))));
        float fy = (float)((double)0.5f * ((double)1.f - cos((double)(two_pi * ndy))));
        float norm = 1 / (r * r);
        float d_fx = d_return * fy * norm, d_fy = d_return * fx * norm;
        float d_norm = d_return * fx * fy;
        float d_ndx = (float)((double)(d_fx * 0.5f) * sin((double)(two_pi * ndx)) * (double)two_pi);
        float d_ndy = (float)((double)(d_fy * 0.5f) * sin((double)(two_pi * ndy)) * (double)two_pi);
        float w2 = (2 * r) * (2 * r);
        return d_ndx * (-2 * dx / w2) + d_ndy * (-2 * dy / w2) + (-2) * d_norm / (r * r * r);
    }
}

// diffvg.cpp:817-833
DVG_HD float smoothstep(float d) {
    float t = clampf((d + 1.f) / 2.f, 0.f, 1.f);
    return t * t * (3 - 2 * t);
}
DVG_HD float d_smoothstep(float d, float d_ret) {
    if (d < -1.f || d > 1.f) return 0.f;
    float t = (d + 1.f) / 2.f;
    float d_t = d_ret * (6 * t - 6 * t * t);
    return d_t / 2.f;
}

// diffvg.h:110-126 (10-bit interleave)
DVG_HD uint32_t expand_bits(uint32_t x) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < 10; i++) r |= (x & (1u << i)) << i;
    return r;
}

}  // namespace dvg
