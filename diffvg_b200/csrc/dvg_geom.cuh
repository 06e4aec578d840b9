// dvg_geom.cuh -- per-primitive geometric predicates: "is pt within stroke radius of this
// segment" (within_distance.h) and "signed crossings of the +x ray from pt with this segment"
// (winding_number.h).  Arithmetic follows the reference expression by expression (types,
// operation order, float literals inside double expressions) because one flipped sample
// changes a pixel by 1/spp.
#pragma once
#include "dvg_scene.cuh"
#include "dvg_crmath.cuh"

namespace dvg {

#ifdef DVG_FMA_QUINTIC
#ifndef DVG_FQ_SELECT
#define DVG_FQ_HORNER
#define DVG_FQ_RECIP
#define DVG_FQ_NEWTON
#define DVG_FQ_ISOL
#endif
#endif

DVG_HD F2 eval_quad(F2 p0, F2 p1, F2 p2, float t) {  // within_distance.h:75-78
    float tt = 1 - t;
    return (tt * tt) * p0 + (2 * tt * t) * p1 + (t * t) * p2;
}
DVG_HD F2 eval_cubic(F2 p0, F2 p1, F2 p2, F2 p3, float t) {  // within_distance.h:129-132
    float tt = 1 - t;
    return (tt * tt * tt) * p0 + (3 * tt * tt * t) * p1 + (3 * tt * t * t) * p2 + (t * t * t) * p3;
}

// vector.h:787-817
DVG_HD F2 quadratic_closest_pt_approx(F2 b0, F2 b1, F2 b2, F2 pt, float *t_out) {
    b0 = b0 - pt; b1 = b1 - pt; b2 = b2 - pt;
#define DVG_DET(u, v) ((u).x * (v).y - (u).y * (v).x)
    float a = DVG_DET(b0, b2), b = 2 * DVG_DET(b1, b0), d = 2 * DVG_DET(b2, b1);
    float f = b * d - a * a;
    F2 d21 = b2 - b1, d10 = b1 - b0, d20 = b2 - b0;
    F2 gf = 2 * (b * d21 + d * d10 + a * d20);
    gf = mk2(gf.y, -gf.x);
    F2 pp = (-f * gf) / dot2(gf, gf);
    F2 d0p = b0 - pp;
    float ap = DVG_DET(d0p, d20), bp = 2 * DVG_DET(d10, d0p);
#undef DVG_DET
    float t = clampf((ap + bp) / (2 * a + b + d), 0.f, 1.f);
    float tt = 1 - t;
    if (t_out) *t_out = t;
    return ((tt * tt) * b0 + (2 * tt * t) * b1 + (t * t) * b2) + pt;
}

// The normalised quintic whose roots are the stationary points of |q(t) - pt|^2 for a cubic
// (within_distance.h:161-172).  Coefficients are formed in float and only then widened.
struct Quintic { double B, C, D, E, F; };

DVG_HD Quintic cubic_quintic(F2 p0, F2 p1, F2 p2, F2 p3, F2 pt) {
    F2 q3 = -p0 + 3 * p1 - 3 * p2 + p3;
    F2 q2 = 3 * p0 - 6 * p1 + 3 * p2;
    F2 q1 = -3 * p0 + 3 * p1;
    F2 pp = p0 - pt;
    double A = 3 * sum2(q3 * q3);
    double B = 5 * sum2(q3 * q2);
    double C = 4 * sum2(q3 * q1) + 2 * sum2(q2 * q2);
    double D = 3 * (sum2(q2 * q1) + sum2(q3 * pp));
    double E = sum2(q1 * q1) + 2 * sum2(pp * q2);
    double F = sum2(pp * q1);
    Quintic q;
#ifdef DVG_FQ_RECIP
    // one division + five multiplies instead of five divisions (within_distance.h:168-172); the
    // quotients differ from the reference's by <= 1 ulp of a double, see the note at quintic_eval
    const double inv_A = 1.0 / A;
    q.B = B * inv_A; q.C = C * inv_A; q.D = D * inv_A; q.E = E * inv_A; q.F = F * inv_A;
#else
    q.B = B / A; q.C = C / A; q.D = D / A; q.E = E / A; q.F = F / A;
#endif
    return q;
}
// within_distance.h:211-225 evaluate the monic quintic and its derivative term by term in double
// (19 + 17 rounded operations).  Here: Horner with fused multiply-adds (5 + 4 DFMA).  The two
// differ by ~1e-16 relative, and every consumer rounds to float (the Newton iterate) or compares
// against 1e-5 / 0, so the float iterates -- and hence the classification -- are identical except
// when a double lands within ~1e-9 relative of a float rounding boundary (DESIGN.md "arithmetic
// contract"; checked bit-for-bit against the reference on the full-size configs).
#ifdef DVG_FQ_HORNER
DVG_HD double quintic_eval(const Quintic &q, double t) {
    return fma(fma(fma(fma(t + q.B, t, q.C), t, q.D), t, q.E), t, q.F);
}
DVG_HD double quintic_deriv(const Quintic &q, double t) {
    return fma(fma(fma(fma(5.0, t, 4.0 * q.B), t, 3.0 * q.C), t, 2.0 * q.D), t, q.E);
}
#else
DVG_HD double quintic_eval(const Quintic &q, double t) {  // within_distance.h:211-218
    return t * t * t * t * t + q.B * t * t * t * t + q.C * t * t * t + q.D * t * t + q.E * t + q.F;
}
DVG_HD double quintic_deriv(const Quintic &q, double t) {  // within_distance.h:219-225
    return 5 * t * t * t * t + 4 * q.B * t * t * t + 3 * q.C * t * t + 2 * q.D * t + q.E;
}
#endif
// value / derivative of the Newton update (within_distance.h:261).  The quotient only feeds a
// float (the next iterate), so ~45 correct bits are as good as 53: reciprocal seed in float, one
// Newton step in double.  Falls back to the IEEE division outside the float range.
DVG_HD_NOINLINE double ieee_quotient(double value, double derivative) { return value / derivative; }
// float reciprocal used as a SEED only (one MUFU.RCP on the device; the callers keep |x| within [1e-30, 1e30])
DVG_HD float seed_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
DVG_HD double newton_quotient(double value, double derivative) {
#if defined(DVG_FQ_NEWTON)
    const float df = (float)derivative;
    if (fabsf(df) > 1e-30f && fabsf(df) < 1e30f) {
        const double r0 = (double)seed_rcp(df);
        const double r1 = fma(r0, fma(-derivative, r0, 1.0), r0);
        return value * r1;
    }
#endif
    return ieee_quotient(value, derivative);   // out of line: ~350 instructions, almost never taken
}

// Roots of the isolator cubic (within_distance.h:194, solve.h:29-59 with T = double) for use as
// FLOAT bracket ends.  The reference evaluates the trigonometric / Cardano closed forms in double
// (acos, 3 x cos or pow: several hundred FP64 instructions on the GPU) and then rounds each root to
// float.  Here: the same branch decisions in double, a float-precision closed-form estimate, and two
// Newton steps on the normalised cubic in double.  Both land within ~1e-12 (relative to the root
// scale) of the true root, so the rounded floats agree except when the value sits that close to a
// float rounding boundary (~2e-5 of the roots); such a 1-ulp shift of a bracket end changes the
// final classification only if additionally a quintic root falls inside that 1-ulp gap or the Newton
// iterate path differs at a sample within an ulp of the stroke edge (~1e-7 each): DESIGN.md
// "arithmetic contract" gives the measured agreement (0 of 2e8 tests).
#ifndef DVG_POLISH_STEPS
#define DVG_POLISH_STEPS 2
#endif
DVG_HD double cubic_polish(double b, double c, double d, double x) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int it = 0; it < DVG_POLISH_STEPS; it++) {
        const double f = fma(fma(fma(x, 1.0, b), x, c), x, d);
        const double fp = fma(fma(3.0, x, 2.0 * b), x, c);
#if defined(DVG_FQ_NEWTON)
        // the correction f / fp is ~1e-6 |x| in the first step and ~1e-12 |x| in the second: a quotient good to float
        // precision moves x by less than a double ulp of what the exact quotient would (f itself is formed in double)
        const float fpf = (float)fp, ff = (float)f;
        if (fabsf(fpf) > 1e-30f && fabsf(fpf) < 1e30f && fabsf(ff) < 1e30f) { x -= (double)(ff * seed_rcp(fpf)); continue; }
#endif
        if (fp != 0.0) x -= newton_quotient(f, fp);
    }
    return x;
}
DVG_HD int isolator_roots(double a, double b, double c, double d, double t[3]) {
    if (fabs(a) < 1e-6f) {
        if (solve_quadratic_d(b, c, d, &t[0], &t[1])) return 2;
        return 0;
    }
    { const double inv_a = 1.0 / a; b *= inv_a; c *= inv_a; d *= inv_a; }
    const double Q = (b * b - 3 * c) / 9.f;
    const double R = (2 * b * b * b - 9 * b * c + 27 * d) / 54.f;
    const double Q3 = Q * Q * Q;
    const double b3 = b / 3.0;
    if (R * R < Q3) {
        const float sq = sqrtf((float)Q);
        float x = (float)R / (sq * sq * sq);
        x = x < -1.f ? -1.f : (x > 1.f ? 1.f : x);
        const float theta = acosf(x);
        const float two_pi = 6.28318530717958647692f;
        const double m2sq = -2.0 * (double)sq;
        t[0] = cubic_polish(b, c, d, m2sq * (double)cosf(theta / 3.f) - b3);
        t[1] = cubic_polish(b, c, d, m2sq * (double)cosf((theta + two_pi) / 3.f) - b3);
        t[2] = cubic_polish(b, c, d, m2sq * (double)cosf((theta - two_pi) / 3.f) - b3);
        return 3;
    } else {
        const float s = (float)sqrt(R * R - Q3);
        const float Af = R > 0 ? -cbrtf((float)R + s) : cbrtf((float)(-R) + s);
        const float Bf = fabsf(Af) > 1e-6f ? (float)Q / Af : 0.f;
        t[0] = cubic_polish(b, c, d, (double)(Af + Bf) - b3);
        return 1;
    }
}

// Isolator-polynomial split points (within_distance.h:184-210).  Returns the sorted interval
// ends.  Q10: when q_root is outside [0,1] the reference reads intervals[0]
// uninitialised; we then use -1 ("no split point": negative entries are skipped).
// EXACT_ISOLATOR: the distance queries (dvg_distance.cuh) feed the root parameter into a continuous
// output, so a quintic root the reference finds (or misses) because of where its isolator split points
// fall must be found (or missed) here too; they evaluate the isolator cubic with the reference's own
// closed form.  The stroke classification only needs hit / no hit and uses the fast roots.
template <bool EXACT_ISOLATOR = false>
DVG_HD int quintic_intervals(const Quintic &q, float intervals[4], float stale0 = -1.f) {
    double p1A = ((2 / 5.f) * q.C - (4 / 25.f) * q.B * q.B);
    double p1B = ((3 / 5.f) * q.D - (3 / 25.f) * q.B * q.C);
    double p1C = ((4 / 5.f) * q.E - (2 / 25.f) * q.B * q.D);
    double p1D = q.F - q.B * q.E / 25.f;
    double q_root = -q.B / 5.f;
    double p_roots[3];
#ifdef DVG_FQ_ISOL
    int num_sol = EXACT_ISOLATOR ? solve_cubic_d(p1A, p1B, p1C, p1D, p_roots) : isolator_roots(p1A, p1B, p1C, p1D, p_roots);
#else
    int num_sol = solve_cubic_d(p1A, p1B, p1C, p1D, p_roots);
#endif
    intervals[0] = stale0;
    if (q_root >= 0 && q_root <= 1) intervals[0] = (float)q_root;
    for (int j = 0; j < num_sol; j++) intervals[j + 1] = (float)p_roots[j];
    int n = 1 + num_sol;
    for (int j = 1; j < n; j++) {
        for (int k = j; k > 0 && intervals[k - 1] > intervals[k]; k--) {
            float tmp = intervals[k]; intervals[k] = intervals[k - 1]; intervals[k - 1] = tmp;
        }
    }
    return n;
}
// ---- the same quintic and split points, with everything that does not depend on the sample formed ONCE PER PRIMITIVE.
// Of cubic_quintic's inputs only pp = p0 - pt changes from sample to sample; it enters D, E and F through three float dot
// products.  A, B, C, the float vectors q1 q2 q3, the float sums sum(q2 q1) and sum(q1 q1), the isolator's leading
// coefficient p1A, both reciprocals (1 / A, 1 / p1A) and the split point -B / 5 are the primitive's.  Every float
// expression below is the one cubic_quintic evaluates, in the same order, so D, E, F -- and with DVG_FQ_RECIP the
// normalised coefficients -- are bit-identical to cubic_quintic's (tests/test_cpu_oracle_and_host.py checks that on the
// host build of this header).  The exact-test kernel (dvg_wave.cu) reads this 80-byte record instead of redoing ~100
// float operations and two double divisions per (sample, primitive) pair.
struct alignas(16) PrimQuintic {
    double inv_A, B, C, p1A, inv_p1A;
    float q1x, q1y, q2x, q2y, q3x, q3y, s21, s11;
    float iv0;   // (float)(-B / 5) when that lies in [0, 1], else -1 ("no split point", Q10)
    int degenerate;   // |p1A| < 1e-6: the isolator is (at most) a quadratic (solve.h:30-36)
};
DVG_HD PrimQuintic prim_quintic(F2 p0, F2 p1, F2 p2, F2 p3) {
    const F2 q3 = -p0 + 3 * p1 - 3 * p2 + p3;
    const F2 q2 = 3 * p0 - 6 * p1 + 3 * p2;
    const F2 q1 = -3 * p0 + 3 * p1;
    const double A = 3 * sum2(q3 * q3);
    const double B = 5 * sum2(q3 * q2);
    const double C = 4 * sum2(q3 * q1) + 2 * sum2(q2 * q2);
    PrimQuintic k;
    k.inv_A = 1.0 / A;
    k.B = B * k.inv_A; k.C = C * k.inv_A;
    k.q1x = q1.x; k.q1y = q1.y; k.q2x = q2.x; k.q2y = q2.y; k.q3x = q3.x; k.q3y = q3.y;
    k.s21 = sum2(q2 * q1); k.s11 = sum2(q1 * q1);
    k.p1A = ((2 / 5.f) * k.C - (4 / 25.f) * k.B * k.B);
    k.degenerate = fabs(k.p1A) < 1e-6f ? 1 : 0;
    k.inv_p1A = 1.0 / k.p1A;
    const double q_root = -k.B / 5.f;
    k.iv0 = (q_root >= 0 && q_root <= 1) ? (float)q_root : -1.f;
    return k;
}
DVG_HD Quintic quintic_of(const PrimQuintic &k, F2 p0, F2 pt) {
    const F2 pp = p0 - pt;
    const F2 q1 = mk2(k.q1x, k.q1y), q2 = mk2(k.q2x, k.q2y), q3 = mk2(k.q3x, k.q3y);
    const double D = 3 * (k.s21 + sum2(q3 * pp));
    const double E = k.s11 + 2 * sum2(pp * q2);
    const double F = sum2(pp * q1);
    Quintic q;
    q.B = k.B; q.C = k.C; q.D = D * k.inv_A; q.E = E * k.inv_A; q.F = F * k.inv_A;
    return q;
}
// isolator_roots for the normalised cubic x^3 + b x^2 + c x + d (the caller divided by the leading coefficient): same
// estimate-then-polish scheme, divisions by constants written as multiplications (the closed form is only the
// starting point of cubic_polish).
// Cosine of the closed-form ESTIMATE (|x| <= pi): one MUFU on the device (absolute error ~5e-7), the polish removes it.
DVG_HD float est_cos(float x) {
#if defined(__CUDA_ARCH__)
    return __cosf(x);
#else
    return cosf(x);
#endif
}
// A split point only matters inside [0, 1]: negative ones are skipped and every one above 1 closes the last bracket at 1
// (within_distance.h:229-232), whatever its value.  An estimate that lies outside [0, 1] by far more than the closed
// form's error (<= 2e-4 of the magnitudes it subtracts, reached next to a double root) is therefore used as it is.
DVG_HD double polish_if_in_range(double b, double c, double d, double est, float scale) {
    const float m = 0.02f * scale + 0.01f;
    if (est < -(double)m || est > 1.0 + (double)m) return est;
    return cubic_polish(b, c, d, est);
}
DVG_HD int isolator_roots_monic(double b, double c, double d, double t[3]) {
    const double Q = (b * b - 3 * c) * (1.0 / 9.0);
    const double R = (2 * b * b * b - 9 * b * c + 27 * d) * (1.0 / 54.0);
    const double Q3 = Q * Q * Q;
    const double b3 = b * (1.0 / 3.0);
    if (R * R < Q3) {
        const float sq = sqrtf((float)Q);
        float x = (float)R / (sq * sq * sq);
        x = x < -1.f ? -1.f : (x > 1.f ? 1.f : x);
        const float theta = acosf(x);
        const float two_pi = 6.28318530717958647692f;
        const double m2sq = -2.0 * (double)sq;
        const float scale = 2.f * sq + fabsf((float)b3);
        t[0] = polish_if_in_range(b, c, d, m2sq * (double)est_cos(theta / 3.f) - b3, scale);
        t[1] = polish_if_in_range(b, c, d, m2sq * (double)est_cos((theta + two_pi) / 3.f) - b3, scale);
        t[2] = polish_if_in_range(b, c, d, m2sq * (double)est_cos((theta - two_pi) / 3.f) - b3, scale);
        return 3;
    } else {
        const float s = (float)sqrt(R * R - Q3);
        const float Af = R > 0 ? -cbrtf((float)R + s) : cbrtf((float)(-R) + s);
        const float Bf = fabsf(Af) > 1e-6f ? (float)Q / Af : 0.f;
        t[0] = polish_if_in_range(b, c, d, (double)(Af + Bf) - b3, fabsf(Af) + fabsf(Bf) + fabsf((float)b3));
        return 1;
    }
}
// quintic_intervals (fast isolator roots) from the per-primitive record
DVG_HD int quintic_intervals_of(const PrimQuintic &k, const Quintic &q, float intervals[4]) {
    const double p1B = ((3 / 5.f) * q.D - (3 / 25.f) * q.B * q.C);
    const double p1C = ((4 / 5.f) * q.E - (2 / 25.f) * q.B * q.D);
    const double p1D = q.F - (q.B * q.E) * (1.0 / 25.0);   // (a coefficient of the cubic whose roots are polished: 1 ulp is immaterial)
    double p_roots[3];
    int num_sol;
    if (k.degenerate) num_sol = solve_quadratic_d(p1B, p1C, p1D, &p_roots[0], &p_roots[1]) ? 2 : 0;
    else num_sol = isolator_roots_monic(p1B * k.inv_p1A, p1C * k.inv_p1A, p1D * k.inv_p1A, p_roots);
    intervals[0] = k.iv0;
    const int n = 1 + num_sol;
    // fixed trip counts (everything stays in registers); the insertion sort of within_distance.h:201-209 step for step,
    // including where it stops
    for (int j = 0; j < 3; j++) intervals[j + 1] = j < num_sol ? (float)p_roots[j] : 0.f;
    for (int j = 1; j < 4; j++) {
        bool go = j < n;
        for (int kk = j; kk > 0; kk--) {
            const float a = intervals[kk - 1], b = intervals[kk];
            go = go && a > b;
            intervals[kk - 1] = go ? b : a; intervals[kk] = go ? a : b;
        }
    }
    return n;
}

// Safeguarded Newton inside one bracket (within_distance.h:233-262).  Returns false when the
// bracket holds no sign change.
DVG_HD bool quintic_root_in(const Quintic &q, float lower, float upper, float *t_out) {
    float lb = lower, ub = upper;
    double lb_eval = quintic_eval(q, lb);
    double ub_eval = quintic_eval(q, ub);
    if (lb_eval * ub_eval > 0) return false;
    if (lb_eval > ub_eval) { float tmp = lb; lb = ub; ub = tmp; }
    float t = 0.5f * (lb + ub);
    for (int it = 0; it < 20; it++) {
        if (!(t >= lb && t <= ub)) t = 0.5f * (lb + ub);
        double value = quintic_eval(q, t);
        if (fabs(value) < 1e-5f || it == 19) break;
        if (value > 0.f) ub = t; else lb = t;
        double derivative = quintic_deriv(q, t);
        t = (float)((double)t - newton_quotient(value, derivative));
    }
    *t_out = t;
    return true;
}

// within_distance.h:119-272 (cubic leaf).  r[] = radius at the four control points.
DVG_HD bool stroke_hit_cubic(F2 p0, F2 p1, F2 p2, F2 p3, F4 r, F2 pt, float stale0 = -1.f) {
    if (dist_sq(p0, pt) < r.x * r.x) return true;  // eval(0) == p0 exactly
    if (dist_sq(p3, pt) < r.w * r.w) return true;  // eval(1) == p3 exactly
    Quintic q = cubic_quintic(p0, p1, p2, p3, pt);
    float intervals[4];
    int n = quintic_intervals(q, intervals, stale0);
    float lower_bound = 0.f;
    for (int j = 0; j < n + 1; j++) {
        if (j < n && intervals[j] < 0.f) continue;
        float upper_bound = j < n ? rminf(intervals[j], 1.f) : 1.f;
        float t;
        if (quintic_root_in(q, lower_bound, upper_bound, &t)) {
            float tt = 1 - t;
            float rr = (tt * tt * tt) * r.x + (3 * tt * tt * t) * r.y + (3 * tt * t * t) * r.z + (t * t * t) * r.w;
            if (dist_sq(eval_cubic(p0, p1, p2, p3, t), pt) < rr * rr) return true;
            if (upper_bound >= 1.f) break;
            lower_bound = upper_bound;
        }
        // note: when the bracket has no root the reference `continue`s WITHOUT advancing lower_bound
    }
    return false;
}

// within_distance.h:63-118 (quadratic leaf).  *decided is set when use_distance_approx makes
// the reference return from the whole path traversal (Q9).
DVG_HD_NOINLINE bool stroke_hit_quad(F2 p0, F2 p1, F2 p2, F4 r, float r_shape, bool approx, F2 pt, bool *decided) {
    if (approx) {
        F2 cp = quadratic_closest_pt_approx(p0, p1, p2, pt, nullptr);
        *decided = true;
        return dist_sq(cp, pt) < r_shape * r_shape;
    }
    if (dist_sq(p0, pt) < r.x * r.x) return true;
    if (dist_sq(p2, pt) < r.z * r.z) return true;
    F2 a2 = p0 - 2 * p1 + p2;
    F2 a1 = -p0 + p1;
    float A = sum2(a2 * a2);
    float B = sum2(3 * a2 * a1);
    float C = sum2(2 * a1 * a1 + a2 * (p0 - pt));
    float D = sum2(a1 * (p0 - pt));
    float t[3];
    int num_sol = solve_cubic_f(A, B, C, D, t);
    for (int j = 0; j < num_sol; j++) {
        if (t[j] >= 0 && t[j] <= 1) {
            float tt = 1 - t[j];
            float rr = (tt * tt) * r.x + (2 * tt * t[j]) * r.y + (t[j] * t[j]) * r.z;
            F2 p = eval_quad(p0, p1, p2, t[j]);
            if (dist_sq(p, pt) < rr * rr) return true;
        }
    }
    return false;
}

// within_distance.h:34-62 (line leaf); also the rect edge test 295-312 with r0 == r1.
DVG_HD bool stroke_hit_line(F2 p0, F2 p1, float r0, float r1, F2 pt) {
    float t = dot2(pt - p0, p1 - p0) / dot2(p1 - p0, p1 - p0);
    if (t < 0) {
        return dist_sq(p0, pt) < r0 * r0;
    } else if (t > 1) {
        return dist_sq(p1, pt) < r1 * r1;
    } else {
        float r = r0 + t * (r1 - r0);
        return dist_sq(p0 + t * (p1 - p0), pt) < r * r;
    }
}

// Polyline bracket test (see dvg_scene.cuh): -1 = certainly no hit, +1 = hit, 0 = undecided.
// `cap` = DVG_CAP_N records of 8 floats.  The chord distance evaluated at a slightly inexact t only
// over-estimates the distance by O(|d|^2 dt^2), far below the 1e-2 px margin folded into R_out / R_in.
DVG_HD int capsule_classify(const float *cap, F2 pt) {
    bool all_out = true, any_in = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < DVG_CAP_N; i++) {
#if defined(__CUDA_ARCH__)
        // the records are 16-byte aligned (SceneView::prim_cap is an F4 array): two 128-bit loads per piece
        const float4 ca = reinterpret_cast<const float4 *>(cap)[2 * i], cb = reinterpret_cast<const float4 *>(cap)[2 * i + 1];
        const float c0 = ca.x, c1 = ca.y, c2 = ca.z, c3 = ca.w, c4 = cb.x, c5 = cb.y, c6 = cb.z;
#else
        const float *c = cap + 8 * i;
        const float c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4], c5 = c[5], c6 = c[6];
#endif
        const float wx = pt.x - c0, wy = pt.y - c1;
        float t = (wx * c2 + wy * c3) * c4;
        t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
        const float ex = wx - t * c2, ey = wy - t * c3;
        const float d2 = ex * ex + ey * ey;
        all_out = all_out && (d2 > c5);   // false for NaN
        any_in = any_in || (d2 < c6);
    }
    return all_out ? -1 : (any_in ? 1 : 0);
}
DVG_HD bool capsule_reject(const float *cap, F2 pt) { return capsule_classify(cap, pt) < 0; }

// Stroke test of one primitive.  `r_shape` is shape.stroke_width (used by circle/rect and
// the distance-approx path).  *decided: see stroke_hit_quad.
DVG_HD bool prim_stroke_hit(int type, bool approx, F4 p01, F4 p23, F4 rad, float r_shape, F2 pt, bool *decided) {
    switch (type) {
        case PRIM_LINE:
            return stroke_hit_line(mk2(p01.x, p01.y), mk2(p01.z, p01.w), rad.x, rad.y, pt);
        case PRIM_QUAD:
            return stroke_hit_quad(mk2(p01.x, p01.y), mk2(p01.z, p01.w), mk2(p23.x, p23.y), rad, r_shape, approx, pt, decided);
        case PRIM_CUBIC:
            return stroke_hit_cubic(mk2(p01.x, p01.y), mk2(p01.z, p01.w), mk2(p23.x, p23.y), mk2(p23.z, p23.w), rad, pt);
        case PRIM_CIRCLE: {  // within_distance.h:8-16
            float d = distance2(mk2(p01.x, p01.y), pt);
            return fabsf(d - p01.z) < r_shape;
        }
        case PRIM_RECT: {  // within_distance.h:292-334
            F2 lt = mk2(p01.x, p01.y), rt = mk2(p01.z, p01.y), lb = mk2(p01.x, p01.w), rb = mk2(p01.z, p01.w);
            if (stroke_hit_line(lt, lb, r_shape, r_shape, pt)) return true;
            if (stroke_hit_line(lt, rt, r_shape, r_shape, pt)) return true;
            if (stroke_hit_line(rt, rb, r_shape, r_shape, pt)) return true;
            if (stroke_hit_line(lb, rb, r_shape, r_shape, pt)) return true;
            return false;
        }
        default:  // stroked ellipses are rejected at scene creation (Q6)
            return false;
    }
}

// The same without the cubic case: the render kernels answer cubic strokes with the warp solver
// (dvg_wsolve.cuh) and keep this rarely used remainder out of line (the quadratic's closed-form roots go
// through double acos / cos / pow: ~2 k instructions that would otherwise sit in the middle of the hot loop).
DVG_HD_NOINLINE bool prim_stroke_hit_nocubic(int type, bool approx, F4 p01, F4 p23, F4 rad, float r_shape, F2 pt, bool *decided) {
    if (type == PRIM_CUBIC) return false;
    return prim_stroke_hit(type, approx, p01, p23, rad, r_shape, pt, decided);
}

// solve.h:29-59 with T = double, the reference's operation sequence with correctly rounded acos / cos / pow (cold path of
// the winding test, see prim_winding).
DVG_HD_NOINLINE int solve_cubic_cr(double a, double b, double c, double d, double t[3]) {
    if (fabs(a) < 1e-6f) {
        if (solve_quadratic_d(b, c, d, &t[0], &t[1])) return 2;
        return 0;
    }
    b /= a; c /= a; d /= a;
    double Q = (b * b - 3 * c) / 9.f;
    double R = (2 * b * b * b - 9 * b * c + 27 * d) / 54.f;
    if (R * R < Q * Q * Q) {
        double theta = cr_acos(R / sqrt(Q * Q * Q));
        t[0] = -2.f * sqrt(Q) * cr_cos(theta / 3.f) - b / 3.f;
        t[1] = -2.f * sqrt(Q) * cr_cos((theta + 2.f * DVG_PI_D) / 3.f) - b / 3.f;
        t[2] = -2.f * sqrt(Q) * cr_cos((theta - 2.f * DVG_PI_D) / 3.f) - b / 3.f;
        return 3;
    } else {
        double A = R > 0 ? -cr_pow13(R + sqrt(R * R - Q * Q * Q)) : cr_pow13(-R + sqrt(R * R - Q * Q * Q));
        double B = fabs(A) > 1e-6f ? Q / A : 0.0;
        t[0] = (A + B) - b / 3.0;
        return 1;
    }
}

// winding_number.h:118-156: the reference's operation sequence for a cubic segment.
DVG_HD_NOINLINE int cubic_winding_exact(F2 p0, F2 p1, F2 p2, F2 p3, F2 pt) {
    double t[3];
    const double ca = (double)(-p0.y + 3 * p1.y - 3 * p2.y + p3.y), cb = (double)(3 * p0.y - 6 * p1.y + 3 * p2.y),
                 cc = (double)(-3 * p0.y + 3 * p1.y), cd = (double)(p0.y - pt.y);
    int num_sol = solve_cubic_d(ca, cb, cc, cd, t);
    // float coefficient * double t: the products are formed in double (winding_number.h:142-149)
    float cx3 = -p0.x + 3 * p1.x - 3 * p2.x + p3.x, cx2 = 3 * p0.x - 6 * p1.x + 3 * p2.x, cx1 = -3 * p0.x + 3 * p1.x;
    float cy3 = -p0.y + 3 * p1.y - 3 * p2.y + p3.y, cy2 = 3 * p0.y - 6 * p1.y + 3 * p2.y, cy1 = -3 * p0.y + 3 * p1.y;
    // A root within 1e-9 of a decision boundary (t = 0, t = 1, crossing exactly at the sample, horizontal tangent at
    // the crossing): the verdict hangs on the last bit of acos / cos / pow; solve again with correctly rounded
    // ones (dvg_crmath.cuh), which is what glibc returns for 99.9% of the arguments
    bool near = false;
    for (int j = 0; j < num_sol; j++) {
        const double tj = t[j];
        if (fabs(tj) < 1e-9 || fabs(tj - 1.0) < 1e-9) near = true;
        else if (tj > 0 && tj < 1) {
            const double tp = (double)cx3 * tj * tj * tj + (double)cx2 * tj * tj + (double)cx1 * tj + (double)p0.x - (double)pt.x;
            const double dy = (double)(3 * cy3) * tj * tj + (double)(2 * cy2) * tj + (double)cy1;
            if (fabs(tp) < 1e-9 * (1.0 + fabs((double)pt.x)) || fabs(dy) < 1e-9 * (fabs((double)cy1) + fabs((double)cy2) + fabs((double)cy3))) near = true;
        }
    }
    if (near) num_sol = solve_cubic_cr(ca, cb, cc, cd, t);
    int w = 0;
    for (int j = 0; j < num_sol; j++) {
        if (t[j] >= 0 && t[j] <= 1) {
            double tp = (double)cx3 * t[j] * t[j] * t[j] + (double)cx2 * t[j] * t[j] + (double)cx1 * t[j] +
                        (double)p0.x - (double)pt.x;
            if (tp > 0) {  // Q13: strict here, >= for lines and quadratics
                if ((double)(3 * cy3) * t[j] * t[j] + (double)(2 * cy2) * t[j] + (double)cy1 > 0) w += 1;
                else w -= 1;
            }
        }
    }
    return w;
}

// Winding contribution of a cubic segment, FAST form.  The reference (winding_number.h:118-156) solves y(t) = pt.y with the
// closed-form double solve_cubic (acos / cos / pow in double: several hundred FP64 instructions) and then takes four
// DECISIONS per root: t in [0, 1], crossing to the right of the sample (tp > 0), crossing direction (dy > 0).  Only the
// decisions matter.  Here the roots inside a window slightly wider than [0, 1] are found directly: the critical points
// of y cut the window into monotone pieces, a sign change across a piece is one root, found by a bracketed Newton in
// double (closer to the true root than the closed form, for any size of the leading coefficient).  The answer is
// taken only when every decision is clear by a margin 100 x wider than the band in which the reference's own verdict
// depends on its last bits (the 1e-9 band of cubic_winding_exact's correctly-rounded fall-back): roots 1e-7 away from
// 0 and 1, |tp| and |dy| 1e-7 (relative) away from 0 -- which also excludes near-double roots --, the Newton
// converged.  Anything else (about 1% of the pairs of an SVG asset rendered on a regular sample grid) returns false and
// takes the reference's operation sequence.
// `shortcut`: also try the no-root answer below (worth it where the lanes of a warp test the SAME segment -- the prefiltered
// kernels: -6% at flower.svg --, not in the pair-per-thread kernel of the sampled path, where it only adds divergence: +3%
// at tiger.svg).
DVG_HD bool cubic_winding_fast(F2 p0, F2 p1, F2 p2, F2 p3, F2 pt, int *w_out, bool shortcut = true) {
#if defined(DVG_FQ_ISOL) && !defined(DVG_WINDING_EXACT_ONLY)
    const float cx3 = -p0.x + 3 * p1.x - 3 * p2.x + p3.x, cx2 = 3 * p0.x - 6 * p1.x + 3 * p2.x, cx1 = -3 * p0.x + 3 * p1.x;
    const float cy3 = -p0.y + 3 * p1.y - 3 * p2.y + p3.y, cy2 = 3 * p0.y - 6 * p1.y + 3 * p2.y, cy1 = -3 * p0.y + 3 * p1.y;
    const double a = (double)cy3, b = (double)cy2, c = (double)cy1, d = (double)(p0.y - pt.y);
    if (!(fabs(a) >= 1e-5f)) return false;       // (at or next to the quadratic branch of solve.h:30-36 -- cheap there --, or NaN)
    // Where the reference's closed form is ill-conditioned its roots are NOT close to the true ones and the verdict is
    // whatever its arithmetic gives: R / sqrt(Q^3) next to +-1 (acos loses half its digits), which is a near-double root
    // or a leading coefficient that is small against the others (one root far away).  Qn, Rn: Q a^2 and R a^3.
    {
        const double Qn = (b * b - 3.0 * a * c) * (1.0 / 9.0);
        const double Rn = (2.0 * b * b * b - 9.0 * a * b * c + 27.0 * a * a * d) * (1.0 / 54.0);
        const double R2 = Rn * Rn, Q3 = Qn * Qn * Qn;
        // acos' error near +-1 is eps / sqrt(2 x relative discriminant), scaled into t by |b / 3a|: the threshold keeps the
        // reference's roots within ~1e-9 of the true ones (the margins below are 1e-7)
        const double thr_a2 = fmax(1e-10 * a * a, 1e-14 * b * b);
        if (!(fabs(R2 - Q3) * (a * a) > thr_a2 * (R2 + fabs(Q3)))) return false;
    }
    const double dy_scale = (double)(fabsf(cy1) + fabsf(cy2) + fabsf(cy3));
    // The whole segment to the right of the sample and y strictly monotone on [0, 1]: every crossing counts (tp > 0) and
    // there is exactly one if pt.y lies between the end points' y, none if it lies outside: no root needed.  Margins:
    // |y'| >= 1e-6 of its scale on [0, 1] (end values and, when the vertex of y' falls inside, the vertex value), pt.y
    // 1e-5 of that scale away from both end points (the root is then > 3e-6 away from 0 and 1), the sample 1e-6 left
    // of the control polygon.
    {
        const float min_x = fminf(fminf(p0.x, p1.x), fminf(p2.x, p3.x));
        if (shortcut && pt.x < min_x - 1e-6f * (1.f + fabsf(pt.x))) {
            const double d0 = c, d1 = 3.0 * a + 2.0 * b + c;     // y'(0), y'(1)
            const double m = 1e-6 * dy_scale;
            bool mono = (d0 > m && d1 > m) || (d0 < -m && d1 < -m);
            if (mono) {
                const double tv = -b / (3.0 * a);                 // vertex of y'
                if (tv > -0.01 && tv < 1.01) {
                    const double dv = c - b * b / (3.0 * a);
                    mono = d0 > 0 ? dv > m : dv < -m;
                }
            }
            if (mono) {
                const double e0 = d, e1 = a + b + c + d;      // y(0) - pt.y, y(1) - pt.y of the polynomial the reference solves
                const double my = 1e-5 * dy_scale;
                if (fabs(e0) > my && fabs(e1) > my) {
                    *w_out = ((e0 < 0) != (e1 < 0)) ? (d0 > 0 ? 1 : -1) : 0;
                    return true;
                }
            }
        }
    }
    const double lo = -0.01, hi = 1.01;          // roots outside are clearly outside [0, 1]
    // monotone pieces: break the window at the critical points of y (float precision is enough for a break point: two
    // roots closer than that to a critical point are a near-double root, whose |dy| fails the margin below)
    // (scalars, not an array: everything below stays in registers)
    double bk1 = hi, bk2 = hi;       // interior break points in (lo, hi), ascending; unused ones sit at hi
    int nb = 2;
    const double D = b * b - 3.0 * a * c;
    if (D > 0) {
        const double sD = (double)sqrtf((float)D);
        const double q = -(b + (b >= 0 ? sD : -sD));
        double t1 = (double)((float)q / (float)(3.0 * a)), t2 = q != 0.0 ? (double)((float)c / (float)q) : t1;
        if (t1 > t2) { const double tmp = t1; t1 = t2; t2 = tmp; }
        const bool in1 = t1 > lo && t1 < hi, in2 = t2 > lo && t2 < hi && t2 > t1;
        if (in1) { bk1 = t1; nb = 3; if (in2) { bk2 = t2; nb = 4; } }
        else if (in2) { bk1 = t2; nb = 3; }
    }
    // values at the break points; the pieces with a sign change are collected first and THEN solved one per trip, so that
    // the lanes of a warp solve together whichever piece their (usually single) root lies in
    const double yb0 = fma(fma(fma(a, lo, b), lo, c), lo, d);
    const double yb1 = fma(fma(fma(a, bk1, b), bk1, c), bk1, d);
    const double yb2 = fma(fma(fma(a, bk2, b), bk2, c), bk2, d);
    const double yb3 = fma(fma(fma(a, hi, b), hi, c), hi, d);
    // piece k = [bk_k, bk_{k+1}] of the nb - 1 pieces (the last break point is hi)
    if (!(yb0 != 0.0 && yb3 != 0.0 && yb0 == yb0 && yb3 == yb3)) return false;
    if (nb >= 3 && !(yb1 != 0.0 && yb1 == yb1)) return false;
    if (nb >= 4 && !(yb2 != 0.0 && yb2 == yb2)) return false;
    int pieces = 0, np = 0;          // indices of the pieces with a sign change, 2 bits each
    {
        const double e1 = nb >= 3 ? yb1 : yb3;                 // end of piece 0
        if ((yb0 > 0) != (e1 > 0)) { pieces |= 0 << (2 * np); np++; }
        if (nb >= 3) {
            const double e2 = nb >= 4 ? yb2 : yb3;             // end of piece 1
            if ((yb1 > 0) != (e2 > 0)) { pieces |= 1 << (2 * np); np++; }
            if (nb >= 4 && (yb2 > 0) != (yb3 > 0)) { pieces |= 2 << (2 * np); np++; }
        }
    }
    int w = 0;
    for (int sidx = 0; sidx < np; sidx++) {
        const int k = (pieces >> (2 * sidx)) & 3;
        const double u = k == 0 ? lo : (k == 1 ? bk1 : bk2);
        const double v = (k == nb - 2) ? hi : (k == 0 ? bk1 : bk2);
        const double yu0 = k == 0 ? yb0 : (k == 1 ? yb1 : yb2);
        const double yv = (k == nb - 2) ? yb3 : (k == 0 ? yb1 : yb2);
        // the root of this piece: four bracketed Newton steps in float from the secant point (fixed trip count: the lanes
        // of a warp stay together), then three steps in double whose correction needs float precision only, then a check
        // that one more step would not move it
        float uf = (float)u, vf = (float)v, fuf = (float)yu0;
        const float df = (float)d;
        float tf = uf - fuf * (vf - uf) / ((float)yv - fuf);
        for (int it = 0; it < 4; it++) {
            if (!(tf > uf && tf < vf)) tf = 0.5f * (uf + vf);
            const float ff = ((cy3 * tf + cy2) * tf + cy1) * tf + df;
            const float fpf = (3.f * cy3 * tf + 2.f * cy2) * tf + cy1;
            if ((ff > 0.f) == (fuf > 0.f)) { uf = tf; fuf = ff; } else vf = tf;
            if (fabsf(fpf) > 1e-30f) tf -= ff * seed_rcp(fpf);
        }
        double t = (double)tf;
        if (!(t > u && t < v)) t = 0.5 * ((double)uf + (double)vf);
        bool converged = false;
        for (int it = 0; it < 4; it++) {
            const double f = fma(fma(fma(a, t, b), t, c), t, d);
            const double fp = fma(fma(3.0 * a, t, 2.0 * b), t, c);
            if (it == 3) { converged = fabs(f) <= 1e-11 * fabs(fp); break; }
            const float fpf = (float)fp;
            if (!(fabsf(fpf) > 1e-30f && fabsf(fpf) < 1e30f)) break;
            t -= (double)((float)f * seed_rcp(fpf));
        }
        if (!converged) return false;
        if (!(fabs(t) > 1e-7 && fabs(t - 1.0) > 1e-7)) return false;
        if (t < 0 || t > 1) continue;
        const double tp = (double)cx3 * t * t * t + (double)cx2 * t * t + (double)cx1 * t + (double)p0.x - (double)pt.x;
        const double dy = (double)(3 * cy3) * t * t + (double)(2 * cy2) * t + (double)cy1;
        if (!(fabs(tp) > 1e-7 * (1.0 + fabs((double)pt.x)) && fabs(dy) > 1e-7 * dy_scale)) return false;
        if (tp > 0) w += dy > 0 ? 1 : -1;
    }
    *w_out = w;
    return true;
#else
    return false;
#endif
}

// The no-root answer of cubic_winding_fast with its sample-independent part formed once per primitive, for the
// CLASSIFIER (dvg_wave.cu): all lanes of a warp test the same segment there, so answering a winding test in place costs a
// dozen FP64 operations and saves a 16-byte pair record written, read and solved later -- at tiger.svg the winding pairs
// were 5 GB of queue traffic per step.  Valid for primitives flagged DVG_PF_YMONO (y strictly monotone on [0, 1] by
// the margins of cubic_winding_fast, leading coefficient not small).
struct alignas(16) PrimWindCert {
    double Q3, R0, K, thr, s;     // Qn^3; Rn = R0 + K d; threshold on the relative discriminant; a + b + c
    float p0y, my;                // d = p0y - pt.y; margin of pt.y against the end points
};
// returns the flags to add to the primitive (0, or DVG_PF_YMONO [| DVG_PF_YUP])
DVG_HD int prim_wind_cert(F2 p0, F2 p1, F2 p2, F2 p3, PrimWindCert &k) {
    const float cy3 = -p0.y + 3 * p1.y - 3 * p2.y + p3.y, cy2 = 3 * p0.y - 6 * p1.y + 3 * p2.y, cy1 = -3 * p0.y + 3 * p1.y;
    const double a = (double)cy3, b = (double)cy2, c = (double)cy1;
    k.Q3 = k.R0 = k.K = k.thr = k.s = 0.0; k.p0y = p0.y; k.my = 0.f;
    if (!(fabs(a) >= 1e-5f)) return 0;
    const double dy_scale = (double)(fabsf(cy1) + fabsf(cy2) + fabsf(cy3));
    const double d0 = c, d1 = 3.0 * a + 2.0 * b + c;
    const double m = 1e-6 * dy_scale;
    bool mono = (d0 > m && d1 > m) || (d0 < -m && d1 < -m);
    if (mono) {
        const double tv = -b / (3.0 * a);
        if (tv > -0.01 && tv < 1.01) {
            const double dv = c - b * b / (3.0 * a);
            mono = d0 > 0 ? dv > m : dv < -m;
        }
    }
    if (!mono) return 0;
    const double Qn = (b * b - 3.0 * a * c) * (1.0 / 9.0);
    k.Q3 = Qn * Qn * Qn;
    k.R0 = (2.0 * b * b * b - 9.0 * a * b * c) * (1.0 / 54.0);
    k.K = 0.5 * a * a;
    k.thr = fmax(1e-10, 1e-14 * (b * b) / (a * a));
    k.s = a + b + c;
    k.my = (float)(1e-5 * dy_scale) * 1.0001f + 1e-30f;    // (rounded up: never below the double margin)
    return DVG_PF_YMONO | (d0 > 0 ? DVG_PF_YUP : 0);
}
// The classifier's test: true when the winding contribution of the flagged segment is certain (*w = -1, 0, +1).
// `box_x0` = the smallest x of the control points (the primitive's leaf box).
DVG_HD bool wind_cert_answer(const PrimWindCert &k, bool up, float box_x0, F2 pt, int *w) {
    if (!(pt.x < box_x0 - 1e-6f * (1.f + fabsf(pt.x)))) return false;
    const double d = (double)(k.p0y - pt.y);
    const double Rn = k.R0 + k.K * d, R2 = Rn * Rn;
    if (!(fabs(R2 - k.Q3) > k.thr * (R2 + fabs(k.Q3)))) return false;
    const double e1 = k.s + d, my = (double)k.my;
    if (!(fabs(d) > my && fabs(e1) > my)) return false;
    *w = ((d < 0) != (e1 < 0)) ? (up ? 1 : -1) : 0;
    return true;
}

// winding_number.h:62-156 per leaf type + 9-31, 176-186 for the closed-form shapes.
DVG_HD_NOINLINE int prim_winding(int type, F4 p01, F4 p23, F2 pt, bool shortcut = true) {
    switch (type) {
        case PRIM_LINE: {
            F2 p0 = mk2(p01.x, p01.y), p1 = mk2(p01.z, p01.w);
            if (p1.y != p0.y) {
                float t = (pt.y - p0.y) / (p1.y - p0.y);
                if (t >= 0 && t <= 1) {
                    float tp = p0.x - pt.x + t * (p1.x - p0.x);
                    if (tp >= 0) return (p1.y - p0.y > 0) ? 1 : -1;
                }
            }
            return 0;
        }
        case PRIM_QUAD: {
            F2 p0 = mk2(p01.x, p01.y), p1 = mk2(p01.z, p01.w), p2 = mk2(p23.x, p23.y);
            float t[2];
            int w = 0;
            if (solve_quadratic_f(p0.y - 2 * p1.y + p2.y, -2 * p0.y + 2 * p1.y, p0.y - pt.y, &t[0], &t[1])) {
                for (int j = 0; j < 2; j++) {
                    if (t[j] >= 0 && t[j] <= 1) {
                        float tp = (p0.x - 2 * p1.x + p2.x) * t[j] * t[j] + (-2 * p0.x + 2 * p1.x) * t[j] + p0.x - pt.x;
                        if (tp >= 0) {
                            if (2 * (p0.y - 2 * p1.y + p2.y) * t[j] + (-2 * p0.y + 2 * p1.y) > 0) w += 1;
                            else w -= 1;
                        }
                    }
                }
            }
            return w;
        }
        case PRIM_CUBIC: {
            const F2 p0 = mk2(p01.x, p01.y), p1 = mk2(p01.z, p01.w), p2 = mk2(p23.x, p23.y), p3 = mk2(p23.z, p23.w);
            int wf;
            if (cubic_winding_fast(p0, p1, p2, p3, pt, &wf, shortcut)) return wf;
            return cubic_winding_exact(p0, p1, p2, p3, pt);
        }
        case PRIM_CIRCLE:
            return dist_sq(mk2(p01.x, p01.y), pt) < p01.z * p01.z ? 1 : 0;
        case PRIM_ELLIPSE: {
            float ex = p01.x - pt.x, ey = p01.y - pt.y;
            return (ex * ex) / (p01.z * p01.z) + (ey * ey) / (p01.w * p01.w) < 1 ? 1 : 0;
        }
        case PRIM_RECT:
            return (pt.x > p01.x && pt.x < p01.z && pt.y > p01.y && pt.y < p01.w) ? 1 : 0;
    }
    return 0;
}

}  // namespace dvg
