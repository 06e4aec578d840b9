/*
 * oracle/dvg_oracle.c -- TEST INFRASTRUCTURE.  A plain-C restatement of the reference's FORWARD colour path
 * (pydiffvg.RenderFunction.apply -> diffvg.cpp render(), no prefiltering, no SDF output) on the packed scene of
 * include/dvg_scene_format.h.  Single-threaded, no acceleration structure: every sample visits every shape group
 * in ascending group id, which is the order the reference composites in after its fragment sort
 * (diffvg.cpp:605-615); the reference's three BVH levels only cull conservatively, so the fragments are the same.
 *
 * Only tests/ (and through them bench.py / smoke() when oracle/_ref is absent) may load this library; the product
 * never does.  It is pinned against the compiled reference (oracle/_ref) through the committed golden vectors
 * (tests/test_cpu_oracle_and_host.py::test_c_restatement_*).  The backward pass, prefiltering and the SDF output
 * are NOT restated here: dvgo_render returns an error for them and oracle_check falls through to oracle/_ref.
 *
 * Each function cites the reference lines it follows.  Arithmetic types follow the reference: float unless it
 * silently promotes (unqualified sqrt/acos/cos/pow on floats are the double overloads in its build; solve_cubic is
 * instantiated with double for the cubic closest-point and winding tests).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dvg_scene_format.h"

#define EXPORT __attribute__((visibility("default")))

static char g_err[256] = "";
static int fail(const char *m) { snprintf(g_err, sizeof g_err, "%s", m); return 1; }
EXPORT const char *dvgo_last_error(void) { return g_err; }

typedef struct { float x, y; } v2;
static v2 V2(float x, float y) { v2 r; r.x = x; r.y = y; return r; }
static v2 sub(v2 a, v2 b) { return V2(a.x - b.x, a.y - b.y); }
static v2 add(v2 a, v2 b) { return V2(a.x + b.x, a.y + b.y); }
static v2 mul(float s, v2 a) { return V2(s * a.x, s * a.y); }
static float dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static float dist2(v2 a, v2 b) { v2 d = sub(b, a); return d.x * d.x + d.y * d.y; }   /* vector.h:478-481 */

/* ---------------------------------------------------------------- pcg.h:11-40 */
typedef struct { uint64_t state, inc; } pcg32;
static uint32_t pcg_next(pcg32 *r) {
    uint64_t old = r->state;
    r->state = old * 6364136223846793005ULL + (r->inc | 1);
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((0u - rot) & 31));
}
static float pcg_float(pcg32 *r) {
    union { uint32_t u; float f; } x;
    x.u = (pcg_next(r) >> 9) | 0x3f800000u;
    return x.f - 1.0f;
}
static void pcg_init(pcg32 *r, int idx, uint64_t seed) {
    r->state = 0;
    r->inc = (((uint64_t)idx + 1) << 1u) | 1u;
    pcg_next(r);
    r->state += 0x853c49e6748fea9bULL + seed;
    pcg_next(r);
}
EXPORT void dvgo_pcg(int idx, uint64_t seed, uint64_t *state, float *rx, float *ry) {
    pcg32 r;
    pcg_init(&r, idx, seed);
    *state = r.state;
    *rx = pcg_float(&r);
    *ry = pcg_float(&r);
}

/* ---------------------------------------------------------------- solve.h:5-59 */
static int solve_quadratic_d(double a, double b, double c, double *t0, double *t1) {
    double disc = b * b - 4 * a * c;
    if (disc < 0) return 0;
    double rd = sqrt(disc), q = b < 0 ? -0.5f * (b - rd) : -0.5f * (b + rd);
    *t0 = q / a; *t1 = c / q;
    if (*t0 > *t1) { double t = *t0; *t0 = *t1; *t1 = t; }
    return 1;
}
static int solve_quadratic_f(float a, float b, float c, float *t0, float *t1) {
    float disc = b * b - 4 * a * c;
    if (disc < 0) return 0;
    float rd = sqrtf(disc), q = b < 0 ? -0.5f * (b - rd) : -0.5f * (b + rd);
    *t0 = q / a; *t1 = c / q;
    if (*t0 > *t1) { float t = *t0; *t0 = *t1; *t1 = t; }
    return 1;
}
static int solve_cubic_d(double a, double b, double c, double d, double t[3]) {
    if (fabs(a) < 1e-6f) return solve_quadratic_d(b, c, d, &t[0], &t[1]) ? 2 : 0;
    b /= a; c /= a; d /= a;
    double Q = (b * b - 3 * c) / 9.f, R = (2 * b * b * b - 9 * b * c + 27 * d) / 54.f;
    if (R * R < Q * Q * Q) {
        double th = acos(R / sqrt(Q * Q * Q));
        t[0] = -2.f * sqrt(Q) * cos(th / 3.f) - b / 3.f;
        t[1] = -2.f * sqrt(Q) * cos((th + 2.f * M_PI) / 3.f) - b / 3.f;
        t[2] = -2.f * sqrt(Q) * cos((th - 2.f * M_PI) / 3.f) - b / 3.f;
        return 3;
    }
    double A = R > 0 ? -pow(R + sqrt(R * R - Q * Q * Q), 1. / 3.) : pow(-R + sqrt(R * R - Q * Q * Q), 1. / 3.);
    double B = fabs(A) > 1e-6f ? Q / A : 0.0;
    t[0] = (A + B) - b / 3.0;
    return 1;
}
static int solve_cubic_f(float a, float b, float c, float d, float t[3]) {   /* T = float: libm calls are still double */
    if (fabsf(a) < 1e-6f) return solve_quadratic_f(b, c, d, &t[0], &t[1]) ? 2 : 0;
    b /= a; c /= a; d /= a;
    float Q = (b * b - 3 * c) / 9.f, R = (2 * b * b * b - 9 * b * c + 27 * d) / 54.f;
    if (R * R < Q * Q * Q) {
        float th = (float)acos((double)R / sqrt((double)(Q * Q * Q)));
        double m = (double)(-2.f) * sqrt((double)Q);
        float pi_f = (float)M_PI;
        t[0] = (float)(m * cos((double)(th / 3.f)) - (double)(b / 3.f));
        t[1] = (float)(m * cos((double)((th + 2.f * pi_f) / 3.f)) - (double)(b / 3.f));
        t[2] = (float)(m * cos((double)((th - 2.f * pi_f) / 3.f)) - (double)(b / 3.f));
        return 3;
    }
    double third = (double)(float)(1. / 3.), sq = sqrt((double)(R * R - Q * Q * Q));
    float A = R > 0 ? (float)(-pow((double)R + sq, third)) : (float)pow((double)(-R) + sq, third);
    float B = fabsf(A) > 1e-6f ? Q / A : 0.f;
    t[0] = (A + B) - b / 3.f;
    return 1;
}

/* ---------------------------------------------------------------- within_distance.h:34-272 (per segment) */
static int near_line(v2 p0, v2 p1, float r0, float r1, v2 pt) {            /* :34-62 */
    float t = dot(sub(pt, p0), sub(p1, p0)) / dot(sub(p1, p0), sub(p1, p0));
    if (t < 0) return dist2(p0, pt) < r0 * r0;
    if (t > 1) return dist2(p1, pt) < r1 * r1;
    float r = r0 + t * (r1 - r0);
    return dist2(add(p0, mul(t, sub(p1, p0))), pt) < r * r;
}
static int near_quad(v2 p0, v2 p1, v2 p2, float r0, float r1, float r2, v2 pt) {   /* :63-118 */
    if (dist2(p0, pt) < r0 * r0) return 1;
    if (dist2(p2, pt) < r2 * r2) return 1;
    v2 a2 = add(sub(p0, mul(2, p1)), p2), a1 = sub(p1, p0), pp = sub(p0, pt);
    float A = a2.x * a2.x + a2.y * a2.y;
    float B = 3 * a2.x * a1.x + 3 * a2.y * a1.y;
    float C = (2 * a1.x * a1.x + a2.x * pp.x) + (2 * a1.y * a1.y + a2.y * pp.y);
    float D = a1.x * pp.x + a1.y * pp.y;
    float t[3];
    int n = solve_cubic_f(A, B, C, D, t);
    for (int j = 0; j < n; j++) {
        if (t[j] >= 0 && t[j] <= 1) {
            float tt = 1 - t[j];
            float r = (tt * tt) * r0 + (2 * tt * t[j]) * r1 + (t[j] * t[j]) * r2;
            v2 p = add(add(mul(tt * tt, p0), mul(2 * tt * t[j], p1)), mul(t[j] * t[j], p2));
            if (dist2(p, pt) < r * r) return 1;
        }
    }
    return 0;
}
static v2 cubic_at(v2 p0, v2 p1, v2 p2, v2 p3, float t) {                 /* :129-132 */
    float tt = 1 - t;
    return add(add(mul(tt * tt * tt, p0), mul(3 * tt * tt * t, p1)), add(mul(3 * tt * t * t, p2), mul(t * t * t, p3)));
}
static int near_cubic(v2 p0, v2 p1, v2 p2, v2 p3, const float r[4], v2 pt) {       /* :119-272 */
    if (dist2(p0, pt) < r[0] * r[0]) return 1;
    if (dist2(p3, pt) < r[3] * r[3]) return 1;
    v2 q3 = add(add(mul(-1, p0), mul(3, p1)), add(mul(-3, p2), p3));
    v2 q2 = add(add(mul(3, p0), mul(-6, p1)), mul(3, p2));
    v2 q1 = add(mul(-3, p0), mul(3, p1));
    v2 pp = sub(p0, pt);
    /* coefficients of d/dt |q(t) - pt|^2 / 2, formed in float, normalised in double (:161-172) */
    double A = 3 * (q3.x * q3.x + q3.y * q3.y);
    double B = 5 * (q3.x * q2.x + q3.y * q2.y);
    double C = 4 * (q3.x * q1.x + q3.y * q1.y) + 2 * (q2.x * q2.x + q2.y * q2.y);
    double D = 3 * ((q2.x * q1.x + q2.y * q1.y) + (q3.x * pp.x + q3.y * pp.y));
    double E = (q1.x * q1.x + q1.y * q1.y) + 2 * (pp.x * q2.x + pp.y * q2.y);
    double F = pp.x * q1.x + pp.y * q1.y;
    B /= A; C /= A; D /= A; E /= A; F /= A;
    /* isolator polynomials (:184-210): the roots of the cubic p1 and of the linear q split [0,1] so that every
       piece holds at most one root of the quintic */
    double p1A = (2 / 5.f) * C - (4 / 25.f) * B * B;
    double p1B = (3 / 5.f) * D - (3 / 25.f) * B * C;
    double p1C = (4 / 5.f) * E - (2 / 25.f) * B * D;
    double p1D = F - B * E / 25.f;
    double q_root = -B / 5.f;
    double pr[3];
    int ns = solve_cubic_d(p1A, p1B, p1C, p1D, pr);
    float iv[4];
    /* Q10: the reference leaves intervals[0] unwritten when q_root is outside [0,1]; "no split point" here (the
       parity build of the reference pins that read to 0, which yields an empty bracket) */
    iv[0] = (q_root >= 0 && q_root <= 1) ? (float)q_root : -1.f;
    for (int j = 0; j < ns; j++) iv[j + 1] = (float)pr[j];
    int n = 1 + ns;
    for (int j = 1; j < n; j++)
        for (int k = j; k > 0 && iv[k - 1] > iv[k]; k--) { float t = iv[k]; iv[k] = iv[k - 1]; iv[k - 1] = t; }
#define QUINTIC(t) ((t) * (t) * (t) * (t) * (t) + B * (t) * (t) * (t) * (t) + C * (t) * (t) * (t) + D * (t) * (t) + E * (t) + F)
#define DQUINTIC(t) (5 * (t) * (t) * (t) * (t) + 4 * B * (t) * (t) * (t) + 3 * C * (t) * (t) + 2 * D * (t) + E)
    float lower = 0.f;
    for (int j = 0; j < n + 1; j++) {
        if (j < n && iv[j] < 0.f) continue;
        float upper = j < n ? (iv[j] < 1.f ? iv[j] : 1.f) : 1.f;
        float lb = lower, ub = upper;
        double lbe = QUINTIC((double)lb), ube = QUINTIC((double)ub);
        if (lbe * ube > 0) continue;                       /* no sign change: `lower` stays where it is (:238) */
        if (lbe > ube) { float t = lb; lb = ub; ub = t; }
        float t = 0.5f * (lb + ub);
        for (int it = 0; it < 20; it++) {                  /* safeguarded Newton (:244-262) */
            if (!(t >= lb && t <= ub)) t = 0.5f * (lb + ub);
            double v = QUINTIC((double)t);
            if (fabs(v) < 1e-5f || it == 19) break;
            if (v > 0.f) ub = t; else lb = t;
            double dv = DQUINTIC((double)t);
            t = (float)((double)t - v / dv);
        }
        float tt = 1 - t;
        float rr = (tt * tt * tt) * r[0] + (3 * tt * tt * t) * r[1] + (3 * tt * t * t) * r[2] + (t * t * t) * r[3];
        if (dist2(cubic_at(p0, p1, p2, p3, t), pt) < rr * rr) return 1;
        if (upper >= 1.f) break;
        lower = upper;
    }
#undef QUINTIC
#undef DQUINTIC
    return 0;
}

/* ---------------------------------------------------------------- winding_number.h:62-156 (per segment) */
static int wind_line(v2 p0, v2 p1, v2 pt) {
    if (p1.y != p0.y) {
        float t = (pt.y - p0.y) / (p1.y - p0.y);
        if (t >= 0 && t <= 1) {
            float tp = p0.x - pt.x + t * (p1.x - p0.x);
            if (tp >= 0) return (p1.y - p0.y > 0) ? 1 : -1;
        }
    }
    return 0;
}
static int wind_quad(v2 p0, v2 p1, v2 p2, v2 pt) {
    float t[2];
    int w = 0;
    if (solve_quadratic_f(p0.y - 2 * p1.y + p2.y, -2 * p0.y + 2 * p1.y, p0.y - pt.y, &t[0], &t[1])) {
        for (int j = 0; j < 2; j++) {
            if (t[j] >= 0 && t[j] <= 1) {
                float tp = (p0.x - 2 * p1.x + p2.x) * t[j] * t[j] + (-2 * p0.x + 2 * p1.x) * t[j] + p0.x - pt.x;
                if (tp >= 0) w += (2 * (p0.y - 2 * p1.y + p2.y) * t[j] + (-2 * p0.y + 2 * p1.y) > 0) ? 1 : -1;
            }
        }
    }
    return w;
}
static int wind_cubic(v2 p0, v2 p1, v2 p2, v2 p3, v2 pt) {
    float cy3 = -p0.y + 3 * p1.y - 3 * p2.y + p3.y, cy2 = 3 * p0.y - 6 * p1.y + 3 * p2.y, cy1 = -3 * p0.y + 3 * p1.y;
    float cx3 = -p0.x + 3 * p1.x - 3 * p2.x + p3.x, cx2 = 3 * p0.x - 6 * p1.x + 3 * p2.x, cx1 = -3 * p0.x + 3 * p1.x;
    double t[3];
    int n = solve_cubic_d((double)cy3, (double)cy2, (double)cy1, (double)(p0.y - pt.y), t), w = 0;
    for (int j = 0; j < n; j++) {
        if (t[j] >= 0 && t[j] <= 1) {
            double tp = (double)cx3 * t[j] * t[j] * t[j] + (double)cx2 * t[j] * t[j] + (double)cx1 * t[j] + (double)p0.x - (double)pt.x;
            if (tp > 0)   /* strict here, >= for lines and quadratics (Q13) */
                w += ((double)(3 * cy3) * t[j] * t[j] + (double)(2 * cy2) * t[j] + (double)cy1 > 0) ? 1 : -1;
        }
    }
    return w;
}

/* ---------------------------------------------------------------- scene access */
typedef struct {
    const int32_t *topo;
    const float *params;
    int cw, ch, ns, ng;
} scene_t;

static const int32_t *srec(const scene_t *s, int i) { return s->topo + s->topo[DVG_H_OFF_SHAPES] + i * DVG_SHAPE_REC_LEN; }
static const int32_t *grec(const scene_t *s, int g) { return s->topo + s->topo[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN; }

/* matrix.h:209-231, 505-512 */
static void inverse3(const float *m, float *o) {
    float det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    float id = 1 / det;
    o[0] = (m[4] * m[8] - m[7] * m[5]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[3] * m[2] - m[0] * m[5]) * id;
    o[6] = (m[3] * m[7] - m[6] * m[4]) * id; o[7] = (m[6] * m[1] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[3] * m[1]) * id;
}
static v2 xform_pt(const float *m, v2 p) {
    float t0 = m[0] * p.x + m[1] * p.y + m[2], t1 = m[3] * p.x + m[4] * p.y + m[5], t2 = m[6] * p.x + m[7] * p.y + m[8];
    return V2(t0 / t2, t1 / t2);
}

static v2 path_pt(const float *pts, int i) { return V2(pts[2 * i], pts[2 * i + 1]); }

/* within_distance.h:337-432 for one shape (all its segments; the path BVH only culls) */
static int shape_stroke_hit(const scene_t *s, const int32_t *sr, v2 pt) {
    const float *p = s->params + sr[DVG_S_PARAM_OFF];
    float sw = sr[DVG_S_WIDTH_OFF] >= 0 ? s->params[sr[DVG_S_WIDTH_OFF]] : 0.f;
    switch (sr[DVG_S_TYPE]) {
        case DVG_SHAPE_CIRCLE: {                               /* :8-16 */
            float d = sqrtf(dist2(V2(p[1], p[2]), pt));
            return fabsf(d - p[0]) < sw;
        }
        case DVG_SHAPE_RECT: {                                 /* :292-334 */
            v2 lt = V2(p[0], p[1]), rt = V2(p[2], p[1]), lb = V2(p[0], p[3]), rb = V2(p[2], p[3]);
            return near_line(lt, lb, sw, sw, pt) || near_line(lt, rt, sw, sw, pt) || near_line(rt, rb, sw, sw, pt) ||
                   near_line(lb, rb, sw, sw, pt);
        }
        case DVG_SHAPE_PATH: {
            const float *th = sr[DVG_S_THICK_OFF] >= 0 ? s->params + sr[DVG_S_THICK_OFF] : NULL;
            const int32_t *ncp = s->topo + s->topo[DVG_H_OFF_NCP] + sr[DVG_S_NCP_OFF];
            int np = sr[DVG_S_NUM_POINTS], pid = 0;
            for (int k = 0; k < sr[DVG_S_NUM_SEGS]; k++) {
                int i0 = pid, i1, i2, i3;
                if (ncp[k] == 0) {
                    i1 = (pid + 1) % np;
                    if (near_line(path_pt(p, i0), path_pt(p, i1), th ? th[i0] : sw, th ? th[i1] : sw, pt)) return 1;
                    pid += 1;
                } else if (ncp[k] == 1) {
                    i1 = pid + 1; i2 = (pid + 2) % np;
                    if (near_quad(path_pt(p, i0), path_pt(p, i1), path_pt(p, i2), th ? th[i0] : sw, th ? th[i1] : sw,
                                  th ? th[i2] : sw, pt)) return 1;
                    pid += 2;
                } else {
                    i1 = pid + 1; i2 = pid + 2; i3 = (pid + 3) % np;
                    float r[4] = {th ? th[i0] : sw, th ? th[i1] : sw, th ? th[i2] : sw, th ? th[i3] : sw};
                    if (near_cubic(path_pt(p, i0), path_pt(p, i1), path_pt(p, i2), path_pt(p, i3), r, pt)) return 1;
                    pid += 3;
                }
            }
            return 0;
        }
        default: return 0;   /* stroked ellipses assert in the reference (within_distance.h:342-345) */
    }
}

/* winding_number.h:9-31, 159-202 for one shape */
static int shape_winding(const scene_t *s, const int32_t *sr, v2 pt) {
    const float *p = s->params + sr[DVG_S_PARAM_OFF];
    switch (sr[DVG_S_TYPE]) {
        case DVG_SHAPE_CIRCLE: return dist2(V2(p[1], p[2]), pt) < p[0] * p[0] ? 1 : 0;
        case DVG_SHAPE_ELLIPSE: {
            float ex = p[2] - pt.x, ey = p[3] - pt.y;
            return (ex * ex) / (p[0] * p[0]) + (ey * ey) / (p[1] * p[1]) < 1 ? 1 : 0;
        }
        case DVG_SHAPE_RECT: return (pt.x > p[0] && pt.x < p[2] && pt.y > p[1] && pt.y < p[3]) ? 1 : 0;
        default: {
            const int32_t *ncp = s->topo + s->topo[DVG_H_OFF_NCP] + sr[DVG_S_NCP_OFF];
            int np = sr[DVG_S_NUM_POINTS], pid = 0, w = 0;
            for (int k = 0; k < sr[DVG_S_NUM_SEGS]; k++) {
                if (ncp[k] == 0) { w += wind_line(path_pt(p, pid), path_pt(p, (pid + 1) % np), pt); pid += 1; }
                else if (ncp[k] == 1) { w += wind_quad(path_pt(p, pid), path_pt(p, pid + 1), path_pt(p, (pid + 2) % np), pt); pid += 2; }
                else { w += wind_cubic(path_pt(p, pid), path_pt(p, pid + 1), path_pt(p, pid + 2), path_pt(p, (pid + 3) % np), pt); pid += 3; }
            }
            return w;
        }
    }
}

/* diffvg.cpp:276-368 */
static void eval_color(int type, const float *c, int stops, v2 pt, float out[4]) {
    if (type == DVG_COLOR_CONSTANT) { memcpy(out, c, 16); return; }
    float t;
    if (type == DVG_COLOR_LINEAR) {
        v2 beg = V2(c[0], c[1]), end = V2(c[2], c[3]);
        float l = dot(sub(end, beg), sub(end, beg));
        t = dot(sub(pt, beg), sub(end, beg)) / (l > 1e-3f ? l : 1e-3f);
    } else {
        v2 o = sub(pt, V2(c[0], c[1]));
        float nx = o.x / c[2], ny = o.y / c[3];
        t = sqrtf(nx * nx + ny * ny);
    }
    const float *off = c + 4, *col = c + 4 + stops;
    if (t < off[0]) { memcpy(out, col, 16); return; }
    for (int i = 0; i < stops - 1; i++) {
        if (t >= off[i] && t < off[i + 1]) {
            float tt = (t - off[i]) / (off[i + 1] - off[i]);
            for (int k = 0; k < 4; k++) out[k] = col[4 * i + k] * (1 - tt) + col[4 * (i + 1) + k] * tt;
            return;
        }
    }
    memcpy(out, col + 4 * (stops - 1), 16);
}

/* diffvg.cpp:525-653 without the EdgeQuery */
static void sample_color(const scene_t *s, const float *const *c2s, const float *bg, v2 npt, float out[4]) {
    v2 pt = V2(npt.x * s->cw, npt.y * s->ch);
    float acc[4] = {0, 0, 0, 0};
    int nfrag = 0;
    if (bg) memcpy(acc, bg, 16);
    for (int g = 0; g < s->ng; g++) {
        const int32_t *gr = grec(s, g);
        const int32_t *ids = s->topo + s->topo[DVG_H_OFF_GSHAPES] + gr[DVG_G_SHAPES_OFF];
        v2 lp = xform_pt(c2s[g], pt);
        for (int pass = 0; pass < 2; pass++) {           /* stroke fragment first, then fill (:555-582) */
            int type = pass == 0 ? gr[DVG_G_STROKE_TYPE] : gr[DVG_G_FILL_TYPE];
            if (type < 0) continue;
            int hit = 0;
            if (pass == 0) {
                for (int k = 0; k < gr[DVG_G_NUM_SHAPES] && !hit; k++) hit = shape_stroke_hit(s, srec(s, ids[k]), lp);
            } else {
                int w = 0;
                for (int k = 0; k < gr[DVG_G_NUM_SHAPES]; k++) w += shape_winding(s, srec(s, ids[k]), lp);
                hit = gr[DVG_G_EVEN_ODD] ? (abs(w) % 2 == 1) : (w != 0);       /* :82-86 */
            }
            if (!hit) continue;
            float c[4];
            eval_color(type, s->params + (pass == 0 ? gr[DVG_G_STROKE_OFF] : gr[DVG_G_FILL_OFF]),
                       pass == 0 ? gr[DVG_G_STROKE_STOPS] : gr[DVG_G_FILL_STOPS], pt, c);
            float a = c[3];
            acc[0] = acc[0] * (1 - a) + a * c[0];           /* premultiplied "over" (:644-645) */
            acc[1] = acc[1] * (1 - a) + a * c[1];
            acc[2] = acc[2] * (1 - a) + a * c[2];
            acc[3] = acc[3] * (1 - a) + a;
            nfrag++;
        }
    }
    if (nfrag == 0) {
        if (bg) memcpy(out, bg, 16); else memset(out, 0, 16);
        return;
    }
    if (acc[3] > 1e-6f) { float inv = 1.f / acc[3]; acc[0] *= inv; acc[1] *= inv; acc[2] *= inv; }
    memcpy(out, acc, 16);
}

/* filter.h:22-48 */
static float filter_weight(int type, float radius, float dx, float dy) {
    if (fabsf(dx) > radius || fabsf(dy) > radius) return 0;
    if (type == DVG_FILTER_BOX) { float w = 2 * radius; return 1.f / (w * w); }
    if (type == DVG_FILTER_TENT) { float r2 = radius * radius; return (radius - fabsf(dx)) * (radius - fabsf(dy)) / (r2 * r2); }
    if (type == DVG_FILTER_PARABOLIC) {
        float sx = dx / radius, sy = dy / radius;
        return (4.f / 3.f) * (1 - sx * sx) * (4.f / 3.f) * (1 - sy * sy);
    }
    float ndx = (dx / (2 * radius)) + 0.5f, ndy = (dy / (2 * radius)) + 0.5f, two_pi = (float)(2 * M_PI);
    double a = (double)0.5f * ((double)1.f - cos((double)(two_pi * ndx)));
    double b = a * (double)0.5f * ((double)1.f - cos((double)(two_pi * ndy)));
    return (float)(b / (double)(radius * radius));
}

/* diffvg.cpp:1115-1272 (weight_kernel + render_kernel, forward, colour) */
EXPORT int dvgo_render(const int32_t *topo, const float *params, const float *background, float *image, float *sdf,
                       int width, int height, int nsx, int nsy, uint64_t seed, float *d_background,
                       const float *d_render_image, const float *d_render_sdf, float *d_translation,
                       int use_prefiltering, const float *eval_positions, int n_eval, float *d_params, int nthreads) {
    (void)d_background; (void)d_translation; (void)eval_positions; (void)n_eval; (void)d_params; (void)nthreads;
    if (!topo || !params || topo[DVG_H_MAGIC] != DVG_TOPO_MAGIC) return fail("bad scene");
    if (d_render_image || d_render_sdf || sdf || use_prefiltering || !image)
        return fail("oracle/dvg_oracle.c restates the forward colour path only; build oracle/_ref for the rest");
    scene_t s;
    s.topo = topo; s.params = params;
    s.cw = topo[DVG_H_CANVAS_W]; s.ch = topo[DVG_H_CANVAS_H]; s.ns = topo[DVG_H_NUM_SHAPES]; s.ng = topo[DVG_H_NUM_GROUPS];
    for (int i = 0; i < s.ns; i++)
        if (srec(&s, i)[DVG_S_FLAGS] & DVG_SF_DISTANCE_APPROX) return fail("use_distance_approx is not restated");
    float *inv = (float *)malloc(sizeof(float) * 9 * (size_t)s.ng);
    const float **c2s = (const float **)malloc(sizeof(float *) * (size_t)s.ng);
    for (int g = 0; g < s.ng; g++) {                      /* shape.h:122-124 */
        inverse3(params + grec(&s, g)[DVG_G_XFORM_OFF], inv + 9 * g);
        c2s[g] = inv + 9 * g;
    }
    const int ftype = topo[DVG_H_FILTER_TYPE];
    const float radius = params[topo[DVG_H_FILTER_RADIUS_OFF]];
    const int ri = (int)ceilf(radius);
    float *wimg = (float *)calloc((size_t)width * height, sizeof(float));
    const int n = width * height * nsx * nsy;
    for (int pass = 0; pass < 2; pass++) {                /* 0: weights (:1115-1158), 1: colours (:1161-1249) */
        for (int idx = 0; idx < n; idx++) {
            int sx = idx % nsx, sy = (idx / nsx) % nsy, x = (idx / (nsx * nsy)) % width, y = idx / (nsx * nsy * width);
            pcg32 rng;
            pcg_init(&rng, idx, seed);
            float rx = pcg_float(&rng), ry = pcg_float(&rng);
            v2 pt = V2(x + ((float)sx + rx) / nsx, y + ((float)sy + ry) / nsy);
            float color[4] = {0, 0, 0, 0};
            if (pass == 1) {
                v2 npt = pt;
                npt.x /= width; npt.y /= height;
                sample_color(&s, c2s, background ? background + 4 * (y * width + x) : NULL, npt, color);
            }
            for (int dy = -ri; dy <= ri; dy++) {
                for (int dx = -ri; dx <= ri; dx++) {
                    int xx = x + dx, yy = y + dy;
                    if (xx < 0 || xx >= width || yy < 0 || yy >= height) continue;
                    float w = filter_weight(ftype, radius, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
                    if (pass == 0) { wimg[yy * width + xx] += w; continue; }
                    float ws = wimg[yy * width + xx];
                    if (ws > 0) {
                        float inv_ws = 1.f / ws;            /* Vector4 / scalar multiplies by the reciprocal (vector.h) */
                        for (int k = 0; k < 4; k++) image[4 * (yy * width + xx) + k] += (w * color[k]) * inv_ws;
                    }
                }
            }
        }
    }
    free(wimg); free(inv); free((void *)c2s);
    return 0;
}

EXPORT int64_t dvgo_scene_dump(const int32_t *topo, const float *params, int what, int index, uint32_t *out, int64_t cap) {
    (void)topo; (void)params; (void)what; (void)index; (void)out; (void)cap;
    fail("scene dumps (BVH / CDF) are not restated; build oracle/_ref");
    return -1;
}
