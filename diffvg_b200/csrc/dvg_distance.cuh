// dvg_distance.cuh -- closest-point / distance queries and their adjoints: the SDF-prefiltering
// path (sample_color_prefiltered, diffvg.cpp:835-1113) and the SDF output (sample_distance,
// diffvg.cpp:709-775).  Follows compute_distance.h:18-439 (closest_point, compute_distance) and
// 441-949 (d_closest_point, d_compute_distance).
//
// The reference searches three BVH levels with a shrinking radius; the result it returns is
//   per shape : argmin over segments of the LOCAL-space distance, if that minimum is < max_radius
//               (rect: always found, Q7; circle: never found, Q5; ellipse: asserts);
//   per group : argmin over found shapes of the CANVAS-space distance of the shape's closest point.
// Pruning with boxes is conservative, so both minima are independent of the traversal order
// (except on exact float ties, where the winner only matters for gradients and the tied
// candidates -- shared end points of adjacent segments -- produce the same gradient).  The flat
// primitive lists of this implementation therefore give the same answer; they prune with the
// segment's leaf box against the running minimum exactly like compute_distance.h:285-294.
#pragma once
#include "dvg_scene.cuh"
#include "dvg_geom.cuh"
#include "dvg_color.cuh"
#include "dvg_trace.cuh"

namespace dvg {

constexpr int DVG_MAXPF = 64;  // prefilter fragment capacity per sample (diffvg.cpp:847)

// aabb.h:62-67
DVG_HD bool box_within_distance(Box b, F2 p, float r) {
    return p.x >= b.x0 - r && p.x <= b.x1 + r && p.y >= b.y0 - r && p.y <= b.y1 + r;
}

// Closest point of one path segment / rect to `pt` (local space).  Returns the distance the
// reference compares (compute_distance.h:46-267 per leaf type; rect 312-352 returns no distance,
// the caller measures it in canvas space).
DVG_HD float prim_closest(int type, bool approx, F4 p01, F4 p23, F2 pt, F2 &cp, float &t_root) {
    switch (type) {
        case PRIM_LINE: {  // compute_distance.h:49-70
            F2 p0 = mk2(p01.x, p01.y), p1 = mk2(p01.z, p01.w);
            float t = dot2(pt - p0, p1 - p0) / dot2(p1 - p0, p1 - p0);
            if (t < 0) { cp = p0; t_root = 0; return distance2(p0, pt); }
            if (t > 1) { cp = p1; t_root = 1; return distance2(p1, pt); }
            cp = p0 + t * (p1 - p0); t_root = t;
            return distance2(p0 + t * (p1 - p0), pt);
        }
        case PRIM_QUAD: {  // compute_distance.h:71-126
            F2 p0 = mk2(p01.x, p01.y), p1 = mk2(p01.z, p01.w), p2 = mk2(p23.x, p23.y);
            if (approx) {
                cp = quadratic_closest_pt_approx(p0, p1, p2, pt, &t_root);
                return distance2(cp, pt);
            }
            F2 pt0 = eval_quad(p0, p1, p2, 0.f), pt1 = eval_quad(p0, p1, p2, 1.f);
            float dist = distance2(pt0, pt);
            cp = pt0; t_root = 0;
            float dist1 = distance2(pt1, pt);
            if (dist1 < dist) { dist = dist1; cp = pt1; t_root = 1; }
            F2 a2 = p0 - 2 * p1 + p2, a1 = -p0 + p1;
            float A = sum2(a2 * a2);
            float B = sum2(3 * a2 * a1);
            float C = sum2(2 * a1 * a1 + a2 * (p0 - pt));
            float D = sum2(a1 * (p0 - pt));
            float t[3];
            int num_sol = solve_cubic_f(A, B, C, D, t);
            for (int j = 0; j < num_sol; j++) {
                if (t[j] >= 0 && t[j] <= 1) {
                    F2 p = eval_quad(p0, p1, p2, t[j]);
                    float distp = distance2(p, pt);
                    if (distp < dist) { dist = distp; cp = p; t_root = t[j]; }
                }
            }
            return dist;
        }
        case PRIM_CUBIC: {  // compute_distance.h:127-267
            F2 p0 = mk2(p01.x, p01.y), p1 = mk2(p01.z, p01.w), p2 = mk2(p23.x, p23.y), p3 = mk2(p23.z, p23.w);
            F2 pt0 = eval_cubic(p0, p1, p2, p3, 0.f), pt1 = eval_cubic(p0, p1, p2, p3, 1.f);
            float dist = distance2(pt0, pt);
            cp = pt0; t_root = 0;
            float dist1 = distance2(pt1, pt);
            if (dist1 < dist) { dist = dist1; cp = pt1; t_root = 1; }
            Quintic q = cubic_quintic(p0, p1, p2, p3, pt);
            float intervals[4];
            int n = quintic_intervals<true>(q, intervals);
            float lower_bound = 0.f;
            for (int j = 0; j < n + 1; j++) {
                if (j < n && intervals[j] < 0.f) continue;
                float upper_bound = j < n ? rminf(intervals[j], 1.f) : 1.f;
                float t;
                if (quintic_root_in(q, lower_bound, upper_bound, &t)) {
                    F2 p = eval_cubic(p0, p1, p2, p3, t);
                    float distp = distance2(p, pt);
                    if (distp < dist) { dist = distp; cp = p; t_root = t; }
                    if (upper_bound >= 1.f) break;
                    lower_bound = upper_bound;
                }
            }
            return dist;
        }
        case PRIM_RECT: {  // compute_distance.h:312-352; Q7: the interior case stores p0, not p
            F2 lt = mk2(p01.x, p01.y), rt = mk2(p01.z, p01.y), lb = mk2(p01.x, p01.w), rb = mk2(p01.z, p01.w);
            const F2 e0[4] = {lt, lt, rt, lb};
            const F2 e1[4] = {lb, rt, rb, rb};
            float min_dist = 0.f;
            cp = mk2(0, 0);
            for (int k = 0; k < 4; k++) {
                F2 a = e0[k], b = e1[k];
                float t = dot2(pt - a, b - a) / dot2(b - a, b - a);
                float d; F2 c;
                if (t < 0) { d = distance2(a, pt); c = a; }
                else if (t > 1) { d = distance2(b, pt); c = b; }
                else { d = distance2(a + t * (b - a), pt); c = a; }
                if (k == 0 || d < min_dist) { min_dist = d; cp = c; }
            }
            t_root = 0;
            return min_dist;
        }
        default:
            cp = mk2(0, 0); t_root = 0;
            return INFINITY;
    }
}

// Result of a group-level search (compute_distance.h:375-439).
struct DistHit {
    bool found;
    float dist;       // canvas space
    F2 cp;            // canvas space
    int inst;         // shape instance of the closest shape
    int base_id, point_id;
    float t_root;
};

DVG_HD void dist_hit_init(DistHit &h, float max_radius) {
    h.found = false; h.dist = max_radius; h.cp = mk2(0, 0); h.inst = -1; h.base_id = -1; h.point_id = -1; h.t_root = 0.f;
}

// ------------------------------------------------------------------------------------------
// Adjoint of closest_point for one path segment (compute_distance.h:456-793).  Adds into the
// path's point gradients through `sk` (parameter offsets `poff + 2*i`) and into d_pt.
// Gradient arithmetic only has to agree to the 1e-4 relative gradient tolerance, so products are
// grouped freely; the branch structure (t == 0 / t == 1 / interior, the 1e-6 and 1e-10 guards,
// Q8) follows the reference.
template <typename Sink>
DVG_D void d_closest_point_path(const SceneView &sc, const int *srec, int base_id, int point_id, float t_root,
                                F2 pt, F2 d_cp, const Sink &sk, F2 &d_pt) {
    const int np = srec[DVG_S_NUM_POINTS];
    const int poff = srec[DVG_S_PARAM_OFF];
    const float *P = sc.params + poff;
    const int ncp = sc.topo[sc.topo[DVG_H_OFF_NCP] + srec[DVG_S_NCP_OFF] + base_id];
    if (ncp == 0) {
        const int i0 = point_id, i1 = (point_id + 1) % np;
        F2 p0 = mk2(P[2 * i0], P[2 * i0 + 1]), p1 = mk2(P[2 * i1], P[2 * i1 + 1]);
        float t = dot2(pt - p0, p1 - p0) / dot2(p1 - p0, p1 - p0);
        F2 d_p0 = mk2(0, 0), d_p1 = mk2(0, 0);
        if (t < 0) d_p0 = d_cp;
        else if (t > 1) d_p1 = d_cp;
        else { d_p0 = d_cp * (1 - t); d_p1 = d_cp * t; }
        sk.add(poff + 2 * i0, d_p0.x); sk.add(poff + 2 * i0 + 1, d_p0.y);
        sk.add(poff + 2 * i1, d_p1.x); sk.add(poff + 2 * i1 + 1, d_p1.y);
    } else if (ncp == 1) {
        const int i0 = point_id, i1 = point_id + 1, i2 = (point_id + 2) % np;
        F2 p0 = mk2(P[2 * i0], P[2 * i0 + 1]), p1 = mk2(P[2 * i1], P[2 * i1 + 1]), p2 = mk2(P[2 * i2], P[2 * i2 + 1]);
        F2 d_p0 = mk2(0, 0), d_p1 = mk2(0, 0), d_p2 = mk2(0, 0);
        const float t = t_root;
        if (t == 0) d_p0 = d_cp;
        else if (t == 1) d_p2 = d_cp;
        else {
            // Q8: the reference declares fresh d_p0..2 inside this branch (compute_distance.h:531-533), so
            // nothing of it reaches the points; only d_pt is updated.
            F2 a2 = p0 - 2 * p1 + p2, a1 = -p0 + p1;
            float A = sum2(a2 * a2);
            float B = sum2(3 * a2 * a1);
            float C = sum2(2 * a1 * a1 + a2 * (p0 - pt));
            float tt = 1 - t;
            float d_tt = 2 * tt * dot2(d_cp, p0) + 2 * t * dot2(d_cp, p1);
            float d_t = -d_tt + 2 * tt * dot2(d_cp, p1) + 2 * t * dot2(d_cp, p2);
            float poly_deriv_t = 3 * A * t * t + 2 * B * t + C;
            if (fabsf(poly_deriv_t) > 1e-6f) {
                float d_C = -(d_t / poly_deriv_t) * t;
                float d_D = -(d_t / poly_deriv_t);
                d_pt = d_pt + d_C * (-a2) + d_D * (-a1);
            }
        }
        sk.add(poff + 2 * i0, d_p0.x); sk.add(poff + 2 * i0 + 1, d_p0.y);
        sk.add(poff + 2 * i1, d_p1.x); sk.add(poff + 2 * i1 + 1, d_p1.y);
        sk.add(poff + 2 * i2, d_p2.x); sk.add(poff + 2 * i2 + 1, d_p2.y);
    } else {
        const int i0 = point_id, i1 = point_id + 1, i2 = point_id + 2, i3 = (point_id + 3) % np;
        F2 p0 = mk2(P[2 * i0], P[2 * i0 + 1]), p1 = mk2(P[2 * i1], P[2 * i1 + 1]);
        F2 p2 = mk2(P[2 * i2], P[2 * i2 + 1]), p3 = mk2(P[2 * i3], P[2 * i3 + 1]);
        F2 d_p0 = mk2(0, 0), d_p1 = mk2(0, 0), d_p2 = mk2(0, 0), d_p3 = mk2(0, 0);
        const float t = t_root;
        if (t == 0) d_p0 = d_cp;
        else if (t == 1) d_p3 = d_cp;
        else {
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
            B /= A; C /= A; D /= A; E /= A; F /= A;
            float tt = 1 - t;
            float d_tt = 3 * tt * tt * dot2(d_cp, p0) + 6 * tt * t * dot2(d_cp, p1) + 3 * t * t * dot2(d_cp, p2);
            float d_t = -d_tt + 3 * tt * tt * dot2(d_cp, p1) + 6 * tt * t * dot2(d_cp, p2) + 3 * t * t * dot2(d_cp, p3);
            d_p0 = d_cp * (tt * tt * tt);
            d_p1 = d_cp * (3 * tt * tt * t);
            d_p2 = d_cp * (3 * tt * t * t);
            d_p3 = d_cp * (t * t * t);
            const double td = (double)t;
            double poly_deriv_t = 5 * td * td * td * td + 4 * B * td * td * td + 3 * C * td * td + 2 * D * td + E;
            if (fabs(poly_deriv_t) > 1e-10f) {
                const double k = -((double)d_t / poly_deriv_t);
                double dB = k * td * td * td * td, dC = k * td * td * td, dD = k * td * td, dE = k * td, dF = k;
                double dA = -dB * B / A - dC * C / A - dD * D / A - dE * E / A - dF * F / A;
                dB /= A; dC /= A; dD /= A; dE /= A; dF /= A;
                // vector coefficients are float in the reference (Vector2f * double -> the scalar is
                // narrowed by the TVector2 operator templates); keep the sums in float
                const float fA = (float)dA, fB = (float)dB, fC = (float)dC, fD = (float)dD, fE = (float)dE, fF = (float)dF;
                d_p0 = d_p0 + (fA * 3 * (-1) * 2) * q3;
                d_p1 = d_p1 + (fA * 3 * 3 * 2) * q3;
                d_p2 = d_p2 + (fA * 3 * (-3) * 2) * q3;
                d_p3 = d_p3 + (fA * 3 * 1 * 2) * q3;
                d_p0 = d_p0 + (fB * 5) * ((-1) * q2 + 3 * q3);
                d_p1 = d_p1 + (fB * 5) * (3 * q2 + (-6) * q3);
                d_p2 = d_p2 + (fB * 5) * ((-3) * q2 + 3 * q3);
                d_p3 = d_p3 + (fB * 5) * q2;
                d_p0 = d_p0 + (fC * 4) * ((-1) * q1 + (-3) * q3) + (fC * 2) * ((3 * 2) * q2);
                d_p1 = d_p1 + (fC * 4) * (3 * q1 + 3 * q3) + (fC * 2) * ((-6 * 2) * q2);
                d_p2 = d_p2 + (fC * 4) * ((-3) * q1) + (fC * 2) * ((3 * 2) * q2);
                d_p3 = d_p3 + (fC * 4) * q1;
                d_p0 = d_p0 + (fD * 3) * (3 * q1 + (-3) * q2) + (fD * 3) * ((-1) * pp + q3);
                d_p1 = d_p1 + (fD * 3) * ((-6) * q1 + 3 * q2) + (fD * 3) * (3 * pp);
                d_p2 = d_p2 + (fD * 3) * (3 * q1) + (fD * 3) * ((-3) * pp);
                d_pt = d_pt + (fD * 3) * ((-1) * q3);
                d_p0 = d_p0 + fE * ((-3 * 2) * q1) + (fE * 2) * (q2 + 3 * pp);
                d_p1 = d_p1 + fE * ((3 * 2) * q1) + (fE * 2) * ((-6) * pp);
                d_p2 = d_p2 + (fE * 2) * (3 * pp);
                d_pt = d_pt + (fE * 2) * ((-1) * q2);
                d_p0 = d_p0 + fF * q1 + fF * ((-3) * pp);
                d_p1 = d_p1 + fF * (3 * pp);
                d_pt = d_pt + fF * ((-1) * q1);
            }
        }
        sk.add(poff + 2 * i0, d_p0.x); sk.add(poff + 2 * i0 + 1, d_p0.y);
        sk.add(poff + 2 * i1, d_p1.x); sk.add(poff + 2 * i1 + 1, d_p1.y);
        sk.add(poff + 2 * i2, d_p2.x); sk.add(poff + 2 * i2 + 1, d_p2.y);
        sk.add(poff + 2 * i3, d_p3.x); sk.add(poff + 2 * i3 + 1, d_p3.y);
    }
}

// compute_distance.h:795-871
template <typename Sink>
DVG_D void d_closest_point_rect(const float *p, int poff, F2 pt, F2 d_cp, const Sink &sk, F2 &d_pt) {
    F2 lt = mk2(p[0], p[1]), rt = mk2(p[2], p[1]), lb = mk2(p[0], p[3]), rb = mk2(p[2], p[3]);
    const F2 e0[4] = {lt, lt, rt, lb};
    const F2 e1[4] = {lb, rt, rb, rb};
    int min_id = 0;
    float min_dist = 0.f;
    for (int k = 0; k < 4; k++) {
        F2 a = e0[k], b = e1[k];
        float t = dot2(pt - a, b - a) / dot2(b - a, b - a);
        float d = t < 0 ? distance2(a, pt) : (t > 1 ? distance2(b, pt) : distance2(a + t * (b - a), pt));
        if (k == 0 || d < min_dist) { min_dist = d; min_id = k; }
    }
    F2 a = e0[min_id], b = e1[min_id];
    F2 d_a = mk2(0, 0), d_b = mk2(0, 0);
    float t = dot2(pt - a, b - a) / dot2(b - a, b - a);
    if (t < 0) d_a = d_cp;
    else if (t > 1) d_b = d_cp;
    else {
        d_a = d_cp * (1 - t); d_b = d_cp * t;
        float d_t = sum2(d_cp * (b - a));
        float den = dot2(b - a, b - a);
        float d_num = d_t / den, d_den = d_t * (-t) / den;
        d_pt = d_pt + d_num * (b - a);
        d_b = d_b + d_num * (pt - a);
        d_a = d_a + d_num * ((a - b) + (a - pt));
        d_b = d_b + (2 * d_den) * (b - a);
        d_a = d_a + (2 * d_den) * (a - b);
    }
    // corner -> (p_min, p_max) components: lt = (min.x, min.y), rt = (max.x, min.y), lb = (min.x, max.y), rb = max
    F2 d_lt = mk2(0, 0), d_rt = mk2(0, 0), d_lb = mk2(0, 0), d_rb = mk2(0, 0);
    if (min_id == 0) { d_lt = d_a; d_lb = d_b; }
    else if (min_id == 1) { d_lt = d_a; d_rt = d_b; }
    else if (min_id == 2) { d_rt = d_a; d_rb = d_b; }
    else { d_lb = d_a; d_rb = d_b; }
    sk.add(poff + 0, d_lt.x + d_lb.x);
    sk.add(poff + 1, d_lt.y + d_rt.y);
    sk.add(poff + 2, d_rt.x + d_rb.x);
    sk.add(poff + 3, d_lb.y + d_rb.y);
}

// d_compute_distance (compute_distance.h:899-949).  `pt`, `cp` in canvas space.
template <typename Sink>
DVG_D void d_compute_distance(const SceneView &sc, const GroupInfo &g, int inst, F2 pt, F2 cp, int base_id, int point_id,
                              float t_root, float d_dist, const Sink &sk, float *d_translation) {
    if (dist_sq(pt, cp) < 1e-10f) return;  // the derivative at distance 0 is undefined
    const InstInfo &ii = sc.insts[inst];
    const int *srec = sc.topo + sc.topo[DVG_H_OFF_SHAPES] + ii.shape * DVG_SHAPE_REC_LEN;
    F2 local_pt = xform_pt(g.c2s, pt);
    F2 local_cp = xform_pt(g.c2s, cp);
    // d_distance(closest_pt, pt, d_dist, d_closest_pt, d_pt): vector.h:556-564
    F2 v = pt - cp;
    float l = sqrtf(v.x * v.x + v.y * v.y);
    float d_l_sq = 0.5f * d_dist / l;
    F2 dv = (2 * d_l_sq) * v;
    F2 d_cp = -dv, d_pt = dv;
    float d_s2c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    F2 d_local_cp = mk2(0, 0);
    d_xform_pt(g.s2c, local_cp, d_cp, d_s2c, d_local_cp);
    F2 d_local_pt = mk2(0, 0);
    const int type = srec[DVG_S_TYPE];
    if (type == DVG_SHAPE_PATH) {
        d_closest_point_path(sc, srec, base_id, point_id, t_root, local_pt, d_local_cp, sk, d_local_pt);
    } else if (type == DVG_SHAPE_RECT) {
        d_closest_point_rect(sc.params + srec[DVG_S_PARAM_OFF], srec[DVG_S_PARAM_OFF], local_pt, d_local_cp, sk, d_local_pt);
    }  // circle: never found (Q5); ellipse: unsupported
    float d_c2s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    d_xform_pt(g.c2s, pt, d_local_pt, d_c2s, d_pt);
    // d_shape_to_canvas += -c2s^T * d_c2s * c2s^T   (compute_distance.h:942-944)
    float tmp[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            float s = 0.f;
            for (int k = 0; k < 3; k++) s += g.c2s[3 * k + r] * d_c2s[3 * k + c];   // (c2s^T * d_c2s)[r][c]
            tmp[3 * r + c] = s;
        }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            float s = 0.f;
            for (int k = 0; k < 3; k++) s += tmp[3 * r + k] * g.c2s[3 * c + k];     // (... * c2s^T)[r][c]
            sk.add(g.xform_off + 3 * r + c, d_s2c[3 * r + c] - s);
        }
    if (d_translation) {
        DVG_ATOMIC_ADD(d_translation + 0, -d_pt.x);
        DVG_ATOMIC_ADD(d_translation + 1, -d_pt.y);
    }
}

// ------------------------------------------------------------------------------------------
// Per-sample state machine for the prefiltered colour (diffvg.cpp:835-1113): consumes candidate
// primitives in ascending id order like SampleTracer (dvg_trace.cuh).
struct PfFragment {   // what the backward pass needs per fragment
    int key;          // group << 1 | is_stroke
    int inst;         // closest shape instance
    float d;          // signed distance (fill) / distance (stroke)
    F2 cp;
    int base_id, point_id;
    float t_root;
    bool within;
    F4 prev;          // premultiplied accumulation before this fragment
};

// Colour and coverage of one prefiltered fragment (diffvg.cpp:861-932): what end_group composites, and what the cached
// backward kernel (dvg_prefilter.cu) replays from a fragment record -- one function, so both form the same floats.
// `inst`: the closest shape instance (its stroke width enters a stroke fragment's coverage).  Returns the coverage.
DVG_HD float pf_coverage(const SceneView &sc, bool is_stroke, int inst, float d) {
    if (!is_stroke) return smoothstep(d);
    const InstInfo &ii = sc.insts[inst];
    const int *srec = sc.topo + sc.topo[DVG_H_OFF_SHAPES] + ii.shape * DVG_SHAPE_REC_LEN;
    const float sw = srec[DVG_S_WIDTH_OFF] >= 0 ? sc.params[srec[DVG_S_WIDTH_OFF]] : 0.f;
    return smoothstep(fabsf(d) + sw) - smoothstep(fabsf(d) - sw);
}
DVG_HD F4 pf_fragment_color(const SceneView &sc, const GroupInfo &g, bool is_stroke, float w, F2 cpt) {
    F4 c = is_stroke ? eval_color(g.stroke_type, sc.params + g.stroke_off, g.stroke_stops, cpt)
                     : eval_color(g.fill_type, sc.params + g.fill_off, g.fill_stops, cpt);
    c.w *= w;
    return c;
}
DVG_HD void pf_composite(F4 &accum, F4 c) {   // "over", premultiplied (diffvg.cpp:944-955)
    const float oma = 1 - c.w;
    accum.x = accum.x * oma + c.w * c.x;
    accum.y = accum.y * oma + c.w * c.y;
    accum.z = accum.z * oma + c.w * c.z;
    accum.w = accum.w * oma + c.w;
}

// Fragment cache of the prefiltered path: the forward kernel leaves the first DVG_PFC_K fragments of every sample (32 bytes
// each: PfFragment without `prev`, which a replay of the compositing gives back) and the fragment count; the backward pass
// of the same (scene, size, samples, rows) differentiates from them without walking the candidate lists again.  Samples
// with more fragments than that go through the full backward kernel.  Layout: thread t of the launch (warp w = t / 32)
// owns the records [(w * K + slot) * 2 + half][lane] (U4 each: a warp's store is 512 contiguous bytes).
#define DVG_PFC_K 4
struct PfCache {
    U4 *recs;   // null: no cache
    int *count;    // [threads of the launch] fragments emitted (may exceed DVG_PFC_K)
};
DVG_HD void pf_cache_pack(const PfFragment &f, U4 &a, U4 &b) {
    a.x = (unsigned)((f.key << 1) | (f.within ? 1 : 0)); a.y = (unsigned)f.inst; a.z = dvg_float_bits(f.d); a.w = dvg_float_bits(f.cp.x);
    b.x = dvg_float_bits(f.cp.y); b.y = (unsigned)f.base_id; b.z = (unsigned)f.point_id; b.w = dvg_float_bits(f.t_root);
}
DVG_HD void pf_cache_unpack(U4 a, U4 b, PfFragment &f) {
    f.key = (int)(a.x >> 1); f.within = (a.x & 1u) != 0u; f.inst = (int)a.y; f.d = dvg_bits_float(a.z); f.cp.x = dvg_bits_float(a.w);
    f.cp.y = dvg_bits_float(b.x); f.base_id = (int)b.y; f.point_id = (int)b.z; f.t_root = dvg_bits_float(b.w);
}

// RECORD: 0 colour only; 1 fragment records into `frags` (the full backward kernel); 2 the first DVG_PFC_K fragment records
// into the cache slots of this thread (the forward kernel; `crec` may be null)
template <int RECORD>
struct PrefilterTracer {
    F2 cpt;
    bool active;
    PfFragment *frags;
    U4 *crec;             // RECORD == 2: this thread's first cache record (slot stride 64 U4), or null
    F4 accum;
    int nfrag, sp;
    // current group
    int cur_g, cur_inst;
    const GroupInfo *gp;
    F2 lpt;
    bool g_visit, g_fill_ok, s_fill_ok, has_stroke, has_fill, multi;
    int winding, w_shape;
    DistHit gs, gf;          // group-level results: stroke search (radius inf), fill search (radius 1)
    // current shape: local-space running minimum
    bool sh_found; float sh_min; F2 sh_cp; int sh_base, sh_pid; float sh_t; bool sh_rect;
    float sh_radius;

    DVG_HD void init(F2 cpt_, bool active_, F4 first, PfFragment *frags_) {
        cpt = cpt_; active = active_; frags = frags_; crec = nullptr; accum = first; nfrag = 0; sp = 0;
        cur_g = -1; cur_inst = -1; gp = nullptr; lpt = cpt_;
        g_visit = g_fill_ok = s_fill_ok = has_stroke = has_fill = multi = false;
        winding = w_shape = 0;
        dist_hit_init(gs, INFINITY); dist_hit_init(gf, 1.f);
        sh_found = false; sh_min = 0.f; sh_cp = mk2(0, 0); sh_base = sh_pid = -1; sh_t = 0.f; sh_rect = false; sh_radius = 0.f;
    }

    DVG_HD void emit(F4 c, int is_stroke, const DistHit &h, float d) {
        if (RECORD == 1) {
            if (sp < DVG_MAXPF) {
                PfFragment &f = frags[sp];
                f.key = (cur_g << 1) | is_stroke; f.inst = h.inst; f.d = d; f.cp = h.cp;
                f.base_id = h.base_id; f.point_id = h.point_id; f.t_root = h.t_root; f.within = h.found; f.prev = accum;
                sp++;
            }
        }
        if (RECORD == 2) {
            if (crec != nullptr && nfrag < DVG_PFC_K) {
                PfFragment f;
                f.key = (cur_g << 1) | is_stroke; f.inst = h.inst; f.d = d; f.cp = h.cp;
                f.base_id = h.base_id; f.point_id = h.point_id; f.t_root = h.t_root; f.within = h.found;
                U4 a, b;
                pf_cache_pack(f, a, b);
                crec[nfrag * 64] = a; crec[nfrag * 64 + 32] = b;
            }
        }
        pf_composite(accum, c);
        nfrag++;
    }

    DVG_HD void end_shape() {
        if (cur_inst < 0) return;
        if (sh_found) {  // compute_distance.h:402-417
            const F2 ccp = (gp->flags & DVG_GF_IDENTITY) ? sh_cp : xform_pt(gp->s2c, sh_cp);
            const float dist = distance2(ccp, cpt);
            if (has_stroke && (!gs.found || dist < gs.dist)) {
                gs.found = true; gs.dist = dist; gs.cp = ccp; gs.inst = cur_inst; gs.base_id = sh_base; gs.point_id = sh_pid; gs.t_root = sh_t;
            }
            // the fill search runs with max_radius 1: a path is found only if its local minimum is < 1
            if (has_fill && (sh_rect || sh_min < 1.f) && (!gf.found || dist < gf.dist)) {
                gf.found = true; gf.dist = dist; gf.cp = ccp; gf.inst = cur_inst; gf.base_id = sh_base; gf.point_id = sh_pid; gf.t_root = sh_t;
            }
        }
        winding += w_shape; w_shape = 0;
        sh_found = false;
    }

    DVG_HD void end_group(const SceneView &sc) {
        if (cur_g < 0) return;
        end_shape();
        if (!g_visit) return;
        if (has_stroke && gs.found) {  // diffvg.cpp:861-890
            const InstInfo &ii = sc.insts[gs.inst];
            const int *srec = sc.topo + sc.topo[DVG_H_OFF_SHAPES] + ii.shape * DVG_SHAPE_REC_LEN;
            const float sw = srec[DVG_S_WIDTH_OFF] >= 0 ? sc.params[srec[DVG_S_WIDTH_OFF]] : 0.f;
            const float d = gs.dist;
            const float w = smoothstep(fabsf(d) + sw) - smoothstep(fabsf(d) - sw);   // (== pf_coverage(sc, true, gs.inst, d))
            if (w > 0) {
                DistHit h = gs; h.found = true;
                emit(pf_fragment_color(sc, *gp, true, w, cpt), 1, h, d);
            }
        }
        if (has_fill) {  // diffvg.cpp:891-932
            const int aw = winding < 0 ? -winding : winding;
            const bool inside = g_fill_ok && ((gp->flags & DVG_GF_EVEN_ODD) ? (aw % 2 == 1) : (winding != 0));
            if (gf.found || inside) {
                float d = gf.dist;   // == 1.f (the search radius) when nothing was found
                if (!inside) d = -d;
                const float w = smoothstep(d);
                if (w > 0) emit(pf_fragment_color(sc, *gp, false, w, cpt), 0, gf, d);
            }
        }
    }

    DVG_HD void begin_group(const SceneView &sc, int g) {
        cur_g = g; cur_inst = -1;
        gp = &sc.groups[g];
        winding = 0; w_shape = 0;
        has_stroke = gp->stroke_type >= 0; has_fill = gp->fill_type >= 0;
        multi = gp->num_shapes > 1;
        dist_hit_init(gs, INFINITY); dist_hit_init(gf, 1.f);
        g_visit = active && (sc.num_groups == 1 || box_inside_r(gp->scene_box, cpt, gp->scene_r));  // diffvg.cpp:934-942
        lpt = (gp->flags & DVG_GF_IDENTITY) ? cpt : xform_pt(gp->c2s, cpt);
        g_fill_ok = g_visit && has_fill && box_inside(gp->local_box, lpt);  // is_inside bbox reject, diffvg.cpp:42-45
    }

    DVG_HD void begin_shape(const SceneView &sc, int inst) {
        cur_inst = inst;
        const InstInfo &ii = sc.insts[inst];
        s_fill_ok = g_fill_ok && (!multi || box_inside(ii.box, lpt));
        sh_found = false; sh_rect = false;
        sh_radius = has_stroke ? INFINITY : 1.f;   // the wider of the two searches this group needs
        sh_min = sh_radius;
        sh_base = sh_pid = -1; sh_t = 0.f; sh_cp = mk2(0, 0);
    }

    // WORDS: the winding contribution of this candidate was answered by the winding pre-pass (dvg_wave.cu: classify ->
    // k_wave_solve_fill) and arrives as `nib`; it is zero wherever the gating tests below would not have asked for it.
    template <bool WORDS = false>
    DVG_HD void step(const SceneView &sc, const PrimRef &pr, int nib = 0) {
        if (pr.group != cur_g) { end_group(sc); begin_group(sc, pr.group); }
        if (pr.inst != cur_inst) { end_shape(); begin_shape(sc, pr.inst); }
        if (!g_visit) return;
        const int tf = pr.tf;
        const int type = tf & DVG_PF_TYPE_MASK;
        if (type <= PRIM_CUBIC) {
            // path-BVH pruning against the running minimum (compute_distance.h:285-294); a lone leaf is the root
            if ((tf & DVG_PF_SINGLE) || box_within_distance(pr.box, lpt, sh_min)) {
                // fill-only group: the search radius is 1 (sh_min <= 1), and the polyline bracket of the segment (built for
                // radius 1, dvg_buildfn.cuh) may prove that every point of the curve is farther than that.  The reference's
                // answer is the distance to SOME point of the curve (a root it found, or an end point), so it could not
                // pass `dist < sh_min` either: the solve is skipped, nothing changes.
                const bool far = !has_stroke && pr.cap != nullptr && type >= PRIM_QUAD && capsule_reject(pr.cap, lpt);
                if (!far) {
                    F2 cp; float t_root;
                    const float dist = prim_closest(type, (tf & DVG_PF_APPROX) != 0, pr.p01, pr.p23, lpt, cp, t_root);
                    if (dist < sh_min) {
                        sh_min = dist; sh_cp = cp; sh_t = t_root; sh_found = true;
                        sh_base = pr.base_id; sh_pid = pr.point_id;
                    }
                }
            }
        } else if (type == PRIM_RECT) {
            F2 cp; float t_root;
            prim_closest(type, false, pr.p01, pr.p23, lpt, cp, t_root);
            sh_cp = cp; sh_found = true; sh_rect = true; sh_base = -1; sh_pid = -1; sh_t = 0.f;
        }
        if (WORDS) {
            w_shape += nib;
        } else if (s_fill_ok) {
            if ((tf & DVG_PF_SINGLE) || box_ray_intersect(pr.box, lpt)) w_shape += prim_winding(type, pr.p01, pr.p23, lpt);
        }
    }

    DVG_HD void finish(const SceneView &sc) { end_group(sc); cur_g = -1; }

    DVG_HD F4 resolve(const float *bg_px) const {
        if (nfrag <= 0) {
            if (bg_px) return mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]);
            return mk4(0, 0, 0, 0);
        }
        F4 c = accum;
        if (c.w > 1e-6f) {
            const float inv = 1.f / c.w;
            c.x *= inv; c.y *= inv; c.z *= inv;
        }
        return c;
    }
};

// Backward of sample_color_prefiltered for one sample (diffvg.cpp:985-1111).  d_color is the
// gathered image gradient; returns the gradient w.r.t. the background in d_bg (when there were fragments).
template <typename Sink>
DVG_D void prefilter_backward(const SceneView &sc, const PrefilterTracer<true> &tr, F4 color, F4 d_color, const Sink &sk,
                              float *d_translation, F4 &d_bg) {
    float dcr = d_color.x, dcg = d_color.y, dcb = d_color.z, dca = d_color.w;
    if (tr.accum.w > 1e-6f) {
        const float inv = 1.f / tr.accum.w;
        dca -= (d_color.x * color.x + d_color.y * color.y + d_color.z * color.z) / tr.accum.w;
        dcr = d_color.x * inv; dcg = d_color.y * inv; dcb = d_color.z * inv;
    }
    for (int i = tr.sp - 1; i >= 0; i--) {
        const PfFragment &f = tr.frags[i];
        const int g = f.key >> 1;
        const bool is_stroke = (f.key & 1) != 0;
        const GroupInfo &gi = sc.groups[g];
        const int ctype = is_stroke ? gi.stroke_type : gi.fill_type;
        const int coff = is_stroke ? gi.stroke_off : gi.fill_off;
        const int cstops = is_stroke ? gi.stroke_stops : gi.fill_stops;
        const InstInfo *ii = f.inst >= 0 ? &sc.insts[f.inst] : nullptr;
        const int *srec = ii ? sc.topo + sc.topo[DVG_H_OFF_SHAPES] + ii->shape * DVG_SHAPE_REC_LEN : nullptr;
        const float sw = (srec && srec[DVG_S_WIDTH_OFF] >= 0) ? sc.params[srec[DVG_S_WIDTH_OFF]] : 0.f;
        const float d = f.d;
        float w, apw = 0.f, amw = 0.f;
        if (is_stroke) { apw = fabsf(d) + sw; amw = fabsf(d) - sw; w = smoothstep(apw) - smoothstep(amw); }
        else w = smoothstep(d);
        const F4 base = eval_color(ctype, sc.params + coff, cstops, tr.cpt);
        const float alpha = base.w * w;   // fragments[i].alpha
        const F4 prev = f.prev;
        const float d_prev_alpha = dca * (1.f - alpha);
        float d_alpha_i = dca * (1.f - prev.w);
        d_alpha_i += (dcr * (base.x - prev.x) + dcg * (base.y - prev.y)) + dcb * (base.z - prev.z);
        const F4 dci = mk4(dcr * alpha, dcg * alpha, dcb * alpha, 0.f);
        if (w != 0) {
            const float d_w = w > 0 ? (alpha / w) * d_alpha_i : 0.f;
            d_alpha_i *= w;
            const F4 dc = mk4(dci.x, dci.y, dci.z, d_alpha_i);
            if (ctype == 0) add4(sk, coff, dc);
            // Q4: gradient STROKE colours have no gradient storage in the reference (it faults there: scene.cpp:866-889);
            // accumulated like a fill's
            else d_eval_gradient(ctype, sc.params + coff, cstops, tr.cpt, dc, sk, coff, d_translation);
            if (is_stroke) {
                const float d_apw = d_smoothstep(apw, d_w);
                const float d_amw = -d_smoothstep(amw, d_w);
                float d_d = d_apw + d_amw;
                if (d < 0) d_d = -d_d;
                const float d_sw = d_apw - d_amw;
                if (fabsf(d_d) > 1e-10f)
                    d_compute_distance(sc, gi, f.inst, tr.cpt, f.cp, f.base_id, f.point_id, f.t_root, d_d, sk, d_translation);
                if (srec && srec[DVG_S_WIDTH_OFF] >= 0) sk.add(srec[DVG_S_WIDTH_OFF], d_sw);
            } else {
                float d_d = d_smoothstep(d, d_w);
                if (d < 0) d_d = -d_d;
                if (fabsf(d_d) > 1e-10f && f.within)
                    d_compute_distance(sc, gi, f.inst, tr.cpt, f.cp, f.base_id, f.point_id, f.t_root, d_d, sk, d_translation);
            }
        }
        dcr = dcr * (1 - alpha); dcg = dcg * (1 - alpha); dcb = dcb * (1 - alpha);
        dca = d_prev_alpha;
    }
    d_bg = mk4(dcr, dcg, dcb, dca);
}

// ------------------------------------------------------------------------------------------
// is_inside(scene, group, pt, nullptr) (diffvg.cpp:33-87) as a plain loop over the group's
// primitives; used by the SDF output where no candidate list exists.
DVG_HD bool group_is_inside(const SceneView &sc, int g, F2 cpt) {
    const GroupInfo &gi = sc.groups[g];
    const F2 lpt = (gi.flags & DVG_GF_IDENTITY) ? cpt : xform_pt(gi.c2s, cpt);
    if (!box_inside(gi.local_box, lpt)) return false;
    const bool multi = gi.num_shapes > 1;
    int winding = 0;
    for (int k = 0; k < gi.num_shapes; k++) {
        const int inst = gi.inst_begin + k;
        const InstInfo &ii = sc.insts[inst];
        if (multi && !box_inside(ii.box, lpt)) continue;
        const int e1 = (k + 1 < gi.num_shapes) ? sc.insts[inst + 1].prim_begin : gi.prim_end;
        for (int e = ii.prim_begin; e < e1; e++) {
            const int tf = sc.prim_meta[e].type_flags;
            if ((tf & DVG_PF_SINGLE) || box_ray_intersect(sc.prim_box[e], lpt))
                winding += prim_winding(tf & DVG_PF_TYPE_MASK, sc.prim_p01[e], sc.prim_p23[e], lpt);
        }
    }
    const int aw = winding < 0 ? -winding : winding;
    return (gi.flags & DVG_GF_EVEN_ODD) ? (aw % 2 == 1) : (winding != 0);
}

// compute_distance(scene, group, pt, infinity, ...) as a plain loop (SDF output).
DVG_HD void group_distance(const SceneView &sc, int g, F2 cpt, DistHit &out) {
    const GroupInfo &gi = sc.groups[g];
    const F2 lpt = (gi.flags & DVG_GF_IDENTITY) ? cpt : xform_pt(gi.c2s, cpt);
    dist_hit_init(out, INFINITY);
    for (int k = 0; k < gi.num_shapes; k++) {
        const int inst = gi.inst_begin + k;
        const InstInfo &ii = sc.insts[inst];
        const int e1 = (k + 1 < gi.num_shapes) ? sc.insts[inst + 1].prim_begin : gi.prim_end;
        bool found = false; float mn = INFINITY; F2 mcp = mk2(0, 0); int mb = -1, mp = -1; float mt = 0.f;
        for (int e = ii.prim_begin; e < e1; e++) {
            const PrimMeta pm = sc.prim_meta[e];
            const int type = pm.type_flags & DVG_PF_TYPE_MASK;
            if (type <= PRIM_CUBIC) {
                if ((pm.type_flags & DVG_PF_SINGLE) || box_within_distance(sc.prim_box[e], lpt, mn)) {
                    F2 cp; float tr;
                    const float dist = prim_closest(type, (pm.type_flags & DVG_PF_APPROX) != 0, sc.prim_p01[e], sc.prim_p23[e], lpt, cp, tr);
                    if (dist < mn) { mn = dist; mcp = cp; mt = tr; mb = pm.base_id; mp = pm.point_id; found = true; }
                }
            } else if (type == PRIM_RECT) {
                F2 cp; float tr;
                prim_closest(type, false, sc.prim_p01[e], sc.prim_p23[e], lpt, cp, tr);
                mcp = cp; found = true; mb = mp = -1; mt = 0.f;
            }
        }
        if (found) {
            const F2 ccp = (gi.flags & DVG_GF_IDENTITY) ? mcp : xform_pt(gi.s2c, mcp);
            const float dist = distance2(ccp, cpt);
            if (!out.found || dist < out.dist) {
                out.found = true; out.dist = dist; out.cp = ccp; out.inst = inst; out.base_id = mb; out.point_id = mp; out.t_root = mt;
            }
        }
    }
}

}  // namespace dvg
