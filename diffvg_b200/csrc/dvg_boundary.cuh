// dvg_boundary.cuh -- boundary (edge) sampling: pick a shape by length CDF, map t in [0,1) to a
// point / normal / pdf on its outline, and the Reynolds-transport scatter weights.
// Follows cdf.h:5-29, sample_boundary.h:20-453, diffvg.cpp:1325-1386 (sample_boundary_kernel)
// and diffvg.cpp:89-274 (accumulate_boundary_gradient).
#pragma once
#include "dvg_scene.cuh"
#include "dvg_geom.cuh"
#include "dvg_color.cuh"

namespace dvg {

// cdf.h:5-29
DVG_HD int cdf_sample(const float *cdf, int num_entries, float u, float *updated_u) {
    int lb = 0;
    int len = num_entries - 1 - lb;
    while (len > 0) {
        int half_len = len / 2;
        int mid = lb + half_len;
        if (u < cdf[mid]) {
            len = half_len;
        } else {
            lb = mid + 1;
            len = len - half_len - 1;
        }
    }
    lb = clampi(lb, 0, num_entries - 1);
    if (updated_u) {
        if (lb > 0) *updated_u = (u - cdf[lb - 1]) / (cdf[lb] - cdf[lb - 1]);
        else *updated_u = u / cdf[lb];
    }
    return lb;
}

// What render_edge_kernel needs from one boundary sample (diffvg.cpp:1314-1323), 48 bytes.
struct BoundarySample {
    F2 pt;        // normalised [0,1) canvas position
    F2 local_pt;  // shape-local position
    F2 normal;    // canvas space
    float t;      // the *un-remapped* second random number (diffvg.cpp:1373)
    float pdf;
    float path_t; // BoundaryData.path.t
    int inst;     // shape instance (-> group, shape); -1 = invalid
    int base_point_id;
    int point_id_stroke;  // point_id | (is_stroke << 31)
};

// cos/sin of 2*pi*t as the reference computes them: float argument, double ::cos/::sin,
// The same search started from a GUIDE table: guide[k] = the answer for u = k / DVG_CDF_GUIDE (k = 0 .. DVG_CDF_GUIDE),
// so u in [k / G, (k + 1) / G) lies in [guide[k], guide[k + 1]] -- the cdf is non-decreasing -- and the bisection runs over
// that handful of entries instead of all of them (11 dependent loads at 2048 shapes: a fifth of the sample-generation
// kernel).  Same result as cdf_sample for every u in [0, 1).
#define DVG_CDF_GUIDE 2048
DVG_HD int cdf_sample_guided(const float *cdf, int num_entries, float u, const int *guide) {
    int k = (int)(u * (float)DVG_CDF_GUIDE);
    k = k < 0 ? 0 : (k > DVG_CDF_GUIDE - 1 ? DVG_CDF_GUIDE - 1 : k);
    int lb = guide[k];
    int len = guide[k + 1] - lb;
    while (len > 0) {
        int half_len = len / 2;
        int mid = lb + half_len;
        if (u < cdf[mid]) {
            len = half_len;
        } else {
            lb = mid + 1;
            len = len - half_len - 1;
        }
    }
    return clampi(lb, 0, num_entries - 1);
}
// multiplied by a float radius in double, rounded on store (sample_boundary.h:31-34).
DVG_HD_NOINLINE F2 circle_offset(float radius, float t) {
    float arg = 2 * (float)DVG_PI_D * t;
#if defined(__CUDA_ARCH__)
    double sn, cs;
    sincos((double)arg, &sn, &cs);   // one argument reduction for both
    return mk2((float)((double)radius * cs), (float)((double)radius * sn));
#else
    return mk2((float)((double)radius * cos((double)arg)), (float)((double)radius * sin((double)arg)));
#endif
}

// sample_boundary.h:80-307.  `pts` = path points, `thick` = per-point thickness or null.
DVG_HD F2 sample_boundary_path(const float *pts, const float *thick, const int *ncp, int num_points, int num_segs,
                               bool is_closed, const float *seg_cdf, const float *seg_pmf, const int *seg_point_id,
                               float path_length, float t, F2 &normal, float &pdf, int &base_point_id, int &point_id,
                               float &path_t, float dir, float stroke_radius) {
    if (dir != 0.f && !is_closed) {
        float cap_length;
        if (thick) cap_length = (float)DVG_PI_D * (thick[0] + thick[num_points - 1]);
        else cap_length = 2 * (float)DVG_PI_D * stroke_radius;
        float cap_prob = cap_length / (cap_length + path_length);
        if (t < cap_prob) {
            t = t / cap_prob;
            pdf *= cap_prob;
            float r0 = stroke_radius, r1 = stroke_radius;
            if (thick) { r0 = thick[0]; r1 = thick[num_points - 1]; }
            if (dir < 0) {
                F2 p0 = mk2(pts[0], pts[1]);
                F2 offset = circle_offset(r0, t);
                normal = normalize2(offset);
                pdf /= (2 * (float)DVG_PI_D * r0);
                base_point_id = 0; point_id = 0; path_t = 0;
                return p0 + offset;
            } else {
                F2 p0 = mk2(pts[2 * (num_points - 1)], pts[2 * (num_points - 1) + 1]);
                F2 offset = circle_offset(r1, t);
                normal = normalize2(offset);
                pdf /= (2 * (float)DVG_PI_D * r1);
                base_point_id = num_segs - 1;
                point_id = num_points - 2 - ncp[base_point_id];
                path_t = 1;
                return p0 + offset;
            }
        } else {
            t = (t - cap_prob) / (1 - cap_prob);
            pdf *= (1 - cap_prob);
        }
    }
    int sample_id = cdf_sample(seg_cdf, num_segs, t, &t);
    int pid = seg_point_id[sample_id];
    base_point_id = sample_id; point_id = pid; path_t = t;
    int n = ncp[sample_id];
    if (t < -1e-3f || t > 1 + 1e-3f) { pdf = 0; return mk2(0, 0); }
    if (n == 0) {
        int i0 = pid, i1 = (i0 + 1) % num_points;
        F2 p0 = mk2(pts[2 * i0], pts[2 * i0 + 1]), p1 = mk2(pts[2 * i1], pts[2 * i1 + 1]);
        F2 tangent = p1 - p0;
        float tan_len = length2(tangent);
        if (tan_len == 0) { pdf = 0; return mk2(0, 0); }
        normal = mk2(-tangent.y, tangent.x) / tan_len;
        pdf *= seg_pmf[sample_id] / tan_len;
        F2 ret = p0 + t * (p1 - p0);
        if (dir != 0.f) {
            float r0 = stroke_radius, r1 = stroke_radius;
            if (thick) { r0 = thick[i0]; r1 = thick[i1]; }
            float r = r0 + t * (r1 - r0);
            ret = ret + (dir * r) * normal;
            if (dir < 0) normal = -normal;
        }
        return ret;
    } else if (n == 1) {
        int i0 = pid, i1 = i0 + 1, i2 = (i0 + 2) % num_points;
        F2 p0 = mk2(pts[2 * i0], pts[2 * i0 + 1]), p1 = mk2(pts[2 * i1], pts[2 * i1 + 1]), p2 = mk2(pts[2 * i2], pts[2 * i2 + 1]);
        F2 tangent = (2 * (1 - t)) * (p1 - p0) + (2 * t) * (p2 - p1);
        float tan_len = length2(tangent);
        if (tan_len == 0) { pdf = 0; return mk2(0, 0); }
        normal = mk2(-tangent.y, tangent.x) / tan_len;
        pdf *= seg_pmf[sample_id] / tan_len;
        F2 ret = eval_quad(p0, p1, p2, t);
        if (dir != 0.f) {
            float r0 = stroke_radius, r1 = stroke_radius, r2 = stroke_radius;
            if (thick) { r0 = thick[i0]; r1 = thick[i1]; r2 = thick[i2]; }
            float tt = 1 - t;
            float r = (tt * tt) * r0 + (2 * tt * t) * r1 + (t * t) * r2;
            ret = ret + (dir * r) * normal;
            if (dir < 0) normal = -normal;
        }
        return ret;
    } else {
        int i0 = pid, i1 = pid + 1, i2 = pid + 2, i3 = (pid + 3) % num_points;
        F2 p0 = mk2(pts[2 * i0], pts[2 * i0 + 1]), p1 = mk2(pts[2 * i1], pts[2 * i1 + 1]);
        F2 p2 = mk2(pts[2 * i2], pts[2 * i2 + 1]), p3 = mk2(pts[2 * i3], pts[2 * i3 + 1]);
        float omt = 1 - t;
        F2 tangent = (3 * (omt * omt)) * (p1 - p0) + (6 * omt * t) * (p2 - p1) + (3 * t * t) * (p3 - p2);
        float tan_len = length2(tangent);
        if (tan_len == 0) { pdf = 0; return mk2(0, 0); }
        normal = mk2(-tangent.y, tangent.x) / tan_len;
        pdf *= seg_pmf[sample_id] / tan_len;
        F2 ret = eval_cubic(p0, p1, p2, p3, t);
        if (dir != 0.f) {
            float r0 = stroke_radius, r1 = stroke_radius, r2 = stroke_radius, r3 = stroke_radius;
            if (thick) { r0 = thick[i0]; r1 = thick[i1]; r2 = thick[i2]; r3 = thick[i3]; }
            float tt = 1 - t;
            float r = (tt * tt * tt) * r0 + (3 * tt * tt * t) * r1 + (3 * tt * t * t) * r2 + (t * t * t) * r3;
            ret = ret + (dir * r) * normal;
            if (dir < 0) normal = -normal;
        }
        return ret;
    }
}

// Apply the stroke offset shared by circle/ellipse/rect (sample_boundary.h:38-44 etc.)
DVG_HD F2 stroke_offset(F2 ret, F2 &normal, float dir, float stroke_radius) {
    if (dir != 0.f) {
        ret = ret + (dir * stroke_radius) * normal;
        if (dir < 0) normal = -normal;
    }
    return ret;
}

// diffvg.cpp:1325-1386 + sample_boundary.h:386-453.  Fills `bs`; bs.inst = -1 when invalid.
// `scene`: which scene of a batch (SceneView): its CDF / instance / shape / segment tables and parameters; `idx` and `seed`
// are that scene's own.  bs.inst comes back batch-wide.
DVG_HD void make_boundary_sample(const SceneView &sc, int idx, uint64_t seed, BoundarySample &bs, const float *shape_cdf = nullptr,
                                 int scene = 0, const int *guide = nullptr) {
    bs.inst = -1;
    bs.pt = mk2(0, 0);
    Pcg32 rng = pcg32_init(idx, seed);
    float u = pcg32_next_float(rng);
    const int inst_base = scene * sc.num_insts;
    int sample_id = (guide && shape_cdf) ? cdf_sample_guided(shape_cdf, sc.num_insts, u, guide)
                                         : cdf_sample(shape_cdf ? shape_cdf : sc.shape_cdf + inst_base, sc.num_insts, u, nullptr);   // (a staged copy of the same table)
    const InstInfo ii = sc.insts[inst_base + sample_id];
    int shape_id = ii.shape;
    // Q11: the pmf is looked up by *shape id*, not by sample id (diffvg.cpp:1343).
    // shape_id < num_shapes <= num_insts is not guaranteed by the reference either; clamp the read.
    float shape_pmf = sc.shape_pmf[inst_base + (shape_id < sc.num_insts ? shape_id : sc.num_insts - 1)];
    if (shape_pmf <= 0) return;
    float t = pcg32_next_float(rng);
    const float t_orig = t;
    const GroupInfo &g = sc.groups[ii.group];
    const int *srec = sc.topo + sc.topo[DVG_H_OFF_SHAPES] + shape_id * DVG_SHAPE_REC_LEN;
    const float *params = sc.params + (size_t)scene * sc.num_params;   // this scene's parameters
    float stroke_width = srec[DVG_S_WIDTH_OFF] >= 0 ? params[srec[DVG_S_WIDTH_OFF]] : 0.f;
    float pdf = 1;
    bool stroke_perturb = false;
    bool has_fill = g.fill_type >= 0, has_stroke = g.stroke_type >= 0;
    if (has_fill && has_stroke) {
        if (t < 0.5f) { stroke_perturb = false; t = 2 * t; pdf = 0.5f; }
        else { stroke_perturb = true; t = 2 * (t - 0.5f); pdf = 0.5f; }
    } else if (has_stroke) {
        stroke_perturb = true;
    }
    float dir = 0.f;
    if (stroke_perturb) {
        if (t < 0.5f) { dir = -1.f; t = 2 * t; pdf *= 0.5f; }
        else { dir = 1.f; t = 2 * (t - 0.5f); pdf *= 0.5f; }
    }
    F2 normal = mk2(0, 0);
    F2 local;
    int base_point_id = 0, point_id = 0;
    float path_t = 0;
    const float *p = params + srec[DVG_S_PARAM_OFF];
    const float two_pi = 2 * (float)DVG_PI_D;
    switch (srec[DVG_S_TYPE]) {
        case DVG_SHAPE_CIRCLE: {  // sample_boundary.h:20-46
            F2 offset = circle_offset(p[0], t);
            normal = normalize2(offset);
            pdf /= (two_pi * p[0]);
            local = stroke_offset(mk2(p[1], p[2]) + offset, normal, dir, stroke_width);
            break;
        }
        case DVG_SHAPE_ELLIPSE: {  // sample_boundary.h:48-78
            float arg = two_pi * t;
            double c = cos((double)arg), s = sin((double)arg);
            F2 offset = mk2((float)((double)p[0] * c), (float)((double)p[1] * s));
            float dxdt = (float)((double)(-p[0]) * s * (double)2 * (double)(float)DVG_PI_D);
            float dydt = (float)((double)p[1] * c * (double)2 * (double)(float)DVG_PI_D);
            normal = normalize2(mk2(dydt, -dxdt));
            pdf /= sqrtf(dxdt * dxdt + dydt * dydt);
            local = stroke_offset(mk2(p[2], p[3]) + offset, normal, dir, stroke_width);
            break;
        }
        case DVG_SHAPE_PATH: {
            const float *thick = srec[DVG_S_THICK_OFF] >= 0 ? params + srec[DVG_S_THICK_OFF] : nullptr;
            const int *ncp = sc.topo + sc.topo[DVG_H_OFF_NCP] + srec[DVG_S_NCP_OFF];
            int so = scene * sc.total_segs + srec[DVG_S_NCP_OFF];
            local = sample_boundary_path(p, thick, ncp, srec[DVG_S_NUM_POINTS], srec[DVG_S_NUM_SEGS],
                                         (srec[DVG_S_FLAGS] & DVG_SF_CLOSED) != 0, sc.seg_cdf + so, sc.seg_pmf + so,
                                         sc.seg_point_id + so, sc.shapes_length[scene * sc.num_shapes + shape_id], t, normal, pdf,
                                         base_point_id, point_id, path_t, dir, stroke_width);
            break;
        }
        default: {  // rect, sample_boundary.h:309-384
            F2 pmin = mk2(p[0], p[1]), pmax = mk2(p[2], p[3]);
            float w = pmax.x - pmin.x, h = pmax.y - pmin.y;
            pdf /= (2 * (w + h));
            if (t <= w / (w + h)) {
                t *= (w + h) / w;
                if (t < 0.5f) { normal = mk2(0, -1); local = pmin + (2 * t) * mk2(pmax.x - pmin.x, 0.f); }
                else { normal = mk2(0, 1); local = mk2(pmin.x, pmax.y) + (2 * (t - 0.5f)) * mk2(pmax.x - pmin.x, 0.f); }
            } else {
                t = (t - w / (w + h)) * (w + h) / h;
                if (t < 0.5f) { normal = mk2(-1, 0); local = pmin + (2 * t) * mk2(0.f, pmax.y - pmin.y); }
                else { normal = mk2(1, 0); local = mk2(pmax.x, pmin.y) + (2 * (t - 0.5f)) * mk2(0.f, pmax.y - pmin.y); }
            }
            local = stroke_offset(local, normal, dir, stroke_width);
            break;
        }
    }
    if (pdf <= 0) return;
    // (an exactly-identity transform maps the point and the normal onto themselves bit for bit: skip the 18 loads)
    const bool ident = (g.flags & DVG_GF_IDENTITY) != 0;
    F2 bpt = ident ? local : xform_pt(g.s2c, local);
    normal = ident ? normalize2(normal) : xform_normal(g.c2s, normal);
    bpt.x /= sc.canvas_w;
    bpt.y /= sc.canvas_h;
    bs.pt = bpt;
    bs.local_pt = local;
    bs.normal = normal;
    bs.t = t_orig;
    bs.pdf = shape_pmf * pdf;
    bs.path_t = path_t;
    bs.inst = inst_base + sample_id;
    bs.base_point_id = base_point_id;
    bs.point_id_stroke = point_id | (stroke_perturb ? (int)0x80000000 : 0);
}

// gather_d_color (diffvg.cpp:779-815)
DVG_HD F4 gather_d_color(const Filter &f, const float *d_img, const float *wimg, int width, int height, F2 pt) {
    const int x = (int)pt.x, y = (int)pt.y;
    const int ri = (int)ceilf(f.radius);
    F4 d = mk4(0, 0, 0, 0);
    for (int dy = -ri; dy <= ri; dy++) {
        for (int dx = -ri; dx <= ri; dx++) {
            int xx = x + dx, yy = y + dy;
            if (xx >= 0 && xx < width && yy >= 0 && yy < height) {
                float fw = filter_weight(f, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
                float ws = wimg[yy * width + xx];
                if (ws > 0 && fw != 0.f) {
                    const float *px = d_img + 4 * (yy * width + xx);
                    d = d + (fw / ws) * mk4(px[0], px[1], px[2], px[3]);
                }
            }
        }
    }
    return d;
}


// accumulate_boundary_gradient (diffvg.cpp:89-274), scattering into the flat d_params.
template <typename Sink>
DVG_D void accumulate_boundary_gradient(const SceneView &sc, const RenderArgs &ra, const BoundarySample &bs,
                                        const InstInfo &ii, const GroupInfo &g, float contrib, F2 normal, const Sink &sk) {
    const int *srec = sc.topo + sc.topo[DVG_H_OFF_SHAPES] + ii.shape * DVG_SHAPE_REC_LEN;
    const bool is_stroke = bs.point_id_stroke < 0;
    const int point_id = bs.point_id_stroke & 0x7fffffff;
    const int type = srec[DVG_S_TYPE];
    const int poff = srec[DVG_S_PARAM_OFF];
    float w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0, nw = 0;
    if (type == DVG_SHAPE_PATH) {
        const int np = srec[DVG_S_NUM_POINTS];
        const int ncp = sc.topo[sc.topo[DVG_H_OFF_NCP] + srec[DVG_S_NCP_OFF] + bs.base_point_id];
        const float t = bs.path_t;
        if (ncp == 0) {
            nw = 2; i0 = point_id; i1 = (point_id + 1) % np;
            w0 = 1 - t; w1 = t;
        } else if (ncp == 1) {
            nw = 3; i0 = point_id; i1 = point_id + 1; i2 = (point_id + 2) % np;
            w0 = (1 - t) * (1 - t); w1 = 2 * (1 - t) * t; w2 = t * t;
        } else {
            nw = 4; i0 = point_id; i1 = point_id + 1; i2 = point_id + 2; i3 = (point_id + 3) % np;
            const float omt = 1 - t;
            w0 = omt * omt * omt; w1 = 3 * (omt * omt) * t; w2 = 3 * omt * t * t; w3 = t * t * t;
        }
    }
    if (is_stroke) {
        const int toff = type == DVG_SHAPE_PATH ? srec[DVG_S_THICK_OFF] : -1;
        if (toff >= 0) {  // diffvg.cpp:110-144
            sk.add(toff + i0, w0 * contrib);
            sk.add(toff + i1, w1 * contrib);
            if (nw > 2) sk.add(toff + i2, w2 * contrib);
            if (nw > 3) sk.add(toff + i3, w3 * contrib);
        } else if (srec[DVG_S_WIDTH_OFF] >= 0) {
            sk.add(srec[DVG_S_WIDTH_OFF], contrib);
        }
    }
    switch (type) {
        case DVG_SHAPE_CIRCLE:
            sk.add(poff + 1, normal.x * contrib);
            sk.add(poff + 2, normal.y * contrib);
            sk.add(poff + 0, contrib);
            break;
        case DVG_SHAPE_ELLIPSE: {
            sk.add(poff + 2, normal.x * contrib);
            sk.add(poff + 3, normal.y * contrib);
            // the reference uses the UN-remapped random number t here (diffvg.cpp:166-167, 1373)
            const float arg = 2 * (float)DVG_PI_D * bs.t;
            sk.add(poff + 0, cosf(arg) * normal.x * contrib);
            sk.add(poff + 1, sinf(arg) * normal.y * contrib);
            break;
        }
        case DVG_SHAPE_PATH: {
            const float nx = normal.x, ny = normal.y;
            sk.add(poff + 2 * i0 + 0, w0 * nx * contrib); sk.add(poff + 2 * i0 + 1, w0 * ny * contrib);
            sk.add(poff + 2 * i1 + 0, w1 * nx * contrib); sk.add(poff + 2 * i1 + 1, w1 * ny * contrib);
            if (nw > 2) { sk.add(poff + 2 * i2 + 0, w2 * nx * contrib); sk.add(poff + 2 * i2 + 1, w2 * ny * contrib); }
            if (nw > 3) { sk.add(poff + 2 * i3 + 0, w3 * nx * contrib); sk.add(poff + 2 * i3 + 1, w3 * ny * contrib); }
            break;
        }
        default: {  // rect (diffvg.cpp:232-255): exact normal compare in LOCAL orientation
            if (normal.x == -1.f && normal.y == 0.f) sk.add(poff + 0, -contrib);
            else if (normal.x == 1.f && normal.y == 0.f) sk.add(poff + 2, contrib);
            else if (normal.x == 0.f && normal.y == -1.f) sk.add(poff + 1, -contrib);
            else if (normal.x == 0.f && normal.y == 1.f) sk.add(poff + 3, contrib);
            break;
        }
    }
}

// The same scatter as accumulate_boundary_gradient, as a fixed-slot record that the kernel fills inside its divergent
// "this lane scatters" branch and adds to the gradient buffer after the warp has re-converged (dvg_wave.cu
// scatter_record).  Slots: 0-7 point coordinates (circle: centre.x, centre.y, radius; ellipse: centre.xy, radius.xy;
// rect: the one edge coordinate), 8 stroke width, 9-12 per-point thickness.  addr < 0 = unused.  key < 0 = nothing to
// add; lanes with equal keys have equal addr[].
#define DVG_GREC_N 13
struct GradRec {
    int key;
    int addr[DVG_GREC_N];
    float val[DVG_GREC_N];
};
DVG_HD void boundary_gradient_record(const SceneView &sc, const BoundarySample &bs, const InstInfo &ii, float contrib,
                                     F2 normal, GradRec &gr) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < DVG_GREC_N; j++) { gr.addr[j] = -1; gr.val[j] = 0.f; }
    const int *srec = sc.topo + sc.topo[DVG_H_OFF_SHAPES] + ii.shape * DVG_SHAPE_REC_LEN;
    const bool is_stroke = bs.point_id_stroke < 0;
    const int point_id = bs.point_id_stroke & 0x7fffffff;
    const int type = srec[DVG_S_TYPE];
    const int pbase = ii.scene * sc.num_params;   // the scene's slice of the gradient buffer (0 outside batches)
    const int poff = pbase + srec[DVG_S_PARAM_OFF];
    const float nx = normal.x, ny = normal.y;
    if (type == DVG_SHAPE_PATH) {
        const int np = srec[DVG_S_NUM_POINTS];
        const int ncp = sc.topo[sc.topo[DVG_H_OFF_NCP] + srec[DVG_S_NCP_OFF] + bs.base_point_id];
        const float t = bs.path_t;
        float w0, w1, w2 = 0.f, w3 = 0.f;
        int i0 = point_id, i1, i2 = -1, i3 = -1;
        if (ncp == 0) {
            i1 = (point_id + 1) % np;
            w0 = 1 - t; w1 = t;
        } else if (ncp == 1) {
            i1 = point_id + 1; i2 = (point_id + 2) % np;
            w0 = (1 - t) * (1 - t); w1 = 2 * (1 - t) * t; w2 = t * t;
        } else {
            i1 = point_id + 1; i2 = point_id + 2; i3 = (point_id + 3) % np;
            const float omt = 1 - t;
            w0 = omt * omt * omt; w1 = 3 * (omt * omt) * t; w2 = 3 * omt * t * t; w3 = t * t * t;
        }
        gr.addr[0] = poff + 2 * i0; gr.addr[1] = poff + 2 * i0 + 1; gr.val[0] = w0 * nx * contrib; gr.val[1] = w0 * ny * contrib;
        gr.addr[2] = poff + 2 * i1; gr.addr[3] = poff + 2 * i1 + 1; gr.val[2] = w1 * nx * contrib; gr.val[3] = w1 * ny * contrib;
        if (i2 >= 0) { gr.addr[4] = poff + 2 * i2; gr.addr[5] = poff + 2 * i2 + 1; gr.val[4] = w2 * nx * contrib; gr.val[5] = w2 * ny * contrib; }
        if (i3 >= 0) { gr.addr[6] = poff + 2 * i3; gr.addr[7] = poff + 2 * i3 + 1; gr.val[6] = w3 * nx * contrib; gr.val[7] = w3 * ny * contrib; }
        if (is_stroke) {
            const int toff = srec[DVG_S_THICK_OFF];
            if (toff >= 0) {  // diffvg.cpp:110-144
                gr.addr[9] = pbase + toff + i0; gr.val[9] = w0 * contrib;
                gr.addr[10] = pbase + toff + i1; gr.val[10] = w1 * contrib;
                if (i2 >= 0) { gr.addr[11] = pbase + toff + i2; gr.val[11] = w2 * contrib; }
                if (i3 >= 0) { gr.addr[12] = pbase + toff + i3; gr.val[12] = w3 * contrib; }
            } else if (srec[DVG_S_WIDTH_OFF] >= 0) {
                gr.addr[8] = pbase + srec[DVG_S_WIDTH_OFF]; gr.val[8] = contrib;
            }
        }
    } else {
        if (is_stroke && srec[DVG_S_WIDTH_OFF] >= 0) { gr.addr[8] = pbase + srec[DVG_S_WIDTH_OFF]; gr.val[8] = contrib; }
        if (type == DVG_SHAPE_CIRCLE) {
            gr.addr[0] = poff + 1; gr.val[0] = nx * contrib;
            gr.addr[1] = poff + 2; gr.val[1] = ny * contrib;
            gr.addr[2] = poff + 0; gr.val[2] = contrib;
        } else if (type == DVG_SHAPE_ELLIPSE) {
            gr.addr[0] = poff + 2; gr.val[0] = nx * contrib;
            gr.addr[1] = poff + 3; gr.val[1] = ny * contrib;
            // the reference uses the UN-remapped random number t here (diffvg.cpp:166-167, 1373)
            const float arg = 2 * (float)DVG_PI_D * bs.t;
            gr.addr[2] = poff + 0; gr.val[2] = cosf(arg) * nx * contrib;
            gr.addr[3] = poff + 1; gr.val[3] = sinf(arg) * ny * contrib;
        } else {  // rect (diffvg.cpp:232-255): exact normal compare in LOCAL orientation
            if (nx == -1.f && ny == 0.f) { gr.addr[0] = poff + 0; gr.val[0] = -contrib; }
            else if (nx == 1.f && ny == 0.f) { gr.addr[0] = poff + 2; gr.val[0] = contrib; }
            else if (nx == 0.f && ny == -1.f) { gr.addr[0] = poff + 1; gr.val[0] = -contrib; }
            else if (nx == 0.f && ny == 1.f) { gr.addr[0] = poff + 3; gr.val[0] = contrib; }
        }
    }
    // lanes with equal keys have equal addr[]: the first point address identifies the segment / shape / rect
    // edge, bit 0 the stroke side (which decides slots 8-12); rect samples off the four exact normals add nothing
    gr.key = gr.addr[0] >= 0 ? gr.addr[0] * 2 + (is_stroke ? 1 : 0) : (gr.addr[8] >= 0 ? gr.addr[8] * 2 + 1 : -1);
}

// d_xform_pt part of accumulate_boundary_gradient (diffvg.cpp:256-273): the 9 terms for
// d_shape_to_canvas.  Kept separate so that the kernel can reduce them across the warp first
// (groups very often share one transform, so all lanes target the same 9 floats).
DVG_HD void boundary_xform_gradient(const BoundarySample &bs, const GroupInfo &g, float contrib, F2 normal, float dm[9]) {
    F2 dpt = mk2(0, 0);
    d_xform_pt(g.s2c, bs.local_pt, mk2(normal.x * contrib, normal.y * contrib), dm, dpt);
}

}  // namespace dvg
