// dvg_buildfn.cuh -- per-item bodies of the scene-build kernels (dvg_build.cu), written as
// host/device functions so the same arithmetic can be exercised by the host-side test
// harness (tests/host_emul/).  See dvg_build.cu for the launch structure.
#pragma once
#include "dvg_scene.cuh"
#include "dvg_geom.cuh"

namespace dvg {

// Mutable view used by the build kernels (same arrays as SceneView, non-const).
struct BuildView {
    int canvas_w, canvas_h;
    int num_shapes, num_groups, num_insts, num_prims;   // per scene
    int batch, num_params, total_segs;                  // scenes of this topology laid out back to back (SceneView)
    const int *topo;
    const float *params;
    // topology-only maps (host-built at scene creation)
    const int *inst_group, *inst_shape, *inst_prim_begin;  // inst_prim_begin has num_insts+1 entries
    const int *prim_inst, *prim_seg, *prim_point_id;
    // per shape
    float *shapes_length; Box *shape_box; float *shape_r0;
    // per segment
    float *seg_cdf, *seg_pmf; int *seg_point_id;
    // per instance / group / primitive
    InstInfo *insts; GroupInfo *groups;
    F4 *prim_p01, *prim_p23, *prim_rad; Box *prim_box; float *prim_thick; PrimMeta *prim_meta; Box *prim_cbox; Box *prim_cbox_pf; F4 *prim_cap; PrimQuintic *prim_quint; PrimWindCert *prim_wcert; int *shape_guide;
    float *shape_cdf, *shape_pmf;
    int *error_flag; float *total_length;
};

// ------------------------------------------------------------------ shapes
// One path segment (scene.cpp:132-191, 602-618): its length estimate `d` (chord of a line, two / three chords of a
// quadratic / cubic through the curve points at 1/2 resp. 1/3 and 2/3), the centre `yc` of its y-extent (the key of the
// reference's y-sort) and the largest control-point radius `th`.  `n` = number of control points, `pid` = first point.
DVG_HD void path_segment_measure(const float *p, const float *thick, int np, int n, int pid, float stroke_width,
                                 float &d, float &yc, float &th) {
    float ymin, ymax;
    if (n == 0) {
        int i0 = pid, i1 = (i0 + 1) % np;
        F2 p0 = mk2(p[2 * i0], p[2 * i0 + 1]), p1 = mk2(p[2 * i1], p[2 * i1 + 1]);
        d = distance2(p1, p0);
        ymin = rminf(p1.y, rminf(p0.y, INFINITY)); ymax = rmaxf(p1.y, rmaxf(p0.y, -INFINITY));
        th = thick ? rmaxf(thick[i0], thick[i1]) : stroke_width;
    } else if (n == 1) {
        int i0 = pid, i1 = i0 + 1, i2 = (i0 + 2) % np;
        F2 p0 = mk2(p[2 * i0], p[2 * i0 + 1]), p1 = mk2(p[2 * i1], p[2 * i1 + 1]), p2 = mk2(p[2 * i2], p[2 * i2 + 1]);
        F2 v1 = eval_quad(p0, p1, p2, 0.5f);
        d = distance2(v1, p0) + distance2(v1, p2);
        ymin = rminf(p2.y, rminf(p1.y, rminf(p0.y, INFINITY)));
        ymax = rmaxf(p2.y, rmaxf(p1.y, rmaxf(p0.y, -INFINITY)));
        th = thick ? rmaxf(rmaxf(thick[i0], thick[i1]), thick[i2]) : stroke_width;
    } else {
        int i0 = pid, i1 = i0 + 1, i2 = i0 + 2, i3 = (i0 + 3) % np;
        F2 p0 = mk2(p[2 * i0], p[2 * i0 + 1]), p1 = mk2(p[2 * i1], p[2 * i1 + 1]);
        F2 p2 = mk2(p[2 * i2], p[2 * i2 + 1]), p3 = mk2(p[2 * i3], p[2 * i3 + 1]);
        F2 v1 = eval_cubic(p0, p1, p2, p3, 1.f / 3.f), v2 = eval_cubic(p0, p1, p2, p3, 2.f / 3.f);
        d = distance2(v1, p0) + distance2(v1, v2) + distance2(v2, p3);
        ymin = rminf(p3.y, rminf(p2.y, rminf(p1.y, rminf(p0.y, INFINITY))));
        ymax = rmaxf(p3.y, rmaxf(p2.y, rmaxf(p1.y, rmaxf(p0.y, -INFINITY))));
        th = thick ? rmaxf(rmaxf(rmaxf(thick[i0], thick[i1]), thick[i2]), thick[i3]) : stroke_width;
    }
    yc = 0.5f * (ymin + ymax);
}

// shapes_length (scene.cpp:113-205), shapes_bbox (499-629), per-path segment pmf/cdf/point-id
// map (248-333) and the "first leaf after the y-sort" radius the reference uses as the group
// radius of thickness paths (scene.cpp:602-618, 650-667).
DVG_HD_NOINLINE void build_shape(const BuildView &bv, int s_batch) {
    const int *topo = bv.topo;
    const int scene = s_batch / bv.num_shapes, s = s_batch - scene * bv.num_shapes;
    const float *P = bv.params + (size_t)scene * bv.num_params;
    const int seg_base = scene * bv.total_segs;
    const int *r = topo + topo[DVG_H_OFF_SHAPES] + s * DVG_SHAPE_REC_LEN;
    const float *p = P + r[DVG_S_PARAM_OFF];
    float stroke_width = r[DVG_S_WIDTH_OFF] >= 0 ? P[r[DVG_S_WIDTH_OFF]] : 0.f;
    float len = 0.f;
    Box box;
    float r0q = stroke_width;
    const float pi_f = (float)DVG_PI_D;
    switch (r[DVG_S_TYPE]) {
        case DVG_SHAPE_CIRCLE:
            len += (float)(2.f * DVG_PI_D) * p[0];
            box.x0 = p[1] - p[0]; box.y0 = p[2] - p[0]; box.x1 = p[1] + p[0]; box.y1 = p[2] + p[0];
            break;
        case DVG_SHAPE_ELLIPSE: {
            float a = p[0], b = p[1];
            // scene.cpp:130: the unqualified sqrt is ::sqrt(double), so the difference and the product with float(M_PI)
            // are formed in double and rounded once, by the += into the float length
            len = (float)((double)len + (double)pi_f * ((double)(3 * (a + b)) - sqrt((double)((3 * a + b) * (a + 3 * b)))));
            box.x0 = p[2] - p[0]; box.y0 = p[3] - p[1]; box.x1 = p[2] + p[0]; box.y1 = p[3] + p[1];
            break;
        }
        case DVG_SHAPE_RECT:
            len += 2 * (p[2] - p[0] + p[3] - p[1]);
            box.x0 = p[0]; box.y0 = p[1]; box.x1 = p[2]; box.y1 = p[3];
            break;
        default: {
            const int np = r[DVG_S_NUM_POINTS], nseg = r[DVG_S_NUM_SEGS];
            const int *ncp = topo + topo[DVG_H_OFF_NCP] + r[DVG_S_NCP_OFF];
            const float *thick = r[DVG_S_THICK_OFF] >= 0 ? P + r[DVG_S_THICK_OFF] : nullptr;
            float *seg_pmf = bv.seg_pmf + seg_base + r[DVG_S_NCP_OFF];
            float *seg_cdf = bv.seg_cdf + seg_base + r[DVG_S_NCP_OFF];
            int *seg_pid = bv.seg_point_id + seg_base + r[DVG_S_NCP_OFF];
            box.x0 = box.y0 = INFINITY; box.x1 = box.y1 = -INFINITY;
            if (np > 0) { box.x0 = box.x1 = p[0]; box.y0 = box.y1 = p[1]; }
            for (int i = 1; i < np; i++) {
                float x = p[2 * i], y = p[2 * i + 1];
                box.x0 = rminf(x, box.x0); box.y0 = rminf(y, box.y0);
                box.x1 = rmaxf(x, box.x1); box.y1 = rmaxf(y, box.y1);
            }
            float length = 0.f;
            int pid = 0;
            float best_y = INFINITY;
            // pass 1: total length (scene.cpp:132-191); raw segment lengths parked in seg_pmf
            for (int i = 0; i < nseg; i++) {
                seg_pid[i] = pid;
                float d, yc, th;
                path_segment_measure(p, thick, np, ncp[i], pid, stroke_width, d, yc, th);
                pid += ncp[i] + 1;
                length += d;
                seg_pmf[i] = d;
                if (yc < best_y) { best_y = yc; if (thick) r0q = th; }
            }
            len += length;
            // pass 2: pmf / cdf with the reciprocal of the total (scene.cpp:257-326)
            float inv_length = 1.f / len;
            float c = 0.f;
            for (int i = 0; i < nseg; i++) {
                float d = seg_pmf[i] * inv_length;
                seg_pmf[i] = d;
                c = (i == 0) ? d : d + c;
                seg_cdf[i] = c;
            }
            break;
        }
    }
    bv.shapes_length[s_batch] = len;
    bv.shape_box[s_batch] = box;
    bv.shape_r0[s_batch] = r0q;
}

DVG_HD Box box_merge(Box a, Box b) {
    Box o;
    o.x0 = rminf(a.x0, b.x0); o.y0 = rminf(a.y0, b.y0); o.x1 = rmaxf(a.x1, b.x1); o.y1 = rmaxf(a.y1, b.y1);
    return o;
}
DVG_HD Box box_merge_pt(Box a, F2 p) {
    Box o;
    o.x0 = rminf(p.x, a.x0); o.y0 = rminf(p.y, a.y0); o.x1 = rmaxf(p.x, a.x1); o.y1 = rmaxf(p.y, a.y1);
    return o;
}
DVG_HD Box box_transform(const float *m, Box b) {  // aabb.h:52-60
    Box o; o.x0 = o.y0 = INFINITY; o.x1 = o.y1 = -INFINITY;
    o = box_merge_pt(o, xform_pt(m, mk2(b.x0, b.y0)));
    o = box_merge_pt(o, xform_pt(m, mk2(b.x0, b.y1)));
    o = box_merge_pt(o, xform_pt(m, mk2(b.x1, b.y0)));
    o = box_merge_pt(o, xform_pt(m, mk2(b.x1, b.y1)));
    return o;
}


// ------------------------------------------------------------------ groups
// transforms, group root box, scene-BVH leaf box and radius (scene.cpp:632-682, shape.h:122-124)
DVG_HD_NOINLINE void build_group(const BuildView &bv, int g_batch) {
    const int *topo = bv.topo;
    const int scene = g_batch / bv.num_groups, g = g_batch - scene * bv.num_groups;
    const int pbase = scene * bv.num_params;   // colour / transform offsets are stored absolute (see SceneView)
    const float *P = bv.params + pbase;
    const int *r = topo + topo[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN;
    const int *ids = topo + topo[DVG_H_OFF_GSHAPES] + r[DVG_G_SHAPES_OFF];
    const Box *shape_box = bv.shape_box + scene * bv.num_shapes;
    const float *shape_r0 = bv.shape_r0 + scene * bv.num_shapes;
    GroupInfo gi;
    gi.fill_type = r[DVG_G_FILL_TYPE]; gi.fill_off = pbase + r[DVG_G_FILL_OFF]; gi.fill_stops = r[DVG_G_FILL_STOPS];
    gi.stroke_type = r[DVG_G_STROKE_TYPE]; gi.stroke_off = pbase + r[DVG_G_STROKE_OFF]; gi.stroke_stops = r[DVG_G_STROKE_STOPS];
    gi.num_shapes = r[DVG_G_NUM_SHAPES];
    gi.inst_begin = scene * bv.num_insts + r[DVG_G_SHAPES_OFF];
    gi.prim_begin = scene * bv.num_prims + bv.inst_prim_begin[r[DVG_G_SHAPES_OFF]];
    gi.prim_end = scene * bv.num_prims + bv.inst_prim_begin[r[DVG_G_SHAPES_OFF] + gi.num_shapes];
    gi.xform_off = pbase + r[DVG_G_XFORM_OFF];
    gi.pad = 0;
    const float *m = P + r[DVG_G_XFORM_OFF];
    for (int k = 0; k < 9; k++) gi.s2c[k] = m[k];
    inverse3(gi.s2c, gi.c2s);
    bool ident = m[0] == 1.f && m[1] == 0.f && m[2] == 0.f && m[3] == 0.f && m[4] == 1.f && m[5] == 0.f &&
                 m[6] == 0.f && m[7] == 0.f && m[8] == 1.f;
    bool affine = m[6] == 0.f && m[7] == 0.f && m[8] == 1.f;
    gi.flags = (r[DVG_G_EVEN_ODD] ? DVG_GF_EVEN_ODD : 0) | (ident ? DVG_GF_IDENTITY : 0) | (affine ? DVG_GF_AFFINE : 0);
    Box lb = shape_box[ids[0]];
    float max_radius = shape_r0[ids[0]];
    for (int k = 1; k < gi.num_shapes; k++) {
        lb = box_merge(lb, shape_box[ids[k]]);
        float rr = shape_r0[ids[k]];
        max_radius = max_radius > rr ? max_radius : rr;  // std::max(a, b): (a < b) ? b : a
    }
    gi.local_box = lb;
    gi.scene_box = box_transform(gi.s2c, lb);
    gi.scene_r = gi.stroke_type < 0 ? 0.f : max_radius;
    bv.groups[g_batch] = gi;
}

// ------------------------------------------------------------------ reject capsules
DVG_HD float seg_dist(F2 a, F2 d, float inv_len2, F2 p) {
    F2 w = p - a;
    float t = clampf(dot2(w, d) * inv_len2, 0.f, 1.f);
    F2 e = w - t * d;
    return sqrtf(dot2(e, e));
}
DVG_HD void cap_piece(const F2 *q, float rmax, float rmin, float *out) {
    F2 d = q[3] - q[0];
    float len2 = dot2(d, d);
    float inv = len2 > 1e-12f ? 1.f / len2 : 0.f;
    if (inv == 0.f) d = mk2(0, 0);
    float dev = rmaxf(seg_dist(q[0], d, inv, q[1]), seg_dist(q[0], d, inv, q[2]));
    // margin: float rounding of the de Casteljau split, of eval_cubic in the exact test and of this
    // test itself are all < 1e-4 px at canvas scales; 1e-2 px + 1e-4 relative is far above that
    float Ro = (dev + rmax) * 1.0001f + 1e-2f;
    float Ri = (rmin - dev) * 0.9999f - 1e-2f;
    out[0] = q[0].x; out[1] = q[0].y; out[2] = d.x; out[3] = d.y; out[4] = inv; out[5] = Ro * Ro;
    out[6] = Ri > 0.f ? Ri * Ri : -1.f;
    out[7] = 0.f;
}
DVG_HD void cap_split(const F2 *p, F2 *l, F2 *r) {  // de Casteljau at 1/2
    F2 a = 0.5f * (p[0] + p[1]), b = 0.5f * (p[1] + p[2]), c = 0.5f * (p[2] + p[3]);
    F2 ab = 0.5f * (a + b), bc = 0.5f * (b + c);
    F2 m = 0.5f * (ab + bc);
    l[0] = p[0]; l[1] = a; l[2] = ab; l[3] = m;
    r[0] = m; r[1] = bc; r[2] = c; r[3] = p[3];
}
// type: PRIM_QUAD / PRIM_CUBIC get real brackets; everything else a never-decide record.
DVG_HD void build_capsules(int type, F4 p01, F4 p23, float rmax, float rmin, float *out) {
    bool ok = (type == PRIM_QUAD || type == PRIM_CUBIC) && rmax == rmax && rmin == rmin;
    F2 c[4];
    c[0] = mk2(p01.x, p01.y);
    if (type == PRIM_CUBIC) { c[1] = mk2(p01.z, p01.w); c[2] = mk2(p23.x, p23.y); c[3] = mk2(p23.z, p23.w); }
    else {  // degree elevation of the quadratic
        F2 q1 = mk2(p01.z, p01.w), q2 = mk2(p23.x, p23.y);
        c[1] = c[0] + (2.f / 3.f) * (q1 - c[0]); c[2] = q2 + (2.f / 3.f) * (q1 - q2); c[3] = q2;
    }
    for (int k = 0; k < 4; k++) ok = ok && fabsf(c[k].x) < 1e18f && fabsf(c[k].y) < 1e18f;   // finite, no overflow below
    if (!ok) {
        for (int i = 0; i < DVG_CAP_N; i++) {
            float *o = out + 8 * i;
            o[0] = o[1] = o[2] = o[3] = o[4] = o[7] = 0.f; o[5] = INFINITY; o[6] = -1.f;
        }
        return;
    }
    // three levels of halving -> 8 pieces, written in curve order
    F2 h[2][4], q[4][4], e[2][4];
    cap_split(c, h[0], h[1]);
    cap_split(h[0], q[0], q[1]);
    cap_split(h[1], q[2], q[3]);
    for (int i = 0; i < 4; i++) {
        cap_split(q[i], e[0], e[1]);
        cap_piece(e[0], rmax, rmin, out + 8 * (2 * i));
        cap_piece(e[1], rmax, rmin, out + 8 * (2 * i + 1));
    }
}

// Tight binning of curved strokes.  A diagonal 40-px stroke has a 900 px^2 bounding box and a 100 px^2
// footprint; binned by box, two thirds of a tile's candidates are strokes that no sample of the tile can touch,
// and every one of them costs each sample a leaf test + an 8-piece bracket test in the classify kernels.  The
// polyline bracket (dvg_scene.cuh) proves "farther than R_out from every chord => the exact stroke test returns
// false"; here the same statement is made for a whole tile: the tile is cut into squares, and a square whose
// centre is farther than R_out + half-diagonal from a chord cannot contain such a point.  NaN records keep.
DVG_HD bool bracket_reaches_tile(const F4 *cap, float x0, float y0, float x1, float y1) {
    const float w = x1 - x0, h = y1 - y0;
    const bool wide = w >= h;
    const float side = wide ? h : w;
    int nsq = (int)ceilf((wide ? w : h) / (side > 1e-6f ? side : 1e-6f));
    nsq = nsq < 1 ? 1 : (nsq > 8 ? 8 : nsq);
    const float step = (wide ? w : h) / nsq;
    const float hd = 0.5f * sqrtf(step * step + side * side);   // half diagonal of one piece of the tile
    bool far_all = true;
    for (int i = 0; i < DVG_CAP_N; i++) {
        const F4 ca = cap[2 * i], cb = cap[2 * i + 1];   // A.xy, d.xy | 1/|d|^2, R_out^2, R_in^2, pad
        const float R = sqrtf(cb.y) + hd;
        const float thr = R * R;
        for (int q = 0; q < nsq; q++) {
            const float cx = wide ? x0 + (q + 0.5f) * step : 0.5f * (x0 + x1);
            const float cy = wide ? 0.5f * (y0 + y1) : y0 + (q + 0.5f) * step;
            const float wx = cx - ca.x, wy = cy - ca.y;
            float t = (wx * ca.z + wy * ca.w) * cb.x;
            t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
            const float ex = wx - t * ca.z, ey = wy - t * ca.w;
            far_all = far_all && (ex * ex + ey * ey > thr);   // false for NaN / inf
        }
    }
    return !far_all;
}

// ------------------------------------------------------------------ primitives
// Leaf boxes and radii follow scene.cpp:527-600 (topology-only maps prim -> inst / segment /
// first point come from the host).
DVG_HD_NOINLINE void build_prim(const BuildView &bv, int e_batch) {
    const int *topo = bv.topo;
    const int scene = e_batch / bv.num_prims, e_local = e_batch - scene * bv.num_prims;
    const float *P = bv.params + (size_t)scene * bv.num_params;
    const int inst_local = bv.prim_inst[e_local];
    const int inst = scene * bv.num_insts + inst_local;                       // batch-wide ids from here on
    const int g = scene * bv.num_groups + bv.inst_group[inst_local], s = bv.inst_shape[inst_local];   // s: id inside the scene
    const int e = e_batch;
    const Box *shape_boxes = bv.shape_box + scene * bv.num_shapes;
    const GroupInfo &gi = bv.groups[g];
    const int *r = topo + topo[DVG_H_OFF_SHAPES] + s * DVG_SHAPE_REC_LEN;
    const float *p = P + r[DVG_S_PARAM_OFF];
    const float sw = r[DVG_S_WIDTH_OFF] >= 0 ? P[r[DVG_S_WIDTH_OFF]] : 0.f;
    const bool has_stroke = gi.stroke_type >= 0, has_fill = gi.fill_type >= 0;
    F4 p01 = mk4(0, 0, 0, 0), p23 = mk4(0, 0, 0, 0), rad = mk4(sw, sw, sw, sw);
    Box box;
    float thick = sw;
    PrimMeta pm;
    pm.inst = inst; pm.point_id = 0; pm.base_id = 0;
    int tf;
    const bool first_in_inst = (e_local == bv.inst_prim_begin[inst_local]);
    switch (r[DVG_S_TYPE]) {
        case DVG_SHAPE_CIRCLE:
            tf = PRIM_CIRCLE | DVG_PF_SINGLE;
            p01 = mk4(p[1], p[2], p[0], 0.f);
            box = shape_boxes[s];
            break;
        case DVG_SHAPE_ELLIPSE:
            tf = PRIM_ELLIPSE | DVG_PF_SINGLE;
            p01 = mk4(p[2], p[3], p[0], p[1]);
            box = shape_boxes[s];
            break;
        case DVG_SHAPE_RECT:
            tf = PRIM_RECT | DVG_PF_SINGLE;
            p01 = mk4(p[0], p[1], p[2], p[3]);
            box = shape_boxes[s];
            break;
        default: {
            const int np = r[DVG_S_NUM_POINTS], nseg = r[DVG_S_NUM_SEGS];
            const int seg = bv.prim_seg[e_local], pid = bv.prim_point_id[e_local];
            const int *ncp = topo + topo[DVG_H_OFF_NCP] + r[DVG_S_NCP_OFF];
            const float *th = r[DVG_S_THICK_OFF] >= 0 ? P + r[DVG_S_THICK_OFF] : nullptr;
            pm.point_id = pid; pm.base_id = seg;
            const int n = ncp[seg];
            tf = n | (nseg == 1 ? DVG_PF_SINGLE : 0) | (th ? DVG_PF_THICK : 0) |
                 ((r[DVG_S_FLAGS] & DVG_SF_DISTANCE_APPROX) ? DVG_PF_APPROX : 0);
            int i0 = pid, i1, i2 = 0, i3 = 0;
            if (n == 0) { i1 = (i0 + 1) % np; }
            else if (n == 1) { i1 = i0 + 1; i2 = (i0 + 2) % np; }
            else { i1 = i0 + 1; i2 = i0 + 2; i3 = (i0 + 3) % np; }
            F2 q0 = mk2(p[2 * i0], p[2 * i0 + 1]), q1 = mk2(p[2 * i1], p[2 * i1 + 1]);
            box.x0 = box.y0 = INFINITY; box.x1 = box.y1 = -INFINITY;
            box = box_merge_pt(box, q0);
            box = box_merge_pt(box, q1);
            p01 = mk4(q0.x, q0.y, q1.x, q1.y);
            if (th) { rad.x = th[i0]; rad.y = th[i1]; thick = rmaxf(rad.x, rad.y); }
            if (n >= 1) {
                F2 q2 = mk2(p[2 * i2], p[2 * i2 + 1]);
                box = box_merge_pt(box, q2);
                p23.x = q2.x; p23.y = q2.y;
                if (th) { rad.z = th[i2]; thick = rmaxf(thick, rad.z); }
            }
            if (n >= 2) {
                F2 q3 = mk2(p[2 * i3], p[2 * i3 + 1]);
                box = box_merge_pt(box, q3);
                p23.z = q3.x; p23.w = q3.y;
                if (th) { rad.w = th[i3]; thick = rmaxf(thick, rad.w); }
            }
            break;
        }
    }
    if (first_in_inst) tf |= DVG_PF_FIRST;
    if (e == gi.prim_begin) tf |= DVG_PF_GFIRST;
    {
        const int pt0 = tf & DVG_PF_TYPE_MASK;
        if (has_stroke && !has_fill && (gi.flags & DVG_GF_IDENTITY) && (pt0 == PRIM_CUBIC || pt0 == PRIM_QUAD) && !(tf & DVG_PF_APPROX))
            tf |= DVG_PF_TIGHT;
    }
    if ((tf & DVG_PF_TYPE_MASK) == PRIM_CUBIC && has_fill) {   // winding certificate of the classifier (dvg_geom.cuh)
        PrimWindCert wc;
        tf |= prim_wind_cert(mk2(p01.x, p01.y), mk2(p01.z, p01.w), mk2(p23.x, p23.y), mk2(p23.z, p23.w), wc);
        bv.prim_wcert[e] = wc;
    }
    pm.type_flags = tf;
    bv.prim_p01[e] = p01; bv.prim_p23[e] = p23; bv.prim_rad[e] = rad;
    bv.prim_box[e] = box; bv.prim_thick[e] = thick; bv.prim_meta[e] = pm;
    if ((tf & DVG_PF_TYPE_MASK) == PRIM_CUBIC && has_stroke)   // read by the exact stroke test of cubic segments only
        bv.prim_quint[e] = prim_quintic(mk2(p01.x, p01.y), mk2(p01.z, p01.w), mk2(p23.x, p23.y), mk2(p23.z, p23.w));
    {
        float cap[DVG_CAP_N * 8];
        const int ptype = tf & DVG_PF_TYPE_MASK;
        float rmin = thick;   // smallest control-point radius (per-point thickness) or the stroke width
        if (tf & DVG_PF_THICK) {
            rmin = rminf(rad.x, rad.y);
            if (ptype >= PRIM_QUAD) rmin = rminf(rmin, rad.z);
            if (ptype == PRIM_CUBIC) rmin = rminf(rmin, rad.w);
        }
        // a fill-only curve gets the bracket of the radius-1 closest-point search of the prefiltered path
        // (compute_distance(..., 1.f, ...), diffvg.cpp:891-932): PrefilterTracer skips the solve of a segment that is
        // certainly farther than 1 from the sample
        if (has_stroke) build_capsules(ptype, p01, p23, thick, rmin, cap);
        else build_capsules(has_fill ? ptype : -1, p01, p23, 1.f, 0.f, cap);
        for (int k = 0; k < DVG_CAP_F4; k++) bv.prim_cap[(size_t)e * DVG_CAP_F4 + k] = mk4(cap[4 * k], cap[4 * k + 1], cap[4 * k + 2], cap[4 * k + 3]);
    }
    if (first_in_inst) {
        InstInfo ii;
        ii.box = shape_boxes[s];
        ii.r = has_stroke ? sw : 0.f;   // scene.cpp:638
        ii.group = g; ii.shape = s; ii.prim_begin = e; ii.scene = scene;
        bv.insts[inst] = ii;
    }
    // Conservative canvas-space bound of the region where this primitive can change a sample
    // (used only for binning; the exact per-sample predicates are evaluated in the kernels).
    Box reg; reg.x0 = reg.y0 = INFINITY; reg.x1 = reg.y1 = -INFINITY;
    if (has_stroke) {
        Box sb; sb.x0 = box.x0 - thick; sb.y0 = box.y0 - thick; sb.x1 = box.x1 + thick; sb.y1 = box.y1 + thick;
        reg = box_merge(reg, sb);
    }
    if (has_fill) {
        Box fb; fb.x0 = rminf(gi.local_box.x0, box.x0); fb.y0 = box.y0; fb.x1 = box.x1; fb.y1 = box.y1;
        reg = box_merge(reg, fb);
    }
    Box cb;
    const float big = 3.0e38f;
    if (!has_stroke && !has_fill) { cb.x0 = cb.y0 = big; cb.x1 = cb.y1 = -big; }
    else if (gi.flags & DVG_GF_IDENTITY) cb = reg;
    else if ((gi.flags & DVG_GF_AFFINE) && reg.x0 <= reg.x1) cb = box_transform(gi.s2c, reg);
    else { cb.x0 = cb.y0 = -big; cb.x1 = cb.y1 = big; }
    if (bv.num_groups > 1) {  // the group is only visited inside its scene-BVH leaf box (+ radius)
        cb.x0 = rmaxf(cb.x0, gi.scene_box.x0 - gi.scene_r); cb.y0 = rmaxf(cb.y0, gi.scene_box.y0 - gi.scene_r);
        cb.x1 = rminf(cb.x1, gi.scene_box.x1 + gi.scene_r); cb.y1 = rminf(cb.y1, gi.scene_box.y1 + gi.scene_r);
    }
    if (!(cb.x0 == cb.x0 && cb.x1 == cb.x1 && cb.y0 == cb.y0 && cb.y1 == cb.y1)) {  // NaN -> everywhere
        cb.x0 = cb.y0 = -big; cb.x1 = cb.y1 = big;
    }
    bv.prim_cbox[e] = cb;
    // Same for the SDF-prefiltering path (diffvg.cpp:835-1113): a stroked group runs an unbounded
    // closest-point search over all its segments wherever the group is visited; a fill-only group
    // searches within radius 1 (compute_distance(..., 1.f, ...)) and casts the winding ray.
    Box cp;
    if (!has_stroke && !has_fill) { cp.x0 = cp.y0 = big; cp.x1 = cp.y1 = -big; }
    else if (has_stroke) { cp.x0 = cp.y0 = -big; cp.x1 = cp.y1 = big; }
    else {
        Box rp; rp.x0 = rminf(gi.local_box.x0, box.x0 - 1.f); rp.y0 = box.y0 - 1.f; rp.x1 = box.x1 + 1.f; rp.y1 = box.y1 + 1.f;
        if (gi.flags & DVG_GF_IDENTITY) cp = rp;
        else if (gi.flags & DVG_GF_AFFINE) cp = box_transform(gi.s2c, rp);
        else { cp.x0 = cp.y0 = -big; cp.x1 = cp.y1 = big; }
    }
    if (bv.num_groups > 1) {
        cp.x0 = rmaxf(cp.x0, gi.scene_box.x0 - gi.scene_r); cp.y0 = rmaxf(cp.y0, gi.scene_box.y0 - gi.scene_r);
        cp.x1 = rminf(cp.x1, gi.scene_box.x1 + gi.scene_r); cp.y1 = rminf(cp.y1, gi.scene_box.y1 + gi.scene_r);
    }
    if (!(cp.x0 == cp.x0 && cp.x1 == cp.x1 && cp.y0 == cp.y0 && cp.y1 == cp.y1)) {
        cp.x0 = cp.y0 = -big; cp.x1 = cp.y1 = big;
    }
    bv.prim_cbox_pf[e] = cp;
}

// ------------------------------------------------------------------ shape CDF (sequential part)
// scene.cpp:207-235: float prefix sum in the reference's order; returns the normalisation.
DVG_HD_NOINLINE float build_shape_cdf_serial(const BuildView &bv) {
    float c = 0.f;
    for (int i = 0; i < bv.num_insts; i++) {
        float len = bv.shapes_length[bv.inst_shape[i]];
        c = (i == 0) ? len : len + c;
        bv.shape_cdf[i] = c;
        bv.shape_pmf[i] = len;
    }
    if (!(c > 0.f)) *bv.error_flag = 1;            // scene.cpp:231-235 (also catches NaN)
    else if (isinf(c)) *bv.error_flag = 2;          // scene.cpp:236-240
    else *bv.error_flag = 0;
    *bv.total_length = c;
    return c;
}

}  // namespace dvg
