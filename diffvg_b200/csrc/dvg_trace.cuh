// dvg_trace.cuh -- per-sample state machine that consumes candidate primitives in ascending id
// order and produces what sample_color(scene, ...) (diffvg.cpp:525-653) produces: the "over"
// composite of all fragments hit by the sample, plus the EdgeQuery verdict.
//
// Why a state machine instead of the reference's collect-sort-blend:
//   * primitives arrive ordered by (group, shape-in-group, segment) so fragments are produced
//     already sorted by group id with the stroke fragment before the fill fragment, which is
//     exactly the order the reference's stable insertion sort yields (diffvg.cpp:555-582,
//     605-615) -- no fragment array, no sort;
//   * every inner-node test of the reference's three BVH levels is implied by the test at its
//     leaf (boxes and radii are merged monotonically, scene.cpp:476-480), so traversal
//     collapses to flat per-leaf predicates evaluated with the same comparisons:
//       group  leaf: inside(scene_box, pt, scene_r)          unless the scene has one group
//       shape  leaf: inside(shape_box, local_pt[, r])        unless the group has one shape
//       segment leaf: within_distance / intersect(seg_box)   unless the path has one segment
//     (a BVH root is always visited without a test).
#pragma once
#include "dvg_scene.cuh"
#include "dvg_geom.cuh"
#include "dvg_color.cuh"

namespace dvg {

constexpr int DVG_MAXF = 256;  // fragment capacity per sample (diffvg.cpp:542)

enum { TM_IMMEDIATE = 0, TM_CLASSIFY = 1, TM_CONSUME = 2 };
#define DVG_NEED_STROKE 1
#define DVG_NEED_FILL 2

// One staged primitive, as the tracer sees it.
struct PrimRef {
    F4 p01, p23, rad;
    Box box;
    float thick;
    int tf;     // type | flags
    int inst;
    int group;
    int base_id, point_id;  // path segment index / first point index (prefilter path only)
    const float *cap;  // DVG_CAP_N * 6 floats (reject capsules)
};

template <bool EDGE, bool RECORD>
struct SampleTracer {
    // inputs
    F2 cpt;          // sample position, canvas space
    bool active;
    int q_group, q_shape;   // EdgeQuery target (EDGE only)
    int *fkey; F4 *fprev;   // fragment record for the backward pass (RECORD only)
    // composite
    F4 accum;        // premultiplied rgb + alpha after the last fragment
    int nfrag, sp;
    // current group / shape
    int cur_g, cur_inst, cur_shape;
    const GroupInfo *gp;
    F2 lpt;
    bool g_visit, g_fill_ok, s_stroke_ok, s_fill_ok;
    bool stroke_hit, sh_hit, sh_done, has_stroke, multi;
    int winding, w_shape;
    float shape_r;
    // EdgeQuery bookkeeping
    bool hit0, seen_q, opaque_after_q, any_opaque;

    DVG_HD void init(F2 cpt_, bool active_, F4 first, int qg, int qs, int *fkey_, F4 *fprev_) {
        cpt = cpt_; active = active_; q_group = qg; q_shape = qs; fkey = fkey_; fprev = fprev_;
        accum = first; nfrag = 0; sp = 0;
        cur_g = -1; cur_inst = -1; cur_shape = -1; gp = nullptr; lpt = cpt_;
        g_visit = g_fill_ok = s_stroke_ok = s_fill_ok = false;
        stroke_hit = sh_hit = sh_done = has_stroke = multi = false;
        winding = 0; w_shape = 0; shape_r = 0.f;
        hit0 = seen_q = opaque_after_q = any_opaque = false;
    }

    // diffvg.cpp:628-648 for one fragment (+ the EdgeQuery rules of 634-643, deferred: see q_hit()).
    DVG_HD void emit(F4 c, int is_stroke) {
        if (RECORD) {
            if (sp < DVG_MAXF) { fkey[sp] = (cur_g << 1) | is_stroke; fprev[sp] = accum; sp++; }
        }
        if (EDGE) {
            if (c.w >= 1.f) { any_opaque = true; opaque_after_q = true; }
            if (cur_g == q_group) { seen_q = true; opaque_after_q = false; }
        }
        const float oma = 1 - c.w;
        accum.x = accum.x * oma + c.w * c.x;
        accum.y = accum.y * oma + c.w * c.y;
        accum.z = accum.z * oma + c.w * c.z;
        accum.w = accum.w * oma + c.w;
        nfrag++;
    }

    DVG_HD void end_shape() {
        if (cur_inst >= 0) {
            if (EDGE && cur_g == q_group && cur_shape == q_shape) {  // diffvg.cpp:60-68
                const bool eo = (gp->flags & DVG_GF_EVEN_ODD) != 0;
                const int aw = w_shape < 0 ? -w_shape : w_shape;
                if ((eo && (aw % 2 == 1)) || (!eo && w_shape != 0)) hit0 = true;
            }
            winding += w_shape;
            w_shape = 0;
        }
    }

    DVG_HD void end_group(const float *params) {
        if (cur_g < 0) return;
        end_shape();
        if (has_stroke && stroke_hit)  // diffvg.cpp:555-568
            emit(eval_color(gp->stroke_type, params + gp->stroke_off, gp->stroke_stops, cpt), 1);
        if (g_fill_ok) {  // diffvg.cpp:569-582 with is_inside 82-86
            const int aw = winding < 0 ? -winding : winding;
            const bool inside = (gp->flags & DVG_GF_EVEN_ODD) ? (aw % 2 == 1) : (winding != 0);
            if (inside) emit(eval_color(gp->fill_type, params + gp->fill_off, gp->fill_stops, cpt), 0);
        }
    }

    DVG_HD void begin_group(const SceneView &sc, int g) {
        cur_g = g; cur_inst = -1; cur_shape = -1;
        gp = &sc.groups[g];
        stroke_hit = false; winding = 0; w_shape = 0;
        has_stroke = gp->stroke_type >= 0;
        multi = gp->num_shapes > 1;
        // scene-BVH leaf test (diffvg.cpp:585-592); a lone root is visited unconditionally
        g_visit = active && (sc.num_groups == 1 || box_inside_r(gp->scene_box, cpt, gp->scene_r));
        lpt = (gp->flags & DVG_GF_IDENTITY) ? cpt : xform_pt(gp->c2s, cpt);
        g_fill_ok = g_visit && gp->fill_type >= 0 && box_inside(gp->local_box, lpt);  // diffvg.cpp:42-45
    }

    DVG_HD void begin_shape(const SceneView &sc, int inst) {
        cur_inst = inst;
        const InstInfo &ii = sc.insts[inst];
        cur_shape = ii.shape;
        shape_r = ii.r;
        sh_hit = false; sh_done = false;
        // group-BVH leaf tests (within_distance.h:382-388, diffvg.cpp:71-78)
        s_stroke_ok = g_visit && has_stroke && (!multi || box_inside_r(ii.box, lpt, ii.r));
        s_fill_ok = g_fill_ok && (!multi || box_inside(ii.box, lpt));
    }

    // Consume one candidate primitive.  Group / shape changes are detected here; in the kernels
    // they are uniform across the warp because all lanes walk the same list.
    //
    // Three modes (the kernels split the work so that the expensive exact tests can be
    // re-packed across the lanes of a warp, see dvg_render.cu):
    //   TM_IMMEDIATE  evaluate the exact predicates inline (host harness; reference order);
    //   TM_CLASSIFY   only the cheap leaf tests: returns which exact tests this sample needs for
    //                 this primitive (DVG_NEED_*), no compositing state is touched;
    //   TM_CONSUME    replay with the results of the exact tests (`need`, `res_hit`, `res_wind`).
    // Skipping a test because the shape / group was already hit (IMMEDIATE) and evaluating it
    // anyway (CLASSIFY + CONSUME) give the same flags: hits are OR-ed, windings only summed.
    template <int MODE>
    DVG_HD int step(const SceneView &sc, const PrimRef &pr, int need = 0, bool res_hit = false, int res_wind = 0) {
        if (pr.group != cur_g) {
            if (MODE != TM_CLASSIFY) end_group(sc.params);
            begin_group(sc, pr.group);
        }
        if (pr.inst != cur_inst) {
            if (MODE != TM_CLASSIFY) end_shape();
            begin_shape(sc, pr.inst);
        }
        const int tf = pr.tf;
        if (MODE == TM_CONSUME) {
            if ((need & DVG_NEED_STROKE) && res_hit) {
                sh_hit = true; stroke_hit = true;
                if (EDGE && cur_g == q_group && cur_shape == q_shape) hit0 = true;  // within_distance.h:426-429
            }
            if (need & DVG_NEED_FILL) w_shape += res_wind;
            return 0;
        }
        int out = 0;
        const bool is_q_group = EDGE && cur_g == q_group;
        bool test_stroke = s_stroke_ok && !sh_done;
        if (MODE == TM_IMMEDIATE) test_stroke = test_stroke && !sh_hit && (!stroke_hit || is_q_group);
        if (test_stroke) {
            // path-BVH leaf test (within_distance.h:278-285)
            if ((tf & DVG_PF_SINGLE) || box_inside_r(pr.box, lpt, pr.thick)) {
                if (MODE == TM_CLASSIFY) {
                    if (tf & DVG_PF_APPROX) sh_done = true;  // Q9: the reference returns after the first candidate
                    out |= DVG_NEED_STROKE;
                } else {
                    bool decided = false;
                    const int ptype = tf & DVG_PF_TYPE_MASK;
                    // conservative early-out for curved segments (exact answer `false`, no root solve)
#ifdef DVG_NO_CAPSULE
                    const bool skip = false; const int cls = 0;
#else
                    const int cls = ((ptype == PRIM_CUBIC || ptype == PRIM_QUAD) && !(tf & DVG_PF_APPROX)) ? capsule_classify(pr.cap, lpt) : 0;
                    const bool skip = cls < 0;   // cls > 0 still runs the exact solve (default mode, see dvg_render.cu)
#endif
#ifdef DVG_CAPSULE_STATS
                    {
                        bool dd = false;
                        const bool ex = prim_stroke_hit(ptype, (tf & DVG_PF_APPROX) != 0, pr.p01, pr.p23, pr.rad, shape_r, lpt, &dd);
                        dvg_capsule_stats(cls, ex);
                        if (cls > 0 && !ex && ptype == PRIM_CUBIC) dvg_capsule_dump(pr.p01, pr.p23, pr.rad, lpt);
                        if (cls > 0 && ptype == PRIM_CUBIC) dvg_cert_stats(pr.p01, pr.p23, pr.cap, lpt, ex);
                    }
#endif
                    const bool h = skip ? false : prim_stroke_hit(ptype, (tf & DVG_PF_APPROX) != 0, pr.p01, pr.p23, pr.rad,
                                                                      shape_r, lpt, &decided);
                    if (decided) sh_done = true;
                    if (h) {
                        sh_hit = true; stroke_hit = true;
                        if (is_q_group && cur_shape == q_shape) hit0 = true;  // within_distance.h:426-429
                    }
                }
            }
        }
        if (s_fill_ok) {
            if ((tf & DVG_PF_SINGLE) || box_ray_intersect(pr.box, lpt)) {  // winding_number.h:162-169
                if (MODE == TM_CLASSIFY) out |= DVG_NEED_FILL;
                else w_shape += prim_winding(tf & DVG_PF_TYPE_MASK, pr.p01, pr.p23, lpt);
            }
        }
        return out;
    }
    DVG_HD void step(const SceneView &sc, const PrimRef &pr) { step<TM_IMMEDIATE>(sc, pr); }

    DVG_HD void finish(const SceneView &sc) { end_group(sc.params); cur_g = -1; }

    // EdgeQuery.hit after compositing (diffvg.cpp:634-643), in closed form: with L the last
    // fragment of the query group, hit = no opaque fragment after L; without such a fragment,
    // hit = (traversal hit) and no opaque fragment at all.
    DVG_HD bool q_hit() const { return seen_q ? !opaque_after_q : (hit0 && !any_opaque); }

    // diffvg.cpp:596-603, 649-653
    DVG_HD F4 resolve(const float *bg_px) const {
        if (nfrag <= 0) {
            if (bg_px) return mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]);
            return mk4(0, 0, 0, 0);
        }
        F4 c = accum;
        if (c.w > 1e-6f) {
            const float inv = 1.f / c.w;  // operator/= multiplies by the reciprocal (vector.h:428-436)
            c.x *= inv; c.y *= inv; c.z *= inv;
        }
        return c;
    }
};

// Sample position of pixel-sample (x, y, sx, sy) with global index idx (diffvg.cpp:1168-1191,
// 539-541): returns the screen-space point and the canvas-space point.
DVG_HD void sample_position(int canvas_w, int canvas_h, int width, int height, int nsx, int nsy, uint64_t seed,
                            bool use_prefiltering, int x, int y, int sx, int sy, int idx, F2 &pt, F2 &cpt) {
    Pcg32 rng = pcg32_init(idx, seed);
    float rx = pcg32_next_float(rng);
    float ry = pcg32_next_float(rng);
    if (use_prefiltering) rx = ry = 0.5f;
    pt = mk2(x + ((float)sx + rx) / nsx, y + ((float)sy + ry) / nsy);
    F2 npt = pt;
    npt.x /= width;
    npt.y /= height;
    cpt = mk2(npt.x * canvas_w, npt.y * canvas_h);
}

}  // namespace dvg
