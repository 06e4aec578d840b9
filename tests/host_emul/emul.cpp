// emul.cpp -- TEST-ONLY host harness.  Compiles the product's host/device arithmetic headers
// (diffvg_b200/csrc/*.cuh) with g++ and drives them with plain loops, so that the geometric
// predicates, the per-sample state machine, the scene-build functions and the boundary
// sampler can be checked against the reference on a machine without a GPU.
// It is NOT a CPU fallback: nothing in the product imports or links it, the kernels'
// orchestration (tiles, shared memory, warp reductions) is not represented here, and the
// product fails loudly when the CUDA library is missing.
#ifdef DVG_CAPSULE_STATS
#include <atomic>
static std::atomic<long long> g_cs[6];   // [cls+1][exact]
static inline void dvg_capsule_stats(int cls, bool ex) { g_cs[(cls + 1) * 2 + (ex ? 1 : 0)]++; }
static std::atomic<long long> g_cert[4];
static std::atomic<long long> g_cert2[4];   // [0] bernstein gmin > 0, [1] sampled g > 0 everywhere, [2] total, [3] unused   // [certificate holds][exact]  (research: DESIGN.md section 10, item 1)
namespace dvg { struct F4; struct F2; }
static void dvg_capsule_dump(const dvg::F4 &p01, const dvg::F4 &p23, const dvg::F4 &rad, const dvg::F2 &pt);
static void dvg_cert_stats(const dvg::F4 &p01, const dvg::F4 &p23, const float *cap, const dvg::F2 &pt, bool ex);
#endif
#include "../../diffvg_b200/csrc/dvg_common.cuh"
#include "../../diffvg_b200/csrc/dvg_scene.cuh"
#include "../../diffvg_b200/csrc/dvg_geom.cuh"
#include "../../diffvg_b200/csrc/dvg_color.cuh"
#include "../../diffvg_b200/csrc/dvg_boundary.cuh"
#include "../../diffvg_b200/csrc/dvg_trace.cuh"
#include "../../diffvg_b200/csrc/dvg_distance.cuh"
#include "../../diffvg_b200/csrc/dvg_buildfn.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

using namespace dvg;

namespace {

struct HostScene {
    std::vector<int> topo;
    std::vector<float> params;
    std::vector<int> inst_group, inst_shape, inst_prim_begin, prim_inst, prim_seg, prim_point_id;
    std::vector<float> shapes_length, shape_r0, seg_cdf, seg_pmf, prim_thick, shape_cdf, shape_pmf;
    std::vector<Box> shape_box, prim_box, prim_cbox, prim_cbox_pf;
    std::vector<int> seg_point_id;
    std::vector<InstInfo> insts;
    std::vector<GroupInfo> groups;
    std::vector<F4> p01, p23, rad, cap;
    std::vector<PrimQuintic> quint;
    std::vector<PrimWindCert> wcert;
    std::vector<PrimMeta> meta;
    int error_flag = 0;
    float total_length = 0;
    SceneView sc;
};

void build(HostScene &hs, const int32_t *topo, const float *params) {
    const int ns = topo[DVG_H_NUM_SHAPES], ng = topo[DVG_H_NUM_GROUPS];
    int topo_len = topo[DVG_H_OFF_GSHAPES] + topo[DVG_H_TOTAL_GSHAPES];
    topo_len = std::max(topo_len, topo[DVG_H_OFF_NCP] + topo[DVG_H_TOTAL_SEGS]);
    hs.topo.assign(topo, topo + topo_len);
    hs.params.assign(params, params + topo[DVG_H_NUM_PARAMS]);
    const int *t = hs.topo.data();
    for (int g = 0; g < ng; g++) {
        const int *r = t + t[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN;
        const int *ids = t + t[DVG_H_OFF_GSHAPES] + r[DVG_G_SHAPES_OFF];
        for (int k = 0; k < r[DVG_G_NUM_SHAPES]; k++) {
            const int *sr = t + t[DVG_H_OFF_SHAPES] + ids[k] * DVG_SHAPE_REC_LEN;
            int inst = (int)hs.inst_group.size();
            hs.inst_group.push_back(g); hs.inst_shape.push_back(ids[k]);
            hs.inst_prim_begin.push_back((int)hs.prim_inst.size());
            if (sr[DVG_S_TYPE] == DVG_SHAPE_PATH) {
                const int *ncp = t + t[DVG_H_OFF_NCP] + sr[DVG_S_NCP_OFF];
                int pid = 0;
                for (int seg = 0; seg < sr[DVG_S_NUM_SEGS]; seg++) {
                    hs.prim_inst.push_back(inst); hs.prim_seg.push_back(seg); hs.prim_point_id.push_back(pid);
                    pid += ncp[seg] + 1;
                }
            } else {
                hs.prim_inst.push_back(inst); hs.prim_seg.push_back(0); hs.prim_point_id.push_back(0);
            }
        }
    }
    hs.inst_prim_begin.push_back((int)hs.prim_inst.size());
    const int ni = (int)hs.inst_group.size(), np = (int)hs.prim_inst.size(), nsg = std::max(t[DVG_H_TOTAL_SEGS], 1);
    hs.shapes_length.resize(ns); hs.shape_box.resize(ns); hs.shape_r0.resize(ns);
    hs.seg_cdf.resize(nsg); hs.seg_pmf.resize(nsg); hs.seg_point_id.resize(nsg);
    hs.insts.resize(ni); hs.groups.resize(ng);
    hs.p01.resize(np); hs.p23.resize(np); hs.rad.resize(np); hs.prim_box.resize(np); hs.prim_thick.resize(np);
    hs.meta.resize(np); hs.prim_cbox.resize(np); hs.prim_cbox_pf.resize(np); hs.cap.resize((size_t)np * DVG_CAP_F4); hs.quint.resize(np); hs.wcert.resize(np); hs.shape_cdf.resize(ni); hs.shape_pmf.resize(ni);
    BuildView bv;
    bv.canvas_w = t[DVG_H_CANVAS_W]; bv.canvas_h = t[DVG_H_CANVAS_H];
    bv.num_shapes = ns; bv.num_groups = ng; bv.num_insts = ni; bv.num_prims = np;
    bv.topo = t; bv.params = hs.params.data();
    bv.inst_group = hs.inst_group.data(); bv.inst_shape = hs.inst_shape.data(); bv.inst_prim_begin = hs.inst_prim_begin.data();
    bv.prim_inst = hs.prim_inst.data(); bv.prim_seg = hs.prim_seg.data(); bv.prim_point_id = hs.prim_point_id.data();
    bv.shapes_length = hs.shapes_length.data(); bv.shape_box = hs.shape_box.data(); bv.shape_r0 = hs.shape_r0.data();
    bv.seg_cdf = hs.seg_cdf.data(); bv.seg_pmf = hs.seg_pmf.data(); bv.seg_point_id = hs.seg_point_id.data();
    bv.insts = hs.insts.data(); bv.groups = hs.groups.data();
    bv.prim_p01 = hs.p01.data(); bv.prim_p23 = hs.p23.data(); bv.prim_rad = hs.rad.data(); bv.prim_box = hs.prim_box.data();
    bv.prim_thick = hs.prim_thick.data(); bv.prim_meta = hs.meta.data(); bv.prim_cbox = hs.prim_cbox.data(); bv.prim_cbox_pf = hs.prim_cbox_pf.data(); bv.prim_cap = hs.cap.data(); bv.prim_quint = hs.quint.data(); bv.prim_wcert = hs.wcert.data();
    bv.shape_cdf = hs.shape_cdf.data(); bv.shape_pmf = hs.shape_pmf.data();
    bv.error_flag = &hs.error_flag; bv.total_length = &hs.total_length; bv.shape_guide = nullptr;
    for (int s = 0; s < ns; s++) build_shape(bv, s);
    for (int g = 0; g < ng; g++) build_group(bv, g);
    for (int e = 0; e < np; e++) build_prim(bv, e);
    float norm = build_shape_cdf_serial(bv);
    for (int i = 0; i < ni; i++) { hs.shape_cdf[i] /= norm; hs.shape_pmf[i] /= norm; }
    SceneView &sc = hs.sc;
    sc.canvas_w = bv.canvas_w; sc.canvas_h = bv.canvas_h;
    sc.num_shapes = ns; sc.num_groups = ng; sc.num_insts = ni; sc.num_prims = np;
    sc.filter.type = t[DVG_H_FILTER_TYPE]; sc.filter.radius = hs.params[t[DVG_H_FILTER_RADIUS_OFF]];
    sc.filter_radius_off = t[DVG_H_FILTER_RADIUS_OFF];
    sc.topo = t; sc.params = hs.params.data();
    sc.prim_p01 = hs.p01.data(); sc.prim_p23 = hs.p23.data(); sc.prim_rad = hs.rad.data(); sc.prim_box = hs.prim_box.data();
    sc.prim_thick = hs.prim_thick.data(); sc.prim_meta = hs.meta.data(); sc.prim_cbox = hs.prim_cbox.data(); sc.prim_cbox_pf = hs.prim_cbox_pf.data(); sc.prim_cap = hs.cap.data(); sc.shape_guide = nullptr; sc.prim_quint = hs.quint.data(); sc.prim_wcert = hs.wcert.data();
    sc.insts = hs.insts.data(); sc.groups = hs.groups.data();
    sc.shapes_length = hs.shapes_length.data(); sc.shape_cdf = hs.shape_cdf.data(); sc.shape_pmf = hs.shape_pmf.data();
    sc.seg_cdf = hs.seg_cdf.data(); sc.seg_pmf = hs.seg_pmf.data(); sc.seg_point_id = hs.seg_point_id.data();
    sc.error_flag = &hs.error_flag;
}

PrimRef prim_ref(const HostScene &hs, int e) {
    PrimRef pr;
    pr.p01 = hs.p01[e]; pr.p23 = hs.p23[e]; pr.rad = hs.rad[e]; pr.box = hs.prim_box[e]; pr.thick = hs.prim_thick[e];
    pr.tf = hs.meta[e].type_flags; pr.inst = hs.meta[e].inst; pr.group = hs.insts[hs.meta[e].inst].group;
    pr.base_id = hs.meta[e].base_id; pr.point_id = hs.meta[e].point_id;
    pr.cap = reinterpret_cast<const float *>(&hs.cap[(size_t)e * DVG_CAP_F4]);
    return pr;
}

// candidate primitives for a canvas-space rectangle (mirrors the tile-bin test of dvg_build.cu)
void candidates(const HostScene &hs, float x0, float y0, float x1, float y1, std::vector<int> &out, bool pf = false) {
    out.clear();
    for (int e = 0; e < hs.sc.num_prims; e++) {
        const Box &b = pf ? hs.prim_cbox_pf[e] : hs.prim_cbox[e];
        if (b.x0 <= x1 && b.x1 >= x0 && b.y0 <= y1 && b.y1 >= y0) out.push_back(e);
    }
}

template <typename F>
void parallel_rows(int n, int nthreads, F f) {
    std::vector<std::thread> th;
    std::atomic<int> next{0};
    for (int t = 0; t < nthreads; t++) th.emplace_back([&] { for (int y; (y = next++) < n;) f(y); });
    for (auto &t : th) t.join();
}

}  // namespace

#define EXPORT extern "C" __attribute__((visibility("default")))

// Forward (d_image == null) or backward.  Images are host arrays.  Backward accumulation is
// serialised with a mutex per call to keep the harness simple.
// Gradients are accumulated in DOUBLE here (the per-sample terms are the product's float
// arithmetic; only the summation is wider), so this harness is also the accuracy yardstick for
// sums where the reference's sequential float atomics saturate (e.g. d_filter.radius: 4 M terms of
// ~1e-8 added one by one to a float of magnitude ~1).
struct DoubleSink {
    double *D;
    void add(int idx, float v) const { D[idx] += (double)v; }
};
static float *g_debug_out = nullptr;
EXPORT void emul_set_boundary_dump(float *buf) { g_debug_out = buf; }

EXPORT int emul_render(const int32_t *topo, const float *params, const float *background, float *image,
                       int W, int H, int nsx, int nsy, uint64_t seed, const float *d_image,
                       float *d_params, float *d_background, int skip_xform_grad, int nthreads) {
    HostScene hs;
    build(hs, topo, params);
    if (hs.error_flag) return 3;
    const SceneView &sc = hs.sc;
    const int spp = nsx * nsy;
    std::vector<float> weight((size_t)W * H, 0.f);
    const int ri = (int)ceilf(sc.filter.radius);
    for (int idx = 0; idx < W * H * spp; idx++) {  // weight_kernel (diffvg.cpp:1115-1158)
        const int sx = idx % nsx, sy = (idx / nsx) % nsy, x = (idx / spp) % W, y = idx / (spp * W);
        F2 pt, cpt;
        sample_position(sc.canvas_w, sc.canvas_h, W, H, nsx, nsy, seed, false, x, y, sx, sy, idx, pt, cpt);
        for (int dy = -ri; dy <= ri; dy++)
            for (int dx = -ri; dx <= ri; dx++) {
                int xx = x + dx, yy = y + dy;
                if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                    weight[yy * W + xx] += filter_weight(sc.filter, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
            }
    }
    const float cw = (float)sc.canvas_w, ch = (float)sc.canvas_h;
    const float margin = 4e-4f * std::max(cw, ch) + 1e-4f;
    std::mutex mu;
    RenderArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.width = W; ra.height = H; ra.nsx = nsx; ra.nsy = nsy; ra.seed = seed;
    ra.background = background; ra.d_render_image = d_image; ra.d_params = d_params; ra.d_background = d_background;
    ra.weight_image = weight.data();
    ra.flags = skip_xform_grad ? 1u : 0u;

    std::vector<double> acc(d_params ? hs.params.size() : 0, 0.0);
    const DoubleSink dsink{acc.data()};
    parallel_rows(H, nthreads, [&](int y) {
        std::vector<int> cand;
        std::vector<int> fkey(DVG_MAXF);
        std::vector<F4> fprev(DVG_MAXF);
        std::vector<float> row(d_image ? 0 : (size_t)(2 * ri + 1) * W * 4, 0.f);  // rows y-ri..y+ri
        for (int x = 0; x < W; x++) {
            candidates(hs, (float)x / W * cw - margin, (float)y / H * ch - margin, (float)(x + 1) / W * cw + margin,
                       (float)(y + 1) / H * ch + margin, cand);
            for (int s = 0; s < spp; s++) {
                const int sx = s % nsx, sy = s / nsx;
                const int idx = ((y * W + x) * nsy + sy) * nsx + sx;
                F2 pt, cpt;
                sample_position(sc.canvas_w, sc.canvas_h, W, H, nsx, nsy, seed, false, x, y, sx, sy, idx, pt, cpt);
                const float *bg_px = background ? background + 4 * (y * W + x) : nullptr;
                F4 first = bg_px ? mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]) : mk4(0, 0, 0, 0);
                if (!d_image) {
                    SampleTracer<false, false> tr;
                    tr.init(cpt, true, first, -1, -1, nullptr, nullptr);
                    for (int e : cand) tr.step(sc, prim_ref(hs, e));
                    tr.finish(sc);
                    const F4 color = tr.resolve(bg_px);
                    for (int dy = -ri; dy <= ri; dy++)
                        for (int dx = -ri; dx <= ri; dx++) {
                            int xx = x + dx, yy = y + dy;
                            if (xx >= 0 && xx < W && yy >= 0 && yy < H && weight[yy * W + xx] > 0) {
                                float fw = filter_weight(sc.filter, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
                                float inv_ws = 1.f / weight[yy * W + xx];
                                float *d = &row[((size_t)(dy + ri) * W + xx) * 4];
                                d[0] += (fw * color.x) * inv_ws; d[1] += (fw * color.y) * inv_ws;
                                d[2] += (fw * color.z) * inv_ws; d[3] += (fw * color.w) * inv_ws;
                            }
                        }
                } else {
                    SampleTracer<false, true> tr;
                    tr.init(cpt, true, first, -1, -1, fkey.data(), fprev.data());
                    for (int e : cand) tr.step(sc, prim_ref(hs, e));
                    tr.finish(sc);
                    const F4 color = tr.resolve(bg_px);
                    const F4 d_color = gather_d_color(sc.filter, d_image, weight.data(), W, H, pt);
                    std::lock_guard<std::mutex> lk(mu);
                    float dcr = d_color.x, dcg = d_color.y, dcb = d_color.z, dca = d_color.w;
                    if (tr.nfrag > 0) {
                        if (tr.accum.w > 1e-6f) {
                            const float inv = 1.f / tr.accum.w;
                            dca -= (d_color.x * color.x + d_color.y * color.y + d_color.z * color.z) / tr.accum.w;
                            dcr = d_color.x * inv; dcg = d_color.y * inv; dcb = d_color.z * inv;
                        }
                        for (int i = tr.sp - 1; i >= 0; i--) {
                            const int key = fkey[i];
                            const GroupInfo &g = sc.groups[key >> 1];
                            const int ctype = (key & 1) ? g.stroke_type : g.fill_type;
                            const int coff = (key & 1) ? g.stroke_off : g.fill_off;
                            const int cstops = (key & 1) ? g.stroke_stops : g.fill_stops;
                            const F4 prev = fprev[i];
                            const F4 fc = eval_color(ctype, sc.params + coff, cstops, cpt);
                            const float d_prev_alpha = dca * (1.f - fc.w);
                            float d_alpha_i = dca * (1.f - prev.w);
                            d_alpha_i += (dcr * (fc.x - prev.x) + dcg * (fc.y - prev.y)) + dcb * (fc.z - prev.z);
                            F4 dc = mk4(dcr * fc.w, dcg * fc.w, dcb * fc.w, d_alpha_i);
                            dcr = dcr * (1 - fc.w); dcg = dcg * (1 - fc.w); dcb = dcb * (1 - fc.w);
                            dca = d_prev_alpha;
                            if (ctype == 0) { for (int k = 0; k < 4; k++) acc[coff + k] += (double)(&dc.x)[k]; }
                            else if (!(key & 1)) d_eval_gradient(ctype, sc.params + coff, cstops, cpt, dc, dsink, coff, nullptr);
                        }
                        if (bg_px && d_background) {
                            float *d = d_background + 4 * (y * W + x);
                            d[0] += dcr; d[1] += dcg; d[2] += dcb; d[3] += dca;
                        }
                    } else if (bg_px && d_background) {
                        float *d = d_background + 4 * (y * W + x);
                        d[0] += d_color.x; d[1] += d_color.y; d[2] += d_color.z; d[3] += d_color.w;
                    }
                    for (int dy = -ri; dy <= ri; dy++)
                        for (int dx = -ri; dx <= ri; dx++) {
                            int xx = x + dx, yy = y + dy;
                            if (xx >= 0 && xx < W && yy >= 0 && yy < H && weight[yy * W + xx] > 0) {
                                const float ws = weight[yy * W + xx];
                                const float ddx = (xx + 0.5f) - pt.x, ddy = (yy + 0.5f) - pt.y;
                                const float fw = filter_weight(sc.filter, ddx, ddy);
                                const float *dp = d_image + 4 * (yy * W + xx);
                                const float dotv = dp[0] * color.x + dp[1] * color.y + dp[2] * color.z + dp[3] * color.w;
                                const float d_weight = (dotv * ws - fw * dotv * (ws - fw)) / (ws * ws);
                                acc[sc.filter_radius_off] += (double)d_filter_weight_radius(sc.filter, ddx, ddy, d_weight);
                            }
                        }
                }
            }
        }
        if (!d_image) {
            std::lock_guard<std::mutex> lk(mu);
            for (int dy = -ri; dy <= ri; dy++) {
                int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int i = 0; i < W * 4; i++) image[(size_t)yy * W * 4 + i] += row[(size_t)(dy + ri) * W * 4 + i];
            }
        }
    });

    if (d_image) {
        // boundary term (diffvg.cpp:1558-1626), one boundary sample per pixel sample index
        const int n = W * H * spp;
        parallel_rows((n + 4095) / 4096, nthreads, [&](int blk) {
            std::vector<int> cand;
            for (int idx = blk * 4096; idx < std::min(n, (blk + 1) * 4096); idx++) {
                BoundarySample bs;
                make_boundary_sample(sc, idx, seed, bs);
                if (bs.inst < 0) continue;
                const int bx = (int)(bs.pt.x * W), by = (int)(bs.pt.y * H);
                if (bx < 0 || bx >= W || by < 0 || by >= H) continue;
                const InstInfo &ii = sc.insts[bs.inst];
                const float px = bs.pt.x * cw, py = bs.pt.y * ch;
                candidates(hs, px - 2 * margin, py - 2 * margin, px + 2 * margin, py + 2 * margin, cand);
                const float *bg_px = background ? background + 4 * (by * W + bx) : nullptr;
                F4 first = bg_px ? mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]) : mk4(0, 0, 0, 0);
                F4 col[2]; bool hit[2];
                for (int side = 0; side < 2; side++) {
                    const F2 off = 1e-4f * bs.normal;
                    const F2 npt = side ? bs.pt + off : bs.pt - off;
                    SampleTracer<true, false> tr;
                    tr.init(mk2(npt.x * sc.canvas_w, npt.y * sc.canvas_h), true, first, ii.group, ii.shape, nullptr, nullptr);
                    for (int e : cand) tr.step(sc, prim_ref(hs, e));
                    tr.finish(sc);
                    col[side] = tr.resolve(bg_px); hit[side] = tr.q_hit();
                }
                if (!hit[0] && !hit[1]) continue;
                F4 c_in = col[0], c_out = col[1];
                F2 normal = bs.normal;
                if (!hit[0]) { normal = -normal; c_in = col[1]; c_out = col[0]; }
                F4 d_color = gather_d_color(sc.filter, d_image, weight.data(), W, H, mk2(bs.pt.x * W, bs.pt.y * H));
                d_color = d_color * (1.f / (float)(sc.canvas_w * sc.canvas_h));
                const F4 diff = c_in - c_out;
                const float contrib = (diff.x * d_color.x + diff.y * d_color.y + diff.z * d_color.z + diff.w * d_color.w) / bs.pdf;
                std::lock_guard<std::mutex> lk(mu);
                if (g_debug_out) {
                    float *o = g_debug_out + 4 * (size_t)idx;
                    o[0] = contrib; o[1] = (float)((hit[0] ? 1 : 0) | (hit[1] ? 2 : 0)); o[2] = normal.x; o[3] = normal.y;
                }
                accumulate_boundary_gradient(sc, ra, bs, ii, sc.groups[ii.group], contrib, normal, dsink);
                if (!skip_xform_grad) {
                    float dm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                    boundary_xform_gradient(bs, sc.groups[ii.group], contrib, normal, dm);
                    for (int k = 0; k < 9; k++) acc[sc.groups[ii.group].xform_off + k] += (double)dm[k];
                }
            }
        });
    }
    if (d_params) for (size_t i = 0; i < acc.size(); i++) d_params[i] = (float)acc[i];
    return 0;
}

// Scene-build tables for the bit-exact CDF / indexing checks (same selectors as dvg_scene_dump
// where applicable: 3 shapes_length, 4 shape cdf, 5 shape pmf, 6/7/8 per-path tables).
EXPORT int64_t emul_scene_dump(const int32_t *topo, const float *params, int what, int index, uint32_t *out, int64_t cap) {
    HostScene hs;
    build(hs, topo, params);
    std::vector<uint32_t> w;
    auto pushf = [&](float f) { uint32_t u; memcpy(&u, &f, 4); w.push_back(u); };
    const int *sr = hs.topo.data() + hs.topo[DVG_H_OFF_SHAPES] + index * DVG_SHAPE_REC_LEN;
    switch (what) {
        case 3: for (float f : hs.shapes_length) pushf(f); break;
        case 4: for (float f : hs.shape_cdf) pushf(f); break;
        case 5: for (float f : hs.shape_pmf) pushf(f); break;
        case 6: for (int i = 0; i < sr[DVG_S_NUM_SEGS]; i++) pushf(hs.seg_cdf[sr[DVG_S_NCP_OFF] + i]); break;
        case 7: for (int i = 0; i < sr[DVG_S_NUM_SEGS]; i++) pushf(hs.seg_pmf[sr[DVG_S_NCP_OFF] + i]); break;
        case 8: for (int i = 0; i < sr[DVG_S_NUM_SEGS]; i++) w.push_back((uint32_t)hs.seg_point_id[sr[DVG_S_NCP_OFF] + i]); break;
        default: return -1;
    }
    if ((int64_t)w.size() > cap) return -1;
    memcpy(out, w.data(), w.size() * 4);
    return (int64_t)w.size();
}

// Counterpart of dvg_debug_prim_tests (csrc/dvg_debug.cu): the same raw per-primitive predicates on the host.
EXPORT void emul_debug_prim_tests(const int32_t *topo, const float *params, int W, int H, int nsx, int nsy, uint64_t seed, int x, int y,
                                  int32_t *out, float *pos) {
    HostScene hs;
    build(hs, topo, params);
    const SceneView &sc = hs.sc;
    for (int s = 0; s < nsx * nsy; s++) {
        const int sx = s % nsx, sy = s / nsx;
        const int idx = ((y * W + x) * nsy + sy) * nsx + sx;
        F2 pt, cpt;
        sample_position(sc.canvas_w, sc.canvas_h, W, H, nsx, nsy, seed, false, x, y, sx, sy, idx, pt, cpt);
        pos[2 * s] = cpt.x; pos[2 * s + 1] = cpt.y;
        for (int e = 0; e < sc.num_prims; e++) {
            const PrimMeta pm = sc.prim_meta[e];
            const InstInfo &ii = sc.insts[pm.inst];
            const GroupInfo &g = sc.groups[ii.group];
            const F2 lp = (g.flags & DVG_GF_IDENTITY) ? cpt : xform_pt(g.c2s, cpt);
            const int type = pm.type_flags & DVG_PF_TYPE_MASK;
            int r = 0;
            if (g.stroke_type >= 0) {
                bool decided = false;
                r |= 2;
                if (type != PRIM_ELLIPSE && prim_stroke_hit(type, (pm.type_flags & DVG_PF_APPROX) != 0, sc.prim_p01[e], sc.prim_p23[e], sc.prim_rad[e], ii.r, lp, &decided)) r |= 1;
            }
            if (g.fill_type >= 0) {
                r |= 4;
                r |= (prim_winding(type, sc.prim_p01[e], sc.prim_p23[e], lp) & 0xff) << 8;
            }
            out[(size_t)s * sc.num_prims + e] = r;
        }
    }
}

// dvg_crmath.cuh on the host, for the accuracy test against 200-bit arithmetic (which: 0 cos, 1 acos, 2 pow(x, 1./3.))
EXPORT void emul_crmath(int which, const double *x, double *y, int n) {
    for (int i = 0; i < n; i++) y[i] = which == 0 ? cr_cos(x[i]) : (which == 1 ? cr_acos(x[i]) : cr_pow13(x[i]));
}
EXPORT int emul_solve_cubic(int accurate, double a, double b, double c, double d, double *t) {
    return accurate ? solve_cubic_cr(a, b, c, d, t) : solve_cubic_d(a, b, c, d, t);
}

// PCG known-answer helper: state after init and the first two floats.
// Test support for the tight tile binning (dvg_buildfn.cuh bracket_reaches_tile): brackets of one cubic stroke,
// the tile test, and the per-point bracket classification the kernels use.
EXPORT void emul_bracket_build(const float *pts8, float rmax, float rmin, float *cap_out) {
    build_capsules(PRIM_CUBIC, mk4(pts8[0], pts8[1], pts8[2], pts8[3]), mk4(pts8[4], pts8[5], pts8[6], pts8[7]), rmax, rmin, cap_out);
}
EXPORT int emul_bracket_reaches_tile(const float *cap, float x0, float y0, float x1, float y1) {
    return bracket_reaches_tile(reinterpret_cast<const F4 *>(cap), x0, y0, x1, y1) ? 1 : 0;
}
EXPORT int emul_bracket_classify(const float *cap, float x, float y) { return capsule_classify(cap, mk2(x, y)); }
EXPORT int emul_stroke_hit_cubic(const float *pts8, const float *rad4, float x, float y) {
    return stroke_hit_cubic(mk2(pts8[0], pts8[1]), mk2(pts8[2], pts8[3]), mk2(pts8[4], pts8[5]), mk2(pts8[6], pts8[7]),
                            mk4(rad4[0], rad4[1], rad4[2], rad4[3]), mk2(x, y)) ? 1 : 0;
}

// The per-primitive split of the closest-point quintic (dvg_geom.cuh prim_quintic / quintic_of / quintic_intervals_of)
// against cubic_quintic / quintic_intervals on the same inputs.  out[0]: pairs whose five normalised coefficients are not
// bit-identical; out[1]: pairs whose split points inside [0, 1] differ (count or any float); out[2]: pairs whose verdict through the
// kernel's bracket enumeration differs from stroke_hit_cubic's.
EXPORT void emul_quintic_split_check(const float *pts8, const float *rad4, const float *xy, int n, long long *out) {
    const F2 p0 = mk2(pts8[0], pts8[1]), p1 = mk2(pts8[2], pts8[3]), p2 = mk2(pts8[4], pts8[5]), p3 = mk2(pts8[6], pts8[7]);
    const F4 r = mk4(rad4[0], rad4[1], rad4[2], rad4[3]);
    const PrimQuintic k = prim_quintic(p0, p1, p2, p3);
    out[0] = out[1] = out[2] = 0;
    for (int i = 0; i < n; i++) {
        const F2 pt = mk2(xy[2 * i], xy[2 * i + 1]);
        const Quintic a = cubic_quintic(p0, p1, p2, p3, pt), b = quintic_of(k, p0, pt);
        if (memcmp(&a, &b, sizeof a) != 0) out[0]++;
        float ia[4] = {0, 0, 0, 0}, ib[4] = {0, 0, 0, 0};
        const int na = quintic_intervals(a, ia), nb = quintic_intervals_of(k, b, ib);
        // only split points inside [0, 1] matter (negative: skipped; above 1: the bracket ends at 1): the per-primitive form
        // leaves roots far outside unpolished
        for (int j = 0; j < 4; j++) {
            if (ia[j] < 0.f) ia[j] = -1.f; else if (ia[j] >= 1.f) ia[j] = 1.f;
            if (ib[j] < 0.f) ib[j] = -1.f; else if (ib[j] >= 1.f) ib[j] = 1.f;
        }
        if (na != nb || memcmp(ia, ib, sizeof(float) * na) != 0) out[1]++;
        // the kernel's enumeration: brackets from the signs at the split points, answer = OR of the radius tests
        bool hit = dist_sq(p0, pt) < r.x * r.x || dist_sq(p3, pt) < r.w * r.w;
        if (!hit) {
            float lower = 0.f;
            for (int j = 0; j < nb + 1 && !hit; j++) {
                if (j < nb && ib[j] < 0.f) continue;
                const float upper = j < nb ? rminf(ib[j], 1.f) : 1.f;
                float t;
                if (quintic_root_in(b, lower, upper, &t)) {
                    const float tt = 1 - t;
                    const float rr = (tt * tt * tt) * r.x + (3 * tt * tt * t) * r.y + (3 * tt * t * t) * r.z + (t * t * t) * r.w;
                    if (dist_sq(eval_cubic(p0, p1, p2, p3, t), pt) < rr * rr) hit = true;
                    if (upper >= 1.f) break;
                    lower = upper;
                }
            }
        }
        if (hit != stroke_hit_cubic(p0, p1, p2, p3, r, pt)) out[2]++;
    }
}

// cubic_winding_fast against cubic_winding_exact (dvg_geom.cuh): out[0] = pairs the fast form answered, out[1] = answered
// pairs whose winding differs from the reference's operation sequence; out[2], out[3]: the same for the classifier's
// per-primitive certificate (prim_wind_cert + wind_cert_answer).
EXPORT void emul_winding_fast_check(const float *pts8, const float *xy, int n, long long *out) {
    const F2 p0 = mk2(pts8[0], pts8[1]), p1 = mk2(pts8[2], pts8[3]), p2 = mk2(pts8[4], pts8[5]), p3 = mk2(pts8[6], pts8[7]);
    out[0] = out[1] = out[2] = out[3] = 0;
    PrimWindCert wc;
    const int flags = prim_wind_cert(p0, p1, p2, p3, wc);      // the classifier's per-primitive certificate (dvg_wave.cu)
    const float box_x0 = rminf(rminf(p0.x, p1.x), rminf(p2.x, p3.x));
    for (int i = 0; i < n; i++) {
        const F2 pt = mk2(xy[2 * i], xy[2 * i + 1]);
        int wf = 0;
        if (cubic_winding_fast(p0, p1, p2, p3, pt, &wf)) {
            out[0]++;
            if (wf != cubic_winding_exact(p0, p1, p2, p3, pt)) out[1]++;
        }
        int wcw = 0;
        if ((flags & DVG_PF_YMONO) && wind_cert_answer(wc, (flags & DVG_PF_YUP) != 0, box_x0, pt, &wcw)) {
            out[2]++;
            if (wcw != cubic_winding_exact(p0, p1, p2, p3, pt)) out[3]++;
        }
    }
}

// cdf_sample_guided (the sample-generation kernel's shape pick) against cdf_sample: returns the number of u for which
// the two differ.  The guide table is built the way k_build_shape_cdf builds it.
EXPORT long long emul_cdf_guided_check(const float *cdf, int n, const float *us, int m) {
    std::vector<int> guide(DVG_CDF_GUIDE + 1);
    for (int k = 0; k <= DVG_CDF_GUIDE; k++) guide[k] = cdf_sample(cdf, n, (float)k / (float)DVG_CDF_GUIDE, nullptr);
    long long bad = 0;
    for (int i = 0; i < m; i++)
        if (cdf_sample_guided(cdf, n, us[i], guide.data()) != cdf_sample(cdf, n, us[i], nullptr)) bad++;
    return bad;
}

EXPORT void emul_pcg(int idx, uint64_t seed, uint64_t *state, float *rx, float *ry) {
    Pcg32 r = pcg32_init(idx, seed);
    *state = r.state;
    *rx = pcg32_next_float(r);
    *ry = pcg32_next_float(r);
}

// Debug: per-sample report for one pixel (which primitives' stroke tests hit, final colour) and
// a scan of the reference's uninitialised intervals[0] (Q10) over [0,1].
EXPORT void emul_debug_pixel(const int32_t *topo, const float *params, int W, int H, int nsx, int nsy, uint64_t seed, int x, int y) {
    HostScene hs;
    build(hs, topo, params);
    const SceneView &sc = hs.sc;
    const float cw = (float)sc.canvas_w, ch = (float)sc.canvas_h;
    const float margin = 4e-4f * std::max(cw, ch) + 1e-4f;
    std::vector<int> cand;
    candidates(hs, (float)x / W * cw - margin, (float)y / H * ch - margin, (float)(x + 1) / W * cw + margin, (float)(y + 1) / H * ch + margin, cand);
    printf("pixel %d %d: %d candidates\n", x, y, (int)cand.size());
    for (int s = 0; s < nsx * nsy; s++) {
        const int sx = s % nsx, sy = s / nsx;
        const int idx = ((y * W + x) * nsy + sy) * nsx + sx;
        F2 pt, cpt;
        sample_position(sc.canvas_w, sc.canvas_h, W, H, nsx, nsy, seed, false, x, y, sx, sy, idx, pt, cpt);
        SampleTracer<false, false> tr;
        tr.init(cpt, true, mk4(0, 0, 0, 0), -1, -1, nullptr, nullptr);
        for (int e : cand) tr.step(sc, prim_ref(hs, e));
        tr.finish(sc);
        F4 c = tr.resolve(nullptr);
        printf(" sample %2d idx %d pt (%.6f %.6f) nfrag %d color %.6f %.6f %.6f %.6f\n", s, idx, cpt.x, cpt.y, tr.nfrag, c.x, c.y, c.z, c.w);
        for (int e : cand) {
            PrimRef pr = prim_ref(hs, e);
            if ((pr.tf & DVG_PF_TYPE_MASK) != PRIM_CUBIC) continue;
            if (!((pr.tf & DVG_PF_SINGLE) || box_inside_r(pr.box, cpt, pr.thick))) continue;
            F2 p0 = mk2(pr.p01.x, pr.p01.y), p1 = mk2(pr.p01.z, pr.p01.w), p2 = mk2(pr.p23.x, pr.p23.y), p3 = mk2(pr.p23.z, pr.p23.w);
            bool base = stroke_hit_cubic(p0, p1, p2, p3, pr.rad, cpt);
            Quintic q = cubic_quintic(p0, p1, p2, p3, cpt);
            double q_root = -q.B / 5.f;
            int flips = 0; float first_flip = -1;
            if (!(q_root >= 0 && q_root <= 1)) {
                for (int k = 0; k <= 2000; k++) {
                    float st = k / 2000.f;
                    if (stroke_hit_cubic(p0, p1, p2, p3, pr.rad, cpt, st) != base) { if (!flips) first_flip = st; flips++; }
                }
            }
            float iv[4]; int n = quintic_intervals(q, iv);
            printf("   prim %d group %d hit %d q_root %.4f intervals(%d) %.5f %.5f %.5f %.5f stale-flips %d (first %.4f)\n", e, pr.group, (int)base, q_root, n, iv[0], n > 1 ? iv[1] : 9.f, n > 2 ? iv[2] : 9.f, n > 3 ? iv[3] : 9.f, flips, first_flip);
        }
    }
}

EXPORT void emul_debug_boundary(const int32_t *topo, const float *params, int W, int H, uint64_t seed, int idx) {
    HostScene hs;
    build(hs, topo, params);
    const SceneView &sc = hs.sc;
    BoundarySample bs;
    make_boundary_sample(sc, idx, seed, bs);
    printf("idx %d inst %d pt (%.7f %.7f) -> px (%.5f %.5f) normal (%.6f %.6f) pdf %g path_t %g base %d pid %d stroke %d\n", idx, bs.inst,
           bs.pt.x, bs.pt.y, bs.pt.x * W, bs.pt.y * H, bs.normal.x, bs.normal.y, bs.pdf, bs.path_t, bs.base_point_id,
           bs.point_id_stroke & 0x7fffffff, bs.point_id_stroke < 0);
    if (bs.inst < 0) return;
    const InstInfo &ii = sc.insts[bs.inst];
    printf("  group %d shape %d prim_begin %d\n", ii.group, ii.shape, ii.prim_begin);
    const float cw = (float)sc.canvas_w, ch = (float)sc.canvas_h;
    for (int side = 0; side < 2; side++) {
        const F2 off = 1e-4f * bs.normal;
        const F2 npt = side ? bs.pt + off : bs.pt - off;
        F2 cpt = mk2(npt.x * sc.canvas_w, npt.y * sc.canvas_h);
        for (int e = 0; e < sc.num_prims; e++) {
            PrimRef pr = prim_ref(hs, e);
            if (pr.group != ii.group) continue;
            bool inb = (pr.tf & DVG_PF_SINGLE) || box_inside_r(pr.box, cpt, pr.thick);
            bool dec = false;
            bool h = prim_stroke_hit(pr.tf & DVG_PF_TYPE_MASK, false, pr.p01, pr.p23, pr.rad, ii.r, cpt, &dec);
            const Box &cb = hs.prim_cbox[e];
            printf("  side %d cpt (%.6f %.6f) prim %d inbox %d hit %d caprej %d cbox (%.4f %.4f %.4f %.4f)\n", side, cpt.x, cpt.y, e, (int)inb, (int)h,
                   (int)capsule_reject(pr.cap, cpt), cb.x0, cb.y0, cb.x1, cb.y1);
        }
    }
}

#ifdef DVG_CAPSULE_STATS
// RESEARCH (not used by the product): a per-pair certificate that the reference's closest-point solve succeeds for a
// sample the bracket proves inside.  g(t) = |q'(t)|^2 + (q(t) - p).q''(t) is the derivative of the quintic
// f(t) = (q(t) - p).q'(t); if its Bernstein coefficients on [0,1] are all > 0, f is strictly increasing: one root, no
// near-double root for Newton's |f| < 1e-5 stop to land on.  The stop then leaves |t - t*| <= 1e-5 A / min g, i.e. a
// displacement of at most max|q'| * that along the curve, which must stay inside the slack the bracket leaves.
static bool cert_monotone(const F4 &p01, const F4 &p23, const float *cap, const F2 &pt) {
    const double P[4][2] = {{p01.x, p01.y}, {p01.z, p01.w}, {p23.x, p23.y}, {p23.z, p23.w}};
    double c[4][2], d[3][2], e[2][2];
    for (int i = 0; i < 4; i++) { c[i][0] = P[i][0] - pt.x; c[i][1] = P[i][1] - pt.y; }
    for (int j = 0; j < 3; j++) { d[j][0] = 3 * (P[j + 1][0] - P[j][0]); d[j][1] = 3 * (P[j + 1][1] - P[j][1]); }
    for (int k = 0; k < 2; k++) { e[k][0] = 6 * (P[k + 2][0] - 2 * P[k + 1][0] + P[k][0]); e[k][1] = 6 * (P[k + 2][1] - 2 * P[k + 1][1] + P[k][1]); }
    const double C2[3] = {1, 2, 1}, C3[4] = {1, 3, 3, 1}, C1[2] = {1, 1}, C4[5] = {1, 4, 6, 4, 1};
    double gmin = 1e300, dmax = 0;
    for (int k = 0; k <= 4; k++) {
        double g = 0;
        for (int i = 0; i <= 2; i++) { int j = k - i; if (j < 0 || j > 2) continue; g += C2[i] * C2[j] / C4[k] * (d[i][0] * d[j][0] + d[i][1] * d[j][1]); }
        for (int i = 0; i <= 3; i++) { int j = k - i; if (j < 0 || j > 1) continue; g += C3[i] * C1[j] / C4[k] * (c[i][0] * e[j][0] + c[i][1] * e[j][1]); }
        gmin = g < gmin ? g : gmin;
    }
    for (int j = 0; j < 3; j++) { double l = sqrt(d[j][0] * d[j][0] + d[j][1] * d[j][1]); dmax = l > dmax ? l : dmax; }
    g_cert2[2]++;
    {
        bool pos = true;
        for (int i = 0; i <= 64 && pos; i++) {
            const double t = i / 64.0, u = 1 - t;
            const double qx = u*u*u*c[0][0] + 3*u*u*t*c[1][0] + 3*u*t*t*c[2][0] + t*t*t*c[3][0], qy = u*u*u*c[0][1] + 3*u*u*t*c[1][1] + 3*u*t*t*c[2][1] + t*t*t*c[3][1];
            const double dx = u*u*d[0][0] + 2*u*t*d[1][0] + t*t*d[2][0], dy = u*u*d[0][1] + 2*u*t*d[1][1] + t*t*d[2][1];
            const double ex = u*e[0][0] + t*e[1][0], ey = u*e[0][1] + t*e[1][1];
            pos = dx*dx + dy*dy + qx*ex + qy*ey > 0;
        }
        if (pos) g_cert2[1]++;
    }
    if (!(gmin > 0)) return false;
    g_cert2[0]++;
    const double q3x = -P[0][0] + 3 * P[1][0] - 3 * P[2][0] + P[3][0], q3y = -P[0][1] + 3 * P[1][1] - 3 * P[2][1] + P[3][1];
    const double A = 3 * (q3x * q3x + q3y * q3y);
    double slack = 0;   // r_min - (distance to chord + deviation), best piece
    for (int i = 0; i < DVG_CAP_N; i++) {
        const float *cc = cap + 8 * i;
        const double wx = pt.x - cc[0], wy = pt.y - cc[1];
        double t = (wx * cc[2] + wy * cc[3]) * cc[4];
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
        const double ex = wx - t * cc[2], ey = wy - t * cc[3];
        const double di = sqrt(ex * ex + ey * ey);
        if (cc[6] > 0) { const double sl = sqrt((double)cc[6]) - di + 1e-2; slack = sl > slack ? sl : slack; }
    }
    return dmax * 1e-5 * A / gmin < 0.5 * slack;
}
static std::atomic<int> g_dump_left{12};
static void dvg_capsule_dump(const F4 &p01, const F4 &p23, const F4 &rad, const F2 &pt) {
    if (g_dump_left-- <= 0) return;
    F2 p0 = mk2(p01.x, p01.y), p1 = mk2(p01.z, p01.w), p2 = mk2(p23.x, p23.y), p3 = mk2(p23.z, p23.w);
    Quintic q = cubic_quintic(p0, p1, p2, p3, pt);
    float iv[4]; int n = quintic_intervals(q, iv);
    // brute force closest point
    double best = 1e30, bt = 0;
    for (int i = 0; i <= 100000; i++) { float t = i / 100000.f; F2 e = eval_cubic(p0, p1, p2, p3, t); double d = sqrt((double)dist_sq(e, pt)); if (d < best) { best = d; bt = t; } }
    printf("FAIL pts (%.4f %.4f)(%.4f %.4f)(%.4f %.4f)(%.4f %.4f) r %.4f pt (%.5f %.5f) brute d %.5f at t %.5f | q_root %.5f intervals", p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y, rad.x, pt.x, pt.y, best, bt, -q.B / 5.f);
    for (int j = 0; j < n; j++) printf(" %.6f", iv[j]);
    printf(" | coeffs B %.4g C %.4g D %.4g E %.4g F %.4g\n   evals:", q.B, q.C, q.D, q.E, q.F);
    for (int j = 0; j <= 20; j++) printf(" %.3g", quintic_eval(q, j / 20.0));
    printf("\n   brackets:");
    float lower = 0.f;
    for (int j = 0; j < n + 1; j++) {
        if (j < n && iv[j] < 0.f) continue;
        float upper = j < n ? rminf(iv[j], 1.f) : 1.f;
        float t; bool ok = quintic_root_in(q, lower, upper, &t);
        printf(" [%.6f,%.6f]->%s", lower, upper, ok ? "root" : "none");
        if (ok) { printf("(t=%.6f d=%.5f)", t, sqrt((double)dist_sq(eval_cubic(p0, p1, p2, p3, t), pt))); if (upper >= 1.f) break; lower = upper; }
    }
    printf("\n");
}
static void dvg_cert_stats(const F4 &p01, const F4 &p23, const float *cap, const F2 &pt, bool ex) {
    g_cert[(cert_monotone(p01, p23, cap, pt) ? 2 : 0) + (ex ? 1 : 0)]++;
}
EXPORT void emul_capsule_stats(long long *out, int reset) {
    for (int i = 0; i < 6; i++) { out[i] = g_cs[i]; if (reset) g_cs[i] = 0; }
    for (int i = 0; i < 4; i++) { out[6 + i] = g_cert[i]; if (reset) g_cert[i] = 0; }
    for (int i = 0; i < 4; i++) { out[10 + i] = g_cert2[i]; if (reset) g_cert2[i] = 0; }
}
#endif

// ------------------------------------------------------------------------------------------
// SDF prefiltering (use_prefiltering = true): forward (d_image == null) or backward.  No boundary
// pass (diffvg.cpp:1558).  The weight image uses JITTERED positions (Q1), the colour samples use
// sub-pixel centres.
EXPORT int emul_render_pf(const int32_t *topo, const float *params, const float *background, float *image,
                          int W, int H, int nsx, int nsy, uint64_t seed, const float *d_image,
                          float *d_params, float *d_background, float *d_translation, int nthreads) {
    HostScene hs;
    build(hs, topo, params);
    if (hs.error_flag) return 3;
    const SceneView &sc = hs.sc;
    const int spp = nsx * nsy;
    std::vector<float> weight((size_t)W * H, 0.f);
    const int ri = (int)ceilf(sc.filter.radius);
    for (int idx = 0; idx < W * H * spp; idx++) {
        const int sx = idx % nsx, sy = (idx / nsx) % nsy, x = (idx / spp) % W, y = idx / (spp * W);
        F2 pt, cpt;
        sample_position(sc.canvas_w, sc.canvas_h, W, H, nsx, nsy, seed, false, x, y, sx, sy, idx, pt, cpt);
        for (int dy = -ri; dy <= ri; dy++)
            for (int dx = -ri; dx <= ri; dx++) {
                int xx = x + dx, yy = y + dy;
                if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                    weight[yy * W + xx] += filter_weight(sc.filter, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
            }
    }
    const float cw = (float)sc.canvas_w, ch = (float)sc.canvas_h;
    const float margin = 4e-4f * std::max(cw, ch) + 1e-4f;
    std::mutex mu;
    std::vector<double> acc(d_params ? hs.params.size() : 0, 0.0);
    const DoubleSink dsink{acc.data()};
    std::vector<float> img_acc(d_image ? 0 : (size_t)W * H * 4, 0.f);
    parallel_rows(H, nthreads, [&](int y) {
        std::vector<int> cand;
        std::vector<PfFragment> frags(DVG_MAXPF);
        for (int x = 0; x < W; x++) {
            candidates(hs, (float)x / W * cw - margin, (float)y / H * ch - margin, (float)(x + 1) / W * cw + margin,
                       (float)(y + 1) / H * ch + margin, cand, true);
            for (int s = 0; s < spp; s++) {
                const int sx = s % nsx, sy = s / nsx;
                const int idx = ((y * W + x) * nsy + sy) * nsx + sx;
                F2 pt, cpt;
                sample_position(sc.canvas_w, sc.canvas_h, W, H, nsx, nsy, seed, true, x, y, sx, sy, idx, pt, cpt);
                const float *bg_px = background ? background + 4 * (y * W + x) : nullptr;
                F4 first = bg_px ? mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]) : mk4(0, 0, 0, 0);
                PrefilterTracer<true> tr;
                tr.init(cpt, true, first, frags.data());
                for (int e : cand) tr.step(sc, prim_ref(hs, e));
                tr.finish(sc);
                const F4 color = tr.resolve(bg_px);
                std::lock_guard<std::mutex> lk(mu);
                if (!d_image) {
                    for (int dy = -ri; dy <= ri; dy++)
                        for (int dx = -ri; dx <= ri; dx++) {
                            int xx = x + dx, yy = y + dy;
                            if (xx >= 0 && xx < W && yy >= 0 && yy < H && weight[yy * W + xx] > 0) {
                                float fw = filter_weight(sc.filter, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
                                float inv_ws = 1.f / weight[yy * W + xx];
                                float *d = &image[((size_t)yy * W + xx) * 4];
                                d[0] += (fw * color.x) * inv_ws; d[1] += (fw * color.y) * inv_ws;
                                d[2] += (fw * color.z) * inv_ws; d[3] += (fw * color.w) * inv_ws;
                            }
                        }
                } else {
                    const F4 d_color = gather_d_color(sc.filter, d_image, weight.data(), W, H, pt);
                    float *dtr = d_translation ? d_translation + 2 * (y * W + x) : nullptr;
                    if (tr.nfrag > 0) {
                        F4 d_bg;
                        prefilter_backward(sc, tr, color, d_color, dsink, dtr, d_bg);
                        if (bg_px && d_background) {
                            float *d = d_background + 4 * (y * W + x);
                            d[0] += d_bg.x; d[1] += d_bg.y; d[2] += d_bg.z; d[3] += d_bg.w;
                        }
                    } else if (bg_px && d_background) {
                        float *d = d_background + 4 * (y * W + x);
                        d[0] += d_color.x; d[1] += d_color.y; d[2] += d_color.z; d[3] += d_color.w;
                    }
                    for (int dy = -ri; dy <= ri; dy++)
                        for (int dx = -ri; dx <= ri; dx++) {
                            int xx = x + dx, yy = y + dy;
                            if (xx >= 0 && xx < W && yy >= 0 && yy < H && weight[yy * W + xx] > 0) {
                                const float ws = weight[yy * W + xx];
                                const float ddx = (xx + 0.5f) - pt.x, ddy = (yy + 0.5f) - pt.y;
                                const float fw = filter_weight(sc.filter, ddx, ddy);
                                const float *dp = d_image + 4 * (yy * W + xx);
                                const float dotv = dp[0] * color.x + dp[1] * color.y + dp[2] * color.z + dp[3] * color.w;
                                const float d_weight = (dotv * ws - fw * dotv * (ws - fw)) / (ws * ws);
                                acc[sc.filter_radius_off] += (double)d_filter_weight_radius(sc.filter, ddx, ddy, d_weight);
                            }
                        }
                }
            }
        }
    });
    if (d_params) for (size_t i = 0; i < acc.size(); i++) d_params[i] = (float)acc[i];
    return 0;
}

// SDF output (sample_distance, diffvg.cpp:709-775).  eval_positions == null: one sample per pixel
// sample, accumulated into sdf[H*W]; else sdf[n_eval].  Backward when d_sdf != null.
EXPORT int emul_sdf(const int32_t *topo, const float *params, float *sdf, int W, int H, int nsx, int nsy, uint64_t seed,
                    const float *eval_positions, int n_eval, const float *d_sdf, float *d_params, float *d_translation) {
    HostScene hs;
    build(hs, topo, params);
    if (hs.error_flag) return 3;
    const SceneView &sc = hs.sc;
    const int spp = nsx * nsy;
    const int n = eval_positions ? n_eval : W * H * spp;
    std::vector<double> acc(d_params ? hs.params.size() : 0, 0.0);
    const DoubleSink dsink{acc.data()};
    for (int idx = 0; idx < n; idx++) {
        F2 pt, cpt;
        int x, y;
        if (!eval_positions) {
            const int sx = idx % nsx, sy = (idx / nsx) % nsy;
            x = (idx / spp) % W; y = idx / (spp * W);
            sample_position(sc.canvas_w, sc.canvas_h, W, H, nsx, nsy, seed, false, x, y, sx, sy, idx, pt, cpt);
        } else {
            pt = mk2(eval_positions[2 * idx], eval_positions[2 * idx + 1]);
            x = (int)pt.x; y = (int)pt.y;
            F2 npt = pt; npt.x /= W; npt.y /= H;
            cpt = mk2(npt.x * sc.canvas_w, npt.y * sc.canvas_h);
        }
        const float weight = eval_positions ? 1.f : 1.f / spp;
        int min_g = -1; DistHit best; dist_hit_init(best, 0.f);
        for (int g = sc.num_groups - 1; g >= 0; g--) {
            DistHit h;
            group_distance(sc, g, cpt, h);
            if (h.found && (min_g == -1 || h.dist < best.dist)) { best = h; min_g = g; }
        }
        float dist = 0.f;
        if (min_g >= 0) {
            dist = best.dist * weight;
            bool inside = false;
            if (sc.groups[min_g].fill_type >= 0) {
                inside = group_is_inside(sc, min_g, cpt);
                if (inside) dist = -dist;
            }
            if (d_sdf) {
                const float dd = eval_positions ? d_sdf[idx] : d_sdf[y * W + x];
                const float d_abs = inside ? -dd : dd;
                d_compute_distance(sc, sc.groups[min_g], best.inst, cpt, best.cp, best.base_id, best.point_id, best.t_root, d_abs, dsink,
                                   d_translation ? d_translation + 2 * (y * W + x) : nullptr);
            }
        }
        if (sdf) { if (eval_positions) sdf[idx] += dist; else sdf[y * W + x] += dist; }
    }
    if (d_params) for (size_t i = 0; i < acc.size(); i++) d_params[i] = (float)acc[i];
    return 0;
}
