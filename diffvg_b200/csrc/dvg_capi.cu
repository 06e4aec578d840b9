// dvg_capi.cu -- host side of the C ABI declared in include/diffvg_b200.h.
#include "dvg_internal.h"
#include "../../include/diffvg_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace dvg {
long long g_launch_count = 0;
int prof_report(char *buf, long long cap);
double peak_probe_flops(int iters, int blocks);
}

using namespace dvg;

namespace {

thread_local std::string g_err;
float *g_debug_out = nullptr;  // dvg_debug_set_boundary_dump
long long g_debug_edge_pass_samples = 0;   // dvg_debug_set_limits
long long g_debug_pair_capacity = 0;
bool g_pf_nocache = getenv("DVG_PF_NOCACHE") != nullptr && getenv("DVG_PF_NOCACHE")[0] == '1';   // dvg_debug_set_prefilter_inline bit 1
bool g_pf_inline = getenv("DVG_PF_INLINE") != nullptr && getenv("DVG_PF_INLINE")[0] == '1';   // dvg_debug_set_prefilter_inline (env: measurements only)
bool g_fast_accept = getenv("DVG_FAST_ACCEPT") != nullptr && getenv("DVG_FAST_ACCEPT")[0] == '1';   // dvg_set_fast_stroke_accept (env: measurements only)

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(DVG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
    } while (0)

// Grow-only device buffer: no cudaMalloc on the steady-state path (the reference allocates
// and frees managed memory inside every render(), diffvg.cpp:1509, 1567-1572).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return (T *)p; }
};

struct DeviceGuard {
    int old = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&old); if (old != dev) cudaSetDevice(dev); else old = -1; }
    ~DeviceGuard() { if (old >= 0) cudaSetDevice(old); }
};

}  // namespace

struct DvgScene {
    int device = 0;
    std::vector<int32_t> topo;
    int canvas_w = 0, canvas_h = 0, num_shapes = 0, num_groups = 0, num_insts = 0, num_prims = 0;
    int num_params = 0, total_segs = 0;   // per scene
    // BATCH (dvg_scene_create_batch): `batch` scenes of this topology live back to back in every table below (SceneView);
    // the counts above stay per scene.  Seeds: one per scene, uploaded when they change; `seeds_version` stands in for
    // the seed in the reuse keys of the weight image and of the forward pass's result words.
    int batch = 1;
    DevBuf d_seeds;
    DevBuf d_scan_ws;   // launch_scan workspace (zeroed once; every scan leaves it zeroed)
    std::vector<uint64_t> seeds_host;
    uint64_t seeds_version = 0;
    // host topology maps
    std::vector<int> inst_group, inst_shape, inst_prim_begin, prim_inst, prim_seg, prim_point_id;
    // device: topology
    DevBuf d_topo, d_inst_group, d_inst_shape, d_inst_prim_begin, d_prim_inst, d_prim_seg, d_prim_point_id;
    // device: parameters + derived tables
    DevBuf d_params, d_shapes_length, d_shape_box, d_shape_r0, d_seg_cdf, d_seg_pmf, d_seg_point_id;
    DevBuf d_insts, d_groups, d_p01, d_p23, d_rad, d_box, d_thick, d_meta, d_cbox, d_cbox_pf, d_cap, d_quint, d_wcert, d_shape_cdf, d_shape_pmf, d_shape_guide;
    DevBuf d_flags;  // [0] error flag, [1] total length (float bits)
    // bins
    DevBuf d_bin_counts, d_bin_offsets, d_bin_items, d_sbin_counts, d_sbin_items;
    int bin_w = 0, bin_h = 0, bin_tw = 0, bin_th = 0, bin_pf = 0;  // configuration the bins were built for (0 = none)
    int bin_r0 = 0, bin_r1 = 0;                                    // tile rows that were binned
    // per-render workspaces
    DevBuf d_weight;
    int w_w = 0, w_h = 0, w_nsx = 0, w_nsy = 0, w_ftype = -1, w_r0 = 0, w_r1 = 0;
    uint64_t w_seed = 0; float w_radius = 0; bool w_valid = false;
    DevBuf d_keys, d_tile_counts, d_tile_offsets, d_tile_fill, d_blk_counts, d_blk_offsets, d_sorted;
    // wavefront passes (dvg_wave.cu)
    DevBuf d_wave_hit, d_wave_wind, d_wave_pairs_s, d_wave_pairs_f, d_wave_counters, d_tile_nch, d_tile_choff,
        d_edge_chunks, d_edge_choff, d_wave_max, d_bsamples, d_bsamples_raw, d_item_tile, d_grad_rep;
    DevBuf d_bvh_path, d_bvh_group, d_bvh_scene, d_bvh_keys;   // reference-topology trees (dvg_bvh.cu), built on demand by dvg_scene_dump
    int total_chunks = 0, max_nch = 0;   // of the current bins (read back with the bin total)
    bool has_fills = false;
    // which pixel pass the result words currently hold (forward's are reused by the interior backward pass)
    bool wpx_valid = false;
    int wpx_w = 0, wpx_h = 0, wpx_nsx = 0, wpx_nsy = 0, wpx_r0 = 0, wpx_r1 = 0, wpx_pf = 0; uint64_t wpx_seed = 0; uint32_t wpx_fast = 0;
    // fragment cache of the prefiltered path (dvg_distance.cuh PfCache): what the forward pass left, and for which render
    DevBuf d_pf_recs, d_pf_count;
    bool pfc_valid = false;
    int pfc_w = 0, pfc_h = 0, pfc_nsx = 0, pfc_nsy = 0, pfc_r0 = 0, pfc_r1 = 0;
    int32_t *h_pinned = nullptr;  // [0] error flag, [1] total bin items
    float *h_params_pinned = nullptr;  // staging for host-resident params (true async H2D)
    cudaEvent_t h_params_free = nullptr;  // recorded after the H2D copy that last read the staging buffer
    int32_t *h_counts = nullptr;         // pinned [2][4]: the wave counters of the last pixel / boundary pass (pair-queue feedback)
    cudaEvent_t ev_counts[2] = {nullptr, nullptr};
    bool counts_pending[2] = {false, false};
    cudaStream_t counts_stream[2] = {nullptr, nullptr};   // where each pending copy was queued
    int64_t want_s = 0, want_f = 0;      // most pairs any pass of this scene asked for so far
    bool slot_seen[2] = {false, false};  // the pixel / boundary pass has reported its counts at least once
    bool params_set = false;
    bool checked = false;
    int scene_error = 0;
    std::string scene_error_msg;
    float filter_radius_host = 0.5f;  // refreshed with the error-flag read-back

    BuildView build_view() {
        BuildView bv;
        bv.canvas_w = canvas_w; bv.canvas_h = canvas_h;
        bv.num_shapes = num_shapes; bv.num_groups = num_groups; bv.num_insts = num_insts; bv.num_prims = num_prims;
        bv.batch = batch; bv.num_params = num_params; bv.total_segs = total_segs;
        bv.topo = d_topo.as<int>(); bv.params = d_params.as<float>();
        bv.inst_group = d_inst_group.as<int>(); bv.inst_shape = d_inst_shape.as<int>();
        bv.inst_prim_begin = d_inst_prim_begin.as<int>();
        bv.prim_inst = d_prim_inst.as<int>(); bv.prim_seg = d_prim_seg.as<int>(); bv.prim_point_id = d_prim_point_id.as<int>();
        bv.shapes_length = d_shapes_length.as<float>(); bv.shape_box = d_shape_box.as<Box>(); bv.shape_r0 = d_shape_r0.as<float>();
        bv.seg_cdf = d_seg_cdf.as<float>(); bv.seg_pmf = d_seg_pmf.as<float>(); bv.seg_point_id = d_seg_point_id.as<int>();
        bv.insts = d_insts.as<InstInfo>(); bv.groups = d_groups.as<GroupInfo>();
        bv.prim_p01 = d_p01.as<F4>(); bv.prim_p23 = d_p23.as<F4>(); bv.prim_rad = d_rad.as<F4>();
        bv.prim_box = d_box.as<Box>(); bv.prim_thick = d_thick.as<float>(); bv.prim_meta = d_meta.as<PrimMeta>();
        bv.prim_cbox = d_cbox.as<Box>(); bv.prim_cbox_pf = d_cbox_pf.as<Box>(); bv.prim_cap = d_cap.as<F4>(); bv.prim_quint = d_quint.as<PrimQuintic>(); bv.prim_wcert = d_wcert.as<PrimWindCert>(); bv.shape_guide = d_shape_guide.as<int>();
        bv.shape_cdf = d_shape_cdf.as<float>(); bv.shape_pmf = d_shape_pmf.as<float>();
        bv.error_flag = d_flags.as<int>(); bv.total_length = d_flags.as<float>() + 1;
        return bv;
    }
    SceneView view() {
        SceneView sc;
        sc.canvas_w = canvas_w; sc.canvas_h = canvas_h;
        sc.num_shapes = num_shapes; sc.num_groups = num_groups; sc.num_insts = num_insts; sc.num_prims = num_prims;
        sc.batch = batch; sc.num_params = num_params; sc.total_segs = total_segs;
        sc.filter.type = topo[DVG_H_FILTER_TYPE];
        sc.filter.radius = filter_radius_host;
        sc.filter_radius_off = topo[DVG_H_FILTER_RADIUS_OFF];
        sc.topo = d_topo.as<int>(); sc.params = d_params.as<float>();
        sc.prim_p01 = d_p01.as<F4>(); sc.prim_p23 = d_p23.as<F4>(); sc.prim_rad = d_rad.as<F4>();
        sc.prim_box = d_box.as<Box>(); sc.prim_thick = d_thick.as<float>(); sc.prim_meta = d_meta.as<PrimMeta>();
        sc.prim_cbox = d_cbox.as<Box>(); sc.prim_cbox_pf = d_cbox_pf.as<Box>(); sc.prim_cap = d_cap.as<F4>(); sc.prim_quint = d_quint.as<PrimQuintic>(); sc.prim_wcert = d_wcert.as<PrimWindCert>(); sc.shape_guide = d_shape_guide.as<int>();
        sc.insts = d_insts.as<InstInfo>(); sc.groups = d_groups.as<GroupInfo>();
        sc.shapes_length = d_shapes_length.as<float>();
        sc.shape_cdf = d_shape_cdf.as<float>(); sc.shape_pmf = d_shape_pmf.as<float>();
        sc.seg_cdf = d_seg_cdf.as<float>(); sc.seg_pmf = d_seg_pmf.as<float>(); sc.seg_point_id = d_seg_point_id.as<int>();
        sc.error_flag = d_flags.as<int>();
        return sc;
    }
    BinView bin_view() {
        BinView b;
        b.tile_w = bin_tw; b.tile_h = bin_th; b.batch = batch;
        b.tiles_x = (bin_w + bin_tw - 1) / bin_tw; b.tiles_y = (bin_h + bin_th - 1) / bin_th;
        b.offsets = d_bin_offsets.as<int>(); b.items = d_bin_items.as<int>();
        return b;
    }
    void release_all() {
        DevBuf *all[] = {&d_topo, &d_inst_group, &d_inst_shape, &d_inst_prim_begin, &d_prim_inst, &d_prim_seg,
                         &d_prim_point_id, &d_params, &d_shapes_length, &d_shape_box, &d_shape_r0, &d_seg_cdf, &d_seg_pmf,
                         &d_seg_point_id, &d_insts, &d_groups, &d_p01, &d_p23, &d_rad, &d_box, &d_thick, &d_meta, &d_cbox, &d_cbox_pf, &d_cap, &d_quint, &d_wcert, &d_shape_guide,
                         &d_shape_cdf, &d_shape_pmf, &d_flags, &d_bin_counts, &d_bin_offsets, &d_bin_items, &d_sbin_counts, &d_sbin_items, &d_weight,
                         &d_keys, &d_tile_counts, &d_tile_offsets, &d_tile_fill, &d_blk_counts, &d_blk_offsets, &d_sorted,
                         &d_wave_hit, &d_wave_wind, &d_wave_pairs_s, &d_wave_pairs_f, &d_wave_counters, &d_tile_nch, &d_tile_choff,
                         &d_edge_chunks, &d_edge_choff, &d_wave_max, &d_bsamples, &d_bsamples_raw, &d_item_tile, &d_grad_rep,
                         &d_bvh_path, &d_bvh_group, &d_bvh_scene, &d_bvh_keys, &d_seeds, &d_scan_ws};
        for (DevBuf *b : all) b->release();
        if (h_pinned) cudaFreeHost(h_pinned);
        h_pinned = nullptr;
        if (h_params_pinned) cudaFreeHost(h_params_pinned);
        h_params_pinned = nullptr;
        if (h_params_free) cudaEventDestroy(h_params_free);
        h_params_free = nullptr;
        if (h_counts) cudaFreeHost(h_counts);
        h_counts = nullptr;
        for (int k = 0; k < 2; k++) { if (ev_counts[k]) cudaEventDestroy(ev_counts[k]); ev_counts[k] = nullptr; }
    }
};

namespace {

template <typename T>
int upload(DevBuf &buf, const std::vector<T> &v) {
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CK(buf.ensure(bytes));
    if (!v.empty()) CK(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return DVG_OK;
}

int validate_topo(const int32_t *t, int64_t len) {
    if (len < DVG_TOPO_HEADER_LEN) return fail(DVG_ERR_INVALID, "topo shorter than its header");
    if (t[DVG_H_MAGIC] != DVG_TOPO_MAGIC) return fail(DVG_ERR_INVALID, "bad topo magic");
    const int ns = t[DVG_H_NUM_SHAPES], ng = t[DVG_H_NUM_GROUPS], np = t[DVG_H_NUM_PARAMS];
    if (ns <= 0 || ng <= 0) return fail(DVG_ERR_INVALID, "scene needs at least one shape and one shape group");
    if (t[DVG_H_CANVAS_W] <= 0 || t[DVG_H_CANVAS_H] <= 0) return fail(DVG_ERR_INVALID, "bad canvas size");
    auto in = [&](int64_t off, int64_t n) { return off >= DVG_TOPO_HEADER_LEN && n >= 0 && off + n <= len; };
    if (!in(t[DVG_H_OFF_SHAPES], (int64_t)ns * DVG_SHAPE_REC_LEN) || !in(t[DVG_H_OFF_GROUPS], (int64_t)ng * DVG_GROUP_REC_LEN) ||
        !in(t[DVG_H_OFF_NCP], t[DVG_H_TOTAL_SEGS]) || !in(t[DVG_H_OFF_GSHAPES], t[DVG_H_TOTAL_GSHAPES]))
        return fail(DVG_ERR_INVALID, "topo section out of bounds");
    auto pin = [&](int off, int n) { return off >= 0 && n >= 0 && (int64_t)off + n <= np; };
    if (t[DVG_H_FILTER_TYPE] < 0 || t[DVG_H_FILTER_TYPE] > 3) return fail(DVG_ERR_INVALID, "bad filter type");
    if (!pin(t[DVG_H_FILTER_RADIUS_OFF], 1)) return fail(DVG_ERR_INVALID, "filter radius offset out of range");
    for (int i = 0; i < ns; i++) {
        const int32_t *r = t + t[DVG_H_OFF_SHAPES] + i * DVG_SHAPE_REC_LEN;
        int nfl;
        switch (r[DVG_S_TYPE]) {
            case DVG_SHAPE_CIRCLE: nfl = 3; break;
            case DVG_SHAPE_ELLIPSE: nfl = 4; break;
            case DVG_SHAPE_RECT: nfl = 4; break;
            case DVG_SHAPE_PATH: {
                const int npts = r[DVG_S_NUM_POINTS], nseg = r[DVG_S_NUM_SEGS];
                if (npts <= 0 || nseg <= 0) return fail(DVG_ERR_INVALID, "path with no points or segments");
                if (r[DVG_S_NCP_OFF] < 0 || r[DVG_S_NCP_OFF] + nseg > t[DVG_H_TOTAL_SEGS])
                    return fail(DVG_ERR_INVALID, "path ncp range out of bounds");
                const int32_t *ncp = t + t[DVG_H_OFF_NCP] + r[DVG_S_NCP_OFF];
                int need = 0;
                for (int k = 0; k < nseg; k++) {
                    if (ncp[k] < 0 || ncp[k] > 2) return fail(DVG_ERR_INVALID, "num_control_points must be 0, 1 or 2");
                    need += ncp[k] + 1;
                }
                // closed paths wrap the last end point onto point 0; open ones need it explicitly
                // segment point indices wrap modulo num_points (closed paths), as in the reference
                if (need > npts) return fail(DVG_ERR_INVALID, "path has fewer points than its segments need");
                if (r[DVG_S_THICK_OFF] >= 0 && !pin(r[DVG_S_THICK_OFF], npts))
                    return fail(DVG_ERR_INVALID, "path thickness out of params range");
                nfl = 2 * npts;
                break;
            }
            default: return fail(DVG_ERR_INVALID, "bad shape type");
        }
        if (!pin(r[DVG_S_PARAM_OFF], nfl)) return fail(DVG_ERR_INVALID, "shape params out of range");
        if (r[DVG_S_WIDTH_OFF] >= 0 && !pin(r[DVG_S_WIDTH_OFF], 1)) return fail(DVG_ERR_INVALID, "stroke width offset out of range");
    }
    for (int g = 0; g < ng; g++) {
        const int32_t *r = t + t[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN;
        if (r[DVG_G_NUM_SHAPES] <= 0 || r[DVG_G_SHAPES_OFF] < 0 || r[DVG_G_SHAPES_OFF] + r[DVG_G_NUM_SHAPES] > t[DVG_H_TOTAL_GSHAPES])
            return fail(DVG_ERR_INVALID, "group shape list out of bounds");
        const int32_t *ids = t + t[DVG_H_OFF_GSHAPES] + r[DVG_G_SHAPES_OFF];
        for (int k = 0; k < r[DVG_G_NUM_SHAPES]; k++) {
            if (ids[k] < 0 || ids[k] >= ns) return fail(DVG_ERR_INVALID, "group references a shape id out of range");
            const int32_t *sr = t + t[DVG_H_OFF_SHAPES] + ids[k] * DVG_SHAPE_REC_LEN;
            if (sr[DVG_S_TYPE] == DVG_SHAPE_ELLIPSE && r[DVG_G_STROKE_TYPE] >= 0)
                return fail(DVG_ERR_UNSUPPORTED, "stroked ellipses are not supported (the reference asserts: within_distance.h:342-345)");
        }
        for (int which = 0; which < 2; which++) {
            const int type = r[which ? DVG_G_STROKE_TYPE : DVG_G_FILL_TYPE];
            const int off = r[which ? DVG_G_STROKE_OFF : DVG_G_FILL_OFF];
            const int stops = r[which ? DVG_G_STROKE_STOPS : DVG_G_FILL_STOPS];
            if (type < -1 || type > 2) return fail(DVG_ERR_INVALID, "bad colour type");
            if (type == 0 && !pin(off, 4)) return fail(DVG_ERR_INVALID, "colour out of params range");
            if (type > 0 && (stops <= 0 || !pin(off, 4 + 5 * stops))) return fail(DVG_ERR_INVALID, "gradient out of params range");
        }
        if (!pin(r[DVG_G_XFORM_OFF], 9)) return fail(DVG_ERR_INVALID, "shape_to_canvas out of params range");
    }
    return DVG_OK;
}

void choose_tile(int spp, int *tw, int *th) {
    if (spp >= 16) { *tw = 8; *th = 2; }
    else if (spp >= 4) { *tw = 8; *th = 8; }
    else if (spp >= 2) { *tw = 16; *th = 8; }
    else { *tw = 16; *th = 16; }
}

// After the build the host needs (a) whether the scene is degenerate (scene.cpp:231-240 throws) and (b) the filter
// radius.  Both are read back together with the bin totals by ensure_bins -- ONE synchronisation per set_params, early
// in the forward pass -- or here when no bins are needed (SDF output) or they are sized for the worst case.
int parse_build_flags(DvgScene *s) {
    s->checked = true;
    s->scene_error = s->h_pinned[0];
    float total; memcpy(&total, &s->h_pinned[1], 4);
    memcpy(&s->filter_radius_host, &s->h_pinned[2], 4);
    if (s->scene_error) {
        char buf[256];
        if (s->scene_error == 1)
            snprintf(buf, sizeof buf, "The total length of the shape boundaries in the scene is equal or less than 0. Length = %f", total);
        else
            snprintf(buf, sizeof buf, "The total length of the shape boundaries in the scene is not a number. Length = %f", total);
        s->scene_error_msg = buf;
        return fail(DVG_ERR_SCENE, buf);
    }
    return DVG_OK;
}

int queue_build_flags(DvgScene *s, cudaStream_t st) {
    CK(cudaMemcpyAsync(s->h_pinned, s->d_flags.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(s->h_pinned + 2, s->d_params.as<float>() + s->topo[DVG_H_FILTER_RADIUS_OFF], 4,
                       cudaMemcpyDeviceToHost, st));
    return DVG_OK;
}

int finish_build(DvgScene *s, cudaStream_t st) {
    if (s->checked) return s->scene_error ? fail(DVG_ERR_SCENE, s->scene_error_msg) : DVG_OK;
    int rc = queue_build_flags(s, st);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    return parse_build_flags(s);
}

void wave_feedback_poll(DvgScene *s, cudaStream_t synced = nullptr, bool have_synced = false);   // (pair-queue sizing, below)

constexpr int64_t kSmallBins = 1 << 20;    // tiles x primitives below which bins are sized for the worst case
constexpr int64_t kSmallPairs = 1 << 21;   // worst-case exact tests of a pass below which the pair queues are, too

// `row_begin, row_end`: pixel rows the caller will render.  The prefiltered path has no boundary pass, so a row shard
// only ever looks at the tiles of its own rows and only those are binned (at 8 GPUs binning the whole 2048^2 image
// on every rank was 10% of the step); the boundary pass of the sampled path lands anywhere, so it bins everything.
// Also completes the scene build (finish_build) with the same synchronisation.
int ensure_bins(DvgScene *s, int width, int height, int spp, int pf, cudaStream_t st, int row_begin, int row_end) {
    int tw, th;
    choose_tile(spp, &tw, &th);
    const int tiles_y_all = (height + th - 1) / th;
    const int r0 = pf ? row_begin / th : 0, r1 = pf ? std::min(tiles_y_all, (row_end + th - 1) / th) : tiles_y_all;
    if (s->bin_w == width && s->bin_h == height && s->bin_tw == tw && s->bin_th == th && s->bin_pf == pf &&
        s->bin_r0 <= r0 && s->bin_r1 >= r1) return finish_build(s, st);
    BinBuild bb;
    bb.width = width; bb.height = height; bb.tile_w = tw; bb.tile_h = th; bb.prefilter = pf;
    bb.batch = s->batch;
    bb.scan_ws = s->d_scan_ws.as<int>();
    bb.flat = s->num_prims <= 4 * s->num_groups ? 1 : 0;
    bb.tile_row0 = r0; bb.tile_row1 = r1;
    bb.tiles_x = (width + tw - 1) / tw; bb.tiles_y = (height + th - 1) / th;
    const int ntiles = bb.tiles_x * bb.tiles_y * s->batch;   // a batch: every scene's tiles, scene after scene
    CK(s->d_bin_counts.ensure(sizeof(int) * ntiles));
    CK(s->d_bin_offsets.ensure(sizeof(int) * (ntiles + 1)));
    bb.counts = s->d_bin_counts.as<int>(); bb.offsets = s->d_bin_offsets.as<int>(); bb.items = nullptr;
    // two-level binning when the supertile lists fit a fixed stride of num_prims entries each (<= 128 MB)
    bb.super = 8;
    bb.stiles_x = (bb.tiles_x + bb.super - 1) / bb.super; bb.stiles_y = (bb.tiles_y + bb.super - 1) / bb.super;
    const int64_t sentries = (int64_t)bb.stiles_x * bb.stiles_y * s->num_prims;
    if (ntiles < 256 || s->num_prims < 64 || sentries > ((int64_t)1 << 25) || s->batch > 1) bb.super = 0;
    bb.s_counts = nullptr; bb.s_items = nullptr;
    if (bb.super) {
        CK(s->d_sbin_counts.ensure(sizeof(int) * (size_t)bb.stiles_x * bb.stiles_y));
        CK(s->d_sbin_items.ensure(sizeof(int) * (size_t)sentries));
        bb.s_counts = s->d_sbin_counts.as<int>(); bb.s_items = s->d_sbin_items.as<int>();
    }
    BuildView bv = s->build_view();
    launch_bin_coarse(bv, bb, st);
    launch_bin_count(bv, bb, st);
    CK(s->d_tile_nch.ensure(sizeof(int) * ntiles));
    CK(s->d_tile_choff.ensure(sizeof(int) * (ntiles + 1)));
    CK(s->d_wave_max.ensure(sizeof(int)));
    launch_wave_tile_chunks(bb.offsets, s->d_tile_nch.as<int>(), s->d_tile_choff.as<int>(), s->d_wave_max.as<int>(), ntiles, bb.scan_ws, st);
    int total;
    const int64_t nbin = s->batch > 1 ? ntiles : (int64_t)(r1 - r0) * bb.tiles_x;
    if (nbin * s->num_prims <= kSmallBins) {
        // small scene (batched 64x64 scenes, single shapes): size everything for the worst case -- every primitive in
        // every tile -- and skip the read-back of the totals
        total = (int)(nbin * s->num_prims);
        s->max_nch = (s->num_prims + 31) / 32;
        s->total_chunks = (int)nbin * s->max_nch;
        int rc = finish_build(s, st);
        if (rc) return rc;
    } else {
        const bool flags_too = !s->checked;
        if (flags_too) { int rc = queue_build_flags(s, st); if (rc) return rc; }
        CK(cudaMemcpyAsync(s->h_pinned + 3, bb.offsets + ntiles, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(s->h_pinned + 4, s->d_tile_choff.as<int>() + ntiles, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(s->h_pinned + 5, s->d_wave_max.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));   // the one synchronisation of an iteration
        wave_feedback_poll(s, st, true);
        total = s->h_pinned[3];
        s->total_chunks = s->h_pinned[4];
        s->max_nch = s->h_pinned[5];
        int rc = flags_too ? parse_build_flags(s) : finish_build(s, st);
        if (rc) return rc;
    }
    s->wpx_valid = false; s->pfc_valid = false;
    CK(s->d_bin_items.ensure(sizeof(int) * std::max(total, 1)));
    bb.items = s->d_bin_items.as<int>();
    launch_bin_fill(bv, bb, st);
    CK(cudaGetLastError());
    s->bin_w = width; s->bin_h = height; s->bin_tw = tw; s->bin_th = th; s->bin_pf = pf; s->bin_r0 = r0; s->bin_r1 = r1;
    return DVG_OK;
}

// rows [r0, r1): the pixel rows whose weights the caller reads.  A forward row shard and the prefiltered backward pass
// read their own rows; the boundary pass of the sampled backward path gathers d_image anywhere: whole image.
int ensure_weight(DvgScene *s, const SceneView &sc, RenderArgs &ra, int r0, int r1, cudaStream_t st) {
    const size_t wpx = (size_t)ra.width * ra.height * s->batch;
    CK(s->d_weight.ensure(sizeof(float) * wpx));
    ra.weight_image = s->d_weight.as<float>();
    if (s->w_valid && s->w_w == ra.width && s->w_h == ra.height && s->w_nsx == ra.nsx && s->w_nsy == ra.nsy &&
        s->w_seed == ra.seed && s->w_ftype == sc.filter.type && s->w_radius == sc.filter.radius &&
        s->w_r0 <= r0 && s->w_r1 >= r1)
        return DVG_OK;  // weights depend only on (size, spp, seed, filter, rows): reuse forward's in backward
    CK(cudaMemsetAsync(ra.weight_image, 0, sizeof(float) * wpx, st));
    launch_weight(sc, ra, r0, r1, st);
    s->w_valid = true; s->w_w = ra.width; s->w_h = ra.height; s->w_nsx = ra.nsx; s->w_nsy = ra.nsy;
    s->w_seed = ra.seed; s->w_ftype = sc.filter.type; s->w_radius = sc.filter.radius; s->w_r0 = r0; s->w_r1 = r1;
    return DVG_OK;
}

// ---- wavefront passes: workspace sizing and the classify -> solve sequence.
// Nothing is read back inside a pass.  The pair queues keep the capacity that earlier passes asked for: every pass copies
// its counters to pinned memory when it ends, and the next pass -- whenever that copy has landed -- grows a queue that
// was too small.  A pass that still overflows (the first one of a scene; a sudden change of the geometry) is followed by
// a retry kernel that answers every pair in place and rewrites the result words (dvg_wave.cu wave_classify<true>):
// slower, same results; when nothing overflowed that kernel exits at once.
// `synced`: a stream the caller has just synchronised (copies queued on it have landed whatever the event query says:
// under a profiler that serialises launches the query was seen to stay "not ready"), or null.
void wave_feedback_poll(DvgScene *s, cudaStream_t synced, bool have_synced) {
    for (int slot = 0; slot < 2; slot++) {
        if (!s->counts_pending[slot]) continue;
        const bool landed = (have_synced && s->counts_stream[slot] == synced) || cudaEventQuery(s->ev_counts[slot]) == cudaSuccess;
        if (!landed) continue;
        s->counts_pending[slot] = false;
        const int32_t *c = s->h_counts + 4 * slot;
        // a counter that wrapped negative asked for more than 2^31 pairs: keep the in-place path for the surplus
        s->slot_seen[slot] = true;
        s->want_s = std::max<int64_t>(s->want_s, c[0] < 0 ? 0x7fffffff : c[0]);
        s->want_f = std::max<int64_t>(s->want_f, c[1] < 0 ? 0x7fffffff : c[1]);
    }
}

int wave_view(DvgScene *s, int64_t chunk_slots, int64_t evals, WaveView *out, int slot = 0) {
    const int64_t words = chunk_slots * 32;
    if (words >= ((int64_t)1 << 27)) return fail(DVG_ERR_UNSUPPORTED, "render too large for the 27-bit result-word index of the pair queue");
    CK(s->d_wave_hit.ensure(sizeof(unsigned) * (size_t)std::max<int64_t>(words, 1)));
    if (s->has_fills) CK(s->d_wave_wind.ensure(sizeof(unsigned) * 4 * (size_t)std::max<int64_t>(words, 1)));
    CK(s->d_wave_counters.ensure(sizeof(int) * 4));
    CK(s->d_edge_choff.ensure(sizeof(int) * 4));   // real size set by the boundary pass
    wave_feedback_poll(s);
    // worst case: every lane of every chunk slot asks for an exact test of all 32 candidates.  Small passes get queues
    // of that size (they can not overflow); otherwise what earlier passes asked for plus a margin, and for the first pass
    // a guess of three stroke tests / four winding tests per evaluation
    const int64_t worst = chunk_slots * 32 * 32;
    const int64_t lim = (int64_t)1 << 30;
    int64_t need_s, need_f;
    if (worst <= kSmallPairs) need_s = need_f = std::max<int64_t>(worst, 1);
    else {
        need_s = s->want_s ? s->want_s + s->want_s / 8 + 4096 : 0;
        need_f = s->want_f ? s->want_f + s->want_f / 8 + 4096 : 0;
        if (!s->slot_seen[slot]) {   // this kind of pass has no history yet (the boundary pass asks for ~3x the pixel pass's pairs)
            need_s = std::max<int64_t>(need_s, std::max<int64_t>(3 * evals + evals / 2, 1 << 16));
            need_f = std::max<int64_t>(need_f, std::max<int64_t>(4 * evals, 1 << 16));
        }
        need_s = std::min(std::min(need_s, worst), lim);
        need_f = std::min(std::min(need_f, worst), lim);
    }
    if (g_debug_pair_capacity > 0 && worst > kSmallPairs) need_s = need_f = g_debug_pair_capacity;
    if ((int64_t)(s->d_wave_pairs_s.cap / sizeof(WavePair)) < need_s) CK(s->d_wave_pairs_s.ensure(sizeof(WavePair) * (size_t)need_s));
    if (s->has_fills && (int64_t)(s->d_wave_pairs_f.cap / sizeof(WavePair)) < need_f) CK(s->d_wave_pairs_f.ensure(sizeof(WavePair) * (size_t)need_f));
    WaveView wv;
    wv.hit = s->d_wave_hit.as<unsigned>();
    wv.wind = s->has_fills ? s->d_wave_wind.as<unsigned>() : nullptr;
    wv.pairs_s = s->d_wave_pairs_s.as<WavePair>(); wv.cap_s = (int)std::min<int64_t>(s->d_wave_pairs_s.cap / sizeof(WavePair), lim);
    wv.pairs_f = s->d_wave_pairs_f.as<WavePair>(); wv.cap_f = s->has_fills ? (int)std::min<int64_t>(s->d_wave_pairs_f.cap / sizeof(WavePair), lim) : 0;
    if (g_debug_pair_capacity > 0 && worst > kSmallPairs) {
        wv.cap_s = (int)std::min<int64_t>(wv.cap_s, g_debug_pair_capacity);
        wv.cap_f = (int)std::min<int64_t>(wv.cap_f, g_debug_pair_capacity);
    }
    wv.counters = s->d_wave_counters.as<int>();
    wv.tile_choff = s->d_tile_choff.as<int>();
    wv.edge_choff = s->d_edge_choff.as<int>();
    *out = wv;
    return DVG_OK;
}

// Runs `classify` (W1) and the exact tests (W2); `slot`: 0 pixel pass, 1 boundary pass (where its counters are parked).
template <typename Classify, typename Retry>
int wave_classify_and_solve(DvgScene *s, const SceneView &sc, WaveView &wv, int slot, cudaStream_t st, const Classify &classify,
                            const Retry &retry, bool small) {
    CK(cudaMemsetAsync(wv.counters, 0, sizeof(int) * 2, st));
    classify(wv);
    CK(cudaGetLastError());
    launch_wave_solve(sc, wv, true, s->has_fills, st);
    CK(cudaGetLastError());
    if (!small) {   // worst-case-sized queues can not overflow
        retry(wv);  // exits at once unless a queue overflowed
        CK(cudaGetLastError());
    }
    if (!s->counts_pending[slot]) {   // an unread copy of an earlier pass is still in flight: skip this one
        CK(cudaMemcpyAsync(s->h_counts + 4 * slot, wv.counters, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(s->ev_counts[slot], st));
        s->counts_pending[slot] = true;
        s->counts_stream[slot] = st;
    }
    return DVG_OK;
}

// Pixel pass (forward, or the interior term of the backward pass): the result words of a forward pass are
// reused by the backward pass of the same (scene, size, samples, seed, rows).
int wave_pixel_pass(DvgScene *s, const SceneView &sc, const BinView &bins, const RenderArgs &ra, bool backward, cudaStream_t st) {
    const int64_t wpt = wave_items_per_tile(bins, ra);
    const int items = wave_pixel_items(bins, ra);
    WaveView wv;
    int rc = wave_view(s, (int64_t)s->total_chunks * wpt, (int64_t)items * 32, &wv);
    if (rc) return rc;
    const uint32_t fast = ra.flags & DVG_RF_FAST_ACCEPT;
    const bool reuse = s->wpx_valid && s->wpx_w == ra.width && s->wpx_h == ra.height && s->wpx_nsx == ra.nsx && s->wpx_nsy == ra.nsy &&
                       s->wpx_seed == ra.seed && s->wpx_r0 == ra.row_begin && s->wpx_r1 == ra.row_end &&
                       s->wpx_pf == ra.use_prefiltering && s->wpx_fast == fast;
    if (!reuse) {
        s->wpx_valid = false;
        const bool small = (int64_t)s->total_chunks * wpt * 32 * 32 <= kSmallPairs;
        rc = wave_classify_and_solve(s, sc, wv, 0, st, [&](const WaveView &v) { launch_wave_classify_px(sc, bins, ra, v, st); },
                                     [&](const WaveView &v) { launch_wave_retry_px(sc, bins, ra, v, st); }, small);
        if (rc) return rc;
        s->wpx_valid = true; s->wpx_w = ra.width; s->wpx_h = ra.height; s->wpx_nsx = ra.nsx; s->wpx_nsy = ra.nsy;
        s->wpx_seed = ra.seed; s->wpx_r0 = ra.row_begin; s->wpx_r1 = ra.row_end; s->wpx_pf = ra.use_prefiltering; s->wpx_fast = fast;
    }
    launch_wave_composite_px(sc, bins, ra, wv, backward, st);
    CK(cudaGetLastError());
    return DVG_OK;
}

// Prefiltered pixel pass (sample_color_prefiltered).  Two things are kept from the forward pass for the backward pass of the
// same (scene, size, samples, rows) -- the prefiltered sample positions do not depend on the seed:
//  * the winding numbers of the filled groups, answered by the wavefront pair (classify with the stroke side masked ->
//    k_wave_solve_fill) and read by k_render_pf as words.  Scenes without fills, and renders whose words would not fit the
//    27-bit index, run the winding test inline;
//  * the first DVG_PFC_K fragment records of every sample (PfCache): the backward pass differentiates those samples from
//    the records (k_pf_backward_cached) and walks the candidate lists only for samples with more fragments.
constexpr int64_t kPfCacheMaxBytes = (int64_t)8 << 30;
int wave_pf_pass(DvgScene *s, const SceneView &sc, const BinView &bins, const RenderArgs &ra, bool backward, cudaStream_t st) {
    const int64_t wpt = wave_items_per_tile(bins, ra);
    const int items = wave_pixel_items(bins, ra);
    const int64_t nwords = (int64_t)s->total_chunks * wpt * 32;
    const bool words = s->has_fills && !g_pf_inline && nwords < ((int64_t)1 << 27) && items > 0;
    WaveView wv;
    memset(&wv, 0, sizeof wv);
    if (words) {
        int rc = wave_view(s, (int64_t)s->total_chunks * wpt, (int64_t)items * 32, &wv);
        if (rc) return rc;
        const bool reuse = s->wpx_valid && s->wpx_w == ra.width && s->wpx_h == ra.height && s->wpx_nsx == ra.nsx && s->wpx_nsy == ra.nsy &&
                           s->wpx_r0 == ra.row_begin && s->wpx_r1 == ra.row_end && s->wpx_pf == ra.use_prefiltering;
        if (!reuse) {
            s->wpx_valid = false;
            const bool small = (int64_t)s->total_chunks * wpt * 32 * 32 <= kSmallPairs;
            CK(cudaMemsetAsync(wv.counters, 0, sizeof(int) * 2, st));
            launch_wave_classify_px(sc, bins, ra, wv, st);
            CK(cudaGetLastError());
            launch_wave_solve(sc, wv, false, true, st);
            CK(cudaGetLastError());
            if (!small) { launch_wave_retry_px(sc, bins, ra, wv, st); CK(cudaGetLastError()); }
            if (!s->counts_pending[0]) {
                CK(cudaMemcpyAsync(s->h_counts, wv.counters, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaEventRecord(s->ev_counts[0], st));
                s->counts_pending[0] = true;
                s->counts_stream[0] = st;
            }
            s->wpx_valid = true; s->wpx_w = ra.width; s->wpx_h = ra.height; s->wpx_nsx = ra.nsx; s->wpx_nsy = ra.nsy;
            s->wpx_seed = ra.seed; s->wpx_r0 = ra.row_begin; s->wpx_r1 = ra.row_end; s->wpx_pf = ra.use_prefiltering;
            s->wpx_fast = ra.flags & DVG_RF_FAST_ACCEPT;
        }
    }
    const int64_t threads = pf_launch_threads(bins, ra);
    const bool cacheable = !g_pf_nocache && threads > 0 && threads * DVG_PFC_K * 32 <= kPfCacheMaxBytes;
    PfCache pc;
    pc.recs = nullptr; pc.count = nullptr;
    if (!backward) {
        s->pfc_valid = false;
        if (cacheable) {
            CK(s->d_pf_recs.ensure((size_t)threads * DVG_PFC_K * 32));
            CK(s->d_pf_count.ensure(sizeof(int) * (size_t)threads));
            pc.recs = s->d_pf_recs.as<U4>(); pc.count = s->d_pf_count.as<int>();
        }
        launch_render_pf(sc, bins, ra, wv.wind, wv.hit, wv.tile_choff, pc, false, st);
        CK(cudaGetLastError());
        if (cacheable) {
            s->pfc_valid = true; s->pfc_w = ra.width; s->pfc_h = ra.height; s->pfc_nsx = ra.nsx; s->pfc_nsy = ra.nsy;
            s->pfc_r0 = ra.row_begin; s->pfc_r1 = ra.row_end;
        }
        return DVG_OK;
    }
    const bool cached = cacheable && s->pfc_valid && s->pfc_w == ra.width && s->pfc_h == ra.height && s->pfc_nsx == ra.nsx &&
                        s->pfc_nsy == ra.nsy && s->pfc_r0 == ra.row_begin && s->pfc_r1 == ra.row_end;
    if (cached) {
        pc.recs = s->d_pf_recs.as<U4>(); pc.count = s->d_pf_count.as<int>();
        launch_pf_backward_cached(sc, bins, ra, pc, st);
        CK(cudaGetLastError());
    }
    launch_render_pf(sc, bins, ra, wv.wind, wv.hit, wv.tile_choff, pc, true, st);   // (with a cache: the samples it could not hold)
    CK(cudaGetLastError());
    return DVG_OK;
}

// Boundary pass (diffvg.cpp:1558-1626) over the boundary-sample indices [bw.sample_begin, + bw.num_samples).  `bw` comes
// with its sort buffers bound.
int wave_edge_pass(DvgScene *s, const SceneView &sc, const BinView &bins, const RenderArgs &ra, BoundaryWork &bw, cudaStream_t st) {
    const int ntiles = bin_total_tiles(bins);
    const int spi = wave_edge_samples_per_item();
    bw.max_blocks = bw.num_samples / spi + ntiles;        // upper bound on the boundary items
    CK(s->d_edge_chunks.ensure(sizeof(int) * ntiles));
    CK(s->d_edge_choff.ensure(sizeof(int) * (ntiles + 1)));
    CK(s->d_bsamples.ensure(sizeof(BoundarySample) * (size_t)bw.num_samples));
    CK(s->d_bsamples_raw.ensure(sizeof(BoundarySample) * (size_t)bw.num_samples));
    CK(s->d_item_tile.ensure(sizeof(int) * (size_t)bw.max_blocks));
    bw.samples = s->d_bsamples.as<BoundarySample>();
    bw.samples_unsorted = s->d_bsamples_raw.as<BoundarySample>();
    bw.item_tile = s->d_item_tile.as<int>();
    WaveView wv;
    // every item of a tile has that tile's chunk count: bounded by max_nch without another read-back
    int rc = wave_view(s, (int64_t)bw.max_blocks * std::max(s->max_nch, 1), (int64_t)bw.max_blocks * 32, &wv, 1);
    if (rc) return rc;
    s->wpx_valid = false;   // the result words are about to be overwritten
    launch_wave_boundary_sort(sc, bins, ra, bw, wv, s->d_edge_chunks.as<int>(), st);
    CK(cudaGetLastError());
    const bool small = (int64_t)bw.max_blocks * std::max(s->max_nch, 1) * 32 * 32 <= kSmallPairs;
    rc = wave_classify_and_solve(s, sc, wv, 1, st, [&](const WaveView &v) { launch_wave_classify_edge(sc, bins, ra, bw, v, st); },
                                 [&](const WaveView &v) { launch_wave_retry_edge(sc, bins, ra, bw, v, st); }, small);
    if (rc) return rc;
    launch_wave_composite_edge(sc, bins, ra, bw, wv, st);
    CK(cudaGetLastError());
    return DVG_OK;
}

// How many boundary samples one boundary pass may take: its result words are sized for the bound
// (samples / 16 + tiles) items x the chunk count of the densest tile, and must stay below 2^26 words (256 MB of hit
// words, 1 GB of winding words with fills; the pair records address 2^27).  Larger renders (2048^2 at 4x4 spp) run
// several passes over consecutive sample ranges.
int64_t edge_pass_samples(DvgScene *s, int ntiles) {
    const int64_t word_cap = (int64_t)1 << 26;
    const int64_t items = word_cap / 32 / std::max(s->max_nch, 1) - ntiles;
    return items * wave_edge_samples_per_item();
}

int check_render_args(DvgScene *s, int width, int height, int nsx, int nsy) {
    if (!s) return fail(DVG_ERR_INVALID, "null scene");
    if (!s->params_set) return fail(DVG_ERR_INVALID, "dvg_scene_set_params has not been called");
    if (width <= 0 || height <= 0 || nsx <= 0 || nsy <= 0) return fail(DVG_ERR_INVALID, "bad render size / sample counts");
    if ((int64_t)width * height * nsx * nsy >= (int64_t)1 << 31)
        return fail(DVG_ERR_INVALID, "width*height*samples must fit a 32-bit index (reference: parallel.h:39-44)");
    return DVG_OK;
}

}  // namespace

extern "C" {

int dvg_abi_version(void) { return 1; }

const char *dvg_last_error(void) { return g_err.c_str(); }

int64_t dvg_kernel_launch_count(void) { return dvg::g_launch_count; }

int dvg_scene_create(const int32_t *topo, int64_t topo_len, int device, DvgScene **out_scene) {
    return dvg_scene_create_batch(topo, topo_len, device, 1, out_scene);
}

int dvg_scene_create_batch(const int32_t *topo, int64_t topo_len, int device, int batch, DvgScene **out_scene) {
    if (!topo || !out_scene) return fail(DVG_ERR_INVALID, "null argument");
    if (batch < 1) return fail(DVG_ERR_INVALID, "batch must be at least 1");
    int rc = validate_topo(topo, topo_len);
    if (rc) return rc;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(DVG_ERR_INVALID, "bad CUDA device ordinal");
    DeviceGuard guard(device);
    DvgScene *s = new DvgScene();
    s->device = device;
    s->batch = batch;
    s->topo.assign(topo, topo + topo_len);
    const int32_t *t = s->topo.data();
    s->canvas_w = t[DVG_H_CANVAS_W]; s->canvas_h = t[DVG_H_CANVAS_H];
    s->num_shapes = t[DVG_H_NUM_SHAPES]; s->num_groups = t[DVG_H_NUM_GROUPS];
    s->num_params = t[DVG_H_NUM_PARAMS]; s->total_segs = t[DVG_H_TOTAL_SEGS];
    s->num_insts = t[DVG_H_TOTAL_GSHAPES];
    for (int g = 0; g < s->num_groups; g++)
        if (t[t[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN + DVG_G_FILL_TYPE] >= 0) s->has_fills = true;
    // instances and primitives in (group, shape-in-group, segment) order
    for (int g = 0; g < s->num_groups; g++) {
        const int32_t *r = t + t[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN;
        if (r[DVG_G_SHAPES_OFF] != (int)s->inst_group.size()) {
            delete s;
            return fail(DVG_ERR_INVALID, "group shape lists must be stored consecutively in group order");
        }
        const int32_t *ids = t + t[DVG_H_OFF_GSHAPES] + r[DVG_G_SHAPES_OFF];
        for (int k = 0; k < r[DVG_G_NUM_SHAPES]; k++) {
            const int sh = ids[k];
            const int32_t *sr = t + t[DVG_H_OFF_SHAPES] + sh * DVG_SHAPE_REC_LEN;
            const int inst = (int)s->inst_group.size();
            s->inst_group.push_back(g);
            s->inst_shape.push_back(sh);
            s->inst_prim_begin.push_back((int)s->prim_inst.size());
            if (sr[DVG_S_TYPE] == DVG_SHAPE_PATH) {
                const int32_t *ncp = t + t[DVG_H_OFF_NCP] + sr[DVG_S_NCP_OFF];
                int pid = 0;
                for (int seg = 0; seg < sr[DVG_S_NUM_SEGS]; seg++) {
                    s->prim_inst.push_back(inst);
                    s->prim_seg.push_back(seg);
                    s->prim_point_id.push_back(pid);
                    pid += ncp[seg] + 1;
                }
            } else {
                s->prim_inst.push_back(inst);
                s->prim_seg.push_back(0);
                s->prim_point_id.push_back(0);
            }
        }
    }
    s->inst_prim_begin.push_back((int)s->prim_inst.size());
    s->num_prims = (int)s->prim_inst.size();
    if (s->prim_inst.size() * (size_t)batch >= ((size_t)1 << 28)) { delete s; return fail(DVG_ERR_UNSUPPORTED, "more than 2^28 primitives (the pair queue packs the primitive id in 28 bits)"); }
    rc = upload(s->d_topo, s->topo);
    if (!rc) rc = upload(s->d_inst_group, s->inst_group);
    if (!rc) rc = upload(s->d_inst_shape, s->inst_shape);
    if (!rc) rc = upload(s->d_inst_prim_begin, s->inst_prim_begin);
    if (!rc) rc = upload(s->d_prim_inst, s->prim_inst);
    if (!rc) rc = upload(s->d_prim_seg, s->prim_seg);
    if (!rc) rc = upload(s->d_prim_point_id, s->prim_point_id);
    auto ens = [&](DevBuf &b, size_t bytes) { if (!rc && b.ensure(std::max<size_t>(bytes, 16)) != cudaSuccess) rc = fail(DVG_ERR_CUDA, "cudaMalloc failed"); };
    const size_t B = (size_t)batch;   // every derived table holds the scenes of a batch back to back
    const size_t ns = s->num_shapes * B, ng = s->num_groups * B, ni = s->num_insts * B, npr = s->num_prims * B, nsg = std::max(s->total_segs, 1) * B;
    ens(s->d_params, sizeof(float) * s->num_params * B);
    ens(s->d_shapes_length, 4 * ns); ens(s->d_shape_box, sizeof(Box) * ns); ens(s->d_shape_r0, 4 * ns);
    ens(s->d_seg_cdf, 4 * nsg); ens(s->d_seg_pmf, 4 * nsg); ens(s->d_seg_point_id, 4 * nsg);
    ens(s->d_insts, sizeof(InstInfo) * ni); ens(s->d_groups, sizeof(GroupInfo) * ng);
    ens(s->d_p01, 16 * npr); ens(s->d_p23, 16 * npr); ens(s->d_rad, 16 * npr); ens(s->d_box, 16 * npr);
    ens(s->d_thick, 4 * npr); ens(s->d_meta, sizeof(PrimMeta) * npr); ens(s->d_cbox, 16 * npr); ens(s->d_cbox_pf, 16 * npr); ens(s->d_cap, 16 * DVG_CAP_F4 * npr); ens(s->d_quint, sizeof(PrimQuintic) * npr); ens(s->d_wcert, sizeof(PrimWindCert) * (s->has_fills ? npr : 1));
    ens(s->d_shape_cdf, 4 * ni); ens(s->d_shape_pmf, 4 * ni); ens(s->d_flags, 16);
    ens(s->d_shape_guide, sizeof(int) * (DVG_CDF_GUIDE + 1));
    ens(s->d_scan_ws, sizeof(int) * (DVG_SCAN_WS_BLOCKS + 1));
    if (!rc && cudaMemset(s->d_scan_ws.p, 0, sizeof(int) * (DVG_SCAN_WS_BLOCKS + 1)) != cudaSuccess) rc = fail(DVG_ERR_CUDA, "cudaMemset failed");
    if (!rc && cudaMallocHost((void **)&s->h_pinned, 64) != cudaSuccess) rc = fail(DVG_ERR_CUDA, "cudaMallocHost failed");
    if (!rc && cudaMallocHost((void **)&s->h_counts, 32) != cudaSuccess) rc = fail(DVG_ERR_CUDA, "cudaMallocHost failed");
    for (int k = 0; k < 2 && !rc; k++)
        if (cudaEventCreateWithFlags(&s->ev_counts[k], cudaEventDisableTiming) != cudaSuccess) rc = fail(DVG_ERR_CUDA, "cudaEventCreate failed");
    if (!rc) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) dvg::g_num_sms = sms;
    }
    if (rc) { s->release_all(); delete s; return rc; }
    *out_scene = s;
    return DVG_OK;
}

int dvg_scene_set_params(DvgScene *s, const float *params, int64_t num_params, int params_on_device, void *stream) {
    if (!s || !params) return fail(DVG_ERR_INVALID, "null argument");
    if (num_params != (int64_t)s->num_params * s->batch) return fail(DVG_ERR_INVALID, "params length does not match the topology (x batch)");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (params_on_device) {
        CK(cudaMemcpyAsync(s->d_params.p, params, sizeof(float) * num_params, cudaMemcpyDeviceToDevice, st));
    } else {
        // pageable host memory -> pinned staging -> async H2D (the caller may reuse `params` on return)
        if (!s->h_params_pinned) {
            CK(cudaMallocHost((void **)&s->h_params_pinned, sizeof(float) * std::max<size_t>((size_t)num_params, 1)));
            CK(cudaEventCreateWithFlags(&s->h_params_free, cudaEventDisableTiming));
        } else {
            CK(cudaEventSynchronize(s->h_params_free));
        }
        memcpy(s->h_params_pinned, params, sizeof(float) * num_params);
        CK(cudaMemcpyAsync(s->d_params.p, s->h_params_pinned, sizeof(float) * num_params, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(s->h_params_free, st));
    }
    launch_build(s->build_view(), st);
    CK(cudaGetLastError());
    s->params_set = true;
    s->checked = false;
    s->scene_error = 0;
    s->bin_w = s->bin_h = 0;   // bins depend on the geometry
    s->wpx_valid = false; s->pfc_valid = false;
    s->w_valid = s->w_valid && true;  // the weight image does not depend on the scene, only on the filter (checked later)
    return DVG_OK;
}

static int sdf_args_check(const float *eval_positions, int num_eval_positions) {
    if ((eval_positions != nullptr) != (num_eval_positions > 0) || num_eval_positions < 0)
        return fail(DVG_ERR_INVALID, "eval_positions and num_eval_positions must be given together");
    return DVG_OK;
}

// Batch scenes: one seed per scene (host array), uploaded when it differs from the last call's.
static int bind_seeds(DvgScene *s, const uint64_t *seeds, RenderArgs &ra, cudaStream_t st) {
    if (s->batch == 1) return DVG_OK;
    if (!seeds) return fail(DVG_ERR_INVALID, "a batch scene needs one seed per scene (dvg_render_*_batch)");
    if (s->seeds_host.size() != (size_t)s->batch || memcmp(s->seeds_host.data(), seeds, 8 * (size_t)s->batch) != 0) {
        s->seeds_host.assign(seeds, seeds + s->batch);
        CK(s->d_seeds.ensure(8 * (size_t)s->batch));
        CK(cudaMemcpyAsync(s->d_seeds.p, s->seeds_host.data(), 8 * (size_t)s->batch, cudaMemcpyHostToDevice, st));
        s->seeds_version++;
    }
    ra.seeds = s->d_seeds.as<uint64_t>();
    ra.seed = s->seeds_version;   // stands in for the seeds in the reuse keys (weight image, forward result words)
    return DVG_OK;
}

static int check_batch_args(DvgScene *s, int width, int height, int nsx, int nsy, bool plain) {
    if (s->batch == 1) return DVG_OK;
    if (!plain) return fail(DVG_ERR_UNSUPPORTED, "batch scenes render colour images with the sampled path only (no prefiltering, SDF, eval_positions, d_translation or row ranges)");
    if ((int64_t)width * height * nsx * nsy * s->batch >= (int64_t)1 << 31)
        return fail(DVG_ERR_INVALID, "batch*width*height*samples must fit a 32-bit index");
    return DVG_OK;
}

static int render_forward_impl(DvgScene *s, const float *background, float *render_image, float *render_sdf,
                               int width, int height, int nsx, int nsy, uint64_t seed, int use_prefiltering,
                               const float *eval_positions, int num_eval_positions, int row_begin, int row_end, void *stream,
                               const uint64_t *seeds = nullptr) {
    int rc = check_render_args(s, width, height, nsx, nsy);
    if (rc) return rc;
    rc = check_batch_args(s, width, height, nsx, nsy, render_image && !render_sdf && !use_prefiltering && !eval_positions &&
                                                          row_begin == 0 && row_end == height);
    if (rc) return rc;
    if (!render_image && !render_sdf) return fail(DVG_ERR_INVALID, "render_image and render_sdf are both null");
    rc = sdf_args_check(eval_positions, num_eval_positions);
    if (rc) return rc;
    // diffvg.cpp:1504-1520: no weight image exists with eval_positions, so colour output is impossible there
    if (eval_positions && render_image) return fail(DVG_ERR_INVALID, "eval_positions can only be used with the SDF output");
    if (row_begin < 0 || row_end > height || row_begin > row_end) return fail(DVG_ERR_INVALID, "bad row range");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    // bins first: their read-back also completes the scene build (error flag, filter radius) -- one synchronisation
    rc = render_image ? ensure_bins(s, width, height, nsx * nsy, use_prefiltering ? 1 : 0, st, row_begin, row_end) : finish_build(s, st);
    if (rc) return rc;
    SceneView sc = s->view();
    RenderArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.width = width; ra.height = height; ra.nsx = nsx; ra.nsy = nsy; ra.seed = seed;
    ra.use_prefiltering = use_prefiltering; ra.row_begin = row_begin; ra.row_end = row_end;
    ra.background = background; ra.render_image = render_image;
    if (g_fast_accept) ra.flags |= DVG_RF_FAST_ACCEPT;
    rc = bind_seeds(s, seeds, ra, st);
    if (rc) return rc;
    if (render_image) {
        if (row_begin % s->bin_th != 0) return fail(DVG_ERR_INVALID, "row_begin must be a multiple of the tile height");
        rc = ensure_weight(s, sc, ra, row_begin, row_end, st);
        if (rc) return rc;
        CK(cudaMemsetAsync(render_image + 4 * (size_t)row_begin * width, 0,
                           sizeof(float) * 4 * (size_t)width * (row_end - row_begin) * s->batch, st));
        rc = use_prefiltering ? wave_pf_pass(s, sc, s->bin_view(), ra, false, st) : wave_pixel_pass(s, sc, s->bin_view(), ra, false, st);
        if (rc) return rc;
        CK(cudaGetLastError());
    }
    if (render_sdf) {
        if (row_begin != 0 || row_end != height) return fail(DVG_ERR_UNSUPPORTED, "the SDF output is not row-sharded");
        SdfArgs sa;
        sa.sdf = render_sdf; sa.d_sdf = nullptr; sa.eval_positions = eval_positions; sa.num_eval = num_eval_positions;
        const size_t n_out = eval_positions ? (size_t)num_eval_positions : (size_t)width * height;
        CK(cudaMemsetAsync(render_sdf, 0, sizeof(float) * n_out, st));
        launch_sdf(sc, ra, sa, false, st);
        CK(cudaGetLastError());
    }
    return DVG_OK;
}

int dvg_render_forward_rows(DvgScene *s, const float *background, float *render_image,
                            int width, int height, int nsx, int nsy, uint64_t seed,
                            int use_prefiltering, int row_begin, int row_end, void *stream) {
    if (!render_image) return fail(DVG_ERR_INVALID, "render_image is null");
    return render_forward_impl(s, background, render_image, nullptr, width, height, nsx, nsy, seed, use_prefiltering,
                               nullptr, 0, row_begin, row_end, stream);
}

int dvg_render_forward(DvgScene *s, const float *background, float *render_image, float *render_sdf,
                       int width, int height, int nsx, int nsy, uint64_t seed,
                       int use_prefiltering, const float *eval_positions, int num_eval_positions, void *stream) {
    return render_forward_impl(s, background, render_image, render_sdf, width, height, nsx, nsy, seed, use_prefiltering,
                               eval_positions, num_eval_positions, 0, height, stream);
}

static int render_backward_impl(DvgScene *s, const float *background, const float *d_render_image, const float *d_render_sdf,
                                int width, int height, int nsx, int nsy, uint64_t seed, int use_prefiltering,
                                const float *eval_positions, int num_eval_positions, int row_begin, int row_end,
                                float *d_params, float *d_background, float *d_translation, uint32_t flags, void *stream,
                                const uint64_t *seeds = nullptr) {
    int rc = check_render_args(s, width, height, nsx, nsy);
    if (rc) return rc;
    rc = check_batch_args(s, width, height, nsx, nsy, d_render_image && !d_render_sdf && !use_prefiltering && !eval_positions &&
                                                          !d_translation && row_begin == 0 && row_end == height);
    if (rc) return rc;
    if (!d_params) return fail(DVG_ERR_INVALID, "d_params is null");
    if (!d_render_image && !d_render_sdf) return fail(DVG_ERR_INVALID, "d_render_image and d_render_sdf are both null");
    if (d_render_image && ((uintptr_t)d_render_image & 15) != 0) return fail(DVG_ERR_INVALID, "d_render_image must be 16-byte aligned");
    rc = sdf_args_check(eval_positions, num_eval_positions);
    if (rc) return rc;
    if (eval_positions && d_render_image) return fail(DVG_ERR_INVALID, "eval_positions can only be used with the SDF output");
    if (row_begin < 0 || row_end > height || row_begin > row_end) return fail(DVG_ERR_INVALID, "bad row range");
    const bool whole = row_begin == 0 && row_end == height;
    if (!whole && (d_render_sdf || d_translation)) return fail(DVG_ERR_UNSUPPORTED, "the SDF output / d_translation are not row-sharded");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    rc = d_render_image ? ensure_bins(s, width, height, nsx * nsy, use_prefiltering ? 1 : 0, st, row_begin, row_end) : finish_build(s, st);
    if (rc) return rc;
    SceneView sc = s->view();
    RenderArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.width = width; ra.height = height; ra.nsx = nsx; ra.nsy = nsy; ra.seed = seed;
    ra.use_prefiltering = use_prefiltering; ra.row_begin = row_begin; ra.row_end = row_end;
    ra.flags = flags | (g_fast_accept ? DVG_RF_FAST_ACCEPT : 0u);
    ra.background = background; ra.d_render_image = d_render_image;
    ra.d_params = d_params; ra.d_background = d_background; ra.d_translation = d_translation;
    ra.debug_out = g_debug_out;
    rc = bind_seeds(s, seeds, ra, st);
    if (rc) return rc;
    const size_t all_params = (size_t)s->num_params * s->batch;   // a batch: scene b's gradients at b * num_params
    if (!(flags & DVG_BWD_ACCUMULATE)) CK(cudaMemsetAsync(d_params, 0, sizeof(float) * all_params, st));
    const bool wave_grads = d_render_image != nullptr;
    if (wave_grads) {   // private copies of the gradient buffer for the composite kernels (dvg_wave.cu grad_replica)
        // the copies exist to spread same-address atomics; the scenes of a batch already spread them
        ra.grad_reps = s->batch >= 8 ? 4 : 32; ra.num_params = (int)all_params;
        CK(s->d_grad_rep.ensure(sizeof(float) * (size_t)ra.grad_reps * all_params));
        ra.d_params_rep = s->d_grad_rep.as<float>();
        CK(cudaMemsetAsync(ra.d_params_rep, 0, sizeof(float) * (size_t)ra.grad_reps * all_params, st));
    }
    if (d_background)
        CK(cudaMemsetAsync(d_background + 4 * (size_t)row_begin * width, 0,
                           sizeof(float) * 4 * (size_t)width * (row_end - row_begin) * s->batch, st));
    if (d_translation) CK(cudaMemsetAsync(d_translation, 0, sizeof(float) * 2 * (size_t)width * height, st));
    if (d_render_image) {
        if (row_begin % s->bin_th != 0) return fail(DVG_ERR_INVALID, "row_begin must be a multiple of the tile height");
        BinView bins = s->bin_view();
        rc = use_prefiltering ? ensure_weight(s, sc, ra, row_begin, row_end, st) : ensure_weight(s, sc, ra, 0, height, st);
        if (rc) return rc;
        if (use_prefiltering) {
            // interior term only: the SDF coverage is differentiable, no boundary pass (diffvg.cpp:1558)
            rc = wave_pf_pass(s, sc, bins, ra, true, st);
            if (rc) return rc;
            launch_wave_reduce_grads(ra, st);
            CK(cudaGetLastError());
        } else {
            rc = wave_pixel_pass(s, sc, bins, ra, true, st);
            if (rc) return rc;
            // boundary term (diffvg.cpp:1558-1626): boundary-sample indices of the owned rows
            const int spp = nsx * nsy;
            const int ntiles = bin_total_tiles(bins);
            const int64_t all_begin = (int64_t)row_begin * width * spp;
            const int64_t all_count = (int64_t)(row_end - row_begin) * width * spp * s->batch;   // (batch: whole images, scene after scene)
            int64_t per_pass = edge_pass_samples(s, ntiles);
            if (g_debug_edge_pass_samples > 0) per_pass = std::min<int64_t>(per_pass, g_debug_edge_pass_samples);
            if (all_count > 0 && per_pass < 4096)
                return fail(DVG_ERR_UNSUPPORTED, "a tile holds too many candidate primitives for the boundary pass at this render size");
            for (int64_t done = 0; done < all_count; done += per_pass) {
                BoundaryWork bw;
                bw.samples = nullptr; bw.samples_unsorted = nullptr; bw.item_tile = nullptr;
                bw.scan_ws = s->d_scan_ws.as<int>();
                bw.sample_begin = (int)(all_begin + done);
                bw.num_samples = (int)std::min<int64_t>(per_pass, all_count - done);
                CK(s->d_keys.ensure(sizeof(int) * (size_t)bw.num_samples));
                CK(s->d_sorted.ensure(sizeof(int) * (size_t)bw.num_samples));
                CK(s->d_tile_counts.ensure(sizeof(int) * ntiles)); CK(s->d_tile_fill.ensure(sizeof(int) * ntiles));
                CK(s->d_blk_counts.ensure(sizeof(int) * ntiles));
                CK(s->d_tile_offsets.ensure(sizeof(int) * (ntiles + 1))); CK(s->d_blk_offsets.ensure(sizeof(int) * (ntiles + 1)));
                bw.keys = s->d_keys.as<int>(); bw.sorted_idx = s->d_sorted.as<int>();
                bw.tile_counts = s->d_tile_counts.as<int>(); bw.tile_fill = s->d_tile_fill.as<int>();
                bw.blk_counts = s->d_blk_counts.as<int>();
                bw.tile_offsets = s->d_tile_offsets.as<int>(); bw.blk_offsets = s->d_blk_offsets.as<int>();
                rc = wave_edge_pass(s, sc, bins, ra, bw, st);
                if (rc) return rc;
            }
            launch_wave_reduce_grads(ra, st);
            CK(cudaGetLastError());
        }
    }
    if (d_render_sdf) {
        SdfArgs sa;
        sa.sdf = nullptr; sa.d_sdf = d_render_sdf; sa.eval_positions = eval_positions; sa.num_eval = num_eval_positions;
        launch_sdf(sc, ra, sa, true, st);
        CK(cudaGetLastError());
    }
    return DVG_OK;
}

int dvg_render_backward_rows(DvgScene *s, const float *background, const float *d_render_image,
                             int width, int height, int nsx, int nsy, uint64_t seed,
                             int use_prefiltering, int row_begin, int row_end,
                             float *d_params, float *d_background, uint32_t flags, void *stream) {
    if (!d_render_image) return fail(DVG_ERR_INVALID, "d_render_image is null");
    return render_backward_impl(s, background, d_render_image, nullptr, width, height, nsx, nsy, seed, use_prefiltering,
                                nullptr, 0, row_begin, row_end, d_params, d_background, nullptr, flags, stream);
}

int dvg_render_backward(DvgScene *s, const float *background, const float *d_render_image, const float *d_render_sdf,
                        int width, int height, int nsx, int nsy, uint64_t seed,
                        int use_prefiltering, const float *eval_positions, int num_eval_positions,
                        float *d_params, float *d_background, float *d_translation, uint32_t flags, void *stream) {
    return render_backward_impl(s, background, d_render_image, d_render_sdf, width, height, nsx, nsy, seed, use_prefiltering,
                                eval_positions, num_eval_positions, 0, height, d_params, d_background, d_translation, flags, stream);
}

int dvg_render_forward_batch(DvgScene *s, const float *background, float *render_image, int width, int height, int nsx, int nsy,
                             const uint64_t *seeds, void *stream) {
    if (!render_image) return fail(DVG_ERR_INVALID, "render_image is null");
    if (s && s->batch == 1) return render_forward_impl(s, background, render_image, nullptr, width, height, nsx, nsy, seeds ? seeds[0] : 0, 0,
                                                       nullptr, 0, 0, height, stream);
    return render_forward_impl(s, background, render_image, nullptr, width, height, nsx, nsy, 0, 0, nullptr, 0, 0, height, stream, seeds);
}

int dvg_render_backward_batch(DvgScene *s, const float *background, const float *d_render_image, int width, int height, int nsx, int nsy,
                              const uint64_t *seeds, float *d_params, float *d_background, uint32_t flags, void *stream) {
    if (!d_render_image) return fail(DVG_ERR_INVALID, "d_render_image is null");
    if (s && s->batch == 1) return render_backward_impl(s, background, d_render_image, nullptr, width, height, nsx, nsy, seeds ? seeds[0] : 0, 0,
                                                        nullptr, 0, 0, height, d_params, d_background, nullptr, flags, stream);
    return render_backward_impl(s, background, d_render_image, nullptr, width, height, nsx, nsy, 0, 0, nullptr, 0, 0, height,
                                d_params, d_background, nullptr, flags, stream, seeds);
}

int dvg_scene_row_costs(DvgScene *s, int width, int height, int nsx, int nsy, int use_prefiltering, float *out_host, int cap,
                        int *tile_h_out, void *stream) {
    int rc = check_render_args(s, width, height, nsx, nsy);
    if (rc) return rc;
    if (!out_host || !tile_h_out) return fail(DVG_ERR_INVALID, "null argument");
    if (s->batch > 1) return fail(DVG_ERR_UNSUPPORTED, "batch scenes are split by scene, not by rows");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    rc = ensure_bins(s, width, height, nsx * nsy, use_prefiltering ? 1 : 0, st, 0, height);   // whole image
    if (rc) return rc;
    const BinView bins = s->bin_view();
    if (cap < bins.tiles_y) return fail(DVG_ERR_INVALID, "row-cost buffer too small");
    CK(s->d_item_tile.ensure(sizeof(float) * (size_t)bins.tiles_y));   // (scratch: rebound by the next boundary pass)
    launch_tile_row_costs(bins.offsets, bins.tiles_x, bins.tiles_y, s->d_item_tile.as<float>(), st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_host, s->d_item_tile.p, sizeof(float) * (size_t)bins.tiles_y, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *tile_h_out = bins.tile_h;
    return bins.tiles_y > 0 ? DVG_OK : fail(DVG_ERR_INVALID, "empty image");
}

int dvg_debug_set_boundary_dump(float *device_buf) { g_debug_out = device_buf; return DVG_OK; }

int dvg_debug_set_limits(int64_t pair_capacity, int64_t edge_pass_samples) {
    g_debug_pair_capacity = pair_capacity; g_debug_edge_pass_samples = edge_pass_samples;
    return DVG_OK;
}

int dvg_debug_prim_tests(DvgScene *s, int width, int height, int nsx, int nsy, uint64_t seed, int x, int y,
                         int32_t *out_host, float *pos_host, void *stream) {
    int rc = check_render_args(s, width, height, nsx, nsy);
    if (rc) return rc;
    if (!out_host || !pos_host || x < 0 || y < 0 || x >= width || y >= height) return fail(DVG_ERR_INVALID, "bad argument");
    if (s->batch > 1) return fail(DVG_ERR_UNSUPPORTED, "dvg_debug_prim_tests reads single scenes");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    rc = finish_build(s, st);
    if (rc) return rc;
    rc = ensure_bins(s, width, height, nsx * nsy, 0, st, 0, height);
    if (rc) return rc;
    RenderArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.width = width; ra.height = height; ra.nsx = nsx; ra.nsy = nsy; ra.seed = seed; ra.row_end = height;
    const size_t n = (size_t)nsx * nsy * s->num_prims;
    int *d_out = nullptr; float *d_pos = nullptr;
    CK(cudaMalloc((void **)&d_out, n * 4));
    CK(cudaMalloc((void **)&d_pos, (size_t)nsx * nsy * 8));
    launch_debug_prim_tests(s->view(), s->bin_view(), ra, x, y, d_out, d_pos, st);
    CK(cudaMemcpyAsync(out_host, d_out, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pos_host, d_pos, (size_t)nsx * nsy * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_out); cudaFree(d_pos);
    return DVG_OK;
}

int dvg_set_fast_stroke_accept(int on) { g_fast_accept = on != 0; return DVG_OK; }
int dvg_debug_set_prefilter_inline(int on) { g_pf_inline = (on & 1) != 0; g_pf_nocache = (on & 2) != 0; return DVG_OK; }

int dvg_profile_enable(int on) {
    dvg::g_profile_on = on != 0;
    return DVG_OK;
}

int64_t dvg_profile_report(char *buf, int64_t cap) {
    if (!buf || cap <= 0) { fail(DVG_ERR_INVALID, "null buffer"); return -1; }
    int n = dvg::prof_report(buf, cap);
    if (n < 0) fail(DVG_ERR_INVALID, "profile buffer too small");
    return n;
}

int dvg_measure_peak(int which, int device, double *tflops) {
    if (!tflops || which < 0 || which > 1) return fail(DVG_ERR_INVALID, "bad argument");
    DeviceGuard guard(device);
    float *d = nullptr;
    CK(cudaMalloc((void **)&d, 64));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int iters = which == 0 ? 4096 : 2048;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(a, 0));
        launch_peak_probe(which, d, iters, 0);
        CK(cudaEventRecord(b, 0));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        double tf = dvg::peak_probe_flops(iters, 148 * 8) / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    *tflops = best;
    return DVG_OK;
}

int64_t dvg_scene_dump(DvgScene *s, int what, int index, uint32_t *out, int64_t cap, void *stream) {
    // every selector is copied back from DEVICE memory: the tables the kernels read (3-10) and the
    // reference-topology trees built by dvg_bvh.cu (0-2)
    auto bad = [&](const char *msg) { fail(DVG_ERR_INVALID, msg); return (int64_t)-1; };
    if (!s || !out) return bad("null argument");
    if (!s->params_set) return bad("dvg_scene_set_params has not been called");
    if (s->batch > 1) return bad("dvg_scene_dump reads single scenes");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t *t = s->topo.data();
    const void *src = nullptr;
    int64_t words = 0;
    auto shape_rec = [&](int sh) { return t + t[DVG_H_OFF_SHAPES] + sh * DVG_SHAPE_REC_LEN; };
    if (what >= 0 && what <= 2) {
        DevBuf &dp = s->d_bvh_path, &dg = s->d_bvh_group, &ds = s->d_bvh_scene, &dk = s->d_bvh_keys;
        if (dp.ensure(sizeof(BvhNode) * 2 * (size_t)std::max(s->total_segs, 1)) != cudaSuccess ||
            dg.ensure(sizeof(BvhNode) * 2 * (size_t)s->num_insts) != cudaSuccess ||
            ds.ensure(sizeof(BvhNode) * 2 * (size_t)s->num_groups) != cudaSuccess ||
            dk.ensure(8 * bvh_key_words(s->total_segs, s->num_insts, s->num_groups)) != cudaSuccess) {
            fail(DVG_ERR_CUDA, "cudaMalloc failed");
            return -1;
        }
        launch_bvh_build(s->build_view(), dp.as<BvhNode>(), dg.as<BvhNode>(), ds.as<BvhNode>(), dk.as<unsigned long long>(), st);
        if (what == 0) { src = ds.p; words = 7 * (2 * (int64_t)s->num_groups - 1); }
        else if (what == 1) {
            if (index < 0 || index >= s->num_groups) return bad("group index out of range");
            const int32_t *r = t + t[DVG_H_OFF_GROUPS] + index * DVG_GROUP_REC_LEN;
            src = dg.as<BvhNode>() + 2 * r[DVG_G_SHAPES_OFF];
            words = 7 * (2 * (int64_t)r[DVG_G_NUM_SHAPES] - 1);
        } else {
            if (index < 0 || index >= s->num_shapes || shape_rec(index)[DVG_S_TYPE] != DVG_SHAPE_PATH) return bad("shape index is not a path");
            src = dp.as<BvhNode>() + 2 * shape_rec(index)[DVG_S_NCP_OFF];
            words = 7 * (2 * (int64_t)shape_rec(index)[DVG_S_NUM_SEGS] - 1);
        }
    } else if (what >= 3 && what <= 5) {
        src = what == 3 ? s->d_shapes_length.p : (what == 4 ? s->d_shape_cdf.p : s->d_shape_pmf.p);
        words = what == 3 ? s->num_shapes : s->num_insts;
    } else if (what >= 6 && what <= 8) {
        if (index < 0 || index >= s->num_shapes || shape_rec(index)[DVG_S_TYPE] != DVG_SHAPE_PATH) return bad("shape index is not a path");
        const int off = shape_rec(index)[DVG_S_NCP_OFF];
        src = what == 6 ? (const void *)(s->d_seg_cdf.as<float>() + off)
                        : (what == 7 ? (const void *)(s->d_seg_pmf.as<float>() + off) : (const void *)(s->d_seg_point_id.as<int>() + off));
        words = shape_rec(index)[DVG_S_NUM_SEGS];
    } else if (what == 9 || what == 10) {
        src = what == 9 ? s->d_inst_shape.p : s->d_inst_group.p;
        words = s->num_insts;
    } else {
        return bad("bad dump selector");
    }
    if (words > cap) return bad("dump buffer too small");
    if (cudaMemcpyAsync(out, src, (size_t)words * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        fail(DVG_ERR_CUDA, std::string("dvg_scene_dump: ") + cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    return words;
}

int dvg_scene_destroy(DvgScene *s) {
    if (!s) return DVG_OK;
    {
        DeviceGuard guard(s->device);
        s->release_all();
    }
    delete s;
    return DVG_OK;
}

}  // extern "C"
