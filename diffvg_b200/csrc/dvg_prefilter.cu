// dvg_prefilter.cu -- the SDF-prefiltering variant of the render kernel
// (sample_color_prefiltered, diffvg.cpp:835-1113; selected by use_prefiltering) and the SDF
// output (sample_distance, diffvg.cpp:709-775; OutputType.sdf / eval_positions).
//
// Same tiling as dvg_render.cu: a block owns (part of) a pixel tile, a thread owns one sample and
// walks the tile's candidate list (prefilter bins: a stroked group contributes all its segments
// wherever the group is visited, because its closest-point search is unbounded; a fill-only
// group the segments within radius 1 plus those its winding ray can cross).  Per sample the state
// is PrefilterTracer (dvg_distance.cuh); the backward variant keeps the <= 64 fragment records
// of the reference in local memory and scatters through the block's shared-memory GradCache.
#include "dvg_internal.h"
#include "dvg_kernel_util.cuh"

namespace dvg {

// (an opportunistic warp-level sum before the atomics -- all 32 lanes adding to one address inside a large filled shape --
// measured no faster: 4.42 vs 4.46 ms for the cached backward kernel at flower.svg 2048^2)
typedef GlobalSink PfSink;
constexpr int PB = 256;  // threads per block
#ifndef DVG_PF_MINB
#define DVG_PF_MINB 2
#endif
#ifndef DVG_PF_FWD_MINB
#define DVG_PF_FWD_MINB 2
#endif
// resident blocks per SM of the forward kernel with the winding words: two (with the fragment-cache stores three blocks spill:
// 3.72 vs 3.56 ms at flower.svg 2048^2; without the stores three were faster, 2.95 vs 3.40); the full backward kernel
// (fragment records, distance gradients) and the inline forms are slower with three; the cached backward kernel takes three

DVG_D PrimRef load_prim(const SceneView &sc, int e) {
    PrimRef pr;
    const PrimMeta pm = sc.prim_meta[e];
    pr.p01 = sc.prim_p01[e]; pr.p23 = sc.prim_p23[e];
    pr.rad = mk4(0, 0, 0, 0);
    pr.box = sc.prim_box[e]; pr.thick = 0.f;
    pr.tf = pm.type_flags; pr.inst = pm.inst; pr.group = sc.insts[pm.inst].group;
    pr.base_id = pm.base_id; pr.point_id = pm.point_id;
    pr.cap = reinterpret_cast<const float *>(sc.prim_cap + (size_t)e * DVG_CAP_F4);
    return pr;
}

// Geometry of one thread's sample: block -> (tile, part), thread -> sample of the tile (the layout of pixel_item, dvg_wave.cu).
struct PfSample {
    int tile, l, x, y, gthread;
    bool active;
    F2 pt, cpt;
    const float *bg_px;
    F4 first;
};
DVG_D PfSample pf_sample(const SceneView &sc, const BinView &bins, const RenderArgs &ra) {
    PfSample ps;
    const int tile_row0 = ra.row_begin / bins.tile_h;
    const int spp = ra.nsx * ra.nsy;
    const int ns = bins.tile_w * bins.tile_h * spp;
    const int parts = (ns + PB - 1) / PB;
    ps.tile = blockIdx.x / parts + tile_row0 * bins.tiles_x;
    const int part = blockIdx.x % parts;
    const int tx = ps.tile % bins.tiles_x, ty = ps.tile / bins.tiles_x;
    ps.l = part * PB + threadIdx.x;
    ps.gthread = blockIdx.x * PB + threadIdx.x;
    const int s = ps.l % spp, p = ps.l / spp;
    ps.x = tx * bins.tile_w + p % bins.tile_w;
    ps.y = ty * bins.tile_h + p / bins.tile_w;
    ps.active = ps.l < ns && ps.x < ra.width && ps.y < ra.height && ps.y >= ra.row_begin && ps.y < ra.row_end;
    ps.pt = mk2(0, 0); ps.cpt = mk2(0, 0);
    ps.bg_px = nullptr;
    ps.first = mk4(0, 0, 0, 0);
    if (ps.active) {
        const int sx = s % ra.nsx, sy = s / ra.nsx;
        const int idx = ((ps.y * ra.width + ps.x) * ra.nsy + sy) * ra.nsx + sx;
        sample_position(sc.canvas_w, sc.canvas_h, ra.width, ra.height, ra.nsx, ra.nsy, ra.seed, true, ps.x, ps.y, sx, sy, idx, ps.pt, ps.cpt);
        if (ra.background) {
            ps.bg_px = ra.background + 4 * (ps.y * ra.width + ps.x);
            ps.first = mk4(ps.bg_px[0], ps.bg_px[1], ps.bg_px[2], ps.bg_px[3]);
        }
    }
    return ps;
}

// The backward pass of one sample from its fragment records (diffvg.cpp:985-1111) and what surrounds it in the kernels.
template <typename Tracer, typename Sink>
DVG_D void pf_sample_backward(const SceneView &sc, const RenderArgs &ra, const PfSample &ps, const Tracer &tr, F4 color, const Sink &sk) {
    const F4 d_color = gather_d_color(sc.filter, ra.d_render_image, ra.weight_image, ra.width, ra.height, ps.pt);
    float *dtr = ra.d_translation ? ra.d_translation + 2 * (ps.y * ra.width + ps.x) : nullptr;
    if (tr.nfrag > 0) {
        F4 d_bg;
        prefilter_backward(sc, tr, color, d_color, sk, dtr, d_bg);
        if (ps.bg_px && ra.d_background) {
            float *d = ra.d_background + 4 * (ps.y * ra.width + ps.x);
            atomicAdd(d + 0, d_bg.x); atomicAdd(d + 1, d_bg.y); atomicAdd(d + 2, d_bg.z); atomicAdd(d + 3, d_bg.w);
        }
    } else if (ps.bg_px && ra.d_background) {
        float *d = ra.d_background + 4 * (ps.y * ra.width + ps.x);
        atomicAdd(d + 0, d_color.x); atomicAdd(d + 1, d_color.y); atomicAdd(d + 2, d_color.z); atomicAdd(d + 3, d_color.w);
    }
}

// WORDS: `relevant` (one word per (sample, chunk), the same in every lane of a warp: the candidates some sample of the warp
// can be affected by, wave_classify<.., PF>) and the winding contributions come from the winding pre-pass
// (wave_classify<.., FILLS, PF> -> k_wave_solve_fill, dvg_wave.cu): `wind` holds one 4-bit answer per (sample, candidate) in
// the layout of the wavefront passes (warp = item of 32 samples; word (cb + chunk) * 32 + lane, four 32-bit words each).
// Inline, the FP64 root solves of the winding test ran with the lanes that happened to need them, in a 128-register kernel
// at 22% occupancy, and the backward kernel repeated all of them; now they run one per lane in k_wave_solve_fill and the
// backward pass re-uses the forward pass's words.
// `pc` (fragment cache, dvg_distance.cuh): the FORWARD kernel leaves every sample's first DVG_PFC_K fragment records and its
// fragment count there; the BACKWARD kernel, when given one, only differentiates the samples with more fragments than that
// (k_pf_backward_cached has done the others) and its warps leave at once when they hold none.
template <bool BACKWARD, bool WORDS>
__global__ void __launch_bounds__(PB, (!BACKWARD && WORDS) ? DVG_PF_FWD_MINB : DVG_PF_MINB) k_render_pf(SceneView sc, BinView bins, RenderArgs ra, const uint4 *wind, const unsigned *relevant, const int *tile_choff, PfCache pc) {
    // gradients go straight to one of the private copies of the gradient buffer (dvg_kernel_util.cuh grad_replica):
    // the per-block shared-memory hash + barrier + flush this kernel used before was 40% of its time at 2048^2
    const PfSink sk{BACKWARD ? grad_replica(ra) : nullptr};
    const int spp = ra.nsx * ra.nsy;
    const int ns = bins.tile_w * bins.tile_h * spp;
    const int tid = threadIdx.x;
    const bool pow2 = (spp & (spp - 1)) == 0;
    const int grp = pow2 ? (spp < 32 ? spp : 32) : 1;
    PfFragment frags[BACKWARD ? DVG_MAXPF : 1];
    float d_radius_acc = 0.f;
    PfSample ps = pf_sample(sc, bins, ra);
    if (BACKWARD && pc.count) {   // the samples the cached kernel could not take
        ps.active = ps.active && pc.count[ps.gthread] > DVG_PFC_K;
        if (!__any_sync(0xffffffffu, ps.active)) return;
    }
    const int tile = ps.tile, l = ps.l, x = ps.x, y = ps.y;
    const bool active = ps.active;
    const F2 pt = ps.pt, cpt = ps.cpt;
    const float *bg_px = ps.bg_px;
    PrefilterTracer<BACKWARD ? 1 : 2> tr;
    tr.init(cpt, active, ps.first, frags);
    if (!BACKWARD && pc.recs) tr.crec = reinterpret_cast<U4 *>(pc.recs) + ((size_t)(ps.gthread >> 5) * (DVG_PFC_K * 64) + (tid & 31));
    const int beg = bins.offsets[tile], end = bins.offsets[tile + 1];
    if constexpr (WORDS) {
        const int wpt = (ns + 31) / 32;
        const int part_w = l >> 5;                 // this warp's item within the tile (pixel_item, dvg_wave.cu)
        const int c0 = tile_choff[tile];
        const int nch = tile_choff[tile + 1] - c0;
        const int64_t w0 = ((int64_t)c0 * wpt + (int64_t)part_w * nch) * 32 + (tid & 31);
        const bool have_words = part_w < wpt;
        for (int c = 0; c < nch; c++) {
            // candidates that matter to some sample of this warp (the same word in every lane); the others change no
            // state: no distance within reach, no winding contribution, and a group or shape none of whose candidates is
            // visited emits nothing
            unsigned m = have_words ? relevant[w0 + (int64_t)c * 32] : 0u;
            if (m == 0u) continue;
            const uint4 wd = wind[w0 + (int64_t)c * 32];
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1u;
                const PrimRef pr = load_prim(sc, bins.items[beg + c * 32 + k]);   // warp-uniform loads
                const unsigned ww = (k < 8 ? wd.x : (k < 16 ? wd.y : (k < 24 ? wd.z : wd.w)));
                const int nib = (int)((ww >> (4 * (k & 7))) & 15u);
                tr.template step<true>(sc, pr, (nib ^ 8) - 8);
            }
        }
    } else {
        for (int i = beg; i < end; i++) {
            const PrimRef pr = load_prim(sc, bins.items[i]);   // block-uniform loads
            tr.step(sc, pr);
        }
    }
    tr.finish(sc);
    const F4 color = tr.resolve(bg_px);
    if constexpr (!BACKWARD) {
        if (pc.count) pc.count[ps.gthread] = active ? tr.nfrag : 0;
        splat_color(sc, ra, x, y, pt, color, active, grp, tid);
    } else {
        if (active) pf_sample_backward(sc, ra, ps, tr, color, sk);
        if (!(ra.flags & 4u)) {   // DVG_BWD_SKIP_FILTER_GRAD (block-uniform: every lane of the warp is here)
            // (the box form sums over the lanes of a pixel: only when every sample of the launch is taken here)
            const bool box_fast = !pc.count && sc.filter.type == 0 && pow2 && (int)ceilf(sc.filter.radius) == 1;
            if (box_fast) d_radius_acc = filter_radius_grad_box(sc, ra, x, y, pt, color, active, grp, tid & 31);
            else if (active) d_radius_acc = filter_radius_grad(sc, ra, x, y, pt, color);
        }
        d_radius_acc = warp_sum(d_radius_acc);
        if ((tid & 31) == 0) sk.add(sc.filter_radius_off, d_radius_acc);
    }
}

// Backward pass of the samples whose fragments the forward kernel cached (at most DVG_PFC_K of them: 99-100% of the samples
// of the fill scenes): no candidate walk, no closest-point search -- the records are read back, a replay of the compositing
// gives every fragment's `prev` (the same functions on the same floats as the forward kernel's, pf_fragment_color /
// pf_composite), and prefilter_backward differentiates as it does in the full kernel.
#ifndef DVG_PFC_MINB
#define DVG_PFC_MINB 3
#endif
__global__ void __launch_bounds__(PB, DVG_PFC_MINB) k_pf_backward_cached(SceneView sc, BinView bins, RenderArgs ra, PfCache pc) {
    const PfSink sk{grad_replica(ra)};
    const int spp = ra.nsx * ra.nsy;
    const int tid = threadIdx.x;
    const bool pow2 = (spp & (spp - 1)) == 0;
    const int grp = pow2 ? (spp < 32 ? spp : 32) : 1;
    const PfSample ps = pf_sample(sc, bins, ra);
    const int n = ps.active ? pc.count[ps.gthread] : 0;
    const bool mine = ps.active && n <= DVG_PFC_K;
    PfFragment frags[DVG_PFC_K];
    PrefilterTracer<1> tr;
    tr.init(ps.cpt, mine, ps.first, frags);
    if (mine) {
        const U4 *rec = reinterpret_cast<const U4 *>(pc.recs) + ((size_t)(ps.gthread >> 5) * (DVG_PFC_K * 64) + (tid & 31));
#pragma unroll
        for (int j = 0; j < DVG_PFC_K; j++) {
            if (j < n) {
                pf_cache_unpack(rec[j * 64], rec[j * 64 + 32], frags[j]);
                frags[j].prev = tr.accum;
                const bool is_stroke = (frags[j].key & 1) != 0;
                const float w = pf_coverage(sc, is_stroke, frags[j].inst, frags[j].d);
                pf_composite(tr.accum, pf_fragment_color(sc, sc.groups[frags[j].key >> 1], is_stroke, w, ps.cpt));
                tr.nfrag++; tr.sp++;
            }
        }
    }
    const F4 color = tr.resolve(ps.bg_px);
    if (mine) pf_sample_backward(sc, ra, ps, tr, color, sk);
    float d_radius_acc = 0.f;
    if (!(ra.flags & 4u)) {   // DVG_BWD_SKIP_FILTER_GRAD
        const bool box_fast = sc.filter.type == 0 && pow2 && (int)ceilf(sc.filter.radius) == 1;
        // (box form: the neighbour sum of a pixel is formed by ALL the lanes of the pixel, also those the full kernel takes)
        if (box_fast) d_radius_acc = filter_radius_grad_box(sc, ra, ps.x, ps.y, ps.pt, color, ps.active, grp, tid & 31);
        else if (mine) d_radius_acc = filter_radius_grad(sc, ra, ps.x, ps.y, ps.pt, color);
        if (!mine) d_radius_acc = 0.f;
    }
    d_radius_acc = warp_sum(d_radius_acc);
    if ((tid & 31) == 0) sk.add(sc.filter_radius_off, d_radius_acc);
}

// sample_distance for every pixel sample (eval_positions == null) or every evaluation position.
// One thread per sample; groups are searched from the last to the first with a strict `<`, as in
// diffvg.cpp:726-739.  SDF scenes are small (the reference loops over ALL groups without any
// culling here), so there is no binning.
template <bool BACKWARD>
__global__ void __launch_bounds__(PB) k_sdf(SceneView sc, RenderArgs ra, SdfArgs sa) {
    GradCache *gcp = nullptr;
    if constexpr (BACKWARD) {
        __shared__ GradCache s_gc;
        gcp = &s_gc;
        grad_cache_init(s_gc);
        __syncthreads();
    }
    const CacheSink sk{gcp, ra.d_params};
    const int spp = ra.nsx * ra.nsy;
    const int n = sa.eval_positions ? sa.num_eval : ra.width * ra.height * spp;
    const int idx = blockIdx.x * PB + threadIdx.x;
    if (idx < n) {
        F2 pt, cpt;
        int x, y;
        if (!sa.eval_positions) {
            const int sx = idx % ra.nsx, sy = (idx / ra.nsx) % ra.nsy;
            x = (idx / spp) % ra.width; y = idx / (spp * ra.width);
            sample_position(sc.canvas_w, sc.canvas_h, ra.width, ra.height, ra.nsx, ra.nsy, ra.seed, ra.use_prefiltering != 0,
                            x, y, sx, sy, idx, pt, cpt);
        } else {
            pt = mk2(sa.eval_positions[2 * idx], sa.eval_positions[2 * idx + 1]);
            x = (int)pt.x; y = (int)pt.y;
            F2 npt = pt;
            npt.x /= ra.width; npt.y /= ra.height;
            cpt = mk2(npt.x * sc.canvas_w, npt.y * sc.canvas_h);
        }
        const float weight = sa.eval_positions ? 1.f : 1.f / spp;
        int min_g = -1;
        DistHit best;
        dist_hit_init(best, 0.f);
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
            if constexpr (BACKWARD) {
                const float dd = sa.eval_positions ? sa.d_sdf[idx] : sa.d_sdf[y * ra.width + x];
                const float d_abs = inside ? -dd : dd;
                // d_translation is indexed by the pixel (diffvg.cpp:1281); out-of-image evaluation positions
                // would write out of bounds in the reference -- skipped here
                float *dtr = nullptr;
                if (ra.d_translation && x >= 0 && x < ra.width && y >= 0 && y < ra.height) dtr = ra.d_translation + 2 * (y * ra.width + x);
                d_compute_distance(sc, sc.groups[min_g], best.inst, cpt, best.cp, best.base_id, best.point_id, best.t_root, d_abs, sk, dtr);
            }
        }
        if (!BACKWARD && sa.sdf) {
            if (sa.eval_positions) sa.sdf[idx] = dist;   // one sample per slot
            else atomicAdd(sa.sdf + y * ra.width + x, dist);
        }
    }
    if (BACKWARD) {
        __syncthreads();
        grad_cache_flush(*gcp, ra.d_params);
    }
}

constexpr auto kpf_fwd_words = k_render_pf<false, true>, kpf_fwd_inline = k_render_pf<false, false>;
constexpr auto kpf_bwd_words = k_render_pf<true, true>, kpf_bwd_inline = k_render_pf<true, false>;

static int pf_blocks(const BinView &bins, const RenderArgs &ra) {
    const int ns = bins.tile_w * bins.tile_h * ra.nsx * ra.nsy;
    const int parts = (ns + PB - 1) / PB;
    const int r0 = ra.row_begin / bins.tile_h;
    const int r1 = (ra.row_end + bins.tile_h - 1) / bins.tile_h;
    return (r1 - r0) * bins.tiles_x * parts;
}

// `wind` / `relevant` / `tile_choff`: the words of the winding pre-pass and the chunk offsets they are laid out by, or null
// (the winding test then runs inline: scenes without fills, renders beyond the 27-bit word index).  `pc`: fragment cache to
// fill (forward) or to skip the cached samples by (backward), or nulls.
void launch_render_pf(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const unsigned *wind, const unsigned *relevant,
                      const int *tile_choff, const PfCache &pc, bool backward, cudaStream_t st) {
    const int nblk = pf_blocks(bins, ra);
    if (nblk <= 0) return;
    const uint4 *w4 = reinterpret_cast<const uint4 *>(wind);
    if (backward) {
        if (wind) DVG_LAUNCH_AS("k_render_pf<true>", kpf_bwd_words, dim3(nblk), dim3(PB), 0, st, sc, bins, ra, w4, relevant, tile_choff, pc);
        else DVG_LAUNCH_AS("k_render_pf<true>", kpf_bwd_inline, dim3(nblk), dim3(PB), 0, st, sc, bins, ra, w4, relevant, tile_choff, pc);
    } else {
        if (wind) DVG_LAUNCH_AS("k_render_pf<false>", kpf_fwd_words, dim3(nblk), dim3(PB), 0, st, sc, bins, ra, w4, relevant, tile_choff, pc);
        else DVG_LAUNCH_AS("k_render_pf<false>", kpf_fwd_inline, dim3(nblk), dim3(PB), 0, st, sc, bins, ra, w4, relevant, tile_choff, pc);
    }
}
void launch_pf_backward_cached(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const PfCache &pc, cudaStream_t st) {
    const int nblk = pf_blocks(bins, ra);
    if (nblk <= 0) return;
    DVG_LAUNCH(k_pf_backward_cached, dim3(nblk), dim3(PB), 0, st, sc, bins, ra, pc);
}
int64_t pf_launch_threads(const BinView &bins, const RenderArgs &ra) { return (int64_t)pf_blocks(bins, ra) * PB; }

void launch_sdf(const SceneView &sc, const RenderArgs &ra, const SdfArgs &sa, bool backward, cudaStream_t st) {
    const int n = sa.eval_positions ? sa.num_eval : ra.width * ra.height * ra.nsx * ra.nsy;
    if (n <= 0) return;
    if (backward) DVG_LAUNCH(k_sdf<true>, dim3((n + PB - 1) / PB), dim3(PB), 0, st, sc, ra, sa);
    else DVG_LAUNCH(k_sdf<false>, dim3((n + PB - 1) / PB), dim3(PB), 0, st, sc, ra, sa);
}

}  // namespace dvg
