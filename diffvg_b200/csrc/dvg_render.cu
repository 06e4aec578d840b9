// dvg_render.cu -- the hot kernels: filter-weight splat, forward render, interior backward,
// and the boundary (edge-sampling) pass.
//
// Execution model (differs from the reference's one-thread-one-sample recursion with three
// int stacks and a 256-entry fragment array in 11 KB of local memory, diffvg.cpp:525-707):
//   * a thread block owns one pixel TILE; its candidate primitives come from the tile bin
//     (ascending primitive id == compositing order), are staged through shared memory in
//     chunks and walked by all threads in lockstep;
//   * each thread owns one sample and keeps only O(1) state: the running "over" composite,
//     the current group's stroke-hit flag and winding number.  The reference's fragment
//     array + insertion sort disappears because candidates already arrive sorted;
//   * the reference's three BVH levels collapse into flat per-leaf predicates (every inner
//     node test is implied by its leaf's test because boxes/radii are merged monotonically),
//     evaluated with exactly the reference's comparisons so classification is identical;
//   * pixel sums are accumulated in shared memory and flushed once per tile; colour
//     gradients are reduced across the warp before touching global memory.
#include "dvg_internal.h"
#include "dvg_kernel_util.cuh"

namespace dvg {

#ifndef DVG_RB
#define DVG_RB 256
#endif
#ifndef DVG_MINB
#define DVG_MINB 2
#endif
constexpr int RB = DVG_RB;   // threads per render block (warps are independent; small blocks keep the tail short)
constexpr int NWARP = RB / 32;
constexpr int MAXF = DVG_MAXF;
constexpr int EDGE_SPB = RB / 2;  // boundary samples per block (two lanes per sample)

// Per-warp scratch in shared memory for the re-packing of exact tests (see traverse()).
struct WarpScratch {
    unsigned int hit[32];            // [lane] bit k: stroke test of candidate k hit
    unsigned int wind[32][4];        // [lane] 4-bit signed winding contribution of candidate k
    unsigned short queue[32 * 32];   // work items (owner lane << 5 | candidate k)
};

// ------------------------------------------------------------------------------------------
// weight_kernel (diffvg.cpp:1115-1158).  Q1: always uses the jittered position, even when
// the render kernel uses pixel centres for prefiltering.
__global__ void k_weight(SceneView sc, RenderArgs ra, int idx_begin, int idx_end) {
    const int idx = idx_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= idx_end) return;
    Pcg32 rng = pcg32_init(idx, ra.seed);
    const int sx = idx % ra.nsx;
    const int sy = (idx / ra.nsx) % ra.nsy;
    const int x = (idx / (ra.nsx * ra.nsy)) % ra.width;
    const int y = idx / (ra.nsx * ra.nsy * ra.width);
    float rx = pcg32_next_float(rng);
    float ry = pcg32_next_float(rng);
    F2 pt = mk2(x + ((float)sx + rx) / ra.nsx, y + ((float)sy + ry) / ra.nsy);
    const int ri = (int)ceilf(sc.filter.radius);
    for (int dy = -ri; dy <= ri; dy++) {
        for (int dx = -ri; dx <= ri; dx++) {
            int xx = x + dx, yy = y + dy;
            if (xx >= 0 && xx < ra.width && yy >= 0 && yy < ra.height) {
                float w = filter_weight(sc.filter, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
                if (w != 0.f) atomicAdd(&ra.weight_image[yy * ra.width + xx], w);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tile traversal, one WARP at a time (no block-level barriers).  All 32 lanes (= 32 samples) walk
// the tile's candidate list (ascending primitive id == compositing order) in chunks of 32:
//   A  classify : lane k loads candidate k's leaf record; the records are broadcast with shuffles
//                 and every lane runs the cheap leaf tests of the reference's three BVH levels
//                 (SampleTracer<TM_CLASSIFY>) -> per-lane bit masks "needs exact stroke test" /
//                 "needs exact winding test".
//   D  solve    : the (sample, candidate) pairs that need an exact test are written to a
//                 candidate-major queue and handed out 32 at a time, so the quintic / cubic root
//                 solves (FP64, ~1-2 k instructions each) run with all lanes busy instead of the
//                 1-in-4 lane occupancy of one-thread-one-sample traversal.  Curved strokes first
//                 go through the conservative capsule early-out and the survivors are re-packed.
//   E  consume  : every lane replays the candidates in order with its results
//                 (SampleTracer<TM_CONSUME>): fragments, compositing, EdgeQuery bookkeeping.
// This reproduces sample_color(scene, ...) of diffvg.cpp:525-653.
DVG_D F2 local_point(const GroupInfo &g, F2 cpt) {
    return (g.flags & DVG_GF_IDENTITY) ? cpt : xform_pt(g.c2s, cpt);
}

template <bool EDGE, bool RECORD>
DVG_D void traverse(const SceneView &sc, const BinView &bins, const int tile, WarpScratch &ws,
                    SampleTracer<EDGE, RECORD> &tr, const bool fast_accept) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int beg = bins.offsets[tile], end = bins.offsets[tile + 1];
    const F2 cpt = tr.cpt;
    SampleTracer<false, false> ct;  // classification-only state
    ct.init(cpt, tr.active, mk4(0, 0, 0, 0), -1, -1, nullptr, nullptr);
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        // lane k holds candidate k
        int e = 0, tf = 0, inst = 0, group = 0;
        Box box; box.x0 = box.y0 = box.x1 = box.y1 = 0.f;
        float thick = 0.f;
        if (lane < n) {
            e = bins.items[base + lane];
            const PrimMeta pm = sc.prim_meta[e];
            tf = pm.type_flags; inst = pm.inst;
            group = sc.insts[inst].group;
            box = sc.prim_box[e];
            thick = sc.prim_thick[e];
        }
        // ---- A: classify
        unsigned need_s = 0, need_f = 0;
        for (int k = 0; k < n; k++) {
            PrimRef pr;
            pr.box.x0 = __shfl_sync(FULL, box.x0, k); pr.box.y0 = __shfl_sync(FULL, box.y0, k);
            pr.box.x1 = __shfl_sync(FULL, box.x1, k); pr.box.y1 = __shfl_sync(FULL, box.y1, k);
            pr.thick = __shfl_sync(FULL, thick, k);
            pr.tf = __shfl_sync(FULL, tf, k); pr.inst = __shfl_sync(FULL, inst, k); pr.group = __shfl_sync(FULL, group, k);
            const int nd = ct.template step<TM_CLASSIFY>(sc, pr);
            need_s |= (unsigned)(nd & 1) << k;
            need_f |= (unsigned)((nd >> 1) & 1) << k;
        }
        ws.hit[lane] = 0u;
        ws.wind[lane][0] = 0u; ws.wind[lane][1] = 0u; ws.wind[lane][2] = 0u; ws.wind[lane][3] = 0u;
        // ---- D (strokes, then fills)
#pragma unroll 1
        for (int kind = 0; kind < 2; kind++) {
            const unsigned need = kind == 0 ? need_s : need_f;
            if (!__any_sync(FULL, need != 0u)) continue;
            int qn = 0;
            for (int k = 0; k < n; k++) {  // candidate-major queue: lanes of a round mostly share the primitive
                const bool mine = (need >> k) & 1u;
                const unsigned m = __ballot_sync(FULL, mine);
                if (mine) ws.queue[qn + __popc(m & lt)] = (unsigned short)((lane << 5) | k);
                qn += __popc(m);
            }
            __syncwarp();
            if (kind == 0) {
                // D1: polyline bracket (dvg_scene.cuh): decided pairs are answered here, the rest re-packed in place
                int wn = 0;
                for (int r = 0; r < qn; r += 32) {
                    const bool have = r + lane < qn;
                    const int it = have ? ws.queue[r + lane] : 0;
                    const int k = it & 31, owner = it >> 5;
                    const int ek = __shfl_sync(FULL, e, k), tfk = __shfl_sync(FULL, tf, k), gk = __shfl_sync(FULL, group, k);
                    const F2 op = mk2(__shfl_sync(FULL, cpt.x, owner), __shfl_sync(FULL, cpt.y, owner));
                    bool keep = have;
#ifndef DVG_NO_CAPSULE
                    if (have) {
                        const int ptype = tfk & DVG_PF_TYPE_MASK;
                        if ((ptype == PRIM_CUBIC || ptype == PRIM_QUAD) && !(tfk & DVG_PF_APPROX)) {
                            const int cls = capsule_classify(reinterpret_cast<const float *>(sc.prim_cap + (size_t)ek * DVG_CAP_F4),
                                                             local_point(sc.groups[gk], op));
                            // cls > 0 ("certainly inside") is trusted only in the opt-in fast mode: the
                            // reference's solver has false negatives there (~6e-6 of the samples, Q21)
                            // which the default mode reproduces by running the exact solve
                            keep = cls == 0 || (cls > 0 && !fast_accept);
                            if (cls > 0 && fast_accept) atomicOr(&ws.hit[owner], 1u << k);
                        }
                    }
#endif
                    __syncwarp();
                    const unsigned m = __ballot_sync(FULL, keep);
                    if (keep) ws.queue[wn + __popc(m & lt)] = (unsigned short)it;
                    wn += __popc(m);
                    __syncwarp();
                }
                qn = wn;
            }
            // D2: exact tests, 32 per round
            for (int r = 0; r < qn; r += 32) {
                const bool have = r + lane < qn;
                const int it = have ? ws.queue[r + lane] : 0;
                const int k = it & 31, owner = it >> 5;
                const int ek = __shfl_sync(FULL, e, k), tfk = __shfl_sync(FULL, tf, k), gk = __shfl_sync(FULL, group, k);
                const int ik = __shfl_sync(FULL, inst, k);
                const F2 op = mk2(__shfl_sync(FULL, cpt.x, owner), __shfl_sync(FULL, cpt.y, owner));
                if (have) {
                    const F2 lp = local_point(sc.groups[gk], op);
                    const F4 p01 = sc.prim_p01[ek], p23 = sc.prim_p23[ek];
                    if (kind == 0) {
                        bool decided = false;
                        const bool h = prim_stroke_hit(tfk & DVG_PF_TYPE_MASK, (tfk & DVG_PF_APPROX) != 0, p01, p23, sc.prim_rad[ek],
                                                       sc.insts[ik].r, lp, &decided);
                        if (h) atomicOr(&ws.hit[owner], 1u << k);
                    } else {
                        const int w = prim_winding(tfk & DVG_PF_TYPE_MASK, p01, p23, lp);
                        if (w != 0) atomicOr(&ws.wind[owner][k >> 3], (unsigned)(w & 15) << (4 * (k & 7)));
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
        // ---- E: consume
        const unsigned hitm = ws.hit[lane];
        const unsigned w0 = ws.wind[lane][0], w1 = ws.wind[lane][1], w2 = ws.wind[lane][2], w3 = ws.wind[lane][3];
        for (int k = 0; k < n; k++) {
            PrimRef pr;
            pr.tf = __shfl_sync(FULL, tf, k); pr.inst = __shfl_sync(FULL, inst, k); pr.group = __shfl_sync(FULL, group, k);
            const int nd = (int)((need_s >> k) & 1u) | ((int)((need_f >> k) & 1u) << 1);
            const unsigned ww = (k < 8 ? w0 : (k < 16 ? w1 : (k < 24 ? w2 : w3)));
            const int nib = (int)((ww >> (4 * (k & 7))) & 15u);
            tr.template step<TM_CONSUME>(sc, pr, nd, ((hitm >> k) & 1u) != 0u, (nib ^ 8) - 8);
        }
        __syncwarp();
    }
    tr.finish(sc);
}

// ------------------------------------------------------------------------------------------
// render_kernel (diffvg.cpp:1161-1272), colour output.  BACKWARD = the d_render_image != null
// variant: recompute the forward, then d_sample_color (diffvg.cpp:656-705), d_background and
// the filter-radius gradient (1250-1268).
template <bool BACKWARD>
__global__ void __launch_bounds__(RB, DVG_MINB) k_render(SceneView sc, BinView bins, RenderArgs ra) {
    __shared__ WarpScratch s_ws[NWARP];
    GradCache *gcp = nullptr;
    if constexpr (BACKWARD) {
        __shared__ GradCache s_gc;
        gcp = &s_gc;
        grad_cache_init(s_gc);
        __syncthreads();
    }
    const CacheSink sk{gcp, ra.d_params};
    WarpScratch &ws = s_ws[threadIdx.x >> 5];
    const int tile_row0 = ra.row_begin / bins.tile_h;
    const int spp = ra.nsx * ra.nsy;
    const int npix = bins.tile_w * bins.tile_h;
    const int ns = npix * spp;
    const int parts = (ns + RB - 1) / RB;   // blocks per tile
    const int tile = blockIdx.x / parts + tile_row0 * bins.tiles_x;
    const int part = blockIdx.x % parts;
    const int tx = tile % bins.tiles_x, ty = tile / bins.tiles_x;
    const int tid = threadIdx.x;
    // samples of one pixel sit in `grp` consecutive lanes when spp is a power of two (<= 32, or a
    // multiple of 32): their splat onto the own pixel is pre-reduced with shuffles
    const bool pow2 = (spp & (spp - 1)) == 0;
    const int grp = pow2 ? (spp < 32 ? spp : 32) : 1;
    int fkey[BACKWARD ? MAXF : 1];
    F4 fprev[BACKWARD ? MAXF : 1];
    float d_radius_acc = 0.f;

    {
        const int l = part * RB + tid;
        const int s = l % spp, p = l / spp;
        const int px = p % bins.tile_w, py = p / bins.tile_w;
        const int x = tx * bins.tile_w + px, y = ty * bins.tile_h + py;
        const bool active = l < ns && x < ra.width && y < ra.height && y >= ra.row_begin && y < ra.row_end;
        struct { F2 pt, cpt; } pos;
        pos.pt = mk2(0, 0); pos.cpt = mk2(0, 0);
        const float *bg_px = nullptr;
        F4 first = mk4(0, 0, 0, 0);
        F4 d_color = mk4(0, 0, 0, 0);
        if (active) {
            const int sx = s % ra.nsx, sy = s / ra.nsx;
            const int idx = ((y * ra.width + x) * ra.nsy + sy) * ra.nsx + sx;
            sample_position(sc.canvas_w, sc.canvas_h, ra.width, ra.height, ra.nsx, ra.nsy, ra.seed,
                            ra.use_prefiltering != 0, x, y, sx, sy, idx, pos.pt, pos.cpt);
            if (ra.background) {
                bg_px = ra.background + 4 * (y * ra.width + x);
                first = mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]);
            }
            if (BACKWARD) d_color = gather_d_color(sc.filter, ra.d_render_image, ra.weight_image, ra.width, ra.height, pos.pt);
        }
        SampleTracer<false, BACKWARD> tr;
        tr.init(pos.cpt, active, first, -1, -1, fkey, fprev);
        traverse<false, BACKWARD>(sc, bins, tile, ws, tr, (ra.flags & DVG_RF_FAST_ACCEPT) != 0);
        const F4 color = tr.resolve(bg_px);
        struct { F4 accum; int nfrag, sp; } to;
        to.accum = tr.accum; to.nfrag = tr.nfrag; to.sp = tr.sp;

        if (!BACKWARD) {
            splat_color(sc, ra, x, y, pos.pt, color, active, grp, tid);
        } else {
            // ---- interior backward.  All 32 lanes stay converged: fragments are popped in
            // warp-uniform steps so that lanes sharing a (group, stroke/fill) key are reduced
            // with shuffles and scattered by ONE lane (atomic.h:23-51 does one global atomic
            // per component per sample).
            float dcr = d_color.x, dcg = d_color.y, dcb = d_color.z, dca = d_color.w;
            int sp = to.sp;
            if (to.nfrag > 0) {
                if (to.accum.w > 1e-6f) {
                    const float inv = 1.f / to.accum.w;
                    dca -= (d_color.x * color.x + d_color.y * color.y + d_color.z * color.z) / to.accum.w;
                    dcr = d_color.x * inv; dcg = d_color.y * inv; dcb = d_color.z * inv;
                }
            } else {
                sp = 0;
                if (active && bg_px && ra.d_background) {  // diffvg.cpp:598-600 (Q2: accumulated, not assigned)
                    float *d = ra.d_background + 4 * (y * ra.width + x);
                    atomicAdd(d + 0, d_color.x); atomicAdd(d + 1, d_color.y); atomicAdd(d + 2, d_color.z); atomicAdd(d + 3, d_color.w);
                }
            }
            const bool had_frags = sp > 0;
            while (true) {
                const int mykey = sp > 0 ? fkey[sp - 1] : -1;
                const unsigned m = __ballot_sync(0xffffffffu, mykey >= 0);
                if (!m) break;
                const int key = __shfl_sync(0xffffffffu, mykey, __ffs(m) - 1);
                const GroupInfo &g = sc.groups[key >> 1];
                const int ctype = (key & 1) ? g.stroke_type : g.fill_type;
                const int coff = (key & 1) ? g.stroke_off : g.fill_off;
                const int cstops = (key & 1) ? g.stroke_stops : g.fill_stops;
                F4 dc = mk4(0, 0, 0, 0);
                if (mykey == key) {
                    sp--;
                    const F4 prev = fprev[sp];
                    const F4 fc = eval_color(ctype, sc.params + coff, cstops, pos.cpt);
                    // diffvg.cpp:673-679
                    const float d_prev_alpha = dca * (1.f - fc.w);
                    float d_alpha_i = dca * (1.f - prev.w);
                    d_alpha_i += (dcr * (fc.x - prev.x) + dcg * (fc.y - prev.y)) + dcb * (fc.z - prev.z);
                    dc = mk4(dcr * fc.w, dcg * fc.w, dcb * fc.w, d_alpha_i);
                    dcr = dcr * (1 - fc.w); dcg = dcg * (1 - fc.w); dcb = dcb * (1 - fc.w);
                    dca = d_prev_alpha;
                    if (ctype != 0 && !(key & 1)) {
                        // gradient FILL colours: per-lane scatter (diffvg.cpp:382-499)
                        d_eval_gradient(ctype, sc.params + coff, cstops, pos.cpt, dc, sk, coff,
                                        ra.d_translation ? ra.d_translation + 2 * (y * ra.width + x) : nullptr);
                    }
                    // Q4: gradient STROKE colours have no gradient storage in the reference
                    // (scene.cpp:868,887) -> nothing is accumulated for them.
                }
                if (ctype == 0) {
                    dc.x = warp_sum(dc.x); dc.y = warp_sum(dc.y); dc.z = warp_sum(dc.z); dc.w = warp_sum(dc.w);
                    if ((tid & 31) == 0) {
                        sk.add(coff + 0, dc.x); sk.add(coff + 1, dc.y); sk.add(coff + 2, dc.z); sk.add(coff + 3, dc.w);
                    }
                }
            }
            if (active && had_frags && bg_px && ra.d_background) {  // diffvg.cpp:699-704
                float *d = ra.d_background + 4 * (y * ra.width + x);
                atomicAdd(d + 0, dcr); atomicAdd(d + 1, dcg); atomicAdd(d + 2, dcb); atomicAdd(d + 3, dca);
            }
            if (active) d_radius_acc += filter_radius_grad(sc, ra, x, y, pos.pt, color);
        }
    }
    if (BACKWARD) {
        d_radius_acc = warp_sum(d_radius_acc);
        if ((tid & 31) == 0) sk.add(sc.filter_radius_off, d_radius_acc);
        __syncthreads();
        grad_cache_flush(*gcp, ra.d_params);
    }
}

// ------------------------------------------------------------------------------------------
// Boundary pass, step 1: sample_boundary_kernel (diffvg.cpp:1325-1386) reduced to what the
// ordering step needs -- the destination tile of every boundary sample.  The 56-byte
// BoundarySample records + Morton keys + thrust::sort_by_key round trip (268 MB at 512^2 x
// 16 spp, diffvg.cpp:1560-1595) is replaced by a 4-byte key, a counting sort by tile, and
// re-deriving the sample from its index (same RNG stream) inside the edge kernel.
__global__ void k_boundary_keys(SceneView sc, BinView bins, RenderArgs ra, BoundaryWork bw) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bw.num_samples) return;
    BoundarySample bs;
    make_boundary_sample(sc, bw.sample_begin + k, ra.seed, bs);
    int key = -1;
    if (bs.inst >= 0) {
        const int bx = (int)(bs.pt.x * ra.width), by = (int)(bs.pt.y * ra.height);  // diffvg.cpp:1405-1409
        if (bx >= 0 && bx < ra.width && by >= 0 && by < ra.height) {
            key = (by / bins.tile_h) * bins.tiles_x + bx / bins.tile_w;
            atomicAdd(&bw.tile_counts[key], 1);
        }
    }
    bw.keys[k] = key;
    if (bw.samples) bw.samples[k] = bs;
}

__global__ void k_boundary_blocks(BoundaryWork bw, int ntiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles) bw.blk_counts[t] = (bw.tile_counts[t] + EDGE_SPB - 1) / EDGE_SPB;
}

__global__ void k_boundary_scatter(BoundaryWork bw) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bw.num_samples) return;
    const int key = bw.keys[k];
    if (key < 0) return;
    const int pos = bw.tile_offsets[key] + atomicAdd(&bw.tile_fill[key], 1);
    bw.sorted_idx[pos] = bw.sample_begin + k;
}

// Boundary pass, step 2: render_edge_kernel (diffvg.cpp:1388-1475).  One block per
// (tile, chunk of EDGE_SPB samples); lanes 2k / 2k+1 evaluate the two sides of sample k.
__global__ void __launch_bounds__(RB, DVG_MINB) k_edge(SceneView sc, BinView bins, RenderArgs ra, BoundaryWork bw) {
    __shared__ WarpScratch s_ws[NWARP];
    __shared__ GradCache s_gc;
    __shared__ int s_tile;
    const int ntiles = bins.tiles_x * bins.tiles_y;
    const int blk = blockIdx.x;
    if (blk >= bw.blk_offsets[ntiles]) return;  // uniform for the whole block
    grad_cache_init(s_gc);
    if (threadIdx.x == 0) {
        int lo = 0, hi = ntiles;  // largest t with blk_offsets[t] <= blk
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (bw.blk_offsets[mid] <= blk) lo = mid; else hi = mid;
        }
        s_tile = lo;
    }
    __syncthreads();
    const int tile = s_tile;
    const int chunk = blk - bw.blk_offsets[tile];
    const int k = chunk * EDGE_SPB + (threadIdx.x >> 1);
    const bool valid = k < bw.tile_counts[tile];
    const int side = threadIdx.x & 1;
    BoundarySample bs;
    bs.inst = -1; bs.pt = mk2(0, 0); bs.normal = mk2(0, 0);
    if (valid) make_boundary_sample(sc, bw.sorted_idx[bw.tile_offsets[tile] + k], ra.seed, bs);
    const bool active = valid && bs.inst >= 0;
    int q_group = -1, q_shape = -1;
    int bx = 0, by = 0;
    F2 cpt = mk2(0, 0);
    const float *bg_px = nullptr;
    F4 first = mk4(0, 0, 0, 0);
    if (active) {
        const InstInfo &ii = sc.insts[bs.inst];
        q_group = ii.group; q_shape = ii.shape;
        bx = (int)(bs.pt.x * ra.width); by = (int)(bs.pt.y * ra.height);
        const F2 off = 1e-4f * bs.normal;
        const F2 npt = side ? bs.pt + off : bs.pt - off;  // diffvg.cpp:1416,1420
        cpt = mk2(npt.x * sc.canvas_w, npt.y * sc.canvas_h);
        if (ra.background) {
            bg_px = ra.background + 4 * (by * ra.width + bx);
            first = mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]);
        }
    }
    SampleTracer<true, false> tr;
    tr.init(cpt, active, first, q_group, q_shape, nullptr, nullptr);
    traverse<true, false>(sc, bins, tile, s_ws[threadIdx.x >> 5], tr, (ra.flags & DVG_RF_FAST_ACCEPT) != 0);
    const F4 mine = tr.resolve(bg_px);
    const int my_hit = tr.q_hit() ? 1 : 0;
    F4 other;
    other.x = __shfl_xor_sync(0xffffffffu, mine.x, 1);
    other.y = __shfl_xor_sync(0xffffffffu, mine.y, 1);
    other.z = __shfl_xor_sync(0xffffffffu, mine.z, 1);
    other.w = __shfl_xor_sync(0xffffffffu, mine.w, 1);
    const int other_hit = __shfl_xor_sync(0xffffffffu, my_hit, 1);
    // lane `side == 0` evaluated pt - eps*n ("inside"); it owns the scatter of its sample.
    // occluded samples contribute nothing (diffvg.cpp:1422-1425)
    const bool scatter = active && side == 0 && (my_hit || other_hit);
    const CacheSink sk{&s_gc, ra.d_params};
    float dm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int xoff = -1;
    if (scatter) {
        F4 c_in = mine, c_out = other;
        F2 normal = bs.normal;
        if (!my_hit) { normal = -normal; c_in = other; c_out = mine; }
        const F2 spt = mk2(bs.pt.x * ra.width, bs.pt.y * ra.height);
        F4 d_color = gather_d_color(sc.filter, ra.d_render_image, ra.weight_image, ra.width, ra.height, spt);
        const float inv_area = 1.f / (float)(sc.canvas_w * sc.canvas_h);
        d_color = d_color * inv_area;
        const F4 diff = c_in - c_out;
        const float contrib = (diff.x * d_color.x + diff.y * d_color.y + diff.z * d_color.z + diff.w * d_color.w) / bs.pdf;
        const InstInfo &ii = sc.insts[bs.inst];
        const GroupInfo &g = sc.groups[ii.group];
        accumulate_boundary_gradient(sc, ra, bs, ii, g, contrib, normal, sk);
        if (ra.debug_out) {
            float *o = ra.debug_out + 4 * (size_t)bw.sorted_idx[bw.tile_offsets[tile] + k];
            o[0] = contrib; o[1] = (float)(my_hit | (other_hit << 1)); o[2] = normal.x; o[3] = normal.y;
        }
        if (!(ra.flags & 1u)) {  // DVG_BWD_SKIP_XFORM_GRAD
            boundary_xform_gradient(bs, g, contrib, normal, dm);
            xoff = g.xform_off;
        }
        if (ra.d_translation) {  // diffvg.cpp:1454-1461
            atomicAdd(ra.d_translation + 2 * (by * ra.width + bx) + 0, normal.x * contrib);
            atomicAdd(ra.d_translation + 2 * (by * ra.width + bx) + 1, normal.y * contrib);
        }
    }
    // d_shape_to_canvas: warp-reduce when every scattering lane targets the same transform
    {
        const unsigned am = __ballot_sync(0xffffffffu, xoff >= 0);
        if (am) {
            const int x0 = __shfl_sync(0xffffffffu, xoff, __ffs(am) - 1);
            const bool uniform = __all_sync(0xffffffffu, xoff < 0 || xoff == x0);
            if (uniform) {
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const float v = warp_sum(dm[k]);
                    if ((threadIdx.x & 31) == 0) sk.add(x0 + k, v);
                }
            } else if (xoff >= 0) {
#pragma unroll
                for (int k = 0; k < 9; k++) sk.add(xoff + k, dm[k]);
            }
        }
    }
    __syncthreads();
    grad_cache_flush(s_gc, ra.d_params);
}

// ------------------------------------------------------------------------------------------
int edge_samples_per_block() { return EDGE_SPB; }

// Weights of the pixels of rows [row_begin, row_end): the samples of those rows and of ceil(radius) rows either side
// splat onto them (a row shard does not need the rest of the image).
void launch_weight(const SceneView &sc, const RenderArgs &ra, int row_begin, int row_end, cudaStream_t st) {
    // pixels up to ri rows outside the band are read (gather_d_color, splat); their weights need samples ri rows further
    const int ri = 2 * (int)ceilf(sc.filter.radius);
    const int y0 = max(0, row_begin - ri), y1 = min(ra.height, row_end + ri);
    const int per_row = ra.width * ra.nsx * ra.nsy;
    const int n = (y1 - y0) * per_row;
    if (n <= 0) return;
    DVG_LAUNCH(k_weight, dim3((n + 255) / 256), dim3(256), 0, st, sc, ra, y0 * per_row, y1 * per_row);
}

static int parts_per_tile(const BinView &bins, const RenderArgs &ra) {
    const int ns = bins.tile_w * bins.tile_h * ra.nsx * ra.nsy;
    return (ns + RB - 1) / RB;
}

static int tile_rows_in(const BinView &bins, const RenderArgs &ra) {
    const int r0 = ra.row_begin / bins.tile_h;
    const int r1 = (ra.row_end + bins.tile_h - 1) / bins.tile_h;
    return r1 - r0;
}

void launch_render_forward(const SceneView &sc, const BinView &bins, const RenderArgs &ra, cudaStream_t st) {
    const int nblk = tile_rows_in(bins, ra) * bins.tiles_x * parts_per_tile(bins, ra);
    if (nblk <= 0) return;
    DVG_LAUNCH(k_render<false>, dim3(nblk), dim3(RB), 0, st, sc, bins, ra);
}

void launch_render_backward(const SceneView &sc, const BinView &bins, const RenderArgs &ra, cudaStream_t st) {
    const int nblk = tile_rows_in(bins, ra) * bins.tiles_x * parts_per_tile(bins, ra);
    if (nblk <= 0) return;
    DVG_LAUNCH(k_render<true>, dim3(nblk), dim3(RB), 0, st, sc, bins, ra);
}

// tile keys + counting sort of the boundary samples by tile (fills tile_counts, tile_offsets, sorted_idx)
void launch_boundary_sort(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw, cudaStream_t st) {
    const int ntiles = bins.tiles_x * bins.tiles_y;
    if (bw.num_samples <= 0) return;
    cudaMemsetAsync(bw.tile_counts, 0, sizeof(int) * ntiles, st);
    cudaMemsetAsync(bw.tile_fill, 0, sizeof(int) * ntiles, st);
    DVG_LAUNCH(k_boundary_keys, dim3((bw.num_samples + 255) / 256), dim3(256), 0, st, sc, bins, ra, bw);
    launch_scan(bw.tile_counts, bw.tile_offsets, ntiles, st);
    DVG_LAUNCH(k_boundary_scatter, dim3((bw.num_samples + 255) / 256), dim3(256), 0, st, bw);
}

void launch_boundary(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw, cudaStream_t st) {
    const int ntiles = bins.tiles_x * bins.tiles_y;
    if (bw.num_samples <= 0) return;
    cudaMemsetAsync(bw.tile_counts, 0, sizeof(int) * ntiles, st);
    cudaMemsetAsync(bw.tile_fill, 0, sizeof(int) * ntiles, st);
    DVG_LAUNCH(k_boundary_keys, dim3((bw.num_samples + 255) / 256), dim3(256), 0, st, sc, bins, ra, bw);
    DVG_LAUNCH(k_boundary_blocks, dim3((ntiles + 255) / 256), dim3(256), 0, st, bw, ntiles);
    launch_scan(bw.tile_counts, bw.tile_offsets, ntiles, st);
    launch_scan(bw.blk_counts, bw.blk_offsets, ntiles, st);
    DVG_LAUNCH(k_boundary_scatter, dim3((bw.num_samples + 255) / 256), dim3(256), 0, st, bw);
    DVG_LAUNCH(k_edge, dim3(bw.max_blocks), dim3(RB), 0, st, sc, bins, ra, bw);
}

}  // namespace dvg
