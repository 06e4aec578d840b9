// dvg_wave.cu -- the render passes as a WAVEFRONT of small kernels.
//
// One fused kernel per pass (classify -> exact root solves -> composite in one warp program, dvg_render.cu)
// measured on B200: 128 registers (25% occupancy), 13-15 k SASS instructions (stall_no_inst 27%: the
// warps of an SM sit in different phases and thrash the instruction cache), 9-14 of 32 lanes active in the
// FP64 root solves, 19% of warp samples waiting at the block barrier before the gradient flush.
// None of that is compulsory: HBM is idle (< 1% of 6.4 TB/s), so the phases are separated and talk through
// global memory instead:
//
//   W1 classify   one warp per ITEM (32 evaluations = pixel samples, or boundary-sample sides, on one
//                 tile).  Walks the tile's candidate list in chunks of 32, runs the flat leaf predicates of
//                 the reference's three BVH levels and the polyline bracket, and APPENDS every
//                 (evaluation, candidate) pair that needs an exact test to a global queue (16 B records,
//                 warp-aggregated atomic append, coalesced).  Writes one `hit` word per (evaluation, chunk).
//   W2 solve      one THREAD per queued pair: the exact stroke test (within_distance.h) or winding
//                 contribution (winding_number.h); a hit is OR-ed into the evaluation's word.  Every lane
//                 has work, whatever the mix of decided / undecided samples in a tile.
//   W3 composite  one warp per item again: replays the candidate list with the result words (fragments,
//                 "over" compositing, EdgeQuery), then splat (forward), d_sample_color (interior backward)
//                 or the Reynolds term (boundary pass).
//
// Traffic at the painterly config, forward: 4.4 M words written + read twice, ~6 M pairs x 16 B written +
// read: < 200 MB per pass, ~30 us at the measured HBM rate.  Arithmetic is the same functions as before
// (dvg_geom.cuh / dvg_trace.cuh), so results are unchanged.
#include "dvg_internal.h"
#include "dvg_kernel_util.cuh"

#include <algorithm>

namespace dvg {

int g_num_sms = 148;   // set from the device properties at scene creation (dvg_capi.cu)

#ifndef DVG_WB_MIN
#define DVG_WB_MIN 4
#endif
// minimum resident blocks per SM of the classify kernels (64 registers); the composite kernels take one more block
// (48 registers): measured 4 / 5 / 6 blocks -> 6.91 / 6.82 / 7.02 ms per step when applied to all of them
constexpr int WB = 256;            // threads per block of the per-item kernels (8 items)
constexpr int WNW = WB / 32;
constexpr int W_EDGE_SPI = 16;     // boundary samples per item (two lanes per sample)
constexpr int W_MAXF = DVG_MAXF;

struct WaveScratch {
    unsigned int hit[32];            // [lane] bit k: candidate k answered "hit" without an exact test
    unsigned short queue[32 * 32];   // (owner lane << 5 | candidate k)
};

DVG_D F2 w_local_point(const GroupInfo &g, F2 cpt) {
    return (g.flags & DVG_GF_IDENTITY) ? cpt : xform_pt(g.c2s, cpt);
}

// Append `count` records of this warp to a global queue: one atomic per warp, returns the base index.
DVG_D int warp_reserve(int *counter, int count) {
    int base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, count);
    return __shfl_sync(0xffffffffu, base, 0);
}

// Gradient scatter.  The fused kernels pre-reduce in a shared-memory hash (GradCache) that needs a block
// barrier before its flush and, having no native shared-memory float add, spins on CAS: 30% of the boundary
// composite.  Here lanes that target the same segment are summed with shuffles and the leader issues one
// fire-and-forget `red.global.add.f32` per address (atomic.h:23-51 does one atomic per component per sample).
__global__ void k_wave_reduce_grads(const float *rep, int reps, int n, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int r = 0; r < reps; r++) s += rep[(size_t)r * n + i];
    if (s != 0.f) out[i] += s;
}

// Shape gradients of one boundary sample: every lane issues its own fire-and-forget `red.global.add.f32` into the
// block's gradient replica.  Measured against summing the lanes of a warp that target the same segment with shuffles
// first (4x fewer atomics): 0.95 vs 1.34 ms for the boundary composite -- the 32 replicas keep the L2 atomic units
// far from saturation, while the shuffle chains were a third of the kernel.
DVG_D void scatter_record(const GradRec &gr, float *D) {
    if (gr.key < 0) return;
#pragma unroll
    for (int j = 0; j < DVG_GREC_N; j++)
        if (gr.addr[j] >= 0 && gr.val[j] != 0.f) atomicAdd(D + gr.addr[j], gr.val[j]);
}

// Cold path of W1: the pair queue is full, so the exact test runs in the classifying lane (same functions as W2).
__device__ __noinline__ void wave_exact_in_place(const SceneView &sc, const WaveView &wv, int kind, int e, int tf, int inst, F2 lp,
                                                 unsigned *hit_word, int64_t word, int k) {
    const int ptype = tf & DVG_PF_TYPE_MASK;
    if (kind == 0) {
        bool decided = false;
        if (prim_stroke_hit(ptype, (tf & DVG_PF_APPROX) != 0, sc.prim_p01[e], sc.prim_p23[e], sc.prim_rad[e], sc.insts[inst].r, lp, &decided))
            atomicOr(hit_word, 1u << k);
    } else {
        const int w = prim_winding(ptype, sc.prim_p01[e], sc.prim_p23[e], lp);
        if (w != 0) atomicOr(&wv.wind[(size_t)word * 4 + (k >> 3)], (unsigned)(w & 15) << (4 * (k & 7)));
    }
}

// ------------------------------------------------------------------------------------------ W1
// Classification of one item.  `cb` = first chunk slot of the item; the word of (lane, chunk c) is
// (cb + c) * 32 + lane.  Reproduces the tests of sample_color's traversal (diffvg.cpp:544-594,
// within_distance.h:278-285, 362-388, winding_number.h:162-169) up to, not including, the exact
// per-segment tests.
// INPLACE: the retry form.  It runs after W1 + W2 and only does anything when a pair queue overflowed (a pass whose
// geometry asked for more exact tests than any pass before it): every pair is then answered in the classifying lane and
// the result words of the whole pass are rewritten.  Keeping this out of the hot form keeps that one free of calls
// (the call alone cost it 300 bytes of spills and 60% of its speed).
// FILLS: the scene has filled groups.  The winding test of a y-monotone cubic (DVG_PF_YMONO) whose control points all lie
// to the right of the sample is answered here from the primitive's PrimWindCert record (dvg_geom.cuh wind_cert_answer:
// a dozen FP64 operations, warp-coherent because all lanes test the same segment) instead of being queued.
// PF: the winding pre-pass of the prefiltered path (sample_color_prefiltered, diffvg.cpp:835-1113).  Only the winding
// numbers are taken from here -- the gating tests are those of the sampled path (diffvg.cpp:42-45, 71-78,
// winding_number.h:162-169) -- and the `hit` word of a chunk carries, the same in every lane, the candidates that matter
// to SOME sample of the item: those with a winding test, and those a closest-point search can reach (rects; path segments
// whose leaf box is within the search radius -- 1 for a fill-only group, unbounded for a stroked one: compute_distance.h:
// 285-294 prunes against a running minimum that starts there).  k_render_pf walks only those.
template <bool INPLACE, bool FILLS, bool PF = false>
DVG_D void wave_classify(const SceneView &sc, const BinView &bins, const WaveView &wv, int tile, int64_t cb,
                         F2 cpt, bool active, WaveScratch &ws, bool fast_accept) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int beg = bins.offsets[tile], end = bins.offsets[tile + 1];
    SampleTracer<false, false> ct;
    ct.init(cpt, active, mk4(0, 0, 0, 0), -1, -1, nullptr, nullptr);
    int c = 0;
    for (int base = beg; base < end; base += 32, c++) {
        const int n = min(32, end - base);
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
        unsigned need_s = 0, need_f = 0, relevant = 0;
        unsigned wn0 = 0u, wn1 = 0u, wn2 = 0u, wn3 = 0u;   // winding nibbles answered here
        for (int k = 0; k < n; k++) {
            PrimRef pr;
            pr.box.x0 = __shfl_sync(FULL, box.x0, k); pr.box.y0 = __shfl_sync(FULL, box.y0, k);
            pr.box.x1 = __shfl_sync(FULL, box.x1, k); pr.box.y1 = __shfl_sync(FULL, box.y1, k);
            pr.thick = __shfl_sync(FULL, thick, k);
            pr.tf = __shfl_sync(FULL, tf, k); pr.inst = __shfl_sync(FULL, inst, k); pr.group = __shfl_sync(FULL, group, k);
            const int nd = ct.template step<TM_CLASSIFY>(sc, pr);
            if (!PF) need_s |= (unsigned)(nd & 1) << k;
            bool nf = ((nd >> 1) & 1) != 0;
            if (PF) {
                const int type = pr.tf & DVG_PF_TYPE_MASK;
                const bool reach = type == PRIM_RECT ||
                    (type <= PRIM_CUBIC && ((pr.tf & DVG_PF_SINGLE) || box_within_distance(pr.box, ct.lpt, ct.has_stroke ? INFINITY : 1.f)));
                relevant |= (unsigned)((ct.g_visit && (reach || nf)) ? 1u : 0u) << k;
            }
            if (FILLS && (pr.tf & DVG_PF_YMONO)) {   // (uniform: every lane holds the same candidate)
                const int ek = __shfl_sync(FULL, e, k);
                int w = 0;
                if (nf && wind_cert_answer(sc.prim_wcert[ek], (pr.tf & DVG_PF_YUP) != 0, pr.box.x0, ct.lpt, &w)) {
                    nf = false;
                    const unsigned bits = (unsigned)(w & 15) << (4 * (k & 7));
                    if (k < 8) wn0 |= bits; else if (k < 16) wn1 |= bits; else if (k < 24) wn2 |= bits; else wn3 |= bits;
                }
            }
            need_f |= (unsigned)(nf ? 1u : 0u) << k;
        }
        ws.hit[lane] = 0u;
        const int64_t word0 = (cb + c) * 32;
        if (FILLS) {   // written first: the exact tests (and a pair that finds its queue full and is answered in place) OR into these words
            uint4 z; z.x = wn0; z.y = wn1; z.z = wn2; z.w = wn3;
            reinterpret_cast<uint4 *>(wv.wind)[word0 + lane] = z;
        }
        __syncwarp();
#pragma unroll 1
        for (int kind = 0; kind < 2; kind++) {
            const unsigned need = kind == 0 ? need_s : need_f;
            if (!__any_sync(FULL, need != 0u)) continue;
            int qn = 0;
            for (int k = 0; k < n; k++) {  // candidate-major queue
                const bool mine = (need >> k) & 1u;
                const unsigned m = __ballot_sync(FULL, mine);
                if (mine) ws.queue[qn + __popc(m & lt)] = (unsigned short)((lane << 5) | k);
                qn += __popc(m);
            }
            __syncwarp();
            WavePair *out = kind == 0 ? wv.pairs_s : wv.pairs_f;
            const int cap = kind == 0 ? wv.cap_s : wv.cap_f;
            // winding pairs are all queued: ONE reservation for the chunk's qn of them (a fill-heavy scene queues hundreds
            // per chunk; a reservation per 32 was 6 M atomics on one address per pass at tiger.svg, two thirds of this kernel)
            int fill_base = 0, kq = 0;
            if (!INPLACE && kind == 1) fill_base = warp_reserve(&wv.counters[1], qn);
            for (int r = 0; r < qn; r += 32) {
                const bool have = r + lane < qn;
                const int it = have ? ws.queue[r + lane] : 0;
                const int k = it & 31, owner = it >> 5;
                const int ek = __shfl_sync(FULL, e, k), tfk = __shfl_sync(FULL, tf, k), gk = __shfl_sync(FULL, group, k);
                int ik = 0;
                if (INPLACE) ik = __shfl_sync(FULL, inst, k);
                const F2 op = mk2(__shfl_sync(FULL, cpt.x, owner), __shfl_sync(FULL, cpt.y, owner));
                bool keep = have;
                F2 lp = mk2(0, 0);
                if (have) {
                    lp = w_local_point(sc.groups[gk], op);
                    const int ptype = tfk & DVG_PF_TYPE_MASK;
                    if (kind == 0 && (ptype == PRIM_CUBIC || ptype == PRIM_QUAD) && !(tfk & DVG_PF_APPROX)) {
                        // polyline bracket (dvg_scene.cuh): "certainly outside" is exact; "certainly inside" is
                        // trusted only in the opt-in fast mode (the reference's solver has false negatives, Q21)
                        const int cls = capsule_classify(reinterpret_cast<const float *>(sc.prim_cap + (size_t)ek * DVG_CAP_F4), lp);
                        keep = cls == 0 || (cls > 0 && !fast_accept);
                        if (cls > 0 && fast_accept) atomicOr(&ws.hit[owner], 1u << k);
                    }
                }
                if (INPLACE) {
                    if (keep) wave_exact_in_place(sc, wv, kind, ek, tfk, ik, lp, &ws.hit[owner], word0 + owner, k);
                    continue;
                }
                if (kind == 1) {
                    const int pos = fill_base + r + lane;
                    if (keep && pos < cap) {
                        WavePair p;
                        p.x = lp.x; p.y = lp.y; p.prim = ek | ((tfk & DVG_PF_TYPE_MASK) << 28);   // type rides along: W2 needs no meta load
                        p.ref = ((unsigned)(word0 + owner) << 5) | (unsigned)k;
                        out[pos] = p;
                    }
                    continue;
                }
                // stroke pairs: the survivors of the bracket are compacted in place (kq <= r: only entries already read are
                // overwritten) and queued after the loop with ONE reservation for the chunk
                const unsigned m = __ballot_sync(FULL, keep);
                if (keep) ws.queue[kq + __popc(m & lt)] = (unsigned short)it;
                kq += __popc(m);
            }
            __syncwarp();
            if (!INPLACE && kind == 0 && kq > 0) {
                // the counter keeps counting past the capacity: that is how the retry form and the host learn of it
                const int base_pos = warp_reserve(&wv.counters[0], kq);
                for (int r = 0; r < kq; r += 32) {
                    const bool have = r + lane < kq;
                    const int it = have ? ws.queue[r + lane] : 0;
                    const int k = it & 31, owner = it >> 5;
                    const int ek = __shfl_sync(FULL, e, k), tfk = __shfl_sync(FULL, tf, k), gk = __shfl_sync(FULL, group, k);
                    const F2 op = mk2(__shfl_sync(FULL, cpt.x, owner), __shfl_sync(FULL, cpt.y, owner));
                    const int pos = base_pos + r + lane;
                    if (have && pos < cap) {
                        const F2 lp = w_local_point(sc.groups[gk], op);
                        WavePair p;
                        p.x = lp.x; p.y = lp.y; p.prim = ek | ((tfk & DVG_PF_TYPE_MASK) << 28);
                        p.ref = ((unsigned)(word0 + owner) << 5) | (unsigned)k;
                        out[pos] = p;
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
        wv.hit[word0 + lane] = PF ? __reduce_or_sync(FULL, relevant) : ws.hit[lane];
        __syncwarp();
    }
}

// Geometry of a pixel item: item -> (scene, tile), lane -> sample.  `tile` is the batch-wide tile (the index of its bin
// and of its chunk offsets); x, y, idx are the scene's own pixel and sample index (RNG streams are per scene).
struct PixelItem {
    int tile, scene, x, y, sx, sy, idx;
    bool active;
    int64_t cb;
};
DVG_D PixelItem pixel_item(const BinView &bins, const RenderArgs &ra, const WaveView &wv, int item) {
    PixelItem pi;
    const int spp = ra.nsx * ra.nsy;
    const int ns = bins.tile_w * bins.tile_h * spp;
    const int wpt = (ns + 31) / 32;
    const int tile_row0 = ra.row_begin / bins.tile_h;
    pi.tile = item / wpt + tile_row0 * bins.tiles_x;
    const int part = item % wpt;
    const int tiles_scene = bin_scene_tiles(bins);
    pi.scene = pi.tile / tiles_scene;
    const int ltile = pi.tile - pi.scene * tiles_scene;
    const int tx = ltile % bins.tiles_x, ty = ltile / bins.tiles_x;
    const int l = part * 32 + (threadIdx.x & 31);
    const int s = l % spp, p = l / spp;
    pi.x = tx * bins.tile_w + p % bins.tile_w;
    pi.y = ty * bins.tile_h + p / bins.tile_w;
    pi.sx = s % ra.nsx; pi.sy = s / ra.nsx;
    pi.active = l < ns && pi.x < ra.width && pi.y < ra.height && pi.y >= ra.row_begin && pi.y < ra.row_end;
    pi.idx = ((pi.y * ra.width + pi.x) * ra.nsy + pi.sy) * ra.nsx + pi.sx;
    const int c0 = wv.tile_choff[pi.tile];
    const int nch = wv.tile_choff[pi.tile + 1] - c0;
    pi.cb = (int64_t)c0 * wpt + (int64_t)part * nch;
    return pi;
}

DVG_D bool wave_overflowed(const WaveView &wv) { return wv.counters[0] > wv.cap_s || wv.counters[1] > wv.cap_f; }

template <bool INPLACE, bool FILLS, bool PF = false>
__global__ void __launch_bounds__(WB, DVG_WB_MIN) k_wave_classify_px(SceneView sc, BinView bins, RenderArgs ra, WaveView wv, int num_items) {
    __shared__ WaveScratch s_ws[WNW];
    if (INPLACE && !wave_overflowed(wv)) return;
    for (int item = blockIdx.x * WNW + (threadIdx.x >> 5); item < num_items; item += gridDim.x * WNW) {
        const PixelItem pi = pixel_item(bins, ra, wv, item);
        F2 pt = mk2(0, 0), cpt = mk2(0, 0);
        if (pi.active)
            sample_position(sc.canvas_w, sc.canvas_h, ra.width, ra.height, ra.nsx, ra.nsy, ra.seeds ? ra.seeds[pi.scene] : ra.seed,
                            ra.use_prefiltering != 0, pi.x, pi.y, pi.sx, pi.sy, pi.idx, pt, cpt);
        wave_classify<INPLACE, FILLS, PF>(sc, bins, wv, pi.tile, pi.cb, cpt, pi.active, s_ws[threadIdx.x >> 5], (ra.flags & DVG_RF_FAST_ACCEPT) != 0);
    }
}

// Geometry of a boundary item: 16 boundary samples of one tile; lanes 2k / 2k+1 = the two sides of sample k.
struct EdgeItem {
    int tile, k;
    bool valid;
    int64_t cb;
};
DVG_D EdgeItem edge_item(const BinView &bins, const BoundaryWork &bw, const WaveView &wv, int item) {
    EdgeItem ei;
    const int lo = bw.item_tile[item];
    ei.tile = lo;
    const int chunk = item - bw.blk_offsets[lo];
    ei.k = chunk * W_EDGE_SPI + ((threadIdx.x & 31) >> 1);
    ei.valid = ei.k < bw.tile_counts[lo];
    const int nch = wv.tile_choff[lo + 1] - wv.tile_choff[lo];
    ei.cb = (int64_t)wv.edge_choff[lo] + (int64_t)chunk * nch;
    return ei;
}

struct EdgeLane {
    BoundarySample bs;
    bool active;
    F2 cpt;
    int bx, by;
};
DVG_D EdgeLane edge_lane(const SceneView &sc, const RenderArgs &ra, const BoundaryWork &bw, const EdgeItem &ei) {
    EdgeLane el;
    el.bs.inst = -1; el.bs.pt = mk2(0, 0); el.bs.normal = mk2(0, 0);
    if (ei.valid) el.bs = bw.samples[bw.tile_offsets[ei.tile] + ei.k];   // made by k_boundary_keys, moved into tile order by k_boundary_scatter
    el.active = ei.valid && el.bs.inst >= 0;
    el.cpt = mk2(0, 0); el.bx = el.by = 0;
    if (el.active) {
        el.bx = (int)(el.bs.pt.x * ra.width); el.by = (int)(el.bs.pt.y * ra.height);
        const F2 off = 1e-4f * el.bs.normal;
        const F2 npt = (threadIdx.x & 1) ? el.bs.pt + off : el.bs.pt - off;  // diffvg.cpp:1416,1420
        el.cpt = mk2(npt.x * sc.canvas_w, npt.y * sc.canvas_h);
    }
    return el;
}

template <bool INPLACE, bool FILLS>
__global__ void __launch_bounds__(WB, DVG_WB_MIN) k_wave_classify_edge(SceneView sc, BinView bins, RenderArgs ra, BoundaryWork bw, WaveView wv) {
    __shared__ WaveScratch s_ws[WNW];
    if (INPLACE && !wave_overflowed(wv)) return;
    const int ntiles = bin_total_tiles(bins);
    const int num_items = bw.blk_offsets[ntiles];
    for (int item = blockIdx.x * WNW + (threadIdx.x >> 5); item < num_items; item += gridDim.x * WNW) {
        const EdgeItem ei = edge_item(bins, bw, wv, item);
        const EdgeLane el = edge_lane(sc, ra, bw, ei);
        wave_classify<INPLACE, FILLS>(sc, bins, wv, ei.tile, ei.cb, el.cpt, el.active, s_ws[threadIdx.x >> 5], (ra.flags & DVG_RF_FAST_ACCEPT) != 0);
    }
}

// (the launch macro takes one token per argument: names for the instantiations)
constexpr auto kc_px = k_wave_classify_px<false, false>, kc_px_fills = k_wave_classify_px<false, true>;
constexpr auto kr_px = k_wave_classify_px<true, false>, kr_px_fills = k_wave_classify_px<true, true>;
constexpr auto kc_px_pf = k_wave_classify_px<false, true, true>, kr_px_pf = k_wave_classify_px<true, true, true>;
constexpr auto kc_edge = k_wave_classify_edge<false, false>, kc_edge_fills = k_wave_classify_edge<false, true>;
constexpr auto kr_edge = k_wave_classify_edge<true, false>, kr_edge_fills = k_wave_classify_edge<true, true>;

// ------------------------------------------------------------------------------------------ W2
// Exact stroke tests of the queued pairs (within_distance.h:119-272 for cubic segments).  One WARP takes 32 * SV_PPL pairs at a
// time through two phases that both keep its lanes busy and talk through the warp's slice of shared memory only (no
// block-level barrier: the warps of a block drift apart freely):
//   A (lane = pair, SV_PPL times)  end-point checks; the monic quintic of the stationary points of the squared distance -- its
//        sample-independent part comes from the primitive's PrimQuintic record (dvg_geom.cuh), the sample adds three
//        float dot products; the isolator split points; the sign tests of ALL brackets.  Which brackets hold a root, and
//        where each starts (`lower` only advances past a bracket that held one, :233-271), is a function of those signs
//        alone, so the brackets of a pair are independent UNITS; the pair's answer is the OR of their radius tests (the
//        reference's early return only skips work).  Other segment types are answered here.
//   B (lane = unit)         the warp's units, compacted (ascending brackets from the front of the list, descending ones
//        from its back): the reference's safeguarded Newton on one bracket (<= 20 evaluations), then the radius test at
//        the root found.  After the reference's swap a descending bracket has lb > ub, its "t in [lb, ub]" guard never
//        holds and it bisects (~3x the trips, no derivative needed): keeping the two kinds apart keeps the trip counts
//        of a round alike.
// Earlier forms: (1) A and B as separate kernels with the units in a global queue -- B re-derived the quintic from a random
// 16-byte pair gather plus three 16-byte primitive gathers (L2 hit rate 12-27%, 0.9 GB of DRAM reads per step) and A
// serialised a per-lane append: 2.58 ms per step at the painterly config; (2) one kernel, one block round = 256 pairs with
// a block-wide compaction: 2.12 ms, a fifth of it at the two barriers.
// Per-bracket arithmetic is dvg_geom.cuh's, evaluated on the same inputs: results are unchanged.
constexpr int SV_B = 256;                 // threads per block
constexpr int SV_NW = SV_B / 32;
#ifndef DVG_SOLVE_PPL
#define DVG_SOLVE_PPL 1      // measured (tools/sweep.sh): 1 pair per lane and round at 4 blocks per SM 1.50 ms, 2 at 3 blocks 1.68 ms
#endif
constexpr int SV_PPL = DVG_SOLVE_PPL;     // pairs per lane and round
constexpr int SV_PAIRS = 32 * SV_PPL;     // pairs per warp and round
constexpr int SV_MAXU = SV_PAIRS * 5;     // a quintic has at most five brackets
#ifndef DVG_SOLVE_PER_SM
#define DVG_SOLVE_PER_SM 64
#endif
#ifndef DVG_SOLVE_MINB
#define DVG_SOLVE_MINB 4
#endif

struct SolveWarp {
    double qD[SV_PAIRS], qE[SV_PAIRS], qF[SV_PAIRS];   // (B and C are the primitive's: read from its record)
    float px[SV_PAIRS], py[SV_PAIRS];
    int prim[SV_PAIRS];
    float ulb[SV_MAXU], uub[SV_MAXU];
    unsigned short upair[SV_MAXU];
    unsigned hit[SV_PPL];                 // bit l of word h: pair h * 32 + l answered "hit"
};

// The number of pairs is read from the device counter (nothing is read back to size a launch): the grid is a fixed
// multiple of the SM count and every warp strides over the queue.
__global__ void __launch_bounds__(SV_B, DVG_SOLVE_MINB) k_wave_stroke_solve(SceneView sc, WaveView wv) {
    __shared__ SolveWarp s_sw[SV_NW];
    const unsigned FULL = 0xffffffffu;
    const int count = min(wv.counters[0], wv.cap_s);
    const int lane = threadIdx.x & 31;
        SolveWarp &sw = s_sw[threadIdx.x >> 5];
    const int gwarp = blockIdx.x * SV_NW + (threadIdx.x >> 5), nwarps = gridDim.x * SV_NW;
    for (int base = gwarp * SV_PAIRS; base < count; base += nwarps * SV_PAIRS) {
        // ---- phase A
        int na = 0, nd = 0;               // units so far: ascending from the front, descending from the back
        unsigned refs[SV_PPL];
#pragma unroll 1
        for (int h = 0; h < SV_PPL; h++) {
            const int i = base + h * 32 + lane, slot = h * 32 + lane;
            float lbs[5], ubs[5];
            unsigned valid = 0u, desc = 0u;   // bit j: bracket j holds a root / is descending
            bool hit = false;
            unsigned ref = 0u;
            if (i < count) {
                const WavePair p = wv.pairs_s[i];
                const int ptype = (int)((unsigned)p.prim >> 28);
                const int e = p.prim & 0x0fffffff;
                ref = p.ref;
                const F2 pt = mk2(p.x, p.y);
                if (ptype != PRIM_CUBIC) {
                    const PrimMeta pm = sc.prim_meta[e];
                    bool decided = false;
                    hit = prim_stroke_hit_nocubic(ptype, (pm.type_flags & DVG_PF_APPROX) != 0, sc.prim_p01[e], sc.prim_p23[e], sc.prim_rad[e],
                                                  sc.insts[pm.inst].r, pt, &decided);
                } else {
                    const F4 p01 = sc.prim_p01[e], p23 = sc.prim_p23[e], rad = sc.prim_rad[e];
                    const F2 p0 = mk2(p01.x, p01.y), p3 = mk2(p23.z, p23.w);
                    if (dist_sq(p0, pt) < rad.x * rad.x || dist_sq(p3, pt) < rad.w * rad.w) {
                        hit = true;
                    } else {
                        const PrimQuintic k = sc.prim_quint[e];
                        const Quintic q = quintic_of(k, p0, pt);
                        float iv[4];
                        const int n = quintic_intervals_of(k, q, iv);
                        float lower = 0.f;
                        double f_lower = quintic_eval(q, lower);
                        bool open = true;
#pragma unroll
                        for (int j = 0; j < 5; j++) {
                            lbs[j] = 0.f; ubs[j] = 0.f;
                            const float ivj = iv[j < 4 ? j : 3];
                            if (open && j < n + 1 && !(j < n && ivj < 0.f)) {
                                const float upper = j < n ? rminf(ivj, 1.f) : 1.f;
                                const double f_upper = quintic_eval(q, upper);
                                if (!(f_lower * f_upper > 0)) {                 // :238 (a NaN product counts as a bracket)
                                    const bool d = f_lower > f_upper;         // :239-242
                                    lbs[j] = d ? upper : lower; ubs[j] = d ? lower : upper;
                                    valid |= 1u << j;
                                    if (d) desc |= 1u << j;
                                    if (upper >= 1.f) open = false;            // :268
                                    lower = upper; f_lower = f_upper;
                                }
                            }
                        }
                        sw.qD[slot] = q.D; sw.qE[slot] = q.E; sw.qF[slot] = q.F;
                        sw.px[slot] = p.x; sw.py[slot] = p.y; sw.prim[slot] = e;
                    }
                }
            }
            refs[h] = ref;
            const unsigned hm = __ballot_sync(FULL, hit);
            if (lane == 0) sw.hit[h] = hm;
            // compaction of this half's units (counts packed 16 + 16 bits)
            const unsigned mine = (unsigned)__popc(valid & ~desc) | ((unsigned)__popc(valid & desc) << 16);
            unsigned incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += u;
            }
            const unsigned total = __shfl_sync(FULL, incl, 31), excl = incl - mine;
            int pa = na + (int)(excl & 0xffffu), pd = SV_MAXU - 1 - (nd + (int)(excl >> 16));
#pragma unroll
            for (int j = 0; j < 5; j++) {
                if (!((valid >> j) & 1u)) continue;
                const int pos = ((desc >> j) & 1u) ? pd-- : pa++;
                sw.ulb[pos] = lbs[j]; sw.uub[pos] = ubs[j]; sw.upair[pos] = (unsigned short)slot;
            }
            na += (int)(total & 0xffffu); nd += (int)(total >> 16);
        }
        __syncwarp();
        // ---- phase B: ascending units in rounds of 32, then the descending ones
#pragma unroll 1
        for (int rr0 = 0; rr0 < ((na + 31) & ~31) + nd; rr0 += 32) {
            const bool descending = rr0 >= ((na + 31) & ~31);   // (a round never mixes the two kinds)
            const int k = (descending ? rr0 - ((na + 31) & ~31) : rr0) + lane;
            bool have = k < (descending ? nd : na);
            const int u = descending ? SV_MAXU - 1 - k : k;
            int pr = 0;
            if (have) {
                pr = sw.upair[u];
                have = !((*(volatile unsigned *)&sw.hit[pr >> 5] >> (pr & 31)) & 1u);   // another bracket of the pair already answered "hit"
            }
            bool pass = false;
            if (have) {
                const int e = sw.prim[pr];
                Quintic q;
                q.B = sc.prim_quint[e].B; q.C = sc.prim_quint[e].C; q.D = sw.qD[pr]; q.E = sw.qE[pr]; q.F = sw.qF[pr];
                float lb = sw.ulb[u], ub = sw.uub[u];
                float t = 0.5f * (lb + ub);
                for (int it = 0; it < 20; it++) {                              // within_distance.h:244-262
                    if (descending || !(t >= lb && t <= ub)) t = 0.5f * (lb + ub);
                    const double value = quintic_eval(q, t);
                    if (fabs(value) < 1e-5f || it == 19) break;
                    if (value > 0.f) ub = t; else lb = t;
                    if (!descending) {   // (a descending bracket replaces the Newton iterate by the midpoint before using it)
                        const double derivative = quintic_deriv(q, t);
                        t = (float)((double)t - newton_quotient(value, derivative));
                    }
                }
                const F4 p01 = sc.prim_p01[e], p23 = sc.prim_p23[e], rad = sc.prim_rad[e];
                const F2 pt = mk2(sw.px[pr], sw.py[pr]);
                const float tt = 1 - t;                                        // :263-267
                const float rr = (tt * tt * tt) * rad.x + (3 * tt * tt * t) * rad.y + (3 * tt * t * t) * rad.z + (t * t * t) * rad.w;
                pass = dist_sq(eval_cubic(mk2(p01.x, p01.y), mk2(p01.z, p01.w), mk2(p23.x, p23.y), mk2(p23.z, p23.w), t), pt) < rr * rr;
            }
            if (pass) atomicOr(&sw.hit[pr >> 5], 1u << (pr & 31));
            __syncwarp();
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < SV_PPL; h++)
            if ((sw.hit[h] >> lane) & 1u) atomicOr(&wv.hit[refs[h] >> 5], 1u << (refs[h] & 31u));
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) k_wave_solve_fill(SceneView sc, WaveView wv) {
    const int count = min(wv.counters[1], wv.cap_f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        WavePair p = wv.pairs_f[i];
        const int ptype = (int)((unsigned)p.prim >> 28);
        p.prim &= 0x0fffffff;
        const int w = prim_winding(ptype, sc.prim_p01[p.prim], sc.prim_p23[p.prim], mk2(p.x, p.y), false);
        const unsigned k = p.ref & 31u;
        if (w != 0) atomicOr(&wv.wind[(size_t)(p.ref >> 5) * 4 + (k >> 3)], (unsigned)(w & 15) << (4 * (k & 7)));
    }
}

// ------------------------------------------------------------------------------------------ W3
// Replay of one item's candidate list with the result words: sample_color's fragment collection and
// compositing (diffvg.cpp:555-653).
template <bool EDGE, bool RECORD>
DVG_D void wave_consume(const SceneView &sc, const BinView &bins, const WaveView &wv, int tile, int64_t cb,
                        SampleTracer<EDGE, RECORD> &tr) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int beg = bins.offsets[tile], end = bins.offsets[tile + 1];
    int c = 0;
    for (int base = beg; base < end; base += 32, c++) {
        const int n = min(32, end - base);
        int tf = 0, inst = 0, group = 0;
        if (lane < n) {
            const PrimMeta pm = sc.prim_meta[bins.items[base + lane]];
            tf = pm.type_flags; inst = pm.inst;
            group = sc.insts[inst].group;
        }
        const int64_t word = (cb + c) * 32 + lane;
        const unsigned hitm = wv.hit[word];
        uint4 wd; wd.x = wd.y = wd.z = wd.w = 0u;
        if (wv.wind) wd = reinterpret_cast<const uint4 *>(wv.wind)[word];
        for (int k = 0; k < n; k++) {
            PrimRef pr;
            pr.tf = __shfl_sync(FULL, tf, k); pr.inst = __shfl_sync(FULL, inst, k); pr.group = __shfl_sync(FULL, group, k);
            const unsigned ww = (k < 8 ? wd.x : (k < 16 ? wd.y : (k < 24 ? wd.z : wd.w)));
            const int nib = (int)((ww >> (4 * (k & 7))) & 15u);
            tr.template step<TM_CONSUME>(sc, pr, DVG_NEED_STROKE | DVG_NEED_FILL, ((hitm >> k) & 1u) != 0u, (nib ^ 8) - 8);
        }
    }
    tr.finish(sc);
}

// render_kernel (diffvg.cpp:1161-1272), colour output, after the candidates have been answered.
template <bool BACKWARD>
__global__ void __launch_bounds__(WB, DVG_WB_MIN + 1) k_wave_composite_px(SceneView sc, BinView bins, RenderArgs ra_all, WaveView wv, int num_items) {
    const GlobalSink sk{BACKWARD ? grad_replica(ra_all) : nullptr};
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * WNW + (threadIdx.x >> 5);
    const int spp = ra_all.nsx * ra_all.nsy;
    const bool pow2 = (spp & (spp - 1)) == 0;
    const int grp = pow2 ? (spp < 32 ? spp : 32) : 1;
    const bool box_fast = BACKWARD && sc.filter.type == 0 && pow2 && (int)ceilf(sc.filter.radius) == 1;   // filter_radius_grad_box
    int fkey[BACKWARD ? W_MAXF : 1];
    F4 fprev[BACKWARD ? W_MAXF : 1];
    float d_radius_acc = 0.f;
    int radius_off = sc.filter_radius_off;
    if (item < num_items) {
        const PixelItem pi = pixel_item(bins, ra_all, wv, item);
        const RenderArgs ra = args_of_scene(ra_all, pi.scene);   // the scene's seed and image slices (batch)
        radius_off += pi.scene * sc.num_params;
        const int x = pi.x, y = pi.y;
        const bool active = pi.active;
        F2 pt = mk2(0, 0), cpt = mk2(0, 0);
        const float *bg_px = nullptr;
        F4 first = mk4(0, 0, 0, 0);
        F4 d_color = mk4(0, 0, 0, 0);
        if (active) {
            sample_position(sc.canvas_w, sc.canvas_h, ra.width, ra.height, ra.nsx, ra.nsy, ra.seed, ra.use_prefiltering != 0,
                            x, y, pi.sx, pi.sy, pi.idx, pt, cpt);
            if (ra.background) {
                bg_px = ra.background + 4 * (y * ra.width + x);
                first = mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]);
            }
            if (BACKWARD) d_color = gather_d_color(sc.filter, ra.d_render_image, ra.weight_image, ra.width, ra.height, pt);
        }
        SampleTracer<false, BACKWARD> tr;
        tr.init(cpt, active, first, -1, -1, fkey, fprev);
        wave_consume<false, BACKWARD>(sc, bins, wv, pi.tile, pi.cb, tr);
        const F4 color = tr.resolve(bg_px);
        if (!BACKWARD) {
            splat_color(sc, ra, x, y, pt, color, active, grp, lane);
        } else {
            // interior backward, d_sample_color (diffvg.cpp:656-705)
            float dcr = d_color.x, dcg = d_color.y, dcb = d_color.z, dca = d_color.w;
            int sp = tr.sp;
            if (tr.nfrag > 0) {
                if (tr.accum.w > 1e-6f) {
                    const float inv = 1.f / tr.accum.w;
                    dca -= (d_color.x * color.x + d_color.y * color.y + d_color.z * color.z) / tr.accum.w;
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
            // every lane walks its own fragment stack (diffvg.cpp:656-705) and issues its own atomics: reducing lanes that
            // share a colour record with shuffles first measured 10% slower (the shuffle chains stall the warp)
            while (sp > 0) {
                const int key = fkey[--sp];
                const GroupInfo &g = sc.groups[key >> 1];
                const int ctype = (key & 1) ? g.stroke_type : g.fill_type;
                const int coff = (key & 1) ? g.stroke_off : g.fill_off;
                const int cstops = (key & 1) ? g.stroke_stops : g.fill_stops;
                const F4 prev = fprev[sp];
                const F4 fc = eval_color(ctype, sc.params + coff, cstops, cpt);
                const float d_prev_alpha = dca * (1.f - fc.w);
                float d_alpha_i = dca * (1.f - prev.w);
                d_alpha_i += (dcr * (fc.x - prev.x) + dcg * (fc.y - prev.y)) + dcb * (fc.z - prev.z);
                const F4 dc = mk4(dcr * fc.w, dcg * fc.w, dcb * fc.w, d_alpha_i);
                dcr = dcr * (1 - fc.w); dcg = dcg * (1 - fc.w); dcb = dcb * (1 - fc.w);
                dca = d_prev_alpha;
                if (ctype == 0) { sk.add(coff + 0, dc.x); sk.add(coff + 1, dc.y); sk.add(coff + 2, dc.z); sk.add(coff + 3, dc.w); }
                // (Q4: the reference has no storage for the gradient of a gradient-typed STROKE colour -- scene.cpp:866-889
                // never assigns d_shape_groups[g].stroke_color -- and faults there; the gradient is accumulated like a fill's)
                else d_eval_gradient(ctype, sc.params + coff, cstops, cpt, dc, sk, coff,
                                                     ra.d_translation ? ra.d_translation + 2 * (y * ra.width + x) : nullptr);
            }
            if (active && had_frags && bg_px && ra.d_background) {  // diffvg.cpp:699-704
                float *d = ra.d_background + 4 * (y * ra.width + x);
                atomicAdd(d + 0, dcr); atomicAdd(d + 1, dcg); atomicAdd(d + 2, dcb); atomicAdd(d + 3, dca);
            }
            if (!(ra.flags & 4u)) {   // DVG_BWD_SKIP_FILTER_GRAD
                // (warp-uniform branch: every lane of the warp is in this block, see below)
                if (box_fast) d_radius_acc += filter_radius_grad_box(sc, ra, x, y, pt, color, active, grp, lane);
                else if (active) d_radius_acc += filter_radius_grad(sc, ra, x, y, pt, color);
            }
        }
    }
    if (BACKWARD) {   // every sample adds to d_filter.radius
        d_radius_acc = warp_sum(d_radius_acc);
        if (lane == 0) sk.add(radius_off, d_radius_acc);
    }
}

// render_edge_kernel (diffvg.cpp:1388-1475) after the candidates of both sides have been answered.
__global__ void __launch_bounds__(WB, DVG_WB_MIN + 1) k_wave_composite_edge(SceneView sc, BinView bins, RenderArgs ra_all, BoundaryWork bw, WaveView wv) {
    float *const D = grad_replica(ra_all);
    const GlobalSink sk{D};
    const int lane = threadIdx.x & 31;
    const int ntiles = bin_total_tiles(bins);
    const int item = blockIdx.x * WNW + (threadIdx.x >> 5);
    if (item >= bw.blk_offsets[ntiles]) return;
    {
        const EdgeItem ei = edge_item(bins, bw, wv, item);
        const RenderArgs ra = args_of_scene(ra_all, ei.tile / bin_scene_tiles(bins));   // the scene's image slices (batch)
        const EdgeLane el = edge_lane(sc, ra, bw, ei);
        const BoundarySample &bs = el.bs;
        const bool active = el.active;
        const int side = lane & 1;
        int q_group = -1, q_shape = -1;
        const float *bg_px = nullptr;
        F4 first = mk4(0, 0, 0, 0);
        if (active) {
            const InstInfo &ii = sc.insts[bs.inst];
            q_group = ii.group; q_shape = ii.shape;
            if (ra.background) {
                bg_px = ra.background + 4 * (el.by * ra.width + el.bx);
                first = mk4(bg_px[0], bg_px[1], bg_px[2], bg_px[3]);
            }
        }
        SampleTracer<true, false> tr;
        tr.init(el.cpt, active, first, q_group, q_shape, nullptr, nullptr);
        wave_consume<true, false>(sc, bins, wv, ei.tile, ei.cb, tr);
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
        float dm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        int xoff = -1;
        GradRec gr;
        gr.key = -1;
#pragma unroll
        for (int j = 0; j < DVG_GREC_N; j++) { gr.addr[j] = -1; gr.val[j] = 0.f; }
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
            boundary_gradient_record(sc, bs, ii, contrib, normal, gr);
            if (ra.debug_out) {
                float *o = ra.debug_out + 4 * (size_t)bw.sorted_idx[bw.tile_offsets[ei.tile] + ei.k];
                o[0] = contrib; o[1] = (float)(my_hit | (other_hit << 1)); o[2] = normal.x; o[3] = normal.y;
            }
            if (!(ra.flags & 1u)) {  // DVG_BWD_SKIP_XFORM_GRAD
                boundary_xform_gradient(bs, g, contrib, normal, dm);
                xoff = g.xform_off;
            }
            if (ra.d_translation) {  // diffvg.cpp:1454-1461
                atomicAdd(ra.d_translation + 2 * (el.by * ra.width + el.bx) + 0, normal.x * contrib);
                atomicAdd(ra.d_translation + 2 * (el.by * ra.width + el.bx) + 1, normal.y * contrib);
            }
        }
        scatter_record(gr, D);
        // d_shape_to_canvas: warp-reduce when every scattering lane targets the same transform
        const unsigned am = __ballot_sync(0xffffffffu, xoff >= 0);
        if (am) {
            const int x0 = __shfl_sync(0xffffffffu, xoff, __ffs(am) - 1);
            const bool uniform = __all_sync(0xffffffffu, xoff < 0 || xoff == x0);
            if (uniform) {
#pragma unroll
                for (int c = 0; c < 9; c++) {
                    const float v = warp_sum(dm[c]);
                    if (lane == 0) sk.add(x0 + c, v);
                }
            } else if (xoff >= 0) {
#pragma unroll
                for (int c = 0; c < 9; c++) sk.add(xoff + c, dm[c]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
// chunks per tile (ceil(count / 32)) and their maximum, for sizing the result words
__global__ void k_wave_tile_chunks(const int *bin_offsets, int *nch, int *max_nch, int ntiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int c = (bin_offsets[t + 1] - bin_offsets[t] + 31) / 32;
    nch[t] = c;
    atomicMax(max_nch, c);
}

// per tile: boundary items (ceil(samples / 16)) and items * chunks
__global__ void k_wave_edge_counts(BoundaryWork bw, const int *tile_choff, int *edge_chunks, int ntiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int items = (bw.tile_counts[t] + W_EDGE_SPI - 1) / W_EDGE_SPI;
    bw.blk_counts[t] = items;
    edge_chunks[t] = items * (tile_choff[t + 1] - tile_choff[t]);
}

// item -> tile table (one thread per tile writes its run of items)
__global__ void k_wave_edge_items(BoundaryWork bw, int ntiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int b = bw.blk_offsets[t], e = bw.blk_offsets[t + 1];
    for (int i = b; i < e; i++) bw.item_tile[i] = t;
}

void launch_wave_tile_chunks(const int *bin_offsets, int *nch, int *choff, int *max_nch, int ntiles, int *scan_ws, cudaStream_t st) {
    cudaMemsetAsync(max_nch, 0, sizeof(int), st);
    DVG_LAUNCH(k_wave_tile_chunks, dim3((ntiles + 255) / 256), dim3(256), 0, st, bin_offsets, nch, max_nch, ntiles);
    launch_scan(nch, choff, ntiles, scan_ws, st);
}

void launch_wave_reduce_grads(const RenderArgs &ra, cudaStream_t st) {
    DVG_LAUNCH(k_wave_reduce_grads, dim3((ra.num_params + 255) / 256), dim3(256), 0, st, ra.d_params_rep, ra.grad_reps, ra.num_params,
               ra.d_params);
}

int wave_pixel_items(const BinView &bins, const RenderArgs &ra) {
    const int ns = bins.tile_w * bins.tile_h * ra.nsx * ra.nsy;
    const int wpt = (ns + 31) / 32;
    const int r0 = ra.row_begin / bins.tile_h;
    const int r1 = (ra.row_end + bins.tile_h - 1) / bins.tile_h;
    return bins.batch > 1 ? bin_total_tiles(bins) * wpt : (r1 - r0) * bins.tiles_x * wpt;   // (a batch renders whole images)
}
int wave_items_per_tile(const BinView &bins, const RenderArgs &ra) {
    return (bins.tile_w * bins.tile_h * ra.nsx * ra.nsy + 31) / 32;
}
int wave_edge_samples_per_item() { return W_EDGE_SPI; }

void launch_wave_classify_px(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const WaveView &wv, cudaStream_t st) {
    const int items = wave_pixel_items(bins, ra);
    if (items <= 0) return;
    if (ra.use_prefiltering) DVG_LAUNCH_AS("k_wave_classify_px<false>", kc_px_pf, dim3((items + WNW - 1) / WNW), dim3(WB), 0, st, sc, bins, ra, wv, items);
    else if (wv.wind) DVG_LAUNCH_AS("k_wave_classify_px<false>", kc_px_fills, dim3((items + WNW - 1) / WNW), dim3(WB), 0, st, sc, bins, ra, wv, items);
    else DVG_LAUNCH_AS("k_wave_classify_px<false>", kc_px, dim3((items + WNW - 1) / WNW), dim3(WB), 0, st, sc, bins, ra, wv, items);
}
void launch_wave_retry_px(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const WaveView &wv, cudaStream_t st) {
    const int items = wave_pixel_items(bins, ra);
    if (items <= 0) return;
    const dim3 grid(std::min((items + WNW - 1) / WNW, g_num_sms * DVG_WB_MIN));
    if (ra.use_prefiltering) DVG_LAUNCH_AS("k_wave_classify_px<true>", kr_px_pf, grid, dim3(WB), 0, st, sc, bins, ra, wv, items);
    else if (wv.wind) DVG_LAUNCH_AS("k_wave_classify_px<true>", kr_px_fills, grid, dim3(WB), 0, st, sc, bins, ra, wv, items);
    else DVG_LAUNCH_AS("k_wave_classify_px<true>", kr_px, grid, dim3(WB), 0, st, sc, bins, ra, wv, items);
}

// Grids are a fixed multiple of the SM count (bounded by the queue capacity); the counts stay on the device.
static int stride_grid(int64_t cap, int block, int per_sm) {
    const int64_t need = (cap + block - 1) / block;
    const int64_t persistent = (int64_t)g_num_sms * per_sm;
    return (int)std::max<int64_t>(1, std::min(need, persistent));
}

void launch_wave_solve(const SceneView &sc, const WaveView &wv, bool strokes, bool fills, cudaStream_t st) {
    if (strokes && wv.cap_s > 0)
        DVG_LAUNCH(k_wave_stroke_solve, dim3(stride_grid(wv.cap_s, SV_NW * SV_PAIRS, DVG_SOLVE_PER_SM)), dim3(SV_B), 0, st, sc, wv);
    if (fills && wv.cap_f > 0) DVG_LAUNCH(k_wave_solve_fill, dim3(stride_grid(wv.cap_f, 128, 32)), dim3(128), 0, st, sc, wv);
}

void launch_wave_composite_px(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const WaveView &wv, bool backward,
                              cudaStream_t st) {
    const int items = wave_pixel_items(bins, ra);
    if (items <= 0) return;
    if (backward) DVG_LAUNCH(k_wave_composite_px<true>, dim3((items + WNW - 1) / WNW), dim3(WB), 0, st, sc, bins, ra, wv, items);
    else DVG_LAUNCH(k_wave_composite_px<false>, dim3((items + WNW - 1) / WNW), dim3(WB), 0, st, sc, bins, ra, wv, items);
}

// Boundary pass, ordering step (shared with the fused path's kernels in dvg_render.cu): tile keys, counting
// sort, then per-tile item counts and chunk-slot offsets.
void launch_wave_boundary_sort(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw,
                               const WaveView &wv, int *edge_chunks, cudaStream_t st) {
    const int ntiles = bin_total_tiles(bins);
    launch_boundary_sort(sc, bins, ra, bw, st);
    DVG_LAUNCH(k_wave_edge_counts, dim3((ntiles + 255) / 256), dim3(256), 0, st, bw, wv.tile_choff, edge_chunks, ntiles);
    launch_scan(bw.blk_counts, bw.blk_offsets, ntiles, bw.scan_ws, st);
    launch_scan(edge_chunks, wv.edge_choff, ntiles, bw.scan_ws, st);
    DVG_LAUNCH(k_wave_edge_items, dim3((ntiles + 255) / 256), dim3(256), 0, st, bw, ntiles);
}

void launch_wave_classify_edge(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw,
                               const WaveView &wv, cudaStream_t st) {
    if (wv.wind) DVG_LAUNCH_AS("k_wave_classify_edge<false>", kc_edge_fills, dim3((bw.max_blocks + WNW - 1) / WNW), dim3(WB), 0, st, sc, bins, ra, bw, wv);
    else DVG_LAUNCH_AS("k_wave_classify_edge<false>", kc_edge, dim3((bw.max_blocks + WNW - 1) / WNW), dim3(WB), 0, st, sc, bins, ra, bw, wv);
}
void launch_wave_retry_edge(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw,
                            const WaveView &wv, cudaStream_t st) {
    const dim3 grid(std::min((bw.max_blocks + WNW - 1) / WNW, g_num_sms * DVG_WB_MIN));
    if (wv.wind) DVG_LAUNCH_AS("k_wave_classify_edge<true>", kr_edge_fills, grid, dim3(WB), 0, st, sc, bins, ra, bw, wv);
    else DVG_LAUNCH_AS("k_wave_classify_edge<true>", kr_edge, grid, dim3(WB), 0, st, sc, bins, ra, bw, wv);
}

void launch_wave_composite_edge(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw,
                                const WaveView &wv, cudaStream_t st) {
    DVG_LAUNCH(k_wave_composite_edge, dim3((bw.max_blocks + WNW - 1) / WNW), dim3(WB), 0, st, sc, bins, ra, bw, wv);
}

}  // namespace dvg
