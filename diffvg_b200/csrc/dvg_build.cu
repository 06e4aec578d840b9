// dvg_build.cu -- per-iteration scene build on the GPU: shape lengths, boundary-sampling
// CDFs/PMFs, bounding boxes, primitive table and the tile bins the render kernels traverse.
//
// Replaces the single-threaded host work the reference redoes in every forward
// (scene.cpp:113-333 lengths/CDFs, 496-684 boxes + three BVH levels).  Float prefix sums are
// kept in the reference's sequential order (one thread per path; one thread for the shape
// CDF) so that CDF entries -- and therefore which shape/segment a boundary sample picks --
// are bit-identical.
#include "dvg_internal.h"

namespace dvg {

// (indices run over all scenes of a batch: scene = index / per-scene count, see SceneView)
// One WARP per shape.  Paths: the lanes measure the segments (dvg_buildfn.cuh path_segment_measure: the arithmetic of the
// sequential build_shape) and take the bounding box in parallel; lane 0 then forms the float sums in the reference's
// sequential order (scene.cpp:132-191, 257-326: which segment a boundary sample picks depends on them bit for bit).  One
// thread per shape walking a 100-segment path alone took 0.12 ms at flower.svg, on every rank of a row-sharded render.
__global__ void k_build_shapes(BuildView bv) {
    const int s_batch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s_batch >= bv.num_shapes * bv.batch) return;
    const unsigned FULL = 0xffffffffu;
    const int *topo = bv.topo;
    const int scene = s_batch / bv.num_shapes, s = s_batch - scene * bv.num_shapes;
    const int *r = topo + topo[DVG_H_OFF_SHAPES] + s * DVG_SHAPE_REC_LEN;
    if (r[DVG_S_TYPE] != DVG_SHAPE_PATH) {
        if (lane == 0) build_shape(bv, s_batch);
        return;
    }
    const float *P = bv.params + (size_t)scene * bv.num_params;
    const int seg_base = scene * bv.total_segs;
    const float *p = P + r[DVG_S_PARAM_OFF];
    const float stroke_width = r[DVG_S_WIDTH_OFF] >= 0 ? P[r[DVG_S_WIDTH_OFF]] : 0.f;
    const int np = r[DVG_S_NUM_POINTS], nseg = r[DVG_S_NUM_SEGS];
    const int *ncp = topo + topo[DVG_H_OFF_NCP] + r[DVG_S_NCP_OFF];
    const float *thick = r[DVG_S_THICK_OFF] >= 0 ? P + r[DVG_S_THICK_OFF] : nullptr;
    float *seg_pmf = bv.seg_pmf + seg_base + r[DVG_S_NCP_OFF];
    float *seg_cdf = bv.seg_cdf + seg_base + r[DVG_S_NCP_OFF];
    int *seg_pid = bv.seg_point_id + seg_base + r[DVG_S_NCP_OFF];
    // bounding box of the control points
    Box box; box.x0 = box.y0 = INFINITY; box.x1 = box.y1 = -INFINITY;
    for (int i = lane; i < np; i += 32) {
        const float x = p[2 * i], y = p[2 * i + 1];
        box.x0 = rminf(x, box.x0); box.y0 = rminf(y, box.y0);
        box.x1 = rmaxf(x, box.x1); box.y1 = rmaxf(y, box.y1);
    }
    for (int o = 16; o > 0; o >>= 1) {
        box.x0 = rminf(__shfl_xor_sync(FULL, box.x0, o), box.x0); box.y0 = rminf(__shfl_xor_sync(FULL, box.y0, o), box.y0);
        box.x1 = rmaxf(__shfl_xor_sync(FULL, box.x1, o), box.x1); box.y1 = rmaxf(__shfl_xor_sync(FULL, box.y1, o), box.y1);
    }
    // segments: first point by a running scan of the control-point counts, raw lengths parked in seg_pmf, the y-sort
    // key of thickness paths in seg_cdf
    int carry = 0;
    for (int base = 0; base < nseg; base += 32) {
        const int i = base + lane;
        const int n = i < nseg ? ncp[i] : -1;
        int incl = n + 1;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += u;
        }
        const int pid = carry + incl - (n + 1);
        if (i < nseg) {
            float d, yc, th;
            path_segment_measure(p, thick, np, n, pid, stroke_width, d, yc, th);
            seg_pid[i] = pid;
            seg_pmf[i] = d;
            if (thick) seg_cdf[i] = yc;
        }
        carry += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    float inv_length = 0.f, len = 0.f, r0q = stroke_width;
    if (lane == 0) {
        float length = 0.f;
        for (int i = 0; i < nseg; i++) length += seg_pmf[i];
        len += length;
        if (thick) {   // radius of the first leaf after the reference's y-sort (scene.cpp:602-618): the first minimum
            float best_y = INFINITY;
            int best = -1;
            for (int i = 0; i < nseg; i++) {
                const float yc = seg_cdf[i];
                if (yc < best_y) { best_y = yc; best = i; }
            }
            if (best >= 0) {
                float d, yc, th;
                path_segment_measure(p, thick, np, ncp[best], seg_pid[best], stroke_width, d, yc, th);
                r0q = th;
            }
        }
        inv_length = 1.f / len;
    }
    inv_length = __shfl_sync(FULL, inv_length, 0);
    for (int i = lane; i < nseg; i += 32) seg_pmf[i] = seg_pmf[i] * inv_length;
    __syncwarp();
    if (lane == 0) {
        float c = 0.f;
        for (int i = 0; i < nseg; i++) {
            const float d = seg_pmf[i];
            c = (i == 0) ? d : d + c;
            seg_cdf[i] = c;
        }
        bv.shapes_length[s_batch] = len;
        bv.shape_box[s_batch] = box;
        bv.shape_r0[s_batch] = r0q;
    }
}
__global__ void k_build_groups(BuildView bv) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < bv.num_groups * bv.batch) build_group(bv, g);
}
__global__ void k_build_prims(BuildView bv) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < bv.num_prims * bv.batch) build_prim(bv, e);
}

// scene.cpp:207-246.  Sequential float prefix sum in the reference's order (single thread),
// then the normalisation in parallel by the rest of the block.
__global__ void k_build_shape_cdf(BuildView bv) {
    // The prefix sum itself must run in the reference's sequential float order (which shape a boundary sample
    // picks depends on it bit for bit), but its operands need not be fetched by the summing thread: the block
    // stages 4096 lengths at a time in shared memory (two dependent global loads per element would cost the
    // lone thread ~0.3 ms at 2048 shapes), thread 0 sums them there, and the block writes the results back.
    // One block per scene of a batch.
    constexpr int CH = 4096;
    __shared__ __align__(16) float s_len[CH];
    __shared__ __align__(16) float s_cdf[CH];
    __shared__ float s_carry;
    const int scene = blockIdx.x;
    const float *shapes_length = bv.shapes_length + scene * bv.num_shapes;
    float *shape_cdf = bv.shape_cdf + scene * bv.num_insts, *shape_pmf = bv.shape_pmf + scene * bv.num_insts;
    if (threadIdx.x == 0) s_carry = 0.f;
    for (int base = 0; base < bv.num_insts; base += CH) {
        const int n = min(CH, bv.num_insts - base);
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_len[i] = shapes_length[bv.inst_shape[base + i]];
        __syncthreads();
        if (threadIdx.x == 0) {
            float c = s_carry;
            int i = 0;
            if (base == 0 && n > 0) { c = s_len[0]; s_cdf[0] = c; i = 1; }   // scene.cpp:217-222: the first entry is assigned
            // eight operands at a time through registers: the adds stay one dependent chain in the reference's order, the
            // shared-memory loads and stores no longer sit inside it (45 -> ~10 cycles per element)
            for (; i < n && (i & 3); i++) { c = s_len[i] + c; s_cdf[i] = c; }    // (up to a 16-byte boundary)
            for (; i + 8 <= n; i += 8) {
                const float4 a = *reinterpret_cast<const float4 *>(&s_len[i]), b4 = *reinterpret_cast<const float4 *>(&s_len[i + 4]);
                float4 o0, o1;
                c = a.x + c; o0.x = c; c = a.y + c; o0.y = c; c = a.z + c; o0.z = c; c = a.w + c; o0.w = c;
                c = b4.x + c; o1.x = c; c = b4.y + c; o1.y = c; c = b4.z + c; o1.z = c; c = b4.w + c; o1.w = c;
                *reinterpret_cast<float4 *>(&s_cdf[i]) = o0; *reinterpret_cast<float4 *>(&s_cdf[i + 4]) = o1;
            }
            for (; i < n; i++) {
                c = s_len[i] + c;
                s_cdf[i] = c;
            }
            s_carry = c;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) { shape_cdf[base + i] = s_cdf[i]; shape_pmf[base + i] = s_len[i]; }
        __syncthreads();
    }
    const float norm = s_carry;
    if (threadIdx.x == 0) {
        // error flag of the batch: the first degenerate scene wins (flags start at 0, see launch_build)
        if (!(norm > 0.f)) atomicCAS(bv.error_flag, 0, 1);            // scene.cpp:231-235 (also catches NaN)
        else if (isinf(norm)) atomicCAS(bv.error_flag, 0, 2);         // scene.cpp:236-240
        if (scene == 0 || !(norm > 0.f) || isinf(norm)) *bv.total_length = norm;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bv.num_insts; i += blockDim.x) {
        shape_cdf[i] /= norm;
        shape_pmf[i] /= norm;
    }
    // guide table of the normalised cdf for the sample-generation kernel (scene 0: batches search their tables plainly)
    if (scene == 0 && bv.shape_guide) {
        __syncthreads();
        for (int k = threadIdx.x; k <= DVG_CDF_GUIDE; k += blockDim.x)
            bv.shape_guide[k] = cdf_sample(shape_cdf, bv.num_insts, (float)k / (float)DVG_CDF_GUIDE, nullptr);
    }
}

// ------------------------------------------------------------------ tile bins
// One warp per tile.  Pass 0 counts, pass 1 fills (identical traversal; ballot prefix keeps
// the primitive ids ascending inside a tile, i.e. already in compositing order).
DVG_D bool overlaps(Box a, float x0, float y0, float x1, float y1) {
    return a.x0 <= x1 && a.x1 >= x0 && a.y0 <= y1 && a.y1 >= y0;
}

// Canvas-space rectangle of tiles [tx0, tx1] x [ty0, ty1] (inclusive), with a margin that also covers the +-1e-4
// (normalised) offsets of boundary samples (diffvg.cpp:1416,1420) and float rounding of pt/W*canvas_w.
// int(-0.9) == 0: the reference attributes boundary samples lying up to one pixel left of / above the image to pixel
// column / row 0 (diffvg.cpp:1405-1409), so border tiles reach out 1 px.
DVG_D void tiles_rect(const BuildView &bv, const BinBuild &bb, int tx0, int ty0, int tx1, int ty1, float &x0, float &y0, float &x1, float &y1) {
    const float cw = (float)bv.canvas_w, ch = (float)bv.canvas_h;
    const float margin = 4e-4f * (cw > ch ? cw : ch) + 1e-4f;
    x0 = ((float)(tx0 * bb.tile_w - (tx0 == 0 ? 1 : 0)) / (float)bb.width) * cw - margin;
    x1 = ((float)((tx1 + 1) * bb.tile_w) / (float)bb.width) * cw + margin;
    y0 = ((float)(ty0 * bb.tile_h - (ty0 == 0 ? 1 : 0)) / (float)bb.height) * ch - margin;
    y1 = ((float)((ty1 + 1) * bb.tile_h) / (float)bb.height) * ch + margin;
}

// Level 1 of the two-level binning: one BLOCK per supertile walks ALL primitives (bounding boxes only); the per-tile
// kernel then walks ~1% of them.  Without it every one of 16 k (512^2) .. 65 k (2048^2) tiles walked every primitive:
// 0.42 ms of a 7.2 ms step, and the largest cost a row shard repeats on every GPU.  The eight warps of the block take
// consecutive slices of the primitive list (count, block offsets, write: the supertile's list stays ascending); only the
// supertile rows of the tile rows that will be binned are visited.  (One warp per supertile walking 10 k primitives alone
// took 0.1 ms whatever the size of the band.)
constexpr int BC_B = 256;
__global__ void __launch_bounds__(BC_B) k_bin_coarse(BuildView bv, BinBuild bb, int srow0) {
    __shared__ int s_cnt[BC_B / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int si = blockIdx.x + srow0 * bb.stiles_x;
    const int sx = si % bb.stiles_x, sy = si / bb.stiles_x;
    float x0, y0, x1, y1;
    tiles_rect(bv, bb, sx * bb.super, sy * bb.super, min((sx + 1) * bb.super, bb.tiles_x) - 1, min((sy + 1) * bb.super, bb.tiles_y) - 1, x0, y0, x1, y1);
    int *out = bb.s_items + (size_t)si * bv.num_prims;
    const int slice = ((bv.num_prims + BC_B - 1) / BC_B) * 32;   // primitives per warp, a multiple of 32
    const int e_begin = w * slice, e_end = min(bv.num_prims, e_begin + slice);
    const Box *boxes = bb.prefilter ? bv.prim_cbox_pf : bv.prim_cbox;
    int count = 0;
    for (int e0 = e_begin; e0 < e_end; e0 += 32) {
        const int e = e0 + lane;
        count += __popc(__ballot_sync(0xffffffffu, e < e_end && overlaps(boxes[e], x0, y0, x1, y1)));
    }
    if (lane == 0) s_cnt[w] = count;
    __syncthreads();
    int at = 0, total = 0;
    for (int j = 0; j < BC_B / 32; j++) { if (j < w) at += s_cnt[j]; total += s_cnt[j]; }
    for (int e0 = e_begin; e0 < e_end; e0 += 32) {
        const int e = e0 + lane;
        const bool ph = e < e_end && overlaps(boxes[e], x0, y0, x1, y1);
        const unsigned pmask = __ballot_sync(0xffffffffu, ph);
        if (ph) out[at + __popc(pmask & ((1u << lane) - 1))] = e;
        at += __popc(pmask);
    }
    if (threadIdx.x == 0) bb.s_counts[si] = total;
}

template <int PASS>
__global__ void k_bin(BuildView bv, BinBuild bb) {
    const int lane = threadIdx.x & 31;
    const int tiles_scene = bb.tiles_x * bb.tiles_y;
    const int warp = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + bb.tile_row0 * bb.tiles_x;   // = tile, over all scenes of a batch
    if (warp >= (bb.batch - 1) * tiles_scene + bb.tile_row1 * bb.tiles_x) return;
    const int scene = warp / tiles_scene, ltile = warp - scene * tiles_scene;
    const int tx = ltile % bb.tiles_x, ty = ltile / bb.tiles_x;
    const int e_base = scene * bv.num_prims, g_base = scene * bv.num_groups;   // this scene's slice of the primitive / group tables
    float x0, y0, x1, y1;
    tiles_rect(bv, bb, tx, ty, tx, ty, x0, y0, x1, y1);
    int count = 0;
    int *out = nullptr;
    if (PASS == 1) out = bb.items + bb.offsets[warp];
    if (bb.super) {   // (single scenes only)
        const int si = (ty / bb.super) * bb.stiles_x + tx / bb.super;
        const int *list = bb.s_items + (size_t)si * bv.num_prims;
        const int n = bb.s_counts[si];
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int e = i0 + lane < n ? list[i0 + lane] : -1;
            bool ph = e >= 0 && overlaps(bb.prefilter ? bv.prim_cbox_pf[e] : bv.prim_cbox[e], x0, y0, x1, y1);
            if (ph && !bb.prefilter && (bv.prim_meta[e].type_flags & DVG_PF_TIGHT))
                ph = bracket_reaches_tile(bv.prim_cap + (size_t)e * DVG_CAP_F4, x0, y0, x1, y1);
            const unsigned pmask = __ballot_sync(0xffffffffu, ph);
            if (PASS == 1 && ph) out[count + __popc(pmask & ((1u << lane) - 1))] = e;
            count += __popc(pmask);
        }
    } else if (bb.flat) {
        // few primitives per group (painterly strokes: 2 per group): walk the primitives directly, 32 per trip;
        // their canvas boxes are already clipped to the group's scene-BVH leaf box, so the group test adds nothing
        for (int e0 = 0; e0 < bv.num_prims; e0 += 32) {
            const int e = e_base + e0 + lane;
            bool ph = e0 + lane < bv.num_prims && overlaps(bb.prefilter ? bv.prim_cbox_pf[e] : bv.prim_cbox[e], x0, y0, x1, y1);
            if (ph && !bb.prefilter && (bv.prim_meta[e].type_flags & DVG_PF_TIGHT))
                ph = bracket_reaches_tile(bv.prim_cap + (size_t)e * DVG_CAP_F4, x0, y0, x1, y1);
            const unsigned pmask = __ballot_sync(0xffffffffu, ph);
            if (PASS == 1 && ph) out[count + __popc(pmask & ((1u << lane) - 1))] = e;
            count += __popc(pmask);
        }
    } else
    for (int g0 = 0; g0 < bv.num_groups; g0 += 32) {
        int g = g0 + lane;
        bool hit = false;
        if (g < bv.num_groups) {
            if (bv.num_groups == 1) hit = true;
            else {
                const GroupInfo &gi = bv.groups[g_base + g];
                Box b = gi.scene_box; float r = gi.scene_r;
                b.x0 -= r; b.y0 -= r; b.x1 += r; b.y1 += r;
                hit = overlaps(b, x0, y0, x1, y1) || !(b.x0 == b.x0 && b.x1 == b.x1 && b.y0 == b.y0 && b.y1 == b.y1);
            }
        }
        unsigned gm = __ballot_sync(0xffffffffu, hit);
        while (gm) {
            int gl = __ffs(gm) - 1;
            gm &= gm - 1;
            const GroupInfo &gi = bv.groups[g_base + g0 + gl];
            for (int e0 = gi.prim_begin; e0 < gi.prim_end; e0 += 32) {
                int e = e0 + lane;
                bool ph = e < gi.prim_end && overlaps(bb.prefilter ? bv.prim_cbox_pf[e] : bv.prim_cbox[e], x0, y0, x1, y1);
                if (ph && !bb.prefilter && (bv.prim_meta[e].type_flags & DVG_PF_TIGHT))
                    ph = bracket_reaches_tile(bv.prim_cap + (size_t)e * DVG_CAP_F4, x0, y0, x1, y1);
                unsigned pmask = __ballot_sync(0xffffffffu, ph);
                if (PASS == 1 && ph) out[count + __popc(pmask & ((1u << lane) - 1))] = e;
                count += __popc(pmask);
            }
        }
    }
    if (PASS == 0 && lane == 0) bb.counts[warp] = count;
}

// Single-block exclusive scan of `n` ints (tile counts: 16 k at 512^2, 262 k at 2048^2); out has n+1 entries.  The block
// walks the array in rounds of 1024 x 4 elements: coalesced 16-byte loads, a shuffle scan per warp, one shared-memory step
// across the 32 warps, a running carry.
__global__ void __launch_bounds__(1024) k_exclusive_scan(const int *in, int *out, int n) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 4096) {
        const int i = base + 4 * t;
        int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(in) & 15) == 0)) {
            const int4 q = *reinterpret_cast<const int4 *>(in + i);
            v0 = q.x; v1 = q.y; v2 = q.z; v3 = q.w;
        } else {
            if (i < n) v0 = in[i];
            if (i + 1 < n) v1 = in[i + 1];
            if (i + 2 < n) v2 = in[i + 2];
            if (i + 3 < n) v3 = in[i + 3];
        }
        const int mine = v0 + v1 + v2 + v3;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        if (w == 0) {
            int x = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += u;
            }
            s_warp[lane] = x;   // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        int run = carry + (w ? s_warp[w - 1] : 0) + incl - mine;
        if (i < n) out[i] = run;
        run += v0;
        if (i + 1 < n) out[i + 1] = run;
        run += v1;
        if (i + 2 < n) out[i + 2] = run;
        run += v2;
        if (i + 3 < n) out[i + 3] = run;
        __syncthreads();
        if (t == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (t == 0) out[n] = s_carry;
}

// Candidate count of every tile ROW (one warp per row): the cost profile a row partition is balanced with.
__global__ void k_tile_row_costs(const int *offsets, int tiles_x, int tiles_y, float *out) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= tiles_y) return;
    int sum = 0;
    for (int t = lane; t < tiles_x; t += 32) sum += offsets[row * tiles_x + t + 1] - offsets[row * tiles_x + t];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) out[row] = (float)sum;
}
void launch_tile_row_costs(const int *offsets, int tiles_x, int tiles_y, float *out, cudaStream_t st) {
    DVG_LAUNCH(k_tile_row_costs, dim3((tiles_y * 32 + 127) / 128), dim3(128), 0, st, offsets, tiles_x, tiles_y, out);
}

void launch_build(const BuildView &bv, cudaStream_t st) {
    const int B = 128;
    cudaMemsetAsync(bv.error_flag, 0, sizeof(int), st);
    DVG_LAUNCH(k_build_shapes, dim3((bv.num_shapes * bv.batch * 32 + B - 1) / B), dim3(B), 0, st, bv);   // a warp per shape
    DVG_LAUNCH(k_build_groups, dim3((bv.num_groups * bv.batch + B - 1) / B), dim3(B), 0, st, bv);
    DVG_LAUNCH(k_build_prims, dim3((bv.num_prims * bv.batch + B - 1) / B), dim3(B), 0, st, bv);
    DVG_LAUNCH(k_build_shape_cdf, dim3(bv.batch), dim3(256), 0, st, bv);
}

void launch_bin_coarse(const BuildView &bv, const BinBuild &bb, cudaStream_t st) {
    if (!bb.super) return;
    const int srow0 = bb.tile_row0 / bb.super, srow1 = (bb.tile_row1 + bb.super - 1) / bb.super;   // supertile rows of the binned tile rows
    const int ns = bb.stiles_x * (srow1 - srow0);
    if (ns > 0) DVG_LAUNCH(k_bin_coarse, dim3(ns), dim3(BC_B), 0, st, bv, bb, srow0);
}
void launch_bin_count(const BuildView &bv, const BinBuild &bb, cudaStream_t st) {
    const int ntiles = bb.tiles_x * bb.tiles_y * bb.batch;
    const int nbin = bb.batch > 1 ? ntiles : (bb.tile_row1 - bb.tile_row0) * bb.tiles_x;
    const int B = 128;  // 4 warps = 4 tiles per block
    if (nbin < ntiles) cudaMemsetAsync(bb.counts, 0, sizeof(int) * ntiles, st);
    if (nbin > 0) DVG_LAUNCH(k_bin<0>, dim3((nbin * 32 + B - 1) / B), dim3(B), 0, st, bv, bb);
    launch_scan(bb.counts, bb.offsets, ntiles, bb.scan_ws, st);
}
void launch_bin_fill(const BuildView &bv, const BinBuild &bb, cudaStream_t st) {
    const int nbin = bb.batch > 1 ? bb.tiles_x * bb.tiles_y * bb.batch : (bb.tile_row1 - bb.tile_row0) * bb.tiles_x;
    const int B = 128;
    if (nbin > 0) DVG_LAUNCH(k_bin<1>, dim3((nbin * 32 + B - 1) / B), dim3(B), 0, st, bv, bb);
}
// Multi-block form: block b scans its 4096 elements and parks its total in ws[b]; the LAST block to finish (ticket in
// ws[DVG_SCAN_WS_BLOCKS]) turns the totals into exclusive block offsets and writes the grand total to out[n]; a second
// kernel adds the block offsets.  Two short launches instead of one block walking 64 chunks at 2048^2 (33 -> ~8 us per
// scan; five scans per iteration, all of it work every rank of a row-sharded render repeats).
__global__ void __launch_bounds__(1024) k_scan_blocks(const int *in, int *out, int n, int *ws) {
    __shared__ int s_warp[32];
    __shared__ int s_last;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int i = blockIdx.x * 4096 + 4 * t;
    int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(in) & 15) == 0)) {
        const int4 q = *reinterpret_cast<const int4 *>(in + i);
        v0 = q.x; v1 = q.y; v2 = q.z; v3 = q.w;
    } else {
        if (i < n) v0 = in[i];
        if (i + 1 < n) v1 = in[i + 1];
        if (i + 2 < n) v2 = in[i + 2];
        if (i + 3 < n) v3 = in[i + 3];
    }
    const int mine = v0 + v1 + v2 + v3;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        int x = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += u;
        }
        s_warp[lane] = x;   // inclusive over warps
    }
    __syncthreads();
    int run = (w ? s_warp[w - 1] : 0) + incl - mine;
    if (i < n) out[i] = run;
    run += v0;
    if (i + 1 < n) out[i + 1] = run;
    run += v1;
    if (i + 2 < n) out[i + 2] = run;
    run += v2;
    if (i + 3 < n) out[i + 3] = run;
    if (t == 1023) {
        ws[blockIdx.x] = s_warp[31];
        __threadfence();
        s_last = atomicAdd(&ws[DVG_SCAN_WS_BLOCKS], 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // last block: exclusive scan of the block totals (<= 1024 of them), in place
    __threadfence();
    const int nb = gridDim.x;
    const int mine_b = t < nb ? *(volatile int *)&ws[t] : 0;
    int inc_b = mine_b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc_b, o);
        if (lane >= o) inc_b += u;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = inc_b;
    __syncthreads();
    if (w == 0) {
        int x = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += u;
        }
        s_warp[lane] = x;
    }
    __syncthreads();
    const int excl_b = (w ? s_warp[w - 1] : 0) + inc_b - mine_b;
    if (t < nb) ws[t] = excl_b;
    if (t == nb - 1) out[n] = excl_b + mine_b;
    if (t == 0) ws[DVG_SCAN_WS_BLOCKS] = 0;   // the ticket, for the next scan
}
__global__ void __launch_bounds__(1024) k_scan_add(int *out, int n, const int *ws) {
    const int add = ws[blockIdx.x + 1];
    const int i = (blockIdx.x + 1) * 4096 + 4 * threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (i + k < n) out[i + k] += add;
}

void launch_scan(const int *in, int *out, int n, int *ws, cudaStream_t st) {
    const int nb = (n + 4095) / 4096;
    if (!ws || nb > DVG_SCAN_WS_BLOCKS || nb <= 4) {   // (up to 16 k elements the single block is as fast as two launches)
        DVG_LAUNCH(k_exclusive_scan, dim3(1), dim3(1024), 0, st, in, out, n);
        return;
    }
    DVG_LAUNCH(k_scan_blocks, dim3(nb), dim3(1024), 0, st, in, out, n, ws);
    if (nb > 1) DVG_LAUNCH(k_scan_add, dim3(nb - 1), dim3(1024), 0, st, out, n, ws);
}

}  // namespace dvg
