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

namespace dvg {

constexpr int RB = 256;      // threads per render block
constexpr int CHUNK = 32;    // primitives staged per step
constexpr int MAXF = DVG_MAXF;
constexpr int MAX_TILE_PIX = 256;
constexpr int EDGE_SPB = RB / 2;  // boundary samples per block (two lanes per sample)

struct Stage {
    F4 p01[CHUNK], p23[CHUNK], rad[CHUNK];
    Box box[CHUNK];
    float thick[CHUNK];
    int tf[CHUNK], inst[CHUNK], group[CHUNK];
    F4 cap[CHUNK * DVG_CAP_F4];
};

// ------------------------------------------------------------------------------------------
// weight_kernel (diffvg.cpp:1115-1158).  Q1: always uses the jittered position, even when
// the render kernel uses pixel centres for prefiltering.
__global__ void k_weight(SceneView sc, RenderArgs ra) {
    const int n = ra.width * ra.height * ra.nsx * ra.nsy;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
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
// The lockstep tile traversal: all threads of the block walk the tile's candidate list
// together (staged through shared memory in chunks of CHUNK primitives) and feed it to their
// own SampleTracer (dvg_trace.cuh), which reproduces sample_color(scene, ...) of
// diffvg.cpp:525-653 including the EdgeQuery bookkeeping.
template <bool EDGE, bool RECORD>
DVG_D void traverse(const SceneView &sc, const BinView &bins, const int tile, Stage &st,
                    SampleTracer<EDGE, RECORD> &tr) {
    const int beg = bins.offsets[tile], end = bins.offsets[tile + 1];
    for (int base = beg; base < end; base += CHUNK) {
        const int n = min(CHUNK, end - base);
        __syncthreads();
        if ((int)threadIdx.x < n) {
            const int t = threadIdx.x;
            const int e = bins.items[base + t];
            st.p01[t] = sc.prim_p01[e];
            st.p23[t] = sc.prim_p23[e];
            st.rad[t] = sc.prim_rad[e];
            st.box[t] = sc.prim_box[e];
            st.thick[t] = sc.prim_thick[e];
            const PrimMeta pm = sc.prim_meta[e];
            st.tf[t] = pm.type_flags;
            st.inst[t] = pm.inst;
            st.group[t] = sc.insts[pm.inst].group;
        }
        if ((int)threadIdx.x < n * DVG_CAP_F4) {
            const int t = threadIdx.x;
            const int e = bins.items[base + t / DVG_CAP_F4];
            st.cap[t] = sc.prim_cap[(size_t)e * DVG_CAP_F4 + t % DVG_CAP_F4];
        }
        __syncthreads();
        for (int j = 0; j < n; j++) {
            PrimRef pr;
            pr.p01 = st.p01[j]; pr.p23 = st.p23[j]; pr.rad = st.rad[j]; pr.box = st.box[j];
            pr.thick = st.thick[j]; pr.tf = st.tf[j]; pr.inst = st.inst[j]; pr.group = st.group[j];
            pr.cap = reinterpret_cast<const float *>(&st.cap[j * DVG_CAP_F4]);
            tr.step(sc, pr);
        }
    }
    tr.finish(sc);
}

DVG_D float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// render_kernel (diffvg.cpp:1161-1272), colour output.  BACKWARD = the d_render_image != null
// variant: recompute the forward, then d_sample_color (diffvg.cpp:656-705), d_background and
// the filter-radius gradient (1250-1268).
template <bool BACKWARD>
__global__ void __launch_bounds__(RB) k_render(SceneView sc, BinView bins, RenderArgs ra) {
    __shared__ Stage st;
    __shared__ float s_pix[BACKWARD ? 4 : MAX_TILE_PIX * 4];
    GradCache *gcp = nullptr;
    if constexpr (BACKWARD) {
        __shared__ GradCache s_gc;
        gcp = &s_gc;
        grad_cache_init(s_gc);
        __syncthreads();
    }
    const CacheSink sk{gcp, ra.d_params};
    const int tile_row0 = ra.row_begin / bins.tile_h;
    const int tile = blockIdx.x + tile_row0 * bins.tiles_x;
    const int tx = tile % bins.tiles_x, ty = tile / bins.tiles_x;
    const int spp = ra.nsx * ra.nsy;
    const int npix = bins.tile_w * bins.tile_h;
    const int ns = npix * spp;
    const int rounds = (ns + RB - 1) / RB;
    const int tid = threadIdx.x;
    if (!BACKWARD) {
        for (int i = tid; i < npix * 4; i += RB) s_pix[i] = 0.f;
    }
    int fkey[BACKWARD ? MAXF : 1];
    F4 fprev[BACKWARD ? MAXF : 1];
    float d_radius_acc = 0.f;

    for (int round = 0; round < rounds; round++) {
        const int l = round * RB + tid;
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
        traverse<false, BACKWARD>(sc, bins, tile, st, tr);
        const F4 color = tr.resolve(bg_px);
        struct { F4 accum; int nfrag, sp; } to;
        to.accum = tr.accum; to.nfrag = tr.nfrag; to.sp = tr.sp;

        if (!BACKWARD) {
            if (active) {
                // splat (diffvg.cpp:1224-1249)
                const int ri = (int)ceilf(sc.filter.radius);
                for (int dy = -ri; dy <= ri; dy++) {
                    for (int dx = -ri; dx <= ri; dx++) {
                        const int xx = x + dx, yy = y + dy;
                        if (xx >= 0 && xx < ra.width && yy >= 0 && yy < ra.height) {
                            const float fw = filter_weight(sc.filter, (xx + 0.5f) - pos.pt.x, (yy + 0.5f) - pos.pt.y);
                            if (fw == 0.f) continue;
                            const float ws = ra.weight_image[yy * ra.width + xx];
                            if (!(ws > 0)) continue;
                            const float inv_ws = 1.f / ws;  // Vector4 / scalar == * (1.f / s)
                            const F4 wc = mk4((fw * color.x) * inv_ws, (fw * color.y) * inv_ws,
                                              (fw * color.z) * inv_ws, (fw * color.w) * inv_ws);
                            const int lx = xx - tx * bins.tile_w, ly = yy - ty * bins.tile_h;
                            if (lx >= 0 && lx < bins.tile_w && ly >= 0 && ly < bins.tile_h) {
                                float *d = &s_pix[4 * (ly * bins.tile_w + lx)];
                                atomicAdd(d + 0, wc.x); atomicAdd(d + 1, wc.y); atomicAdd(d + 2, wc.z); atomicAdd(d + 3, wc.w);
                            } else {
                                float *d = ra.render_image + 4 * (yy * ra.width + xx);
                                atomicAdd(d + 0, wc.x); atomicAdd(d + 1, wc.y); atomicAdd(d + 2, wc.z); atomicAdd(d + 3, wc.w);
                            }
                        }
                    }
                }
            }
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
            if (active) {
                // filter-radius gradient (diffvg.cpp:1250-1268).  The reference evaluates
                // d_compute_filter_weight for every in-range pixel with weight > 0, even where
                // the filter weight itself is zero.
                const int ri = (int)ceilf(sc.filter.radius);
                for (int dy = -ri; dy <= ri; dy++) {
                    for (int dx = -ri; dx <= ri; dx++) {
                        const int xx = x + dx, yy = y + dy;
                        if (xx >= 0 && xx < ra.width && yy >= 0 && yy < ra.height) {
                            const float ws = ra.weight_image[yy * ra.width + xx];
                            if (!(ws > 0)) continue;
                            const float ddx = (xx + 0.5f) - pos.pt.x, ddy = (yy + 0.5f) - pos.pt.y;
                            const float fw = filter_weight(sc.filter, ddx, ddy);
                            const float4 dp = *reinterpret_cast<const float4 *>(ra.d_render_image + 4 * (yy * ra.width + xx));
                            const float dotv = dp.x * color.x + dp.y * color.y + dp.z * color.z + dp.w * color.w;
                            const float d_weight = (dotv * ws - fw * dotv * (ws - fw)) / (ws * ws);
                            d_radius_acc += d_filter_weight_radius(sc.filter, ddx, ddy, d_weight);
                        }
                    }
                }
            }
        }
    }
    if (!BACKWARD) {
        __syncthreads();
        for (int i = tid; i < npix; i += RB) {
            const int lx = i % bins.tile_w, ly = i / bins.tile_w;
            const int x = tx * bins.tile_w + lx, y = ty * bins.tile_h + ly;
            if (x < ra.width && y < ra.height) {
                float *d = ra.render_image + 4 * (y * ra.width + x);
                const float *s = &s_pix[4 * i];
                if (s[0] != 0.f) atomicAdd(d + 0, s[0]);
                if (s[1] != 0.f) atomicAdd(d + 1, s[1]);
                if (s[2] != 0.f) atomicAdd(d + 2, s[2]);
                if (s[3] != 0.f) atomicAdd(d + 3, s[3]);
            }
        }
    } else {
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
__global__ void __launch_bounds__(RB) k_edge(SceneView sc, BinView bins, RenderArgs ra, BoundaryWork bw) {
    __shared__ Stage st;
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
    traverse<true, false>(sc, bins, tile, st, tr);
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
void launch_weight(const SceneView &sc, const RenderArgs &ra, cudaStream_t st) {
    const int n = ra.width * ra.height * ra.nsx * ra.nsy;
    DVG_LAUNCH(k_weight, dim3((n + 255) / 256), dim3(256), 0, st, sc, ra);
}

static int tile_rows_in(const BinView &bins, const RenderArgs &ra) {
    const int r0 = ra.row_begin / bins.tile_h;
    const int r1 = (ra.row_end + bins.tile_h - 1) / bins.tile_h;
    return r1 - r0;
}

void launch_render_forward(const SceneView &sc, const BinView &bins, const RenderArgs &ra, cudaStream_t st) {
    const int nblk = tile_rows_in(bins, ra) * bins.tiles_x;
    if (nblk <= 0) return;
    DVG_LAUNCH(k_render<false>, dim3(nblk), dim3(RB), 0, st, sc, bins, ra);
}

void launch_render_backward(const SceneView &sc, const BinView &bins, const RenderArgs &ra, cudaStream_t st) {
    const int nblk = tile_rows_in(bins, ra) * bins.tiles_x;
    if (nblk <= 0) return;
    DVG_LAUNCH(k_render<true>, dim3(nblk), dim3(RB), 0, st, sc, bins, ra);
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
