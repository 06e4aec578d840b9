// dvg_bvh.cu -- the reference's three BVH levels, built on the GPU in the reference's node order.
//
// The render kernels do not traverse these trees (they walk per-tile primitive lists, dvg_build.cu); the
// trees exist for the bit-exact BVH / indexing leg of the parity contract (dvg_scene_dump selectors 0-2,
// SURVEY 8c): same leaves, same sort keys, same bottom-up pairing, so that every node -- children, box,
// max_radius -- can be compared word for word with Scene::bvh_nodes / shape_groups_bvh_nodes / path_bvhs.
//
//   leaves      path BVH : one per segment, box of its control points, max control-point radius, sorted
//                          by box-centre y (scene.cpp:522-619);
//               group BVH: one per shape of the group, shapes_bbox + stroke radius, sorted by the 20-bit
//                          Morton code of the box centre (scene.cpp:632-645, 431-437);
//               scene BVH: one per group, root box of its group BVH moved to canvas space + the group's
//                          radius (scene.cpp:647-683), same Morton sort.
//   sort        one block per tree, bitonic sort of (key << 32 | leaf index) in global scratch.  The
//               reference's std::sort leaves the order of equal keys to the implementation; here ties
//               keep their input order.
//   pairing     scene.cpp:455-491: level by level, an odd node is carried as "leftover" to a later
//               level; the root lands at index 2n-2.  Child indices depend on n alone; the boxes of a
//               level are merged in parallel.
#include "dvg_internal.h"

namespace dvg {

namespace {

constexpr int BVH_B = 256;

DVG_D uint32_t expand_bits10(uint32_t x) {   // diffvg.h:111-126: bits 0..9 spread to the even positions
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 10; b++) r |= (x & (1u << b)) << b;
    return r;
}

// float -> uint32_t as the reference's host build does it (x86-64: cvttss2si into a 64-bit register, low
// word kept): negative values wrap instead of saturating to 0 as the GPU's conversion would.
DVG_D uint32_t host_f2u(float f) {
    if (!(fabsf(f) < 9.2e18f)) return 0u;   // x86 "integer indefinite" 0x8000000000000000 -> low word 0
    return (uint32_t)(long long)f;
}

DVG_D uint32_t morton2d(F2 p, int cw, int ch) {   // scene.cpp:431-437
    const float px = p.x / (float)cw, py = p.y / (float)ch;
    const uint32_t ix = host_f2u(px * 1023), iy = host_f2u(py * 1023);
    return (expand_bits10(ix) << 1u) | expand_bits10(iy);
}

DVG_D uint32_t float_key(float f) {   // order-preserving map of a float to an unsigned key
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

DVG_D int pow2_at_least(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// Sort keys[0, n) ascending (whole block; m = padded power of two, entries [n, m) hold ~0).
DVG_D void block_bitonic(unsigned long long *keys, int m) {
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// scene.cpp:455-491 for a tree whose n sorted leaves are already in nodes[0, n).
DVG_D void block_pair_levels(BvhNode *nodes, int n) {
    if (2 * n - 1 <= 1) return;
    __shared__ int s_beg, s_end, s_left, s_len, s_done;
    if (threadIdx.x == 0) { s_beg = 0; s_end = n; s_left = (n % 2 == 0) ? -1 : n - 1; s_done = 0; }
    __syncthreads();
    while (true) {
        if (threadIdx.x == 0) {
            int length = (s_end - s_beg) / 2;
            if ((s_end - s_beg) % 2 == 1 && s_left != -1 && s_left != s_end - 1) length += 1;
            s_len = length;
            if (!(s_end - s_beg >= 1 || s_left != -1)) s_done = 1;
        }
        __syncthreads();
        if (s_done) break;
        const int beg = s_beg, end = s_end, left = s_left, length = s_len;
        for (int i = threadIdx.x; i < length; i += blockDim.x) {
            BvhNode nd;
            nd.child0 = beg + 2 * i;
            nd.child1 = beg + 2 * i + 1;
            if (nd.child1 >= end) nd.child1 = left;   // only the last node of a level; consumes the leftover
            const BvhNode a = nodes[nd.child0], b = nodes[nd.child1];
            nd.box.x0 = rminf(a.box.x0, b.box.x0); nd.box.y0 = rminf(a.box.y0, b.box.y0);   // aabb.h:25-28
            nd.box.x1 = rmaxf(a.box.x1, b.box.x1); nd.box.y1 = rmaxf(a.box.y1, b.box.y1);
            nd.max_radius = a.max_radius < b.max_radius ? b.max_radius : a.max_radius;      // std::max
            nodes[end + i] = nd;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int l = left;
            if (length > 0 && beg + 2 * (length - 1) + 1 >= end) l = -1;
            if (length == 1 && l == -1) s_done = 1;
            else {
                s_beg = end; s_end = end + length;
                if (length % 2 == 1 && l == -1) l = s_end - 1;
            }
            s_left = l;
        }
        __syncthreads();
        if (s_done) break;
    }
}

// Leaf of one path segment (scene.cpp:527-600): box of its control points, max control-point radius.
DVG_D void segment_leaf(const int *topo, const float *P, const int *srec, int seg, int pid, Box &box, float &thick) {
    const float *p = P + srec[DVG_S_PARAM_OFF];
    const int np = srec[DVG_S_NUM_POINTS];
    const int n = topo[topo[DVG_H_OFF_NCP] + srec[DVG_S_NCP_OFF] + seg];
    const float *th = srec[DVG_S_THICK_OFF] >= 0 ? P + srec[DVG_S_THICK_OFF] : nullptr;
    const float sw = srec[DVG_S_WIDTH_OFF] >= 0 ? P[srec[DVG_S_WIDTH_OFF]] : 0.f;
    int ids[4];
    ids[0] = pid;
    if (n == 0) ids[1] = (pid + 1) % np;
    else if (n == 1) { ids[1] = pid + 1; ids[2] = (pid + 2) % np; }
    else { ids[1] = pid + 1; ids[2] = pid + 2; ids[3] = (pid + 3) % np; }
    box.x0 = box.y0 = INFINITY; box.x1 = box.y1 = -INFINITY;
    thick = sw;
    for (int k = 0; k < n + 2; k++) {
        const float x = p[2 * ids[k]], y = p[2 * ids[k] + 1];
        box.x0 = rminf(x, box.x0); box.y0 = rminf(y, box.y0);       // aabb.h:18-21
        box.x1 = rmaxf(x, box.x1); box.y1 = rmaxf(y, box.y1);
        if (th) thick = k == 0 ? th[ids[0]] : rmaxf(thick, th[ids[k]]);   // diffvg.h max(a, b): a > b ? a : b
    }
}

// One block per shape; non-path shapes have no tree.
__global__ void __launch_bounds__(BVH_B) k_bvh_paths(BuildView bv, BvhNode *nodes, unsigned long long *keys) {
    const int s = blockIdx.x;
    const int *srec = bv.topo + bv.topo[DVG_H_OFF_SHAPES] + s * DVG_SHAPE_REC_LEN;
    if (srec[DVG_S_TYPE] != DVG_SHAPE_PATH) return;
    const int n = srec[DVG_S_NUM_SEGS], off = 2 * srec[DVG_S_NCP_OFF];
    const int *pid = bv.seg_point_id + srec[DVG_S_NCP_OFF];   // first point of every segment (scene build)
    unsigned long long *k = keys + off;
    BvhNode *nd = nodes + off;
    const int m = pow2_at_least(n);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < n) {
            Box b; float th;
            segment_leaf(bv.topo, bv.params, srec, i, pid[i], b, th);
            key = ((unsigned long long)float_key(0.5f * (b.y0 + b.y1)) << 32) | (unsigned)i;   // scene.cpp:602-612
        }
        k[i] = key;
    }
    __syncthreads();
    block_bitonic(k, m);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int seg = (int)(k[i] & 0xffffffffu);
        BvhNode leaf;
        segment_leaf(bv.topo, bv.params, srec, seg, pid[seg], leaf.box, leaf.max_radius);
        leaf.child0 = seg; leaf.child1 = -(pid[seg] + 1);          // scene.cpp:613-618
        nd[i] = leaf;
    }
    __syncthreads();
    block_pair_levels(nd, n);
}

// One block per group (scene.cpp:632-645).
__global__ void __launch_bounds__(BVH_B) k_bvh_groups(BuildView bv, BvhNode *nodes, unsigned long long *keys) {
    const int g = blockIdx.x;
    const int *grec = bv.topo + bv.topo[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN;
    const int n = grec[DVG_G_NUM_SHAPES], off = 2 * grec[DVG_G_SHAPES_OFF];
    const int *ids = bv.topo + bv.topo[DVG_H_OFF_GSHAPES] + grec[DVG_G_SHAPES_OFF];
    unsigned long long *k = keys + off;
    BvhNode *nd = nodes + off;
    const int m = pow2_at_least(n);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < n) {
            const Box b = bv.shape_box[ids[i]];
            const F2 c = 0.5f * mk2(b.x0 + b.x1, b.y0 + b.y1);
            key = ((unsigned long long)morton2d(c, bv.canvas_w, bv.canvas_h) << 32) | (unsigned)i;
        }
        k[i] = key;
    }
    __syncthreads();
    block_bitonic(k, m);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int sh = ids[(int)(k[i] & 0xffffffffu)];
        const int *srec = bv.topo + bv.topo[DVG_H_OFF_SHAPES] + sh * DVG_SHAPE_REC_LEN;
        BvhNode leaf;
        leaf.child0 = sh; leaf.child1 = -1;
        leaf.box = bv.shape_box[sh];
        const float sw = srec[DVG_S_WIDTH_OFF] >= 0 ? bv.params[srec[DVG_S_WIDTH_OFF]] : 0.f;
        leaf.max_radius = grec[DVG_G_STROKE_TYPE] < 0 ? 0.f : sw;   // scene.cpp:638
        nd[i] = leaf;
    }
    __syncthreads();
    block_pair_levels(nd, n);
}

// One block (scene.cpp:647-683).  path_nodes / group_nodes: the trees built by the two kernels above.
__global__ void __launch_bounds__(BVH_B) k_bvh_scene(BuildView bv, const BvhNode *path_nodes, const BvhNode *group_nodes,
                                                     BvhNode *nodes, unsigned long long *keys) {
    const int n = bv.num_groups;
    const int m = pow2_at_least(n);
    // leaves first, in group order, parked behind the keys: m keys of 8 bytes, then n node records
    for (int g = threadIdx.x; g < n; g += blockDim.x) {
        const int *grec = bv.topo + bv.topo[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN;
        const int ns = grec[DVG_G_NUM_SHAPES];
        const int *ids = bv.topo + bv.topo[DVG_H_OFF_GSHAPES] + grec[DVG_G_SHAPES_OFF];
        float max_radius = 0.f;
        for (int i = 0; i < ns; i++) {
            const int *srec = bv.topo + bv.topo[DVG_H_OFF_SHAPES] + ids[i] * DVG_SHAPE_REC_LEN;
            float r = srec[DVG_S_WIDTH_OFF] >= 0 ? bv.params[srec[DVG_S_WIDTH_OFF]] : 0.f;
            if (srec[DVG_S_TYPE] == DVG_SHAPE_PATH && srec[DVG_S_THICK_OFF] >= 0)
                r = path_nodes[2 * srec[DVG_S_NCP_OFF]].max_radius;   // first leaf after the y-sort (scene.cpp:653-667)
            max_radius = i == 0 ? r : (max_radius < r ? r : max_radius);
        }
        const BvhNode root = group_nodes[2 * grec[DVG_G_SHAPES_OFF] + 2 * ns - 2];
        BvhNode leaf;
        leaf.child0 = g; leaf.child1 = -1;
        leaf.box = box_transform(bv.params + grec[DVG_G_XFORM_OFF], root.box);
        leaf.max_radius = grec[DVG_G_STROKE_TYPE] < 0 ? 0.f : max_radius;
        const F2 c = 0.5f * mk2(leaf.box.x0 + leaf.box.x1, leaf.box.y0 + leaf.box.y1);
        keys[g] = ((unsigned long long)morton2d(c, bv.canvas_w, bv.canvas_h) << 32) | (unsigned)g;
        reinterpret_cast<BvhNode *>(keys + m)[g] = leaf;
    }
    for (int i = n + threadIdx.x; i < m; i += blockDim.x) keys[i] = ~0ull;
    __syncthreads();
    block_bitonic(keys, m);
    const BvhNode *parked = reinterpret_cast<const BvhNode *>(keys + m);
    for (int i = threadIdx.x; i < n; i += blockDim.x) nodes[i] = parked[(int)(keys[i] & 0xffffffffu)];
    __syncthreads();
    block_pair_levels(nodes, n);
}

}  // namespace

// Scratch: keys need pow2(n) <= 2n entries per tree laid out at 2 * leaf offset; the scene tree parks its n
// leaves behind its keys.
size_t bvh_key_words(int total_segs, int num_insts, int num_groups) {
    const size_t a = 2 * (size_t)(total_segs > 0 ? total_segs : 1), b = 2 * (size_t)num_insts;
    const size_t c = 2 * (size_t)num_groups + 4 * (size_t)num_groups + 8;   // keys + parked leaves (32 B = 4 words each)
    size_t m = a > b ? a : b;
    return m > c ? m : c;
}

void launch_bvh_build(const BuildView &bv, BvhNode *path_nodes, BvhNode *group_nodes, BvhNode *scene_nodes,
                      unsigned long long *keys, cudaStream_t st) {
    DVG_LAUNCH(k_bvh_paths, dim3(bv.num_shapes), dim3(BVH_B), 0, st, bv, path_nodes, keys);
    DVG_LAUNCH(k_bvh_groups, dim3(bv.num_groups), dim3(BVH_B), 0, st, bv, group_nodes, keys);
    DVG_LAUNCH(k_bvh_scene, dim3(1), dim3(BVH_B), 0, st, bv, path_nodes, group_nodes, scene_nodes, keys);
}

}  // namespace dvg
