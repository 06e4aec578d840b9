// dvg_debug.cu -- test support (sample-level parity debugging; no reference counterpart).
//
// k_debug_prim_tests evaluates, for every sample of ONE pixel, the two exact per-primitive predicates on EVERY primitive
// of the scene -- the stroke test (within_distance.h) and the winding contribution (winding_number.h) -- with no
// culling, queues or result words in between, and reports which primitives the pixel's tile bin holds.  Compared on the
// host with the same arithmetic compiled by g++ (tests/host_emul) it separates "the arithmetic differs on the device"
// from "the traversal dropped a test".
#include "dvg_internal.h"

namespace dvg {

// out[s * num_prims + e] = stroke hit (bit 0) | group strokes (bit 1) | group fills (bit 2) | primitive is in the tile's
// bin (bit 3) | (winding & 0xff) << 8;  pos[2 * s] = canvas-space sample position
__global__ void k_debug_prim_tests(SceneView sc, BinView bins, RenderArgs ra, int x, int y, int *out, float *pos) {
    const int s = blockIdx.x;
    const int sx = s % ra.nsx, sy = s / ra.nsx;
    const int idx = ((y * ra.width + x) * ra.nsy + sy) * ra.nsx + sx;
    F2 pt, cpt;
    sample_position(sc.canvas_w, sc.canvas_h, ra.width, ra.height, ra.nsx, ra.nsy, ra.seed, ra.use_prefiltering != 0, x, y, sx, sy, idx, pt, cpt);
    if (threadIdx.x == 0) { pos[2 * s] = cpt.x; pos[2 * s + 1] = cpt.y; }
    const int tile = (y / bins.tile_h) * bins.tiles_x + x / bins.tile_w;
    const int beg = bins.offsets[tile], end = bins.offsets[tile + 1];
    for (int e = threadIdx.x; e < sc.num_prims; e += blockDim.x) {
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
        for (int i = beg; i < end; i++) if (bins.items[i] == e) { r |= 8; break; }
        out[(size_t)s * sc.num_prims + e] = r;
    }
}

void launch_debug_prim_tests(const SceneView &sc, const BinView &bins, const RenderArgs &ra, int x, int y, int *out, float *pos, cudaStream_t st) {
    DVG_LAUNCH(k_debug_prim_tests, dim3(ra.nsx * ra.nsy), dim3(128), 0, st, sc, bins, ra, x, y, out, pos);
}

}  // namespace dvg
