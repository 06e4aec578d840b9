// dvg_scene.cuh -- device-resident scene layout (SoA, 16-byte records) built every iteration
// from the flat `params` by the kernels in dvg_build.cu.
//
// The reference keeps an AoS object graph behind pointers (scene.h:18-67) and three levels of
// BVH.  Here the unit of work is a *primitive*: one path segment, or one whole
// circle/ellipse/rect, instanced per (group, shape-in-group).  Primitives are numbered in
// (group, shape-in-group, segment) order, so any ascending list of primitive ids is already
// in compositing order (diffvg.cpp:605-615 sorts fragments by group id).
#pragma once
#include "dvg_common.cuh"
#include "../../include/dvg_scene_format.h"

namespace dvg {

enum PrimType { PRIM_LINE = 0, PRIM_QUAD = 1, PRIM_CUBIC = 2, PRIM_CIRCLE = 3, PRIM_ELLIPSE = 4, PRIM_RECT = 5 };

// prim flags (upper bits of PrimMeta.type_flags)
#define DVG_PF_TYPE_MASK 0xF
#define DVG_PF_SINGLE 0x10    // leaf == root of its path BVH: the reference performs no box test
#define DVG_PF_THICK 0x20     // per-point thickness
#define DVG_PF_APPROX 0x40    // use_distance_approx
#define DVG_PF_FIRST 0x80     // first primitive of its shape instance
#define DVG_PF_GFIRST 0x100   // first primitive of its group
#define DVG_PF_YMONO 0x400    // filled cubic whose y(t) is strictly monotone on [0, 1] by a margin: prim_wcert is valid
#define DVG_PF_YUP 0x800      // ... and increasing
#define DVG_PF_TIGHT 0x200    // stroke-only curved primitive in an untransformed group: tiles are binned against its
                              // polyline bracket (prim_cap) instead of its bounding box (dvg_build.cu k_bin)

// Polyline bracket of a curved stroke primitive: the curve is cut into DVG_CAP_N pieces; piece i lies
// within `dev_i` of its chord A_i -> A_i + d_i and every chord point has a curve point within dev_i.
// With D = distance(pt, chord_i):
//   D > max_radius + dev_i + margin for ALL i  => no curve point is within the stroke radius: the
//       quintic / cubic root solve of within_distance.h:63-272 would return false (always exact);
//   D < min_radius - dev_i - margin for SOME i => a curve point is within the stroke radius, i.e. the
//       reference's closest-point solve returns true (see DESIGN.md "polyline bracket" for the
//       measured agreement);
//   otherwise the exact solve runs.  8 floats per piece: A.xy, d.xy, 1/|d|^2, R_out^2, R_in^2, pad.
#define DVG_CAP_N 8
#define DVG_CAP_F4 (2 * DVG_CAP_N)

struct PrimMeta {
    int type_flags;  // PrimType | flags
    int inst;        // shape instance (index into inst_* arrays)
    int point_id;    // path: index of the segment's first point
    int base_id;     // path: segment index (base_point_id)
};

// One record per shape group.  96 bytes... read through uniform (broadcast) loads.
struct GroupInfo {
    Box scene_box;   // leaf box of the reference scene BVH (canvas space), scene.cpp:672-676
    Box local_box;   // root box of the reference group BVH (local space), diffvg.cpp:42
    float scene_r;   // leaf max_radius of the scene BVH, scene.cpp:650-681
    int flags;       // DVG_GF_*
    int fill_type, fill_off, fill_stops;
    int stroke_type, stroke_off, stroke_stops;
    int num_shapes, inst_begin, prim_begin, prim_end;
    float c2s[9];    // canvas_to_shape = inverse(shape_to_canvas), shape.h:122-124
    float s2c[9];
    int xform_off;
    int pad;
};
#define DVG_GF_EVEN_ODD 1
#define DVG_GF_IDENTITY 2   // shape_to_canvas is exactly the identity: xform_pt is exact, skip it
#define DVG_GF_AFFINE 4     // last row is exactly (0,0,1)

// Per shape *instance* (a shape referenced by a group).  In a batch (SceneView::batch scenes of one topology laid out
// back to back, see SceneView) `group` and `prim_begin` are batch-wide indices, `shape` stays the id inside its scene (it
// indexes the shared topology) and `scene` says which scene: scene * num_shapes + shape, scene * num_params + offset, ...
// address that scene's slice of the per-shape tables and of the parameter / gradient buffers.
struct InstInfo {
    Box box;        // shapes_bbox[shape], scene.cpp:499-629
    float r;        // leaf radius in the group BVH: stroke_width if the group strokes else 0, scene.cpp:638
    int group;
    int shape;
    int prim_begin;
    int scene;
};

struct PrimQuintic;
struct PrimWindCert;

// Everything the kernels need, passed by value.
//
// BATCH: `batch` scenes that share one topology (the stroke scenes of apps/generative_models/rendering.py:170-307: same
// shape and group lists, different parameters and seeds) live in ONE set of tables, scene b occupying slice b of every
// array: parameters [b * num_params, ...), shapes [b * num_shapes, ...), instances, groups, primitives, segment tables.
// The counts below are PER SCENE.  Primitive, instance and group ids that travel through bins, pair queues and fragment
// records are batch-wide (slice offset included), colour / transform offsets in GroupInfo are absolute, so the traversal
// and the exact-test kernels never need to know about scenes; only the code that maps work items to pixels (one tile list
// per (scene, tile)), draws random numbers (per-scene seed, sample index local to the scene) or reads the shared
// topology does.  batch == 1 is the plain single-scene case: every slice offset is 0.
struct SceneView {
    int canvas_w, canvas_h;
    int num_shapes, num_groups, num_insts, num_prims;
    int batch, num_params, total_segs;
    Filter filter;
    int filter_radius_off;
    const int *topo;       // device copy of the topology blob
    const float *params;   // device copy of the flat parameters
    // primitives
    const F4 *prim_p01;    // (p0.x,p0.y,p1.x,p1.y) | circle (cx,cy,r,0) | ellipse (cx,cy,rx,ry) | rect (min,max)
    const F4 *prim_p23;
    const F4 *prim_rad;    // stroke radius at each control point (thickness or stroke_width broadcast)
    const Box *prim_box;   // local-space leaf box, scene.cpp:536-585
    const float *prim_thick;  // leaf max_radius, scene.cpp:546,569,597
    const PrimMeta *prim_meta;
    const Box *prim_cbox;  // canvas-space conservative bound of where this primitive can matter (binning only)
    const Box *prim_cbox_pf;  // same for the prefiltering path (binning only)
    const F4 *prim_cap;    // DVG_CAP_F4 float4 per primitive: conservative stroke-reject capsules (dvg_geom.cuh)
    const PrimQuintic *prim_quint;
    const int *shape_guide;          // [DVG_CDF_GUIDE + 1] guide table of shape_cdf (scene 0 of a batch), dvg_boundary.cuh cdf_sample_guided
    const PrimWindCert *prim_wcert;  // filled cubic segments flagged DVG_PF_YMONO: what the classifier needs to answer their winding test itself (dvg_geom.cuh)   // cubic segments: the sample-independent part of the closest-point quintic (dvg_geom.cuh)
    const InstInfo *insts;
    const GroupInfo *groups;
    // boundary sampling tables (scene.cpp:207-333)
    const float *shapes_length;   // [num_shapes]
    const float *shape_cdf;       // [num_insts]
    const float *shape_pmf;       // [num_insts]
    const float *seg_cdf;         // [total_segs], indexed by shape.ncp_off + i
    const float *seg_pmf;
    const int *seg_point_id;
    int *error_flag;
};

// Tile bins for one render configuration: tiles_x * tiles_y tiles per scene, scene b's tiles at b * tiles_x * tiles_y.
struct BinView {
    int tile_w, tile_h, tiles_x, tiles_y;
    int batch;
    const int *offsets;   // [tiles+1]
    const int *items;     // ascending primitive ids per tile
};
DVG_HD int bin_scene_tiles(const BinView &b) { return b.tiles_x * b.tiles_y; }
DVG_HD int bin_total_tiles(const BinView &b) { return b.tiles_x * b.tiles_y * b.batch; }

// RenderArgs.flags: low bits = DVG_BWD_* of the C ABI; internal bits from 16 up
#define DVG_RF_FAST_ACCEPT (1u << 16)

// Arguments of the render / boundary kernels.
struct RenderArgs {
    int width, height, nsx, nsy;
    uint64_t seed;
    const uint64_t *seeds;         // batch: one seed per scene (device), else null; images below are then [batch, H, W, C]
    int use_prefiltering;
    int row_begin, row_end;        // pixel rows owned by this call (sharding); whole image by default
    uint32_t flags;
    const float *background;       // [H,W,4] or null
    float *weight_image;           // [H,W]
    float *render_image;           // [H,W,4] forward output (accumulated)
    const float *d_render_image;   // [H,W,4] backward input
    float *d_params;
    float *d_params_rep;           // wavefront path: grad_reps private copies of the gradient buffer (see dvg_wave.cu), summed at the end
    int grad_reps, num_params;
    float *d_background;
    float *d_translation;
    float *debug_out;              // [n,4] per boundary sample (contrib, hit bits, normal) -- debug builds of the tests only
};

// The arguments as scene `b` of a batch sees them: its own seed and its own slice of every image.
DVG_HD RenderArgs args_of_scene(const RenderArgs &ra, int b) {
    if (!ra.seeds) return ra;
    RenderArgs r = ra;
    const size_t px = (size_t)ra.width * ra.height * (size_t)b;
    r.seed = ra.seeds[b];
    if (ra.background) r.background = ra.background + 4 * px;
    if (ra.weight_image) r.weight_image = ra.weight_image + px;
    if (ra.render_image) r.render_image = ra.render_image + 4 * px;
    if (ra.d_render_image) r.d_render_image = ra.d_render_image + 4 * px;
    if (ra.d_background) r.d_background = ra.d_background + 4 * px;
    return r;
}

}  // namespace dvg
