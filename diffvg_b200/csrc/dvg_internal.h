// dvg_internal.h -- declarations shared by the .cu translation units of libdiffvg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dvg_common.cuh"
#include "dvg_scene.cuh"
#include "dvg_geom.cuh"
#include "dvg_color.cuh"
#include "dvg_boundary.cuh"
#include "dvg_trace.cuh"
#include "dvg_distance.cuh"
#include "dvg_buildfn.cuh"

namespace dvg {

extern long long g_launch_count;  // kernels launched since load (dvg_kernel_launch_count)

// Optional per-kernel timing with CUDA events on the launching stream (dvg_profile_enable):
// bench.py uses it to time the dominant kernel live; off by default (zero overhead).
extern bool g_profile_on;
void prof_begin(const char *name, cudaStream_t st);
void prof_end(cudaStream_t st);

#define DVG_LAUNCH(kernel, grid, block, smem, stream, ...)            \
    do {                                                              \
        if (::dvg::g_profile_on) ::dvg::prof_begin(#kernel, stream);  \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);   \
        if (::dvg::g_profile_on) ::dvg::prof_end(stream);             \
        ::dvg::g_launch_count++;                                      \
    } while (0)

// the same with the reported name given apart (template instantiations with several arguments do not survive as ONE macro argument)
#define DVG_LAUNCH_AS(name, kernel, grid, block, smem, stream, ...)   \
    do {                                                              \
        if (::dvg::g_profile_on) ::dvg::prof_begin(name, stream);     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);   \
        if (::dvg::g_profile_on) ::dvg::prof_end(stream);             \
        ::dvg::g_launch_count++;                                      \
    } while (0)

struct BinBuild {
    int width, height, tile_w, tile_h, tiles_x, tiles_y;
    int batch;      // scenes of one topology back to back (SceneView): tiles_x * tiles_y tiles each; row ranges only with batch 1
    int prefilter;  // 1: bin with prim_cbox_pf (SDF-prefiltering candidate regions)
    int tile_row0, tile_row1;  // tile rows to bin (the others get empty lists)
    // two-level binning: supertiles of `super` x `super` tiles get a candidate list first (fixed stride of num_prims
    // entries per supertile, so no scan / read-back), tiles then test only their supertile's list.  super = 0: off
    int super, stiles_x, stiles_y;
    int *s_counts, *s_items;
    int flat;       // 1: test every primitive (few primitives per group); 0: groups first, then their primitives
    int *scan_ws;  // launch_scan workspace (DVG_SCAN_WS_BLOCKS + 1 ints, zeroed at allocation), or null
    int *counts;   // [tiles]
    int *offsets;  // [tiles+1]
    int *items;    // [capacity]
};

struct BoundaryWork {
    int num_samples;      // = sample_end - sample_begin
    int sample_begin;
    int *keys;            // [num_samples] tile id or -1
    int *tile_counts;     // [tiles]
    int *tile_offsets;    // [tiles+1]
    int *tile_fill;       // [tiles]
    int *blk_counts;      // [tiles]
    int *blk_offsets;     // [tiles+1]
    int *sorted_idx;      // [num_samples]
    int max_blocks;
    BoundarySample *samples;           // [num_samples] in TILE order (wavefront path: made once, sorted, read twice), or null
    BoundarySample *samples_unsorted;  // [num_samples] by index - sample_begin: staging of the counting sort, or null
    int *item_tile;           // [max_blocks] tile of every boundary item, or null
    int *scan_ws;             // launch_scan workspace, or null
};

// Wavefront passes (dvg_wave.cu): queues and result words in global memory.
struct WavePair { float x, y; int prim; unsigned ref; };   // shape-local sample position, PrimType << 28 | primitive, word << 5 | candidate
struct WaveView {
    unsigned *hit;        // one word per (evaluation, chunk of 32 candidates): bit k = stroke test of candidate k hit
    unsigned *wind;       // four words per (evaluation, chunk): 4-bit signed winding per candidate; null without fills
    WavePair *pairs_s, *pairs_f;   // exact stroke tests / winding tests still to run
    int cap_s, cap_f;
    int *counters;        // [0] stroke pairs wanted, [1] fill pairs wanted (may exceed the capacity: the surplus is answered in
                          // place by the retry form of W1)
    int *tile_choff;      // [tiles+1] exclusive scan of chunks per tile
    int *edge_choff;      // [tiles+1] exclusive scan of boundary items * chunks per tile
};

// SDF output (OutputType.sdf): device pointers, NULL = absent
struct SdfArgs {
    float *sdf;                   // [H*W] or [num_eval]
    const float *d_sdf;           // backward input, same shape
    const float *eval_positions;  // [2*num_eval] or null
    int num_eval;
};

// Node of the reference-topology BVHs (scene.h:12-16), 28 bytes; built by dvg_bvh.cu for dvg_scene_dump.
struct BvhNode { int child0, child1; Box box; float max_radius; };
size_t bvh_key_words(int total_segs, int num_insts, int num_groups);
void launch_bvh_build(const BuildView &bv, BvhNode *path_nodes, BvhNode *group_nodes, BvhNode *scene_nodes,
                      unsigned long long *keys, cudaStream_t st);

void launch_debug_prim_tests(const SceneView &sc, const BinView &bins, const RenderArgs &ra, int x, int y, int *out, float *pos, cudaStream_t st);
void launch_peak_probe(int which, float *out, int iters, cudaStream_t st);
void launch_build(const BuildView &bv, cudaStream_t st);
void launch_bin_coarse(const BuildView &bv, const BinBuild &bb, cudaStream_t st);
void launch_bin_count(const BuildView &bv, const BinBuild &bb, cudaStream_t st);
void launch_bin_fill(const BuildView &bv, const BinBuild &bb, cudaStream_t st);
#define DVG_SCAN_WS_BLOCKS 1024   // launch_scan workspace: this many block totals + one ticket (ints, zeroed once)
void launch_scan(const int *in, int *out, int n, int *ws, cudaStream_t st);
void launch_tile_row_costs(const int *offsets, int tiles_x, int tiles_y, float *out, cudaStream_t st);

void launch_weight(const SceneView &sc, const RenderArgs &ra, int row_begin, int row_end, cudaStream_t st);
void launch_render_pf(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const unsigned *wind, const unsigned *relevant,
                      const int *tile_choff, const PfCache &pc, bool backward, cudaStream_t st);
void launch_pf_backward_cached(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const PfCache &pc, cudaStream_t st);
int64_t pf_launch_threads(const BinView &bins, const RenderArgs &ra);
void launch_sdf(const SceneView &sc, const RenderArgs &ra, const SdfArgs &sa, bool backward, cudaStream_t st);
void launch_boundary_sort(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw, cudaStream_t st);

void launch_wave_tile_chunks(const int *bin_offsets, int *nch, int *choff, int *max_nch, int ntiles, int *scan_ws, cudaStream_t st);
int wave_pixel_items(const BinView &bins, const RenderArgs &ra);
void launch_wave_reduce_grads(const RenderArgs &ra, cudaStream_t st);
int wave_items_per_tile(const BinView &bins, const RenderArgs &ra);
int wave_edge_samples_per_item();
void launch_wave_classify_px(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const WaveView &wv, cudaStream_t st);
void launch_wave_solve(const SceneView &sc, const WaveView &wv, bool strokes, bool fills, cudaStream_t st);
void launch_wave_retry_px(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const WaveView &wv, cudaStream_t st);
void launch_wave_retry_edge(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw,
                            const WaveView &wv, cudaStream_t st);
extern int g_num_sms;
void launch_wave_composite_px(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const WaveView &wv, bool backward,
                              cudaStream_t st);
void launch_wave_boundary_sort(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw,
                               const WaveView &wv, int *edge_chunks, cudaStream_t st);
void launch_wave_classify_edge(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw,
                               const WaveView &wv, cudaStream_t st);
void launch_wave_composite_edge(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw,
                                const WaveView &wv, cudaStream_t st);

}  // namespace dvg
