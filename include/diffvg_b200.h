/*
 * diffvg_b200.h -- C ABI of the B200-native differentiable vector-graphics rasteriser.
 *
 * This is the drop-in boundary for the hot path BASELINE.json:north_star names.  Each
 * entry point cites the reference interface it replaces (paths relative to the reference
 * tree).  Plain pointers and sizes only; no C++/torch types.  Every function returns 0
 * on success or a non-zero DvgStatus; dvg_last_error() gives the message (thread-local).
 * Nothing here ever calls exit() (the reference does on CUDA errors, cuda_utils.h:11-16).
 *
 * Threading / streams: calls are asynchronous on the cudaStream_t passed as `stream`
 * (0 = legacy default stream); no device-wide synchronisation is performed (the reference
 * ends every render() with cudaDeviceSynchronize, diffvg.cpp:1639-1641).  A DvgScene may
 * be used from one thread at a time; distinct scenes are independent.
 */
#ifndef DIFFVG_B200_H
#define DIFFVG_B200_H

#include <stdint.h>
#include "dvg_scene_format.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct DvgScene DvgScene;

typedef enum DvgStatus {
    DVG_OK = 0,
    DVG_ERR_INVALID = 1,  /* malformed topo / arguments                     */
    DVG_ERR_CUDA = 2,     /* a CUDA runtime call failed                      */
    DVG_ERR_SCENE = 3,    /* degenerate scene (scene.cpp:231-240)            */
    DVG_ERR_UNSUPPORTED = 4 /* e.g. stroked ellipse (within_distance.h:342-345 asserts) */
} DvgStatus;

/* flags for dvg_render_backward */
#define DVG_BWD_SKIP_XFORM_GRAD 1u   /* do not accumulate d(shape_to_canvas) (saves 9 scatters / boundary sample) */
#define DVG_BWD_ACCUMULATE      2u   /* add into d_params instead of overwriting it                               */
#define DVG_BWD_SKIP_FILTER_GRAD 4u   /* do not accumulate d(filter.radius) (a 3x3-pixel gather per sample at radius 0.5): its entry of d_params stays 0 */

/* Version of this ABI; bumped on incompatible change. */
int dvg_abi_version(void);

/* Thread-local message of the last failing call on this thread. */
const char *dvg_last_error(void);

/*
 * Create a scene from its topology.  Replaces the structural half of
 * `diffvg.Scene(canvas_w, canvas_h, shapes, shape_groups, filter, use_gpu, gpu_index)`
 * (scene.cpp:919-998, allocate_buffers 686-917) and of the pybind11 object graph built in
 * render_pytorch.py:206-363.  `topo` is a HOST array in dvg_scene_format.h layout; it is
 * copied.  `device` is the CUDA ordinal the scene lives on (reference: gpu_index).
 */
int dvg_scene_create(const int32_t *topo, int64_t topo_len, int device, DvgScene **out_scene);

/*
 * Upload the continuous parameters and (re)build everything derived from them on the
 * GPU: segment table, shape lengths, shape/path sampling CDFs and PMFs, bounding boxes,
 * reference-topology BVH levels and the tile bins the kernels traverse.  Replaces
 * copy_and_init_shapes / copy_and_init_shape_groups / compute_shape_length /
 * build_shape_cdfs / build_path_cdfs / compute_bounding_boxes (scene.cpp:45-684), which
 * the reference runs single-threaded on the host in every forward.
 * `params` may be a host pointer (params_on_device = 0: one cudaMemcpyAsync H2D) or a
 * device pointer on the scene's device (params_on_device = 1: D2D copy).
 */
int dvg_scene_set_params(DvgScene *scene, const float *params, int64_t num_params,
                         int params_on_device, void *stream);

/*
 * Forward render.  Replaces `diffvg.render(scene, background, render_image, render_sdf,
 * width, height, nsx, nsy, seed, 0,0,0,0, use_prefiltering, eval_positions, n)`
 * (diffvg.cpp:1477-1492 with the d_* pointers null; kernels 1115-1312).
 * All image pointers are DEVICE memory on the scene's device; NULL = not requested.
 *   background     float[H*W*4] or NULL
 *   render_image   float[H*W*4] (overwritten)
 *   render_sdf     float[H*W] or float[num_eval_positions] (overwritten)
 *   eval_positions float[2*num_eval_positions] or NULL (sdf only, render_kernel 1181-1186)
 */
int dvg_render_forward(DvgScene *scene, const float *background, float *render_image, float *render_sdf,
                       int width, int height, int num_samples_x, int num_samples_y, uint64_t seed,
                       int use_prefiltering, const float *eval_positions, int num_eval_positions,
                       void *stream);

/*
 * Backward render.  Replaces the second `diffvg.render(...)` call with d_render_image /
 * d_render_sdf set (render_pytorch.py:692-707; diffvg.cpp:1193-1272 interior term,
 * 1558-1626 boundary term) AND the per-field gradient read-back through
 * Scene::get_d_shape / get_d_shape_group / get_d_filter_radius (scene.cpp:1024-1034,
 * render_pytorch.py:713-866): gradients arrive as ONE flat array with the layout of
 * `params`.
 *   d_render_image float[H*W*4] or NULL;  d_render_sdf float[H*W] / float[n_eval] or NULL
 *   d_params       float[num_params]  (device; overwritten unless DVG_BWD_ACCUMULATE)
 *   d_background   float[H*W*4] or NULL (overwritten)
 *   d_translation  float[H*W*2] or NULL (overwritten; RenderFunction.render_grad)
 */
int dvg_render_backward(DvgScene *scene, const float *background,
                        const float *d_render_image, const float *d_render_sdf,
                        int width, int height, int num_samples_x, int num_samples_y, uint64_t seed,
                        int use_prefiltering, const float *eval_positions, int num_eval_positions,
                        float *d_params, float *d_background, float *d_translation,
                        uint32_t flags, void *stream);

/*
 * BATCHED SCENES.  Replaces the per-sample Python loops of the reference's batched front end
 * (apps/generative_models/rendering.py:170-237 line_render, :239-307 bezier_render, :101-167 stroke2diffvg): `batch`
 * scenes that share ONE topology (same shapes, groups and segment counts; only the continuous parameters and the seeds
 * differ) are built, rendered and differentiated by one set of kernel launches, where the reference constructs a
 * diffvg.Scene and calls diffvg.render twice per sample.
 *   dvg_scene_create_batch   as dvg_scene_create; `batch` >= 1 (1 = dvg_scene_create)
 *   dvg_scene_set_params     takes float[batch * num_params]: scene b's parameters at b * num_params (same layout each).
 *                            The pixel-filter radius is scene 0's.
 *   images                   float[batch * H * W * 4], scene after scene (background, render_image, d_render_image,
 *                            d_background)
 *   seeds                    HOST uint64[batch]: scene b renders exactly what a single scene with seed seeds[b] renders
 *   d_params                 float[batch * num_params]
 * Colour images of the sampled path only: no prefiltering, SDF output, eval_positions, d_translation or row ranges
 * (DVG_ERR_UNSUPPORTED).  A degenerate scene anywhere in the batch fails the call (DVG_ERR_SCENE).
 */
int dvg_scene_create_batch(const int32_t *topo, int64_t topo_len, int device, int batch, DvgScene **out_scene);
int dvg_render_forward_batch(DvgScene *scene, const float *background, float *render_image,
                             int width, int height, int num_samples_x, int num_samples_y, const uint64_t *seeds,
                             void *stream);
int dvg_render_backward_batch(DvgScene *scene, const float *background, const float *d_render_image,
                              int width, int height, int num_samples_x, int num_samples_y, const uint64_t *seeds,
                              float *d_params, float *d_background, uint32_t flags, void *stream);

/* Destroy the scene and its device buffers (Scene::~Scene, scene.cpp:1000-1022). */
int dvg_scene_destroy(DvgScene *scene);

/*
 * Debug export for the bit-exact BVH / CDF / indexing checks (the reference
 * keeps these private in Scene).  Synchronises `stream`.  `what`:
 *   0 scene BVH nodes, 1 group BVH (index = group), 2 path BVH (index = shape): records of 7
 *     32-bit words {child0, child1, box.min.x, box.min.y, box.max.x, box.max.y, max_radius}
 *     in the reference's node order (scene.cpp:439-494);
 *   3 shapes_length, 4 sample_shapes_cdf, 5 sample_shapes_pmf,
 *   6 path_length_cdf (index = shape), 7 path_length_pmf, 8 path_point_id_map,
 *   9 sample_shape_id, 10 sample_group_id.
 * Writes raw 32-bit words to the HOST buffer `out`; returns the word count or -1.
 */
int64_t dvg_scene_dump(DvgScene *scene, int what, int index, uint32_t *out, int64_t cap, void *stream);

/*
 * Number of kernels this library launched on behalf of the calling thread's process since
 * load (monotonic); bench.py uses the delta for `gpu_launches`.
 */
int64_t dvg_kernel_launch_count(void);

/*
 * Measurement support (no reference counterpart; the reference only prints wall-clock times,
 * render_pytorch.py:8-12).  dvg_profile_enable(1) makes every kernel launch record a CUDA-event
 * pair on its stream; dvg_profile_report synchronises, writes "kernel,launches,total_ms\n" lines
 * for the launches since the last report and returns the text length (-1 on error).
 * dvg_measure_peak runs an FMA-chain probe (which: 0 = FP32, 1 = FP64) and returns the achieved
 * TFLOP/s: the roofline denominator for this CUDA-core-bound path.
 */
int dvg_profile_enable(int on);
/* Opt-in speed mode (off by default; process-wide).  A sample that the polyline bracket of a curved
 * stroke proves to be inside the stroke radius is accepted without running the reference's quintic
 * closest-point solve.  Geometrically that answer is right, but the reference's solver has false
 * negatives at near-tangent quintics (~6e-6 of the samples at the painterly config), so
 * with this on roughly 20-30 pixels per 512x512 image differ from the reference by one sample's weight. */
int dvg_set_fast_stroke_accept(int on);
/* Test support: when non-NULL, the boundary pass also writes (contrib, hit bits, normal.xy) per boundary
 * sample index into this DEVICE buffer of 4*W*H*spp floats (sample-level parity debugging). */
int dvg_debug_set_boundary_dump(float *device_buf);
/* Test support: cap the exact-test pair queues at `pair_capacity` records (0 = off: sized from earlier passes), which
 * forces the answer-in-place path of the classifier, and the boundary pass at `edge_pass_samples` samples per sub-pass
 * (0 = off: as many as the result words allow).  Process-wide. */
int dvg_debug_set_limits(int64_t pair_capacity, int64_t edge_pass_samples);
/* Test support, bit mask: 1 = the prefiltered path runs its winding tests inline in the render kernel (what it does anyway
 * for scenes without fills and for renders beyond the 27-bit word index) instead of through the winding pre-pass; 2 = its
 * backward pass walks the candidate lists of every sample again instead of differentiating from the fragment records the
 * forward pass cached.  Same results either way (gradients up to the order of the atomic sums); process-wide. */
int dvg_debug_set_prefilter_inline(int on);
/* Test support: for every sample of pixel (x, y) and EVERY primitive of the scene (no culling), the exact stroke test and
 * winding contribution as the device computes them: out_host[s * num_prims + e] = hit (bit 0) | group strokes (bit 1) |
 * group fills (bit 2) | primitive is in the pixel's tile bin (bit 3) | (winding & 0xff) << 8; pos_host[2 s .. 2 s + 1] =
 * canvas-space sample position.  HOST buffers of nsx*nsy*num_prims ints and 2*nsx*nsy floats.  Synchronises `stream`. */
int dvg_debug_prim_tests(DvgScene *scene, int width, int height, int num_samples_x, int num_samples_y, uint64_t seed,
                         int x, int y, int32_t *out_host, float *pos_host, void *stream);
int64_t dvg_profile_report(char *buf, int64_t cap);
int dvg_measure_peak(int which, int device, double *tflops);

/*
 * Sharded variants for one-process-per-GPU runs.  Pixel samples
 * [sample_begin, sample_end) of the global index space idx = ((y*W+x)*nsy+sy)*nsx+sx and the
 * same range of boundary-sample indices are processed; RNG streams depend on the global
 * idx only (pcg.h:32-40), so the union over ranks reproduces the single-GPU sample set.
 * Row ranges are in pixels: rows [row_begin, row_end) of the image.
 * The caller all-reduces d_params (NCCL) afterwards.
 */
int dvg_render_forward_rows(DvgScene *scene, const float *background, float *render_image,
                            int width, int height, int num_samples_x, int num_samples_y, uint64_t seed,
                            int use_prefiltering, int row_begin, int row_end, void *stream);
int dvg_render_backward_rows(DvgScene *scene, const float *background, const float *d_render_image,
                             int width, int height, int num_samples_x, int num_samples_y, uint64_t seed,
                             int use_prefiltering, int row_begin, int row_end,
                             float *d_params, float *d_background, uint32_t flags, void *stream);

/*
 * Cost profile for a BALANCED row partition (no reference counterpart: scene.cpp / diffvg.cpp are single-device).  Bins
 * the whole image for the scene's current parameters and writes, per tile row (tile_h_out pixel rows each, top to
 * bottom), the number of (tile, candidate primitive) entries of that row to the HOST array `out_costs` (`cap` >= number
 * of tile rows = ceil(height / tile height)).  The render passes cost about that much per row plus a constant per tile;
 * diffvg_b200/sharded.py cuts the bands so that every rank gets the same share.  Synchronises `stream`.
 */
int dvg_scene_row_costs(DvgScene *scene, int width, int height, int num_samples_x, int num_samples_y,
                        int use_prefiltering, float *out_costs, int cap, int *tile_h_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFVG_B200_H */
