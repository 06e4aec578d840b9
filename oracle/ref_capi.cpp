// ref_capi.cpp -- TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// A thin C-ABI driver around the UNMODIFIED reference renderer (compiled from the
// sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/).  It accepts the same packed (topo, params) scene the product C ABI
// takes (include/dvg_scene_format.h), builds the reference's own
// Shape/ShapeGroup/Scene objects (shape.h:9-169, scene.cpp:919-998) and calls the
// reference's `render` (diffvg.cpp:1477-1649).  Gradients are read back out of
// Scene::d_shapes / d_shape_groups / d_filter into the flat d_params layout, the
// same way render_pytorch.py:713-866 reads them field by field.
//
// No reference code is copied here: this file only *calls* it.
#include "scene.h"
#include "shape.h"
#include "color.h"
#include "filter.h"
#include "ptr.h"
#include "../include/dvg_scene_format.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// diffvg.cpp:1477 (non-static, external linkage)
void render(std::shared_ptr<Scene> scene,
            ptr<float> background_image,
            ptr<float> render_image,
            ptr<float> render_sdf,
            int width,
            int height,
            int num_samples_x,
            int num_samples_y,
            uint64_t seed,
            ptr<float> d_background_image,
            ptr<float> d_render_image,
            ptr<float> d_render_sdf,
            ptr<float> d_translation,
            bool use_prefiltering,
            ptr<float> eval_positions,
            int num_eval_positions);

namespace {

thread_local std::string g_err;

struct RefScene {
    // storage that must outlive Scene construction (Scene deep-copies, scene.cpp:959-966)
    std::vector<Circle> circles;
    std::vector<Ellipse> ellipses;
    std::vector<Path> paths;
    std::vector<Rect> rects;
    std::vector<Shape> shapes;
    std::vector<Constant> constants;
    std::vector<LinearGradient> linears;
    std::vector<RadialGradient> radials;
    std::vector<ShapeGroup> groups;
    std::vector<float> params;  // mutable copy (reference takes non-const pointers)
    std::vector<int> ints;      // mutable copy of topo
    std::shared_ptr<Scene> scene;
};

void *color_ptr(RefScene &rs, int type, int off, int stops) {
    float *p = rs.params.data();
    switch (type) {
        case DVG_COLOR_NONE:
            return nullptr;
        case DVG_COLOR_CONSTANT:
            rs.constants.push_back(Constant{Vector4f{p[off], p[off + 1], p[off + 2], p[off + 3]}});
            return &rs.constants.back();
        case DVG_COLOR_LINEAR:
            rs.linears.push_back(LinearGradient(Vector2f{p[off], p[off + 1]}, Vector2f{p[off + 2], p[off + 3]},
                                                stops, ptr<float>(p + off + 4), ptr<float>(p + off + 4 + stops)));
            return &rs.linears.back();
        case DVG_COLOR_RADIAL:
            rs.radials.push_back(RadialGradient(Vector2f{p[off], p[off + 1]}, Vector2f{p[off + 2], p[off + 3]},
                                                stops, ptr<float>(p + off + 4), ptr<float>(p + off + 4 + stops)));
            return &rs.radials.back();
    }
    throw std::runtime_error("bad colour type");
}

std::unique_ptr<RefScene> build(const int32_t *topo, const float *params) {
    if (topo[DVG_H_MAGIC] != DVG_TOPO_MAGIC) throw std::runtime_error("bad topo magic");
    auto rs = std::make_unique<RefScene>();
    const int ns = topo[DVG_H_NUM_SHAPES], ng = topo[DVG_H_NUM_GROUPS];
    rs->params.assign(params, params + topo[DVG_H_NUM_PARAMS]);
    int topo_len = topo[DVG_H_OFF_GSHAPES] + topo[DVG_H_TOTAL_GSHAPES];
    topo_len = std::max(topo_len, topo[DVG_H_OFF_NCP] + topo[DVG_H_TOTAL_SEGS]);
    topo_len = std::max(topo_len, topo[DVG_H_OFF_GROUPS] + ng * DVG_GROUP_REC_LEN);
    topo_len = std::max(topo_len, topo[DVG_H_OFF_SHAPES] + ns * DVG_SHAPE_REC_LEN);
    rs->ints.assign(topo, topo + topo_len);
    float *p = rs->params.data();
    int *ti = rs->ints.data();
    // reserve so that pointers into the vectors stay valid
    rs->circles.reserve(ns); rs->ellipses.reserve(ns); rs->paths.reserve(ns); rs->rects.reserve(ns);
    rs->shapes.reserve(ns);
    rs->constants.reserve(2 * ng); rs->linears.reserve(2 * ng); rs->radials.reserve(2 * ng);
    rs->groups.reserve(ng);
    for (int i = 0; i < ns; i++) {
        const int *r = ti + topo[DVG_H_OFF_SHAPES] + i * DVG_SHAPE_REC_LEN;
        int off = r[DVG_S_PARAM_OFF];
        float sw = r[DVG_S_WIDTH_OFF] >= 0 ? p[r[DVG_S_WIDTH_OFF]] : 0.f;
        void *sp = nullptr;
        switch (r[DVG_S_TYPE]) {
            case DVG_SHAPE_CIRCLE:
                rs->circles.push_back(Circle{p[off], Vector2f{p[off + 1], p[off + 2]}});
                sp = &rs->circles.back();
                break;
            case DVG_SHAPE_ELLIPSE:
                rs->ellipses.push_back(Ellipse{Vector2f{p[off], p[off + 1]}, Vector2f{p[off + 2], p[off + 3]}});
                sp = &rs->ellipses.back();
                break;
            case DVG_SHAPE_PATH:
                rs->paths.push_back(Path(ptr<int>(ti + topo[DVG_H_OFF_NCP] + r[DVG_S_NCP_OFF]),
                                         ptr<float>(p + off),
                                         ptr<float>(r[DVG_S_THICK_OFF] >= 0 ? p + r[DVG_S_THICK_OFF] : nullptr),
                                         r[DVG_S_NUM_SEGS], r[DVG_S_NUM_POINTS],
                                         (r[DVG_S_FLAGS] & DVG_SF_CLOSED) != 0,
                                         (r[DVG_S_FLAGS] & DVG_SF_DISTANCE_APPROX) != 0));
                sp = &rs->paths.back();
                break;
            case DVG_SHAPE_RECT:
                rs->rects.push_back(Rect{Vector2f{p[off], p[off + 1]}, Vector2f{p[off + 2], p[off + 3]}});
                sp = &rs->rects.back();
                break;
            default:
                throw std::runtime_error("bad shape type");
        }
        rs->shapes.push_back(Shape((ShapeType)r[DVG_S_TYPE], ptr<void>(sp), sw));
    }
    for (int g = 0; g < ng; g++) {
        const int *r = ti + topo[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN;
        void *fc = color_ptr(*rs, r[DVG_G_FILL_TYPE], r[DVG_G_FILL_OFF], r[DVG_G_FILL_STOPS]);
        void *sc = color_ptr(*rs, r[DVG_G_STROKE_TYPE], r[DVG_G_STROKE_OFF], r[DVG_G_STROKE_STOPS]);
        // render_pytorch.py:349-357: absent colours are passed as (constant, null)
        rs->groups.push_back(ShapeGroup(
            ptr<int>(ti + topo[DVG_H_OFF_GSHAPES] + r[DVG_G_SHAPES_OFF]), r[DVG_G_NUM_SHAPES],
            (ColorType)(r[DVG_G_FILL_TYPE] < 0 ? 0 : r[DVG_G_FILL_TYPE]), ptr<void>(fc),
            (ColorType)(r[DVG_G_STROKE_TYPE] < 0 ? 0 : r[DVG_G_STROKE_TYPE]), ptr<void>(sc),
            r[DVG_G_EVEN_ODD] != 0, ptr<float>(p + r[DVG_G_XFORM_OFF])));
    }
    std::vector<const Shape *> sl;
    std::vector<const ShapeGroup *> gl;
    for (auto &s : rs->shapes) sl.push_back(&s);
    for (auto &g : rs->groups) gl.push_back(&g);
    Filter filt{(FilterType)topo[DVG_H_FILTER_TYPE], p[topo[DVG_H_FILTER_RADIUS_OFF]]};
#ifdef DVGREF_CUDA   // the reference's own CUDA path (oracle/Makefile target ref_cuda): timing only, see dvgref_cuda_bench
    rs->scene = std::make_shared<Scene>(topo[DVG_H_CANVAS_W], topo[DVG_H_CANVAS_H], sl, gl, filt,
                                        /*use_gpu*/ true, /*gpu_index*/ -1);
#else
    rs->scene = std::make_shared<Scene>(topo[DVG_H_CANVAS_W], topo[DVG_H_CANVAS_H], sl, gl, filt,
                                        /*use_gpu*/ false, /*gpu_index*/ -1);
#endif
    return rs;
}

void read_color_grad(int type, int off, int stops, void *d_color, float *d) {
    switch (type) {
        case DVG_COLOR_NONE: return;
        case DVG_COLOR_CONSTANT: {
            auto *c = (Constant *)d_color;
            for (int k = 0; k < 4; k++) d[off + k] += c->color[k];
            return;
        }
        case DVG_COLOR_LINEAR: {
            auto *c = (LinearGradient *)d_color;
            d[off + 0] += c->begin.x; d[off + 1] += c->begin.y;
            d[off + 2] += c->end.x;   d[off + 3] += c->end.y;
            for (int k = 0; k < stops; k++) d[off + 4 + k] += c->stop_offsets[k];
            for (int k = 0; k < 4 * stops; k++) d[off + 4 + stops + k] += c->stop_colors[k];
            return;
        }
        case DVG_COLOR_RADIAL: {
            auto *c = (RadialGradient *)d_color;
            d[off + 0] += c->center.x; d[off + 1] += c->center.y;
            d[off + 2] += c->radius.x; d[off + 3] += c->radius.y;
            for (int k = 0; k < stops; k++) d[off + 4 + k] += c->stop_offsets[k];
            for (int k = 0; k < 4 * stops; k++) d[off + 4 + stops + k] += c->stop_colors[k];
            return;
        }
    }
}

// Mirrors the field-by-field read-back of render_pytorch.py:727-866 (shapes shared by
// several groups accumulate once, as in the reference where d_shapes is per shape).
void read_grads(const int32_t *topo, const Scene &scene, float *d) {
    const int ns = topo[DVG_H_NUM_SHAPES], ng = topo[DVG_H_NUM_GROUPS];
    for (int i = 0; i < ns; i++) {
        const int *r = topo + topo[DVG_H_OFF_SHAPES] + i * DVG_SHAPE_REC_LEN;
        int off = r[DVG_S_PARAM_OFF];
        const Shape &ds = scene.d_shapes[i];
        switch (r[DVG_S_TYPE]) {
            case DVG_SHAPE_CIRCLE: {
                auto *c = (Circle *)ds.ptr;
                d[off] += c->radius; d[off + 1] += c->center.x; d[off + 2] += c->center.y;
                break;
            }
            case DVG_SHAPE_ELLIPSE: {
                auto *c = (Ellipse *)ds.ptr;
                d[off] += c->radius.x; d[off + 1] += c->radius.y;
                d[off + 2] += c->center.x; d[off + 3] += c->center.y;
                break;
            }
            case DVG_SHAPE_PATH: {
                auto *c = (Path *)ds.ptr;
                for (int k = 0; k < 2 * r[DVG_S_NUM_POINTS]; k++) d[off + k] += c->points[k];
                if (r[DVG_S_THICK_OFF] >= 0)
                    for (int k = 0; k < r[DVG_S_NUM_POINTS]; k++) d[r[DVG_S_THICK_OFF] + k] += c->thickness[k];
                break;
            }
            case DVG_SHAPE_RECT: {
                auto *c = (Rect *)ds.ptr;
                d[off] += c->p_min.x; d[off + 1] += c->p_min.y;
                d[off + 2] += c->p_max.x; d[off + 3] += c->p_max.y;
                break;
            }
        }
        if (r[DVG_S_WIDTH_OFF] >= 0) d[r[DVG_S_WIDTH_OFF]] += ds.stroke_width;
    }
    for (int g = 0; g < ng; g++) {
        const int *r = topo + topo[DVG_H_OFF_GROUPS] + g * DVG_GROUP_REC_LEN;
        const ShapeGroup &dg = scene.d_shape_groups[g];
        read_color_grad(r[DVG_G_FILL_TYPE], r[DVG_G_FILL_OFF], r[DVG_G_FILL_STOPS], dg.fill_color, d);
        // Q4: the reference never allocates d_stroke_color for gradient
        // stroke colours (scene.cpp:868,887) -> only constant stroke colours are read.
        if (r[DVG_G_STROKE_TYPE] == DVG_COLOR_CONSTANT)
            read_color_grad(r[DVG_G_STROKE_TYPE], r[DVG_G_STROKE_OFF], r[DVG_G_STROKE_STOPS], dg.stroke_color, d);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) d[r[DVG_G_XFORM_OFF] + 3 * a + b] += dg.shape_to_canvas(a, b);
    }
    d[topo[DVG_H_FILTER_RADIUS_OFF]] += scene.d_filter->radius;
}

}  // namespace

#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT const char *dvgref_last_error() { return g_err.c_str(); }

// One call = Scene construction + render(), exactly what RenderFunction.forward or
// .backward do (render_pytorch.py:366-409 / 692-707).  Null pointers mean "not requested".
// d_params (length num_params) is ACCUMULATED into; pass zeros.
EXPORT int dvgref_render(const int32_t *topo, const float *params,
                         const float *background, float *render_image, float *render_sdf,
                         int width, int height, int nsx, int nsy, uint64_t seed,
                         float *d_background, const float *d_render_image, const float *d_render_sdf,
                         float *d_translation, int use_prefiltering,
                         const float *eval_positions, int num_eval_positions,
                         float *d_params) {
    try {
        auto rs = build(topo, params);
        render(rs->scene, ptr<float>((float *)background), ptr<float>(render_image), ptr<float>(render_sdf),
               width, height, nsx, nsy, seed,
               ptr<float>(d_background), ptr<float>((float *)d_render_image), ptr<float>((float *)d_render_sdf),
               ptr<float>(d_translation), use_prefiltering != 0,
               ptr<float>((float *)eval_positions), num_eval_positions);
        if (d_params) read_grads(topo, *rs->scene, d_params);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// Debug dumps of the Scene internals the pybind module never exposed:
//  what = 0: scene BVH nodes            -> 7 floats/ints per node (child0, child1, box[4], max_radius), 2*ng-1 nodes
//  what = 1: group BVH of group `index` -> same record, 2*num_shapes-1 nodes
//  what = 2: path BVH of shape `index`  -> same record, 2*num_segs-1 nodes
//  what = 3: shapes_length[num_shapes]
//  what = 4: sample_shapes_cdf[num_total_shapes]
//  what = 5: sample_shapes_pmf[num_total_shapes]
//  what = 6: path_length_cdf of shape `index` [num_segs]
//  what = 7: path_length_pmf of shape `index` [num_segs]
//  what = 8: path_point_id_map of shape `index` [num_segs] (as int32 bit patterns)
//  what = 9: sample_shape_id[num_total_shapes] (int32)   what = 10: sample_group_id (int32)
// Records are written as raw 32-bit words (ints stay ints). Returns the number of words, or -1.
EXPORT int64_t dvgref_scene_dump(const int32_t *topo, const float *params, int what, int index,
                                 uint32_t *out, int64_t cap) {
    try {
        auto rs = build(topo, params);
        const Scene &sc = *rs->scene;
        std::vector<uint32_t> w;
        auto pushf = [&](float f) { uint32_t u; memcpy(&u, &f, 4); w.push_back(u); };
        auto pushi = [&](int i) { w.push_back((uint32_t)i); };
        auto dump_nodes = [&](const BVHNode *nodes, int nprim) {
            for (int i = 0; i < 2 * nprim - 1; i++) {
                pushi(nodes[i].child0); pushi(nodes[i].child1);
                pushf(nodes[i].box.p_min.x); pushf(nodes[i].box.p_min.y);
                pushf(nodes[i].box.p_max.x); pushf(nodes[i].box.p_max.y);
                pushf(nodes[i].max_radius);
            }
        };
        auto nsegs = [&](int shape) { return topo[topo[DVG_H_OFF_SHAPES] + shape * DVG_SHAPE_REC_LEN + DVG_S_NUM_SEGS]; };
        switch (what) {
            case 0: dump_nodes(sc.bvh_nodes, sc.num_shape_groups); break;
            case 1: dump_nodes(sc.shape_groups_bvh_nodes[index], sc.shape_groups[index].num_shapes); break;
            case 2: dump_nodes(sc.path_bvhs[index], nsegs(index)); break;
            case 3: for (int i = 0; i < sc.num_shapes; i++) pushf(sc.shapes_length[i]); break;
            case 4: for (int i = 0; i < sc.num_total_shapes; i++) pushf(sc.sample_shapes_cdf[i]); break;
            case 5: for (int i = 0; i < sc.num_total_shapes; i++) pushf(sc.sample_shapes_pmf[i]); break;
            case 6: for (int i = 0; i < nsegs(index); i++) pushf(sc.path_length_cdf[index][i]); break;
            case 7: for (int i = 0; i < nsegs(index); i++) pushf(sc.path_length_pmf[index][i]); break;
            case 8: for (int i = 0; i < nsegs(index); i++) pushi(sc.path_point_id_map[index][i]); break;
            case 9: for (int i = 0; i < sc.num_total_shapes; i++) pushi(sc.sample_shape_id[i]); break;
            case 10: for (int i = 0; i < sc.num_total_shapes; i++) pushi(sc.sample_group_id[i]); break;
            default: throw std::runtime_error("bad dump selector");
        }
        if ((int64_t)w.size() > cap) throw std::runtime_error("dump buffer too small");
        memcpy(out, w.data(), w.size() * 4);
        return (int64_t)w.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

#ifdef DVGREF_CUDA
#include <chrono>
#include <cuda_runtime.h>
// The reference's GPU path (diffvg.cpp:1477-1649 with Scene::use_gpu, the "mega-kernel" build SURVEY 2b names as the
// GPU bar) timed on the same workload as bench.py: per step Scene() + forward render, Scene() + backward render, as
// RenderFunction.forward / .backward do (render_pytorch.py:366-409, 692-707).  Images live in device memory (the
// reference is handed torch CUDA tensors); `d_image_host` is uploaded once.  Returns 0 and the mean wall-clock
// milliseconds per step (every render() ends with a device synchronisation); the last image / gradient come back
// for a sanity check.  TIMING ONLY: the nvcc build contracts FMAs, it is not a parity oracle.
EXPORT int dvgref_cuda_bench(const int32_t *topo, const float *params, int width, int height, int nsx, int nsy,
                             uint64_t seed0, const float *d_image_host, int warmup, int steps, double *ms_per_step,
                             float *image_out, float *d_params_out) {
    try {
        const size_t npx = (size_t)width * height * 4;
        float *img = nullptr, *dimg = nullptr;
        if (cudaMalloc(&img, npx * 4) != cudaSuccess || cudaMalloc(&dimg, npx * 4) != cudaSuccess)
            throw std::runtime_error("cudaMalloc failed");
        cudaMemcpy(dimg, d_image_host, npx * 4, cudaMemcpyHostToDevice);
        std::vector<float> grads(topo[DVG_H_NUM_PARAMS]);
        double total = 0;
        for (int it = 0; it < warmup + steps; it++) {
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            {
                auto rs = build(topo, params);
                cudaMemset(img, 0, npx * 4);
                render(rs->scene, ptr<float>(nullptr), ptr<float>(img), ptr<float>(nullptr), width, height, nsx, nsy, seed0 + it,
                       ptr<float>(nullptr), ptr<float>(nullptr), ptr<float>(nullptr), ptr<float>(nullptr), false, ptr<float>(nullptr), 0);
            }
            {
                auto rs = build(topo, params);
                render(rs->scene, ptr<float>(nullptr), ptr<float>(nullptr), ptr<float>(nullptr), width, height, nsx, nsy, seed0 + it,
                       ptr<float>(nullptr), ptr<float>(dimg), ptr<float>(nullptr), ptr<float>(nullptr), false, ptr<float>(nullptr), 0);
                std::fill(grads.begin(), grads.end(), 0.f);
                read_grads(topo, *rs->scene, grads.data());
            }
            auto t1 = std::chrono::steady_clock::now();
            if (it >= warmup) total += std::chrono::duration<double, std::milli>(t1 - t0).count();
        }
        if (image_out) cudaMemcpy(image_out, img, npx * 4, cudaMemcpyDeviceToHost);
        if (d_params_out) memcpy(d_params_out, grads.data(), grads.size() * 4);
        cudaFree(img); cudaFree(dimg);
        *ms_per_step = total / (steps > 0 ? steps : 1);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
#endif
