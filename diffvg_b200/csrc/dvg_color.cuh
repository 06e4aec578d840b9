// dvg_color.cuh -- colour evaluation (constant / linear / radial gradient) and its adjoint.
// Follows diffvg.cpp:276-368 (sample_color) and 370-504 (d_sample_color).
#pragma once
#include "dvg_scene.cuh"

namespace dvg {

#if defined(__CUDA_ARCH__)
#define DVG_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define DVG_ATOMIC_ADD(ptr, v) (*(ptr) += (v))
#endif

// ---------------------------------------------------------------- gradient sinks
// The reference scatters every per-sample gradient term with one global float atomic
// (atomic.h:23-51): ~23 per pixel sample onto ~57 k addresses at the painterly config, and ALL
// samples onto the same 9 floats when the groups share one shape_to_canvas.  Here terms go to a
// per-block open-addressed table in shared memory keyed by the parameter index (shared-memory
// atomics), and the block issues ONE global reduction per distinct parameter it touched.
// A full table (probe limit) falls back to the global atomic, so the result never depends on it.
struct DirectSink {   // host harness / fallback: straight into d_params
    float *D;
    DVG_HD void add(int idx, float v) const { DVG_ATOMIC_ADD(D + idx, v); }
};

#if defined(__CUDACC__)
#ifndef DVG_GC_LOG2
#define DVG_GC_LOG2 10
#endif
constexpr int DVG_GC_SLOTS = 1 << DVG_GC_LOG2;
struct GradCache {
    int keys[DVG_GC_SLOTS];
    float vals[DVG_GC_SLOTS];
};
struct CacheSink {
    GradCache *gc;
    float *D;
    // out of line on purpose: the boundary kernel has ~30 call sites and is instruction-fetch bound
    __device__ __noinline__ void add_slow(int idx, float v) const {
        unsigned h = ((unsigned)idx * 2654435761u) >> (32 - DVG_GC_LOG2);
#pragma unroll 1
        for (int probe = 0; probe < 16; probe++) {
            int k = gc->keys[h];
            if (k == -1) k = atomicCAS(&gc->keys[h], -1, idx);
            if (k == -1 || k == idx) { atomicAdd(&gc->vals[h], v); return; }
            h = (h + 1) & (DVG_GC_SLOTS - 1);
        }
        atomicAdd(D + idx, v);
    }
    __device__ __forceinline__ void add(int idx, float v) const {
        if (v == 0.f) return;
        add_slow(idx, v);
    }
};
__device__ __forceinline__ void grad_cache_init(GradCache &gc) {
    for (int i = threadIdx.x; i < DVG_GC_SLOTS; i += blockDim.x) { gc.keys[i] = -1; gc.vals[i] = 0.f; }
}
// call after a __syncthreads() that orders all adds before it
__device__ __forceinline__ void grad_cache_flush(GradCache &gc, float *D) {
    for (int i = threadIdx.x; i < DVG_GC_SLOTS; i += blockDim.x) {
        const int k = gc.keys[i];
        const float v = gc.vals[i];
        if (k >= 0 && v != 0.f) atomicAdd(D + k, v);
    }
}
#endif

// Gradient parameter t for a linear / radial gradient record at params[off..].
DVG_HD float gradient_t(int type, const float *c, F2 pt) {
    if (type == 1) {  // diffvg.cpp:288-290
        F2 beg = mk2(c[0], c[1]), end = mk2(c[2], c[3]);
        return dot2(pt - beg, end - beg) / rmaxf(dot2(end - beg, end - beg), 1e-3f);
    } else {  // diffvg.cpp:327-329
        F2 offset = pt - mk2(c[0], c[1]);
        F2 no = mk2(offset.x / c[2], offset.y / c[3]);
        return length2(no);
    }
}

DVG_HD F4 load4(const float *p) { return mk4(p[0], p[1], p[2], p[3]); }

// diffvg.cpp:276-368.  `c` points at the colour record inside params.
DVG_HD_NOINLINE F4 eval_color_gradient(int type, const float *c, int num_stops, F2 pt) {
    float t = gradient_t(type, c, pt);
    const float *offsets = c + 4;
    const float *colors = c + 4 + num_stops;
    if (t < offsets[0]) return load4(colors);
    for (int i = 0; i < num_stops - 1; i++) {
        float oc = offsets[i], on = offsets[i + 1];
        if (t >= oc && t < on) {
            F4 cc = load4(colors + 4 * i), cn = load4(colors + 4 * (i + 1));
            float tt = (t - oc) / (on - oc);
            return cc * (1 - tt) + cn * tt;
        }
    }
    return load4(colors + 4 * (num_stops - 1));
}
DVG_HD F4 eval_color(int type, const float *c, int num_stops, F2 pt) {
    if (type == 0) return load4(c);
    return eval_color_gradient(type, c, num_stops, pt);
}

template <typename Sink>
DVG_D void add4(const Sink &sk, int d, F4 v) {
    sk.add(d + 0, v.x); sk.add(d + 1, v.y); sk.add(d + 2, v.z); sk.add(d + 3, v.w);
}

// diffvg.cpp:370-504 for the two gradient types (the constant case is reduced by the
// caller).  `d` points at the record's slot in d_params.  d_translation: 2 floats or null.
// Q3: the radial branch has no `return` after the matched stop, so d_color is also
// added to the last stop; reproduced.
template <typename Sink>
DVG_D_NOINLINE void d_eval_gradient(int type, const float *c, int num_stops, F2 pt, F4 d_color, const Sink &sk, int d, float *d_translation) {
    float t = gradient_t(type, c, pt);
    const float *offsets = c + 4;
    const float *colors = c + 4 + num_stops;
    const int d_offsets = d + 4;
    const int d_colors = d + 4 + num_stops;
    if (t < offsets[0]) { add4(sk, d_colors, d_color); return; }
    for (int i = 0; i < num_stops - 1; i++) {
        float oc = offsets[i], on = offsets[i + 1];
        if (t >= oc && t < on) {
            F4 cc = load4(colors + 4 * i), cn = load4(colors + 4 * (i + 1));
            float tt = (t - oc) / (on - oc);
            F4 d_cc = d_color * (1 - tt), d_cn = d_color * tt;
            float d_tt = sum4(d_color * (cn - cc));
            float d_on = -d_tt * tt / (on - oc);
            float d_oc = d_tt * ((tt - 1.f) / (on - oc));
            float d_t = d_tt / (on - oc);
            add4(sk, d_colors + 4 * i, d_cc);
            add4(sk, d_colors + 4 * (i + 1), d_cn);
            sk.add(d_offsets + i, d_oc);
            sk.add(d_offsets + i + 1, d_on);
            if (type == 1) {
                F2 beg = mk2(c[0], c[1]), end = mk2(c[2], c[3]);
                float l = rmaxf(dot2(end - beg, end - beg), 1e-3f);
                F2 d_beg = (d_t * (-(pt - beg) - (end - beg))) / l;
                F2 d_end = (d_t * (pt - beg)) / l;
                float d_l = -d_t * t / l;
                if (dot2(end - beg, end - beg) > 1e-3f) {
                    d_beg = d_beg + (2 * d_l) * (beg - end);
                    d_end = d_end + (2 * d_l) * (end - beg);
                }
                sk.add(d + 0, d_beg.x); sk.add(d + 1, d_beg.y);
                sk.add(d + 2, d_end.x); sk.add(d + 3, d_end.y);
                if (d_translation) {
                    DVG_ATOMIC_ADD(d_translation + 0, d_beg.x + d_end.x);
                    DVG_ATOMIC_ADD(d_translation + 1, d_beg.y + d_end.y);
                }
                return;
            } else {
                F2 offset = pt - mk2(c[0], c[1]);
                F2 radius = mk2(c[2], c[3]);
                F2 no = mk2(offset.x / radius.x, offset.y / radius.y);
                // d_length (vector.h:496-503)
                float l_sq = no.x * no.x + no.y * no.y;
                float l = sqrtf(l_sq);
                float d_l_sq = 0.5f * d_t / l;
                F2 d_no = (2 * d_l_sq) * no;
                F2 d_offset = mk2(d_no.x / radius.x, d_no.y / radius.y);
                F2 d_radius = mk2(-d_no.x * offset.x / (radius.x * radius.x), -d_no.y * offset.y / (radius.y * radius.y));
                F2 d_center = -d_offset;
                sk.add(d + 0, d_center.x); sk.add(d + 1, d_center.y);
                sk.add(d + 2, d_radius.x); sk.add(d + 3, d_radius.y);
                if (d_translation) {
                    DVG_ATOMIC_ADD(d_translation + 0, d_center.x);
                    DVG_ATOMIC_ADD(d_translation + 1, d_center.y);
                }
                // no return: falls through the loop (Q3)
            }
        }
    }
    add4(sk, d_colors + 4 * (num_stops - 1), d_color);
}

}  // namespace dvg
