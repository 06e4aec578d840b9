// dvg_kernel_util.cuh -- device helpers shared by the render kernels (dvg_render.cu, dvg_prefilter.cu).
#pragma once
#include "dvg_internal.h"

namespace dvg {

DVG_D float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Gradient scatter straight to global memory (fire-and-forget red.global.add.f32).
struct GlobalSink {
    float *D;
    __device__ __forceinline__ void add(int idx, float v) const {
        if (v != 0.f) atomicAdd(D + idx, v);
    }
};

// Contention: some addresses are hit by EVERY warp of a launch (d_filter.radius; d_shape_to_canvas of a transform
// tensor shared by all groups, typically the default eye(3)).  Instead of reducing them per block in shared memory
// behind a barrier (warps of a block finish at very different times: the barrier was 27% of the interior backward
// kernel), every block adds into one of `grad_reps` private copies of the whole gradient buffer, chosen by block
// index; k_wave_reduce_grads sums the copies.  32 copies x 155 KB at the painterly config.
DVG_D float *grad_replica(const RenderArgs &ra) {
    return ra.d_params_rep + (size_t)(blockIdx.x & (unsigned)(ra.grad_reps - 1)) * (size_t)ra.num_params;
}

// Splat of one sample's colour (diffvg.cpp:1224-1249).  Must be called by all lanes of the warp.
// The sample's own pixel is reduced across the `grp` adjacent lanes that hold the samples of that
// pixel first; the (rare for box 0.5) neighbours go straight to global memory.
DVG_D void splat_color(const SceneView &sc, const RenderArgs &ra, int x, int y, F2 pt, F4 color, bool active, int grp, int tid) {
    F4 own = mk4(0, 0, 0, 0);
    if (active) {
        const int ri = (int)ceilf(sc.filter.radius);
        for (int dy = -ri; dy <= ri; dy++) {
            for (int dx = -ri; dx <= ri; dx++) {
                const int xx = x + dx, yy = y + dy;
                if (xx >= 0 && xx < ra.width && yy >= 0 && yy < ra.height) {
                    const float fw = filter_weight(sc.filter, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
                    if (fw == 0.f) continue;
                    const float wsum = ra.weight_image[yy * ra.width + xx];
                    if (!(wsum > 0)) continue;
                    const float inv_ws = 1.f / wsum;  // Vector4 / scalar == * (1.f / s)
                    const F4 wc = mk4((fw * color.x) * inv_ws, (fw * color.y) * inv_ws,
                                      (fw * color.z) * inv_ws, (fw * color.w) * inv_ws);
                    if (dx == 0 && dy == 0) own = wc;
                    else {
                        float *d = ra.render_image + 4 * (yy * ra.width + xx);
                        atomicAdd(d + 0, wc.x); atomicAdd(d + 1, wc.y); atomicAdd(d + 2, wc.z); atomicAdd(d + 3, wc.w);
                    }
                }
            }
        }
    }
    for (int o = grp >> 1; o > 0; o >>= 1) {
        own.x += __shfl_xor_sync(0xffffffffu, own.x, o); own.y += __shfl_xor_sync(0xffffffffu, own.y, o);
        own.z += __shfl_xor_sync(0xffffffffu, own.z, o); own.w += __shfl_xor_sync(0xffffffffu, own.w, o);
    }
    if (active && (tid & (grp - 1)) == 0) {
        float *d = ra.render_image + 4 * (y * ra.width + x);
        if (own.x != 0.f) atomicAdd(d + 0, own.x);
        if (own.y != 0.f) atomicAdd(d + 1, own.y);
        if (own.z != 0.f) atomicAdd(d + 2, own.z);
        if (own.w != 0.f) atomicAdd(d + 3, own.w);
    }
}

// Filter-radius gradient of one sample (diffvg.cpp:1250-1268).  The reference evaluates
// d_compute_filter_weight for every in-range pixel with weight > 0, even where the filter weight
// itself is zero.
DVG_D float filter_radius_grad(const SceneView &sc, const RenderArgs &ra, int x, int y, F2 pt, F4 color) {
    float acc = 0.f;
    const int ri = (int)ceilf(sc.filter.radius);
    for (int dy = -ri; dy <= ri; dy++) {
        for (int dx = -ri; dx <= ri; dx++) {
            const int xx = x + dx, yy = y + dy;
            if (xx >= 0 && xx < ra.width && yy >= 0 && yy < ra.height) {
                const float ws = ra.weight_image[yy * ra.width + xx];
                if (!(ws > 0)) continue;
                const float ddx = (xx + 0.5f) - pt.x, ddy = (yy + 0.5f) - pt.y;
                const float fw = filter_weight(sc.filter, ddx, ddy);
                const float4 dp = *reinterpret_cast<const float4 *>(ra.d_render_image + 4 * (yy * ra.width + xx));
                const float dotv = dp.x * color.x + dp.y * color.y + dp.z * color.z + dp.w * color.w;
                const float d_weight = (dotv * ws - fw * dotv * (ws - fw)) / (ws * ws);
                acc += d_filter_weight_radius(sc.filter, ddx, ddy, d_weight);
            }
        }
    }
    return acc;
}

}  // namespace dvg
