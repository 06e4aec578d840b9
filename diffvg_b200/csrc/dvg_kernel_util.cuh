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

// The same for the BOX filter with ceil(radius) == 1, the samples of a pixel sitting in `grp` adjacent lanes (a power of
// two; must be called by all lanes of the warp).  Box: d_compute_filter_weight does not depend on the offset (filter.h:
// 52-57), so the sample's gradient is K * sum over the 3x3 pixels of d_weight; for the eight NEIGHBOURS a sample's filter
// weight is 0 (it lies inside its own pixel) and d_weight reduces to (d_pixel / weight_sum) . color -- the vector
// sum_n d_pixel_n / weight_sum_n is the PIXEL's, formed once by the lanes of the group (one neighbour each) instead of
// eight loads and divisions per sample.  A sample on a pixel edge (or outside its own pixel's support when radius < 0.5)
// takes the generic loop.  Summation order differs from the generic form by float rounding only (the reference adds
// these terms with atomics in arbitrary order).
DVG_D float filter_radius_grad_box(const SceneView &sc, const RenderArgs &ra, int x, int y, F2 pt, F4 color, bool active, int grp, int lane) {
    const unsigned FULL = 0xffffffffu;
    float4 V = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        for (int j = lane & (grp - 1); j < 8; j += grp) {
            const int n = j < 4 ? j : j + 1;                       // 3x3 offsets without the centre
            const int xx = x + n % 3 - 1, yy = y + n / 3 - 1;
            if (xx >= 0 && xx < ra.width && yy >= 0 && yy < ra.height) {
                const float ws = ra.weight_image[yy * ra.width + xx];
                if (ws > 0) {
                    const float4 dp = *reinterpret_cast<const float4 *>(ra.d_render_image + 4 * (yy * ra.width + xx));
                    const float inv = ws / (ws * ws);              // (dotv * ws) / (ws * ws), diffvg.cpp:1258-1262 with fw = 0
                    V.x += dp.x * inv; V.y += dp.y * inv; V.z += dp.z * inv; V.w += dp.w * inv;
                }
            }
        }
    }
    for (int o = grp >> 1; o > 0; o >>= 1) {
        V.x += __shfl_xor_sync(FULL, V.x, o); V.y += __shfl_xor_sync(FULL, V.y, o);
        V.z += __shfl_xor_sync(FULL, V.z, o); V.w += __shfl_xor_sync(FULL, V.w, o);
    }
    if (!active) return 0.f;
    const float r = sc.filter.radius;
    const float ddx = (x + 0.5f) - pt.x, ddy = (y + 0.5f) - pt.y;
    // every neighbour outside the sample's support, the own pixel inside it?
    if (!(1.f - fabsf(ddx) > r && 1.f - fabsf(ddy) > r && fabsf(ddx) <= r && fabsf(ddy) <= r)) return filter_radius_grad(sc, ra, x, y, pt, color);
    float sum = V.x * color.x + V.y * color.y + V.z * color.z + V.w * color.w;
    const float ws = ra.weight_image[y * ra.width + x];
    if (ws > 0) {
        const float fw = filter_weight(sc.filter, ddx, ddy);
        const float4 dp = *reinterpret_cast<const float4 *>(ra.d_render_image + 4 * (y * ra.width + x));
        const float dotv = dp.x * color.x + dp.y * color.y + dp.z * color.z + dp.w * color.w;
        sum += (dotv * ws - fw * dotv * (ws - fw)) / (ws * ws);
    }
    return d_filter_weight_radius(sc.filter, ddx, ddy, sum);
}

}  // namespace dvg
