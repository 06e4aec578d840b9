// dvg_render.cu -- per-sample helper kernels of the render passes: the filter-weight splat (weight_kernel) and the
// generation + counting sort of the boundary samples (sample_boundary_kernel and the reference's thrust sort).
// The render passes themselves are the wavefront kernels of dvg_wave.cu (sampled path) and dvg_prefilter.cu.
#include "dvg_internal.h"
#include "dvg_kernel_util.cuh"

#include <algorithm>

namespace dvg {

// ------------------------------------------------------------------------------------------
// weight_kernel (diffvg.cpp:1115-1158).  Q1: always uses the jittered position, even when
// the render kernel uses pixel centres for prefiltering.
// One thread per PIXEL: it draws the pixel's samples (same RNG streams: one per sample index), sums what they add to
// their own pixel in a register and sends only the rest -- nothing at all for a filter of radius <= 0.5 unless a sample
// lands exactly on a pixel edge -- to the neighbours with atomics.  (One thread and one atomic per SAMPLE cost 0.25 ms
// at 2048^2 x 4 spp, all of it repeated by every rank of a row-sharded sampled render.)
__global__ void k_weight(SceneView sc, RenderArgs ra_all, int px_begin, int px_end) {
    const int per_scene = ra_all.width * ra_all.height;
    const int gpx = px_begin + blockIdx.x * blockDim.x + threadIdx.x;   // batch: scenes back to back
    if (gpx >= px_end) return;
    const int scene = gpx / per_scene, pix = gpx - scene * per_scene;
    const RenderArgs ra = args_of_scene(ra_all, scene);
    const int x = pix % ra.width, y = pix / ra.width;
    const int spp = ra.nsx * ra.nsy;
    const int ri = (int)ceilf(sc.filter.radius);
    float own = 0.f;
    for (int s = 0; s < spp; s++) {
        const int idx = pix * spp + s;                 // ((y * W + x) * nsy + sy) * nsx + sx
        Pcg32 rng = pcg32_init(idx, ra.seed);
        const int sx = s % ra.nsx, sy = s / ra.nsx;
        const float rx = pcg32_next_float(rng);
        const float ry = pcg32_next_float(rng);
        const F2 pt = mk2(x + ((float)sx + rx) / ra.nsx, y + ((float)sy + ry) / ra.nsy);
        for (int dy = -ri; dy <= ri; dy++) {
            for (int dx = -ri; dx <= ri; dx++) {
                const int xx = x + dx, yy = y + dy;
                if (xx >= 0 && xx < ra.width && yy >= 0 && yy < ra.height) {
                    const float w = filter_weight(sc.filter, (xx + 0.5f) - pt.x, (yy + 0.5f) - pt.y);
                    if (w == 0.f) continue;
                    if (dx == 0 && dy == 0) own += w;
                    else atomicAdd(&ra.weight_image[yy * ra.width + xx], w);
                }
            }
        }
    }
    if (own != 0.f) atomicAdd(&ra.weight_image[pix], own);
}

// ------------------------------------------------------------------------------------------
// Boundary pass, step 1: sample_boundary_kernel (diffvg.cpp:1325-1386) reduced to what the
// 48-byte record the render passes need plus the destination tile of every boundary sample.  The reference's
// Morton keys + thrust::sort_by_key round trip (diffvg.cpp:1560-1595) is replaced by a counting sort by tile.
// Persistent blocks: the shape CDF (searched by every sample: 11 dependent loads at 2048 shapes) is staged in shared
// memory once per block when it fits.
constexpr int KEYS_B = 256;
constexpr int KEYS_CDF_MAX = 8192;
__global__ void __launch_bounds__(KEYS_B) k_boundary_keys(SceneView sc, BinView bins, RenderArgs ra, BoundaryWork bw) {
    __shared__ float s_cdf[KEYS_CDF_MAX];
    __shared__ int s_guide[DVG_CDF_GUIDE + 1];
    const bool staged = sc.num_insts <= KEYS_CDF_MAX && sc.batch == 1;
    if (staged) {
        for (int i = threadIdx.x; i < sc.num_insts; i += KEYS_B) s_cdf[i] = sc.shape_cdf[i];
        for (int i = threadIdx.x; i <= DVG_CDF_GUIDE; i += KEYS_B) s_guide[i] = sc.shape_guide[i];
        __syncthreads();
    }
    const int per_scene = ra.width * ra.height * ra.nsx * ra.nsy;
    for (int k = blockIdx.x * KEYS_B + threadIdx.x; k < bw.num_samples; k += gridDim.x * KEYS_B) {
        BoundarySample bs;
        // batch: scene b owns the boundary-sample indices [0, W*H*spp) of ITS seed's streams; a staged CDF is scene 0's only
        const int gk = bw.sample_begin + k;
        const int scene = sc.batch > 1 ? gk / per_scene : 0;
        const int idx = gk - scene * per_scene;
        make_boundary_sample(sc, idx, ra.seeds ? ra.seeds[scene] : ra.seed, bs, staged ? s_cdf : nullptr, scene, staged ? s_guide : nullptr);
        int key = -1;
        if (bs.inst >= 0) {
            const int bx = (int)(bs.pt.x * ra.width), by = (int)(bs.pt.y * ra.height);  // diffvg.cpp:1405-1409
            if (bx >= 0 && bx < ra.width && by >= 0 && by < ra.height) {
                key = scene * bin_scene_tiles(bins) + (by / bins.tile_h) * bins.tiles_x + bx / bins.tile_w;
                atomicAdd(&bw.tile_counts[key], 1);
            }
        }
        bw.keys[k] = key;
        if (bw.samples_unsorted) bw.samples_unsorted[k] = bs;
    }
}

// Counting sort, scatter step.  The wavefront path moves the 48-byte sample RECORDS into tile order (one coalesced read,
// one scattered write) so that the two kernels that consume them read them in order.
__global__ void k_boundary_scatter(BoundaryWork bw) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bw.num_samples) return;
    const int key = bw.keys[k];
    if (key < 0) return;
    const int pos = bw.tile_offsets[key] + atomicAdd(&bw.tile_fill[key], 1);
    bw.sorted_idx[pos] = bw.sample_begin + k;
    if (bw.samples) {
        const uint4 *src = reinterpret_cast<const uint4 *>(bw.samples_unsorted + k);
        uint4 *dst = reinterpret_cast<uint4 *>(bw.samples + pos);
        const uint4 a = src[0], b = src[1], c = src[2];
        dst[0] = a; dst[1] = b; dst[2] = c;
    }
}

// Weights of the pixels of rows [row_begin, row_end): the samples of those rows and of ceil(radius) rows either side
// splat onto them (a row shard does not need the rest of the image).
void launch_weight(const SceneView &sc, const RenderArgs &ra, int row_begin, int row_end, cudaStream_t st) {
    // pixels up to ri rows outside the band are read (gather_d_color, splat); their weights need samples ri rows further
    const int ri = 2 * (int)ceilf(sc.filter.radius);
    const int y0 = max(0, row_begin - ri), y1 = min(ra.height, row_end + ri);
    if (sc.batch > 1) {   // every scene, whole images
        const int n = ra.height * ra.width * sc.batch;
        DVG_LAUNCH(k_weight, dim3((n + 255) / 256), dim3(256), 0, st, sc, ra, 0, n);
        return;
    }
    const int n = (y1 - y0) * ra.width;
    if (n <= 0) return;
    DVG_LAUNCH(k_weight, dim3((n + 255) / 256), dim3(256), 0, st, sc, ra, y0 * ra.width, y1 * ra.width);
}

// tile keys + counting sort of the boundary samples by tile (fills tile_counts, tile_offsets, sorted_idx)
void launch_boundary_sort(const SceneView &sc, const BinView &bins, const RenderArgs &ra, const BoundaryWork &bw, cudaStream_t st) {
    const int ntiles = bin_total_tiles(bins);
    if (bw.num_samples <= 0) return;
    cudaMemsetAsync(bw.tile_counts, 0, sizeof(int) * ntiles, st);
    cudaMemsetAsync(bw.tile_fill, 0, sizeof(int) * ntiles, st);
    DVG_LAUNCH(k_boundary_keys, dim3(std::min((bw.num_samples + KEYS_B - 1) / KEYS_B, g_num_sms * 8)), dim3(KEYS_B), 0, st, sc, bins, ra, bw);
    launch_scan(bw.tile_counts, bw.tile_offsets, ntiles, bw.scan_ws, st);
    DVG_LAUNCH(k_boundary_scatter, dim3((bw.num_samples + 255) / 256), dim3(256), 0, st, bw);
}

}  // namespace dvg
