// indices.cu — rasterize_to_indices_in_range.
// Replaces CS/rasterize_to_indices_in_range.cu:17-177 (+ host :179-307).  Semantics kept:
// per pixel, walk the tile's depth-ordered list from batch `range_start` to `range_end`
// (a batch = tile_size² list entries), starting from the given transmittance; a pair is
// listed iff sigma >= 0 and alpha = min(0.999, o·exp(-sigma)) >= 1/255, and the walk stops
// (exclusive) at the first pair with T(1-alpha) <= 1e-4.  Output = for every pixel in
// (camera, row, column) order its listed Gaussians front to back.
//
// Two launches of ONE kernel (count / fill), like the reference, with the per-pixel counts
// prefix-summed in between by the single-pass look-back scan (scan.cuh) — int64 sums, the
// reference's int32 cumsum (:262-264) overflows above 2^31 pairs.  Staging: two 16-byte
// records per list entry ({x, y, opacity, a}, {b, c, gaussian, -}) so the inner loop reads
// two LDS.128 per pair; a warp leaves the batch loop as soon as its 32 pixels are done.
#include "raster_common.cuh"
#include "scan.cuh"

namespace b2s {

template <bool FILL>
__global__ void __launch_bounds__(1024)
raster_indices_kernel(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, uint64_t n_isects,
                      const float2 *__restrict__ means2d, const float *__restrict__ conics,
                      const float *__restrict__ opacities, uint32_t W, uint32_t H, uint32_t tile_size,
                      uint32_t tile_width, uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                      const int32_t *__restrict__ flatten_ids, const float *__restrict__ transmittances,
                      const int64_t *__restrict__ chunk_cum,  // inclusive prefix sums of the counts (FILL)
                      int32_t *__restrict__ chunk_cnts, int64_t *__restrict__ gaussian_ids,
                      int64_t *__restrict__ pixel_ids) {
    const TileCoord tc = tile_coord(tile_size, tile_width, tile_height, W, H);
    const uint32_t n_tiles_total = C * tile_width * tile_height;
    const uint32_t block_size = tile_size * tile_size;  // the reference's batch unit (:77)
    const int32_t isect_start = tile_offsets[tc.tile_lin];
    const int32_t isect_end = (tc.tile_lin == n_tiles_total - 1) ? (int32_t)n_isects : tile_offsets[tc.tile_lin + 1];
    const uint32_t num_batches = ((uint32_t)(isect_end - isect_start) + block_size - 1) / block_size;
    const size_t pix = ((size_t)tc.cam * H + tc.i) * W + tc.j;
    const bool inside = tc.inside;
    if (range_start >= num_batches) {
        // nothing of this tile lies in the range (reference :80-84 returns before the count
        // store; its counts buffer is zero-initialised)
        if (!FILL && inside) chunk_cnts[pix] = 0;
        return;
    }
    extern __shared__ float4 smem4[];
    float4 *rec_a = smem4;               // {x, y, opacity, conic.a}
    float4 *rec_b = smem4 + blockDim.x;  // {conic.b, conic.c, flat id (bits), -}

    const float px = (float)tc.j + 0.5f, py = (float)tc.i + 0.5f;
    bool done = !inside;
    float trans = inside ? transmittances[pix] : 0.f;
    int64_t base = 0;
    if (FILL && inside) base = chunk_cum[pix] - (int64_t)chunk_cnts[pix];
    int32_t cnt = 0;
    const uint32_t tr = threadIdx.x;
    const uint32_t b_end = min(range_end, num_batches);
    for (uint32_t b = range_start; b < b_end; ++b) {
        if (__syncthreads_count(done) >= (int)blockDim.x) break;
        const uint32_t batch_start = (uint32_t)isect_start + block_size * b;
        // threads beyond tile_size² (block rounded up to a warp multiple) stage nothing
        if (tr < block_size && batch_start + tr < (uint32_t)isect_end) {
            const int32_t g = flatten_ids[batch_start + tr];
            const float2 xy = __ldg(means2d + g);
            const float ca = __ldg(conics + 3 * (size_t)g), cb = __ldg(conics + 3 * (size_t)g + 1),
                        cc = __ldg(conics + 3 * (size_t)g + 2);
            rec_a[tr] = make_float4(xy.x, xy.y, __ldg(opacities + g), ca);
            rec_b[tr] = make_float4(cb, cc, __int_as_float(g), 0.f);
        }
        __syncthreads();
        const uint32_t batch_size = min(block_size, (uint32_t)isect_end - batch_start);
        for (uint32_t t = 0; t < batch_size && !done; ++t) {
            const float4 ra = rec_a[t];
            const float4 rb = rec_b[t];
            const float dx = ra.x - px, dy = ra.y - py;
            const float sigma = 0.5f * (ra.w * dx * dx + rb.y * dy * dy) + rb.x * dx * dy;
            const float alpha = fminf(kAlphaMax, ra.z * __expf(-sigma));
            if (sigma < 0.f || alpha < kAlphaMin) continue;
            const float next_trans = trans * (1.f - alpha);
            if (next_trans <= kTransmittanceEps) { done = true; break; }
            if (FILL) {
                const int32_t g = __float_as_int(rb.z);
                gaussian_ids[base + cnt] = (int64_t)((uint32_t)g % N);
                pixel_ids[base + cnt] = (int64_t)pix;
            }
            ++cnt;
            trans = next_trans;
        }
    }
    if (!FILL && inside) chunk_cnts[pix] = cnt;
}

template <bool FILL>
static int launch_indices(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N, uint64_t n_isects,
                          const float *means2d, const float *conics, const float *opacities, uint32_t W, uint32_t H,
                          uint32_t tile_size, uint32_t tile_width, uint32_t tile_height, const int32_t *tile_offsets,
                          const int32_t *flatten_ids, const float *transmittances, const int64_t *chunk_cum,
                          int32_t *chunk_cnts, int64_t *gaussian_ids, int64_t *pixel_ids, cudaStream_t st) {
    const uint32_t threads = ((tile_size * tile_size + 31) / 32) * 32;
    const uint32_t grid = C * tile_width * tile_height;
    const size_t smem = (size_t)threads * 2 * sizeof(float4);
    raster_indices_kernel<FILL><<<grid, threads, smem, st>>>(
        range_start, range_end, C, N, n_isects, reinterpret_cast<const float2 *>(means2d), conics, opacities, W, H,
        tile_size, tile_width, tile_height, tile_offsets, flatten_ids, transmittances, chunk_cum, chunk_cnts,
        gaussian_ids, pixel_ids);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace b2s

using namespace b2s;

extern "C" int b200splat_raster_indices_count(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N,
                                              uint64_t n_isects, const float *means2d, const float *conics,
                                              const float *opacities, uint32_t W, uint32_t H, uint32_t tile_size,
                                              uint32_t tile_width, uint32_t tile_height, const int32_t *tile_offsets,
                                              const int32_t *flatten_ids, const float *transmittances,
                                              int32_t *chunk_cnts, int64_t *chunk_cum, int64_t *n_elems_out,
                                              void *scan_workspace, size_t scan_workspace_bytes_, void *stream) {
    const char *where = "b200splat_raster_indices_count";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    cudaStream_t st = (cudaStream_t)stream;
    B2S_REQUIRE(tile_size >= 1 && tile_size <= 32, where, "tile_size must be in [1, 32]");
    B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
    B2S_REQUIRE((uint64_t)tile_width * tile_size >= W && (uint64_t)tile_height * tile_size >= H, where,
                "tile grid does not cover the image");
    const uint64_t n_pix = (uint64_t)C * H * W;
    if (n_pix == 0 || n_isects == 0 || N == 0) {
        cudaMemsetAsync(n_elems_out, 0, sizeof(int64_t), st);
        if (n_pix) cudaMemsetAsync(chunk_cnts, 0, n_pix * sizeof(int32_t), st);
        return 0;
    }
    if (launch_indices<false>(range_start, range_end, C, N, n_isects, means2d, conics, opacities, W, H, tile_size,
                              tile_width, tile_height, tile_offsets, flatten_ids, transmittances, nullptr, chunk_cnts,
                              nullptr, nullptr, st))
        return fail_cuda(where, cudaGetLastError());
    const int rc = lookback_scan_i32_to_i64(chunk_cnts, chunk_cum, n_pix, n_elems_out, scan_workspace,
                                            scan_workspace_bytes_, st);
    if (rc == 2) return fail(where, "scan workspace too small (see b200splat_scan_workspace_bytes)");
    if (rc) return fail_cuda(where, cudaGetLastError());
    return 0;
}

extern "C" int b200splat_raster_indices_fill(uint32_t range_start, uint32_t range_end, uint32_t C, uint32_t N,
                                             uint64_t n_isects, const float *means2d, const float *conics,
                                             const float *opacities, uint32_t W, uint32_t H, uint32_t tile_size,
                                             uint32_t tile_width, uint32_t tile_height, const int32_t *tile_offsets,
                                             const int32_t *flatten_ids, const float *transmittances,
                                             const int32_t *chunk_cnts, const int64_t *chunk_cum,
                                             int64_t *gaussian_ids, int64_t *pixel_ids, void *stream) {
    const char *where = "b200splat_raster_indices_fill";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    B2S_REQUIRE(tile_size >= 1 && tile_size <= 32, where, "tile_size must be in [1, 32]");
    B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
    if ((uint64_t)C * H * W == 0 || n_isects == 0 || N == 0) return 0;
    if (launch_indices<true>(range_start, range_end, C, N, n_isects, means2d, conics, opacities, W, H, tile_size,
                             tile_width, tile_height, tile_offsets, flatten_ids, transmittances, chunk_cum,
                             const_cast<int32_t *>(chunk_cnts), gaussian_ids, pixel_ids, (cudaStream_t)stream))
        return fail_cuda(where, cudaGetLastError());
    return 0;
}
