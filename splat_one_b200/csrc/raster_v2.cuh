// raster_v2.cuh — shared pieces of the B200 "warp-per-tile" rasterizer (tile_size 16,
// <= 4 channels: the RGB / RGB+depth cases splat_one renders).
//
// Mapping: one warp owns one 16x16 tile; lane l owns column x = l & 15 and the 8 rows
// y = (l >> 4) * 8 + j, j = 0..7.  Consequences:
//   * a Gaussian staged in shared memory is fetched by ONE broadcast load per lane and used
//     for 8 pixels (the v1 kernel fetched it once per pixel);
//   * dx is shared by a lane's 8 pixels, so sigma_j = A + dy_j (B + C dy_j) with
//     A = a/2 dx², B = b dx, C = c/2 costs 2 FMA per pixel;
//   * the backward sums a lane's 8 pixels in registers before any cross-lane traffic, so a
//     (tile, Gaussian) pair needs ONE warp reduction instead of one per 32 pixels;
//   * no __syncthreads anywhere: a warp synchronises only with itself.
// Staging: each lane gathers one packed 48-byte record (3 x LDG.128), tests whether the
// Gaussian can reach alpha >= 1/255 anywhere in this tile (exact minimum of the quadratic
// form over the tile rectangle, with a safety slack) and only survivors are compacted into
// shared memory — the reference's intersection list is a conservative 3-sigma bounding
// square, so a large share of (tile, Gaussian) pairs can never contribute.  Culling never
// changes a result: a culled pair would have been rejected by the per-pixel alpha test.
#pragma once
#include "raster_common.cuh"

namespace b2s {

constexpr int kV2Warps = 4;         // warps (= tiles) per CTA
constexpr int kV2Tile = 16;
constexpr int kV2Rows = 8;          // pixels per lane

// record r of Gaussian g: rec[3g] = {x, y, a/2, b}, rec[3g+1] = {c/2, opacity, c0, c1},
// rec[3g+2] = {c2, c3, 0, 0}
static __global__ void __launch_bounds__(kThreads)
pack_records_kernel(uint32_t n, uint32_t channels, const float2 *__restrict__ means2d,
                    const float *__restrict__ conics, const float *__restrict__ colors,
                    const float *__restrict__ opacities, float4 *__restrict__ rec) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const float2 xy = means2d[g];
    const float a = conics[3 * (size_t)g], b = conics[3 * (size_t)g + 1], c = conics[3 * (size_t)g + 2];
    float col[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint32_t k = 0; k < channels; ++k) col[k] = colors[(size_t)g * channels + k];
    rec[3 * (size_t)g] = make_float4(xy.x, xy.y, 0.5f * a, b);
    rec[3 * (size_t)g + 1] = make_float4(0.5f * c, opacities[g], col[0], col[1]);
    rec[3 * (size_t)g + 2] = make_float4(col[2], col[3], 0.f, 0.f);
}

// Can this Gaussian reach alpha >= 1/255 at any pixel centre in [x0,x1] x [y0,y1]?
// Conservative: answers true whenever unsure (non positive-definite conic, NaNs).
__device__ __forceinline__ bool tile_may_contribute(float gx, float gy, float hA, float b, float hC, float opac,
                                                    float x0, float y0, float x1, float y1) {
    // d = g - p with p in the rectangle
    const float dx0 = gx - x1, dx1 = gx - x0, dy0 = gy - y1, dy1 = gy - y0;
    const bool pd = (hA > 0.f) && (hC > 0.f) && (4.f * hA * hC - b * b > 0.f);
    if (!pd) return true;
    float qmin;
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) {
        qmin = 0.f;
    } else {
        // minimum of the convex form over the boundary: 4 edges, 1-D clamp on each
        const float ihC = -0.5f * b / hC, ihA = -0.5f * b / hA;
        float t, q;
        t = fminf(fmaxf(ihC * dx0, dy0), dy1); qmin = hA * dx0 * dx0 + t * (b * dx0 + hC * t);
        t = fminf(fmaxf(ihC * dx1, dy0), dy1); q = hA * dx1 * dx1 + t * (b * dx1 + hC * t); qmin = fminf(qmin, q);
        t = fminf(fmaxf(ihA * dy0, dx0), dx1); q = hC * dy0 * dy0 + t * (b * dy0 + hA * t); qmin = fminf(qmin, q);
        t = fminf(fmaxf(ihA * dy1, dx0), dx1); q = hC * dy1 * dy1 + t * (b * dy1 + hA * t); qmin = fminf(qmin, q);
    }
    // slack: fp32 rounding of the per-pixel sigma (terms up to `mag`) plus 1 % in alpha
    const float mx = fmaxf(fabsf(dx0), fabsf(dx1)), my = fmaxf(fabsf(dy0), fabsf(dy1));
    const float mag = hA * mx * mx + hC * my * my + fabsf(b) * mx * my;
    const float q = qmin - (0.02f + 2e-6f * mag);
    return !(opac * __expf(-q) < kAlphaMin);  // NaN-safe: keeps on NaN
}

struct V2Tile {
    uint32_t tile_lin, cam;
    uint32_t x, y0;        // this lane's column and first row
    uint32_t row_mask;     // bit j set <=> pixel (x, y0 + j) is inside the image
    float px, py0;
    float rx0, ry0, rx1, ry1;  // rectangle of pixel centres of the tile (clipped to the image)
};

__device__ __forceinline__ V2Tile v2_tile(uint32_t tile_lin, uint32_t tile_width, uint32_t tile_height, uint32_t W,
                                          uint32_t H, unsigned lane) {
    V2Tile t;
    const uint32_t n_tiles = tile_width * tile_height;
    t.tile_lin = tile_lin;
    t.cam = tile_lin / n_tiles;
    const uint32_t tid = tile_lin - t.cam * n_tiles;
    const uint32_t ty = tid / tile_width, tx = tid - ty * tile_width;
    t.x = tx * kV2Tile + (lane & 15);
    t.y0 = ty * kV2Tile + (lane >> 4) * kV2Rows;
    t.px = (float)t.x + 0.5f;
    t.py0 = (float)t.y0 + 0.5f;
    t.row_mask = 0;
    if (t.x < W) {
#pragma unroll
        for (int j = 0; j < kV2Rows; ++j)
            if (t.y0 + j < H) t.row_mask |= 1u << j;
    }
    t.rx0 = (float)(tx * kV2Tile) + 0.5f;
    t.ry0 = (float)(ty * kV2Tile) + 0.5f;
    t.rx1 = (float)min(tx * kV2Tile + kV2Tile - 1, W - 1) + 0.5f;
    t.ry1 = (float)min(ty * kV2Tile + kV2Tile - 1, H - 1) + 0.5f;
    return t;
}

}  // namespace b2s
