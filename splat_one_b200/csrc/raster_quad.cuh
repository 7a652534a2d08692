// raster_quad.cuh — shared pieces of the B200 "warp-per-tile" rasterizer (tile_size 16,
// <= 4 channels: the RGB / RGB+depth cases splat_one renders).
//
// Mapping: a CTA is one 16x16 tile.  The tile is cut into four 8x8 "quads"; a warp owns NQ
// of them (NQ = 4: one warp per tile; NQ = 2: two warps, upper / lower 16x8 half, working
// independently).  Lane l owns the two pixels (l & 7, l >> 3) and (l & 7, (l >> 3) + 4) of
// EVERY quad of its warp; their state (T, colour, last id) lives in registers as float2
// pairs, and all per-pixel arithmetic of a quad is issued as packed fp32x2 instructions
// (FFMA2 / FMUL2 / FADD2, new in sm_100): one issue slot does two pixels.
//
// Why quads: the reference's intersection list is a 3-sigma bounding SQUARE per Gaussian;
// measured on the benchmark scene only 28 % of the (pair, pixel) evaluations it implies
// can pass the alpha >= 1/255 test, and 31 % of the pairs have no such pixel at all.  At
// staging time each lane takes one Gaussian of the batch and computes, exactly (minimum
// of the quadratic form over the rectangle of pixel centres, plus a rounding slack), which
// quads it can reach.  Pairs with an empty mask are dropped; the others are compacted into
// shared memory with their mask, and the compositing loop runs only the quads in the mask
// (a warp-uniform branch per quad — no divergence).  Culling never changes a result: a
// culled (pixel, Gaussian) pair is one the per-pixel alpha test would have rejected.
//
// Arithmetic: the packed record carries the conic pre-multiplied by log2(e) and the
// opacity / colours NEGATED (fp32x2 instructions have no negate modifier), so with
// ndy = -dy:   -sigma' = ndy (B + C dy) - A,   -alpha = max(-0.999, (-o) ex2(-sigma')),
// 1 - alpha = (-alpha) + 1,   colour += (-c)(-alpha T).  A = a'/2 dx², B = b' dx, C = c'/2.
#pragma once
#include "raster_common.cuh"

namespace b2s {

constexpr int kQTile = 16;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kInvLog2e = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// record of Gaussian g (48 bytes, three 16-byte loads):
//   rec[3g]   = {x, y, L·a/2, L·b}          L = log2(e)
//   rec[3g+1] = {L·c/2, -opacity, -c0, -c1}
//   rec[3g+2] = {-c2, -c3, log2(255·opacity), 0}
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
    const float op = opacities[g];
    // alpha = op·2^(-sigma') >= 1/255  <=>  sigma' <= log2(255·op); -inf / NaN for op <= 0 / NaN
    const float tau = log2f(255.f * op);
    rec[3 * (size_t)g] = make_float4(xy.x, xy.y, 0.5f * kLog2e * a, kLog2e * b);
    rec[3 * (size_t)g + 1] = make_float4(0.5f * kLog2e * c, -op, -col[0], -col[1]);
    rec[3 * (size_t)g + 2] = make_float4(-col[2], -col[3], tau, 0.f);
}

// Minimum of q(dx,dy) = hA dx² + b dx dy + hC dy² (positive definite) over the rectangle
// dx in [dx0,dx1], dy in [dy0,dy1].  ihC = -b/(2 hC), ihA = -b/(2 hA).
__device__ __forceinline__ float rect_qmin(float hA, float b, float hC, float ihA, float ihC, float dx0, float dx1,
                                           float dy0, float dy1) {
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) return 0.f;
    // the minimiser lies on the boundary: 4 edges, 1-D clamp on each
    float t, q, qmin;
    t = fminf(fmaxf(ihC * dx0, dy0), dy1); qmin = hA * dx0 * dx0 + t * (b * dx0 + hC * t);
    t = fminf(fmaxf(ihC * dx1, dy0), dy1); q = hA * dx1 * dx1 + t * (b * dx1 + hC * t); qmin = fminf(qmin, q);
    t = fminf(fmaxf(ihA * dy0, dx0), dx1); q = hC * dy0 * dy0 + t * (b * dy0 + hA * t); qmin = fminf(qmin, q);
    t = fminf(fmaxf(ihA * dy1, dx0), dx1); q = hC * dy1 * dy1 + t * (b * dy1 + hA * t); qmin = fminf(qmin, q);
    return qmin;
}

// mask flag: this pair takes the slow compositing path, which applies the reference's two rarely-active
// rules literally — `sigma < 0` rejection (conic not positive definite) and the `min(0.999, .)` clamp of
// alpha (only reachable when opacity > 0.998: alpha = opacity·2^(-sigma') <= opacity otherwise, and
// ex2.approx is within 2 ulp).  Every other pair runs code with both compiled out.
constexpr uint32_t kNonPD = 0x80000000u;
constexpr float kClampFreeOpacity = 0.998f;

// one byte per listed (tile, Gaussian) pair, written by the forward kernel for every pair it stages and
// read back by the backward kernel instead of re-running quad_mask: bits 0-3 = reachable quads
// (geometry only), bit 7 = slow path
__device__ __forceinline__ uint8_t pack_quad_mask(uint32_t m) { return (uint8_t)((m & 0xFu) | ((m & kNonPD) ? 0x80u : 0u)); }
__device__ __forceinline__ uint32_t unpack_quad_mask(uint8_t b) { return (uint32_t)(b & 0xFu) | ((b & 0x80u) ? kNonPD : 0u); }

// Which of the NQ 8x8 quads of the region at pixel origin (ox, oy) can this Gaussian reach
// with alpha >= 1/255 at some pixel centre?  Quad q sits at (ox + 8 (q & 1), oy + 8 (q >> 1)).
// Conservative: answers "yes" whenever unsure (non positive-definite conic, NaNs); never
// "yes" for a quad outside the image.  Bit 31 flags a non positive-definite conic.
template <int NQ>
__device__ __forceinline__ uint32_t quad_mask(float gx, float gy, float hA, float b, float hC, float tau, float nopac,
                                              uint32_t ox, uint32_t oy, uint32_t W, uint32_t H) {
    const bool pd = (hA > 0.f) && (hC > 0.f) && (4.f * hA * hC - b * b > 0.f);
    const float ihC = -0.5f * b / hC, ihA = -0.5f * b / hA;
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const uint32_t bx = ox + 8u * (q & 1), by = oy + 8u * (q >> 1);
        if (bx >= W || by >= H) continue;
        const float x0 = (float)bx + 0.5f, y0 = (float)by + 0.5f;
        const float x1 = (float)min(bx + 7u, W - 1u) + 0.5f, y1 = (float)min(by + 7u, H - 1u) + 0.5f;
        const float dx0 = gx - x1, dx1 = gx - x0, dy0 = gy - y1, dy1 = gy - y0;
        const float qmin = rect_qmin(hA, b, hC, ihA, ihC, dx0, dx1, dy0, dy1);
        // slack: fp32 rounding of the per-pixel sigma' (terms up to `mag`) plus ~2 % in alpha
        const float mx = fmaxf(fabsf(dx0), fabsf(dx1)), my = fmaxf(fabsf(dy0), fabsf(dy1));
        const float mag = hA * mx * mx + hC * my * my + fabsf(b) * mx * my;
        const bool drop = pd && (qmin - (0.03f + 2e-6f * mag) > tau);  // NaN-safe: keeps on NaN
        if (!drop) m |= 1u << q;
    }
    if (m != 0 && (!pd || !(-nopac <= kClampFreeOpacity))) m |= kNonPD;  // NaN opacity: slow path too
    return m;
}

__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }

// Asynchronous 16-byte global -> shared copies (LDGSTS): the next batch's records travel to
// shared memory without passing through registers, so nothing in the compositing loop waits on
// them (with register prefetch ptxas placed a move out of the load's destination right behind
// the load: 6-11 % of the raster kernels' stall samples, ncu r1_d).
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// one 48-byte record -> this lane's slot of the prefetch buffer
__device__ __forceinline__ void prefetch_record(uint32_t slot_addr, const float4 *__restrict__ rec, int32_t g) {
    const float4 *src = rec + 3 * (size_t)g;
    cp_async16(slot_addr, src);
    cp_async16(slot_addr + 16, src + 1);
    cp_async16(slot_addr + 32, src + 2);
    cp_async_commit();
}

struct QuadTile {
    uint32_t cam;
    uint32_t ox, oy;       // pixel origin of this warp's region
    uint32_t x, y;         // this lane's first pixel in quad 0
    float px, py;          // its centre
};

// `sub`: which part of the tile this warp owns (0 when NQ == 4; 0/1 = upper/lower 16x8 half
// when NQ == 2)
template <int NQ>
__device__ __forceinline__ QuadTile quad_tile(uint32_t tile_lin, uint32_t tile_width, uint32_t tile_height,
                                              unsigned lane, uint32_t sub) {
    QuadTile t;
    const uint32_t n_tiles = tile_width * tile_height;
    t.cam = tile_lin / n_tiles;
    const uint32_t tid = tile_lin - t.cam * n_tiles;
    const uint32_t ty = tid / tile_width, tx = tid - ty * tile_width;
    t.ox = tx * kQTile;
    t.oy = ty * kQTile + sub * 8u;
    t.x = t.ox + (lane & 7);
    t.y = t.oy + (lane >> 3);
    t.px = (float)t.x + 0.5f;
    t.py = (float)t.y + 0.5f;
    return t;
}

}  // namespace b2s
