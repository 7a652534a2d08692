// raster_bwd.cu — backward of the tile rasterizer (a9).
// Replaces CS/rasterize_to_pixels_bwd.cu:16-277.  Semantics kept exactly: back-to-front
// replay from T_final = 1 - alpha_out, T <- T/(1-alpha); pairs behind last_ids skipped;
// same accept rules as the forward; v_rgb = alpha·T·v_out; v_alpha as in :203-219;
// conic / mean / opacity gradients only when opac·vis <= 0.999; absgrad = |v_xy|.
//
// Design (B200): the reference reduces each of the 9(+2) per-pair gradient values with
// its own 5-step shuffle tree (45–55 SHFL per pair per warp) and then issues 9–11
// scalar atomics from lane 0.  Here the values are reduced with a transposed
// (reduce-scatter) butterfly — 16 SHFL for up to 16 values — which leaves value k in
// lane 2k, so a single RED instruction with <= 16 active lanes commits all of them.
#include "raster_common.cuh"
#include "raster_quad.cuh"

namespace b2s {

// Reduce-scatter over the warp: on return, lane l holds sum over lanes of v[l / (32/P)].
// P is a power of two <= 32.
template <int P>
__device__ __forceinline__ float warp_reduce_scatter(float (&v)[P], const unsigned lane) {
    if (P == 1) return warp_sum(v[0]);
    int o = 16;
#pragma unroll
    for (int h = P / 2; h >= 1; h /= 2) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float keep = upper ? v[i + h] : v[i];
            const float send = upper ? v[i] : v[i + h];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        o >>= 1;
    }
    float r = v[0];
    for (; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

constexpr int next_pow2(int n) { return n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }

// Reduce NV per-lane values across the warp and hand value k (k = 0..NV-1) to `commit`
// on exactly one lane.  Values are processed in chunks of <= 32.
template <int NV, int OFF = 0, typename Commit>
__device__ __forceinline__ void warp_reduce_commit(const float *v, const unsigned lane, Commit &&commit) {
    // chunk sizes are powers of two where that saves shuffles: 9..11 values go as 8 + rest
    constexpr int REM = NV - OFF;
    constexpr int CH = REM > 32 ? 32 : (REM > 8 && REM < 12) ? 8 : REM;
    constexpr int P = next_pow2(CH);
    float w[P];
#pragma unroll
    for (int i = 0; i < P; ++i) w[i] = (i < CH) ? v[OFF + i] : 0.f;
    const float r = warp_reduce_scatter<P>(w, lane);
    constexpr int stride = 32 / P;
    const int k = (int)lane / stride;
    if ((lane % stride) == 0 && k < CH) commit(OFF + k, r);
    if constexpr (OFF + CH < NV) warp_reduce_commit<NV, OFF + CH>(v, lane, commit);
}

// Commit of the quad kernels.  Every lane parks its NV partial values in shared memory (row (slot, k),
// 32 lanes wide, rows padded to 36 floats so that the 128-bit row reads below are conflict-free); once
// SLOTS = 32 / NV pairs are parked, lane l sums row l with eight LDS.128 and issues ONE RED: no
// shuffles (a reduce-scatter butterfly is a chain of 5 dependent SHFL stages per pair: ~10 % of the
// stall samples in ncu r1_d), and the sums of a flush are independent of the compositing chain of the
// following pairs.  The parked values are the NEGATED gradients (the kernel accumulates with negated
// opacity / colours, raster_quad.cuh); the sign is restored by the one subtraction of the flush.
template <int NV>
struct SmemSink {
    static constexpr int SLOTS = 32 / NV;
    static constexpr int ROW = 36;
    static constexpr int FLOATS = SLOTS * NV * ROW;
    float *base;       // destination array of this lane's value index (lane % NV); nullptr: none
    uint32_t stride;
    float *buf;        // [SLOTS * NV][ROW]
    int *gid;          // [SLOTS] Gaussian row of each parked pair
    int cnt;
    int my_slot;

    template <int CDIM>
    __device__ __forceinline__ void init(unsigned lane, uint32_t channels, float *v_colors, float *v_conics,
                                         float *v_means2d, float *v_opacities, float *v_means2d_abs, float *smem_buf,
                                         int *smem_gid) {
        buf = smem_buf; gid = smem_gid; cnt = 0;
        my_slot = (int)lane / NV;
        const int k = (int)lane - my_slot * NV;
        base = nullptr; stride = 0;
        if (my_slot >= SLOTS) return;
        if (k < CDIM) {
            if (k < (int)channels) { base = v_colors + k; stride = channels; }
        } else if (k < CDIM + 3) { base = v_conics + (k - CDIM); stride = 3; }
        else if (k < CDIM + 5) { base = v_means2d + (k - CDIM - 3); stride = 2; }
        else if (k == CDIM + 5) { base = v_opacities; stride = 1; }
        else { base = v_means2d_abs + (k - CDIM - 6); stride = 2; }
    }

    __device__ __forceinline__ void flush(unsigned lane) {
        __syncwarp();
        if ((int)lane < cnt * NV) {
            const float4 *r = reinterpret_cast<const float4 *>(buf + lane * ROW);
            const float4 a = r[0], b = r[1], c = r[2], d = r[3], e = r[4], f = r[5], g4 = r[6], h = r[7];
            const float s0 = (a.x + a.y) + (a.z + a.w) + ((b.x + b.y) + (b.z + b.w));
            const float s1 = (c.x + c.y) + (c.z + c.w) + ((d.x + d.y) + (d.z + d.w));
            const float s2 = (e.x + e.y) + (e.z + e.w) + ((f.x + f.y) + (f.z + f.w));
            const float s3 = (g4.x + g4.y) + (g4.z + g4.w) + ((h.x + h.y) + (h.z + h.w));
            const int g = gid[my_slot];
            if (base != nullptr) atomicAdd(base + (size_t)g * stride, -(s0 + s1) - (s2 + s3));
        }
        __syncwarp();
        cnt = 0;
    }

    // nv: NEGATED per-lane partial gradients of one (tile, Gaussian) pair
    __device__ __forceinline__ void commit(const float (&nv)[NV], unsigned lane, int32_t g) {
        float *row = buf + (cnt * NV) * ROW + lane;
#pragma unroll
        for (int k = 0; k < NV; ++k) row[k * ROW] = nv[k];
        gid[cnt] = g;  // every lane stores the same word: cheaper than predicating one
        if (++cnt == SLOTS) flush(lane);
    }

    __device__ __forceinline__ void finish(unsigned lane) {
        if (cnt > 0) flush(lane);
    }
};

template <int CDIM, bool ABS, int MAXT>
__global__ void __launch_bounds__(MAXT)
raster_bwd_kernel(uint32_t C, uint64_t n_isects, uint32_t channels, const float2 *__restrict__ means2d,
                  const float *__restrict__ conics, const float *__restrict__ colors,
                  const float *__restrict__ opacities, const float *__restrict__ backgrounds,
                  const uint8_t *__restrict__ masks, uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width,
                  uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                  const int32_t *__restrict__ flatten_ids, const float *__restrict__ render_alphas,
                  const int32_t *__restrict__ last_ids, const float *__restrict__ v_render_colors,
                  const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d_abs,
                  float *__restrict__ v_means2d, float *__restrict__ v_conics, float *__restrict__ v_colors,
                  float *__restrict__ v_opacities) {
    constexpr int NV = CDIM + 6 + (ABS ? 2 : 0);
    const TileCoord tc = tile_coord(tile_size, tile_width, tile_height, W, H);
    const uint32_t n_tiles_total = C * tile_width * tile_height;
    if (masks != nullptr && !masks[tc.tile_lin]) return;
    const bool inside = tc.inside;
    const size_t pix = ((size_t)tc.cam * H + tc.i) * W + tc.j;
    const float px = (float)tc.j + 0.5f, py = (float)tc.i + 0.5f;

    const int32_t range_start = tile_offsets[tc.tile_lin];
    const int32_t range_end =
        (tc.tile_lin == n_tiles_total - 1) ? (int32_t)n_isects : tile_offsets[tc.tile_lin + 1];
    const int32_t block_size = (int32_t)blockDim.x;
    const int32_t num_batches = (range_end - range_start + block_size - 1) / block_size;

    extern __shared__ float4 smem4[];
    float4 *rec_a = smem4;                   // {x, y, opacity, conic.a}
    float4 *rec_b = smem4 + block_size;      // {conic.b, conic.c, id (bits), -}
    float *rgbs = reinterpret_cast<float *>(smem4 + 2 * block_size);  // [block_size * CDIM]

    float T_final = 1.f, v_render_a = 0.f;
    float v_render_c[CDIM], buffer[CDIM];
    int32_t bin_final = 0;
#pragma unroll
    for (int k = 0; k < CDIM; ++k) { v_render_c[k] = 0.f; buffer[k] = 0.f; }
    float bg_dot = 0.f;  // sum_k background_k * v_render_c_k
    if (inside) {
        T_final = 1.f - render_alphas[pix];
        bin_final = last_ids[pix];
        v_render_a = v_render_alphas[pix];
#pragma unroll
        for (int k = 0; k < CDIM; ++k)
            if (k < (int)channels) v_render_c[k] = v_render_colors[pix * channels + k];
        if (backgrounds != nullptr) {
#pragma unroll
            for (int k = 0; k < CDIM; ++k)
                if (k < (int)channels) bg_dot += backgrounds[(size_t)tc.cam * channels + k] * v_render_c[k];
        }
    }
    float T = T_final;
    const int32_t tr = (int32_t)threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const int32_t warp_bin_final = warp_max(bin_final);

    for (int32_t b = 0; b < num_batches; ++b) {
        __syncthreads();
        const int32_t batch_end = range_end - 1 - block_size * b;
        const int32_t batch_size = min(block_size, batch_end + 1 - range_start);
        const int32_t idx = batch_end - tr;
        if (idx >= range_start) {
            const int32_t g = flatten_ids[idx];
            const float2 xy = __ldg(means2d + g);
            const float opac = __ldg(opacities + g);
            const float ca = __ldg(conics + 3 * (size_t)g), cb = __ldg(conics + 3 * (size_t)g + 1),
                        cc = __ldg(conics + 3 * (size_t)g + 2);
            rec_a[tr] = make_float4(xy.x, xy.y, opac, ca);
            rec_b[tr] = make_float4(cb, cc, __int_as_float(g), 0.f);
#pragma unroll
            for (int k = 0; k < CDIM; ++k)
                rgbs[tr * CDIM + k] = (k < (int)channels) ? __ldg(colors + (size_t)g * channels + k) : 0.f;
        }
        __syncthreads();
        for (int32_t t = max(0, batch_end - warp_bin_final); t < batch_size; ++t) {
            bool valid = inside && (batch_end - t <= bin_final);
            float alpha = 0.f, opac = 0.f, vis = 0.f, dx = 0.f, dy = 0.f, ca = 0.f, cb = 0.f, cc = 0.f;
            if (valid) {
                const float4 ra = rec_a[t];
                const float4 rb = rec_b[t];
                opac = ra.z; ca = ra.w; cb = rb.x; cc = rb.y;
                dx = ra.x - px; dy = ra.y - py;
                const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                vis = __expf(-sigma);
                alpha = fminf(kAlphaMax, opac * vis);
                if (sigma < 0.f || alpha < kAlphaMin) valid = false;
            }
            if (!__any_sync(0xffffffffu, valid)) continue;
            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = 0.f;
            if (valid) {
                const float ra_ = 1.f / (1.f - alpha);
                T *= ra_;
                const float fac = alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < CDIM; ++k) {
                    const float c = rgbs[t * CDIM + k];
                    v[k] = fac * v_render_c[k];
                    v_alpha += (c * T - buffer[k] * ra_) * v_render_c[k];
                    buffer[k] += c * fac;
                }
                v_alpha += T_final * ra_ * v_render_a;
                if (backgrounds != nullptr) v_alpha += -T_final * ra_ * bg_dot;
                if (opac * vis <= kAlphaMax) {
                    const float v_sigma = -opac * vis * v_alpha;
                    v[CDIM + 0] = 0.5f * v_sigma * dx * dx;
                    v[CDIM + 1] = v_sigma * dx * dy;
                    v[CDIM + 2] = 0.5f * v_sigma * dy * dy;
                    const float vx = v_sigma * (ca * dx + cb * dy);
                    const float vy = v_sigma * (cb * dx + cc * dy);
                    v[CDIM + 3] = vx;
                    v[CDIM + 4] = vy;
                    v[CDIM + 5] = vis * v_alpha;
                    if (ABS) { v[CDIM + 6] = fabsf(vx); v[CDIM + 7] = fabsf(vy); }
                }
            }
            const int32_t g = __float_as_int(rec_b[t].z);
            warp_reduce_commit<NV>(v, lane, [&](int k, float val) {
                float *dst;
                if (k < CDIM) {
                    if (k >= (int)channels) return;
                    dst = v_colors + (size_t)g * channels + k;
                } else if (k < CDIM + 3) dst = v_conics + 3 * (size_t)g + (k - CDIM);
                else if (k < CDIM + 5) dst = v_means2d + 2 * (size_t)g + (k - CDIM - 3);
                else if (k == CDIM + 5) dst = v_opacities + g;
                else dst = v_means2d_abs + 2 * (size_t)g + (k - CDIM - 6);
                atomicAdd(dst, val);
            });
        }
    }
}

template <int CDIM, bool ABS>
static void launch_bwd(uint32_t C, uint64_t n_isects, uint32_t channels, const float *means2d, const float *conics,
                       const float *colors, const float *opacities, const float *backgrounds, const uint8_t *masks,
                       uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                       const int32_t *tile_offsets, const int32_t *flatten_ids, const float *render_alphas,
                       const int32_t *last_ids, const float *v_render_colors, const float *v_render_alphas,
                       float *v_means2d_abs, float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                       cudaStream_t st) {
    const uint32_t threads = ((tile_size * tile_size + 31) / 32) * 32;
    const uint32_t grid = C * tile_width * tile_height;
    const size_t smem = (size_t)threads * (2 * sizeof(float4) + CDIM * sizeof(float));
    if (threads <= 256) {
        auto kern = raster_bwd_kernel<CDIM, ABS, 256>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, st>>>(C, n_isects, channels, reinterpret_cast<const float2 *>(means2d), conics,
                                          colors, opacities, backgrounds, masks, W, H, tile_size, tile_width,
                                          tile_height, tile_offsets, flatten_ids, render_alphas, last_ids,
                                          v_render_colors, v_render_alphas, v_means2d_abs, v_means2d, v_conics,
                                          v_colors, v_opacities);
    } else {
        auto kern = raster_bwd_kernel<CDIM, ABS, 1024>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, st>>>(C, n_isects, channels, reinterpret_cast<const float2 *>(means2d), conics,
                                          colors, opacities, backgrounds, masks, W, H, tile_size, tile_width,
                                          tile_height, tile_offsets, flatten_ids, render_alphas, last_ids,
                                          v_render_colors, v_render_alphas, v_means2d_abs, v_means2d, v_conics,
                                          v_colors, v_opacities);
    }
}


// ---------------------------------------------------------------------------------------
// quad kernels: warp-per-tile, 8x8 quads, packed fp32x2 (raster_quad.cuh).
// Per (tile, Gaussian): each lane folds its pixels into CDIM colour sums and, per column
// half (dx is shared by the pixels of a half), the moments
//   W0 = sum v_sigma,  W1 = sum v_sigma·dy,  W2 = sum v_sigma·dy²
// converts them to the 9 gradient values and parks them in shared memory (SmemSink).
// Everything is accumulated NEGATED (the record carries -opacity and -colour, raster_quad.cuh)
// and the sign is restored once, in the flush.
// ---------------------------------------------------------------------------------------
struct BwdAcc {
    float2 nW0, nW1, nW2;  // negated moments of one column half, one entry per pixel row pair
};

// One Gaussian against one quad.  State per pixel: T, ntb = -(T_final (v_alpha_out - bg·v_c) -
// sum_k buffer_k v_c_k), which folds the reference's per-channel `buffer`
// (CS/rasterize_to_pixels_bwd.cu:203-241): v_alpha only ever needs that dot product.
// SLOW: pair flagged kNonPD (raster_quad.cuh) — the `sigma < 0` rejection and the 0.999 clamp of alpha
// (with its `opac·vis <= 0.999` gradient gate, :221) are applied literally; for every other pair neither
// can fire: alpha = opac·vis itself, and v_sigma = -alpha·v_alpha needs no gate.
template <int CDIM, bool ABS, bool SLOW>
__device__ __forceinline__ bool bwd_quad(float2 &T2, float2 &ntb2, const float2 (&v_c2)[CDIM], const int32_t (&binf)[2],
                                         float2 (&nvacc2)[CDIM], BwdAcc &acc, float2 &abs2x, float2 &abs2y,
                                         const float2 dy2, const float2 ndy2, const float dx, const float hA,
                                         const float cb, const float nA, const float B, const float hC,
                                         const float nopac, const float (&ncol)[4], const int32_t idx) {
    const float2 u2 = __ffma2_rn(bc2(hC), dy2, bc2(B));
    const float2 ns2 = __ffma2_rn(ndy2, u2, bc2(nA));  // -sigma'
    const float2 nov2 = __fmul2_rn(bc2(nopac), make_float2(ex2_approx(ns2.x), ex2_approx(ns2.y)));  // -opac·vis
    const float nal0 = SLOW ? fmaxf(-kAlphaMax, nov2.x) : nov2.x, nal1 = SLOW ? fmaxf(-kAlphaMax, nov2.y) : nov2.y;  // -alpha
    bool ok0 = (nal0 <= -kAlphaMin) && (idx <= binf[0]), ok1 = (nal1 <= -kAlphaMin) && (idx <= binf[1]);
    if (SLOW) { ok0 = ok0 && !(ns2.x > 0.f); ok1 = ok1 && !(ns2.y > 0.f); }
    // a rejected pixel runs with alpha = 0: ra = 1, fac = 0, v_sigma = 0
    const float2 nae2 = make_float2(ok0 ? nal0 : 0.f, ok1 ? nal1 : 0.f);
    const float2 om2 = __fadd2_rn(nae2, bc2(1.f));
    const float2 ra2 = make_float2(rcp_approx(om2.x), rcp_approx(om2.y));
    T2 = __fmul2_rn(T2, ra2);
    const float2 nfac2 = __fmul2_rn(nae2, T2);  // -alpha T
    float2 ncv2 = make_float2(0.f, 0.f);        // -sum_k c_k v_c_k
#pragma unroll
    for (int k = 0; k < CDIM; ++k) {
        nvacc2[k] = __ffma2_rn(nfac2, v_c2[k], nvacc2[k]);
        ncv2 = (k == 0) ? __fmul2_rn(bc2(ncol[0]), v_c2[0]) : __ffma2_rn(bc2(ncol[k]), v_c2[k], ncv2);
    }
    const float2 nv_alpha2 = __ffma2_rn(T2, ncv2, __fmul2_rn(ra2, ntb2));  // -v_alpha
    ntb2 = __ffma2_rn(nfac2, ncv2, ntb2);
    float2 nvs2;  // ov·v_alpha = -v_sigma
    if (SLOW) {
        nvs2 = __fmul2_rn(nov2, nv_alpha2);
        nvs2.x = (ok0 && nov2.x >= -kAlphaMax) ? nvs2.x : 0.f;
        nvs2.y = (ok1 && nov2.y >= -kAlphaMax) ? nvs2.y : 0.f;
    } else {
        nvs2 = __fmul2_rn(nae2, nv_alpha2);  // alpha is unclamped and already 0 for a rejected pixel
    }
    const float2 nwy2 = __fmul2_rn(nvs2, dy2);
    acc.nW0 = __fadd2_rn(acc.nW0, nvs2);
    acc.nW1 = __fadd2_rn(acc.nW1, nwy2);
    acc.nW2 = __ffma2_rn(nwy2, dy2, acc.nW2);
    if (ABS) {
        // |v_sigma (a' dx + b' dy)|, |v_sigma (b' dx + c' dy)| (still scaled by log2 e)
        const float2 gx2 = __ffma2_rn(bc2(cb), dy2, bc2(2.f * hA * dx));
        const float2 gy2 = __ffma2_rn(bc2(2.f * hC), dy2, bc2(cb * dx));
        abs2x.x += fabsf(nvs2.x * gx2.x); abs2x.y += fabsf(nvs2.y * gx2.y);
        abs2y.x += fabsf(nvs2.x * gy2.x); abs2y.y += fabsf(nvs2.y * gy2.y);
    }
    return ok0 || ok1;
}

// `quad_masks` (optional): the per-pair quad masks stored by the forward kernel; when given, staging is
// a byte load instead of the exact rectangle test of quad_mask.
template <int CDIM, bool ABS, int MINB>
__global__ void __launch_bounds__(32, MINB)
raster_bwd_quad_kernel(uint32_t n_tiles_total, uint64_t n_isects, uint32_t channels, const float4 *__restrict__ rec,
                       const float *__restrict__ backgrounds, const uint8_t *__restrict__ masks, uint32_t W,
                       uint32_t H, uint32_t tile_width, uint32_t tile_height,
                       const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                       const uint8_t *__restrict__ quad_masks,
                       const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                       const float *__restrict__ v_render_colors, const float *__restrict__ v_render_alphas,
                       float *__restrict__ v_means2d_abs, float *__restrict__ v_means2d, float *__restrict__ v_conics,
                       float *__restrict__ v_colors, float *__restrict__ v_opacities) {
    constexpr int NV = CDIM + 6 + (ABS ? 2 : 0);
    constexpr int NQ = 4;         // one warp owns the whole tile
    constexpr int NQY = NQ / 2;   // quad rows
    __shared__ float4 s_rec[32 * 3];
    __shared__ int4 s_im[32];  // {sorted index, quad mask, Gaussian row, -}
    using SSink = SmemSink<NV>;
    __shared__ __align__(16) float s_red[SSink::FLOATS];
    __shared__ int s_gid[SSink::SLOTS];
    __shared__ float4 s_pre_all[32 * 3];  // records of the batch in flight
    const unsigned lane = threadIdx.x & 31;
    const float4 *s_pre = s_pre_all + 3 * lane;
    const uint32_t pre_addr = (uint32_t)__cvta_generic_to_shared(s_pre);
    const uint32_t tile_lin = blockIdx.x;
    if (masks != nullptr && !masks[tile_lin]) return;
    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end = (tile_lin == n_tiles_total - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    if (range_end <= range_start) return;
    const QuadTile tc = quad_tile<NQ>(tile_lin, tile_width, tile_height, lane, 0);
    const size_t cam_pix = (size_t)tc.cam * H * W;

    float2 T2[NQ], ntb2[NQ], v_c2[NQ][CDIM];
    int32_t binf[NQ][2];
    int32_t max_bin = -1;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float Tq[2] = {1.f, 1.f}, tb[2] = {0.f, 0.f}, vc[2][CDIM];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            binf[q][j] = -1;
#pragma unroll
            for (int k = 0; k < CDIM; ++k) vc[j][k] = 0.f;
            const uint32_t x = tc.x + 8u * (q & 1), y = tc.y + 8u * (q >> 1) + 4u * j;
            if (x < W && y < H) {
                const size_t p = cam_pix + (size_t)y * W + x;
                const float T_final = 1.f - render_alphas[p];
                Tq[j] = T_final;
                binf[q][j] = last_ids[p];
                float bg_dot = 0.f;
#pragma unroll
                for (int k = 0; k < CDIM; ++k) {
                    if (k < (int)channels) {
                        vc[j][k] = v_render_colors[p * channels + k];
                        if (backgrounds != nullptr) bg_dot += backgrounds[(size_t)tc.cam * channels + k] * vc[j][k];
                    }
                }
                tb[j] = T_final * (v_render_alphas[p] - bg_dot);
                max_bin = max(max_bin, binf[q][j]);
            }
        }
        T2[q] = make_float2(Tq[0], Tq[1]);
        ntb2[q] = make_float2(-tb[0], -tb[1]);
#pragma unroll
        for (int k = 0; k < CDIM; ++k) v_c2[q][k] = make_float2(vc[0][k], vc[1][k]);
    }
    max_bin = warp_max(max_bin);
    // nothing behind the last contributor of any pixel of this warp's region matters
    const int32_t hi0 = min(range_end - 1, max_bin);
    if (hi0 < range_start) return;

    float2 pyc2[NQY], npyc2[NQY];
#pragma unroll
    for (int qy = 0; qy < NQY; ++qy) {
        pyc2[qy] = make_float2(tc.py + 8.f * qy, tc.py + 8.f * qy + 4.f);
        npyc2[qy] = make_float2(-pyc2[qy].x, -pyc2[qy].y);
    }
    const float pxa = tc.px, pxb = tc.px + 8.f;

    SSink ssink;
    ssink.template init<CDIM>(lane, channels, v_colors, v_conics, v_means2d, v_opacities, v_means2d_abs, s_red, s_gid);

    // record (and stored quad mask) prefetch one batch ahead, the record straight into shared memory;
    // sorted id two batches ahead
    int32_t my_idx = hi0 - (int32_t)lane, my_g = 0, g_next = 0;
    uint32_t qm_next = 0;
    if (my_idx >= range_start) {
        my_g = flatten_ids[my_idx];
        prefetch_record(pre_addr, rec, my_g);
        if (quad_masks != nullptr) qm_next = quad_masks[my_idx];
    }
    if (my_idx - 32 >= range_start) g_next = flatten_ids[my_idx - 32];
    uint32_t act = 0;  // warp-uniform: quads with a pixel whose last contributor is inside the batches seen so far
    for (int32_t hi = hi0; hi >= range_start; hi -= 32) {
        const int32_t lo = max(hi - 31, range_start);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (!(act >> q & 1) && __any_sync(0xffffffffu, max(binf[q][0], binf[q][1]) >= lo)) act |= 1u << q;
        uint32_t my_mask = 0;
        float4 r0, r1, r2;
        cp_async_wait_all();
        if (my_idx >= range_start) {
            r0 = s_pre[0]; r1 = s_pre[1]; r2 = s_pre[2];
            my_mask = quad_masks != nullptr
                          ? unpack_quad_mask((uint8_t)qm_next)
                          : quad_mask<NQ>(r0.x, r0.y, r0.z, r0.w, r1.x, r2.z, r1.y, tc.ox, tc.oy, W, H);
            if ((my_mask & act) == 0) my_mask = 0; else my_mask &= (act | kNonPD);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, my_mask != 0);
        const int n = __popc(bal);
        __syncwarp();
        if (my_mask != 0) {
            const int pos = __popc(bal & ((1u << lane) - 1u));
            s_rec[3 * pos] = r0; s_rec[3 * pos + 1] = r1;
            s_rec[3 * pos + 2] = make_float4(r2.x, r2.y, -r0.y, 0.f);
            s_im[pos] = make_int4(my_idx, (int)my_mask, my_g, 0);
        }
        __syncwarp();
        my_idx = hi - 32 - (int32_t)lane;
        if (my_idx >= range_start) {
            my_g = g_next;
            prefetch_record(pre_addr, rec, my_g);
            if (quad_masks != nullptr) qm_next = quad_masks[my_idx];
        }
        if (my_idx - 32 >= range_start) g_next = flatten_ids[my_idx - 32];
        for (int t = 0; t < n; ++t) {
            const float4 a = s_rec[3 * t], b4 = s_rec[3 * t + 1], c4 = s_rec[3 * t + 2];
            const int4 im = s_im[t];
            const int32_t idx = im.x;
            const uint32_t m = (uint32_t)im.y;
            const float dxa = a.x - pxa, dxb = a.x - pxb;
            const float hA = a.z, cb = a.w, hC = b4.x, nopac = b4.y;
            const float nAa = -(hA * dxa) * dxa, Ba = cb * dxa, nAb = -(hA * dxb) * dxb, Bb = cb * dxb;
            const float ncol[4] = {b4.z, b4.w, c4.x, c4.y};
            float2 dy2[NQY], ndy2[NQY];
#pragma unroll
            for (int qy = 0; qy < NQY; ++qy) {
                ndy2[qy] = __fadd2_rn(pyc2[qy], bc2(c4.z));   // p_y - g_y
                dy2[qy] = __fadd2_rn(npyc2[qy], bc2(a.y));    // g_y - p_y
            }
            float2 nvacc2[CDIM];
#pragma unroll
            for (int k = 0; k < CDIM; ++k) nvacc2[k] = make_float2(0.f, 0.f);
            BwdAcc acc[2];
            acc[0].nW0 = acc[0].nW1 = acc[0].nW2 = acc[1].nW0 = acc[1].nW1 = acc[1].nW2 = make_float2(0.f, 0.f);
            float2 abs2x = make_float2(0.f, 0.f), abs2y = abs2x;
            bool hit = false;
#define B2S_BQ(SLOW_, q_)                                                                                            \
    hit |= bwd_quad<CDIM, ABS, SLOW_>(T2[q_], ntb2[q_], v_c2[q_], binf[q_], nvacc2, acc[(q_) & 1], abs2x, abs2y,        \
                                      dy2[(q_) >> 1], ndy2[(q_) >> 1], ((q_) & 1) ? dxb : dxa, hA, cb,                 \
                                      ((q_) & 1) ? nAb : nAa, ((q_) & 1) ? Bb : Ba, hC, nopac, ncol, idx)
            if (m & kNonPD) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (m >> q & 1) B2S_BQ(true, q);  // warp-uniform
            } else {
                // both quads of a quad row in one basic block when both are reached (ILP 2)
#pragma unroll
                for (int h = 0; h < NQY; ++h) {
                    const uint32_t mm = (m >> (2 * h)) & 3u;
                    if (mm == 3u) { B2S_BQ(false, 2 * h); B2S_BQ(false, 2 * h + 1); }
                    else if (mm == 1u) B2S_BQ(false, 2 * h);
                    else if (mm == 2u) B2S_BQ(false, 2 * h + 1);
                }
            }
#undef B2S_BQ
            if (!__any_sync(0xffffffffu, hit)) continue;
            // negated moments -> NEGATED gradients (conic entries in the record are scaled by log2 e);
            // the flush restores the sign
            float nv[NV];
#pragma unroll
            for (int k = 0; k < CDIM; ++k) nv[k] = nvacc2[k].x + nvacc2[k].y;
            const float nW0a = acc[0].nW0.x + acc[0].nW0.y, nW0b = acc[1].nW0.x + acc[1].nW0.y;
            const float nW1a = acc[0].nW1.x + acc[0].nW1.y, nW1b = acc[1].nW1.x + acc[1].nW1.y;
            const float nW2s = (acc[0].nW2.x + acc[0].nW2.y) + (acc[1].nW2.x + acc[1].nW2.y);
            const float nua = dxa * nW0a, nub = dxb * nW0b;
            const float nS1x = nua + nub, nS1y = nW1a + nW1b;
            nv[CDIM + 0] = 0.5f * fmaf(dxa, nua, dxb * nub);               // 1/2 sum v_sigma dx²
            nv[CDIM + 1] = fmaf(dxa, nW1a, dxb * nW1b);                    // sum v_sigma dx dy
            nv[CDIM + 2] = 0.5f * nW2s;                                    // 1/2 sum v_sigma dy²
            nv[CDIM + 3] = kInvLog2e * fmaf(2.f * hA, nS1x, cb * nS1y);    // sum v_sigma (a dx + b dy)
            nv[CDIM + 4] = kInvLog2e * fmaf(cb, nS1x, 2.f * hC * nS1y);    // sum v_sigma (b dx + c dy)
            nv[CDIM + 5] = (nW0a + nW0b) * rcp_approx(nopac);              // sum vis·v_alpha = -S0 / opac
            if (ABS) {
                nv[CDIM + 6] = -kInvLog2e * (abs2x.x + abs2x.y);
                nv[CDIM + 7] = -kInvLog2e * (abs2y.x + abs2y.y);
            }
            ssink.commit(nv, lane, im.z);
        }
    }
    ssink.finish(lane);
}

template <int CDIM, bool ABS>
static void launch_bwd_quad(uint32_t C, uint64_t n_isects, uint32_t channels, const float4 *rec,
                            const float *backgrounds, const uint8_t *masks, uint32_t W, uint32_t H,
                            uint32_t tile_width, uint32_t tile_height, const int32_t *tile_offsets,
                            const int32_t *flatten_ids, const uint8_t *quad_masks, const float *render_alphas,
                            const int32_t *last_ids, const float *v_render_colors, const float *v_render_alphas,
                            float *v_means2d_abs, float *v_means2d, float *v_conics, float *v_colors,
                            float *v_opacities, cudaStream_t st) {
    const uint32_t total = C * tile_width * tile_height;
#define B2S_BWDQ(MINB_)                                                                                              \
    raster_bwd_quad_kernel<CDIM, ABS, MINB_><<<total, 32, 0, st>>>(                                                   \
        total, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, \
        quad_masks, render_alphas, last_ids, v_render_colors, v_render_alphas, v_means2d_abs, v_means2d, v_conics,    \
        v_colors, v_opacities)
#ifdef B2S_TUNING
    switch (tuning_variant()) {
        case 1: B2S_BWDQ(18); return;   // 112 registers
        case 2: B2S_BWDQ(20); return;   // 96 registers
        default: break;
    }
#endif
    B2S_BWDQ(16);
#undef B2S_BWDQ
}

}  // namespace b2s

using namespace b2s;

extern "C" int b200splat_rasterize_bwd(uint32_t C, uint32_t n_gauss, uint64_t n_isects, uint32_t channels,
                                       const float *means2d, const float *conics, const float *colors,
                                       const float *opacities, const float *backgrounds, const uint8_t *masks,
                                       uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width,
                                       uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                                       const void *records, const uint8_t *quad_masks,
                                       const float *render_alphas, const int32_t *last_ids,
                                       const float *v_render_colors, const float *v_render_alphas,
                                       float *v_means2d_abs, float *v_means2d, float *v_conics, float *v_colors,
                                       float *v_opacities, void *stream) {
    const char *where = "b200splat_rasterize_bwd";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    (void)n_gauss;
    if (records != nullptr) {
        B2S_REQUIRE(tile_size == kQTile && channels >= 1 && channels <= 4, where,
                    "packed records are only valid for tile_size 16 and <= 4 channels");
        B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
        if ((uint64_t)C * tile_width * tile_height == 0 || n_isects == 0) return 0;
        const float4 *rec = reinterpret_cast<const float4 *>(records);
        cudaStream_t st2 = (cudaStream_t)stream;
        const bool ab = v_means2d_abs != nullptr;
#define B2S_BWD2(D)                                                                                                    \
    case D:                                                                                                            \
        if (ab)                                                                                                        \
            launch_bwd_quad<D, true>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height,      \
                                   tile_offsets, flatten_ids, quad_masks, render_alphas, last_ids, v_render_colors,    \
                                   v_render_alphas, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities, st2);   \
        else                                                                                                           \
            launch_bwd_quad<D, false>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height,     \
                                    tile_offsets, flatten_ids, quad_masks, render_alphas, last_ids, v_render_colors,   \
                                    v_render_alphas, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities, st2);  \
        break;
        switch (channels) { B2S_BWD2(1) B2S_BWD2(2) B2S_BWD2(3) B2S_BWD2(4) }
#undef B2S_BWD2
        B2S_CHECK_LAUNCH(where);
        return 0;
    }
    B2S_REQUIRE(tile_size >= 1 && tile_size <= 32, where, "tile_size must be in [1, 32]");
    B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
    const int cdim = pick_cdim(channels);
    B2S_REQUIRE(channels >= 1 && cdim > 0, where, "unsupported number of color channels (1..33)");
    if ((uint64_t)C * tile_width * tile_height == 0 || n_isects == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool absgrad = v_means2d_abs != nullptr;
#define B2S_BWD(D)                                                                                                  \
    case D:                                                                                                         \
        if (absgrad)                                                                                                \
            launch_bwd<D, true>(C, n_isects, channels, means2d, conics, colors, opacities, backgrounds, masks, W, H, \
                                tile_size, tile_width, tile_height, tile_offsets, flatten_ids, render_alphas,       \
                                last_ids, v_render_colors, v_render_alphas, v_means2d_abs, v_means2d, v_conics,     \
                                v_colors, v_opacities, st);                                                         \
        else                                                                                                        \
            launch_bwd<D, false>(C, n_isects, channels, means2d, conics, colors, opacities, backgrounds, masks, W,  \
                                 H, tile_size, tile_width, tile_height, tile_offsets, flatten_ids, render_alphas,   \
                                 last_ids, v_render_colors, v_render_alphas, v_means2d_abs, v_means2d, v_conics,    \
                                 v_colors, v_opacities, st);                                                        \
        break;
    switch (cdim) {
        B2S_BWD(1) B2S_BWD(2) B2S_BWD(3) B2S_BWD(4) B2S_BWD(5) B2S_BWD(8) B2S_BWD(9) B2S_BWD(16) B2S_BWD(17)
        B2S_BWD(32) B2S_BWD(33)
    }
#undef B2S_BWD
    B2S_CHECK_LAUNCH(where);
    return 0;
}
