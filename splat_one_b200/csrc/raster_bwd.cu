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
#include "raster_v3.cuh"

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
// v3: warp-per-tile, 8 sub-block slots per lane, exact sub-block culling (raster_v3.cuh).
// Per (tile, Gaussian): each lane folds its (up to) 8 pixels into CDIM colour sums and, per
// column half (dx is shared by the four slots of a half), the moments
//   W0 = sum v_sigma,  W1 = sum v_sigma·dy,  W2 = sum v_sigma·dy²
// converts them to the 9 gradient values, and ONE reduce-scatter butterfly + one RED per
// value commits them.
// ---------------------------------------------------------------------------------------
template <int CDIM, bool ABS, int NS, int MINB>
__global__ void __launch_bounds__(32 * (kV3Slots / NS), MINB)
raster_bwd_v3_kernel(uint32_t n_tiles_total, uint64_t n_isects, uint32_t channels, const float4 *__restrict__ rec,
                     const float *__restrict__ backgrounds, const uint8_t *__restrict__ masks, uint32_t W, uint32_t H,
                     uint32_t tile_width, uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                     const int32_t *__restrict__ flatten_ids, const float *__restrict__ render_alphas,
                     const int32_t *__restrict__ last_ids, const float *__restrict__ v_render_colors,
                     const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d_abs,
                     float *__restrict__ v_means2d, float *__restrict__ v_conics, float *__restrict__ v_colors,
                     float *__restrict__ v_opacities) {
    constexpr int NV = CDIM + 6 + (ABS ? 2 : 0);
    // a CTA is one tile; each of its warps owns NS of the 8 sub-blocks and works on its own
    __shared__ float4 s_rec_all[kV3Slots / NS][32 * 3];
    __shared__ int4 s_im_all[kV3Slots / NS][32];  // {sorted index, sub-block mask, Gaussian row, -}
    const unsigned lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
    float4 *s_rec = s_rec_all[sub];
    int4 *s_im = s_im_all[sub];
    const uint32_t tile_lin = blockIdx.x;
    if (masks != nullptr && !masks[tile_lin]) return;
    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end = (tile_lin == n_tiles_total - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    if (range_end <= range_start) return;
    const V3Tile tc = v3_tile<NS>(tile_lin, tile_width, tile_height, lane, sub);
    const size_t cam_pix = (size_t)tc.cam * H * W;

    // per-pixel state.  tb = T_final (v_alpha_out - sum_k bg_k v_c_k) - sum_k buffer_k v_c_k folds
    // the reference's per-channel `buffer` (CS/rasterize_to_pixels_bwd.cu:203-241): v_alpha only
    // ever needs that dot product.
    float T[NS], tb[NS], v_c[NS][CDIM];
    int32_t binf[NS];
    int32_t max_bin = -1;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        T[s] = 1.f; tb[s] = 0.f; binf[s] = -1;
#pragma unroll
        for (int k = 0; k < CDIM; ++k) v_c[s][k] = 0.f;
        const uint32_t x = tc.x + 8u * (s & 1), y = tc.y + 4u * (s >> 1);
        if (x < W && y < H) {
            const size_t p = cam_pix + (size_t)y * W + x;
            const float T_final = 1.f - render_alphas[p];
            T[s] = T_final;
            binf[s] = last_ids[p];
            float bg_dot = 0.f;
#pragma unroll
            for (int k = 0; k < CDIM; ++k) {
                if (k < (int)channels) {
                    v_c[s][k] = v_render_colors[p * channels + k];
                    if (backgrounds != nullptr) bg_dot += backgrounds[(size_t)tc.cam * channels + k] * v_c[s][k];
                }
            }
            tb[s] = T_final * (v_render_alphas[p] - bg_dot);
            max_bin = max(max_bin, binf[s]);
        }
    }
    max_bin = warp_max(max_bin);
    // nothing behind the last contributor of any pixel of this tile matters
    const int32_t hi0 = min(range_end - 1, max_bin);
    if (hi0 < range_start) return;

    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    int32_t my_idx = hi0 - (int32_t)lane, my_g = 0;
    if (my_idx >= range_start) {
        my_g = flatten_ids[my_idx];
        r0 = __ldg(rec + 3 * (size_t)my_g); r1 = __ldg(rec + 3 * (size_t)my_g + 1); r2 = __ldg(rec + 3 * (size_t)my_g + 2);
    }
    uint32_t act = 0;  // warp-uniform: slots with a pixel whose last contributor is inside the batches seen so far
    for (int32_t hi = hi0; hi >= range_start; hi -= 32) {
        const int32_t lo = max(hi - 31, range_start);
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (!(act >> s & 1) && __any_sync(0xffffffffu, binf[s] >= lo)) act |= 1u << s;
        uint32_t my_mask = 0;
        if (my_idx >= range_start)
            my_mask = subblock_mask<NS>(r0.x, r0.y, r0.z, r0.w, r1.x, r2.z, tc.ox, tc.oy, W, H) & act;
        const unsigned bal = __ballot_sync(0xffffffffu, my_mask != 0);
        const int n = __popc(bal);
        __syncwarp();
        if (my_mask != 0) {
            const int pos = __popc(bal & ((1u << lane) - 1u));
            s_rec[3 * pos] = r0; s_rec[3 * pos + 1] = r1; s_rec[3 * pos + 2] = r2;
            s_im[pos] = make_int4(my_idx, (int)my_mask, my_g, 0);
        }
        __syncwarp();
        my_idx = hi - 32 - (int32_t)lane;
        if (my_idx >= range_start) {
            my_g = flatten_ids[my_idx];
            r0 = __ldg(rec + 3 * (size_t)my_g); r1 = __ldg(rec + 3 * (size_t)my_g + 1); r2 = __ldg(rec + 3 * (size_t)my_g + 2);
        }
        for (int t = 0; t < n; ++t) {
            const float4 a = s_rec[3 * t], b4 = s_rec[3 * t + 1], c4 = s_rec[3 * t + 2];
            const int4 im = s_im[t];
            const int32_t idx = im.x;
            const uint32_t m = (uint32_t)im.y;
            const float dxa = a.x - tc.px, dxb = dxa - 8.f, dyv = a.y - tc.py;
            const float hA = a.z, cb = a.w, hC = b4.x, opac = b4.y;
            const float Aa = hA * dxa * dxa, Ba = cb * dxa, Ab = hA * dxb * dxb, Bb = cb * dxb;
            const float col[4] = {b4.z, b4.w, c4.x, c4.y};
            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = 0.f;
            float W0[2] = {0.f, 0.f}, W1[2] = {0.f, 0.f}, W2[2] = {0.f, 0.f};
            bool hit = false;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (m >> s & 1) {  // warp-uniform
                    const float dy = dyv - 4.f * (float)(s >> 1);
                    const float A = (s & 1) ? Ab : Aa, B = (s & 1) ? Bb : Ba;
                    const float sigma = fmaf(dy, fmaf(hC, dy, B), A);
                    const float ov = opac * ex2_approx(-sigma);
                    const float alpha = fminf(kAlphaMax, ov);
                    const bool ok = !(sigma < 0.f) && (alpha >= kAlphaMin) && (idx <= binf[s]);
                    hit |= ok;
                    // a rejected pixel runs with alpha = 0: ra = 1, fac = 0, v_sigma = 0
                    const float a_e = ok ? alpha : 0.f;
                    const float ra = rcp_approx(1.f - a_e);
                    const float Tn = T[s] * ra;
                    T[s] = Tn;
                    const float fac = a_e * Tn;
                    float cv = 0.f;
#pragma unroll
                    for (int k = 0; k < CDIM; ++k) {
                        v[k] = fmaf(fac, v_c[s][k], v[k]);
                        cv = fmaf(col[k], v_c[s][k], cv);
                    }
                    const float v_alpha = fmaf(Tn, cv, ra * tb[s]);
                    tb[s] = fmaf(-fac, cv, tb[s]);
                    const float v_sigma = (ok && ov <= kAlphaMax) ? -ov * v_alpha : 0.f;
                    const float wy = v_sigma * dy;
                    W0[s & 1] += v_sigma;
                    W1[s & 1] += wy;
                    W2[s & 1] = fmaf(wy, dy, W2[s & 1]);
                    if (ABS) {
                        const float dx = (s & 1) ? dxb : dxa;
                        v[CDIM + 6] += fabsf(v_sigma * (2.f * hA * dx + cb * dy));
                        v[CDIM + 7] += fabsf(v_sigma * (cb * dx + 2.f * hC * dy));
                    }
                }
            }
            if (!__any_sync(0xffffffffu, hit)) continue;
            // moments -> gradients (conic entries in the record are scaled by log2 e)
            const float ua = dxa * W0[0], ub = dxb * W0[1];
            const float S1x = ua + ub, S1y = W1[0] + W1[1];
            v[CDIM + 0] = 0.5f * fmaf(dxa, ua, dxb * ub);                 // 1/2 sum v_sigma dx²
            v[CDIM + 1] = fmaf(dxa, W1[0], dxb * W1[1]);                  // sum v_sigma dx dy
            v[CDIM + 2] = 0.5f * (W2[0] + W2[1]);                         // 1/2 sum v_sigma dy²
            v[CDIM + 3] = kInvLog2e * fmaf(2.f * hA, S1x, cb * S1y);      // sum v_sigma (a dx + b dy)
            v[CDIM + 4] = kInvLog2e * fmaf(cb, S1x, 2.f * hC * S1y);      // sum v_sigma (b dx + c dy)
            v[CDIM + 5] = -(W0[0] + W0[1]) * rcp_approx(opac);            // sum vis·v_alpha
            if (ABS) { v[CDIM + 6] *= kInvLog2e; v[CDIM + 7] *= kInvLog2e; }
            const int32_t g = im.z;
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
static void launch_bwd_v3(uint32_t C, uint64_t n_isects, uint32_t channels, const float4 *rec,
                          const float *backgrounds, const uint8_t *masks, uint32_t W, uint32_t H, uint32_t tile_width,
                          uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                          const float *render_alphas, const int32_t *last_ids, const float *v_render_colors,
                          const float *v_render_alphas, float *v_means2d_abs, float *v_means2d, float *v_conics,
                          float *v_colors, float *v_opacities, cudaStream_t st) {
    const uint32_t total = C * tile_width * tile_height;
#define B2S_BWD3(NS_, MINB_)                                                                                         \
    raster_bwd_v3_kernel<CDIM, ABS, NS_, MINB_><<<total, 32 * (kV3Slots / NS_), 0, st>>>(                             \
        total, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, \
        render_alphas, last_ids, v_render_colors, v_render_alphas, v_means2d_abs, v_means2d, v_conics, v_colors,     \
        v_opacities)
    switch (tuning_variant()) {
        case 1: B2S_BWD3(8, 20); break;
        case 2: B2S_BWD3(4, 16); break;
        case 3: B2S_BWD3(4, 10); break;
        default: B2S_BWD3(8, 16); break;
    }
#undef B2S_BWD3
}

}  // namespace b2s

using namespace b2s;

extern "C" int b200splat_rasterize_bwd(uint32_t C, uint32_t n_gauss, uint64_t n_isects, uint32_t channels,
                                       const float *means2d, const float *conics, const float *colors,
                                       const float *opacities, const float *backgrounds, const uint8_t *masks,
                                       uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width,
                                       uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                                       const void *records,
                                       const float *render_alphas, const int32_t *last_ids,
                                       const float *v_render_colors, const float *v_render_alphas,
                                       float *v_means2d_abs, float *v_means2d, float *v_conics, float *v_colors,
                                       float *v_opacities, void *stream) {
    const char *where = "b200splat_rasterize_bwd";
    (void)n_gauss;
    if (records != nullptr) {
        B2S_REQUIRE(tile_size == kV3Tile && channels >= 1 && channels <= 4, where,
                    "packed records are only valid for tile_size 16 and <= 4 channels");
        B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
        if ((uint64_t)C * tile_width * tile_height == 0 || n_isects == 0) return 0;
        const float4 *rec = reinterpret_cast<const float4 *>(records);
        cudaStream_t st2 = (cudaStream_t)stream;
        const bool ab = v_means2d_abs != nullptr;
#define B2S_BWD2(D)                                                                                                    \
    case D:                                                                                                            \
        if (ab)                                                                                                        \
            launch_bwd_v3<D, true>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height,      \
                                   tile_offsets, flatten_ids, render_alphas, last_ids, v_render_colors,                \
                                   v_render_alphas, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities, st2);   \
        else                                                                                                           \
            launch_bwd_v3<D, false>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height,     \
                                    tile_offsets, flatten_ids, render_alphas, last_ids, v_render_colors,               \
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
