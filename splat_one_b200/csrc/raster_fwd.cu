// raster_fwd.cu — per-tile front-to-back alpha compositing (a8).
// Replaces CS/rasterize_to_pixels_fwd.cu:16-186.  Semantics kept exactly: pixel centre
// (j+0.5, i+0.5); sigma = 0.5(a dx² + c dy²) + b dx dy; alpha = min(0.999, o·exp(-sigma));
// skip when sigma < 0 or alpha < 1/255; stop (exclusive) when T(1-alpha) <= 1e-4;
// alpha_out = 1-T; colour += T·background; last_ids = sorted index of the last
// contributor; masked tiles write background only.
//
// Design (B200): one CTA per tile, one thread per pixel.  Each batch of Gaussians is
// gathered once per CTA into shared memory as two 16-byte records (+ one for the colour
// when D <= 4), so the per-pair inner loop is two/three broadcast LDS.128 instead of the
// reference's 7 scalar LDS + a dependent global colour read, and `flatten_ids` is
// consumed with coalesced loads.  The loop body is the fp32/MUFU critical path; see
// DESIGN.md for the measured limits.
#include "raster_common.cuh"
#include "raster_v3.cuh"

namespace b2s {

template <int CDIM, int MAXT>
__global__ void __launch_bounds__(MAXT)
raster_fwd_kernel(uint32_t C, uint64_t n_isects, uint32_t channels, const float2 *__restrict__ means2d,
                  const float *__restrict__ conics, const float *__restrict__ colors,
                  const float *__restrict__ opacities, const float *__restrict__ backgrounds,
                  const uint8_t *__restrict__ masks, uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width,
                  uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                  const int32_t *__restrict__ flatten_ids, float *__restrict__ render_colors,
                  float *__restrict__ render_alphas, int32_t *__restrict__ last_ids) {
    constexpr bool kStageColor = CDIM <= 4;
    const TileCoord tc = tile_coord(tile_size, tile_width, tile_height, W, H);
    const uint32_t n_tiles_total = C * tile_width * tile_height;
    const bool inside = tc.inside;
    const size_t pix = ((size_t)tc.cam * H + tc.i) * W + tc.j;
    const float px = (float)tc.j + 0.5f, py = (float)tc.i + 0.5f;
    if (backgrounds != nullptr) backgrounds += (size_t)tc.cam * channels;

    if (masks != nullptr && !masks[tc.tile_lin]) {
        // masked tile: background only; alpha / last_ids untouched (CS/...fwd.cu:71-77)
        if (inside) {
            for (uint32_t k = 0; k < channels; ++k)
                render_colors[pix * channels + k] = backgrounds == nullptr ? 0.f : backgrounds[k];
        }
        return;
    }

    const int32_t range_start = tile_offsets[tc.tile_lin];
    const int32_t range_end =
        (tc.tile_lin == n_tiles_total - 1) ? (int32_t)n_isects : tile_offsets[tc.tile_lin + 1];
    const uint32_t block_size = blockDim.x;
    const uint32_t num_batches = (uint32_t)(range_end - range_start + block_size - 1) / block_size;

    extern __shared__ float4 smem4[];
    float4 *rec_a = smem4;                  // {x, y, opacity, conic.a}
    float4 *rec_b = smem4 + block_size;     // {conic.b, conic.c, id (bits), -}
    float4 *rec_c = smem4 + 2 * block_size; // colour (only when kStageColor)

    float T = 1.f;
    uint32_t cur_idx = 0;
    bool done = !inside;
    const uint32_t tr = threadIdx.x;
    float pix_out[CDIM];
#pragma unroll
    for (int k = 0; k < CDIM; ++k) pix_out[k] = 0.f;

    for (uint32_t b = 0; b < num_batches; ++b) {
        // also protects the smem records of the previous batch
        if (__syncthreads_count(done) >= (int)block_size) break;
        const uint32_t batch_start = range_start + block_size * b;
        const uint32_t idx = batch_start + tr;
        if (idx < (uint32_t)range_end) {
            const int32_t g = flatten_ids[idx];
            const float2 xy = __ldg(means2d + g);
            const float opac = __ldg(opacities + g);
            const float ca = __ldg(conics + 3 * (size_t)g), cb = __ldg(conics + 3 * (size_t)g + 1),
                        cc = __ldg(conics + 3 * (size_t)g + 2);
            rec_a[tr] = make_float4(xy.x, xy.y, opac, ca);
            rec_b[tr] = make_float4(cb, cc, __int_as_float(g), 0.f);
            if (kStageColor) {
                float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < CDIM; ++k)
                    if (k < (int)channels) c4[k] = __ldg(colors + (size_t)g * channels + k);
                rec_c[tr] = make_float4(c4[0], c4[1], c4[2], c4[3]);
            }
        }
        __syncthreads();
        const uint32_t batch_size = min(block_size, (uint32_t)range_end - batch_start);
        for (uint32_t t = 0; (t < batch_size) && !done; ++t) {
            const float4 ra = rec_a[t];
            const float4 rb = rec_b[t];
            const float dx = ra.x - px, dy = ra.y - py;
            const float sigma = 0.5f * (ra.w * dx * dx + rb.y * dy * dy) + rb.x * dx * dy;
            const float alpha = fminf(kAlphaMax, ra.z * __expf(-sigma));
            if (sigma < 0.f || alpha < kAlphaMin) continue;
            const float next_T = T * (1.f - alpha);
            if (next_T <= kTransmittanceEps) {  // exclusive: this Gaussian is not composited
                done = true;
                break;
            }
            const float vis = alpha * T;
            if (kStageColor) {
                const float4 rc = rec_c[t];
                const float cv[4] = {rc.x, rc.y, rc.z, rc.w};
#pragma unroll
                for (int k = 0; k < CDIM; ++k) pix_out[k] += cv[k] * vis;
            } else {
                const float *c_ptr = colors + (size_t)__float_as_int(rb.z) * channels;
#pragma unroll
                for (int k = 0; k < CDIM; ++k)
                    if (k < (int)channels) pix_out[k] += __ldg(c_ptr + k) * vis;
            }
            cur_idx = batch_start + t;
            T = next_T;
        }
    }

    if (inside) {
        render_alphas[pix] = 1.f - T;
#pragma unroll
        for (int k = 0; k < CDIM; ++k)
            if (k < (int)channels)
                render_colors[pix * channels + k] =
                    backgrounds == nullptr ? pix_out[k] : (pix_out[k] + T * backgrounds[k]);
        last_ids[pix] = (int32_t)cur_idx;
    }
}

template <int CDIM>
static int launch_fwd(uint32_t C, uint64_t n_isects, uint32_t channels, const float *means2d, const float *conics,
                      const float *colors, const float *opacities, const float *backgrounds, const uint8_t *masks,
                      uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                      const int32_t *tile_offsets, const int32_t *flatten_ids, float *render_colors,
                      float *render_alphas, int32_t *last_ids, cudaStream_t st) {
    const uint32_t threads = ((tile_size * tile_size + 31) / 32) * 32;
    const uint32_t grid = C * tile_width * tile_height;
    const size_t smem = (size_t)threads * sizeof(float4) * (CDIM <= 4 ? 3 : 2);
    if (threads <= 256) {
        raster_fwd_kernel<CDIM, 256><<<grid, threads, smem, st>>>(
            C, n_isects, channels, reinterpret_cast<const float2 *>(means2d), conics, colors, opacities, backgrounds,
            masks, W, H, tile_size, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas,
            last_ids);
    } else {
        raster_fwd_kernel<CDIM, 1024><<<grid, threads, smem, st>>>(
            C, n_isects, channels, reinterpret_cast<const float2 *>(means2d), conics, colors, opacities, backgrounds,
            masks, W, H, tile_size, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas,
            last_ids);
    }
    return 0;
}


// ---------------------------------------------------------------------------------------
// v3: warp-per-tile, 8 sub-block slots per lane, exact sub-block culling (raster_v3.cuh)
// ---------------------------------------------------------------------------------------
template <int CDIM, int NS, int MINB>
__global__ void __launch_bounds__(32 * (kV3Slots / NS), MINB)
raster_fwd_v3_kernel(uint32_t n_tiles_total, uint64_t n_isects, uint32_t channels, const float4 *__restrict__ rec,
                     const float *__restrict__ backgrounds, const uint8_t *__restrict__ masks, uint32_t W, uint32_t H,
                     uint32_t tile_width, uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                     const int32_t *__restrict__ flatten_ids, float *__restrict__ render_colors,
                     float *__restrict__ render_alphas, int32_t *__restrict__ last_ids) {
    // a CTA is one tile; each of its warps owns NS of the 8 sub-blocks and works on its own
    // (no block-level synchronisation anywhere)
    __shared__ float4 s_rec_all[kV3Slots / NS][32 * 3];
    __shared__ int2 s_im_all[kV3Slots / NS][32];  // {sorted index, sub-block mask}
    const unsigned lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
    float4 *s_rec = s_rec_all[sub];
    int2 *s_im = s_im_all[sub];
    const uint32_t tile_lin = blockIdx.x;
    const V3Tile tc = v3_tile<NS>(tile_lin, tile_width, tile_height, lane, sub);
    if (backgrounds != nullptr) backgrounds += (size_t)tc.cam * channels;
    const size_t cam_pix = (size_t)tc.cam * H * W;

    // pixel of slot s: (tc.x + 8 (s & 1), tc.y + 4 (s >> 1))
    uint32_t in_mask = 0;  // bit s: this lane's pixel of slot s is inside the image
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (tc.x + 8u * (s & 1) < W && tc.y + 4u * (s >> 1) < H) in_mask |= 1u << s;

    if (masks != nullptr && !masks[tile_lin]) {
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (in_mask >> s & 1) {
                const size_t p = cam_pix + (size_t)(tc.y + 4u * (s >> 1)) * W + tc.x + 8u * (s & 1);
                for (uint32_t k = 0; k < channels; ++k)
                    render_colors[p * channels + k] = backgrounds == nullptr ? 0.f : backgrounds[k];
            }
        return;
    }

    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end = (tile_lin == n_tiles_total - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];

    // the sign of T is the liveness flag: T > 0 while the pixel accumulates, -T_final once it has
    // stopped (or lies outside the image).  A dead pixel then "stops" again on every Gaussian:
    // next_T = T (1 - alpha) < 0 <= 1e-4, so it composites nothing and keeps its T.
    float T[NS], pix[NS][CDIM];
    int32_t cur[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        T[s] = (in_mask >> s & 1) ? 1.f : -1.f;
        cur[s] = 0;
#pragma unroll
        for (int k = 0; k < CDIM; ++k) pix[s][k] = 0.f;
    }
    uint32_t live = __reduce_or_sync(0xffffffffu, in_mask);  // warp-uniform: slots with a live pixel

    // software prefetch of this lane's record for the first batch
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    int32_t my_idx = range_start + (int32_t)lane;
    if (my_idx < range_end) {
        const int32_t g = flatten_ids[my_idx];
        r0 = __ldg(rec + 3 * (size_t)g); r1 = __ldg(rec + 3 * (size_t)g + 1); r2 = __ldg(rec + 3 * (size_t)g + 2);
    }
    for (int32_t base = range_start; base < range_end && live != 0; base += 32) {
        uint32_t my_mask = 0;
        if (my_idx < range_end)
            my_mask = subblock_mask<NS>(r0.x, r0.y, r0.z, r0.w, r1.x, r2.z, tc.ox, tc.oy, W, H) & live;
        const unsigned bal = __ballot_sync(0xffffffffu, my_mask != 0);
        const int n = __popc(bal);
        __syncwarp();  // readers of the previous batch are done
        if (my_mask != 0) {
            const int pos = __popc(bal & ((1u << lane) - 1u));
            s_rec[3 * pos] = r0; s_rec[3 * pos + 1] = r1; s_rec[3 * pos + 2] = r2;
            s_im[pos] = make_int2(my_idx, (int)my_mask);
        }
        __syncwarp();
        // prefetch the next batch while this one is composited
        my_idx = base + 32 + (int32_t)lane;
        if (my_idx < range_end) {
            const int32_t g = flatten_ids[my_idx];
            r0 = __ldg(rec + 3 * (size_t)g); r1 = __ldg(rec + 3 * (size_t)g + 1); r2 = __ldg(rec + 3 * (size_t)g + 2);
        }
        for (int t = 0; t < n; ++t) {
            const float4 a = s_rec[3 * t], b4 = s_rec[3 * t + 1], c4 = s_rec[3 * t + 2];
            const int2 im = s_im[t];
            const int32_t idx = im.x;
            const uint32_t m = (uint32_t)im.y;
            const float dxa = a.x - tc.px, dxb = dxa - 8.f, dyv = a.y - tc.py;
            const float hC = b4.x, opac = b4.y;
            const float Aa = a.z * dxa * dxa, Ba = a.w * dxa, Ab = a.z * dxb * dxb, Bb = a.w * dxb;
            const float col[4] = {b4.z, b4.w, c4.x, c4.y};
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (m >> s & 1) {  // warp-uniform
                    const float dy = dyv - 4.f * (float)(s >> 1);
                    const float A = (s & 1) ? Ab : Aa, B = (s & 1) ? Bb : Ba;
                    const float nsigma = fmaf(-dy, fmaf(hC, dy, B), -A);  // -sigma'
                    const float alpha = fminf(kAlphaMax, opac * ex2_approx(nsigma));
                    const bool ok = !(nsigma > 0.f) && (alpha >= kAlphaMin);
                    const float a_e = ok ? alpha : 0.f;
                    const float next_T = T[s] * (1.f - a_e);         // == T exactly when rejected
                    const bool stop = next_T <= kTransmittanceEps;   // exclusive stop (implies ok, or dead)
                    const float vis = (stop ? 0.f : a_e) * T[s];
#pragma unroll
                    for (int k = 0; k < CDIM; ++k) pix[s][k] = fmaf(col[k], vis, pix[s][k]);
                    cur[s] = (ok && !stop) ? idx : cur[s];
                    T[s] = stop ? -fabsf(T[s]) : next_T;
                }
            }
        }
        // slots whose pixels have all stopped are skipped from now on
        uint32_t nl = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if ((live >> s & 1) && __any_sync(0xffffffffu, T[s] > 0.f)) nl |= 1u << s;
        live = nl;
    }

#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (in_mask >> s & 1) {
            const size_t p = cam_pix + (size_t)(tc.y + 4u * (s >> 1)) * W + tc.x + 8u * (s & 1);
            const float Tf = fabsf(T[s]);
            render_alphas[p] = 1.f - Tf;
#pragma unroll
            for (int k = 0; k < CDIM; ++k)
                if (k < (int)channels)
                    render_colors[p * channels + k] = backgrounds == nullptr ? pix[s][k] : (pix[s][k] + Tf * backgrounds[k]);
            last_ids[p] = cur[s];
        }
    }
}

template <int CDIM>
static void launch_fwd_v3(uint32_t C, uint64_t n_isects, uint32_t channels, const float4 *rec,
                          const float *backgrounds, const uint8_t *masks, uint32_t W, uint32_t H, uint32_t tile_width,
                          uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                          float *render_colors, float *render_alphas, int32_t *last_ids, cudaStream_t st) {
    const uint32_t total = C * tile_width * tile_height;
#define B2S_FWD3(NS_, MINB_)                                                                                        \
    raster_fwd_v3_kernel<CDIM, NS_, MINB_><<<total, 32 * (kV3Slots / NS_), 0, st>>>(                                 \
        total, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, \
        render_colors, render_alphas, last_ids)
    switch (tuning_variant()) {
        case 1: B2S_FWD3(8, 16); break;
        case 2: B2S_FWD3(4, 16); break;
        case 3: B2S_FWD3(4, 10); break;
        default: B2S_FWD3(8, 20); break;
    }
#undef B2S_FWD3
}

}  // namespace b2s

using namespace b2s;

extern "C" size_t b200splat_rasterize_records_bytes(uint32_t n_gauss, uint32_t channels, uint32_t tile_size) {
    if (tile_size != kV3Tile || channels < 1 || channels > 4) return 0;  // generic path: no records
    return (size_t)n_gauss * 3 * sizeof(float4);
}

extern "C" int b200splat_rasterize_pack(uint32_t n_gauss, uint32_t channels, const float *means2d,
                                        const float *conics, const float *colors, const float *opacities,
                                        void *records, void *stream) {
    const char *where = "b200splat_rasterize_pack";
    B2S_REQUIRE(channels >= 1 && channels <= 4, where, "records hold at most 4 channels");
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(records) & 15) == 0, where, "records must be 16-byte aligned");
    if (n_gauss == 0) return 0;
    pack_records_kernel<<<div_up(n_gauss, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        n_gauss, channels, reinterpret_cast<const float2 *>(means2d), conics, colors, opacities,
        reinterpret_cast<float4 *>(records));
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_rasterize_fwd(uint32_t C, uint32_t n_gauss, uint64_t n_isects, uint32_t channels,
                                       const float *means2d, const float *conics, const float *colors,
                                       const float *opacities, const float *backgrounds, const uint8_t *masks,
                                       uint32_t W, uint32_t H, uint32_t tile_size, uint32_t tile_width,
                                       uint32_t tile_height, const int32_t *tile_offsets, const int32_t *flatten_ids,
                                       const void *records,
                                       float *render_colors, float *render_alphas, int32_t *last_ids, void *stream) {
    const char *where = "b200splat_rasterize_fwd";
    (void)n_gauss;
    if (records != nullptr) {
        B2S_REQUIRE(tile_size == kV3Tile && channels >= 1 && channels <= 4, where,
                    "packed records are only valid for tile_size 16 and <= 4 channels");
        B2S_REQUIRE((uint64_t)tile_width * tile_size >= W && (uint64_t)tile_height * tile_size >= H, where,
                    "tile grid does not cover the image");
        B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
        if ((uint64_t)C * tile_width * tile_height == 0) return 0;
        const float4 *rec = reinterpret_cast<const float4 *>(records);
        cudaStream_t st2 = (cudaStream_t)stream;
        switch (channels) {
            case 1: launch_fwd_v3<1>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, st2); break;
            case 2: launch_fwd_v3<2>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, st2); break;
            case 3: launch_fwd_v3<3>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, st2); break;
            default: launch_fwd_v3<4>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, st2); break;
        }
        B2S_CHECK_LAUNCH(where);
        return 0;
    }
    B2S_REQUIRE(tile_size >= 1 && tile_size <= 32, where, "tile_size must be in [1, 32]");
    B2S_REQUIRE((uint64_t)tile_width * tile_size >= W && (uint64_t)tile_height * tile_size >= H, where,
                "tile grid does not cover the image");
    B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
    const int cdim = pick_cdim(channels);
    B2S_REQUIRE(channels >= 1 && cdim > 0, where, "unsupported number of color channels (1..33)");
    if ((uint64_t)C * tile_width * tile_height == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define B2S_FWD(D)                                                                                                   \
    case D:                                                                                                          \
        launch_fwd<D>(C, n_isects, channels, means2d, conics, colors, opacities, backgrounds, masks, W, H, tile_size, \
                      tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids,    \
                      st);                                                                                           \
        break;
    switch (cdim) {
        B2S_FWD(1) B2S_FWD(2) B2S_FWD(3) B2S_FWD(4) B2S_FWD(5) B2S_FWD(8) B2S_FWD(9) B2S_FWD(16) B2S_FWD(17)
        B2S_FWD(32) B2S_FWD(33)
    }
#undef B2S_FWD
    B2S_CHECK_LAUNCH(where);
    return 0;
}
