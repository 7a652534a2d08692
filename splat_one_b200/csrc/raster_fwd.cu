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
#include "raster_quad.cuh"

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
// quad kernels: warp-per-tile (or per half tile), 8x8 quads, packed fp32x2 (raster_quad.cuh)
// ---------------------------------------------------------------------------------------
// One Gaussian against one quad: the lane's two pixels (rows v and v + 4 of the quad).
// SLOW: the pair is flagged kNonPD (raster_quad.cuh): sigma may be negative and is tested like the
// reference does (CS/rasterize_to_pixels_fwd.cu:147), and alpha may reach the 0.999 clamp (:146).
// For every other pair neither rule can fire and both are compiled out.
template <int CDIM, bool SLOW>
__device__ __forceinline__ void fwd_quad(float2 &T2, float2 (&pix2)[CDIM], int32_t (&cur)[2], const float2 dy2,
                                         const float2 ndy2, const float nA, const float B, const float hC,
                                         const float nopac, const float (&ncol)[4], const int32_t idx) {
    const float2 u2 = __ffma2_rn(bc2(hC), dy2, bc2(B));
    const float2 ns2 = __ffma2_rn(ndy2, u2, bc2(nA));  // -sigma'
    const float2 nov2 = __fmul2_rn(bc2(nopac), make_float2(ex2_approx(ns2.x), ex2_approx(ns2.y)));
    const float nal0 = SLOW ? fmaxf(-kAlphaMax, nov2.x) : nov2.x, nal1 = SLOW ? fmaxf(-kAlphaMax, nov2.y) : nov2.y;  // -alpha
    bool ok0 = nal0 <= -kAlphaMin, ok1 = nal1 <= -kAlphaMin;
    if (SLOW) { ok0 = ok0 && !(ns2.x > 0.f); ok1 = ok1 && !(ns2.y > 0.f); }
    const float2 nae2 = make_float2(ok0 ? nal0 : 0.f, ok1 ? nal1 : 0.f);
    const float2 next_T2 = __fmul2_rn(T2, __fadd2_rn(nae2, bc2(1.f)));  // == T exactly when rejected
    // exclusive stop; also fires for a dead pixel (T < 0), which therefore composites nothing
    const bool st0 = next_T2.x <= kTransmittanceEps, st1 = next_T2.y <= kTransmittanceEps;
    const float2 nac2 = make_float2(st0 ? 0.f : nae2.x, st1 ? 0.f : nae2.y);
    const float2 nvis2 = __fmul2_rn(nac2, T2);  // -alpha T
#pragma unroll
    for (int k = 0; k < CDIM; ++k) pix2[k] = __ffma2_rn(bc2(ncol[k]), nvis2, pix2[k]);
    cur[0] = (nac2.x < 0.f) ? idx : cur[0];
    cur[1] = (nac2.y < 0.f) ? idx : cur[1];
    T2.x = st0 ? -fabsf(T2.x) : next_T2.x;  // -|T| (operand modifiers of one FSEL)
    T2.y = st1 ? -fabsf(T2.y) : next_T2.y;
}

// `quad_masks` (optional, [n_isects] bytes): the geometric quad mask of every pair this kernel stages is
// stored for the backward kernel (raster_quad.cuh pack_quad_mask).
template <int CDIM, int MINB>
__global__ void __launch_bounds__(32, MINB)
raster_fwd_quad_kernel(uint32_t n_tiles_total, uint64_t n_isects, uint32_t channels, const float4 *__restrict__ rec,
                       const float *__restrict__ backgrounds, const uint8_t *__restrict__ masks, uint32_t W,
                       uint32_t H, uint32_t tile_width, uint32_t tile_height,
                       const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                       float *__restrict__ render_colors, float *__restrict__ render_alphas,
                       int32_t *__restrict__ last_ids, uint8_t *__restrict__ quad_masks) {
    constexpr int NQ = 4;         // one warp owns the whole tile
    constexpr int NQY = NQ / 2;   // quad rows
    __shared__ float4 s_rec[32 * 3];
    __shared__ int2 s_im[32];     // {sorted index, quad mask}
    __shared__ float4 s_pre_all[32 * 3];  // records of the batch in flight
    const unsigned lane = threadIdx.x & 31;
    const float4 *s_pre = s_pre_all + 3 * lane;
    const uint32_t pre_addr = (uint32_t)__cvta_generic_to_shared(s_pre);
    const uint32_t tile_lin = blockIdx.x;
    const QuadTile tc = quad_tile<NQ>(tile_lin, tile_width, tile_height, lane, 0);
    if (backgrounds != nullptr) backgrounds += (size_t)tc.cam * channels;
    const size_t cam_pix = (size_t)tc.cam * H * W;

    // pixel j of quad q: (tc.x + 8 (q & 1), tc.y + 8 (q >> 1) + 4 j)
    uint32_t in_mask = 0;  // bit 2q + j: that pixel is inside the image
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (tc.x + 8u * (q & 1) < W && tc.y + 8u * (q >> 1) + 4u * j < H) in_mask |= 1u << (2 * q + j);

    if (masks != nullptr && !masks[tile_lin]) {
#pragma unroll
        for (int s = 0; s < 2 * NQ; ++s)
            if (in_mask >> s & 1) {
                const size_t p = cam_pix + (size_t)(tc.y + 8u * (s >> 2) + 4u * (s & 1)) * W + tc.x + 8u * (s >> 1 & 1);
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
    float2 T2[NQ], pix2[NQ][CDIM];
    int32_t cur[NQ][2];
    uint32_t live = 0;  // warp-uniform: quads with a live pixel
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        T2[q] = make_float2((in_mask >> (2 * q) & 1) ? 1.f : -1.f, (in_mask >> (2 * q + 1) & 1) ? 1.f : -1.f);
        cur[q][0] = cur[q][1] = 0;
#pragma unroll
        for (int k = 0; k < CDIM; ++k) pix2[q][k] = make_float2(0.f, 0.f);
        if (__any_sync(0xffffffffu, (in_mask >> (2 * q) & 3) != 0)) live |= 1u << q;
    }
    // per-lane constants: centres of the lane's pixel rows (and their negatives), per quad row
    float2 pyc2[NQY], npyc2[NQY];
#pragma unroll
    for (int qy = 0; qy < NQY; ++qy) {
        pyc2[qy] = make_float2(tc.py + 8.f * qy, tc.py + 8.f * qy + 4.f);
        npyc2[qy] = make_float2(-pyc2[qy].x, -pyc2[qy].y);
    }
    const float pxa = tc.px, pxb = tc.px + 8.f;

    // prefetch of this lane's record: one batch ahead straight into shared memory (cp.async), two
    // batches ahead for the sorted id it is addressed by
    int32_t my_idx = range_start + (int32_t)lane, g_next = 0;
    if (my_idx < range_end) prefetch_record(pre_addr, rec, flatten_ids[my_idx]);
    if (my_idx + 32 < range_end) g_next = flatten_ids[my_idx + 32];
    for (int32_t base = range_start; base < range_end && live != 0; base += 32) {
        uint32_t my_mask = 0;
        float4 r0, r1, r2;
        cp_async_wait_all();
        if (my_idx < range_end) {
            r0 = s_pre[0]; r1 = s_pre[1]; r2 = s_pre[2];
            my_mask = quad_mask<NQ>(r0.x, r0.y, r0.z, r0.w, r1.x, r2.z, r1.y, tc.ox, tc.oy, W, H);
            if (quad_masks != nullptr) quad_masks[my_idx] = pack_quad_mask(my_mask);
            if ((my_mask & live) == 0) my_mask = 0; else my_mask &= (live | kNonPD);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, my_mask != 0);
        const int n = __popc(bal);
        __syncwarp();  // readers of the previous batch are done
        if (my_mask != 0) {
            const int pos = __popc(bal & ((1u << lane) - 1u));
            s_rec[3 * pos] = r0; s_rec[3 * pos + 1] = r1;
            s_rec[3 * pos + 2] = make_float4(r2.x, r2.y, -r0.y, 0.f);
            s_im[pos] = make_int2(my_idx, (int)my_mask);
        }
        __syncwarp();
        // prefetch the next batch while this one is composited
        my_idx = base + 32 + (int32_t)lane;
        if (my_idx < range_end) prefetch_record(pre_addr, rec, g_next);
        if (my_idx + 32 < range_end) g_next = flatten_ids[my_idx + 32];
        for (int t = 0; t < n; ++t) {
            const float4 a = s_rec[3 * t], b4 = s_rec[3 * t + 1], c4 = s_rec[3 * t + 2];
            const int2 im = s_im[t];
            const int32_t idx = im.x;
            const uint32_t m = (uint32_t)im.y;
            const float dxa = a.x - pxa, dxb = a.x - pxb;
            const float hC = b4.x, nopac = b4.y;
            const float nAa = -(a.z * dxa) * dxa, Ba = a.w * dxa, nAb = -(a.z * dxb) * dxb, Bb = a.w * dxb;
            const float ncol[4] = {b4.z, b4.w, c4.x, c4.y};
            float2 dy2[NQY], ndy2[NQY];
#pragma unroll
            for (int qy = 0; qy < NQY; ++qy) {
                ndy2[qy] = __fadd2_rn(pyc2[qy], bc2(c4.z));   // p_y - g_y
                dy2[qy] = __fadd2_rn(npyc2[qy], bc2(a.y));    // g_y - p_y
            }
#define B2S_FQ(SLOW_, q_)                                                                                            \
    fwd_quad<CDIM, SLOW_>(T2[q_], pix2[q_], cur[q_], dy2[(q_) >> 1], ndy2[(q_) >> 1], ((q_) & 1) ? nAb : nAa,          \
                          ((q_) & 1) ? Bb : Ba, hC, nopac, ncol, idx)
            if (m & kNonPD) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (m >> q & 1) B2S_FQ(true, q);  // warp-uniform
            } else {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (m >> q & 1) B2S_FQ(false, q);  // warp-uniform
            }
#undef B2S_FQ
        }
        // quads whose pixels have all stopped are skipped from now on
        uint32_t nl = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if ((live >> q & 1) && __any_sync(0xffffffffu, T2[q].x > 0.f || T2[q].y > 0.f)) nl |= 1u << q;
        live = nl;
    }

#pragma unroll
    for (int q = 0; q < NQ; ++q) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (in_mask >> (2 * q + j) & 1) {
                const size_t p = cam_pix + (size_t)(tc.y + 8u * (q >> 1) + 4u * j) * W + tc.x + 8u * (q & 1);
                const float Tf = fabsf(j ? T2[q].y : T2[q].x);
                render_alphas[p] = 1.f - Tf;
#pragma unroll
                for (int k = 0; k < CDIM; ++k)
                    if (k < (int)channels) {
                        const float c = j ? pix2[q][k].y : pix2[q][k].x;
                        render_colors[p * channels + k] = backgrounds == nullptr ? c : (c + Tf * backgrounds[k]);
                    }
                last_ids[p] = cur[q][j];
            }
        }
    }
}

template <int CDIM>
static void launch_fwd_quad(uint32_t C, uint64_t n_isects, uint32_t channels, const float4 *rec,
                            const float *backgrounds, const uint8_t *masks, uint32_t W, uint32_t H,
                            uint32_t tile_width, uint32_t tile_height, const int32_t *tile_offsets,
                            const int32_t *flatten_ids, float *render_colors, float *render_alphas, int32_t *last_ids,
                            uint8_t *quad_masks, cudaStream_t st) {
    const uint32_t total = C * tile_width * tile_height;
#define B2S_FWDQ(MINB_)                                                                                              \
    raster_fwd_quad_kernel<CDIM, MINB_><<<total, 32, 0, st>>>(total, n_isects, channels, rec, backgrounds, masks, W,  \
                                                               H, tile_width, tile_height, tile_offsets, flatten_ids, \
                                                               render_colors, render_alphas, last_ids, quad_masks)
#ifdef B2S_TUNING
    switch (tuning_variant()) {
        case 1: B2S_FWDQ(16); return;
        case 2: B2S_FWDQ(24); return;
        case 3: B2S_FWDQ(18); return;
        default: break;
    }
#endif
    B2S_FWDQ(20);
#undef B2S_FWDQ
}

}  // namespace b2s

using namespace b2s;

extern "C" size_t b200splat_rasterize_records_bytes(uint32_t n_gauss, uint32_t channels, uint32_t tile_size) {
    if (tile_size != kQTile || channels < 1 || channels > 4) return 0;  // generic path: no records
    return (size_t)n_gauss * 3 * sizeof(float4);
}

extern "C" int b200splat_rasterize_pack(uint32_t n_gauss, uint32_t channels, const float *means2d,
                                        const float *conics, const float *colors, const float *opacities,
                                        void *records, void *stream) {
    const char *where = "b200splat_rasterize_pack";
    B2S_REQUIRE_ALIGNED8(means2d, where);
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
                                       const void *records, uint8_t *quad_masks,
                                       float *render_colors, float *render_alphas, int32_t *last_ids, void *stream) {
    const char *where = "b200splat_rasterize_fwd";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    (void)n_gauss;
    if (records != nullptr) {
        B2S_REQUIRE(tile_size == kQTile && channels >= 1 && channels <= 4, where,
                    "packed records are only valid for tile_size 16 and <= 4 channels");
        B2S_REQUIRE((uint64_t)tile_width * tile_size >= W && (uint64_t)tile_height * tile_size >= H, where,
                    "tile grid does not cover the image");
        B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
        if ((uint64_t)C * tile_width * tile_height == 0) return 0;
        const float4 *rec = reinterpret_cast<const float4 *>(records);
        cudaStream_t st2 = (cudaStream_t)stream;
        switch (channels) {
            case 1: launch_fwd_quad<1>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, quad_masks, st2); break;
            case 2: launch_fwd_quad<2>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, quad_masks, st2); break;
            case 3: launch_fwd_quad<3>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, quad_masks, st2); break;
            default: launch_fwd_quad<4>(C, n_isects, channels, rec, backgrounds, masks, W, H, tile_width, tile_height, tile_offsets, flatten_ids, render_colors, render_alphas, last_ids, quad_masks, st2); break;
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
