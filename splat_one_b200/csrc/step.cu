// step.cu — the training step either side of rasterization().
//
//  * splat activations: `scales = exp(splats["scales"])`, `opacities = sigmoid(splats["opacities"])`
//    (R/utils/gsplat_utils/gsplat_trainer.py:458-459) — one launch instead of two, and one
//    launch for both backward products.
//  * photometric loss: `l1_loss(colors, pixels) * (1 - lambda) + (1 - fused_ssim(colors, pixels,
//    padding="valid")) * lambda` (gsplat_trainer.py:624-628).  fused_ssim is a third-party
//    dependency of the reference that is not vendored; what it computes is the standard SSIM
//    (Wang et al. 2004): 11x11 Gaussian window, sigma 1.5, C1 = 0.01², C2 = 0.03², statistics
//    by zero-padded correlation, mean over the pixels whose window lies inside the image
//    ("valid").  Here both terms and all three derivative maps come out of ONE pass over the
//    two images in their native [C,H,W,3] layout (the reference permutes to NCHW and copies),
//    and the backward is one separable convolution of the derivative maps.
//
// Both kernels are HBM-streaming: 24 B/pixel read + 36 B/pixel of derivative maps written
// (forward), 60 B/pixel read + 12 B/pixel written (backward); the 11-tap separable window
// runs out of shared memory (32x32 pixel tiles with a 5-pixel apron).
#include "common.cuh"

namespace b2s {

// ---------------------------------------------------------------------------------------
// activations
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
splat_activations_fwd_kernel(uint32_t N, const float *__restrict__ scales_raw, const float *__restrict__ opac_raw,
                             float *__restrict__ scales, float *__restrict__ opacities) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * N) scales[i] = expf(scales_raw[i]);
    if (i < N) opacities[i] = 1.f / (1.f + expf(-opac_raw[i]));
}

__global__ void __launch_bounds__(kThreads)
splat_activations_bwd_kernel(uint32_t N, const float *__restrict__ scales, const float *__restrict__ opacities,
                             const float *__restrict__ v_scales, const float *__restrict__ v_opacities,
                             float *__restrict__ v_scales_raw, float *__restrict__ v_opac_raw) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * N && v_scales_raw != nullptr) v_scales_raw[i] = v_scales[i] * scales[i];
    if (i < N && v_opac_raw != nullptr) {
        const float o = opacities[i];
        v_opac_raw[i] = v_opacities[i] * o * (1.f - o);
    }
}

// ---------------------------------------------------------------------------------------
// viewmats = inverse(camtoworlds)  (gsplat_trainer.py:483 `torch.linalg.inv`): general 4x4
// inverse by the adjugate, evaluated in double.  torch.linalg.inv synchronises the host to
// check for singular inputs, which drains the launch queue once per step; this kernel does
// not (a singular matrix yields inf/nan, like 1/0).
// ---------------------------------------------------------------------------------------
__global__ void invert_4x4_kernel(uint32_t C, const float *__restrict__ mats, float *__restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m[16], inv[16];
    for (int i = 0; i < 16; i++) m[i] = (double)mats[16 * c + i];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    const double idet = 1.0 / det;
    for (int i = 0; i < 16; i++) out[16 * c + i] = (float)(inv[i] * idet);
}

// ---------------------------------------------------------------------------------------
// L1 + SSIM
// ---------------------------------------------------------------------------------------
constexpr int kWin = 11, kHalf = 5;
constexpr int kTile = 32;                    // output pixels per block edge
constexpr int kHalo = kTile + 2 * kHalf;     // 42
constexpr int kRowF = kHalo * 3;             // floats per staged row (3 channels interleaved)
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

struct Window { float g[kWin]; };

static Window make_window() {
    Window w;
    double v[kWin], sum = 0.0;
    for (int i = 0; i < kWin; i++) { v[i] = exp(-(double)((i - kHalf) * (i - kHalf)) / (2.0 * 1.5 * 1.5)); sum += v[i]; }
    for (int i = 0; i < kWin; i++) w.g[i] = (float)(v[i] / sum);
    return w;
}

// Shared-memory geometry.  A staged apron row holds kHalo pixels x 3 channels; rows are padded
// to an odd number of elements so that lanes walking DOWN the rows (the horizontal pass) hit
// distinct banks, and the horizontal pass writes its results transposed ([column][row], odd
// row count) so that the vertical pass, whose lanes walk ACROSS the columns, does too.
constexpr int kRowS = kRowF + 1;             // 127 elements per staged row
constexpr int kColS = kHalo + 1;             // 43 rows per transposed column
constexpr int kRun = 6;                      // consecutive outputs per thread, horizontal pass
constexpr int kRunsPerRow = 6;               // 6 runs of 6 cover the 32 columns (starts 0,6,11,16,21,26: the
                                             // overlapping columns are computed twice with identical results)
constexpr int kRowsPerThread = 4;            // consecutive outputs per thread, vertical pass

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ int run_start(int i) { return i < 2 ? 6 * i : 5 * i + 1; }

// Stage the apron tile of up to three [H,W,3] images (zero outside the image): elements of
// images A and B interleaved as float2, image C (optional) as float.
template <bool WITH_C>
__device__ __forceinline__ void stage_tiles(const float *__restrict__ A, const float *__restrict__ B,
                                            const float *__restrict__ Cc, uint32_t H, uint32_t W, int x0, int y0,
                                            float2 *s_ab, float *s_c, int halo_rows = kHalo) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int off[4];   // element offset inside an image row, or -1 when the column is outside the image
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int f = lane + 32 * i, px3 = f / 3, x = x0 - kHalf + px3;
        off[i] = (f < kRowF && x >= 0 && x < (int)W) ? x * 3 + (f - 3 * px3) : -1;
    }
    for (int r = warp; r < halo_rows; r += 8) {
        const int y = y0 - kHalf + r;
        const bool row_ok = y >= 0 && y < (int)H;
        const size_t base = (size_t)(row_ok ? y : 0) * W * 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = lane + 32 * i;
            if (f >= kRowF) continue;
            float a = 0.f, b = 0.f, c = 0.f;
            if (row_ok && off[i] >= 0) {
                a = __ldg(A + base + off[i]);
                b = __ldg(B + base + off[i]);
                if (WITH_C) c = __ldg(Cc + base + off[i]);
            }
            s_ab[r * kRowS + f] = f2(a, b);
            if (WITH_C) s_c[r * kRowS + f] = c;
        }
    }
}

template <int TY>
__global__ void __launch_bounds__(256, TY == 16 ? 4 : 3)
l1_ssim_fwd_kernel(uint32_t H, uint32_t W, const float *__restrict__ img, const float *__restrict__ tgt,
                   const Window win, float *__restrict__ d_mu, float *__restrict__ d_xx, float *__restrict__ d_xy,
                   float *__restrict__ partials) {
    constexpr int HY = TY + 2 * kHalf;        // apron rows
    constexpr int CS = HY + 1;                // transposed column stride (odd)
    constexpr int RPT = TY / 8;               // consecutive output rows per thread, vertical pass
    extern __shared__ float2 smem2[];
    float2 *s_ab = smem2;                        // [HY][kRowS]   {img, tgt}
    float2 *s_m = s_ab + HY * kRowS;          // [kTile][CS]   {sum g a, sum g b}
    float2 *s_q = s_m + kTile * CS;           // [kTile][CS]   {sum g a², sum g b²}
    float *s_p = reinterpret_cast<float *>(s_q + kTile * CS);  // [kTile][CS]  sum g a b
    const uint32_t cam = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * TY;
    const size_t img_off = (size_t)cam * H * W * 3;
    stage_tiles<false>(img + img_off, tgt + img_off, nullptr, H, W, x0, y0, s_ab, nullptr, HY);
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int px = x0 + tx;
    float l1 = 0.f, ss = 0.f;
    float o_mu[RPT][3], o_xx[RPT][3], o_xy[RPT][3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        // horizontal pass: thread = (apron row r, run of kRun output columns); packed fp32x2
        // accumulation of the {a, b} and {a², b²} pairs (FFMA2), scalar a·b
        if (threadIdx.x < HY * kRunsPerRow) {
            const int r = threadIdx.x % HY, xr = run_start(threadIdx.x / HY);
            float2 hm[kRun], hq[kRun];
            float hp[kRun];
#pragma unroll
            for (int j = 0; j < kRun; ++j) { hm[j] = f2(0.f, 0.f); hq[j] = f2(0.f, 0.f); hp[j] = 0.f; }
#pragma unroll
            for (int e = 0; e < kRun + kWin - 1; ++e) {
                const float2 ab = s_ab[r * kRowS + (xr + e) * 3 + ch];
                const float2 sq = __fmul2_rn(ab, ab);
                const float pr = ab.x * ab.y;
#pragma unroll
                for (int j = 0; j < kRun; ++j) {
                    const int k = e - j;   // tap index of element e for output j
                    if (k >= 0 && k < kWin) {
                        const float g = win.g[k];
                        hm[j] = __ffma2_rn(f2(g, g), ab, hm[j]);
                        hq[j] = __ffma2_rn(f2(g, g), sq, hq[j]);
                        hp[j] = fmaf(g, pr, hp[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kRun; ++j) {
                s_m[(xr + j) * CS + r] = hm[j];
                s_q[(xr + j) * CS + r] = hq[j];
                s_p[(xr + j) * CS + r] = hp[j];
            }
        }
        __syncthreads();
        // vertical pass: thread = (column tx, RPT consecutive rows)
        {
            const int ly0 = ty * RPT;
            float2 vm[RPT], vq[RPT];
            float vp[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) { vm[j] = f2(0.f, 0.f); vq[j] = f2(0.f, 0.f); vp[j] = 0.f; }
#pragma unroll
            for (int e = 0; e < RPT + kWin - 1; ++e) {
                const float2 m = s_m[tx * CS + ly0 + e], q = s_q[tx * CS + ly0 + e];
                const float p = s_p[tx * CS + ly0 + e];
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const int k = e - j;
                    if (k >= 0 && k < kWin) {
                        const float g = win.g[k];
                        vm[j] = __ffma2_rn(f2(g, g), m, vm[j]);
                        vq[j] = __ffma2_rn(f2(g, g), q, vq[j]);
                        vp[j] = fmaf(g, p, vp[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int ly = ly0 + j, py = y0 + ly;
                const float mu1 = vm[j].x, mu2 = vm[j].y, exx = vq[j].x, eyy = vq[j].y, exy = vp[j];
                const bool inside = px < (int)W && py < (int)H;
                const bool valid = inside && px >= kHalf && py >= kHalf && px + kHalf < (int)W && py + kHalf < (int)H;
                float dmu = 0.f, dxx = 0.f, dxy = 0.f;
                if (valid) {
                    const float A = mu1 * mu1 + mu2 * mu2 + kC1;
                    const float B = (exx - mu1 * mu1) + (eyy - mu2 * mu2) + kC2;
                    const float Cn = 2.f * mu1 * mu2 + kC1;
                    const float Dn = 2.f * (exy - mu1 * mu2) + kC2;
                    const float iA = 1.f / A, iB = 1.f / B, iAB = iA * iB;
                    const float val = Cn * Dn * iAB;
                    ss += val;
                    // derivatives of val w.r.t. (mu1, E[x²], E[xy]) as independent variables
                    dxx = -val * iB;
                    dxy = 2.f * Cn * iAB;
                    dmu = 2.f * (mu2 * (Dn - Cn) * iAB + mu1 * val * (iB - iA));
                }
                if (inside) {
                    const float2 ab = s_ab[(ly + kHalf) * kRowS + (tx + kHalf) * 3 + ch];
                    l1 += fabsf(ab.x - ab.y);
                }
                o_mu[j][ch] = dmu; o_xx[j][ch] = dxx; o_xy[j][ch] = dxy;
            }
        }
        __syncthreads();
    }
    if (d_mu != nullptr) {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int py = y0 + ty * RPT + j;
            if (px < (int)W && py < (int)H) {
                const size_t o = img_off + ((size_t)py * W + px) * 3;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) { d_mu[o + ch] = o_mu[j][ch]; d_xx[o + ch] = o_xx[j][ch]; d_xy[o + ch] = o_xy[j][ch]; }
            }
        }
    }
    // block reduction of the two sums (deterministic: fixed tree, one slot per block)
    __shared__ float red[2][8];
    for (int o = 16; o >= 1; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (tx == 0) { red[0][ty] = l1; red[1][ty] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
        const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partials[2 * blk] = a;
        partials[2 * blk + 1] = b;
    }
}

// out[0] = loss, out[1] = L1 mean, out[2] = SSIM mean
__global__ void __launch_bounds__(256)
l1_ssim_finalize_kernel(uint32_t n_blocks, const float *__restrict__ partials, double n_l1, double n_ssim,
                        float lambda, float *__restrict__ out) {
    __shared__ double red[2][256];
    double a = 0.0, b = 0.0;
    for (uint32_t i = threadIdx.x; i < n_blocks; i += blockDim.x) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    red[0][threadIdx.x] = a; red[1][threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double l1 = red[0][0] / n_l1, ssim = red[1][0] / n_ssim;
        out[0] = (float)(l1 * (1.0 - (double)lambda) + (1.0 - ssim) * (double)lambda);
        out[1] = (float)l1;
        out[2] = (float)ssim;
    }
}

// v_img = v_loss * [ (1-lambda)/n_l1 * sign(img - tgt)
//                    - lambda/n_ssim * (G*d_mu + 2 img (G*d_xx) + tgt (G*d_xy)) ]      (G* = zero-padded window)
// TY: output rows per block (32, or 16: a 26-row apron = 50 KB of shared memory, 4 blocks/SM instead of 2)
template <int TY>
__global__ void __launch_bounds__(256, TY == 16 ? 4 : 3)
l1_ssim_bwd_kernel(uint32_t H, uint32_t W, const float *__restrict__ img, const float *__restrict__ tgt,
                   const Window win, const float *__restrict__ d_mu, const float *__restrict__ d_xx,
                   const float *__restrict__ d_xy, const float *__restrict__ v_loss, float s_l1, float s_ssim,
                   float *__restrict__ v_img) {
    constexpr int HY = TY + 2 * kHalf;        // apron rows
    constexpr int CS = HY + 1;                // transposed column stride (odd)
    constexpr int RPT = TY / 8;               // consecutive output rows per thread, vertical pass
    extern __shared__ float2 smem2[];
    float2 *s_ab = smem2;                        // [HY][kRowS]  {d_mu, d_xx}
    float2 *s_m = s_ab + HY * kRowS;          // [kTile][CS]  horizontal sums of the pair
    float *s_c = reinterpret_cast<float *>(s_m + kTile * CS);   // [HY][kRowS]  d_xy
    float *s_p = s_c + HY * kRowS;            // [kTile][CS]
    const uint32_t cam = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * TY;
    const size_t img_off = (size_t)cam * H * W * 3;
    stage_tiles<true>(d_mu + img_off, d_xx + img_off, d_xy + img_off, H, W, x0, y0, s_ab, s_c, HY);
    __syncthreads();
    const float gl = v_loss != nullptr ? __ldg(v_loss) : 1.f;
    const float k_l1 = gl * s_l1, k_ss = gl * s_ssim;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int px = x0 + tx;
    float out[RPT][3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        if (threadIdx.x < HY * kRunsPerRow) {
            const int r = threadIdx.x % HY, xr = run_start(threadIdx.x / HY);
            float2 hm[kRun];
            float hp[kRun];
#pragma unroll
            for (int j = 0; j < kRun; ++j) { hm[j] = f2(0.f, 0.f); hp[j] = 0.f; }
#pragma unroll
            for (int e = 0; e < kRun + kWin - 1; ++e) {
                const float2 ab = s_ab[r * kRowS + (xr + e) * 3 + ch];
                const float c = s_c[r * kRowS + (xr + e) * 3 + ch];
#pragma unroll
                for (int j = 0; j < kRun; ++j) {
                    const int k = e - j;
                    if (k >= 0 && k < kWin) {
                        const float g = win.g[k];
                        hm[j] = __ffma2_rn(f2(g, g), ab, hm[j]);
                        hp[j] = fmaf(g, c, hp[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kRun; ++j) {
                s_m[(xr + j) * CS + r] = hm[j];
                s_p[(xr + j) * CS + r] = hp[j];
            }
        }
        __syncthreads();
        {
            const int ly0 = ty * RPT;
            float2 vm[RPT];
            float vp[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) { vm[j] = f2(0.f, 0.f); vp[j] = 0.f; }
#pragma unroll
            for (int e = 0; e < RPT + kWin - 1; ++e) {
                const float2 m = s_m[tx * CS + ly0 + e];
                const float p = s_p[tx * CS + ly0 + e];
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const int k = e - j;
                    if (k >= 0 && k < kWin) {
                        const float g = win.g[k];
                        vm[j] = __ffma2_rn(f2(g, g), m, vm[j]);
                        vp[j] = fmaf(g, p, vp[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int py = y0 + ly0 + j;
                float v = 0.f;
                if (px < (int)W && py < (int)H) {
                    const size_t o = img_off + ((size_t)py * W + px) * 3 + ch;
                    const float a = __ldg(img + o), b = __ldg(tgt + o);
                    const float sgn = a > b ? 1.f : (a < b ? -1.f : 0.f);
                    v = k_l1 * sgn + k_ss * (vm[j].x + 2.f * a * vm[j].y + b * vp[j]);
                }
                out[j][ch] = v;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int py = y0 + ty * RPT + j;
        if (px < (int)W && py < (int)H) {
            const size_t o = img_off + ((size_t)py * W + px) * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) v_img[o + ch] = out[j][ch];
        }
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b200splat_splat_activations_fwd(uint32_t N, const float *scales_raw, const float *opacities_raw,
                                               float *scales, float *opacities, void *stream) {
    if (N == 0) return 0;
    B2S_REQUIRE(N <= 0x55555555u, "b200splat_splat_activations_fwd", "N too large");
    splat_activations_fwd_kernel<<<div_up(3 * (uint64_t)N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        N, scales_raw, opacities_raw, scales, opacities);
    B2S_CHECK_LAUNCH("b200splat_splat_activations_fwd");
    return 0;
}

extern "C" int b200splat_splat_activations_bwd(uint32_t N, const float *scales, const float *opacities,
                                               const float *v_scales, const float *v_opacities, float *v_scales_raw,
                                               float *v_opacities_raw, void *stream) {
    if (N == 0) return 0;
    B2S_REQUIRE(N <= 0x55555555u, "b200splat_splat_activations_bwd", "N too large");
    splat_activations_bwd_kernel<<<div_up(3 * (uint64_t)N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        N, scales, opacities, v_scales, v_opacities, v_scales_raw, v_opacities_raw);
    B2S_CHECK_LAUNCH("b200splat_splat_activations_bwd");
    return 0;
}

extern "C" int b200splat_invert_4x4(uint32_t C, const float *mats, float *out, void *stream) {
    if (C == 0) return 0;
    invert_4x4_kernel<<<div_up(C, 64), 64, 0, (cudaStream_t)stream>>>(C, mats, out);
    B2S_CHECK_LAUNCH("b200splat_invert_4x4");
    return 0;
}

constexpr int kFwdTY = 32;  // output rows per forward block (16 measured slower for the forward: 0.104 vs 0.089 ms)

static inline uint32_t ssim_blocks(uint32_t C, uint32_t H, uint32_t W) {
    return C * div_up(H, kFwdTY) * div_up(W, kTile);
}

extern "C" size_t b200splat_l1_ssim_workspace_bytes(uint32_t C, uint32_t H, uint32_t W) {
    return (size_t)ssim_blocks(C, H, W) * 2 * sizeof(float);
}

extern "C" int b200splat_l1_ssim_fwd(uint32_t C, uint32_t H, uint32_t W, const float *img, const float *target,
                                     float ssim_lambda, float *d_mu, float *d_xx, float *d_xy, float *out3,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    const char *where = "b200splat_l1_ssim_fwd";
    B2S_REQUIRE(H >= (uint32_t)kWin && W >= (uint32_t)kWin, where, "image smaller than the 11x11 SSIM window");
    B2S_REQUIRE(C >= 1 && C <= 65535, where, "1..65535 images");
    B2S_REQUIRE((d_mu != nullptr) == (d_xx != nullptr) && (d_mu != nullptr) == (d_xy != nullptr), where,
                "derivative maps go together");
    B2S_REQUIRE(workspace != nullptr && workspace_bytes >= b200splat_l1_ssim_workspace_bytes(C, H, W), where,
                "workspace too small (see b200splat_l1_ssim_workspace_bytes)");
    static const Window win = make_window();
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(div_up(W, kTile), div_up(H, kFwdTY), C);
    const size_t smem = (size_t)((kFwdTY + 2 * kHalf) * kRowS * 2 + kTile * (kFwdTY + 2 * kHalf + 1) * 5) * sizeof(float);
    cudaFuncSetAttribute(l1_ssim_fwd_kernel<kFwdTY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float *partials = reinterpret_cast<float *>(workspace);
    l1_ssim_fwd_kernel<kFwdTY><<<grid, 256, smem, st>>>(H, W, img, target, win, d_mu, d_xx, d_xy, partials);
    B2S_CHECK_LAUNCH(where);
    const double n_l1 = (double)C * H * W * 3, n_ssim = (double)C * (H - 2 * kHalf) * (W - 2 * kHalf) * 3;
    l1_ssim_finalize_kernel<<<1, 256, 0, st>>>(ssim_blocks(C, H, W), partials, n_l1, n_ssim, ssim_lambda, out3);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_l1_ssim_bwd(uint32_t C, uint32_t H, uint32_t W, const float *img, const float *target,
                                     float ssim_lambda, const float *d_mu, const float *d_xx, const float *d_xy,
                                     const float *v_loss, float *v_img, void *stream) {
    const char *where = "b200splat_l1_ssim_bwd";
    B2S_REQUIRE(H >= (uint32_t)kWin && W >= (uint32_t)kWin, where, "image smaller than the 11x11 SSIM window");
    B2S_REQUIRE(C >= 1 && C <= 65535, where, "1..65535 images");
    static const Window win = make_window();
    cudaStream_t st = (cudaStream_t)stream;
    const bool tall = tuning_variant() == 9;  // A/B: 32-row tiles (2 blocks/SM)
    const int TYv = tall ? 32 : 16;
    const dim3 grid(div_up(W, kTile), div_up(H, TYv), C);
    const size_t smem = (size_t)((TYv + 2 * kHalf) * kRowS * 3 + kTile * (TYv + 2 * kHalf + 1) * 3) * sizeof(float);
    cudaFuncSetAttribute(l1_ssim_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(l1_ssim_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const double n_l1 = (double)C * H * W * 3, n_ssim = (double)C * (H - 2 * kHalf) * (W - 2 * kHalf) * 3;
    const float k1 = (float)((1.0 - (double)ssim_lambda) / n_l1), k2 = (float)(-(double)ssim_lambda / n_ssim);
    if (tall) l1_ssim_bwd_kernel<32><<<grid, 256, smem, st>>>(H, W, img, target, win, d_mu, d_xx, d_xy, v_loss, k1, k2, v_img);
    else l1_ssim_bwd_kernel<16><<<grid, 256, smem, st>>>(H, W, img, target, win, d_mu, d_xx, d_xy, v_loss, k1, k2, v_img);
    B2S_CHECK_LAUNCH(where);
    return 0;
}
