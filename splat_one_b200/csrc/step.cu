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

// stage the (kHalo x kHalo x 3) apron tile of one [H,W,3] image, zero outside the image
__device__ __forceinline__ void stage_tile(const float *__restrict__ img, uint32_t H, uint32_t W, int x0, int y0,
                                           float *s) {
    for (int i = threadIdx.x; i < kHalo * kRowF; i += blockDim.x) {
        const int r = i / kRowF, f = i - r * kRowF;
        const int px3 = f / 3, y = y0 - kHalf + r, x = x0 - kHalf + px3;
        float v = 0.f;
        if (y >= 0 && y < (int)H && x >= 0 && x < (int)W) v = __ldg(img + ((size_t)y * W + x) * 3 + (f - 3 * px3));
        s[i] = v;
    }
}

__global__ void __launch_bounds__(256)
l1_ssim_fwd_kernel(uint32_t H, uint32_t W, const float *__restrict__ img, const float *__restrict__ tgt,
                   const Window win, float *__restrict__ d_mu, float *__restrict__ d_xx, float *__restrict__ d_xy,
                   float *__restrict__ partials) {
    extern __shared__ float smem[];
    float *s_x = smem;                          // [kHalo][kRowF]
    float *s_y = s_x + kHalo * kRowF;
    float *s_h = s_y + kHalo * kRowF;           // [5][kHalo][kTile]
    const uint32_t cam = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t img_off = (size_t)cam * H * W * 3;
    stage_tile(img + img_off, H, W, x0, y0, s_x);
    stage_tile(tgt + img_off, H, W, x0, y0, s_y);
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int px = x0 + tx;
    float l1 = 0.f, ss = 0.f;
    float o_mu[4][3], o_xx[4][3], o_xy[4][3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        // horizontal pass: 5 running statistics per (row, column)
        for (int i = threadIdx.x; i < kHalo * kTile; i += blockDim.x) {
            const int r = i >> 5, x = i & 31;
            float hx = 0.f, hy = 0.f, hxx = 0.f, hyy = 0.f, hxy = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float a = s_x[r * kRowF + (x + k) * 3 + ch], b = s_y[r * kRowF + (x + k) * 3 + ch];
                const float g = win.g[k], ga = g * a, gb = g * b;
                hx += ga; hy += gb; hxx = fmaf(ga, a, hxx); hyy = fmaf(gb, b, hyy); hxy = fmaf(ga, b, hxy);
            }
            s_h[(0 * kHalo + r) * kTile + x] = hx;
            s_h[(1 * kHalo + r) * kTile + x] = hy;
            s_h[(2 * kHalo + r) * kTile + x] = hxx;
            s_h[(3 * kHalo + r) * kTile + x] = hyy;
            s_h[(4 * kHalo + r) * kTile + x] = hxy;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ly = ty + 8 * j, py = y0 + ly;
            float mu1 = 0.f, mu2 = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float g = win.g[k];
                mu1 = fmaf(g, s_h[(0 * kHalo + ly + k) * kTile + tx], mu1);
                mu2 = fmaf(g, s_h[(1 * kHalo + ly + k) * kTile + tx], mu2);
                exx = fmaf(g, s_h[(2 * kHalo + ly + k) * kTile + tx], exx);
                eyy = fmaf(g, s_h[(3 * kHalo + ly + k) * kTile + tx], eyy);
                exy = fmaf(g, s_h[(4 * kHalo + ly + k) * kTile + tx], exy);
            }
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
                const float a = s_x[(ly + kHalf) * kRowF + (tx + kHalf) * 3 + ch];
                const float b = s_y[(ly + kHalf) * kRowF + (tx + kHalf) * 3 + ch];
                l1 += fabsf(a - b);
            }
            o_mu[j][ch] = dmu; o_xx[j][ch] = dxx; o_xy[j][ch] = dxy;
        }
        __syncthreads();
    }
    if (d_mu != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int py = y0 + ty + 8 * j;
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
__global__ void __launch_bounds__(256)
l1_ssim_bwd_kernel(uint32_t H, uint32_t W, const float *__restrict__ img, const float *__restrict__ tgt,
                   const Window win, const float *__restrict__ d_mu, const float *__restrict__ d_xx,
                   const float *__restrict__ d_xy, const float *__restrict__ v_loss, float s_l1, float s_ssim,
                   float *__restrict__ v_img) {
    extern __shared__ float smem[];
    float *s_m[3] = {smem, smem + kHalo * kRowF, smem + 2 * kHalo * kRowF};
    float *s_h = smem + 3 * kHalo * kRowF;      // [3][kHalo][kTile]
    const uint32_t cam = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t img_off = (size_t)cam * H * W * 3;
    stage_tile(d_mu + img_off, H, W, x0, y0, s_m[0]);
    stage_tile(d_xx + img_off, H, W, x0, y0, s_m[1]);
    stage_tile(d_xy + img_off, H, W, x0, y0, s_m[2]);
    __syncthreads();
    const float gl = v_loss != nullptr ? __ldg(v_loss) : 1.f;
    const float k_l1 = gl * s_l1, k_ss = gl * s_ssim;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int px = x0 + tx;
    float out[4][3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        for (int i = threadIdx.x; i < kHalo * kTile; i += blockDim.x) {
            const int r = i >> 5, x = i & 31;
            float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float g = win.g[k];
                const int o = r * kRowF + (x + k) * 3 + ch;
                h0 = fmaf(g, s_m[0][o], h0); h1 = fmaf(g, s_m[1][o], h1); h2 = fmaf(g, s_m[2][o], h2);
            }
            s_h[(0 * kHalo + r) * kTile + x] = h0;
            s_h[(1 * kHalo + r) * kTile + x] = h1;
            s_h[(2 * kHalo + r) * kTile + x] = h2;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ly = ty + 8 * j, py = y0 + ly;
            float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float g = win.g[k];
                c0 = fmaf(g, s_h[(0 * kHalo + ly + k) * kTile + tx], c0);
                c1 = fmaf(g, s_h[(1 * kHalo + ly + k) * kTile + tx], c1);
                c2 = fmaf(g, s_h[(2 * kHalo + ly + k) * kTile + tx], c2);
            }
            float v = 0.f;
            if (px < (int)W && py < (int)H) {
                const size_t o = img_off + ((size_t)py * W + px) * 3 + ch;
                const float a = __ldg(img + o), b = __ldg(tgt + o);
                const float sgn = a > b ? 1.f : (a < b ? -1.f : 0.f);
                v = k_l1 * sgn + k_ss * (c0 + 2.f * a * c1 + b * c2);
            }
            out[j][ch] = v;
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int py = y0 + ty + 8 * j;
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

static inline uint32_t ssim_blocks(uint32_t C, uint32_t H, uint32_t W) {
    return C * div_up(H, kTile) * div_up(W, kTile);
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
    const dim3 grid(div_up(W, kTile), div_up(H, kTile), C);
    const size_t smem = (size_t)(2 * kHalo * kRowF + 5 * kHalo * kTile) * sizeof(float);
    cudaFuncSetAttribute(l1_ssim_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float *partials = reinterpret_cast<float *>(workspace);
    l1_ssim_fwd_kernel<<<grid, 256, smem, st>>>(H, W, img, target, win, d_mu, d_xx, d_xy, partials);
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
    const dim3 grid(div_up(W, kTile), div_up(H, kTile), C);
    const size_t smem = (size_t)(3 * kHalo * kRowF + 3 * kHalo * kTile) * sizeof(float);
    cudaFuncSetAttribute(l1_ssim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const double n_l1 = (double)C * H * W * 3, n_ssim = (double)C * (H - 2 * kHalf) * (W - 2 * kHalf) * 3;
    l1_ssim_bwd_kernel<<<grid, 256, smem, st>>>(H, W, img, target, win, d_mu, d_xx, d_xy, v_loss,
                                                (float)((1.0 - (double)ssim_lambda) / n_l1),
                                                (float)(-(double)ssim_lambda / n_ssim), v_img);
    B2S_CHECK_LAUNCH(where);
    return 0;
}
