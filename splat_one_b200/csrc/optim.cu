// optim.cu — optimizer / densifier-side kernels fed by the rasterization path:
//   selective_adam_update   CS/adam.cu:16-44 (host :46-82), used by gsplat.optimizers.SelectiveAdam
//                           (splat_one: visible_adam, utils/gsplat_utils/gsplat_trainer.py:719-730)
//   compute_relocation      CS/compute_relocation.cu:6-39 (MCMC strategy, Eq. 9 of
//                           "3D Gaussian Splatting as Markov Chain Monte Carlo")
//   strategy_update_state   G/strategy/default.py:239-262 (`DefaultStrategy._update_state`: running
//                           grad2d / count / radii statistics that drive densification), one kernel
//                           instead of ~10 ATen launches; the unpacked, non-absgrad case is also
//                           available folded into the projection backward (projection.cu)
// The first two are streaming maps.  Adam: 16 B read + 12 B written per updated element, nothing at
// all for invisible Gaussians (the visibility byte is read once per element through L1).
#include "common.cuh"

namespace b2s {

// One thread per parameter element; M elements per Gaussian.  Exactly the reference's
// update: no bias correction, step = -lr·m / (sqrt(v) + eps), only where visible.
static __global__ void __launch_bounds__(kThreads)
selective_adam_kernel(float *__restrict__ param, const float *__restrict__ grad, float *__restrict__ exp_avg,
                      float *__restrict__ exp_avg_sq, const uint8_t *__restrict__ visible, float lr, float b1,
                      float b2, float eps, uint64_t total, uint32_t M) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    if (!visible[p / M]) return;
    const float g = __ldcs(grad + p);
    const float m = b1 * exp_avg[p] + (1.0f - b1) * g;
    const float v = b2 * exp_avg_sq[p] + (1.0f - b2) * g * g;
    param[p] += -lr * m / (sqrtf(v) + eps);
    exp_avg[p] = m;
    exp_avg_sq[p] = v;
}

// new_opacity = 1 - (1 - o)^(1/n);  denom = sum_{i=1..n} sum_{k<i} C(i-1,k) (-1)^k / sqrt(k+1) · new_opacity^(k+1);
// new_scale = o / denom · scale.  The reference evaluates two pow() per term (up to 1275 terms);
// the k-dependent factor is hoisted here (n terms) and the double sum keeps its order.
static __global__ void __launch_bounds__(kThreads)
relocation_kernel(uint32_t N, const float *__restrict__ opacities, const float *__restrict__ scales,
                  const int32_t *__restrict__ ratios, const float *__restrict__ binoms, int n_max,
                  float *__restrict__ new_opacities, float *__restrict__ new_scales) {
    constexpr int kMaxN = 64;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N) return;
    const int n = ratios[idx];
    const float o = opacities[idx];
    const float no = 1.0f - powf(1.0f - o, 1.0f / (float)n);
    new_opacities[idx] = no;
    float term[kMaxN];
    float pw = no, sgn = 1.f;
    const int nk = n < kMaxN ? n : kMaxN;
    for (int k = 0; k < nk; ++k) {
        term[k] = sgn / sqrtf((float)(k + 1)) * pw;
        pw *= no;
        sgn = -sgn;
    }
    float denom = 0.f;
    for (int i = 1; i <= nk; ++i)
        for (int k = 0; k < i; ++k) denom += binoms[(i - 1) * n_max + k] * term[k];
    const float coeff = o / denom;
#pragma unroll
    for (int j = 0; j < 3; ++j) new_scales[3 * (size_t)idx + j] = coeff * scales[3 * (size_t)idx + j];
}

// Unpacked layout: one thread per Gaussian loops over the cameras (no atomics, deterministic).
static __global__ void __launch_bounds__(kThreads)
strategy_state_kernel(uint32_t C, uint32_t N, const float2 *__restrict__ grads, const int32_t *__restrict__ radii,
                      float sx, float sy, float max_wh, float *__restrict__ grad2d, float *__restrict__ count,
                      float *__restrict__ radii_state) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float g = 0.f, c = 0.f, r = 0.f;
    for (uint32_t cid = 0; cid < C; ++cid) {
        const uint64_t idx = (uint64_t)cid * N + n;
        const int32_t rad = radii[idx];
        if (rad > 0) {
            const float2 v = __ldcs(grads + idx);
            const float gx = v.x * sx, gy = v.y * sy;
            g += sqrtf(gx * gx + gy * gy);
            c += 1.f;
            r = fmaxf(r, __fdiv_rn((float)rad, max_wh));
        }
    }
    if (c > 0.f) {
        grad2d[n] += g;
        count[n] += c;
        if (radii_state != nullptr) radii_state[n] = fmaxf(radii_state[n], r);
    }
}

// Packed (COO) layout: one thread per visible (camera, Gaussian) pair; a Gaussian seen by several cameras
// is updated with atomics (the radius as an integer max: non-negative floats order like their bit patterns).
static __global__ void __launch_bounds__(kThreads)
strategy_state_packed_kernel(uint32_t nnz, const int64_t *__restrict__ gaussian_ids, const float2 *__restrict__ grads,
                             const int32_t *__restrict__ radii, float sx, float sy, float max_wh,
                             float *__restrict__ grad2d, float *__restrict__ count, float *__restrict__ radii_state) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t n = gaussian_ids[i];
    const float2 v = __ldcs(grads + i);
    const float gx = v.x * sx, gy = v.y * sy;
    atomicAdd(grad2d + n, sqrtf(gx * gx + gy * gy));
    atomicAdd(count + n, 1.f);
    if (radii_state != nullptr) {
        const float r = fmaxf(__fdiv_rn((float)radii[i], max_wh), 0.f);
        atomicMax(reinterpret_cast<int *>(radii_state + n), __float_as_int(r));
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b200splat_strategy_update_state(uint32_t C, uint32_t N, uint32_t nnz, const int64_t *gaussian_ids,
                                               const float *grads, const int32_t *radii, float grad_scale_x,
                                               float grad_scale_y, float max_wh, float *state_grad2d,
                                               float *state_count, float *state_radii, void *stream) {
    const char *where = "b200splat_strategy_update_state";
    B2S_REQUIRE(state_grad2d != nullptr && state_count != nullptr, where, "grad2d and count state arrays are required");
    cudaStream_t st = (cudaStream_t)stream;
    const float2 *g2 = reinterpret_cast<const float2 *>(grads);
    if (gaussian_ids != nullptr) {
        if (nnz == 0) return 0;
        strategy_state_packed_kernel<<<div_up(nnz, kThreads), kThreads, 0, st>>>(
            nnz, gaussian_ids, g2, radii, grad_scale_x, grad_scale_y, max_wh, state_grad2d, state_count, state_radii);
    } else {
        if ((uint64_t)C * N == 0) return 0;
        strategy_state_kernel<<<div_up(N, kThreads), kThreads, 0, st>>>(C, N, g2, radii, grad_scale_x, grad_scale_y,
                                                                        max_wh, state_grad2d, state_count, state_radii);
    }
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_selective_adam_update(float *param, const float *param_grad, float *exp_avg,
                                               float *exp_avg_sq, const uint8_t *visible, float lr, float b1, float b2,
                                               float eps, uint32_t N, uint32_t M, void *stream) {
    const char *where = "b200splat_selective_adam_update";
    const uint64_t total = (uint64_t)N * M;
    if (total == 0) return 0;
    selective_adam_kernel<<<div_up(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        param, param_grad, exp_avg, exp_avg_sq, visible, lr, b1, b2, eps, total, M);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_compute_relocation(uint32_t N, const float *opacities, const float *scales,
                                            const int32_t *ratios, const float *binoms, int n_max,
                                            float *new_opacities, float *new_scales, void *stream) {
    const char *where = "b200splat_compute_relocation";
    B2S_REQUIRE(n_max >= 1 && n_max <= 64, where, "n_max must be in [1, 64]");
    if (N == 0) return 0;
    relocation_kernel<<<div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(N, opacities, scales, ratios, binoms,
                                                                                 n_max, new_opacities, new_scales);
    B2S_CHECK_LAUNCH(where);
    return 0;
}
