// sh.cu — spherical-harmonics colour evaluation (a5), forward and backward.
// Replaces CS/compute_sh_fwd.cu:12-71 and CS/compute_sh_bwd.cu:13-98; the closed forms
// are Sloan's "Efficient Spherical Harmonic Evaluation" (JCGT 2013) as used by
// CS/spherical_harmonics.cuh:17-366, up to degree 4 (25 bases).
//
// Design (B200): the stage is pure HBM streaming — 12·K bytes of coefficients per
// visible element dominate.  One thread owns one element (all three channels, so the
// basis is evaluated once instead of three times as in the reference's
// thread-per-channel mapping), coefficients are fetched with 128-bit read-only loads
// when the row is 16-byte aligned, and the backward writes every byte of v_coeffs
// exactly once (zeros included) so no separate zero-fill pass is needed; v_dirs is
// produced without atomics.
#include "common.cuh"

namespace b2s {

// Calls f(k, B, dB/dx, dB/dy, dB/dz) for every basis k < (deg+1)^2 at unit direction
// (x,y,z).  With WITH_GRAD=false the derivative arguments are 0 and dead-code eliminated.
template <bool WITH_GRAD, typename F>
__device__ __forceinline__ void sh_for_each_basis(const uint32_t deg, const float x, const float y, const float z,
                                                  F &&f) {
    f(0, 0.2820947917738781f, 0.f, 0.f, 0.f);
    if (deg < 1) return;
    constexpr float c1 = 0.48860251190292f;
    f(1, -c1 * y, 0.f, -c1, 0.f);
    f(2, c1 * z, 0.f, 0.f, c1);
    f(3, -c1 * x, -c1, 0.f, 0.f);
    if (deg < 2) return;
    const float z2 = z * z;
    const float fTmp0B = -1.092548430592079f * z;
    const float fC1 = x * x - y * y;
    const float fS1 = 2.f * x * y;
    constexpr float c2 = 0.5462742152960395f;
    const float fC1_x = 2.f * x, fC1_y = -2.f * y, fS1_x = 2.f * y, fS1_y = 2.f * x;
    const float pSH6 = 0.9461746957575601f * z2 - 0.3153915652525201f;
    const float pSH6_z = 2.f * 0.9461746957575601f * z;
    f(4, c2 * fS1, c2 * fS1_x, c2 * fS1_y, 0.f);
    f(5, fTmp0B * y, 0.f, fTmp0B, -1.092548430592079f * y);
    f(6, pSH6, 0.f, 0.f, pSH6_z);
    f(7, fTmp0B * x, fTmp0B, 0.f, -1.092548430592079f * x);
    f(8, c2 * fC1, c2 * fC1_x, c2 * fC1_y, 0.f);
    if (deg < 3) return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float fTmp1B = 1.445305721320277f * z;
    const float fC2 = x * fC1 - y * fS1;
    const float fS2 = x * fS1 + y * fC1;
    constexpr float c3 = -0.5900435899266435f;
    const float fTmp0C_z = -2.285228997322329f * 2.f * z;
    const float fC2_x = fC1 + x * fC1_x - y * fS1_x;
    const float fC2_y = x * fC1_y - fS1 - y * fS1_y;
    const float fS2_x = fS1 + x * fS1_x + y * fC1_x;
    const float fS2_y = x * fS1_y + fC1 + y * fC1_y;
    const float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    const float pSH12_z = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    f(9, c3 * fS2, c3 * fS2_x, c3 * fS2_y, 0.f);
    f(10, fTmp1B * fS1, fTmp1B * fS1_x, fTmp1B * fS1_y, 1.445305721320277f * fS1);
    f(11, fTmp0C * y, 0.f, fTmp0C, fTmp0C_z * y);
    f(12, pSH12, 0.f, 0.f, pSH12_z);
    f(13, fTmp0C * x, fTmp0C, 0.f, fTmp0C_z * x);
    f(14, fTmp1B * fC1, fTmp1B * fC1_x, fTmp1B * fC1_y, 1.445305721320277f * fC1);
    f(15, c3 * fC2, c3 * fC2_x, c3 * fC2_y, 0.f);
    if (deg < 4) return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float fTmp2B = -1.770130769779931f * z;
    const float fC3 = x * fC2 - y * fS2;
    const float fS3 = x * fS2 + y * fC2;
    constexpr float c4 = 0.6258357354491763f;
    const float fTmp0D_z = 3.f * -4.683325804901025f * z2 + 2.007139630671868f;
    const float fTmp1C_z = 2.f * 3.31161143515146f * z;
    const float fC3_x = fC2 + x * fC2_x - y * fS2_x;
    const float fC3_y = x * fC2_y - fS2 - y * fS2_y;
    const float fS3_x = fS2 + y * fC2_x + x * fS2_x;
    const float fS3_y = x * fS2_y + fC2 + y * fC2_y;
    const float pSH20 = 1.984313483298443f * z * pSH12 - 1.006230589874905f * pSH6;
    const float pSH20_z = 1.984313483298443f * (pSH12 + z * pSH12_z) - 1.006230589874905f * pSH6_z;
    f(16, c4 * fS3, c4 * fS3_x, c4 * fS3_y, 0.f);
    f(17, fTmp2B * fS2, fTmp2B * fS2_x, fTmp2B * fS2_y, -1.770130769779931f * fS2);
    f(18, fTmp1C * fS1, fTmp1C * fS1_x, fTmp1C * fS1_y, fTmp1C_z * fS1);
    f(19, fTmp0D * y, 0.f, fTmp0D, fTmp0D_z * y);
    f(20, pSH20, 0.f, 0.f, pSH20_z);
    f(21, fTmp0D * x, fTmp0D, 0.f, fTmp0D_z * x);
    f(22, fTmp1C * fC1, fTmp1C * fC1_x, fTmp1C * fC1_y, fTmp1C_z * fC1);
    f(23, fTmp2B * fC2, fTmp2B * fC2_x, fTmp2B * fC2_y, -1.770130769779931f * fC2);
    f(24, c4 * fC3, c4 * fC3_x, c4 * fC3_y, 0.f);
}

// Row loader: 3*NB floats of one coefficient row into registers (128-bit loads when the
// row start is 16-byte aligned and the length is a multiple of 4).
template <int NB>
__device__ __forceinline__ void load_row(const float *__restrict__ row, float (&c)[NB * 3], bool vec_ok) {
    if ((NB * 3) % 4 == 0 && vec_ok) {
        const float4 *p = reinterpret_cast<const float4 *>(row);
#pragma unroll
        for (int k = 0; k < NB * 3 / 4; k++) {
            const float4 q = __ldg(p + k);
            c[4 * k] = q.x; c[4 * k + 1] = q.y; c[4 * k + 2] = q.z; c[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NB * 3; k++) c[k] = __ldg(row + k);
    }
}

// NB = number of active bases = (deg+1)^2 (compile time: 1,4,9,16,25).
template <int NB>
__global__ void __launch_bounds__(kThreads)
sh_fwd_kernel(uint32_t n_elems, uint32_t n_rows, uint32_t K, uint32_t deg, const float *__restrict__ dirs,
              const float *__restrict__ coeffs, const uint8_t *__restrict__ masks, float *__restrict__ colors) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    if (masks != nullptr && !masks[e]) return;
    const uint32_t row_id = (n_rows == n_elems) ? e : (e % n_rows);
    const float *row = coeffs + (size_t)row_id * K * 3;
    float c[NB * 3];
    load_row<NB>(row, c, ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(coeffs) & 15) == 0));
    float x = 0.f, y = 0.f, z = 0.f;
    if (NB > 1) {
        const float dx = dirs[3 * (size_t)e], dy = dirs[3 * (size_t)e + 1], dz = dirs[3 * (size_t)e + 2];
        const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
        x = dx * inorm; y = dy * inorm; z = dz * inorm;
    }
    float r = 0.f, g = 0.f, b = 0.f;
    sh_for_each_basis<false>(deg, x, y, z, [&](int k, float B, float, float, float) {
        r += B * c[3 * k]; g += B * c[3 * k + 1]; b += B * c[3 * k + 2];
    });
    colors[3 * (size_t)e] = r; colors[3 * (size_t)e + 1] = g; colors[3 * (size_t)e + 2] = b;
}

template <int NB>
__global__ void __launch_bounds__(kThreads)
sh_bwd_kernel(uint32_t n_elems, uint32_t n_rows, uint32_t K, uint32_t deg, const float *__restrict__ dirs,
              const float *__restrict__ coeffs, const uint8_t *__restrict__ masks,
              const float *__restrict__ v_colors, float *__restrict__ v_coeffs, float *__restrict__ v_dirs) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    float *vrow = v_coeffs + (size_t)e * K * 3;
    const bool vec_ok_out = ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(v_coeffs) & 15) == 0);
    const bool active = (masks == nullptr) || masks[e];
    float vc[NB * 3];
    float vx = 0.f, vy = 0.f, vz = 0.f;
    if (active) {
        const float vr = v_colors[3 * (size_t)e], vg = v_colors[3 * (size_t)e + 1], vb = v_colors[3 * (size_t)e + 2];
        float x = 0.f, y = 0.f, z = 0.f, inorm = 0.f;
        if (NB > 1) {
            const float dx = dirs[3 * (size_t)e], dy = dirs[3 * (size_t)e + 1], dz = dirs[3 * (size_t)e + 2];
            inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
            x = dx * inorm; y = dy * inorm; z = dz * inorm;
        }
        if (v_dirs != nullptr && NB > 1) {
            const uint32_t row_id = (n_rows == n_elems) ? e : (e % n_rows);
            float c[NB * 3];
            load_row<NB>(coeffs + (size_t)row_id * K * 3, c,
                         ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(coeffs) & 15) == 0));
            sh_for_each_basis<true>(deg, x, y, z, [&](int k, float B, float Bx, float By, float Bz) {
                vc[3 * k] = B * vr; vc[3 * k + 1] = B * vg; vc[3 * k + 2] = B * vb;
                const float w = c[3 * k] * vr + c[3 * k + 1] * vg + c[3 * k + 2] * vb;
                vx += Bx * w; vy += By * w; vz += Bz * w;
            });
            // d(dir/|dir|)/d(dir): tangent-plane projection scaled by 1/|dir|
            const float d = vx * x + vy * y + vz * z;
            vx = (vx - d * x) * inorm; vy = (vy - d * y) * inorm; vz = (vz - d * z) * inorm;
        } else {
            sh_for_each_basis<false>(deg, x, y, z, [&](int k, float B, float, float, float) {
                vc[3 * k] = B * vr; vc[3 * k + 1] = B * vg; vc[3 * k + 2] = B * vb;
            });
        }
    } else {
#pragma unroll
        for (int k = 0; k < NB * 3; k++) vc[k] = 0.f;
    }
    // write the whole row: active bases, then zeros up to K
    if ((NB * 3) % 4 == 0 && vec_ok_out) {
        float4 *o = reinterpret_cast<float4 *>(vrow);
#pragma unroll
        for (int k = 0; k < NB * 3 / 4; k++) __stcs(o + k, make_float4(vc[4 * k], vc[4 * k + 1], vc[4 * k + 2], vc[4 * k + 3]));
        for (uint32_t k = NB * 3 / 4; k < K * 3 / 4; k++) __stcs(o + k, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
#pragma unroll
        for (int k = 0; k < NB * 3; k++) vrow[k] = vc[k];
        for (uint32_t k = NB * 3; k < K * 3; k++) vrow[k] = 0.f;
    }
    if (v_dirs != nullptr) {
        v_dirs[3 * (size_t)e] = vx; v_dirs[3 * (size_t)e + 1] = vy; v_dirs[3 * (size_t)e + 2] = vz;
    }
}


// ---------------------------------------------------------------------------------------
// Fused view-dependent colours for rasterization() (G/rendering.py:368-392):
//   dirs = means - campos[c];  colors = clamp_min(SH(dirs, coeffs) + 0.5, 0);  masked by
//   radii > 0.  One kernel instead of sub / compare / SH / add / clamp (and, backward,
//   where / SH-bwd / sum over cameras / sub-bwd), and no [C,N,3] `dirs` tensor.
// ---------------------------------------------------------------------------------------
// camera centre = inverse(viewmat)[:3, 3], general 4x4 (adjugate, evaluated in double)
__global__ void camera_centers_kernel(uint32_t C, const float *__restrict__ viewmats, float *__restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) m[i][j] = (double)viewmats[16 * c + 4 * i + j];
    // minor of (row r, col q): determinant of the 3x3 left after deleting them
    auto minor3 = [&](int r, int q) {
        int rr[3], cc[3], a = 0, b = 0;
        for (int i = 0; i < 4; i++) if (i != r) rr[a++] = i;
        for (int j = 0; j < 4; j++) if (j != q) cc[b++] = j;
        return m[rr[0]][cc[0]] * (m[rr[1]][cc[1]] * m[rr[2]][cc[2]] - m[rr[1]][cc[2]] * m[rr[2]][cc[1]]) -
               m[rr[0]][cc[1]] * (m[rr[1]][cc[0]] * m[rr[2]][cc[2]] - m[rr[1]][cc[2]] * m[rr[2]][cc[0]]) +
               m[rr[0]][cc[2]] * (m[rr[1]][cc[0]] * m[rr[2]][cc[1]] - m[rr[1]][cc[1]] * m[rr[2]][cc[0]]);
    };
    double det = 0.0;
    for (int j = 0; j < 4; j++) det += ((j & 1) ? -1.0 : 1.0) * m[0][j] * minor3(0, j);
    // inv[i][3] = cofactor(3, i) / det
    for (int i = 0; i < 3; i++) out[3 * c + i] = (float)((((3 + i) & 1) ? -1.0 : 1.0) * minor3(3, i) / det);
}

template <int NB>
__global__ void __launch_bounds__(kThreads)
sh_colors_fwd_kernel(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, int per_view, const float *__restrict__ means,
                     const float *__restrict__ campos, const float *__restrict__ coeffs,
                     const int32_t *__restrict__ radii, float *__restrict__ colors) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (uint64_t)C * N) return;
    const uint32_t c = e / N, n = e % N;
    float r = 0.f, g = 0.f, b = 0.f;
    // the radius, the coefficient row and the mean are requested together (the row of an invisible
    // Gaussian is wasted bandwidth, but a visible one makes one DRAM round trip instead of two)
    const int32_t radius = __ldcs(radii + e);
    const float *row = coeffs + (per_view ? e : (uint64_t)n) * K * 3;
    float cf[NB * 3];
    load_row<NB>(row, cf, ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(coeffs) & 15) == 0));
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (NB > 1) { mx = __ldg(means + 3 * n); my = __ldg(means + 3 * n + 1); mz = __ldg(means + 3 * n + 2); }
    if (radius > 0) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (NB > 1) {
            const float dx = mx - campos[3 * c], dy = my - campos[3 * c + 1], dz = mz - campos[3 * c + 2];
            const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
            x = dx * inorm; y = dy * inorm; z = dz * inorm;
        }
        sh_for_each_basis<false>(deg, x, y, z, [&](int k, float B, float, float, float) {
            r += B * cf[3 * k]; g += B * cf[3 * k + 1]; b += B * cf[3 * k + 2];
        });
        r = fmaxf(r + 0.5f, 0.f); g = fmaxf(g + 0.5f, 0.f); b = fmaxf(b + 0.5f, 0.f);
    }
    colors[3 * e] = r; colors[3 * e + 1] = g; colors[3 * e + 2] = b;
}

// Where the colour cotangents and the camera centres of camera c come from in the colour backward:
// one local pair of arrays, or — camera-parallel peer exchange (splat_one_b200/distributed.py) — the
// symmetric buffers of the ranks, read IN PLACE over NVLink: block b = c / cams_per_block lives at
// bases[b] + offset_bytes and holds {campos [cams_per_block][3], padding to hdr_floats,
// v_colors [cams_per_block][N][3]} (pre-masked: zero where invisible or clamped).  The all-gather of
// the plain exchange becomes the kernel's own loads, overlapped with its gradient writes.
struct CamSource {
    const float *v_colors;
    const float *campos;
    const unsigned long long *bases;  // device array, one base address per rank; nullptr: local arrays
    unsigned long long offset_bytes;
    uint32_t cams_per_block;
    uint32_t hdr_floats;
};

__device__ __forceinline__ const float *cam_block(const CamSource &s, uint32_t c, uint32_t &local) {
    const uint32_t b = c / s.cams_per_block;
    local = c - b * s.cams_per_block;
    return reinterpret_cast<const float *>(s.bases[b] + s.offset_bytes);
}
// row base of camera c's cotangents ([N,3])
__device__ __forceinline__ const float *cam_cotangents(const CamSource &s, uint32_t c, uint32_t N) {
    if (s.bases == nullptr) return s.v_colors + (size_t)c * N * 3;
    uint32_t l;
    const float *blk = cam_block(s, c, l);
    return blk + s.hdr_floats + (size_t)l * N * 3;
}
__device__ __forceinline__ const float *cam_centre(const CamSource &s, uint32_t c) {
    if (s.bases == nullptr) return s.campos + 3 * (size_t)c;
    uint32_t l;
    const float *blk = cam_block(s, c, l);
    return blk + 3 * l;
}
// CB = cameras whose cotangents are fetched together (loads in flight per thread): 1 for local arrays,
// kPeerCamBatch when they are NVLink loads
constexpr int kPeerCamBatch = 4;

// One thread per Gaussian, loop over cameras: v_coeffs (shared table) and v_means are
// written once, without atomics.  The clamp passes gradient where the clamped colour > 0.
// MASKED: the cotangents arrive pre-masked (peer exchange): no radii / clamp-mask loads are compiled in.
template <int NB, int CB, bool MASKED>
__global__ void __launch_bounds__(kThreads, NB <= 16 ? 2 : 1)   // degree 4 (75 + 75 live values) keeps 1 block
sh_colors_bwd_kernel(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, int per_view, const float *__restrict__ means,
                     const CamSource src, const float *__restrict__ coeffs,
                     const int32_t *__restrict__ radii, const float *__restrict__ colors,
                     float *__restrict__ v_coeffs, float *__restrict__ v_means,
                     uint32_t means_cam_begin, uint32_t means_cam_end) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const bool vec_ok = ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(v_coeffs) & 15) == 0);
    // vc accumulates over the cameras for a shared table; per-view tables flush it per camera
    float vc[NB * 3];
#pragma unroll
    for (int k = 0; k < NB * 3; k++) vc[k] = 0.f;
    float vmx = 0.f, vmy = 0.f, vmz = 0.f;
    const float mx = __ldg(means + 3 * n), my = __ldg(means + 3 * n + 1), mz = __ldg(means + 3 * n + 2);
    auto flush = [&](float *vrow) {
        if ((NB * 3) % 4 == 0 && vec_ok) {
            float4 *o = reinterpret_cast<float4 *>(vrow);
#pragma unroll
            for (int k = 0; k < NB * 3 / 4; k++) __stcs(o + k, make_float4(vc[4 * k], vc[4 * k + 1], vc[4 * k + 2], vc[4 * k + 3]));
            for (uint32_t k = NB * 3 / 4; k < K * 3 / 4; k++) __stcs(o + k, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
#pragma unroll
            for (int k = 0; k < NB * 3; k++) vrow[k] = vc[k];
            for (uint32_t k = NB * 3; k < K * 3; k++) vrow[k] = 0.f;
        }
    };
    // Every load of a camera batch — cotangents, clamp mask, radius, camera centre — and the coefficient row are
    // issued BEFORE anything is consumed: gating them on `visible` (as the arithmetic is) made four dependent
    // DRAM round trips per thread (ncu r2: 56 % long-scoreboard stalls at 16 warps per SM).  The loads of
    // Gaussians that turn out invisible are wasted bandwidth only.
    const bool want_means = v_means != nullptr && NB > 1 && means_cam_begin < means_cam_end;
    const bool cf_vec = ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(coeffs) & 15) == 0);
    // The row is hoisted only where one camera at a time is in flight (CB == 1): with CB cameras' cotangents
    // prefetched (peer exchange) 48 more live registers across the loop spill (measured 0.233 -> 0.41 ms).
    constexpr bool kHoistRow = CB == 1;
    float cf[NB * 3];
    if (kHoistRow && want_means && !per_view) load_row<NB>(coeffs + (uint64_t)n * K * 3, cf, cf_vec);
    for (uint32_t c0 = 0; c0 < C; c0 += CB) {
      float vin[CB][3], cin[CB][3], cpos[CB][3];
      int32_t rad[CB];
#pragma unroll
      for (int jc = 0; jc < CB; ++jc) {
          vin[jc][0] = vin[jc][1] = vin[jc][2] = 0.f;
          cin[jc][0] = cin[jc][1] = cin[jc][2] = 1.f;
          cpos[jc][0] = cpos[jc][1] = cpos[jc][2] = 0.f;
          rad[jc] = 1;
          if (c0 + jc < C) {
              const uint64_t e = (uint64_t)(c0 + jc) * N + n;
              const float *vrow = cam_cotangents(src, c0 + jc, N) + 3 * (size_t)n;
              vin[jc][0] = vrow[0]; vin[jc][1] = vrow[1]; vin[jc][2] = vrow[2];
              if (!MASKED && colors != nullptr) {
                  cin[jc][0] = __ldcs(colors + 3 * e); cin[jc][1] = __ldcs(colors + 3 * e + 1);
                  cin[jc][2] = __ldcs(colors + 3 * e + 2);
              }
              if (!MASKED && radii != nullptr) rad[jc] = __ldcs(radii + e);
              if (NB > 1 && kHoistRow) {
                  const float *cp = cam_centre(src, c0 + jc);
                  cpos[jc][0] = cp[0]; cpos[jc][1] = cp[1]; cpos[jc][2] = cp[2];
              }
          }
      }
#pragma unroll
      for (int jc = 0; jc < CB; ++jc) {
        const uint32_t c = c0 + jc;
        if (c >= C) break;
        const uint64_t e = (uint64_t)c * N + n;
        // radii == NULL / colors == NULL: v_colors is PRE-MASKED (zero where the Gaussian is
        // invisible or the colour was clamped) — the layout the camera-parallel exchange gathers
        const float vr = (MASKED || cin[jc][0] > 0.f) ? vin[jc][0] : 0.f, vg = (MASKED || cin[jc][1] > 0.f) ? vin[jc][1] : 0.f,
                    vb = (MASKED || cin[jc][2] > 0.f) ? vin[jc][2] : 0.f;
        const bool visible = (!MASKED && radii != nullptr) ? (rad[jc] > 0) : (vr != 0.f || vg != 0.f || vb != 0.f);
        if (visible) {
            float x = 0.f, y = 0.f, z = 0.f, inorm = 0.f;
            if (NB > 1) {
                if (!kHoistRow) {  // CB cameras in flight: the centre (a few cached words) is read on use
                    const float *cp = cam_centre(src, c);
                    cpos[jc][0] = cp[0]; cpos[jc][1] = cp[1]; cpos[jc][2] = cp[2];
                }
                const float dx = mx - cpos[jc][0], dy = my - cpos[jc][1], dz = mz - cpos[jc][2];
                inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
                x = dx * inorm; y = dy * inorm; z = dz * inorm;
            }
            if (want_means && c >= means_cam_begin && c < means_cam_end) {
                // the row is read with 128-bit loads (12 per Gaussian at K = 16): a scalar load per
                // coefficient touches 32 cache lines per warp instruction and is L1-wavefront bound
                if (per_view || !kHoistRow) load_row<NB>(coeffs + (per_view ? e : (uint64_t)n) * K * 3, cf, cf_vec);
                float vx = 0.f, vy = 0.f, vz = 0.f;
                sh_for_each_basis<true>(deg, x, y, z, [&](int k, float B, float Bx, float By, float Bz) {
                    vc[3 * k] += B * vr; vc[3 * k + 1] += B * vg; vc[3 * k + 2] += B * vb;
                    const float w = cf[3 * k] * vr + cf[3 * k + 1] * vg + cf[3 * k + 2] * vb;
                    vx += Bx * w; vy += By * w; vz += Bz * w;
                });
                const float d = vx * x + vy * y + vz * z;
                vmx += (vx - d * x) * inorm; vmy += (vy - d * y) * inorm; vmz += (vz - d * z) * inorm;
            } else {
                sh_for_each_basis<false>(deg, x, y, z, [&](int k, float B, float, float, float) {
                    vc[3 * k] += B * vr; vc[3 * k + 1] += B * vg; vc[3 * k + 2] += B * vb;
                });
            }
        }
        if (per_view) {
            flush(v_coeffs + e * K * 3);
#pragma unroll
            for (int k = 0; k < NB * 3; k++) vc[k] = 0.f;
        }
      }
    }
    if (!per_view) flush(v_coeffs + (uint64_t)n * K * 3);
    if (v_means != nullptr) { v_means[3 * n] = vmx; v_means[3 * n + 1] = vmy; v_means[3 * n + 2] = vmz; }
}

// Packed (COO) variants of the fused colour stage: one thread per visible (camera, Gaussian)
// row; the row's Gaussian / camera come from gaussian_ids / camera_ids (int64, as the
// reference's packed projection emits them).  This replaces, for packed=True, the chain
// `means[gaussian_ids] - campos[camera_ids]`, `colors[gaussian_ids]` (a [nnz,K,3] gather copy),
// spherical_harmonics, +0.5, clamp (G/rendering.py:370-392) and — in the backward — the
// sort-based index_put of the gathered coefficient gradient.
template <int NB>
__global__ void __launch_bounds__(kThreads)
sh_colors_packed_fwd_kernel(uint32_t nnz, uint32_t N, uint32_t K, uint32_t deg, int per_view,
                            const float *__restrict__ means, const float *__restrict__ campos,
                            const float *__restrict__ sh0, const float *__restrict__ coeffs,
                            const int64_t *__restrict__ camera_ids,
                            const int64_t *__restrict__ gaussian_ids, float *__restrict__ colors) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const uint64_t n = (uint64_t)gaussian_ids[i], c = (uint64_t)camera_ids[i];
    float cf[NB * 3];
    if (sh0 != nullptr) {
        // split table (sh0 [N,1,3], coeffs = shN [N,K-1,3]; gsplat_trainer.py:474 without the cat)
        cf[0] = __ldg(sh0 + 3 * n); cf[1] = __ldg(sh0 + 3 * n + 1); cf[2] = __ldg(sh0 + 3 * n + 2);
        const float *rowN = coeffs + n * (K - 1) * 3;
#pragma unroll
        for (int k = 3; k < NB * 3; k++) cf[k] = __ldg(rowN + k - 3);
    } else {
        const float *row = coeffs + (per_view ? c * N + n : n) * K * 3;
        load_row<NB>(row, cf, ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(coeffs) & 15) == 0));
    }
    float x = 0.f, y = 0.f, z = 0.f;
    if (NB > 1) {
        const float dx = __ldg(means + 3 * n) - campos[3 * c], dy = __ldg(means + 3 * n + 1) - campos[3 * c + 1],
                    dz = __ldg(means + 3 * n + 2) - campos[3 * c + 2];
        const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
        x = dx * inorm; y = dy * inorm; z = dz * inorm;
    }
    float r = 0.f, g = 0.f, b = 0.f;
    sh_for_each_basis<false>(deg, x, y, z, [&](int k, float B, float, float, float) {
        r += B * cf[3 * k]; g += B * cf[3 * k + 1]; b += B * cf[3 * k + 2];
    });
    colors[3 * (size_t)i] = fmaxf(r + 0.5f, 0.f);
    colors[3 * (size_t)i + 1] = fmaxf(g + 0.5f, 0.f);
    colors[3 * (size_t)i + 2] = fmaxf(b + 0.5f, 0.f);
}

// v_coeffs / v_means are ZERO-INITIALISED by the caller.  `unique` != 0: every destination row
// is produced by at most one thread (one camera, or per-view tables) and is stored directly;
// otherwise rows of one Gaussian seen by several cameras are summed with atomics.
template <int NB>
__global__ void __launch_bounds__(kThreads)
sh_colors_packed_bwd_kernel(uint32_t nnz, uint32_t N, uint32_t K, uint32_t deg, int per_view, int unique,
                            int means_unique, const float *__restrict__ means, const float *__restrict__ campos,
                            const float *__restrict__ sh0, const float *__restrict__ coeffs,
                            const int64_t *__restrict__ camera_ids,
                            const int64_t *__restrict__ gaussian_ids, const float *__restrict__ colors,
                            const float *__restrict__ v_colors, float *__restrict__ v_sh0, float *__restrict__ v_coeffs,
                            float *__restrict__ v_means) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const uint64_t n = (uint64_t)gaussian_ids[i], c = (uint64_t)camera_ids[i];
    const uint64_t row_id = per_view ? c * N + n : n;
    const bool vec_ok = ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(v_coeffs) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(coeffs) & 15) == 0);
    const float vr = colors[3 * (size_t)i] > 0.f ? v_colors[3 * (size_t)i] : 0.f;
    const float vg = colors[3 * (size_t)i + 1] > 0.f ? v_colors[3 * (size_t)i + 1] : 0.f;
    const float vb = colors[3 * (size_t)i + 2] > 0.f ? v_colors[3 * (size_t)i + 2] : 0.f;
    float x = 0.f, y = 0.f, z = 0.f, inorm = 0.f;
    if (NB > 1) {
        const float dx = __ldg(means + 3 * n) - campos[3 * c], dy = __ldg(means + 3 * n + 1) - campos[3 * c + 1],
                    dz = __ldg(means + 3 * n + 2) - campos[3 * c + 2];
        inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
        x = dx * inorm; y = dy * inorm; z = dz * inorm;
    }
    float vc[NB * 3];
    if (v_means != nullptr && NB > 1) {
        float cf[NB * 3];
        if (sh0 != nullptr) {
            cf[0] = cf[1] = cf[2] = 0.f;  // the DC basis has no direction derivative
            const float *rowN = coeffs + n * (K - 1) * 3;
#pragma unroll
            for (int k = 3; k < NB * 3; k++) cf[k] = __ldg(rowN + k - 3);
        } else {
            load_row<NB>(coeffs + row_id * K * 3, cf, vec_ok);
        }
        float vx = 0.f, vy = 0.f, vz = 0.f;
        sh_for_each_basis<true>(deg, x, y, z, [&](int k, float B, float Bx, float By, float Bz) {
            vc[3 * k] = B * vr; vc[3 * k + 1] = B * vg; vc[3 * k + 2] = B * vb;
            const float w = cf[3 * k] * vr + cf[3 * k + 1] * vg + cf[3 * k + 2] * vb;
            vx += Bx * w; vy += By * w; vz += Bz * w;
        });
        const float d = vx * x + vy * y + vz * z;
        const float mx = (vx - d * x) * inorm, my = (vy - d * y) * inorm, mz = (vz - d * z) * inorm;
        if (means_unique) {
            v_means[3 * n] = mx; v_means[3 * n + 1] = my; v_means[3 * n + 2] = mz;
        } else {
            atomicAdd(v_means + 3 * n, mx); atomicAdd(v_means + 3 * n + 1, my); atomicAdd(v_means + 3 * n + 2, mz);
        }
    } else {
        sh_for_each_basis<false>(deg, x, y, z, [&](int k, float B, float, float, float) {
            vc[3 * k] = B * vr; vc[3 * k + 1] = B * vg; vc[3 * k + 2] = B * vb;
        });
    }
    if (sh0 != nullptr) {
        float *v0 = v_sh0 + 3 * n, *vN = v_coeffs + n * (K - 1) * 3;
        if (unique) {
            v0[0] = vc[0]; v0[1] = vc[1]; v0[2] = vc[2];
#pragma unroll
            for (int k = 3; k < NB * 3; k++) vN[k - 3] = vc[k];
        } else {
            atomicAdd(v0, vc[0]); atomicAdd(v0 + 1, vc[1]); atomicAdd(v0 + 2, vc[2]);
#pragma unroll
            for (int k = 3; k < NB * 3; k++) atomicAdd(vN + k - 3, vc[k]);
        }
        return;
    }
    float *vrow = v_coeffs + row_id * K * 3;
    if (unique) {
        if ((NB * 3) % 4 == 0 && vec_ok) {
            float4 *o = reinterpret_cast<float4 *>(vrow);
#pragma unroll
            for (int k = 0; k < NB * 3 / 4; k++) o[k] = make_float4(vc[4 * k], vc[4 * k + 1], vc[4 * k + 2], vc[4 * k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < NB * 3; k++) vrow[k] = vc[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < NB * 3; k++) atomicAdd(vrow + k, vc[k]);
    }
}


// ---------------------------------------------------------------------------------------
// Staged ("split-capable") colour stage.  Same maths as sh_colors_fwd/bwd_kernel for ONE shared
// coefficient table, but (1) the table may arrive as the two tensors splat_one optimises,
// sh0 [N,1,3] and shN [N,K-1,3] (R/utils/gsplat_utils/gsplat_trainer.py:474 concatenates them
// into [N,K,3] on every step: a 2 x 12K-byte-per-Gaussian copy forward and its split
// backward), and (2) coefficient rows move through shared memory: a warp owns 32 consecutive
// Gaussians, whose rows are one contiguous, 16-byte aligned span of global memory, so the span
// is loaded / stored with fully coalesced 128-bit accesses and each lane picks its own row out
// of shared memory (row stride chosen odd in units of the access width: conflict-free).
// The thread-per-row kernels above touch 32 different cache lines per load instruction.
//   sh0 != NULL: `rest` = shN, rows of R = 3 (K-1) floats, basis k >= 1 at column 3 (k-1)
//   sh0 == NULL: `rest` = the whole table, rows of R = 3 K floats, basis k at column 3 k
// ---------------------------------------------------------------------------------------
constexpr int kStageWarps = 4;

__host__ __device__ __forceinline__ uint32_t stage_row_stride(uint32_t R) {
    return (R % 4 == 0) ? 4u * ((R / 4) | 1u) : (R | 1u);
}

// ---- 1-D bulk copies (TMA engine, cp.async.bulk / SASS UBLKCP): when the smem rows are contiguous (RS == R:
// the split table, 45 floats per row) a warp's 32 rows are ONE 5760-byte span on both sides, moved by a single
// asynchronous instruction issued by lane 0 — no registers, no per-lane address arithmetic; completion through a
// per-warp mbarrier (loads) or the bulk-group wait (stores).  Spans whose byte count is not a multiple of 16 (the
// last warp of an odd-sized table) take the per-lane loops below.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_bar_init(uint64_t *bar, unsigned lane) {
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
}

__device__ __forceinline__ void bulk_rows_in(float *s_dst, const float *g_src, uint32_t bytes, uint64_t *bar,
                                             uint32_t parity, unsigned lane) {
    const uint32_t bar_a = smem_u32(bar);
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(s_dst)), "l"(g_src), "r"(bytes), "r"(bar_a) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar_a), "r"(parity) : "memory");
}

__device__ __forceinline__ void bulk_rows_out(float *g_dst, const float *s_src, uint32_t bytes, unsigned lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // every lane's generic smem writes -> async proxy
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g_dst), "r"(smem_u32(s_src)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem may be reused / the block may exit
    }
    __syncwarp();
}

// global span [g, g + cnt*R) -> smem rows of stride RS.  bar: this warp's mbarrier (nullptr: no bulk path);
// parity: its phase, flipped here after every bulk load.
__device__ __forceinline__ void stage_rows_in(const float *__restrict__ g, uint32_t cnt, uint32_t R, uint32_t RS,
                                              float *s, unsigned lane, bool aligned, uint64_t *bar = nullptr,
                                              uint32_t *parity = nullptr) {
    const uint32_t total = cnt * R;
    if (bar != nullptr && aligned && RS == R && total > 0 && (total & 3u) == 0) {
        bulk_rows_in(s, g, total * 4u, bar, *parity, lane);
        *parity ^= 1u;
        return;
    }
    if (aligned && R % 4 == 0) {
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        float4 *s4 = reinterpret_cast<float4 *>(s);
        const uint32_t R4 = R / 4, RS4 = RS / 4;
        for (uint32_t i = lane; i < total / 4; i += 32) {
            const uint32_t r = i / R4;
            s4[r * RS4 + (i - r * R4)] = __ldg(g4 + i);
        }
    } else if (aligned && RS == R) {
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        float4 *s4 = reinterpret_cast<float4 *>(s);
        const uint32_t nv = total / 4;
        for (uint32_t i = lane; i < nv; i += 32) s4[i] = __ldg(g4 + i);
        for (uint32_t i = 4 * nv + lane; i < total; i += 32) s[i] = __ldg(g + i);
    } else {
        for (uint32_t i = lane; i < total; i += 32) {
            const uint32_t r = i / R;
            s[r * RS + (i - r * R)] = __ldg(g + i);
        }
    }
}

// smem rows -> global span (streaming stores: the gradient is read once, by the optimizer)
__device__ __forceinline__ void stage_rows_out(float *__restrict__ g, uint32_t cnt, uint32_t R, uint32_t RS,
                                               const float *s, unsigned lane, bool aligned, bool bulk = false) {
    const uint32_t total = cnt * R;
    if (bulk && aligned && RS == R && total > 0 && (total & 3u) == 0) {
        bulk_rows_out(g, s, total * 4u, lane);
        return;
    }
    if (aligned && R % 4 == 0) {
        float4 *g4 = reinterpret_cast<float4 *>(g);
        const float4 *s4 = reinterpret_cast<const float4 *>(s);
        const uint32_t R4 = R / 4, RS4 = RS / 4;
        for (uint32_t i = lane; i < total / 4; i += 32) {
            const uint32_t r = i / R4;
            __stcs(g4 + i, s4[r * RS4 + (i - r * R4)]);
        }
    } else if (aligned && RS == R) {
        float4 *g4 = reinterpret_cast<float4 *>(g);
        const float4 *s4 = reinterpret_cast<const float4 *>(s);
        const uint32_t nv = total / 4;
        for (uint32_t i = lane; i < nv; i += 32) __stcs(g4 + i, s4[i]);
        for (uint32_t i = 4 * nv + lane; i < total; i += 32) g[i] = s[i];
    } else {
        for (uint32_t i = lane; i < total; i += 32) {
            const uint32_t r = i / R;
            g[i] = s[r * RS + (i - r * R)];
        }
    }
}

// this lane's row: NF floats starting at column 0, smem -> registers
template <int NF>
__device__ __forceinline__ void row_to_regs(const float *srow, float (&c)[NF], bool vec) {
    if (vec && NF % 4 == 0) {
        const float4 *p = reinterpret_cast<const float4 *>(srow);
#pragma unroll
        for (int k = 0; k < NF / 4; k++) {
            const float4 q = p[k];
            c[4 * k] = q.x; c[4 * k + 1] = q.y; c[4 * k + 2] = q.z; c[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NF; k++) c[k] = srow[k];
    }
}

// registers -> this lane's row: NF values, then zeros up to R
template <int NF>
__device__ __forceinline__ void regs_to_row(float *srow, const float (&c)[NF], uint32_t R, bool vec) {
    if (vec && NF % 4 == 0) {
        float4 *p = reinterpret_cast<float4 *>(srow);
#pragma unroll
        for (int k = 0; k < NF / 4; k++) p[k] = make_float4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
        for (uint32_t k = NF / 4; k < R / 4; k++) p[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
        for (int k = 0; k < NF; k++) srow[k] = c[k];
        for (uint32_t k = NF; k < R; k++) srow[k] = 0.f;
    }
}

// SPLIT: sh0/shN layout.  NF = floats of a staged row that the active bases use.
template <int NB, bool SPLIT>
__global__ void __launch_bounds__(32 * kStageWarps)
sh_colors_staged_fwd_kernel(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *__restrict__ means,
                            const float *__restrict__ campos, const float *__restrict__ sh0,
                            const float *__restrict__ rest, const int32_t *__restrict__ radii,
                            float *__restrict__ colors) {
    constexpr int K0 = SPLIT ? 1 : 0;
    constexpr int NF = (NB - K0) * 3;
    extern __shared__ float4 stage_smem4[];
    const uint32_t R = (K - K0) * 3, RS = stage_row_stride(R);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *s = reinterpret_cast<float *>(stage_smem4) + (size_t)warp * 32 * RS;
    const uint32_t n0 = (blockIdx.x * kStageWarps + warp) * 32;
    if (n0 >= N) return;  // warp-uniform
    const uint32_t cnt = min(32u, N - n0), n = n0 + lane, c = blockIdx.y;
    const uint64_t e = (uint64_t)c * N + n;
    const bool act = lane < cnt && radii[e] > 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(rest) & 15) == 0;
    __shared__ uint64_t s_bar[kStageWarps];
    uint32_t parity = 0;
    stage_bar_init(&s_bar[warp], lane);
    float cf[NF > 0 ? NF : 1];
    if (NF > 0) {
        if (__any_sync(0xffffffffu, act))
            stage_rows_in(rest + (size_t)n0 * R, cnt, R, RS, s, lane, aligned, &s_bar[warp], &parity);
        __syncwarp();
        if (act) row_to_regs<(NF > 0 ? NF : 1)>(s + (size_t)lane * RS, cf, R % 4 == 0);
    }
    if (lane >= cnt) return;
    float r = 0.f, g = 0.f, b = 0.f;
    if (act) {
        float dc[3] = {0.f, 0.f, 0.f};
        if (SPLIT) { dc[0] = __ldg(sh0 + 3 * (size_t)n); dc[1] = __ldg(sh0 + 3 * (size_t)n + 1); dc[2] = __ldg(sh0 + 3 * (size_t)n + 2); }
        float x = 0.f, y = 0.f, z = 0.f;
        if (NB > 1) {
            const float dx = __ldg(means + 3 * (size_t)n) - campos[3 * c], dy = __ldg(means + 3 * (size_t)n + 1) - campos[3 * c + 1],
                        dz = __ldg(means + 3 * (size_t)n + 2) - campos[3 * c + 2];
            const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
            x = dx * inorm; y = dy * inorm; z = dz * inorm;
        }
        sh_for_each_basis<false>(deg, x, y, z, [&](int k, float B, float, float, float) {
            if (SPLIT && k == 0) { r += B * dc[0]; g += B * dc[1]; b += B * dc[2]; }
            else { r += B * cf[3 * (k - K0)]; g += B * cf[3 * (k - K0) + 1]; b += B * cf[3 * (k - K0) + 2]; }
        });
        r = fmaxf(r + 0.5f, 0.f); g = fmaxf(g + 0.5f, 0.f); b = fmaxf(b + 0.5f, 0.f);
    }
    colors[3 * e] = r; colors[3 * e + 1] = g; colors[3 * e + 2] = b;
}

// One lane per Gaussian, loop over cameras (like sh_colors_bwd_kernel: same conventions for
// radii / colors == NULL and the means camera window); gradient rows leave through smem.
// CFS: the coefficient rows stay in shared memory (a second buffer) and are read from there
// inside the basis loop instead of living in 3 (NB-1) registers: 140 -> ~96 registers.  Odd row
// lengths (the split table: 45 floats) read them conflict-free; rows that are a multiple of 4 floats (the
// whole table: 48) keep the 16-byte-granular stride of the vector staging (52) and pay 4-way conflicts on
// these scalar reads — still far cheaper than the thread-per-row kernel's 32-lines-per-instruction
// global accesses (ncu r2: that kernel is L1 tag-rate bound at 0.105 ms).
template <int NB, bool SPLIT, bool CFS, int CB>
__global__ void __launch_bounds__(32 * kStageWarps, CFS ? 5 : 1)
sh_colors_staged_bwd_kernel(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *__restrict__ means,
                            const CamSource src, const float *__restrict__ sh0,
                            const float *__restrict__ rest, const int32_t *__restrict__ radii,
                            const float *__restrict__ colors,
                            float *__restrict__ v_sh0, float *__restrict__ v_rest, float *__restrict__ v_means,
                            uint32_t means_cam_begin, uint32_t means_cam_end) {
    constexpr int K0 = SPLIT ? 1 : 0;
    constexpr int NF = (NB - K0) * 3;
    constexpr int NFA = NF > 0 ? NF : 1;
    extern __shared__ float4 stage_smem4[];
    const uint32_t R = (K - K0) * 3, RS = stage_row_stride(R);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *s = reinterpret_cast<float *>(stage_smem4) + (size_t)warp * 32 * RS;
    // CFS: second buffer behind the kStageWarps gradient buffers
    const float *scf = reinterpret_cast<float *>(stage_smem4) + (size_t)(kStageWarps + warp) * 32 * RS + (size_t)lane * RS;
    const uint32_t n0 = (blockIdx.x * kStageWarps + warp) * 32;
    if (n0 >= N) return;  // warp-uniform
    const uint32_t cnt = min(32u, N - n0);
    const bool mine = lane < cnt;
    const uint32_t n = mine ? n0 + lane : n0;
    const bool vec = R % 4 == 0;
    const bool want_means = v_means != nullptr && NB > 1 && means_cam_begin < means_cam_end;
    // the DC basis has no direction derivative: sh0 itself is never read here
    __shared__ uint64_t s_bar[kStageWarps];
    uint32_t parity = 0;
    stage_bar_init(&s_bar[warp], lane);
    float cf[CFS ? 1 : NFA];
    if (want_means && NF > 0) {
        if (CFS) {
            stage_rows_in(rest + (size_t)n0 * R, cnt, R, RS, const_cast<float *>(scf) - (size_t)lane * RS, lane,
                          (reinterpret_cast<uintptr_t>(rest) & 15) == 0, &s_bar[warp], &parity);
            __syncwarp();
        } else {
            stage_rows_in(rest + (size_t)n0 * R, cnt, R, RS, s, lane, (reinterpret_cast<uintptr_t>(rest) & 15) == 0,
                          &s_bar[warp], &parity);
            __syncwarp();
            row_to_regs<CFS ? 1 : NFA>(s + (size_t)lane * RS, cf, vec);
            __syncwarp();  // the buffer is reused for the gradient rows below
        }
    }
    float vc[NFA], vdc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < NFA; k++) vc[k] = 0.f;
    float vmx = 0.f, vmy = 0.f, vmz = 0.f;
    const float mx = __ldg(means + 3 * (size_t)n), my = __ldg(means + 3 * (size_t)n + 1), mz = __ldg(means + 3 * (size_t)n + 2);
    constexpr bool MASKED = CB > 1;   // peer exchange: cotangents arrive pre-masked, radii / colors are NULL
    for (uint32_t c0 = 0; c0 < C && mine; c0 += CB) {
      // all loads of the camera batch are in flight before the first is consumed (see sh_colors_bwd_kernel)
      float vin[CB][3], cin[CB][3], cpos[CB][3];
      int32_t rad[CB];
#pragma unroll
      for (int jc = 0; jc < CB; ++jc) {
          vin[jc][0] = vin[jc][1] = vin[jc][2] = 0.f;
          cin[jc][0] = cin[jc][1] = cin[jc][2] = 1.f;
          cpos[jc][0] = cpos[jc][1] = cpos[jc][2] = 0.f;
          rad[jc] = 1;
          if (c0 + jc < C) {
              const uint64_t e = (uint64_t)(c0 + jc) * N + n;
              const float *vrow = cam_cotangents(src, c0 + jc, N) + 3 * (size_t)n;
              vin[jc][0] = vrow[0]; vin[jc][1] = vrow[1]; vin[jc][2] = vrow[2];
              if (!MASKED && colors != nullptr) {
                  cin[jc][0] = __ldcs(colors + 3 * e); cin[jc][1] = __ldcs(colors + 3 * e + 1);
                  cin[jc][2] = __ldcs(colors + 3 * e + 2);
              }
              if (!MASKED && radii != nullptr) rad[jc] = __ldcs(radii + e);
              if (NB > 1 && !MASKED) {
                  const float *cp = cam_centre(src, c0 + jc);
                  cpos[jc][0] = cp[0]; cpos[jc][1] = cp[1]; cpos[jc][2] = cp[2];
              }
          }
      }
#pragma unroll
      for (int jc = 0; jc < CB; ++jc) {
        const uint32_t c = c0 + jc;
        if (c >= C) break;
        const float vr = (MASKED || cin[jc][0] > 0.f) ? vin[jc][0] : 0.f, vg = (MASKED || cin[jc][1] > 0.f) ? vin[jc][1] : 0.f,
                    vb = (MASKED || cin[jc][2] > 0.f) ? vin[jc][2] : 0.f;
        const bool visible = (!MASKED && radii != nullptr) ? (rad[jc] > 0) : (vr != 0.f || vg != 0.f || vb != 0.f);
        if (!visible) continue;
        float x = 0.f, y = 0.f, z = 0.f, inorm = 0.f;
        if (NB > 1) {
            if (MASKED) {  // CB cameras in flight: the centre (a few cached words) is read on use
                const float *cp = cam_centre(src, c);
                cpos[jc][0] = cp[0]; cpos[jc][1] = cp[1]; cpos[jc][2] = cp[2];
            }
            const float dx = mx - cpos[jc][0], dy = my - cpos[jc][1], dz = mz - cpos[jc][2];
            inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
            x = dx * inorm; y = dy * inorm; z = dz * inorm;
        }
        if (want_means && c >= means_cam_begin && c < means_cam_end) {
            float vx = 0.f, vy = 0.f, vz = 0.f;
            sh_for_each_basis<true>(deg, x, y, z, [&](int k, float B, float Bx, float By, float Bz) {
                if (SPLIT && k == 0) { vdc[0] += B * vr; vdc[1] += B * vg; vdc[2] += B * vb; return; }  // dB0 = 0
                const int j = 3 * (k - K0);
                vc[j] += B * vr; vc[j + 1] += B * vg; vc[j + 2] += B * vb;
                const float w = CFS ? scf[j] * vr + scf[j + 1] * vg + scf[j + 2] * vb
                                    : cf[CFS ? 0 : j] * vr + cf[CFS ? 0 : j + 1] * vg + cf[CFS ? 0 : j + 2] * vb;
                vx += Bx * w; vy += By * w; vz += Bz * w;
            });
            const float d = vx * x + vy * y + vz * z;
            vmx += (vx - d * x) * inorm; vmy += (vy - d * y) * inorm; vmz += (vz - d * z) * inorm;
        } else {
            sh_for_each_basis<false>(deg, x, y, z, [&](int k, float B, float, float, float) {
                if (SPLIT && k == 0) { vdc[0] += B * vr; vdc[1] += B * vg; vdc[2] += B * vb; return; }
                const int j = 3 * (k - K0);
                vc[j] += B * vr; vc[j + 1] += B * vg; vc[j + 2] += B * vb;
            });
        }
      }
    }
    if (R > 0) {
        if (mine) {
            if constexpr (NF > 0) regs_to_row<NF>(s + (size_t)lane * RS, vc, R, vec);
            else for (uint32_t k = 0; k < R; k++) s[(size_t)lane * RS + k] = 0.f;
        }
        __syncwarp();
        stage_rows_out(v_rest + (size_t)n0 * R, cnt, R, RS, s, lane, (reinterpret_cast<uintptr_t>(v_rest) & 15) == 0, true);
    }
    if (!mine) return;
    if (SPLIT) { v_sh0[3 * (size_t)n] = vdc[0]; v_sh0[3 * (size_t)n + 1] = vdc[1]; v_sh0[3 * (size_t)n + 2] = vdc[2]; }
    if (v_means != nullptr) { v_means[3 * (size_t)n] = vmx; v_means[3 * (size_t)n + 1] = vmy; v_means[3 * (size_t)n + 2] = vmz; }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b200splat_sh_fwd(uint32_t n_elems, uint32_t n_rows, uint32_t K, uint32_t deg, const float *dirs,
                                const float *coeffs, const uint8_t *masks, float *colors, void *stream) {
    const char *where = "b200splat_sh_fwd";
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    if (n_elems == 0) return 0;
    B2S_REQUIRE(n_rows > 0 && n_elems % n_rows == 0, where, "n_elems must be a multiple of n_coeff_rows");
    const unsigned grid = div_up(n_elems, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
    switch (deg) {
        case 0: sh_fwd_kernel<1><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, colors); break;
        case 1: sh_fwd_kernel<4><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, colors); break;
        case 2: sh_fwd_kernel<9><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, colors); break;
        case 3: sh_fwd_kernel<16><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, colors); break;
        default: sh_fwd_kernel<25><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, colors); break;
    }
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_sh_bwd(uint32_t n_elems, uint32_t n_rows, uint32_t K, uint32_t deg, const float *dirs,
                                const float *coeffs, const uint8_t *masks, const float *v_colors, float *v_coeffs,
                                float *v_dirs, void *stream) {
    const char *where = "b200splat_sh_bwd";
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    if (n_elems == 0) return 0;
    B2S_REQUIRE(n_rows > 0 && n_elems % n_rows == 0, where, "n_elems must be a multiple of n_coeff_rows");
    const unsigned grid = div_up(n_elems, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
    switch (deg) {
        case 0: sh_bwd_kernel<1><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
        case 1: sh_bwd_kernel<4><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
        case 2: sh_bwd_kernel<9><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
        case 3: sh_bwd_kernel<16><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
        default: sh_bwd_kernel<25><<<grid, kThreads, 0, st>>>(n_elems, n_rows, K, deg, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs); break;
    }
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_camera_centers(uint32_t C, const float *viewmats, float *campos, void *stream) {
    if (C == 0) return 0;
    camera_centers_kernel<<<div_up(C, 64), 64, 0, (cudaStream_t)stream>>>(C, viewmats, campos);
    B2S_CHECK_LAUNCH("b200splat_camera_centers");
    return 0;
}

extern "C" int b200splat_sh_colors_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, int per_view,
                                       const float *means, const float *campos, const float *coeffs,
                                       const int32_t *radii, float *colors, void *stream) {
    const char *where = "b200splat_sh_colors_fwd";
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    if ((uint64_t)C * N == 0) return 0;
    const unsigned grid = div_up((uint64_t)C * N, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
#define B2S_SHC(NBV) sh_colors_fwd_kernel<NBV><<<grid, kThreads, 0, st>>>(C, N, K, deg, per_view, means, campos, coeffs, radii, colors)
    switch (deg) {
        case 0: B2S_SHC(1); break;
        case 1: B2S_SHC(4); break;
        case 2: B2S_SHC(9); break;
        case 3: B2S_SHC(16); break;
        default: B2S_SHC(25); break;
    }
#undef B2S_SHC
    B2S_CHECK_LAUNCH(where);
    return 0;
}

static int launch_colors_bwd(const char *where, uint32_t C, uint32_t N, uint32_t K, uint32_t deg, int per_view,
                             const float *means, const CamSource &src, const float *coeffs, const int32_t *radii,
                             const float *colors, float *v_coeffs, float *v_means, uint32_t means_cam_begin,
                             uint32_t means_cam_end, cudaStream_t st) {
    // (Routing a whole [N,K,3] table with 16-byte rows through the staged kernel — coalesced row traffic, coefficient
    // rows in shared memory — measured 0.107 ms against 0.105 ms for this kernel at config B: both sit at the same
    // 16 warps per SM, so the thread-per-row kernel stays the path of the un-split table.)
    const unsigned grid = div_up(N, kThreads);
#define B2S_SHC(NBV)                                                                                                  \
    do {                                                                                                              \
        if (src.bases != nullptr)                                                                                     \
            sh_colors_bwd_kernel<NBV, kPeerCamBatch, true><<<grid, kThreads, 0, st>>>(                                      \
                C, N, K, deg, per_view, means, src, coeffs, radii, colors, v_coeffs, v_means, means_cam_begin, means_cam_end); \
        else                                                                                                          \
            sh_colors_bwd_kernel<NBV, 1, false><<<grid, kThreads, 0, st>>>(                                                  \
                C, N, K, deg, per_view, means, src, coeffs, radii, colors, v_coeffs, v_means, means_cam_begin, means_cam_end); \
    } while (0)
    switch (deg) {
        case 0: B2S_SHC(1); break;
        case 1: B2S_SHC(4); break;
        case 2: B2S_SHC(9); break;
        case 3: B2S_SHC(16); break;
        default: B2S_SHC(25); break;
    }
#undef B2S_SHC
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_sh_colors_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, int per_view,
                                       const float *means, const float *campos, const float *coeffs,
                                       const int32_t *radii, const float *colors, const float *v_colors,
                                       float *v_coeffs, float *v_means, uint32_t means_cam_begin,
                                       uint32_t means_cam_end, void *stream) {
    const char *where = "b200splat_sh_colors_bwd";
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const CamSource src{v_colors, campos, nullptr, 0ull, 1u, 0u};
    return launch_colors_bwd(where, C, N, K, deg, per_view, means, src, coeffs, radii, colors, v_coeffs, v_means,
                             means_cam_begin, means_cam_end, st);
}

// camera-parallel peer exchange: campos / v_colors of camera block b are read from peer_bases[b] + offset_bytes
// (device array of per-rank base addresses, e.g. the buffer_ptrs_dev of a torch symmetric-memory handle)
extern "C" int b200splat_sh_colors_bwd_peer(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *means,
                                            const float *coeffs, const void *peer_bases, uint64_t offset_bytes,
                                            uint32_t cams_per_block, uint32_t hdr_floats, float *v_coeffs,
                                            float *v_means, uint32_t means_cam_begin, uint32_t means_cam_end,
                                            void *stream) {
    const char *where = "b200splat_sh_colors_bwd_peer";
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    B2S_REQUIRE(peer_bases != nullptr && cams_per_block > 0 && hdr_floats >= 3 * cams_per_block, where,
                "peer table, cameras per block and header size are required");
    if (N == 0) return 0;
    const CamSource src{nullptr, nullptr, reinterpret_cast<const unsigned long long *>(peer_bases), offset_bytes,
                        cams_per_block, hdr_floats};
    return launch_colors_bwd(where, C, N, K, deg, 0, means, src, coeffs, nullptr, nullptr, v_coeffs, v_means,
                             means_cam_begin, means_cam_end, (cudaStream_t)stream);
}

static int packed_fwd_impl(const char *where, uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t deg,
                           int per_view, const float *means, const float *campos, const float *sh0,
                           const float *coeffs, const int64_t *camera_ids, const int64_t *gaussian_ids, float *colors,
                           void *stream) {
    (void)C;
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    if (nnz == 0) return 0;
    const unsigned grid = div_up(nnz, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
#define B2S_SHP(NB) sh_colors_packed_fwd_kernel<NB><<<grid, kThreads, 0, st>>>(nnz, N, K, deg, per_view, means, campos, sh0, coeffs, camera_ids, gaussian_ids, colors)
    switch (deg) {
        case 0: B2S_SHP(1); break;
        case 1: B2S_SHP(4); break;
        case 2: B2S_SHP(9); break;
        case 3: B2S_SHP(16); break;
        default: B2S_SHP(25); break;
    }
#undef B2S_SHP
    B2S_CHECK_LAUNCH(where);
    return 0;
}

static int packed_bwd_impl(const char *where, uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t deg,
                           int per_view, const float *means, const float *campos, const float *sh0,
                           const float *coeffs, const int64_t *camera_ids, const int64_t *gaussian_ids,
                           const float *colors, const float *v_colors, float *v_sh0, float *v_coeffs, float *v_means,
                           void *stream) {
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    if (nnz == 0) return 0;
    const unsigned grid = div_up(nnz, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
    const int unique = (per_view || C == 1) ? 1 : 0, means_unique = (C == 1) ? 1 : 0;
#define B2S_SHP(NB) sh_colors_packed_bwd_kernel<NB><<<grid, kThreads, 0, st>>>(nnz, N, K, deg, per_view, unique, means_unique, means, campos, sh0, coeffs, camera_ids, gaussian_ids, colors, v_colors, v_sh0, v_coeffs, v_means)
    switch (deg) {
        case 0: B2S_SHP(1); break;
        case 1: B2S_SHP(4); break;
        case 2: B2S_SHP(9); break;
        case 3: B2S_SHP(16); break;
        default: B2S_SHP(25); break;
    }
#undef B2S_SHP
    B2S_CHECK_LAUNCH(where);
    return 0;
}

// ---- staged colour stage (shared table, optionally split into sh0 / shN) ------------------
extern "C" size_t b200splat_sh_colors_staged_smem_bytes(uint32_t K, int split) {
    const uint32_t R = (K - (split ? 1 : 0)) * 3;
    return (size_t)kStageWarps * 32 * stage_row_stride(R) * sizeof(float);
}

template <bool SPLIT>
static int launch_staged_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *means, const float *campos,
                             const float *sh0, const float *rest, const int32_t *radii, float *colors, size_t smem,
                             cudaStream_t st) {
    const dim3 grid(div_up(N, 32 * kStageWarps), C);
#define B2S_SHS(NBV)                                                                                                 \
    do {                                                                                                             \
        auto kern = sh_colors_staged_fwd_kernel<NBV, SPLIT>;                                                         \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        kern<<<grid, 32 * kStageWarps, smem, st>>>(C, N, K, deg, means, campos, sh0, rest, radii, colors);           \
    } while (0)
    switch (deg) {
        case 0: B2S_SHS(1); break;
        case 1: B2S_SHS(4); break;
        case 2: B2S_SHS(9); break;
        case 3: B2S_SHS(16); break;
        default: B2S_SHS(25); break;
    }
#undef B2S_SHS
    return 0;
}

template <bool SPLIT, int CB>
static int launch_staged_bwd_cb(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *means, const CamSource &src,
                             const float *sh0, const float *rest, const int32_t *radii, const float *colors,
                             float *v_sh0, float *v_rest, float *v_means, uint32_t cb,
                             uint32_t ce, size_t smem, cudaStream_t st) {
    const unsigned grid = div_up(N, 32 * kStageWarps);
    const uint32_t R = (K - (SPLIT ? 1 : 0)) * 3;
    (void)R;
    const bool cfs = 2 * smem <= 100 * 1024;
#define B2S_SHS(NBV)                                                                                                 \
    do {                                                                                                             \
        if (cfs) {                                                                                                   \
            auto kern_c = sh_colors_staged_bwd_kernel<NBV, SPLIT, true, CB>;                                         \
            if (2 * smem > 48 * 1024)                                                                                \
                cudaFuncSetAttribute(kern_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem));         \
            kern_c<<<grid, 32 * kStageWarps, 2 * smem, st>>>(                                                        \
                C, N, K, deg, means, src, sh0, rest, radii, colors, v_sh0, v_rest, v_means, cb, ce);                 \
        } else {                                                                                                     \
            auto kern = sh_colors_staged_bwd_kernel<NBV, SPLIT, false, CB>;                                              \
            if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            kern<<<grid, 32 * kStageWarps, smem, st>>>(C, N, K, deg, means, src, sh0, rest, radii, colors,           \
                                                       v_sh0, v_rest, v_means, cb, ce);                              \
        }                                                                                                            \
    } while (0)
    switch (deg) {
        case 0: B2S_SHS(1); break;
        case 1: B2S_SHS(4); break;
        case 2: B2S_SHS(9); break;
        case 3: B2S_SHS(16); break;
        default: B2S_SHS(25); break;
    }
#undef B2S_SHS
    return 0;
}

template <bool SPLIT>
static int launch_staged_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *means, const CamSource &src,
                             const float *sh0, const float *rest, const int32_t *radii, const float *colors,
                             float *v_sh0, float *v_rest, float *v_means, uint32_t cb,
                             uint32_t ce, size_t smem, cudaStream_t st) {
    if (src.bases != nullptr)
        return launch_staged_bwd_cb<SPLIT, kPeerCamBatch>(C, N, K, deg, means, src, sh0, rest, radii, colors, v_sh0, v_rest,
                                                          v_means, cb, ce, smem, st);
    return launch_staged_bwd_cb<SPLIT, 1>(C, N, K, deg, means, src, sh0, rest, radii, colors, v_sh0, v_rest, v_means, cb,
                                          ce, smem, st);
}

extern "C" int b200splat_sh_colors_staged_fwd(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *means,
                                              const float *campos, const float *sh0, const float *rest,
                                              const int32_t *radii, float *colors, void *stream) {
    const char *where = "b200splat_sh_colors_staged_fwd";
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    if ((uint64_t)C * N == 0) return 0;
    B2S_REQUIRE(C <= 65535, where, "more than 65535 cameras");
    const size_t smem = b200splat_sh_colors_staged_smem_bytes(K, sh0 != nullptr);
    B2S_REQUIRE(smem <= 200 * 1024, where, "K too large for the staged colour kernels");
    cudaStream_t st = (cudaStream_t)stream;
    if (sh0 != nullptr) launch_staged_fwd<true>(C, N, K, deg, means, campos, sh0, rest, radii, colors, smem, st);
    else launch_staged_fwd<false>(C, N, K, deg, means, campos, sh0, rest, radii, colors, smem, st);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_sh_colors_staged_bwd(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *means,
                                              const float *campos, const float *sh0, const float *rest,
                                              const int32_t *radii, const float *colors, const float *v_colors,
                                              float *v_sh0, float *v_rest, float *v_means, uint32_t means_cam_begin,
                                              uint32_t means_cam_end, void *stream) {
    const char *where = "b200splat_sh_colors_staged_bwd";
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    if (N == 0) return 0;
    B2S_REQUIRE((sh0 != nullptr) == (v_sh0 != nullptr), where, "sh0 and v_sh0 go together");
    const size_t smem = b200splat_sh_colors_staged_smem_bytes(K, sh0 != nullptr);
    B2S_REQUIRE(smem <= 200 * 1024, where, "K too large for the staged colour kernels");
    cudaStream_t st = (cudaStream_t)stream;
    const CamSource src{v_colors, campos, nullptr, 0ull, 1u, 0u};
    if (sh0 != nullptr)
        launch_staged_bwd<true>(C, N, K, deg, means, src, sh0, rest, radii, colors, v_sh0, v_rest, v_means,
                                means_cam_begin, means_cam_end, smem, st);
    else
        launch_staged_bwd<false>(C, N, K, deg, means, src, sh0, rest, radii, colors, v_sh0, v_rest,
                                 v_means, means_cam_begin, means_cam_end, smem, st);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_sh_colors_staged_bwd_peer(uint32_t C, uint32_t N, uint32_t K, uint32_t deg, const float *means,
                                                   const float *sh0, const float *rest, const void *peer_bases,
                                                   uint64_t offset_bytes, uint32_t cams_per_block, uint32_t hdr_floats,
                                                   float *v_sh0, float *v_rest, float *v_means,
                                                   uint32_t means_cam_begin, uint32_t means_cam_end, void *stream) {
    const char *where = "b200splat_sh_colors_staged_bwd_peer";
    B2S_REQUIRE(deg <= 4, where, "degrees_to_use must be <= 4");
    B2S_REQUIRE((deg + 1) * (deg + 1) <= K, where, "K too small for degrees_to_use");
    B2S_REQUIRE(peer_bases != nullptr && cams_per_block > 0 && hdr_floats >= 3 * cams_per_block, where,
                "peer table, cameras per block and header size are required");
    if (N == 0) return 0;
    B2S_REQUIRE((sh0 != nullptr) == (v_sh0 != nullptr), where, "sh0 and v_sh0 go together");
    const size_t smem = b200splat_sh_colors_staged_smem_bytes(K, sh0 != nullptr);
    B2S_REQUIRE(smem <= 200 * 1024, where, "K too large for the staged colour kernels");
    cudaStream_t st = (cudaStream_t)stream;
    const CamSource src{nullptr, nullptr, reinterpret_cast<const unsigned long long *>(peer_bases), offset_bytes,
                        cams_per_block, hdr_floats};
    if (sh0 != nullptr)
        launch_staged_bwd<true>(C, N, K, deg, means, src, sh0, rest, nullptr, nullptr, v_sh0, v_rest, v_means,
                                means_cam_begin, means_cam_end, smem, st);
    else
        launch_staged_bwd<false>(C, N, K, deg, means, src, sh0, rest, nullptr, nullptr, v_sh0, v_rest, v_means,
                                 means_cam_begin, means_cam_end, smem, st);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_sh_colors_packed_fwd(uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t deg,
                                              int per_view, const float *means, const float *campos,
                                              const float *coeffs, const int64_t *camera_ids,
                                              const int64_t *gaussian_ids, float *colors, void *stream) {
    return packed_fwd_impl("b200splat_sh_colors_packed_fwd", nnz, C, N, K, deg, per_view, means, campos, nullptr, coeffs,
                           camera_ids, gaussian_ids, colors, stream);
}

extern "C" int b200splat_sh_colors_packed_bwd(uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t deg,
                                              int per_view, const float *means, const float *campos,
                                              const float *coeffs, const int64_t *camera_ids,
                                              const int64_t *gaussian_ids, const float *colors,
                                              const float *v_colors, float *v_coeffs, float *v_means, void *stream) {
    return packed_bwd_impl("b200splat_sh_colors_packed_bwd", nnz, C, N, K, deg, per_view, means, campos, nullptr, coeffs,
                           camera_ids, gaussian_ids, colors, v_colors, nullptr, v_coeffs, v_means, stream);
}

// split table: sh0 [N,1,3] + shN [N,K-1,3] (shared, never per view); v_sh0 / v_shN must be zero-filled by
// the caller when a Gaussian can be invisible or seen by several cameras (rows are accumulated)
extern "C" int b200splat_sh_colors_packed_split_fwd(uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t deg,
                                                    const float *means, const float *campos, const float *sh0,
                                                    const float *shN, const int64_t *camera_ids,
                                                    const int64_t *gaussian_ids, float *colors, void *stream) {
    const char *where = "b200splat_sh_colors_packed_split_fwd";
    B2S_REQUIRE(sh0 != nullptr && K >= 1, where, "sh0 is required");
    return packed_fwd_impl(where, nnz, C, N, K, deg, 0, means, campos, sh0, shN, camera_ids, gaussian_ids, colors,
                           stream);
}

extern "C" int b200splat_sh_colors_packed_split_bwd(uint32_t nnz, uint32_t C, uint32_t N, uint32_t K, uint32_t deg,
                                                    const float *means, const float *campos, const float *sh0,
                                                    const float *shN, const int64_t *camera_ids,
                                                    const int64_t *gaussian_ids, const float *colors,
                                                    const float *v_colors, float *v_sh0, float *v_shN, float *v_means,
                                                    void *stream) {
    const char *where = "b200splat_sh_colors_packed_split_bwd";
    B2S_REQUIRE(sh0 != nullptr && v_sh0 != nullptr, where, "sh0 and v_sh0 are required");
    return packed_bwd_impl(where, nnz, C, N, K, deg, 0, means, campos, sh0, shN, camera_ids, gaussian_ids, colors,
                           v_colors, v_sh0, v_shN, v_means, stream);
}
