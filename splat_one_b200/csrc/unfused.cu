// unfused.cu — the un-fused exported operators of the projection chain:
//   quat_scale_to_covar_preci  CS/quat_scale_to_covar_preci_{fwd,bwd}.cu, math CS/utils.cuh:15-180
//   world_to_cam               CS/world_to_cam_{fwd,bwd}.cu, math CS/utils.cuh:598-658
//   proj                       CS/proj_{fwd,bwd}.cu, camera models CS/utils.cuh:183-594
// All are streaming, HBM-bound maps; one thread per Gaussian (or per pair), 128-bit loads of
// the quaternions, outputs written exactly once.  The backward of world_to_cam maps a thread
// to a GAUSSIAN and loops over cameras (as csrc/projection.cu does), so v_means / v_covars
// need neither atomics nor a zero fill (reference: warp partition by gid + atomics,
// CS/world_to_cam_bwd.cu:82-108); only the [C,4,4] pose gradient is reduced with atomics.
#include "proj_math.cuh"

namespace b2s {

__device__ __forceinline__ V4 load_quat(const float *__restrict__ quats, uint32_t i) {
    const float4 q = __ldg(reinterpret_cast<const float4 *>(quats) + i);
    return {q.x, q.y, q.z, q.w};  // memory order w, x, y, z
}

__device__ __forceinline__ void store_sym(float *__restrict__ out, uint32_t i, const M3 &m, bool triu) {
    if (triu) {
        float *o = out + 6 * (size_t)i;
        o[0] = m.m[0][0]; o[1] = m.m[0][1]; o[2] = m.m[0][2];
        o[3] = m.m[1][1]; o[4] = m.m[1][2]; o[5] = m.m[2][2];
    } else {
        float *o = out + 9 * (size_t)i;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) o[3 * r + c] = m.m[r][c];
    }
}

// cotangent of a [N,6] triu / [N,3,3] full matrix as a full 3x3 (off-diagonals of the triu
// form are halved, CS/quat_scale_to_covar_preci_bwd.cu:57-68)
__device__ __forceinline__ M3 load_cotangent(const float *__restrict__ v, uint32_t i, bool triu) {
    M3 g;
    if (triu) {
        const float *p = v + 6 * (size_t)i;
        g.m[0][0] = p[0]; g.m[0][1] = g.m[1][0] = p[1] * .5f; g.m[0][2] = g.m[2][0] = p[2] * .5f;
        g.m[1][1] = p[3]; g.m[1][2] = g.m[2][1] = p[4] * .5f; g.m[2][2] = p[5];
    } else {
        const float *p = v + 9 * (size_t)i;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) g.m[r][c] = p[3 * r + c];
    }
    return g;
}

static __global__ void __launch_bounds__(kThreads)
qs2cp_fwd_kernel(uint32_t N, const float *__restrict__ quats, const float *__restrict__ scales, bool triu,
                 float *__restrict__ covars, float *__restrict__ precis) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const V4 q = load_quat(quats, i);
    const V3 s = {scales[3 * (size_t)i], scales[3 * (size_t)i + 1], scales[3 * (size_t)i + 2]};
    if (covars != nullptr) store_sym(covars, i, quat_scale_to_covar(q, s), triu);
    if (precis != nullptr) {
        // P = (R S^-1)(R S^-1)^T, CS/utils.cuh:80-96
        const V3 is = {1.0f / s.x, 1.0f / s.y, 1.0f / s.z};
        store_sym(precis, i, quat_scale_to_covar(q, is), triu);
    }
}

static __global__ void __launch_bounds__(kThreads)
qs2cp_bwd_kernel(uint32_t N, const float *__restrict__ quats, const float *__restrict__ scales,
                 const float *__restrict__ v_covars, const float *__restrict__ v_precis, bool triu,
                 float *__restrict__ v_quats, float *__restrict__ v_scales) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const V4 q = load_quat(quats, i);
    const V3 s = {scales[3 * (size_t)i], scales[3 * (size_t)i + 1], scales[3 * (size_t)i + 2]};
    V4 v_q = {0.f, 0.f, 0.f, 0.f};
    V3 v_s = {0.f, 0.f, 0.f};
    if (v_covars != nullptr) quat_scale_to_covar_vjp(q, s, load_cotangent(v_covars, i, triu), v_q, v_s);
    if (v_precis != nullptr) {
        // same VJP at S^-1, then d(1/s)/ds = -1/s² (CS/utils.cuh:139-180)
        const V3 is = {1.0f / s.x, 1.0f / s.y, 1.0f / s.z};
        V3 v_is = {0.f, 0.f, 0.f};
        quat_scale_to_covar_vjp(q, is, load_cotangent(v_precis, i, triu), v_q, v_is);
        v_s.x += -is.x * is.x * v_is.x;
        v_s.y += -is.y * is.y * v_is.y;
        v_s.z += -is.z * is.z * v_is.z;
    }
    reinterpret_cast<float4 *>(v_quats)[i] = make_float4(v_q.w, v_q.x, v_q.y, v_q.z);
    v_scales[3 * (size_t)i] = v_s.x; v_scales[3 * (size_t)i + 1] = v_s.y; v_scales[3 * (size_t)i + 2] = v_s.z;
}

__device__ __forceinline__ M3 load_m3(const float *__restrict__ p) {
    M3 m;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) m.m[r][c] = p[3 * r + c];
    return m;
}

__device__ __forceinline__ void load_pose(const float *__restrict__ vm, M3 &R, V3 &t) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) R.m[r][c] = vm[4 * r + c];
    t = {vm[3], vm[7], vm[11]};
}

static __global__ void __launch_bounds__(kThreads)
world_to_cam_fwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                        const float *__restrict__ viewmats, float *__restrict__ means_c,
                        float *__restrict__ covars_c) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= N) return;
    const V3 p = {means[3 * (size_t)gid], means[3 * (size_t)gid + 1], means[3 * (size_t)gid + 2]};
    const M3 S = load_m3(covars + 9 * (size_t)gid);
    for (uint32_t cid = 0; cid < C; ++cid) {  // parameters are read once for all cameras
        M3 R; V3 t;
        load_pose(viewmats + 16 * cid, R, t);
        const size_t idx = (size_t)cid * N + gid;
        V3 pc = m3_mulv(R, p);
        means_c[3 * idx] = pc.x + t.x; means_c[3 * idx + 1] = pc.y + t.y; means_c[3 * idx + 2] = pc.z + t.z;
        const M3 Sc = m3_mul_bt(m3_mul(R, S), R);  // R S R^T, CS/utils.cuh:629-636
        float *o = covars_c + 9 * idx;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) o[3 * r + c] = Sc.m[r][c];
    }
}

static __global__ void __launch_bounds__(kThreads)
world_to_cam_bwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                        const float *__restrict__ viewmats, const float *__restrict__ v_means_c,
                        const float *__restrict__ v_covars_c, float *__restrict__ v_means,
                        float *__restrict__ v_covars, float *__restrict__ v_viewmats) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = gid < N;
    V3 p = {0.f, 0.f, 0.f};
    M3 S = m3_zero();
    if (in_range) {
        p = {means[3 * (size_t)gid], means[3 * (size_t)gid + 1], means[3 * (size_t)gid + 2]};
        S = load_m3(covars + 9 * (size_t)gid);
    }
    V3 v_p = {0.f, 0.f, 0.f};
    M3 v_S = m3_zero();
    for (uint32_t cid = 0; cid < C; ++cid) {
        M3 R; V3 t;
        load_pose(viewmats + 16 * cid, R, t);
        const size_t idx = (size_t)cid * N + gid;
        M3 v_R = m3_zero();
        V3 v_t = {0.f, 0.f, 0.f};
        if (in_range) {
            if (v_means_c != nullptr) {  // CS/utils.cuh:608-627
                const V3 g = {v_means_c[3 * idx], v_means_c[3 * idx + 1], v_means_c[3 * idx + 2]};
                const V3 vm = m3_tmulv(R, g);
                v_p.x += vm.x; v_p.y += vm.y; v_p.z += vm.z;
                const float gv[3] = {g.x, g.y, g.z}, pw[3] = {p.x, p.y, p.z};
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) v_R.m[r][c] += gv[r] * pw[c];
                v_t = g;
            }
            if (v_covars_c != nullptr) {  // CS/utils.cuh:638-658
                const M3 G = load_m3(v_covars_c + 9 * idx);
                v_S = m3_add(v_S, m3_mul(m3_mul_at(R, G), R));                        // R^T G R
                v_R = m3_add(v_R, m3_add(m3_mul_bt(m3_mul(G, R), S),                  // G R S^T
                                         m3_mul(m3_mul(m3_transpose(G), R), S)));     // G^T R S
            }
        }
        if (v_viewmats != nullptr) {  // warp-uniform
            float vals[12];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) vals[4 * r + c] = v_R.m[r][c];
            vals[3] = v_t.x; vals[7] = v_t.y; vals[11] = v_t.z;
#pragma unroll
            for (int k = 0; k < 12; k++) vals[k] = warp_sum(vals[k]);
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 12; k++) atomicAdd(v_viewmats + 16 * cid + k, vals[k]);
            }
        }
    }
    if (!in_range) return;
    if (v_means != nullptr) {
        v_means[3 * (size_t)gid] = v_p.x; v_means[3 * (size_t)gid + 1] = v_p.y; v_means[3 * (size_t)gid + 2] = v_p.z;
    }
    if (v_covars != nullptr) {
        float *o = v_covars + 9 * (size_t)gid;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) o[3 * r + c] = v_S.m[r][c];
    }
}

__device__ __forceinline__ Cam intrinsics_only(const float *__restrict__ K) {
    Cam c;
    c.R = m3_zero();
    c.t = {0.f, 0.f, 0.f};
    c.fx = K[0]; c.cx = K[2]; c.fy = K[4]; c.cy = K[5];
    return c;
}

// The reference reads the row-major 3x3 with glm::make_mat3 (CS/proj_fwd.cu:54), i.e. as its
// transpose; identical for the symmetric matrices the API is specified for, kept for exactness.
__device__ __forceinline__ M3 load_m3_t(const float *__restrict__ p) {
    M3 m;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) m.m[r][c] = p[3 * c + r];
    return m;
}

static __global__ void __launch_bounds__(kThreads)
proj_fwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                const float *__restrict__ Ks, uint32_t W, uint32_t H, int camera_model, float *__restrict__ means2d,
                float *__restrict__ covars2d) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)C * N) return;
    const uint32_t cid = (uint32_t)(idx / N);
    const Cam cam = intrinsics_only(Ks + 9 * cid);
    const V3 p = {__ldcs(means + 3 * idx), __ldcs(means + 3 * idx + 1), __ldcs(means + 3 * idx + 2)};
    const M3 S = load_m3_t(covars + 9 * idx);
    M2 cov2d = {0.f, 0.f, 0.f, 0.f};
    V2 m2 = {0.f, 0.f};
    M23 J;
    switch (camera_model) {
        case B200SPLAT_PINHOLE: persp_proj(p, S, cam, W, H, cov2d, m2, J); break;
        case B200SPLAT_ORTHO: ortho_proj(p, S, cam, cov2d, m2); break;
        case B200SPLAT_FISHEYE: fisheye_proj(p, S, cam, cov2d, m2); break;
        default: spherical_proj(p, S, W, H, cov2d, m2); break;
    }
    reinterpret_cast<float2 *>(means2d)[idx] = make_float2(m2.x, m2.y);
    reinterpret_cast<float4 *>(covars2d)[idx] = make_float4(cov2d.a00, cov2d.a01, cov2d.a10, cov2d.a11);
}

static __global__ void __launch_bounds__(kThreads)
proj_bwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                const float *__restrict__ Ks, uint32_t W, uint32_t H, int camera_model,
                const float *__restrict__ v_means2d, const float *__restrict__ v_covars2d,
                float *__restrict__ v_means, float *__restrict__ v_covars) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)C * N) return;
    const uint32_t cid = (uint32_t)(idx / N);
    const Cam cam = intrinsics_only(Ks + 9 * cid);
    const V3 p = {__ldcs(means + 3 * idx), __ldcs(means + 3 * idx + 1), __ldcs(means + 3 * idx + 2)};
    const M3 S = load_m3_t(covars + 9 * idx);
    const float2 g2 = __ldcs(reinterpret_cast<const float2 *>(v_means2d) + idx);
    const float4 g4 = __ldcs(reinterpret_cast<const float4 *>(v_covars2d) + idx);
    const V2 v_m2 = {g2.x, g2.y};
    const M2 G = {g4.x, g4.y, g4.z, g4.w};  // row-major cotangent (CS/proj_bwd.cu:61, 76)
    V3 v_p = {0.f, 0.f, 0.f};
    M3 v_S = m3_zero();
    switch (camera_model) {
        case B200SPLAT_PINHOLE: persp_proj_vjp(p, S, cam, W, H, G, v_m2, v_p, v_S); break;
        case B200SPLAT_ORTHO: ortho_proj_vjp(cam, G, v_m2, v_p, v_S); break;
        case B200SPLAT_FISHEYE: fisheye_proj_vjp(p, S, cam, G, v_m2, v_p, v_S); break;
        default: spherical_proj_vjp(p, W, H, G, v_m2, v_p, v_S); break;
    }
    v_means[3 * idx] = v_p.x; v_means[3 * idx + 1] = v_p.y; v_means[3 * idx + 2] = v_p.z;
    float *o = v_covars + 9 * idx;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) o[3 * r + c] = v_S.m[r][c];
}

}  // namespace b2s

using namespace b2s;

extern "C" int b200splat_quat_scale_to_covar_preci_fwd(uint32_t N, const float *quats, const float *scales, int triu,
                                                       float *covars, float *precis, void *stream) {
    const char *where = "b200splat_quat_scale_to_covar_preci_fwd";
    if (N == 0 || (covars == nullptr && precis == nullptr)) return 0;
    B2S_REQUIRE(((uintptr_t)quats & 15) == 0, where, "quats must be 16-byte aligned");
    qs2cp_fwd_kernel<<<div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(N, quats, scales, triu != 0, covars,
                                                                                 precis);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_quat_scale_to_covar_preci_bwd(uint32_t N, const float *quats, const float *scales,
                                                       const float *v_covars, const float *v_precis, int triu,
                                                       float *v_quats, float *v_scales, void *stream) {
    const char *where = "b200splat_quat_scale_to_covar_preci_bwd";
    if (N == 0) return 0;
    B2S_REQUIRE(((uintptr_t)quats & 15) == 0 && ((uintptr_t)v_quats & 15) == 0, where,
                "quats / v_quats must be 16-byte aligned");
    qs2cp_bwd_kernel<<<div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(N, quats, scales, v_covars, v_precis,
                                                                                 triu != 0, v_quats, v_scales);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_world_to_cam_fwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                                          const float *viewmats, float *means_c, float *covars_c, void *stream) {
    const char *where = "b200splat_world_to_cam_fwd";
    if (C == 0 || N == 0) return 0;
    world_to_cam_fwd_kernel<<<div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(C, N, means, covars, viewmats,
                                                                                        means_c, covars_c);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_world_to_cam_bwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                                          const float *viewmats, const float *v_means_c, const float *v_covars_c,
                                          float *v_means, float *v_covars, float *v_viewmats, void *stream) {
    const char *where = "b200splat_world_to_cam_bwd";
    if (C == 0 || N == 0) return 0;
    world_to_cam_bwd_kernel<<<div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, viewmats, v_means_c, v_covars_c, v_means, v_covars, v_viewmats);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_proj_fwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *Ks,
                                  uint32_t W, uint32_t H, int camera_model, float *means2d, float *covars2d,
                                  void *stream) {
    const char *where = "b200splat_proj_fwd";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    B2S_REQUIRE(camera_model >= 0 && camera_model <= 3, where, "unknown camera model");
    if ((uint64_t)C * N == 0) return 0;
    B2S_REQUIRE(((uintptr_t)covars2d & 15) == 0 && ((uintptr_t)means2d & 7) == 0, where,
                "means2d / covars2d must be 8 / 16-byte aligned");
    proj_fwd_kernel<<<div_up((uint64_t)C * N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, Ks, W, H, camera_model, means2d, covars2d);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_proj_bwd(uint32_t C, uint32_t N, const float *means, const float *covars, const float *Ks,
                                  uint32_t W, uint32_t H, int camera_model, const float *v_means2d,
                                  const float *v_covars2d, float *v_means, float *v_covars, void *stream) {
    const char *where = "b200splat_proj_bwd";
    B2S_REQUIRE(camera_model >= 0 && camera_model <= 3, where, "unknown camera model");
    if ((uint64_t)C * N == 0) return 0;
    B2S_REQUIRE(((uintptr_t)v_covars2d & 15) == 0 && ((uintptr_t)v_means2d & 7) == 0, where,
                "v_means2d / v_covars2d must be 8 / 16-byte aligned");
    proj_bwd_kernel<<<div_up((uint64_t)C * N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, Ks, W, H, camera_model, v_means2d, v_covars2d, v_means, v_covars);
    B2S_CHECK_LAUNCH(where);
    return 0;
}
