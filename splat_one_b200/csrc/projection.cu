// projection.cu — fused 3D→2D projection (a2), its backward (a3) and the packed
// (COO) variants (a4).  Replaces CS/fully_fused_projection_{fwd,bwd}.cu and
// CS/fully_fused_projection_packed_{fwd,bwd}.cu of the reference.
//
// Design (B200): all four kernels are single-pass streaming kernels bound by HBM
// (40 B of parameters in, 28 B out per visible pair).  Parameters are read through the
// read-only path; per-camera matrices are hoisted into registers.  The backward of the
// unpacked layout maps one thread to one *Gaussian* and loops over the C cameras, so
// the parameter gradients are written once, without atomics and without a zero-fill
// pass (the reference uses a labelled warp partition + atomics per (camera, Gaussian)
// pair, CS/fully_fused_projection_bwd.cu:209-254).
#include "proj_math.cuh"
#include "scan.cuh"

namespace b2s {

__device__ __forceinline__ V3 load_v3(const float *__restrict__ p, uint32_t i) {
    return {__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)};
}

__device__ __forceinline__ M3 load_covar(const float *__restrict__ covars, const float *__restrict__ quats,
                                         const float *__restrict__ scales, uint32_t gid, V4 &q, V3 &s) {
    if (covars != nullptr) {
        float c[6];
#pragma unroll
        for (int k = 0; k < 6; k++) c[k] = __ldg(covars + 6 * gid + k);
        return covar_from_triu(c);
    }
    const float4 qq = __ldg(reinterpret_cast<const float4 *>(quats) + gid);
    q = {qq.x, qq.y, qq.z, qq.w};
    s = load_v3(scales, gid);
    return quat_scale_to_covar(q, s);
}

// ---------------------------------------------------------------------------------------
// a2: one thread per (camera, Gaussian) pair.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
projection_fwd_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                      const float *__restrict__ quats, const float *__restrict__ scales,
                      const float *__restrict__ viewmats, const float *__restrict__ Ks, uint32_t W, uint32_t H,
                      float eps2d, float near_plane, float far_plane, float radius_clip, int camera_model,
                      int32_t *__restrict__ radii, float *__restrict__ means2d, float *__restrict__ depths,
                      float *__restrict__ conics, float *__restrict__ compensations) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)C * N) return;
    const uint32_t cid = idx / N, gid = idx % N;
    const Cam cam = load_cam(viewmats + 16 * cid, Ks + 9 * cid);
    const V3 mean = load_v3(means, gid);
    V4 q; V3 s;
    const M3 covar = load_covar(covars, quats, scales, gid, q, s);
    ProjOut o;
    if (!project_one(mean, covar, cam, W, H, eps2d, near_plane, far_plane, radius_clip, camera_model, false, o)) {
        // culled rows are written as zeros HERE (the reference leaves them uninitialised): the caller needs no
        // zero-fill pass over the outputs
        radii[idx] = 0;
        reinterpret_cast<float2 *>(means2d)[idx] = make_float2(0.f, 0.f);
        depths[idx] = 0.f;
        conics[3 * idx] = 0.f; conics[3 * idx + 1] = 0.f; conics[3 * idx + 2] = 0.f;
        if (compensations != nullptr) compensations[idx] = 0.f;
        return;
    }
    radii[idx] = (int32_t)o.radius;
    reinterpret_cast<float2 *>(means2d)[idx] = make_float2(o.mean2d.x, o.mean2d.y);
    depths[idx] = o.depth_norm;  // |mean_c| for every camera model (fork behaviour)
    conics[3 * idx] = o.conic[0];
    conics[3 * idx + 1] = o.conic[1];
    conics[3 * idx + 2] = o.conic[2];
    if (compensations != nullptr) compensations[idx] = o.comp;
}

// Running densification statistics of the reference's DefaultStrategy (G/strategy/default.py:239-262,
// `_update_state`), folded into the projection backward: the kernel already visits
// every (camera, Gaussian) pair with its 2-D mean cotangent and radius in registers, so
//   grad2d[n] += sum_c [radius > 0] |(v_x sx, v_y sy)|,   count[n] += sum_c [radius > 0],
//   radii[n]   = max(radii[n], max_c radius / max(W, H))
// cost three read-modify-writes per Gaussian instead of ~10 ATen launches over [C,N] temporaries.
struct DensifyState {
    float *grad2d;     // nullptr: statistics off
    float *count;
    float *radii;      // nullable (refine_scale2d_stop_iter == 0)
    float sx, sy;      // W/2 * n_cameras, H/2 * n_cameras
    float max_wh;      // max(W, H): radii are divided exactly (IEEE), like the reference
};

// ---------------------------------------------------------------------------------------
// a3: one thread per Gaussian, loop over cameras; no atomics on the parameter grads.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)   // (kThreads, 3) = 80 registers spills: 0.059 vs 0.055 ms
projection_bwd_kernel(const DensifyState dens, uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                      const float *__restrict__ quats, const float *__restrict__ scales,
                      const float *__restrict__ viewmats, const float *__restrict__ Ks, uint32_t W, uint32_t H,
                      float eps2d, int camera_model, const int32_t *__restrict__ radii,
                      const float *__restrict__ conics, const float *__restrict__ compensations,
                      const float *__restrict__ v_means2d, const float *__restrict__ v_depths,
                      const float *__restrict__ v_conics, const float *__restrict__ v_compensations,
                      float *__restrict__ v_means, float *__restrict__ v_covars, float *__restrict__ v_quats,
                      float *__restrict__ v_scales, float *__restrict__ v_viewmats) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = gid < N;
    // Per-camera inputs of one (camera, Gaussian) pair.  They are loaded UNCONDITIONALLY and one camera
    // ahead — before the radius is tested and before the covariance is rebuilt — so that a thread makes one
    // DRAM round trip per camera instead of four dependent ones (ncu r2: 50 % long-scoreboard stalls at 16
    // warps per SM; 0.055 -> see DESIGN.md 3.1).  Rows of culled Gaussians are zero-filled, never garbage.
    struct PairIn {
        int32_t radius;
        float conic[3], vconic[3], comp, v_comp, v_depth;
        float2 vm2;
    };
    const bool has_comp = v_compensations != nullptr;
    auto load_pair = [&](uint32_t cid) {
        PairIn p;
        const uint64_t idx = (uint64_t)cid * N + gid;
        p.radius = __ldcs(radii + idx);
        p.conic[0] = __ldcs(conics + 3 * idx); p.conic[1] = __ldcs(conics + 3 * idx + 1); p.conic[2] = __ldcs(conics + 3 * idx + 2);
        p.vconic[0] = __ldcs(v_conics + 3 * idx); p.vconic[1] = __ldcs(v_conics + 3 * idx + 1);
        p.vconic[2] = __ldcs(v_conics + 3 * idx + 2);
        p.vm2 = __ldcs(reinterpret_cast<const float2 *>(v_means2d) + idx);
        p.v_depth = __ldcs(v_depths + idx);
        p.comp = has_comp ? compensations[idx] : 0.f;
        p.v_comp = has_comp ? v_compensations[idx] : 0.f;
        return p;
    };
    PairIn nxt = {};
    if (in_range && C > 0) nxt = load_pair(0);
    V3 mean = {0.f, 0.f, 0.f};
    V4 q = {1.f, 0.f, 0.f, 0.f};
    V3 s = {1.f, 1.f, 1.f};
    M3 covar = m3_zero();
    if (in_range) {
        mean = load_v3(means, gid);
        covar = load_covar(covars, quats, scales, gid, q, s);
    }
    V3 v_mean = {0.f, 0.f, 0.f};
    M3 v_covar = m3_zero();
    float d_grad = 0.f, d_count = 0.f, d_radius = 0.f;
    for (uint32_t cid = 0; cid < C; ++cid) {
        const PairIn cur = nxt;
        if (in_range && cid + 1 < C) nxt = load_pair(cid + 1);
        const int32_t radius = in_range ? cur.radius : 0;
        const bool valid = radius > 0;
        M3 v_R = m3_zero();
        V3 v_t = {0.f, 0.f, 0.f};
        if (valid) {
            const Cam cam = load_cam(viewmats + 16 * cid, Ks + 9 * cid);
            const V2 v_mean2d = {cur.vm2.x, cur.vm2.y};
            if (dens.grad2d != nullptr) {
                const float gx = cur.vm2.x * dens.sx, gy = cur.vm2.y * dens.sy;
                d_grad += sqrtf(gx * gx + gy * gy);
                d_count += 1.f;
                d_radius = fmaxf(d_radius, __fdiv_rn((float)radius, dens.max_wh));
            }
            float comp = cur.comp, v_comp = cur.v_comp;
            project_one_vjp(mean, covar, cam, W, H, eps2d, camera_model, cur.conic, has_comp ? &comp : nullptr,
                            has_comp ? &v_comp : nullptr, v_mean2d, cur.v_depth, cur.vconic, v_mean, v_covar,
                            v_viewmats ? &v_R : nullptr, v_viewmats ? &v_t : nullptr);
        }
        if (v_viewmats != nullptr) {  // warp-uniform branch
            float vals[12];
#pragma unroll
            for (int i = 0; i < 3; i++) {
#pragma unroll
                for (int j = 0; j < 3; j++) vals[i * 4 + j] = v_R.m[i][j];
            }
            vals[3] = v_t.x; vals[7] = v_t.y; vals[11] = v_t.z;
#pragma unroll
            for (int k = 0; k < 12; k++) vals[k] = warp_sum(vals[k]);
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 12; k++)
                    if (vals[k] != 0.f) atomicAdd(v_viewmats + 16 * cid + k, vals[k]);
            }
        }
    }
    if (!in_range) return;
    if (dens.grad2d != nullptr && d_count > 0.f) {
        dens.grad2d[gid] += d_grad;
        dens.count[gid] += d_count;
        if (dens.radii != nullptr) dens.radii[gid] = fmaxf(dens.radii[gid], d_radius);
    }
    v_means[3 * gid] = v_mean.x; v_means[3 * gid + 1] = v_mean.y; v_means[3 * gid + 2] = v_mean.z;
    if (v_covars != nullptr) {
        // flattened upper triangle (CS/fully_fused_projection_bwd.cu:221-231)
        v_covars[6 * gid + 0] = v_covar.m[0][0];
        v_covars[6 * gid + 1] = v_covar.m[0][1] + v_covar.m[1][0];
        v_covars[6 * gid + 2] = v_covar.m[0][2] + v_covar.m[2][0];
        v_covars[6 * gid + 3] = v_covar.m[1][1];
        v_covars[6 * gid + 4] = v_covar.m[1][2] + v_covar.m[2][1];
        v_covars[6 * gid + 5] = v_covar.m[2][2];
    } else {
        V4 v_q = {0.f, 0.f, 0.f, 0.f};
        V3 v_s = {0.f, 0.f, 0.f};
        quat_scale_to_covar_vjp(q, s, v_covar, v_q, v_s);
        reinterpret_cast<float4 *>(v_quats)[gid] = make_float4(v_q.w, v_q.x, v_q.y, v_q.z);
        v_scales[3 * gid] = v_s.x; v_scales[3 * gid + 1] = v_s.y; v_scales[3 * gid + 2] = v_s.z;
    }
}

// ---------------------------------------------------------------------------------------
// a4: packed forward.  2-D grid (blocks_per_row, C) like the reference so that the COO
// order is (camera, Gaussian) row-major.  MODE 0: per-block counts.  MODE 1: fill.
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads)
projection_packed_kernel(uint32_t C, uint32_t N, const float *__restrict__ means, const float *__restrict__ covars,
                         const float *__restrict__ quats, const float *__restrict__ scales,
                         const float *__restrict__ viewmats, const float *__restrict__ Ks, uint32_t W, uint32_t H,
                         float eps2d, float near_plane, float far_plane, float radius_clip, int camera_model,
                         const int32_t *__restrict__ block_accum, int32_t *__restrict__ block_cnts,
                         int32_t *__restrict__ indptr, int64_t *__restrict__ camera_ids,
                         int64_t *__restrict__ gaussian_ids, int32_t *__restrict__ radii, float *__restrict__ means2d,
                         float *__restrict__ depths, float *__restrict__ conics, float *__restrict__ compensations) {
    const uint32_t blocks_per_row = gridDim.x;
    const uint32_t cid = blockIdx.y;
    const uint32_t block_idx = cid * blocks_per_row + blockIdx.x;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = gid < N;
    ProjOut o;
    if (valid) {
        const Cam cam = load_cam(viewmats + 16 * cid, Ks + 9 * cid);
        const V3 mean = load_v3(means, gid);
        V4 q; V3 s;
        const M3 covar = load_covar(covars, quats, scales, gid, q, s);
        valid = project_one(mean, covar, cam, W, H, eps2d, near_plane, far_plane, radius_clip, camera_model, true, o);
    }
    if (MODE == 0) {
        const int cnt = __syncthreads_count(valid);
        if (threadIdx.x == 0) block_cnts[block_idx] = cnt;
        return;
    }
    // MODE 1: in-block exclusive rank of the valid lanes (ballot + warp prefix in smem)
    __shared__ int warp_cnt[kThreads / 32];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int base = (block_idx > 0) ? block_accum[block_idx - 1] : 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++)
        if (w < (int)wid) base += warp_cnt[w];
    if (valid) {
        const int64_t row = base + __popc(bal & ((1u << lane) - 1u));
        camera_ids[row] = cid;
        gaussian_ids[row] = gid;
        radii[row] = (int32_t)o.radius;
        reinterpret_cast<float2 *>(means2d)[row] = make_float2(o.mean2d.x, o.mean2d.y);
        depths[row] = o.depth_z;  // packed layout stores z (fork behaviour)
        conics[3 * row] = o.conic[0];
        conics[3 * row + 1] = o.conic[1];
        conics[3 * row + 2] = o.conic[2];
        if (compensations != nullptr) compensations[row] = o.comp;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (cid == 0) {
            indptr[0] = 0;
            indptr[C] = block_accum[C * blocks_per_row - 1];
        } else {
            indptr[cid] = block_accum[block_idx - 1];
        }
    }
}

__global__ void copy_last_i32(const int32_t *__restrict__ a, uint32_t n, int32_t *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = n ? a[n - 1] : 0;
}

// a4 backward: one thread per COO row.
__global__ void __launch_bounds__(kThreads)
projection_packed_bwd_kernel(uint32_t C, uint32_t N, uint32_t nnz, const float *__restrict__ means,
                             const float *__restrict__ covars, const float *__restrict__ quats,
                             const float *__restrict__ scales, const float *__restrict__ viewmats,
                             const float *__restrict__ Ks, uint32_t W, uint32_t H, float eps2d, int camera_model,
                             const int64_t *__restrict__ camera_ids, const int64_t *__restrict__ gaussian_ids,
                             const float *__restrict__ conics, const float *__restrict__ compensations,
                             const float *__restrict__ v_means2d, const float *__restrict__ v_depths,
                             const float *__restrict__ v_conics, const float *__restrict__ v_compensations,
                             int sparse_grad, float *__restrict__ v_means, float *__restrict__ v_covars,
                             float *__restrict__ v_quats, float *__restrict__ v_scales,
                             float *__restrict__ v_viewmats) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < nnz;
    V3 v_mean = {0.f, 0.f, 0.f};
    M3 v_covar = m3_zero();
    M3 v_R = m3_zero();
    V3 v_t = {0.f, 0.f, 0.f};
    V4 q = {1.f, 0.f, 0.f, 0.f};
    V3 s = {1.f, 1.f, 1.f};
    uint32_t cid = 0, gid = 0;
    if (valid) {
        cid = (uint32_t)camera_ids[idx];
        gid = (uint32_t)gaussian_ids[idx];
        const Cam cam = load_cam(viewmats + 16 * cid, Ks + 9 * cid);
        const V3 mean = load_v3(means, gid);
        const M3 covar = load_covar(covars, quats, scales, gid, q, s);
        const float conic[3] = {conics[3 * idx], conics[3 * idx + 1], conics[3 * idx + 2]};
        const float vconic[3] = {v_conics[3 * idx], v_conics[3 * idx + 1], v_conics[3 * idx + 2]};
        const V2 v_mean2d = {v_means2d[2 * idx], v_means2d[2 * idx + 1]};
        float comp = 0.f, v_comp = 0.f;
        const bool has_comp = v_compensations != nullptr;
        if (has_comp) { comp = compensations[idx]; v_comp = v_compensations[idx]; }
        project_one_vjp(mean, covar, cam, W, H, eps2d, camera_model, conic, has_comp ? &comp : nullptr,
                        has_comp ? &v_comp : nullptr, v_mean2d, v_depths[idx], vconic, v_mean, v_covar,
                        v_viewmats ? &v_R : nullptr, v_viewmats ? &v_t : nullptr);
        float vq[4] = {0.f, 0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f}, vc[6];
        if (covars != nullptr) {
            vc[0] = v_covar.m[0][0];
            vc[1] = v_covar.m[0][1] + v_covar.m[1][0];
            vc[2] = v_covar.m[0][2] + v_covar.m[2][0];
            vc[3] = v_covar.m[1][1];
            vc[4] = v_covar.m[1][2] + v_covar.m[2][1];
            vc[5] = v_covar.m[2][2];
        } else {
            V4 v_q = {0.f, 0.f, 0.f, 0.f};
            V3 v_s = {0.f, 0.f, 0.f};
            quat_scale_to_covar_vjp(q, s, v_covar, v_q, v_s);
            vq[0] = v_q.w; vq[1] = v_q.x; vq[2] = v_q.y; vq[3] = v_q.z;
            vs[0] = v_s.x; vs[1] = v_s.y; vs[2] = v_s.z;
        }
        if (sparse_grad) {
            v_means[3 * idx] = v_mean.x; v_means[3 * idx + 1] = v_mean.y; v_means[3 * idx + 2] = v_mean.z;
            if (covars != nullptr) {
#pragma unroll
                for (int k = 0; k < 6; k++) v_covars[6 * idx + k] = vc[k];
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) v_quats[4 * idx + k] = vq[k];
#pragma unroll
                for (int k = 0; k < 3; k++) v_scales[3 * idx + k] = vs[k];
            }
        } else {
            atomicAdd(v_means + 3 * gid, v_mean.x);
            atomicAdd(v_means + 3 * gid + 1, v_mean.y);
            atomicAdd(v_means + 3 * gid + 2, v_mean.z);
            if (covars != nullptr) {
#pragma unroll
                for (int k = 0; k < 6; k++) atomicAdd(v_covars + 6 * gid + k, vc[k]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) atomicAdd(v_quats + 4 * gid + k, vq[k]);
#pragma unroll
                for (int k = 0; k < 3; k++) atomicAdd(v_scales + 3 * gid + k, vs[k]);
            }
        }
    }
    if (v_viewmats != nullptr) {
        // rows of one camera are contiguous in COO order: reduce the warp when it is
        // camera-uniform, else fall back to per-lane atomics.
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        const uint32_t cid0 = __shfl_sync(0xffffffffu, cid, __ffs(act ? act : 1u) - 1);
        const bool uniform = __all_sync(0xffffffffu, !valid || cid == cid0);
        float vals[12];
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) vals[i * 4 + j] = v_R.m[i][j];
        }
        vals[3] = v_t.x; vals[7] = v_t.y; vals[11] = v_t.z;
        if (uniform) {
#pragma unroll
            for (int k = 0; k < 12; k++) vals[k] = warp_sum(vals[k]);
            if ((threadIdx.x & 31) == 0 && act) {
#pragma unroll
                for (int k = 0; k < 12; k++) atomicAdd(v_viewmats + 16 * cid0 + k, vals[k]);
            }
        } else if (valid) {
#pragma unroll
            for (int k = 0; k < 12; k++) atomicAdd(v_viewmats + 16 * cid + k, vals[k]);
        }
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b200splat_projection_fwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                                        const float *quats, const float *scales, const float *viewmats,
                                        const float *Ks, uint32_t W, uint32_t H, float eps2d, float near_plane,
                                        float far_plane, float radius_clip, int camera_model, int32_t *radii,
                                        float *means2d, float *depths, float *conics, float *compensations,
                                        void *stream) {
    const char *where = "b200splat_projection_fwd";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    B2S_REQUIRE(camera_model >= 0 && camera_model <= 3, where, "unknown camera model");
    B2S_REQUIRE(covars != nullptr || (quats != nullptr && scales != nullptr), where, "covars or (quats, scales) required");
    if ((uint64_t)C * N == 0) return 0;
    projection_fwd_kernel<<<div_up((uint64_t)C * N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, quats, scales, viewmats, Ks, W, H, eps2d, near_plane, far_plane, radius_clip,
        camera_model, radii, means2d, depths, conics, compensations);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

static int projection_bwd_impl(const char *where, const DensifyState &dens, uint32_t C, uint32_t N, const float *means,
                               const float *covars, const float *quats, const float *scales, const float *viewmats,
                               const float *Ks, uint32_t W, uint32_t H, float eps2d, int camera_model,
                               const int32_t *radii, const float *conics, const float *compensations,
                               const float *v_means2d, const float *v_depths, const float *v_conics,
                               const float *v_compensations, float *v_means, float *v_covars, float *v_quats,
                               float *v_scales, float *v_viewmats, void *stream) {
    B2S_REQUIRE(camera_model >= 0 && camera_model <= 3, where, "unknown camera model");
    B2S_REQUIRE(covars != nullptr || (quats != nullptr && scales != nullptr), where, "covars or (quats, scales) required");
    B2S_REQUIRE(covars == nullptr || v_covars != nullptr, where, "v_covars required with covars");
    B2S_REQUIRE(covars != nullptr || (v_quats != nullptr && v_scales != nullptr), where, "v_quats/v_scales required");
    if (N == 0) return 0;
    projection_bwd_kernel<<<div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        dens, C, N, means, covars, quats, scales, viewmats, Ks, W, H, eps2d, camera_model, radii, conics,
        compensations, v_means2d, v_depths, v_conics, v_compensations, v_means, v_covars, v_quats, v_scales,
        v_viewmats);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_projection_bwd(uint32_t C, uint32_t N, const float *means, const float *covars,
                                        const float *quats, const float *scales, const float *viewmats,
                                        const float *Ks, uint32_t W, uint32_t H, float eps2d, int camera_model,
                                        const int32_t *radii, const float *conics, const float *compensations,
                                        const float *v_means2d, const float *v_depths, const float *v_conics,
                                        const float *v_compensations, float *v_means, float *v_covars,
                                        float *v_quats, float *v_scales, float *v_viewmats, void *stream) {
    const DensifyState off{nullptr, nullptr, nullptr, 0.f, 0.f, 0.f};
    return projection_bwd_impl("b200splat_projection_bwd", off, C, N, means, covars, quats, scales, viewmats, Ks, W, H,
                               eps2d, camera_model, radii, conics, compensations, v_means2d, v_depths, v_conics,
                               v_compensations, v_means, v_covars, v_quats, v_scales, v_viewmats, stream);
}

extern "C" int b200splat_projection_bwd_state(uint32_t C, uint32_t N, const float *means, const float *covars,
                                              const float *quats, const float *scales, const float *viewmats,
                                              const float *Ks, uint32_t W, uint32_t H, float eps2d, int camera_model,
                                              const int32_t *radii, const float *conics, const float *compensations,
                                              const float *v_means2d, const float *v_depths, const float *v_conics,
                                              const float *v_compensations, float *v_means, float *v_covars,
                                              float *v_quats, float *v_scales, float *v_viewmats, float grad_scale_x,
                                              float grad_scale_y, float max_wh, float *state_grad2d,
                                              float *state_count, float *state_radii, void *stream) {
    const char *where = "b200splat_projection_bwd_state";
    B2S_REQUIRE(state_grad2d != nullptr && state_count != nullptr, where, "grad2d and count state arrays are required");
    const DensifyState dens{state_grad2d, state_count, state_radii, grad_scale_x, grad_scale_y, max_wh};
    return projection_bwd_impl(where, dens, C, N, means, covars, quats, scales, viewmats, Ks, W, H, eps2d, camera_model,
                               radii, conics, compensations, v_means2d, v_depths, v_conics, v_compensations, v_means,
                               v_covars, v_quats, v_scales, v_viewmats, stream);
}

extern "C" int b200splat_projection_packed_count(uint32_t C, uint32_t N, const float *means, const float *covars,
                                                 const float *quats, const float *scales, const float *viewmats,
                                                 const float *Ks, uint32_t W, uint32_t H, float eps2d,
                                                 float near_plane, float far_plane, float radius_clip,
                                                 int camera_model, int32_t *block_accum, int32_t *nnz_out,
                                                 void *stream) {
    const char *where = "b200splat_projection_packed_count";
    cudaStream_t st = (cudaStream_t)stream;
    B2S_REQUIRE(camera_model >= 0 && camera_model <= 3, where, "unknown camera model");
    B2S_REQUIRE(C <= 65535, where, "C exceeds the grid.y limit (65535)");
    if ((uint64_t)C * N == 0) {
        cudaMemsetAsync(nnz_out, 0, sizeof(int32_t), st);
        return 0;
    }
    const uint32_t bpr = div_up(N, kThreads);
    dim3 grid(bpr, C, 1);
    projection_packed_kernel<0><<<grid, kThreads, 0, st>>>(C, N, means, covars, quats, scales, viewmats, Ks, W, H,
                                                            eps2d, near_plane, far_plane, radius_clip, camera_model,
                                                            nullptr, block_accum, nullptr, nullptr, nullptr, nullptr,
                                                            nullptr, nullptr, nullptr, nullptr);
    B2S_CHECK_LAUNCH(where);
    // in-place inclusive scan of the per-block counts (small: C*ceil(N/256) entries)
    int rc = small_inclusive_scan_i32(block_accum, (uint64_t)C * bpr, st);
    if (rc) return fail(where, "scan failed");
    copy_last_i32<<<1, 32, 0, st>>>(block_accum, C * bpr, nnz_out);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_projection_packed_fill(uint32_t C, uint32_t N, const float *means, const float *covars,
                                                const float *quats, const float *scales, const float *viewmats,
                                                const float *Ks, uint32_t W, uint32_t H, float eps2d,
                                                float near_plane, float far_plane, float radius_clip,
                                                int camera_model, const int32_t *block_accum, int32_t *indptr,
                                                int64_t *camera_ids, int64_t *gaussian_ids, int32_t *radii,
                                                float *means2d, float *depths, float *conics, float *compensations,
                                                void *stream) {
    const char *where = "b200splat_projection_packed_fill";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    cudaStream_t st = (cudaStream_t)stream;
    B2S_REQUIRE(C <= 65535, where, "C exceeds the grid.y limit (65535)");
    if ((uint64_t)C * N == 0) {
        cudaMemsetAsync(indptr, 0, sizeof(int32_t) * (C + 1), st);
        return 0;
    }
    const uint32_t bpr = div_up(N, kThreads);
    dim3 grid(bpr, C, 1);
    projection_packed_kernel<1><<<grid, kThreads, 0, st>>>(C, N, means, covars, quats, scales, viewmats, Ks, W, H,
                                                            eps2d, near_plane, far_plane, radius_clip, camera_model,
                                                            block_accum, nullptr, indptr, camera_ids, gaussian_ids,
                                                            radii, means2d, depths, conics, compensations);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_projection_packed_bwd(uint32_t C, uint32_t N, uint32_t nnz, const float *means,
                                               const float *covars, const float *quats, const float *scales,
                                               const float *viewmats, const float *Ks, uint32_t W, uint32_t H,
                                               float eps2d, int camera_model, const int64_t *camera_ids,
                                               const int64_t *gaussian_ids, const float *conics,
                                               const float *compensations, const float *v_means2d,
                                               const float *v_depths, const float *v_conics,
                                               const float *v_compensations, int sparse_grad, float *v_means,
                                               float *v_covars, float *v_quats, float *v_scales, float *v_viewmats,
                                               void *stream) {
    const char *where = "b200splat_projection_packed_bwd";
    B2S_REQUIRE(camera_model >= 0 && camera_model <= 3, where, "unknown camera model");
    if (nnz == 0) return 0;
    projection_packed_bwd_kernel<<<div_up(nnz, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        C, N, nnz, means, covars, quats, scales, viewmats, Ks, W, H, eps2d, camera_model, camera_ids, gaussian_ids,
        conics, compensations, v_means2d, v_depths, v_conics, v_compensations, sparse_grad, v_means, v_covars,
        v_quats, v_scales, v_viewmats);
    B2S_CHECK_LAUNCH(where);
    return 0;
}
