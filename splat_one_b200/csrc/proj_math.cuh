// proj_math.cuh — 3D→2D Gaussian projection math (fwd + VJP) for the four camera
// models of the splat_one gsplat fork.  Plain row-major structs, no glm.
//
// Behavioural spec (what is computed, incl. the fork's quirks) comes from
//   CS/utils.cuh  (quat→R :15-37, Σ=MMᵀ :66-97, ortho :183-251, pinhole :254-373,
//                  fisheye :376-517, spherical :520-594, world→cam :598-658,
//                  2x2 inverse :661-679, blur :682-725)
// with CS = /root/reference/submodules/gsplat/gsplat/cuda/csrc.  Notation here:
// M(r,c) row-major; the reference is glm column-major (M[c][r]).
#pragma once
#include "common.cuh"

namespace b2s {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float w, x, y, z; };          // quaternion, wxyz
struct M2 { float a00, a01, a10, a11; };
struct M3 { float m[3][3]; };             // m[row][col]
struct M23 { float m[2][3]; };            // 2 rows x 3 cols (projection Jacobian)

__device__ __forceinline__ M3 m3_zero() {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = 0.f;
    return r;
}
__device__ __forceinline__ M3 m3_mul(const M3 &a, const M3 &b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
__device__ __forceinline__ M3 m3_mul_bt(const M3 &a, const M3 &b) {  // a * b^T
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.m[i][j] = a.m[i][0] * b.m[j][0] + a.m[i][1] * b.m[j][1] + a.m[i][2] * b.m[j][2];
    return r;
}
__device__ __forceinline__ M3 m3_mul_at(const M3 &a, const M3 &b) {  // a^T * b
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.m[i][j] = a.m[0][i] * b.m[0][j] + a.m[1][i] * b.m[1][j] + a.m[2][i] * b.m[2][j];
    return r;
}
__device__ __forceinline__ M3 m3_transpose(const M3 &a) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
    return r;
}
__device__ __forceinline__ M3 m3_add(const M3 &a, const M3 &b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j];
    return r;
}
__device__ __forceinline__ V3 m3_mulv(const M3 &a, const V3 &v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
__device__ __forceinline__ V3 m3_tmulv(const M3 &a, const V3 &v) {  // a^T v
    return {a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z,
            a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z,
            a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z};
}

// cov2d = J Σ Jᵀ   (2x3 · 3x3 · 3x2)
__device__ __forceinline__ M2 j_cov_jt(const M23 &J, const M3 &S) {
    float t[2][3];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            t[i][j] = J.m[i][0] * S.m[0][j] + J.m[i][1] * S.m[1][j] + J.m[i][2] * S.m[2][j];
    M2 r;
    r.a00 = t[0][0] * J.m[0][0] + t[0][1] * J.m[0][1] + t[0][2] * J.m[0][2];
    r.a01 = t[0][0] * J.m[1][0] + t[0][1] * J.m[1][1] + t[0][2] * J.m[1][2];
    r.a10 = t[1][0] * J.m[0][0] + t[1][1] * J.m[0][1] + t[1][2] * J.m[0][2];
    r.a11 = t[1][0] * J.m[1][0] + t[1][1] * J.m[1][1] + t[1][2] * J.m[1][2];
    return r;
}
// v_Σ += Jᵀ G J  (3x2 · 2x2 · 2x3)
__device__ __forceinline__ void jt_g_j_acc(const M23 &J, const M2 &G, M3 &out) {
    float t[3][2];  // Jᵀ G
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t[i][0] = J.m[0][i] * G.a00 + J.m[1][i] * G.a10;
        t[i][1] = J.m[0][i] * G.a01 + J.m[1][i] * G.a11;
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) out.m[i][j] += t[i][0] * J.m[0][j] + t[i][1] * J.m[1][j];
}
// v_J = G J Σᵀ + Gᵀ J Σ   (2x3)
__device__ __forceinline__ M23 vjac(const M2 &G, const M23 &J, const M3 &S) {
    float gj[2][3], gtj[2][3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        gj[0][j] = G.a00 * J.m[0][j] + G.a01 * J.m[1][j];
        gj[1][j] = G.a10 * J.m[0][j] + G.a11 * J.m[1][j];
        gtj[0][j] = G.a00 * J.m[0][j] + G.a10 * J.m[1][j];
        gtj[1][j] = G.a01 * J.m[0][j] + G.a11 * J.m[1][j];
    }
    M23 r;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.m[i][j] = gj[i][0] * S.m[j][0] + gj[i][1] * S.m[j][1] + gj[i][2] * S.m[j][2] +
                        gtj[i][0] * S.m[0][j] + gtj[i][1] * S.m[1][j] + gtj[i][2] * S.m[2][j];
    return r;
}

// ---- quaternion / scale → covariance (CS/utils.cuh:15-37, 66-97) -------------------
__device__ __forceinline__ M3 quat_to_rotmat(const V4 &q, float &inv_norm, V4 &qn) {
    inv_norm = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    float w = q.w * inv_norm, x = q.x * inv_norm, y = q.y * inv_norm, z = q.z * inv_norm;
    qn = {w, x, y, z};
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    M3 R;
    R.m[0][0] = 1.f - 2.f * (y2 + z2); R.m[0][1] = 2.f * (xy - wz);       R.m[0][2] = 2.f * (xz + wy);
    R.m[1][0] = 2.f * (xy + wz);       R.m[1][1] = 1.f - 2.f * (x2 + z2); R.m[1][2] = 2.f * (yz - wx);
    R.m[2][0] = 2.f * (xz - wy);       R.m[2][1] = 2.f * (yz + wx);       R.m[2][2] = 1.f - 2.f * (x2 + y2);
    return R;
}

__device__ __forceinline__ M3 quat_scale_to_covar(const V4 &q, const V3 &s) {
    float inorm; V4 qn;
    M3 R = quat_to_rotmat(q, inorm, qn);
    M3 M;
#pragma unroll
    for (int i = 0; i < 3; i++) { M.m[i][0] = R.m[i][0] * s.x; M.m[i][1] = R.m[i][1] * s.y; M.m[i][2] = R.m[i][2] * s.z; }
    return m3_mul_bt(M, M);
}

__device__ __forceinline__ M3 covar_from_triu(const float *c) {
    M3 S;
    S.m[0][0] = c[0]; S.m[0][1] = c[1]; S.m[0][2] = c[2];
    S.m[1][0] = c[1]; S.m[1][1] = c[3]; S.m[1][2] = c[4];
    S.m[2][0] = c[2]; S.m[2][1] = c[4]; S.m[2][2] = c[5];
    return S;
}

// VJP of Σ = (R S)(R S)ᵀ wrt quat and scale (CS/utils.cuh:39-63, 99-137).
__device__ __forceinline__ void quat_scale_to_covar_vjp(const V4 &q, const V3 &s, const M3 &v_covar,
                                                        V4 &v_q, V3 &v_s) {
    float inorm; V4 qn;
    M3 R = quat_to_rotmat(q, inorm, qn);
    M3 M;
#pragma unroll
    for (int i = 0; i < 3; i++) { M.m[i][0] = R.m[i][0] * s.x; M.m[i][1] = R.m[i][1] * s.y; M.m[i][2] = R.m[i][2] * s.z; }
    M3 G = m3_add(v_covar, m3_transpose(v_covar));
    M3 vM = m3_mul(G, M);
    M3 vR;
#pragma unroll
    for (int i = 0; i < 3; i++) { vR.m[i][0] = vM.m[i][0] * s.x; vR.m[i][1] = vM.m[i][1] * s.y; vR.m[i][2] = vM.m[i][2] * s.z; }
    v_s.x += R.m[0][0] * vM.m[0][0] + R.m[1][0] * vM.m[1][0] + R.m[2][0] * vM.m[2][0];
    v_s.y += R.m[0][1] * vM.m[0][1] + R.m[1][1] * vM.m[1][1] + R.m[2][1] * vM.m[2][1];
    v_s.z += R.m[0][2] * vM.m[0][2] + R.m[1][2] * vM.m[1][2] + R.m[2][2] * vM.m[2][2];
    const float w = qn.w, x = qn.x, y = qn.y, z = qn.z;
    V4 g;
    g.w = 2.f * (x * (vR.m[2][1] - vR.m[1][2]) + y * (vR.m[0][2] - vR.m[2][0]) + z * (vR.m[1][0] - vR.m[0][1]));
    g.x = 2.f * (-2.f * x * (vR.m[1][1] + vR.m[2][2]) + y * (vR.m[1][0] + vR.m[0][1]) +
                 z * (vR.m[2][0] + vR.m[0][2]) + w * (vR.m[2][1] - vR.m[1][2]));
    g.y = 2.f * (x * (vR.m[1][0] + vR.m[0][1]) - 2.f * y * (vR.m[0][0] + vR.m[2][2]) +
                 z * (vR.m[2][1] + vR.m[1][2]) + w * (vR.m[0][2] - vR.m[2][0]));
    g.z = 2.f * (x * (vR.m[2][0] + vR.m[0][2]) + y * (vR.m[2][1] + vR.m[1][2]) -
                 2.f * z * (vR.m[0][0] + vR.m[1][1]) + w * (vR.m[1][0] - vR.m[0][1]));
    // project out the component along the (normalised) quaternion: d(q/|q|)/dq
    float d = g.w * w + g.x * x + g.y * y + g.z * z;
    v_q.w += (g.w - d * w) * inorm;
    v_q.x += (g.x - d * x) * inorm;
    v_q.y += (g.y - d * y) * inorm;
    v_q.z += (g.z - d * z) * inorm;
}

// ---- camera ---------------------------------------------------------------------------
struct Cam {
    M3 R; V3 t;
    float fx, fy, cx, cy;
};
__device__ __forceinline__ Cam load_cam(const float *__restrict__ vm, const float *__restrict__ K) {
    Cam c;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c.R.m[i][j] = vm[i * 4 + j];
    c.t = {vm[3], vm[7], vm[11]};
    c.fx = K[0]; c.cx = K[2]; c.fy = K[4]; c.cy = K[5];
    return c;
}

// Each *_proj returns the Jacobian it used (needed by the VJP) through J.
// pinhole, CS/utils.cuh:254-293: guard-band clamp of x/z, y/z to ±(lim + 0.3 tanfov).
__device__ __forceinline__ void persp_proj(const V3 &p, const M3 &cov, const Cam &c, uint32_t W, uint32_t H,
                                           M2 &cov2d, V2 &mean2d, M23 &J) {
    float x = p.x, y = p.y, z = p.z;
    float tan_fovx = 0.5f * W / c.fx, tan_fovy = 0.5f * H / c.fy;
    float lim_x_pos = (W - c.cx) / c.fx + 0.3f * tan_fovx;
    float lim_x_neg = c.cx / c.fx + 0.3f * tan_fovx;
    float lim_y_pos = (H - c.cy) / c.fy + 0.3f * tan_fovy;
    float lim_y_neg = c.cy / c.fy + 0.3f * tan_fovy;
    float rz = 1.f / z, rz2 = rz * rz;
    float tx = z * fminf(lim_x_pos, fmaxf(-lim_x_neg, x * rz));
    float ty = z * fminf(lim_y_pos, fmaxf(-lim_y_neg, y * rz));
    J.m[0][0] = c.fx * rz; J.m[0][1] = 0.f;       J.m[0][2] = -c.fx * tx * rz2;
    J.m[1][0] = 0.f;       J.m[1][1] = c.fy * rz; J.m[1][2] = -c.fy * ty * rz2;
    cov2d = j_cov_jt(J, cov);
    mean2d = {c.fx * x * rz + c.cx, c.fy * y * rz + c.cy};
}

__device__ __forceinline__ void persp_proj_vjp(const V3 &p, const M3 &cov, const Cam &c, uint32_t W, uint32_t H,
                                               const M2 &v_cov2d, const V2 &v_mean2d, V3 &v_p, M3 &v_cov) {
    float x = p.x, y = p.y, z = p.z;
    float tan_fovx = 0.5f * W / c.fx, tan_fovy = 0.5f * H / c.fy;
    float lim_x_pos = (W - c.cx) / c.fx + 0.3f * tan_fovx;
    float lim_x_neg = c.cx / c.fx + 0.3f * tan_fovx;
    float lim_y_pos = (H - c.cy) / c.fy + 0.3f * tan_fovy;
    float lim_y_neg = c.cy / c.fy + 0.3f * tan_fovy;
    float rz = 1.f / z, rz2 = rz * rz, rz3 = rz2 * rz;
    float tx = z * fminf(lim_x_pos, fmaxf(-lim_x_neg, x * rz));
    float ty = z * fminf(lim_y_pos, fmaxf(-lim_y_neg, y * rz));
    M23 J;
    J.m[0][0] = c.fx * rz; J.m[0][1] = 0.f;       J.m[0][2] = -c.fx * tx * rz2;
    J.m[1][0] = 0.f;       J.m[1][1] = c.fy * rz; J.m[1][2] = -c.fy * ty * rz2;
    jt_g_j_acc(J, v_cov2d, v_cov);
    v_p.x += c.fx * rz * v_mean2d.x;
    v_p.y += c.fy * rz * v_mean2d.y;
    v_p.z += -(c.fx * x * v_mean2d.x + c.fy * y * v_mean2d.y) * rz2;
    M23 vJ = vjac(v_cov2d, J, cov);
    if (x * rz <= lim_x_pos && x * rz >= -lim_x_neg) v_p.x += -c.fx * rz2 * vJ.m[0][2];
    else v_p.z += -c.fx * rz3 * vJ.m[0][2] * tx;
    if (y * rz <= lim_y_pos && y * rz >= -lim_y_neg) v_p.y += -c.fy * rz2 * vJ.m[1][2];
    else v_p.z += -c.fy * rz3 * vJ.m[1][2] * ty;
    v_p.z += -c.fx * rz2 * vJ.m[0][0] - c.fy * rz2 * vJ.m[1][1] + 2.f * c.fx * tx * rz3 * vJ.m[0][2] +
             2.f * c.fy * ty * rz3 * vJ.m[1][2];
}

// orthographic, CS/utils.cuh:183-251
__device__ __forceinline__ void ortho_proj(const V3 &p, const M3 &cov, const Cam &c, M2 &cov2d, V2 &mean2d) {
    M23 J;
    J.m[0][0] = c.fx; J.m[0][1] = 0.f;  J.m[0][2] = 0.f;
    J.m[1][0] = 0.f;  J.m[1][1] = c.fy; J.m[1][2] = 0.f;
    cov2d = j_cov_jt(J, cov);
    mean2d = {c.fx * p.x + c.cx, c.fy * p.y + c.cy};
}
__device__ __forceinline__ void ortho_proj_vjp(const Cam &c, const M2 &v_cov2d, const V2 &v_mean2d, V3 &v_p, M3 &v_cov) {
    M23 J;
    J.m[0][0] = c.fx; J.m[0][1] = 0.f;  J.m[0][2] = 0.f;
    J.m[1][0] = 0.f;  J.m[1][1] = c.fy; J.m[1][2] = 0.f;
    jt_g_j_acc(J, v_cov2d, v_cov);
    v_p.x += c.fx * v_mean2d.x;
    v_p.y += c.fy * v_mean2d.y;
}

// fisheye (equidistant), CS/utils.cuh:376-415
__device__ __forceinline__ M23 fisheye_jac(const V3 &p, const Cam &c) {
    float x = p.x, y = p.y, z = p.z;
    const float eps = 0.0000001f;
    float xy_len = sqrtf(x * x + y * y) + eps;
    float x2 = x * x + eps, y2 = y * y, xy = x * y;
    float x2y2 = x2 + y2;
    float x2y2z2_inv = 1.f / (x2y2 + z * z);
    float b = atan2f(xy_len, z) / xy_len / x2y2;
    float a = z * x2y2z2_inv / x2y2;
    M23 J;
    J.m[0][0] = c.fx * (x2 * a + y2 * b); J.m[0][1] = c.fx * xy * (a - b);       J.m[0][2] = -c.fx * x * x2y2z2_inv;
    J.m[1][0] = c.fy * xy * (a - b);       J.m[1][1] = c.fy * (y2 * a + x2 * b); J.m[1][2] = -c.fy * y * x2y2z2_inv;
    return J;
}
__device__ __forceinline__ void fisheye_proj(const V3 &p, const M3 &cov, const Cam &c, M2 &cov2d, V2 &mean2d) {
    float x = p.x, y = p.y, z = p.z;
    const float eps = 0.0000001f;
    float xy_len = sqrtf(x * x + y * y) + eps;
    float theta = atan2f(xy_len, z + eps);
    mean2d = {x * c.fx * theta / xy_len + c.cx, y * c.fy * theta / xy_len + c.cy};
    M23 J = fisheye_jac(p, c);
    cov2d = j_cov_jt(J, cov);
}
// CS/utils.cuh:417-517 (analytic ∂J/∂mean terms kept)
__device__ __forceinline__ void fisheye_proj_vjp(const V3 &p, const M3 &cov, const Cam &c, const M2 &v_cov2d,
                                                 const V2 &v_mean2d, V3 &v_p, M3 &v_cov) {
    float x = p.x, y = p.y, z = p.z;
    const float fx = c.fx, fy = c.fy;
    const float eps = 0.0000001f;
    float x2 = x * x + eps, y2 = y * y, xy = x * y;
    float x2y2 = x2 + y2;
    float len_xy = sqrtf(x * x + y * y) + eps;
    const float x2y2z2 = x2y2 + z * z;
    float x2y2z2_inv = 1.f / x2y2z2;
    const float theta = atan2f(len_xy, z);
    float b = theta / len_xy / x2y2;
    float a = z * x2y2z2_inv / x2y2;
    v_p.x += fx * (x2 * a + y2 * b) * v_mean2d.x + fy * xy * (a - b) * v_mean2d.y;
    v_p.y += fx * xy * (a - b) * v_mean2d.x + fy * (y2 * a + x2 * b) * v_mean2d.y;
    v_p.z += -fx * x * x2y2z2_inv * v_mean2d.x - fy * y * x2y2z2_inv * v_mean2d.y;
    M23 J;
    J.m[0][0] = fx * (x2 * a + y2 * b); J.m[0][1] = fx * xy * (a - b);       J.m[0][2] = -fx * x * x2y2z2_inv;
    J.m[1][0] = fy * xy * (a - b);       J.m[1][1] = fy * (y2 * a + x2 * b); J.m[1][2] = -fy * y * x2y2z2_inv;
    jt_g_j_acc(J, v_cov2d, v_cov);
    M23 vJ = vjac(v_cov2d, J, cov);
    float l4 = x2y2z2 * x2y2z2;
    float E = -l4 * x2y2 * theta + x2y2z2 * x2y2 * len_xy * z;
    float F = 3.f * l4 * theta - 3.f * x2y2z2 * len_xy * z - 2.f * x2y2 * len_xy * z;
    float A = x * (3.f * E + x2 * F);
    float B = y * (E + x2 * F);
    float Cc = x * (E + y2 * F);
    float D = y * (3.f * E + y2 * F);
    float S1 = x2 - y2 - z * z;
    float S2 = y2 - x2 - z * z;
    float inv1 = x2y2z2_inv * x2y2z2_inv;
    float inv2 = inv1 / (x2y2 * x2y2 * len_xy);
    float dJ_dx00 = fx * A * inv2, dJ_dx01 = fx * B * inv2, dJ_dx02 = fx * S1 * inv1;
    float dJ_dx10 = fy * B * inv2, dJ_dx11 = fy * Cc * inv2, dJ_dx12 = 2.f * fy * xy * inv1;
    float dJ_dy00 = dJ_dx01, dJ_dy01 = fx * Cc * inv2, dJ_dy02 = 2.f * fx * xy * inv1;
    float dJ_dy10 = dJ_dx11, dJ_dy11 = fy * D * inv2, dJ_dy12 = fy * S2 * inv1;
    float dJ_dz00 = dJ_dx02, dJ_dz01 = dJ_dy02, dJ_dz02 = 2.f * fx * x * z * inv1;
    float dJ_dz10 = dJ_dx12, dJ_dz11 = dJ_dy12, dJ_dz12 = 2.f * fy * y * z * inv1;
    v_p.x += dJ_dx00 * vJ.m[0][0] + dJ_dx01 * vJ.m[0][1] + dJ_dx02 * vJ.m[0][2] + dJ_dx10 * vJ.m[1][0] +
             dJ_dx11 * vJ.m[1][1] + dJ_dx12 * vJ.m[1][2];
    v_p.y += dJ_dy00 * vJ.m[0][0] + dJ_dy01 * vJ.m[0][1] + dJ_dy02 * vJ.m[0][2] + dJ_dy10 * vJ.m[1][0] +
             dJ_dy11 * vJ.m[1][1] + dJ_dy12 * vJ.m[1][2];
    v_p.z += dJ_dz00 * vJ.m[0][0] + dJ_dz01 * vJ.m[0][1] + dJ_dz02 * vJ.m[0][2] + dJ_dz10 * vJ.m[1][0] +
             dJ_dz11 * vJ.m[1][1] + dJ_dz12 * vJ.m[1][2];
}

// equirectangular 360°, CS/utils.cuh:520-557.  The fork evaluates the π factors in
// double (`#define M_PI` :10) and rounds once to float; reproduced here.
#define B2S_PI 3.14159265358979323846
__device__ __forceinline__ M23 spherical_jac(const V3 &p, uint32_t W, uint32_t H, float r) {
    float x = p.x, y = p.y, z = p.z;
    float denom_xz = fmaf(z, z, x * x) + 1e-8f;
    float xz_norm = sqrtf(denom_xz);
    float denom_r2 = r * r + 1e-8f;
    M23 J;
    J.m[0][0] = (float)(W / (2 * B2S_PI) * (double)(z / denom_xz));
    J.m[1][0] = (float)(H / B2S_PI * (double)(-(x * y) / (denom_r2 * xz_norm)));
    J.m[0][1] = 0.f;
    J.m[1][1] = (float)(H / B2S_PI * (double)(xz_norm / denom_r2));
    J.m[0][2] = (float)(W / (2 * B2S_PI) * (double)(-x / denom_xz));
    J.m[1][2] = (float)(H / B2S_PI * (double)(-(z * y) / (denom_r2 * xz_norm)));
    return J;
}
__device__ __forceinline__ void spherical_proj(const V3 &p, const M3 &cov, uint32_t W, uint32_t H, M2 &cov2d,
                                               V2 &mean2d) {
    float x = p.x, y = p.y, z = p.z;
    float r = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));  // same contraction as the reference build (see project_one)
    float longitude = atan2f(x, z);
    float latitude = asinf(y / r);
    float normalized_latitude = (float)((double)latitude / (B2S_PI / 2.0));
    float normalized_longitude = (float)((double)longitude / B2S_PI);
    mean2d = {(normalized_longitude + 1) * W / 2, (normalized_latitude + 1) * H / 2};
    M23 J = spherical_jac(p, W, H, r);
    cov2d = j_cov_jt(J, cov);
}
// CS/utils.cuh:559-594: only Jᵀ v_mean2d and Jᵀ v_cov2d J (no ∂J/∂mean term), and r
// carries +1e-8 inside the sqrt here (unlike the forward).
__device__ __forceinline__ void spherical_proj_vjp(const V3 &p, uint32_t W, uint32_t H, const M2 &v_cov2d,
                                                   const V2 &v_mean2d, V3 &v_p, M3 &v_cov) {
    float r = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z + 1e-8f);
    M23 J = spherical_jac(p, W, H, r);
    v_p.x += J.m[0][0] * v_mean2d.x + J.m[1][0] * v_mean2d.y;
    v_p.y += J.m[0][1] * v_mean2d.x + J.m[1][1] * v_mean2d.y;
    v_p.z += J.m[0][2] * v_mean2d.x + J.m[1][2] * v_mean2d.y;
    jt_g_j_acc(J, v_cov2d, v_cov);
}

// ---- full per-pair forward ------------------------------------------------------------
struct ProjOut {
    V2 mean2d;
    float depth_norm;   // |mean_c|   (unpacked path stores this, CS/fully_fused_projection_fwd.cu:204-209)
    float depth_z;      // mean_c.z   (packed path stores this, CS/fully_fused_projection_packed_fwd.cu:249)
    float conic[3];
    float comp;
    float radius;
};

// `packed_rules` selects the packed kernel's validity/radius rules
// (CS/fully_fused_projection_packed_fwd.cu:186-202) instead of the unpacked ones
// (CS/fully_fused_projection_fwd.cu:171-188).  Returns false when culled.
__device__ __forceinline__ bool project_one(const V3 &mean, const M3 &covar, const Cam &cam, uint32_t W, uint32_t H,
                                            float eps2d, float near_plane, float far_plane, float radius_clip,
                                            int camera_model, bool packed_rules, ProjOut &o) {
    V3 pc = m3_mulv(cam.R, mean);
    pc.x += cam.t.x; pc.y += cam.t.y; pc.z += cam.t.z;
    // |mean_c| with the reference build's contraction, fma(z, z, fma(y, y, x·x)) (glm::length in
    // CS/fully_fused_projection_fwd.cu:73-85, 204-209 as nvcc compiles it): the bits of this value are
    // the depth half of the sort keys, and for the spherical model y/|mean_c| feeds asin, whose
    // derivative near the poles turns one ulp into 1e-2 px
    float rnorm = sqrtf(fmaf(pc.z, pc.z, fmaf(pc.y, pc.y, pc.x * pc.x)));
    if (camera_model != B200SPLAT_SPHERICAL) {
        if (pc.z < near_plane || pc.z > far_plane) return false;
    } else {
        if (rnorm < near_plane || rnorm > far_plane) return false;
    }
    M3 covar_c = m3_mul_bt(m3_mul(cam.R, covar), cam.R);
    M2 cov2d; V2 m2; M23 J;
    switch (camera_model) {
        case B200SPLAT_PINHOLE: persp_proj(pc, covar_c, cam, W, H, cov2d, m2, J); break;
        case B200SPLAT_ORTHO: ortho_proj(pc, covar_c, cam, cov2d, m2); break;
        case B200SPLAT_FISHEYE: fisheye_proj(pc, covar_c, cam, cov2d, m2); break;
        default: spherical_proj(pc, covar_c, W, H, cov2d, m2); break;
    }
    // add_blur, CS/utils.cuh:682-689
    float det_orig = cov2d.a00 * cov2d.a11 - cov2d.a01 * cov2d.a10;
    cov2d.a00 += eps2d; cov2d.a11 += eps2d;
    float det = cov2d.a00 * cov2d.a11 - cov2d.a01 * cov2d.a10;
    o.comp = sqrtf(fmaxf(0.f, det_orig / det));
    if (packed_rules && det <= 0.f) return false;
    // inverse, CS/utils.cuh:661-672.  det<=0 leaves the reference's output undefined
    // (unpacked path only); we write zeros there.
    if (det > 0.f) {
        float inv_det = 1.f / det;
        o.conic[0] = cov2d.a11 * inv_det;
        o.conic[1] = -cov2d.a01 * inv_det;
        o.conic[2] = cov2d.a00 * inv_det;
    } else {
        o.conic[0] = o.conic[1] = o.conic[2] = 0.f;
    }
    float b = 0.5f * (cov2d.a00 + cov2d.a11);
    float radius;
    if (!packed_rules) {
        float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
        radius = ceilf(3.f * sqrtf(v1));
    } else {
        float v1 = b + sqrtf(fmaxf(0.1f, b * b - det));
        float v2 = b - sqrtf(fmaxf(0.1f, b * b - det));
        radius = ceilf(3.f * sqrtf(fmaxf(v1, v2)));
    }
    if (radius <= radius_clip) return false;
    if (camera_model != B200SPLAT_SPHERICAL) {
        if (m2.x + radius <= 0 || m2.x - radius >= W || m2.y + radius <= 0 || m2.y - radius >= H) return false;
    }
    o.mean2d = m2;
    o.depth_norm = rnorm;
    o.depth_z = pc.z;
    o.radius = radius;
    return true;
}

// ---- full per-pair backward -----------------------------------------------------------
// Inputs: conic (fwd output), cotangents.  Outputs (accumulated): v_mean (world), v_covar
// (world, full 3x3), v_R, v_t.  CS/fully_fused_projection_bwd.cu:66-206.
__device__ __forceinline__ void project_one_vjp(const V3 &mean, const M3 &covar, const Cam &cam, uint32_t W, uint32_t H,
                                                float eps2d, int camera_model, const float *conic,
                                                const float *comp, const float *v_comp, const V2 &v_mean2d,
                                                float v_depth, const float *v_conic, V3 &v_mean, M3 &v_covar,
                                                M3 *v_R, V3 *v_t) {
    // v_cov2d = -P V P   with V the symmetrised conic cotangent (off-diagonal halved)
    const float p00 = conic[0], p01 = conic[1], p11 = conic[2];
    const float g00 = v_conic[0], g01 = v_conic[1] * .5f, g11 = v_conic[2];
    // T = P V
    float t00 = p00 * g00 + p01 * g01, t01 = p00 * g01 + p01 * g11;
    float t10 = p01 * g00 + p11 * g01, t11 = p01 * g01 + p11 * g11;
    M2 v_cov2d;
    v_cov2d.a00 = -(t00 * p00 + t01 * p01);
    v_cov2d.a01 = -(t00 * p01 + t01 * p11);
    v_cov2d.a10 = -(t10 * p00 + t11 * p01);
    v_cov2d.a11 = -(t10 * p01 + t11 * p11);
    if (v_comp != nullptr) {  // add_blur_vjp, CS/utils.cuh:691-725
        const float compensation = *comp, v_compensation = *v_comp;
        float det_conic = p00 * p11 - p01 * p01;
        float v_sqr_comp = v_compensation * 0.5f / (compensation + 1e-6f);
        float one_minus_sqr_comp = 1.f - compensation * compensation;
        v_cov2d.a00 += v_sqr_comp * (one_minus_sqr_comp * p00 - eps2d * det_conic);
        v_cov2d.a01 += v_sqr_comp * (one_minus_sqr_comp * p01);
        v_cov2d.a10 += v_sqr_comp * (one_minus_sqr_comp * p01);
        v_cov2d.a11 += v_sqr_comp * (one_minus_sqr_comp * p11 - eps2d * det_conic);
    }
    V3 pc = m3_mulv(cam.R, mean);
    pc.x += cam.t.x; pc.y += cam.t.y; pc.z += cam.t.z;
    M3 covar_c = m3_mul_bt(m3_mul(cam.R, covar), cam.R);
    V3 v_pc = {0.f, 0.f, 0.f};
    M3 v_covar_c = m3_zero();
    switch (camera_model) {
        case B200SPLAT_PINHOLE: persp_proj_vjp(pc, covar_c, cam, W, H, v_cov2d, v_mean2d, v_pc, v_covar_c); break;
        case B200SPLAT_ORTHO: ortho_proj_vjp(cam, v_cov2d, v_mean2d, v_pc, v_covar_c); break;
        case B200SPLAT_FISHEYE: fisheye_proj_vjp(pc, covar_c, cam, v_cov2d, v_mean2d, v_pc, v_covar_c); break;
        default: spherical_proj_vjp(pc, W, H, v_cov2d, v_mean2d, v_pc, v_covar_c); break;
    }
    // depth cotangent goes to z for every model (CS/fully_fused_projection_bwd.cu:195)
    v_pc.z += v_depth;
    // world→cam VJPs, CS/utils.cuh:608-658
    V3 vm = m3_tmulv(cam.R, v_pc);
    v_mean.x += vm.x; v_mean.y += vm.y; v_mean.z += vm.z;
    M3 vc = m3_mul(m3_mul_at(cam.R, v_covar_c), cam.R);
    v_covar = m3_add(v_covar, vc);
    if (v_R != nullptr) {
        const float pv[3] = {v_pc.x, v_pc.y, v_pc.z};
        const float mw[3] = {mean.x, mean.y, mean.z};
        M3 a = m3_mul_bt(m3_mul(v_covar_c, cam.R), covar);               // G R Σᵀ
        M3 b = m3_mul(m3_mul(m3_transpose(v_covar_c), cam.R), covar);    // Gᵀ R Σ
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) v_R->m[i][j] += pv[i] * mw[j] + a.m[i][j] + b.m[i][j];
        v_t->x += v_pc.x; v_t->y += v_pc.y; v_t->z += v_pc.z;
    }
}

}  // namespace b2s
