// isect.cu — Gaussian/tile intersection (a6) and tile-offset encode (a7).
// Replaces CS/isect_tiles.cu:17-105 (kernel), :107-307 (host) and :309-390.
//
// Contract (SURVEY.md §8a "Bit-exactness notes"): given identical means2d / radii /
// depths (+ ids) the outputs tiles_per_gauss, isect_ids, flatten_ids and offsets equal
// the reference bit for bit: key = cam << (32+tile_n_bits) | tile << 32 | (int64)(int32
// depth bits), row-major tile enumeration, stable sort on the low
// 32 + tile_n_bits + cam_n_bits bits.
#include "common.cuh"
#include "scan.cuh"

namespace b2s {

struct TileRect { uint32_t x0, y0, x1, y1; };

// CS/isect_tiles.cu:60-70.  The float→uint32 conversion of a negative floor() saturates
// to 0 on the GPU (cvt.rzi.u32.f32), which is what makes the reference's
// `max(0, (uint32_t)...)` a clamp; written explicitly here.  tile_size is a power of two
// in every supported use, so the (fast-math) divisions of the reference are exact.
__device__ __forceinline__ TileRect tile_rect(float mx, float my, float radius, float ts, uint32_t tw,
                                              uint32_t th) {
    // div.approx like the reference's --use_fast_math build; exact for power-of-two ts
    const float tr = __fdividef(radius, ts), tx = __fdividef(mx, ts), ty = __fdividef(my, ts);
    TileRect r;
    r.x0 = min(__float2uint_rz(floorf(tx - tr)), tw);
    r.y0 = min(__float2uint_rz(floorf(ty - tr)), th);
    r.x1 = min(__float2uint_rz(ceilf(tx + tr)), tw);
    r.y1 = min(__float2uint_rz(ceilf(ty + tr)), th);
    return r;
}

// SUM: also accumulate the total (n_isects) with one atomic per block — the depth-first ordering
// needs only the total at this point, not the running sums of CS/isect_tiles.cu:200.
template <bool SUM>
__global__ void __launch_bounds__(kThreads)
isect_count_kernel(uint64_t n_elems, const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                   const float *__restrict__ depths, float ts, uint32_t tw, uint32_t th,
                   int32_t *__restrict__ tiles_per_gauss, int64_t *__restrict__ neg_depth_flag,
                   unsigned long long *__restrict__ total_out) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int32_t cnt = 0;
    if (idx < n_elems) {
        const float radius = (float)radii[idx];
        if (radius > 0.f) {
            const float2 m = reinterpret_cast<const float2 *>(means2d)[idx];
            const TileRect r = tile_rect(m.x, m.y, radius, ts, tw, th);
            cnt = (int32_t)((r.y1 - r.y0) * (r.x1 - r.x0));
            // a set sign bit sign-extends into the tile/camera fields of the reference key
            // (CS/isect_tiles.cu:92): outside the contract, routed to the generic sort
            if (cnt > 0 && depths != nullptr && __float_as_int(depths[idx]) < 0) *neg_depth_flag = 1;
        }
        tiles_per_gauss[idx] = cnt;
    }
    if (SUM) {
        __shared__ unsigned long long s_sum;
        if (threadIdx.x == 0) s_sum = 0;
        __syncthreads();
        const unsigned w = (unsigned)__reduce_add_sync(0xffffffffu, (unsigned)cnt);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_sum, (unsigned long long)w);
        __syncthreads();
        if (threadIdx.x == 0 && s_sum) atomicAdd(total_out, s_sum);
    }
}

__global__ void __launch_bounds__(kThreads)
isect_fill_kernel(int packed, uint32_t N, uint64_t n_elems, const int64_t *__restrict__ camera_ids,
                  const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                  const float *__restrict__ depths, const int64_t *__restrict__ cum_tiles, float ts,
                  uint32_t tw, uint32_t th, uint32_t tile_n_bits, int64_t *__restrict__ isect_ids,
                  int32_t *__restrict__ flatten_ids) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_elems) return;
    const float radius = (float)radii[idx];
    if (radius <= 0.f) return;
    const float2 m = reinterpret_cast<const float2 *>(means2d)[idx];
    const TileRect r = tile_rect(m.x, m.y, radius, ts, tw, th);
    const int64_t cid = packed ? camera_ids[idx] : (int64_t)(idx / N);
    const int64_t cid_enc = cid << (32 + tile_n_bits);
    // sign-extending reinterpretation of the fp32 depth bits (CS/isect_tiles.cu:92)
    const int64_t depth_enc = (int64_t)__float_as_int(depths[idx]);
    int64_t cur = (idx == 0) ? 0 : cum_tiles[idx - 1];
    for (uint32_t i = r.y0; i < r.y1; ++i) {
        for (uint32_t j = r.x0; j < r.x1; ++j) {
            const int64_t tile_id = (int64_t)(i * tw + j);
            isect_ids[cur] = cid_enc | (tile_id << 32) | depth_enc;
            flatten_ids[cur] = (int32_t)idx;
            ++cur;
        }
    }
}

// a7, CS/isect_tiles.cu:309-355: offsets[k] = first sorted index whose (cam,tile) >= k.
__global__ void __launch_bounds__(kThreads)
offset_encode_kernel(uint64_t n_isects, const int64_t *__restrict__ isect_ids, uint32_t total_tiles, uint32_t n_tiles,
                     uint32_t tile_n_bits, int32_t *__restrict__ offsets) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_isects) return;
    const int64_t tile_mask = ((int64_t)1 << tile_n_bits) - 1;
    const int64_t cur_hi = isect_ids[idx] >> 32;
    const int64_t id_curr = (cur_hi >> tile_n_bits) * n_tiles + (cur_hi & tile_mask);
    if (idx == 0) {
        for (int64_t i = 0; i <= id_curr && i < total_tiles; ++i) offsets[i] = 0;
    }
    if (idx == n_isects - 1) {
        for (int64_t i = id_curr + 1; i < total_tiles; ++i) offsets[i] = (int32_t)n_isects;
    }
    if (idx > 0) {
        const int64_t prev_hi = isect_ids[idx - 1] >> 32;
        if (prev_hi == cur_hi) return;
        const int64_t id_prev = (prev_hi >> tile_n_bits) * n_tiles + (prev_hi & tile_mask);
        for (int64_t i = id_prev + 1; i <= id_curr && i < total_tiles; ++i) offsets[i] = (int32_t)idx;
    }
}

static inline uint32_t bit_length(uint32_t v) {
    uint32_t b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

}  // namespace b2s

using namespace b2s;

extern "C" size_t b200splat_scan_workspace_bytes(uint64_t n_elems) { return scan_workspace_bytes(n_elems); }

extern "C" int b200splat_isect_count(int packed, uint32_t C, uint32_t N, uint32_t nnz, const float *means2d,
                                     const int32_t *radii, const float *depths, uint32_t tile_size, uint32_t tile_width,
                                     uint32_t tile_height, int32_t *tiles_per_gauss, int64_t *cum_tiles,
                                     int64_t *n_isects_out, void *scan_workspace, size_t scan_workspace_bytes_,
                                     void *stream) {
    const char *where = "b200splat_isect_count";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    cudaStream_t st = (cudaStream_t)stream;
    B2S_REQUIRE(tile_size > 0, where, "tile_size must be positive");
    const uint64_t n_elems = packed ? (uint64_t)nnz : (uint64_t)C * N;
    cudaMemsetAsync(n_isects_out, 0, 2 * sizeof(int64_t), st);
    if (n_elems == 0) return 0;
    if (cum_tiles == nullptr) {
        // total only (the depth-first ordering scans the counts later, in depth order)
        isect_count_kernel<true><<<div_up(n_elems, kThreads), kThreads, 0, st>>>(
            n_elems, means2d, radii, depths, (float)tile_size, tile_width, tile_height, tiles_per_gauss, n_isects_out + 1,
            reinterpret_cast<unsigned long long *>(n_isects_out));
        B2S_CHECK_LAUNCH(where);
        return 0;
    }
    isect_count_kernel<false><<<div_up(n_elems, kThreads), kThreads, 0, st>>>(
        n_elems, means2d, radii, depths, (float)tile_size, tile_width, tile_height, tiles_per_gauss, n_isects_out + 1,
        nullptr);
    B2S_CHECK_LAUNCH(where);
    const int rc = lookback_scan_i32_to_i64(tiles_per_gauss, cum_tiles, n_elems, n_isects_out, scan_workspace,
                                            scan_workspace_bytes_, st);
    if (rc == 2) return fail(where, "scan workspace too small (see b200splat_scan_workspace_bytes)");
    if (rc) return fail_cuda(where, cudaGetLastError());
    return 0;
}

extern "C" int b200splat_isect_fill(int packed, uint32_t C, uint32_t N, uint32_t nnz, const int64_t *camera_ids,
                                    const float *means2d, const int32_t *radii, const float *depths,
                                    const int64_t *cum_tiles, uint32_t tile_size, uint32_t tile_width,
                                    uint32_t tile_height, int64_t *isect_ids, int32_t *flatten_ids, void *stream) {
    const char *where = "b200splat_isect_fill";
    B2S_REQUIRE_ALIGNED8(means2d, where);
    const uint64_t n_elems = packed ? (uint64_t)nnz : (uint64_t)C * N;
    B2S_REQUIRE(!packed || camera_ids != nullptr, where, "camera_ids required when packed");
    const uint32_t n_tiles = tile_width * tile_height;
    const uint32_t tile_n_bits = bit_length(n_tiles), cam_n_bits = bit_length(C);
    B2S_REQUIRE(tile_n_bits + cam_n_bits <= 32, where, "camera and tile ids do not fit in 32 bits");
    if (n_elems == 0) return 0;
    isect_fill_kernel<<<div_up(n_elems, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        packed, N, n_elems, camera_ids, means2d, radii, depths, cum_tiles, (float)tile_size, tile_width,
        tile_height, tile_n_bits, isect_ids, flatten_ids);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_isect_offset_encode(uint64_t n_isects, const int64_t *isect_ids, uint32_t C,
                                             uint32_t tile_width, uint32_t tile_height, int32_t *offsets,
                                             void *stream) {
    const char *where = "b200splat_isect_offset_encode";
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t n_tiles = tile_width * tile_height;
    const uint64_t total = (uint64_t)C * n_tiles;
    if (total == 0) return 0;
    if (n_isects == 0) {
        cudaMemsetAsync(offsets, 0, total * sizeof(int32_t), st);
        return 0;
    }
    offset_encode_kernel<<<div_up(n_isects, kThreads), kThreads, 0, st>>>(n_isects, isect_ids, (uint32_t)total, n_tiles,
                                                                           bit_length(n_tiles), offsets);
    B2S_CHECK_LAUNCH(where);
    return 0;
}
