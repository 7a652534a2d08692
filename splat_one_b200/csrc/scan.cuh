// scan.cuh — device-wide prefix sums used by the packed projection (a4) and by
// isect_tiles (a6).  The reference calls torch::cumsum for both
// (CS/fully_fused_projection_packed_fwd.cu:352, CS/isect_tiles.cu:200).
//
//  * small_inclusive_scan_i32: in-place, one CTA, no workspace (per-block counters).
//  * lookback_scan_i32_to_i64: single-pass decoupled look-back scan, one read + one
//    write of the data (HBM-bound: 4 B in, 8 B out per element), tile ids handed out by
//    an atomic ticket so that predecessors are always resident.
#pragma once
#include "common.cuh"

namespace b2s {

// ---- in-place single-CTA scan -------------------------------------------------------
static __global__ void __launch_bounds__(1024) small_scan_kernel(int32_t *__restrict__ data, uint64_t n) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        int32_t v = (i < n) ? data[i] : 0;
        int32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int32_t t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t y = __shfl_up_sync(0xffffffffu, t, o);
                if ((int)lane >= o) t += y;
            }
            warp_tot[lane] = t;  // inclusive over warps
        }
        __syncthreads();
        const int32_t carry = carry_s;
        const int32_t incl = x + (wid ? warp_tot[wid - 1] : 0) + carry;
        if (i < n) data[i] = incl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
}

static inline int small_inclusive_scan_i32(int32_t *data, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    small_scan_kernel<<<1, 1024, 0, st>>>(data, n);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---- decoupled look-back scan, int32 -> int64 inclusive ------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096 elements per CTA

// workspace layout: [0] ticket counter (u32, padded to 16 B), then one u64 state per tile
static inline size_t scan_workspace_bytes(uint64_t n) {
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    return 16 + 8 * (tiles + 1);
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

static __global__ void __launch_bounds__(kScanThreads)
lookback_scan_kernel(const int32_t *__restrict__ in, int64_t *__restrict__ out, uint64_t n,
                     unsigned *__restrict__ ticket, unsigned long long *__restrict__ state,
                     int64_t *__restrict__ total_out) {
    constexpr unsigned long long kMask = (1ull << 62) - 1;
    __shared__ unsigned tile_s;
    __shared__ int64_t warp_tot[kScanThreads / 32];
    __shared__ int64_t tile_excl_s;
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = tile_s;
    const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    if (base + kScanItems <= n) {
        const int4 *p = reinterpret_cast<const int4 *>(in + base);
#pragma unroll
        for (int k = 0; k < kScanItems / 4; k++) {
            int4 q = __ldcs(p + k);
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) v[k] = (base + k < n) ? in[base + k] : 0;
    }
    int64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) tsum += v[k];
    // block exclusive scan of thread sums
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((int)lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    int64_t wbase = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        const int64_t t = warp_tot[w];
        if (w < (int)wid) wbase += t;
        tile_total += t;
    }
    const int64_t thread_excl = wbase + x - tsum;
    // look-back by warp 0
    if (wid == 0) {
        if (lane == 0) {
            const unsigned long long flag = (tile == 0) ? 2ull : 1ull;
            atomicExch(state + tile, (flag << 62) | ((unsigned long long)tile_total & kMask));
        }
        int64_t excl = 0;
        int tb = (int)tile - 1;
        while (tb >= 0) {
            const int t = tb - (int)lane;
            unsigned long long s = 2ull << 62;  // virtual tile <0: inclusive prefix 0
            if (t >= 0) {
                do { s = ld_volatile_u64(state + t); } while ((s >> 62) == 0ull);
            }
            const unsigned has_p = __ballot_sync(0xffffffffu, (s >> 62) == 2ull);
            const int first_p = has_p ? (__ffs(has_p) - 1) : 32;
            int64_t val = ((int)lane <= first_p) ? (int64_t)(s & kMask) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            excl += val;
            if (has_p) break;
            tb -= 32;
        }
        if (lane == 0) {
            if (tile > 0)
                atomicExch(state + tile, (2ull << 62) | ((unsigned long long)(excl + tile_total) & kMask));
            tile_excl_s = excl;
            if ((uint64_t)(tile + 1) * kScanTile >= n) *total_out = excl + tile_total;
        }
    }
    __syncthreads();
    int64_t run = tile_excl_s + thread_excl;
    if (base + kScanItems <= n) {
        longlong2 *o2 = reinterpret_cast<longlong2 *>(out + base);
#pragma unroll
        for (int k = 0; k < kScanItems / 2; k++) {
            longlong2 w;
            run += v[2 * k]; w.x = run;
            run += v[2 * k + 1]; w.y = run;
            o2[k] = w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            run += v[k];
            if (base + k < n) out[base + k] = run;
        }
    }
}

// inclusive scan; *total_out (device) = sum of all elements (0 when n == 0).
// `workspace_is_zero`: the caller has already zeroed the workspace on this stream (saves the memset node).
static inline int lookback_scan_i32_to_i64(const int32_t *in, int64_t *out, uint64_t n, int64_t *total_out,
                                           void *workspace, size_t workspace_bytes, cudaStream_t st,
                                           bool workspace_is_zero = false) {
    if (n == 0) {
        return cudaMemsetAsync(total_out, 0, sizeof(int64_t), st) == cudaSuccess ? 0 : 1;
    }
    const size_t need = scan_workspace_bytes(n);
    if (workspace == nullptr || workspace_bytes < need) return 2;
    if (!workspace_is_zero && cudaMemsetAsync(workspace, 0, need, st) != cudaSuccess) return 1;
    const unsigned tiles = (unsigned)((n + kScanTile - 1) / kScanTile);
    unsigned *ticket = reinterpret_cast<unsigned *>(workspace);
    unsigned long long *state = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(workspace) + 16);
    lookback_scan_kernel<<<tiles, kScanThreads, 0, st>>>(in, out, n, ticket, state, total_out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace b2s
