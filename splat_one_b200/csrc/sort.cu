// sort.cu — ordering stage of isect_tiles (a6): produce the (isect_ids, flatten_ids) arrays
// sorted exactly like the reference's stable LSD radix sort of the 64-bit keys
// `cam | tile | depth bits` (CS/isect_tiles.cu:252-300).
//
// Two implementations behind the C ABI:
//
//  * b200splat_isect_sort       — generic: sort already-built (key,value) pairs on bits
//    [0,end_bit).  One library call (CUB onesweep, the kernel the reference itself would
//    instantiate for sm_100a); kept for `sort=True` on caller-provided keys and as the
//    fallback for inputs outside the contract (negative depths).
//
//  * b200splat_isect_sorted     — the B200 path used by rasterization().  An LSD radix
//    sort of `cam|tile|depth` first orders by the 32 depth bits, and those bits are a
//    property of the GAUSSIAN, not of the intersection.  So:
//      1. sort the C·N (or nnz) Gaussians by depth bits       (32-bit keys,  n elements)
//      2. scan tiles_per_gauss in that order                  (                n elements)
//      3. expand every Gaussian into its tiles, in that order (8 B out per intersection)
//      4. stable-sort the intersections by the cam|tile bits  (<= 2 digit passes over I)
//      5. assemble the 64-bit ids (depth bits gathered back)  (12 B out per intersection)
//    The result is bit-identical to sorting the full keys (a stable sort by the high bits
//    of a sequence already ordered by the low bits, and the expansion order equals the
//    reference's tie order), but the intersection-sized traffic drops from
//    ~150 B (6 onesweep passes of 12-byte pairs + histogram) to ~56 B per intersection.
#include "common.cuh"
#include "scan.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace b2s {

struct TileRect2 { uint32_t x0, y0, x1, y1; };

// identical to isect.cu's tile_rect (CS/isect_tiles.cu:60-70)
__device__ __forceinline__ TileRect2 tile_rect2(float mx, float my, float radius, float ts, uint32_t tw,
                                                uint32_t th) {
    const float tr = __fdividef(radius, ts), tx = __fdividef(mx, ts), ty = __fdividef(my, ts);
    TileRect2 r;
    r.x0 = min(__float2uint_rz(floorf(tx - tr)), tw);
    r.y0 = min(__float2uint_rz(floorf(ty - tr)), th);
    r.x1 = min(__float2uint_rz(ceilf(tx + tr)), tw);
    r.y1 = min(__float2uint_rz(ceilf(ty + tr)), th);
    return r;
}

// step 1 input: key = depth bits of visible elements, 0xFFFFFFFF for invisible ones
// (a visible depth can never be 0xFFFFFFFF: the sign bit is excluded by the caller).
__global__ void __launch_bounds__(kThreads)
depth_keys_kernel(uint64_t n_elems, const int32_t *__restrict__ tiles_per_gauss, const float *__restrict__ depths,
                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elems) return;
    keys[i] = tiles_per_gauss[i] > 0 ? (uint32_t)__float_as_int(depths[i]) : 0xFFFFFFFFu;
    vals[i] = (uint32_t)i;
}

// step 2 input: counts in depth order
__global__ void __launch_bounds__(kThreads)
gather_counts_kernel(uint64_t n_elems, const uint32_t *__restrict__ order, const int32_t *__restrict__ tiles_per_gauss,
                     int32_t *__restrict__ counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elems) return;
    counts[i] = tiles_per_gauss[order[i]];
}

// step 3: one warp expands 32 consecutive Gaussians (in depth order).  Their intersections
// occupy one contiguous output range, so the lanes are mapped to OUTPUT positions (lane l
// writes positions l, l + 32, ...: fully coalesced 128-byte stores) and each finds its
// Gaussian by a 5-step binary search over the 32 in-warp start offsets (warp shuffles).
// KeyT: uint16_t when cam|tile fits 16 bits (every BASELINE config) — the two digit passes then
// move 6-byte instead of 8-byte pairs — else uint32_t.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
expand_kernel(int packed, uint32_t N, uint64_t n_elems, const uint32_t *__restrict__ order,
              const int64_t *__restrict__ cum_sorted, const int64_t *__restrict__ camera_ids,
              const float *__restrict__ means2d, const int32_t *__restrict__ radii, float ts, uint32_t tw, uint32_t th,
              uint32_t tile_n_bits, KeyT *__restrict__ tile_keys, uint32_t *__restrict__ vals) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp * 32 >= n_elems) return;  // warp-uniform
    const uint64_t i = warp * 32 + lane;
    uint32_t idx = 0, xy0 = 0, w = 1, cnt = 0, cam_enc = 0;
    if (i < n_elems) {
        idx = order[i];
        const float radius = (float)radii[idx];
        if (radius > 0.f) {
            const float2 m = reinterpret_cast<const float2 *>(means2d)[idx];
            const TileRect2 r = tile_rect2(m.x, m.y, radius, ts, tw, th);
            w = r.x1 - r.x0;
            cnt = (r.y1 - r.y0) * w;
            xy0 = r.x0 | (r.y0 << 16);
            w = max(w, 1u);
            const uint32_t cid = packed ? (uint32_t)camera_ids[idx] : (uint32_t)(idx / N);
            cam_enc = cid << tile_n_bits;
        }
    }
    // cum_sorted is the inclusive scan of the same counts in the same order
    const int64_t end = cum_sorted[i < n_elems ? i : n_elems - 1];
    const int64_t warp_base = __shfl_sync(0xffffffffu, end - (int64_t)cnt, 0);
    const uint32_t total = (uint32_t)(__shfl_sync(0xffffffffu, end, 31) - warp_base);
    const uint32_t rel_start = (uint32_t)(end - (int64_t)cnt - warp_base);
    const float inv_w = 1.f / (float)w;
    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t p = base + lane;
        // largest j with rel_start[j] <= p: Gaussians without tiles share their successor's
        // start and are skipped; lanes past the end have rel_start == total > p
        int pos = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const uint32_t s_ = __shfl_sync(0xffffffffu, rel_start, pos + step);
            if (s_ <= p) pos += step;
        }
        const uint32_t g_idx = __shfl_sync(0xffffffffu, idx, pos), g_xy0 = __shfl_sync(0xffffffffu, xy0, pos);
        const uint32_t g_w = __shfl_sync(0xffffffffu, w, pos), g_rs = __shfl_sync(0xffffffffu, rel_start, pos);
        const uint32_t g_cam = __shfl_sync(0xffffffffu, cam_enc, pos);
        const float g_iw = __shfl_sync(0xffffffffu, inv_w, pos);
        if (p < total) {
            const uint32_t k = p - g_rs;
            uint32_t ry = __float2uint_rz(((float)k + 0.5f) * g_iw);  // k / g_w, fixed up below
            if (ry * g_w > k) --ry;
            if ((ry + 1) * g_w <= k) ++ry;
            const uint32_t rx = k - ry * g_w;
            tile_keys[warp_base + p] = (KeyT)(g_cam | (((g_xy0 >> 16) + ry) * tw + (g_xy0 & 0xffffu) + rx));
            vals[warp_base + p] = g_idx;
        }
    }
}

// step 5: 64-bit ids from the sorted (cam|tile, flat index) pairs, and — fused, optional —
// the per-tile offsets of a7 (same rule as offset_encode_kernel in isect.cu)
template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
assemble_kernel(uint64_t n_isects, const KeyT *__restrict__ tile_keys, const uint32_t *__restrict__ vals,
                const float *__restrict__ depths, int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids,
                uint32_t total_tiles, uint32_t n_tiles, uint32_t tile_n_bits, int32_t *__restrict__ offsets) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_isects) return;
    const uint32_t v = vals[i];
    const uint32_t key = tile_keys[i];
    const uint32_t d = (uint32_t)__float_as_int(__ldg(depths + v));
    isect_ids[i] = (int64_t)(((uint64_t)key << 32) | (uint64_t)d);
    flatten_ids[i] = (int32_t)v;
    if (offsets == nullptr) return;
    const uint32_t tile_mask = (1u << tile_n_bits) - 1u;
    const int64_t id_curr = (int64_t)(key >> tile_n_bits) * n_tiles + (key & tile_mask);
    if (i == 0)
        for (int64_t k = 0; k <= id_curr && k < total_tiles; ++k) offsets[k] = 0;
    if (i == n_isects - 1)
        for (int64_t k = id_curr + 1; k < total_tiles; ++k) offsets[k] = (int32_t)n_isects;
    if (i > 0) {
        const uint32_t prev = tile_keys[i - 1];
        if (prev == key) return;
        const int64_t id_prev = (int64_t)(prev >> tile_n_bits) * n_tiles + (prev & tile_mask);
        for (int64_t k = id_prev + 1; k <= id_curr && k < total_tiles; ++k) offsets[k] = (int32_t)i;
    }
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// workspace of phase 1 (depth order; depends on n_elems only) and of phase 2 (tile order)
struct DepthLayout {
    size_t gkeys_a, gkeys_b, gvals_a, gvals_b, counts, cum, total, scan_ws, scan_ws_bytes, cub, cub_bytes, end;
};
struct TileLayout {
    size_t tkeys_a, tkeys_b, tvals_a, tvals_b, cub, cub_bytes, end;
};

static DepthLayout depth_layout(uint64_t n_elems) {
    DepthLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes); return r; };
    L.gkeys_a = take(4 * n_elems); L.gkeys_b = take(4 * n_elems);
    L.gvals_a = take(4 * n_elems); L.gvals_b = take(4 * n_elems);
    L.counts = take(4 * n_elems);
    L.cum = take(8 * n_elems);
    L.total = take(8);
    L.scan_ws_bytes = scan_workspace_bytes(n_elems);
    L.scan_ws = take(L.scan_ws_bytes);
    size_t b1 = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b1, k, v, (int64_t)(n_elems ? n_elems : 1), 0, 32, (cudaStream_t)0);
    L.cub_bytes = b1 + 256;
    L.cub = take(L.cub_bytes);
    L.end = o;
    return L;
}

// sized for 32-bit keys (the 16-bit variant needs less)
static TileLayout tile_layout(uint64_t n_isects) {
    TileLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes); return r; };
    L.tkeys_a = take(4 * n_isects); L.tkeys_b = take(4 * n_isects);
    L.tvals_a = take(4 * n_isects); L.tvals_b = take(4 * n_isects);
    size_t b2 = 0, b3 = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b2, k, v, (int64_t)(n_isects ? n_isects : 1), 0, 32, (cudaStream_t)0);
    cub::DoubleBuffer<uint16_t> k16(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b3, k16, v, (int64_t)(n_isects ? n_isects : 1), 0, 16, (cudaStream_t)0);
    if (b3 > b2) b2 = b3;
    L.cub_bytes = b2 + 256;
    L.cub = take(L.cub_bytes);
    L.end = o;
    return L;
}

}  // namespace b2s

using namespace b2s;

extern "C" size_t b200splat_sort_workspace_bytes(uint64_t n_isects) {
    if (n_isects == 0) return 0;
    cub::DoubleBuffer<int64_t> keys(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> vals(nullptr, nullptr);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, vals, (int64_t)n_isects, 0, 64, (cudaStream_t)0);
    return bytes + 256;
}

// keys_a/vals_a hold the input and are clobbered; the sorted result lands in buffer
// `*selector_out` (0 = *_a, 1 = *_b), exactly like cub::DoubleBuffer in the reference.
extern "C" int b200splat_isect_sort(uint64_t n_isects, uint32_t end_bit, int64_t *keys_a, int32_t *vals_a,
                                    int64_t *keys_b, int32_t *vals_b, void *workspace, size_t workspace_bytes,
                                    int *selector_out, void *stream) {
    const char *where = "b200splat_isect_sort";
    B2S_REQUIRE(end_bit <= 64, where, "end_bit must be <= 64");
    B2S_REQUIRE(selector_out != nullptr, where, "selector_out is required");
    *selector_out = 0;
    if (n_isects == 0) return 0;
    cub::DoubleBuffer<int64_t> keys(keys_a, keys_b);
    cub::DoubleBuffer<int32_t> vals(vals_a, vals_b);
    size_t bytes = workspace_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(workspace, bytes, keys, vals, (int64_t)n_isects, 0, (int)end_bit,
                                                    (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(where, e);
    *selector_out = keys.selector;
    return 0;
}

// ---- depth-first ordering, phase 1: everything that does not need n_isects -----------------
// (launched by the wrapper BEFORE it reads n_isects back, so the host round trip of that one
// sync is hidden behind ~0.1 ms of useful device work)
extern "C" size_t b200splat_isect_depth_order_workspace_bytes(uint64_t n_elems) {
    return depth_layout(n_elems).end;
}

extern "C" int b200splat_isect_depth_order(uint64_t n_elems, const float *depths, const int32_t *tiles_per_gauss,
                                           void *workspace, size_t workspace_bytes, int *selector_out, void *stream) {
    const char *where = "b200splat_isect_depth_order";
    cudaStream_t st = (cudaStream_t)stream;
    B2S_REQUIRE(n_elems <= 0xffffffffull, where, "more than 2^32 (camera, Gaussian) pairs");
    B2S_REQUIRE(selector_out != nullptr, where, "selector_out is required");
    *selector_out = 0;
    if (n_elems == 0) return 0;
    const DepthLayout L = depth_layout(n_elems);
    B2S_REQUIRE(workspace != nullptr && workspace_bytes >= L.end, where,
                "workspace too small (see b200splat_isect_depth_order_workspace_bytes)");
    char *ws = reinterpret_cast<char *>(workspace);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(ws + off); };
    // 1. Gaussians by depth
    depth_keys_kernel<<<div_up(n_elems, kThreads), kThreads, 0, st>>>(n_elems, tiles_per_gauss, depths, u32(L.gkeys_a),
                                                                      u32(L.gvals_a));
    B2S_CHECK_LAUNCH(where);
    cub::DoubleBuffer<uint32_t> gk(u32(L.gkeys_a), u32(L.gkeys_b)), gv(u32(L.gvals_a), u32(L.gvals_b));
    size_t cb = L.cub_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(ws + L.cub, cb, gk, gv, (int64_t)n_elems, 0, 32, st);
    if (e != cudaSuccess) return fail_cuda(where, e);
    *selector_out = gv.selector;
    const uint32_t *order = gv.Current();
    // 2. offsets in depth order
    int32_t *counts = reinterpret_cast<int32_t *>(ws + L.counts);
    int64_t *cum = reinterpret_cast<int64_t *>(ws + L.cum);
    gather_counts_kernel<<<div_up(n_elems, kThreads), kThreads, 0, st>>>(n_elems, order, tiles_per_gauss, counts);
    B2S_CHECK_LAUNCH(where);
    if (lookback_scan_i32_to_i64(counts, cum, n_elems, reinterpret_cast<int64_t *>(ws + L.total), ws + L.scan_ws,
                                 L.scan_ws_bytes, st))
        return fail(where, "scan failed");
    return 0;
}

template <typename KeyT>
static int tile_order_run(int packed, uint32_t C, uint32_t N, uint64_t n_elems, uint64_t n_isects,
                          const int64_t *camera_ids, const float *means2d, const int32_t *radii, const float *depths,
                          const uint32_t *order, const int64_t *cum, uint32_t tile_size, uint32_t tile_width,
                          uint32_t tile_height, uint32_t tile_n_bits, uint32_t cam_n_bits, int64_t *isect_ids,
                          int32_t *flatten_ids, int32_t *offsets, char *ws, const TileLayout &L, cudaStream_t st,
                          const char *where) {
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(ws + off); };
    KeyT *ka = reinterpret_cast<KeyT *>(ws + L.tkeys_a), *kb = reinterpret_cast<KeyT *>(ws + L.tkeys_b);
    const uint32_t n_tiles = tile_width * tile_height;
    // 3. expand
    expand_kernel<KeyT><<<div_up(n_elems, kThreads), kThreads, 0, st>>>(packed, N, n_elems, order, cum, camera_ids,
                                                                        means2d, radii, (float)tile_size, tile_width,
                                                                        tile_height, tile_n_bits, ka, u32(L.tvals_a));
    B2S_CHECK_LAUNCH(where);
    // 4. stable sort by cam|tile
    cub::DoubleBuffer<KeyT> tk(ka, kb);
    cub::DoubleBuffer<uint32_t> tv(u32(L.tvals_a), u32(L.tvals_b));
    size_t cb = L.cub_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(ws + L.cub, cb, tk, tv, (int64_t)n_isects, 0,
                                                    (int)(tile_n_bits + cam_n_bits), st);
    if (e != cudaSuccess) return fail_cuda(where, e);
    // 5. assemble
    assemble_kernel<KeyT><<<div_up(n_isects, kThreads), kThreads, 0, st>>>(n_isects, tk.Current(), tv.Current(), depths,
                                                                           isect_ids, flatten_ids, C * n_tiles, n_tiles,
                                                                           tile_n_bits, offsets);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

// ---- phase 2: expand, stable tile sort, assemble (+ offsets) ---------------------------------
extern "C" size_t b200splat_isect_tile_order_workspace_bytes(uint64_t n_isects) {
    return tile_layout(n_isects).end;
}

extern "C" int b200splat_isect_tile_order(int packed, uint32_t C, uint32_t N, uint32_t nnz, const int64_t *camera_ids,
                                          const float *means2d, const int32_t *radii, const float *depths,
                                          const void *depth_workspace, int depth_selector, uint64_t n_isects,
                                          uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                                          int64_t *isect_ids, int32_t *flatten_ids, int32_t *offsets, void *workspace,
                                          size_t workspace_bytes, void *stream) {
    const char *where = "b200splat_isect_tile_order";
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n_elems = packed ? (uint64_t)nnz : (uint64_t)C * N;
    B2S_REQUIRE(!packed || camera_ids != nullptr, where, "camera_ids required when packed");
    B2S_REQUIRE(n_elems <= 0xffffffffull, where, "more than 2^32 (camera, Gaussian) pairs");
    const uint32_t n_tiles = tile_width * tile_height;
    uint32_t tile_n_bits = 0, cam_n_bits = 0;
    for (uint32_t v = n_tiles; v; v >>= 1) ++tile_n_bits;
    for (uint32_t v = C; v; v >>= 1) ++cam_n_bits;
    B2S_REQUIRE(tile_n_bits + cam_n_bits <= 32, where, "camera and tile ids do not fit in 32 bits");
    if (n_isects == 0 || n_elems == 0) return 0;
    const DepthLayout D = depth_layout(n_elems);
    const TileLayout L = tile_layout(n_isects);
    B2S_REQUIRE(depth_workspace != nullptr, where, "depth_workspace (from b200splat_isect_depth_order) is required");
    B2S_REQUIRE(workspace != nullptr && workspace_bytes >= L.end, where,
                "workspace too small (see b200splat_isect_tile_order_workspace_bytes)");
    const char *dws = reinterpret_cast<const char *>(depth_workspace);
    char *ws = reinterpret_cast<char *>(workspace);
    const uint32_t *order = reinterpret_cast<const uint32_t *>(dws + (depth_selector ? D.gvals_b : D.gvals_a));
    const int64_t *cum = reinterpret_cast<const int64_t *>(dws + D.cum);
    const int rc = (tile_n_bits + cam_n_bits <= 16 && tuning_variant() != 8)
                       ? tile_order_run<uint16_t>(packed, C, N, n_elems, n_isects, camera_ids, means2d, radii, depths, order,
                                                  cum, tile_size, tile_width, tile_height, tile_n_bits, cam_n_bits,
                                                  isect_ids, flatten_ids, offsets, ws, L, st, where)
                       : tile_order_run<uint32_t>(packed, C, N, n_elems, n_isects, camera_ids, means2d, radii, depths, order,
                                                  cum, tile_size, tile_width, tile_height, tile_n_bits, cam_n_bits,
                                                  isect_ids, flatten_ids, offsets, ws, L, st, where);
    return rc;
}

// both phases in one call (workspace = phase-1 layout followed by phase-2 layout)
extern "C" size_t b200splat_isect_sorted_workspace_bytes(uint64_t n_elems, uint64_t n_isects) {
    return depth_layout(n_elems).end + tile_layout(n_isects).end;
}

extern "C" int b200splat_isect_sorted(int packed, uint32_t C, uint32_t N, uint32_t nnz, const int64_t *camera_ids,
                                      const float *means2d, const int32_t *radii, const float *depths,
                                      const int32_t *tiles_per_gauss, uint64_t n_isects, uint32_t tile_size,
                                      uint32_t tile_width, uint32_t tile_height, int64_t *isect_ids,
                                      int32_t *flatten_ids, int32_t *offsets, void *workspace, size_t workspace_bytes,
                                      void *stream) {
    const char *where = "b200splat_isect_sorted";
    const uint64_t n_elems = packed ? (uint64_t)nnz : (uint64_t)C * N;
    B2S_REQUIRE(n_elems <= 0xffffffffull, where, "more than 2^32 (camera, Gaussian) pairs");
    if (n_isects == 0 || n_elems == 0) return 0;
    const size_t d_bytes = depth_layout(n_elems).end, t_bytes = tile_layout(n_isects).end;
    B2S_REQUIRE(workspace != nullptr && workspace_bytes >= d_bytes + t_bytes, where,
                "workspace too small (see b200splat_isect_sorted_workspace_bytes)");
    char *ws = reinterpret_cast<char *>(workspace);
    int sel = 0;
    int rc = b200splat_isect_depth_order(n_elems, depths, tiles_per_gauss, ws, d_bytes, &sel, stream);
    if (rc) return rc;
    return b200splat_isect_tile_order(packed, C, N, nnz, camera_ids, means2d, radii, depths, ws, sel, n_isects,
                                      tile_size, tile_width, tile_height, isect_ids, flatten_ids, offsets, ws + d_bytes,
                                      t_bytes, stream);
}
