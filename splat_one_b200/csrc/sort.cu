// sort.cu — ordering stage of isect_tiles (a6): produce the (isect_ids, flatten_ids) arrays
// sorted exactly like the reference's stable LSD radix sort of the 64-bit keys
// `cam | tile | depth bits` (CS/isect_tiles.cu:252-300, a cub::DeviceRadixSort::SortPairs
// call there).  Everything here is this library's own code: the digit passes are a
// hand-written single-pass ("onesweep") radix scatter with decoupled look-back, no CUB.
//
// Two implementations behind the C ABI:
//
//  * b200splat_isect_sort       — generic: sort already-built (key,value) pairs on bits
//    [0,end_bit).  ceil(end_bit/8) onesweep passes over 64-bit keys; kept for `sort=True` on
//    caller-provided keys and as the fallback for inputs outside the contract (negative
//    depths).
//
//  * b200splat_isect_sorted     — the B200 path used by rasterization().  An LSD radix
//    sort of `cam|tile|depth` first orders by the 32 depth bits, and those bits are a
//    property of the GAUSSIAN, not of the intersection.  So:
//      1. sort the C·N (or nnz) Gaussians by depth bits: depth_keys (+ the four digit
//         histograms, fused) and four 8-bit passes over 32-bit pairs; the LAST pass drops
//         the keys and gathers tiles_per_gauss into depth order instead   (n elements)
//      2. scan those counts                                               (n elements)
//      3. expand every Gaussian into its tiles, in that order, and histogram the digits
//         of the cam|tile keys in the same kernel              (6 B out per intersection)
//      4. stable-sort the intersections by the cam|tile bits: <= 2 passes for <= 16 bits;
//         the LAST pass writes the final 64-bit ids (depth bits gathered back) and
//         flatten_ids directly — there is no separate assemble pass
//      5. per-tile offsets (a7) by binary search over the sorted ids (C·tiles threads)
//    The result is bit-identical to sorting the full keys (a stable sort by the high bits
//    of a sequence already ordered by the low bits, and the expansion order equals the
//    reference's tie order); intersection-sized traffic is ~36 B per intersection
//    (expand 6 out, pass 1: 6 in + 6 out, pass 2: 6 in + 12 out) against ~150 B for six
//    onesweep passes of 12-byte pairs plus the histogram read.
//
// Onesweep pass (onesweep_kernel): a block takes a tile of TILE consecutive input
// elements (tile ids come from an atomic ticket, so every predecessor of a running block
// is itself running or done), ranks them per warp with match.any (order preserving),
// combines the warp histograms, publishes its per-digit totals and obtains the totals of
// all earlier tiles by decoupled look-back (partial / inclusive flags, one 64-bit word
// per (tile, digit)), stages the tile in shared memory in digit order and writes it out
// in runs that are contiguous per digit.  The global digit histograms are produced by the
// kernels that generate the keys, so a pass reads the data exactly once.
#include "common.cuh"
#include "scan.cuh"

namespace b2s {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kBins = 256;          // 8-bit digits at most
constexpr int kMaxPasses = 8;       // 64-bit keys
constexpr unsigned long long kFlagPartial = 1ull << 62, kFlagInclusive = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

enum SortMode { kModePairs = 0, kModeDepthFinal = 1, kModeTileFinal = 2 };

// digit p of a key = (key >> shift[p]) & mask[p]
struct DigitPlan {
    uint32_t npass;
    uint32_t shift[kMaxPasses];
    uint32_t mask[kMaxPasses];
};

// `bits` key bits in ceil(bits/8) passes of (nearly) equal width: narrower digits mean fewer,
// longer runs per tile in the scatter
static DigitPlan make_plan(uint32_t bits) {
    DigitPlan p;
    p.npass = bits == 0 ? 1 : (bits + 7) / 8;
    uint32_t done = 0;
    for (uint32_t i = 0; i < kMaxPasses; ++i) {
        p.shift[i] = 0;
        p.mask[i] = 0;
    }
    for (uint32_t i = 0; i < p.npass; ++i) {
        const uint32_t left = bits - done, w = bits == 0 ? 1 : (left + (p.npass - i) - 1) / (p.npass - i);
        p.shift[i] = done;
        p.mask[i] = (1u << w) - 1u;
        done += w;
    }
    return p;
}

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

// Increment of the block's digit histograms (shared memory, [NPASS][256]) by the lanes in `active`
// (all of them must call).  Lanes that share the digit of the first active lane are counted by
// that lane in one add (the high digits of consecutive tiles / of depth exponents are mostly
// equal across a warp); the others add individually (low digits are mostly distinct).
template <int NPASS, typename KeyT>
__device__ __forceinline__ void hist_add(uint32_t *s_hist, const uint32_t (&shift)[NPASS], const uint32_t (&mask)[NPASS],
                                         KeyT key, unsigned active, unsigned lane) {
    const int first = __ffs(active) - 1;
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
        const uint32_t d = (uint32_t)(key >> shift[p]) & mask[p];
        const uint32_t d0 = __shfl_sync(active, d, first);
        const unsigned same = __ballot_sync(active, d == d0);
        if ((int)lane == first) atomicAdd(&s_hist[p * kBins + d], (uint32_t)__popc(same));
        else if (d != d0) atomicAdd(&s_hist[p * kBins + d], 1u);
    }
}

__device__ __forceinline__ void hist_flush(const uint32_t *s_hist, uint32_t npass, uint32_t *__restrict__ hist) {
    for (uint32_t i = threadIdx.x; i < npass * kBins; i += blockDim.x) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// step 1 input: key = depth bits of visible elements, 0xFFFFFFFF for invisible ones
// (a visible depth can never be 0xFFFFFFFF: the sign bit is excluded by the caller);
// + the four digit histograms of those keys.
__global__ void __launch_bounds__(kThreads)
depth_keys_kernel(uint32_t n_elems, const int32_t *__restrict__ tiles_per_gauss, const float *__restrict__ depths,
                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, const __grid_constant__ DigitPlan plan,
                  uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_hist[4 * kBins];
    for (uint32_t i = threadIdx.x; i < 4 * kBins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t shift[4] = {plan.shift[0], plan.shift[1], plan.shift[2], plan.shift[3]};
    const uint32_t mask[4] = {plan.mask[0], plan.mask[1], plan.mask[2], plan.mask[3]};
    const unsigned lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (n_elems + stride - 1) / stride;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint64_t i = (uint64_t)r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = i < n_elems;
        const unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const uint32_t key = tiles_per_gauss[i] > 0 ? (uint32_t)__float_as_int(depths[i]) : 0xFFFFFFFFu;
            keys[i] = key;
            vals[i] = (uint32_t)i;
            hist_add<4, uint32_t>(s_hist, shift, mask, key, active, lane);
        }
    }
    __syncthreads();
    hist_flush(s_hist, 4, hist);
}

// digit histograms of caller-provided keys (generic 64-bit sort)
template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
radix_hist_kernel(uint32_t n, const KeyT *__restrict__ keys, const __grid_constant__ DigitPlan plan,
                  uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_hist[kMaxPasses * kBins];
    for (uint32_t i = threadIdx.x; i < kMaxPasses * kBins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint32_t shift[kMaxPasses], mask[kMaxPasses];  // unused passes: mask 0 -> bin 0, never read
#pragma unroll
    for (int p = 0; p < kMaxPasses; ++p) { shift[p] = plan.shift[p]; mask[p] = plan.mask[p]; }
    const unsigned lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (n + stride - 1) / stride;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint64_t i = (uint64_t)r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = i < n;
        const unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) hist_add<kMaxPasses, KeyT>(s_hist, shift, mask, keys[i], active, lane);
    }
    __syncthreads();
    hist_flush(s_hist, plan.npass, hist);
}

// step 3: one warp expands 32 consecutive Gaussians (in depth order).  Their intersections
// occupy one contiguous output range, so the lanes are mapped to OUTPUT positions (lane l
// writes positions l, l + 32, ...: fully coalesced 128-byte stores) and each finds its
// Gaussian by a 5-step binary search over the 32 in-warp start offsets (warp shuffles).
// The digit histograms of the tile sort are accumulated on the way (no second read of the keys).
// KeyT: uint16_t when cam|tile fits 16 bits (every BASELINE config) — the digit passes then
// move 6-byte instead of 8-byte pairs — else uint32_t.
template <typename KeyT, int NPASS>
__global__ void __launch_bounds__(kThreads)
expand_kernel(int packed, uint32_t N, uint32_t n_elems, const uint32_t *__restrict__ order,
              const int64_t *__restrict__ cum_sorted, const int64_t *__restrict__ camera_ids,
              const float *__restrict__ means2d, const int32_t *__restrict__ radii, float ts, uint32_t tw, uint32_t th,
              uint32_t tile_n_bits, KeyT *__restrict__ tile_keys, uint32_t *__restrict__ vals,
              const __grid_constant__ DigitPlan plan, uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_hist[NPASS * kBins];
    for (uint32_t i = threadIdx.x; i < NPASS * kBins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint32_t shift[NPASS], mask[NPASS];
#pragma unroll
    for (int p = 0; p < NPASS; ++p) { shift[p] = plan.shift[p]; mask[p] = plan.mask[p]; }
    const unsigned lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp * 32 < n_elems) {  // warp-uniform
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
            const unsigned active = __ballot_sync(0xffffffffu, p < total);
            if (p < total) {
                const uint32_t k = p - g_rs;
                uint32_t ry = __float2uint_rz(((float)k + 0.5f) * g_iw);  // k / g_w, fixed up below
                if (ry * g_w > k) --ry;
                if ((ry + 1) * g_w <= k) ++ry;
                const uint32_t rx = k - ry * g_w;
                const uint32_t key = g_cam | (((g_xy0 >> 16) + ry) * tw + (g_xy0 & 0xffffu) + rx);
                tile_keys[warp_base + p] = (KeyT)key;
                vals[warp_base + p] = g_idx;
                hist_add<NPASS, uint32_t>(s_hist, shift, mask, key, active, lane);
            }
        }
    }
    __syncthreads();
    hist_flush(s_hist, NPASS, hist);
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <typename KeyT, int ITEMS>
constexpr size_t onesweep_smem_bytes() {
    constexpr size_t head = (2 * kBins + 64) * sizeof(uint32_t);
    constexpr size_t hist = 2 * (size_t)kSortWarps * kBins * sizeof(uint32_t);  // warp histograms + match masks
    constexpr size_t stage = (size_t)kSortThreads * ITEMS * (sizeof(KeyT) + sizeof(uint32_t));
    return head + (hist > stage ? hist : stage);
}

// Order-preserving rank of every element of a warp's slice among the slice's elements with the
// same digit: rank[i] = #earlier elements (items 0..i-1 of any lane, item i of lower lanes) with
// key digit == this one's.  Lanes with equal digits find each other through a per-warp table of
// lane masks in shared memory (atomicOr, read back, cleared by the group's first lane); the group's
// first lane advances the warp histogram.  (The hardware MATCH.ANY instruction iterates over the
// distinct values of the warp — measured ~150-200 cycles per call on random digits — and a ballot
// per digit bit costs ~4 instructions per bit.)  A warp whose lanes all hold the same digit — the
// top byte of depth keys, runs of one tile — skips the table.
template <bool FULL, int ITEMS, typename KeyT>
__device__ __forceinline__ void rank_slice(const KeyT (&key)[ITEMS], uint32_t (&rank)[ITEMS], uint32_t shift,
                                           uint32_t mask, uint32_t *my_hist, uint32_t *my_match, uint32_t n_valid,
                                           unsigned lane) {
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const bool valid = FULL || (uint32_t)(i * 32) + lane < n_valid;
        const unsigned vmask = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
        if (!FULL && vmask == 0u) break;  // warp-uniform
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        const uint32_t d0 = __shfl_sync(0xffffffffu, d, __ffs(vmask) - 1);
        unsigned m;
        const bool uniform = __all_sync(0xffffffffu, !valid || d == d0);
        if (uniform) {
            m = valid ? vmask : 0u;
        } else {
            if (valid) atomicOr(&my_match[d], 1u << lane);
            __syncwarp();
            m = valid ? my_match[d] : 0u;
            __syncwarp();
        }
        const int leader = __ffs(m) - 1;
        uint32_t old = 0;
        if (valid && (int)lane == leader) {
            old = my_hist[d];
            my_hist[d] = old + (uint32_t)__popc(m);
            if (!uniform) my_match[d] = 0u;
        }
        old = __shfl_sync(0xffffffffu, old, leader < 0 ? 0 : leader);
        rank[i] = old + (uint32_t)__popc(m & lt);
        __syncwarp();
    }
}

// One digit pass.  MODE selects what the scatter writes:
//   kModePairs      (key, value) pairs into keys_out / vals_out
//   kModeDepthFinal values into vals_out and tiles_per_gauss[value] (gather_src) into aux_out (int32):
//                   the depth order itself and the tile counts in that order; keys are dropped
//   kModeTileFinal  the final a6 outputs: isect_ids (aux_out, int64) = key << 32 | depth bits of the
//                   value's Gaussian (gather_src), flatten_ids (vals_out) = value
template <typename KeyT, int MODE, int ITEMS>
__global__ void __launch_bounds__(kSortThreads, ITEMS >= 16 ? 3 : 4)
onesweep_kernel(const KeyT *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, KeyT *__restrict__ keys_out,
                uint32_t *__restrict__ vals_out, uint32_t n, uint32_t shift, uint32_t mask,
                const uint32_t *__restrict__ hist, unsigned long long *__restrict__ lookback,
                uint32_t *__restrict__ ticket, const uint32_t *__restrict__ gather_src, void *__restrict__ aux_out) {
    constexpr int TILE = kSortThreads * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_bbs = reinterpret_cast<uint32_t *>(smem_raw);  // [256] start of each digit inside this tile
    uint32_t *s_gbase = s_bbs + kBins;                          // [256] global start of the digit's run - s_bbs
    uint32_t *s_misc = s_gbase + kBins;                         // [64] scan scratch, tile id
    unsigned char *s_alias = reinterpret_cast<unsigned char *>(s_misc + 64);
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_alias);   // [warps][256]; dead once the ranks are final
    uint32_t *s_match = s_hist + kSortWarps * kBins;            // [warps][256] lane masks; dead after ranking
    KeyT *s_keys = reinterpret_cast<KeyT *>(s_alias);           // [TILE] staging, aliases s_hist
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_alias + sizeof(KeyT) * TILE);

    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_misc[63] = atomicAdd(ticket, 1u);
#pragma unroll
    for (int i = 0; i < 2 * kSortWarps; ++i) s_hist[i * kBins + tid] = 0;  // histograms and match masks
    __syncthreads();
    const uint32_t tile = s_misc[63];
    const uint32_t base = tile * (uint32_t)TILE;
    const uint32_t count = min((uint32_t)TILE, n - base);
    const uint32_t wbase = wid * (32u * ITEMS);  // this warp's slice of the tile, lane-strided

    KeyT key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rank[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = wbase + i * 32 + lane;
        key[i] = j < count ? keys_in[base + j] : (KeyT)0;
    }
    const uint32_t n_valid = count > wbase ? count - wbase : 0u;  // valid elements of this warp's slice
    if (count == (uint32_t)TILE)
        rank_slice<true, ITEMS, KeyT>(key, rank, shift, mask, s_hist + wid * kBins, s_match + wid * kBins, n_valid, lane);
    else
        rank_slice<false, ITEMS, KeyT>(key, rank, shift, mask, s_hist + wid * kBins, s_match + wid * kBins, n_valid, lane);
    // the values are not needed before the staging step: their loads overlap the look-back
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = wbase + i * 32 + lane;
        val[i] = j < count ? vals_in[base + j] : 0u;
    }
    __syncthreads();

    // bins: thread t owns digit t.  Warp histograms -> exclusive over warps; tile total; look-back.
    {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = s_hist[w * kBins + tid];
            s_hist[w * kBins + tid] = total;
            total += c;
        }
        unsigned long long *state = lookback + (size_t)tile * kBins + tid;
        st_relaxed_u64(state, (tile == 0 ? kFlagInclusive : kFlagPartial) | (unsigned long long)total);
        // exclusive scans over the digits of (global histogram, tile totals), packed in one word
        unsigned long long x = ((unsigned long long)hist[tid] << 32) | (unsigned long long)total;
        const unsigned long long own = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += y;
        }
        unsigned long long *s_wtot = reinterpret_cast<unsigned long long *>(s_misc);  // [8]
        if (lane == 31) s_wtot[wid] = x;
        __syncthreads();
        unsigned long long pre = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w)
            if (w < (int)wid) pre += s_wtot[w];
        const unsigned long long excl_scan = pre + x - own;
        const uint32_t gstart = (uint32_t)(excl_scan >> 32), bstart = (uint32_t)excl_scan;
        // decoupled look-back: totals of this digit in all earlier tiles.  kProbe predecessors are
        // fetched per round with independent loads (all tiles of a wave publish their partial totals at
        // about the same time, so the walk is a chain of dependent L2 round trips until an inclusive
        // total turns up: probing 4 at a time makes that frontier advance 4x faster)
        uint32_t excl = 0;
        if (tile > 0) {
            constexpr int kProbe = 4;
            int64_t j = (int64_t)tile - 1;
            bool found = false;
            while (!found) {
                unsigned long long sw[kProbe];
#pragma unroll
                for (int w = 0; w < kProbe; ++w)
                    sw[w] = (j - w >= 0) ? ld_relaxed_u64(lookback + (size_t)(j - w) * kBins + tid) : kFlagInclusive;
#pragma unroll
                for (int w = 0; w < kProbe; ++w) {
                    if (!found) {
                        unsigned long long sv = sw[w];
                        while ((sv >> 62) == 0ull) sv = ld_relaxed_u64(lookback + (size_t)(j - w) * kBins + tid);
                        excl += (uint32_t)(sv & kValueMask);
                        found = (sv >> 62) == 2ull;
                    }
                }
                j -= kProbe;
            }
            st_relaxed_u64(state, kFlagInclusive | (unsigned long long)(excl + total));
        }
        s_bbs[tid] = bstart;
        s_gbase[tid] = gstart + excl - bstart;
    }
    __syncthreads();
    // final position inside the tile (digit order, stable)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (wbase + i * 32 + lane < count) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            rank[i] += s_bbs[d] + s_hist[wid * kBins + d];
        }
    }
    __syncthreads();  // s_hist is dead: the staging buffers alias it
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (wbase + i * 32 + lane < count) {
            s_keys[rank[i]] = key[i];
            s_vals[rank[i]] = val[i];
        }
    }
    __syncthreads();
    // write-out: thread t takes staged elements t, t + 256, ...: runs that are contiguous per digit
    auto emit = [&](const KeyT k, const uint32_t v, const uint32_t g, const uint32_t dst) {
        if (MODE == kModePairs) {
            keys_out[dst] = k;
            vals_out[dst] = v;
        } else if (MODE == kModeDepthFinal) {
            vals_out[dst] = v;
            reinterpret_cast<uint32_t *>(aux_out)[dst] = g;
        } else {
            reinterpret_cast<unsigned long long *>(aux_out)[dst] = ((unsigned long long)k << 32) | (unsigned long long)g;
            vals_out[dst] = v;
        }
    };
    if (count == (uint32_t)TILE) {
        // full tile: all staged elements and (final passes) their gathers in flight at once
        KeyT k[ITEMS];
        uint32_t g[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            k[i] = s_keys[tid + i * kSortThreads];
            val[i] = s_vals[tid + i * kSortThreads];
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) g[i] = MODE == kModePairs ? 0u : __ldg(gather_src + val[i]);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            emit(k[i], val[i], g[i], s_gbase[(uint32_t)(k[i] >> shift) & mask] + tid + i * kSortThreads);
    } else {
        for (uint32_t j = tid; j < count; j += kSortThreads) {
            const KeyT k = s_keys[j];
            const uint32_t v = s_vals[j];
            emit(k, v, MODE == kModePairs ? 0u : __ldg(gather_src + v), s_gbase[(uint32_t)(k >> shift) & mask] + j);
        }
    }
}

// a7 from the sorted ids: offsets[k] = first sorted index whose cam|tile >= k (CS/isect_tiles.cu:309-355:
// run starts, with empty tiles inheriting the next start and the tail filled with n_isects).
// One thread per tile; an 8-ary search (7 independent probes per round) keeps the chain of
// dependent loads at ceil(log8 n) rounds.
__global__ void __launch_bounds__(kThreads)
offsets_search_kernel(uint32_t n_isects, const int64_t *__restrict__ isect_ids, uint32_t total_tiles, uint32_t n_tiles,
                      uint32_t tile_n_bits, int32_t *__restrict__ offsets) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total_tiles) return;
    const uint32_t cam = k / n_tiles, tile = k - cam * n_tiles;
    const uint64_t want = ((uint64_t)cam << tile_n_bits) | tile;
    uint32_t lo = 0, hi = n_isects;  // invariant: ids[< lo] < want, ids[>= hi] >= want
    while (hi - lo > 8) {
        const uint32_t step = (hi - lo) >> 3;
        uint64_t probe[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) probe[q] = (uint64_t)__ldg(isect_ids + lo + (q + 1) * step) >> 32;
        uint32_t nlo = lo, nhi = hi;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const uint32_t pos = lo + (q + 1) * step;
            if (probe[q] < want) nlo = pos + 1;
            else if (nhi == hi) nhi = pos;
        }
        lo = nlo;
        hi = nhi;
    }
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (((uint64_t)__ldg(isect_ids + mid) >> 32) < want) lo = mid + 1;
        else hi = mid;
    }
    offsets[k] = (int32_t)lo;
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static inline uint32_t sort_items(uint64_t n) {
    static const int forced = [] { const char *e = getenv("B200SPLAT_SORT_ITEMS"); return e ? atoi(e) : 0; }();  // dev A/B
    if (forced == 8 || forced == 16) return (uint32_t)forced;
    return n <= (1ull << 21) ? 8u : 16u;
}
static inline uint32_t sort_tiles(uint64_t n) {
    const uint64_t t = (uint64_t)kSortThreads * sort_items(n);
    return (uint32_t)((n + t - 1) / t);
}

// zero-initialised control block of a sort: digit histograms, tile tickets, look-back words
struct SortCtl {
    size_t hist, ticket, lookback, per_pass, bytes;
};
static SortCtl sort_ctl(uint64_t n, uint32_t max_passes) {
    SortCtl c;
    c.hist = 0;
    c.ticket = align_up((size_t)kMaxPasses * kBins * sizeof(uint32_t));
    c.lookback = c.ticket + align_up(kMaxPasses * sizeof(uint32_t));
    c.per_pass = align_up((size_t)sort_tiles(n) * kBins * sizeof(unsigned long long));
    c.bytes = c.lookback + c.per_pass * max_passes;
    return c;
}

template <typename KeyT, int MODE, int ITEMS>
static cudaError_t launch_onesweep_t(const KeyT *keys_in, const uint32_t *vals_in, KeyT *keys_out, uint32_t *vals_out,
                                     uint32_t n, uint32_t shift, uint32_t mask, const uint32_t *hist,
                                     unsigned long long *lookback, uint32_t *ticket, const uint32_t *gather_src,
                                     void *aux_out, cudaStream_t st) {
    auto kern = onesweep_kernel<KeyT, MODE, ITEMS>;
    constexpr size_t smem = onesweep_smem_bytes<KeyT, ITEMS>();
    if (smem > 48 * 1024) {  // per device, so not cached in a static
        const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (attr != cudaSuccess) return attr;
    }
    const uint32_t tiles = (n + kSortThreads * ITEMS - 1) / (kSortThreads * ITEMS);
    kern<<<tiles, kSortThreads, smem, st>>>(keys_in, vals_in, keys_out, vals_out, n, shift, mask, hist, lookback,
                                            ticket, gather_src, aux_out);
    return cudaGetLastError();
}

template <typename KeyT, int MODE>
static cudaError_t launch_onesweep(const KeyT *keys_in, const uint32_t *vals_in, KeyT *keys_out, uint32_t *vals_out,
                                   uint32_t n, uint32_t shift, uint32_t mask, const uint32_t *hist,
                                   unsigned long long *lookback, uint32_t *ticket, const uint32_t *gather_src,
                                   void *aux_out, cudaStream_t st) {
    if (sort_items(n) == 8)
        return launch_onesweep_t<KeyT, MODE, 8>(keys_in, vals_in, keys_out, vals_out, n, shift, mask, hist, lookback,
                                                ticket, gather_src, aux_out, st);
    return launch_onesweep_t<KeyT, MODE, 16>(keys_in, vals_in, keys_out, vals_out, n, shift, mask, hist, lookback,
                                             ticket, gather_src, aux_out, st);
}

// workspace of phase 1 (depth order; depends on n_elems only) and of phase 2 (tile order)
struct DepthLayout {
    size_t gkeys_a, gkeys_b, gvals_a, gvals_b, counts, cum, total, zeroed, scan_ws, scan_ws_bytes, ctl, zeroed_bytes, end;
    SortCtl c;
};
struct TileLayout {
    size_t tkeys_a, tkeys_b, tvals_a, tvals_b, ctl, end;
    SortCtl c;
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
    // one memset covers the scan's ticket/state words and the sort's control block
    L.zeroed = o;
    L.scan_ws_bytes = scan_workspace_bytes(n_elems);
    L.scan_ws = take(L.scan_ws_bytes);
    L.c = sort_ctl(n_elems, 4);
    L.ctl = take(L.c.bytes);
    L.zeroed_bytes = o - L.zeroed;
    L.end = o;
    return L;
}

// sized for 32-bit keys in four passes (the 16-bit variant needs less)
static TileLayout tile_layout(uint64_t n_isects) {
    TileLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes); return r; };
    L.tkeys_a = take(4 * n_isects); L.tkeys_b = take(4 * n_isects);
    L.tvals_a = take(4 * n_isects); L.tvals_b = take(4 * n_isects);
    L.c = sort_ctl(n_isects, 4);
    L.ctl = take(L.c.bytes);
    L.end = o;
    return L;
}

}  // namespace b2s

using namespace b2s;

extern "C" size_t b200splat_sort_workspace_bytes(uint64_t n_isects) {
    if (n_isects == 0) return 0;
    return sort_ctl(n_isects, kMaxPasses).bytes + 256;
}

// keys_a/vals_a hold the input and are clobbered; the sorted result lands in buffer
// `*selector_out` (0 = *_a, 1 = *_b), like cub::DoubleBuffer in the reference.
extern "C" int b200splat_isect_sort(uint64_t n_isects, uint32_t end_bit, int64_t *keys_a, int32_t *vals_a,
                                    int64_t *keys_b, int32_t *vals_b, void *workspace, size_t workspace_bytes,
                                    int *selector_out, void *stream) {
    const char *where = "b200splat_isect_sort";
    cudaStream_t st = (cudaStream_t)stream;
    B2S_REQUIRE(end_bit <= 64, where, "end_bit must be <= 64");
    B2S_REQUIRE(selector_out != nullptr, where, "selector_out is required");
    B2S_REQUIRE(n_isects <= 0xffffffffull, where, "more than 2^32 pairs");
    *selector_out = 0;
    if (n_isects == 0 || end_bit == 0) return 0;
    const uint32_t n = (uint32_t)n_isects;
    const DigitPlan plan = make_plan(end_bit);
    const SortCtl c = sort_ctl(n_isects, kMaxPasses);
    B2S_REQUIRE(workspace != nullptr && workspace_bytes >= c.bytes, where,
                "workspace too small (see b200splat_sort_workspace_bytes)");
    char *ws = reinterpret_cast<char *>(workspace);
    cudaError_t e = cudaMemsetAsync(ws, 0, c.lookback + c.per_pass * plan.npass, st);
    if (e != cudaSuccess) return fail_cuda(where, e);
    uint32_t *hist = reinterpret_cast<uint32_t *>(ws + c.hist);
    uint32_t *ticket = reinterpret_cast<uint32_t *>(ws + c.ticket);
    using K = unsigned long long;
    K *ka = reinterpret_cast<K *>(keys_a), *kb = reinterpret_cast<K *>(keys_b);
    uint32_t *va = reinterpret_cast<uint32_t *>(vals_a), *vb = reinterpret_cast<uint32_t *>(vals_b);
    radix_hist_kernel<K><<<min(div_up(n, kThreads), 4u * kNumSMs), kThreads, 0, st>>>(n, ka, plan, hist);
    B2S_CHECK_LAUNCH(where);
    for (uint32_t p = 0; p < plan.npass; ++p) {
        e = launch_onesweep<K, kModePairs>(ka, va, kb, vb, n, plan.shift[p], plan.mask[p], hist + p * kBins,
                                           reinterpret_cast<K *>(ws + c.lookback + c.per_pass * p), ticket + p, nullptr,
                                           nullptr, st);
        if (e != cudaSuccess) return fail_cuda(where, e);
        K *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    *selector_out = (int)(plan.npass & 1u);
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
    *selector_out = 0;  // four passes: the order ends up in the `a` buffer
    if (n_elems == 0) return 0;
    const uint32_t n = (uint32_t)n_elems;
    const DepthLayout L = depth_layout(n_elems);
    B2S_REQUIRE(workspace != nullptr && workspace_bytes >= L.end, where,
                "workspace too small (see b200splat_isect_depth_order_workspace_bytes)");
    char *ws = reinterpret_cast<char *>(workspace);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(ws + off); };
    cudaError_t e = cudaMemsetAsync(ws + L.zeroed, 0, L.zeroed_bytes, st);
    if (e != cudaSuccess) return fail_cuda(where, e);
    const DigitPlan plan = make_plan(32);
    uint32_t *hist = u32(L.ctl + L.c.hist), *ticket = u32(L.ctl + L.c.ticket);
    auto lookback = [&](uint32_t p) {
        return reinterpret_cast<unsigned long long *>(ws + L.ctl + L.c.lookback + L.c.per_pass * p);
    };
    // 1. Gaussians by depth
    depth_keys_kernel<<<min(div_up(n, kThreads), 4u * kNumSMs), kThreads, 0, st>>>(
        n, tiles_per_gauss, depths, u32(L.gkeys_a), u32(L.gvals_a), plan, hist);
    B2S_CHECK_LAUNCH(where);
    uint32_t *ka = u32(L.gkeys_a), *kb = u32(L.gkeys_b), *va = u32(L.gvals_a), *vb = u32(L.gvals_b);
    for (uint32_t p = 0; p < 3; ++p) {
        e = launch_onesweep<uint32_t, kModePairs>(ka, va, kb, vb, n, plan.shift[p], plan.mask[p], hist + p * kBins,
                                                  lookback(p), ticket + p, nullptr, nullptr, st);
        if (e != cudaSuccess) return fail_cuda(where, e);
        uint32_t *t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    // last pass: the order (into gvals_a) and, fused, the tile counts in that order
    int32_t *counts = reinterpret_cast<int32_t *>(ws + L.counts);
    e = launch_onesweep<uint32_t, kModeDepthFinal>(ka, va, kb, vb, n, plan.shift[3], plan.mask[3], hist + 3 * kBins,
                                                   lookback(3), ticket + 3,
                                                   reinterpret_cast<const uint32_t *>(tiles_per_gauss), counts, st);
    if (e != cudaSuccess) return fail_cuda(where, e);
    // 2. offsets in depth order (the scan's workspace was zeroed by the memset above)
    int64_t *cum = reinterpret_cast<int64_t *>(ws + L.cum);
    if (lookback_scan_i32_to_i64(counts, cum, n_elems, reinterpret_cast<int64_t *>(ws + L.total), ws + L.scan_ws,
                                 L.scan_ws_bytes, st, /*workspace_is_zero=*/true))
        return fail(where, "scan failed");
    return 0;
}

template <typename KeyT>
static int tile_order_run(int packed, uint32_t C, uint32_t N, uint32_t n_elems, uint32_t n_isects,
                          const int64_t *camera_ids, const float *means2d, const int32_t *radii, const float *depths,
                          const uint32_t *order, const int64_t *cum, uint32_t tile_size, uint32_t tile_width,
                          uint32_t tile_height, uint32_t tile_n_bits, uint32_t cam_n_bits, int64_t *isect_ids,
                          int32_t *flatten_ids, int32_t *offsets, char *ws, const TileLayout &L, cudaStream_t st,
                          const char *where) {
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(ws + off); };
    KeyT *ka = reinterpret_cast<KeyT *>(ws + L.tkeys_a), *kb = reinterpret_cast<KeyT *>(ws + L.tkeys_b);
    uint32_t *va = u32(L.tvals_a), *vb = u32(L.tvals_b);
    const uint32_t n_tiles = tile_width * tile_height;
    const DigitPlan plan = make_plan(tile_n_bits + cam_n_bits);
    cudaError_t e = cudaMemsetAsync(ws + L.ctl, 0, L.c.lookback + L.c.per_pass * plan.npass, st);
    if (e != cudaSuccess) return fail_cuda(where, e);
    uint32_t *hist = u32(L.ctl + L.c.hist), *ticket = u32(L.ctl + L.c.ticket);
    auto lookback = [&](uint32_t p) {
        return reinterpret_cast<unsigned long long *>(ws + L.ctl + L.c.lookback + L.c.per_pass * p);
    };
    // 3. expand (+ digit histograms)
#define B2S_EXPAND(NP)                                                                                               \
    expand_kernel<KeyT, NP><<<div_up(n_elems, kThreads), kThreads, 0, st>>>(                                          \
        packed, N, n_elems, order, cum, camera_ids, means2d, radii, (float)tile_size, tile_width, tile_height,       \
        tile_n_bits, ka, va, plan, hist)
    switch (plan.npass) {
        case 1: B2S_EXPAND(1); break;
        case 2: B2S_EXPAND(2); break;
        case 3: B2S_EXPAND(3); break;
        default: B2S_EXPAND(4); break;
    }
#undef B2S_EXPAND
    B2S_CHECK_LAUNCH(where);
    // 4. stable sort by cam|tile; the last pass assembles the outputs
    for (uint32_t p = 0; p + 1 < plan.npass; ++p) {
        e = launch_onesweep<KeyT, kModePairs>(ka, va, kb, vb, n_isects, plan.shift[p], plan.mask[p], hist + p * kBins,
                                              lookback(p), ticket + p, nullptr, nullptr, st);
        if (e != cudaSuccess) return fail_cuda(where, e);
        KeyT *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    const uint32_t lp = plan.npass - 1;
    e = launch_onesweep<KeyT, kModeTileFinal>(ka, va, nullptr, reinterpret_cast<uint32_t *>(flatten_ids), n_isects,
                                              plan.shift[lp], plan.mask[lp], hist + lp * kBins, lookback(lp),
                                              ticket + lp, reinterpret_cast<const uint32_t *>(depths), isect_ids, st);
    if (e != cudaSuccess) return fail_cuda(where, e);
    // 5. per-tile offsets
    if (offsets != nullptr) {
        const uint32_t total_tiles = C * n_tiles;
        offsets_search_kernel<<<div_up(total_tiles, kThreads), kThreads, 0, st>>>(n_isects, isect_ids, total_tiles,
                                                                                  n_tiles, tile_n_bits, offsets);
        B2S_CHECK_LAUNCH(where);
    }
    return 0;
}

// ---- phase 2: expand, stable tile sort with fused id assembly, offsets ------------------------
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
    B2S_REQUIRE_ALIGNED8(means2d, where);
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n_elems = packed ? (uint64_t)nnz : (uint64_t)C * N;
    B2S_REQUIRE(!packed || camera_ids != nullptr, where, "camera_ids required when packed");
    B2S_REQUIRE(n_elems <= 0xffffffffull, where, "more than 2^32 (camera, Gaussian) pairs");
    B2S_REQUIRE(n_isects <= 0x7fffffffull, where, "n_isects exceeds int32 offsets");
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
    return (tile_n_bits + cam_n_bits <= 16)
               ? tile_order_run<uint16_t>(packed, C, N, (uint32_t)n_elems, (uint32_t)n_isects, camera_ids, means2d,
                                          radii, depths, order, cum, tile_size, tile_width, tile_height, tile_n_bits,
                                          cam_n_bits, isect_ids, flatten_ids, offsets, ws, L, st, where)
               : tile_order_run<uint32_t>(packed, C, N, (uint32_t)n_elems, (uint32_t)n_isects, camera_ids, means2d,
                                          radii, depths, order, cum, tile_size, tile_width, tile_height, tile_n_bits,
                                          cam_n_bits, isect_ids, flatten_ids, offsets, ws, L, st, where);
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
    B2S_REQUIRE_ALIGNED8(means2d, where);
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
