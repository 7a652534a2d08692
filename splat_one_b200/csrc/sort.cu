// sort.cu — stable LSD radix sort of (int64 key, int32 value) pairs for isect_tiles (a6).
// Replaces the cub::DeviceRadixSort::SortPairs call at CS/isect_tiles.cu:252-300.
//
// Round-1 implementation: CUB 2.8 (CUDA 12.9 toolkit) onesweep with its SM100 tuning
// policy behind the C ABI — the same library kernel the reference would instantiate for
// sm_100a, used here as the correctness anchor and the bar to beat.
// The hand-written depth-first sort (sort Gaussians by depth, expand, then two
// tile-digit passes) plugs in behind the same entry point; see DESIGN.md.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

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
