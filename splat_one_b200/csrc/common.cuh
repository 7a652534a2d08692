// common.cuh — shared host/device helpers for libb200splat (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

#include "../../include/b200splat.h"

namespace b2s {

constexpr int kThreads = 256;     // default 1-D block (CS/bindings.h:8 GSPLAT_N_THREADS)
constexpr int kNumSMs = 148;      // B200

// ---- thread-local error string ----------------------------------------------------
std::string &last_error();
int fail(const char *where, const char *msg);
int fail_cuda(const char *where, cudaError_t e);

#define B2S_CHECK_LAUNCH(where)                                                         \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) return b2s::fail_cuda(where, e__);                      \
    } while (0)

#define B2S_REQUIRE(cond, where, msg)                                                   \
    do {                                                                                \
        if (!(cond)) return b2s::fail(where, msg);                                      \
    } while (0)

// means2d arrays are read and written as float2: an 8-byte aligned base is part of the ABI contract
#define B2S_REQUIRE_ALIGNED8(ptr, where)                                                \
    B2S_REQUIRE((ptr) == nullptr || (reinterpret_cast<uintptr_t>(ptr) & 7) == 0, where, #ptr " must be 8-byte aligned (it is accessed as float2)")

// Developer A/B switch for kernel variants (tools/raster_bench.py); 0 = the shipped default.  The
// environment variable is read ONCE, when the library is loaded (capi.cu); the alternative
// instances are only compiled into builds made with B200SPLAT_TUNING=1 (-DB2S_TUNING).
extern const int g_tuning_variant;
static inline int tuning_variant() { return g_tuning_variant; }

static inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// ---- tiny device helpers --------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_max(int v) {
    return __reduce_max_sync(0xffffffffu, v);
}

// streaming (read-once) loads: keep L1 for the gathered tables
__device__ __forceinline__ float ldg_stream(const float *p) { return __ldcs(p); }

}  // namespace b2s
