// capi.cu — ABI bookkeeping: version, arch string, thread-local error state.
#include "common.cuh"

namespace b2s {

const int g_tuning_variant = [] {
    const char *e = getenv("B200SPLAT_TUNING_VARIANT");
    return e ? atoi(e) : 0;
}();

std::string &last_error() {
    static thread_local std::string e;
    return e;
}

int fail(const char *where, const char *msg) {
    last_error() = std::string(where) + ": " + msg;
    return 1;
}

int fail_cuda(const char *where, cudaError_t e) {
    last_error() = std::string(where) + ": CUDA error: " + cudaGetErrorString(e);
    return 2;
}

}  // namespace b2s

extern "C" int b200splat_abi_version(void) { return 15; }
extern "C" const char *b200splat_last_error(void) { return b2s::last_error().c_str(); }
extern "C" const char *b200splat_arch(void) { return "sm_100a"; }

// Device -> pinned (mapped) host memory, a few words, from a kernel: the n_isects / nnz read-back
// of the two-phase calls without a DMA engine, so it cannot queue behind an unrelated large
// cudaMemcpyAsync that the application has in flight (measured: +0.4 ms per step when a 33 MB
// image download overlapped the next step).  dst must be host memory allocated with
// cudaHostAlloc / torch pin_memory (UVA: same pointer on the device).
namespace b2s {
__global__ void copy_small_kernel(const uint32_t *__restrict__ src, volatile uint32_t *dst, uint32_t n_words) {
    for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
}  // namespace b2s

extern "C" int b200splat_copy_small(const void *src, void *dst_pinned_host, uint32_t n_words, void *stream) {
    const char *where = "b200splat_copy_small";
    B2S_REQUIRE(n_words <= 1024, where, "at most 1024 words");
    if (n_words == 0) return 0;
    b2s::copy_small_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t *>(src),
                                                              reinterpret_cast<volatile uint32_t *>(dst_pinned_host),
                                                              n_words);
    B2S_CHECK_LAUNCH(where);
    return 0;
}
