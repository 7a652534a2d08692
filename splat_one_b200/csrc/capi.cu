// capi.cu — ABI bookkeeping: version, arch string, thread-local error state.
#include "common.cuh"

namespace b2s {

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

extern "C" int b200splat_abi_version(void) { return 11; }
extern "C" const char *b200splat_last_error(void) { return b2s::last_error().c_str(); }
extern "C" const char *b200splat_arch(void) { return "sm_100a"; }
