// peer.cu — the gradient exchange of the camera-parallel mode as our own kernels over NVLink peer
// memory.  No counterpart in the reference, whose multi-GPU mode is
// Gaussian-sharded and NCCL all-to-all based (G/rendering.py:397-478, G/distributed.py).
//
// Every rank maps one SYMMETRIC buffer of every other rank (torch symmetric memory: cuMem
// allocations exchanged at start-up; `bases` below is the device array of the W mapped base
// addresses, the same offsets are valid in all of them) and, on NVSwitch systems, one MULTICAST
// address that aliases all W buffers.  Three kernels:
//
//   peer_publish   this rank's pre-masked colour cotangents + camera centres -> its own block
//   peer_barrier   all ranks have published (flags in the symmetric buffers, release/acquire at
//                  system scope); the colour backward then reads the W blocks in place (sh.cu)
//   peer_allreduce in-place SUM of the flat gradient arena, two-shot: rank r owns slice r, reduces
//                  it (multimem.ld_reduce: the switch adds the W copies; without multicast: W peer
//                  loads added in rank order) and broadcasts it (multimem.st / W peer stores).
//                  Per GPU and direction the multicast form moves (1 + 1/W) arenas, the peer
//                  load/store form 2 (W-1)/W; measured at 44 MB (tools/peer_bench.py): W = 8 0.141 /
//                  0.155 ms against 0.199 ms for the library collective, W = 2 0.157 / 0.092 / 0.103.
//
// Flag protocol (one 32-bit word per (block, source rank) in the DESTINATION rank's buffer):
// signal = CAS 0 -> 1 with release semantics, spinning while the previous signal is unconsumed;
// wait = CAS 1 -> 0 with acquire semantics.  Self-resetting, so kernels can be replayed without
// an epoch argument.  All ranks must launch the same sequence of flag-using kernels.
#include <algorithm>

#include "common.cuh"

namespace b2s {

constexpr int kArThreads = 512;
constexpr int kArBlocks = 2 * kNumSMs;   // all co-resident: 2 x 512 threads per SM
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;  // a dead peer traps, never hangs

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t *p, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t *p, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return old;
}

// Block `slot` of every rank meets block `slot` of every other rank.  Called by all threads of the
// block; threads 0..W-1 each handle one peer.  Everything the block (and, by stream order, the
// kernels before it) wrote is visible to the peers' blocks once they return, and vice versa.
__device__ __forceinline__ void meet_peers(const unsigned long long *flag_bases, unsigned long long flag_off,
                                           uint32_t W, uint32_t rank, uint32_t slot) {
    __syncthreads();
    const uint32_t t = threadIdx.x;
    if (t < W && t != rank) {
        uint32_t *theirs = reinterpret_cast<uint32_t *>(flag_bases[t] + flag_off) + (size_t)slot * W + rank;
        uint32_t *mine = reinterpret_cast<uint32_t *>(flag_bases[rank] + flag_off) + (size_t)slot * W + t;
        const unsigned long long t0 = now_ns();
        while (cas_release_sys(theirs, 0u, 1u) != 0u)
            if (now_ns() - t0 > kSpinTimeoutNs) __trap();
        while (cas_acquire_sys(mine, 1u, 0u) != 1u)
            if (now_ns() - t0 > kSpinTimeoutNs) __trap();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(32)
peer_barrier_kernel(const unsigned long long *__restrict__ flag_bases, unsigned long long flag_off, uint32_t W,
                    uint32_t rank) {
    meet_peers(flag_bases, flag_off, W, rank, 0);
}

// block = {campos [cams_per_block][3], pad to hdr_floats, v [cams_per_block][N][3]}; cotangents are
// masked here (zero where the clamped colour is zero: invisible or clamped) so that the readers need
// neither radii nor colours of other ranks; camera slots >= C (uneven shards) are zero-filled.
__global__ void __launch_bounds__(kThreads)
peer_publish_kernel(uint32_t C, uint64_t n3, uint32_t cams_per_block, uint32_t hdr_floats,
                    const float *__restrict__ campos, const float *__restrict__ colors,
                    const float *__restrict__ v_colors, float *__restrict__ block) {
    const uint64_t live = (uint64_t)C * n3, total = (uint64_t)cams_per_block * n3;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    if (tid < hdr_floats) block[tid] = tid < 3ull * C ? campos[tid] : 0.f;
    float *dst = block + hdr_floats;
    const uint64_t live4 = live / 4;
    const float4 ones = make_float4(1.f, 1.f, 1.f, 1.f);
    for (uint64_t i = tid; i < live4; i += stride) {
        // colors == NULL: the cotangents are already masked (packed layout scattered to [C,N,3])
        const float4 c = colors != nullptr ? __ldcs(reinterpret_cast<const float4 *>(colors) + i) : ones;
        float4 v = __ldcs(reinterpret_cast<const float4 *>(v_colors) + i);
        v.x = c.x > 0.f ? v.x : 0.f; v.y = c.y > 0.f ? v.y : 0.f;
        v.z = c.z > 0.f ? v.z : 0.f; v.w = c.w > 0.f ? v.w : 0.f;
        reinterpret_cast<float4 *>(dst)[i] = v;
    }
    for (uint64_t i = live4 * 4 + tid; i < total; i += stride)
        dst[i] = i < live ? ((colors == nullptr || colors[i] > 0.f) ? v_colors[i] : 0.f) : 0.f;
}

__device__ __forceinline__ float4 multimem_ld_add(unsigned long long addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(unsigned long long addr, const float4 &v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// In-place SUM over ranks of n4 float4 at `off` bytes into every rank's buffer.  Slice r =
// [r*per, (r+1)*per) is reduced and re-broadcast by rank r; between the two meets only rank r writes
// slice r (everywhere) and only rank r reads it (everywhere), so the operation is race-free in place.
// MULTICAST: the switch adds the W copies (multimem.ld_reduce) and replicates the store (multimem.st).
// Otherwise WT = W peers are read with all WT x U loads in flight (WT == 0: any W, runtime loop) and added
// in rank order.
template <bool MULTICAST, int WT, int U>
__global__ void __launch_bounds__(kArThreads, 2)
peer_allreduce_kernel(const unsigned long long *__restrict__ bases, unsigned long long mc_base, unsigned long long off,
                      uint64_t n4, const unsigned long long *__restrict__ flag_bases, unsigned long long flag_off,
                      uint32_t W, uint32_t rank, int dev_mode) {
    // dev_mode (tools/peer_bench.py, tuning builds only): 1 = no handshakes, 2 = handshakes only
    const uint32_t slot = 1 + blockIdx.x;
    if (dev_mode != 1) meet_peers(flag_bases, flag_off, W, rank, slot);          // every arena is complete
    if (dev_mode == 2) n4 = 0;
    const uint64_t per = (n4 + W - 1) / W;
    const uint64_t lo = (uint64_t)rank * per, hi = lo + per < n4 ? lo + per : n4;
    const uint64_t step = (uint64_t)gridDim.x * kArThreads;
    for (uint64_t i0 = lo + (uint64_t)blockIdx.x * kArThreads + threadIdx.x; i0 < hi; i0 += step * U) {
        float4 acc[U];
        if (MULTICAST) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t i = i0 + u * step;
                if (i < hi) acc[u] = multimem_ld_add(mc_base + off + i * 16);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t i = i0 + u * step;
                if (i < hi) multimem_st(mc_base + off + i * 16, acc[u]);
            }
        } else if (WT > 0) {
            float4 v[WT > 0 ? WT : 1][U];
#pragma unroll
            for (int p = 0; p < WT; ++p) {
                const float4 *src = reinterpret_cast<const float4 *>(bases[p] + off);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t i = i0 + u * step;
                    v[p][u] = i < hi ? __ldcg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc[u] = v[0][u];
#pragma unroll
                for (int p = 1; p < WT; ++p) {               // rank order: the same sum on every rank
                    acc[u].x += v[p][u].x; acc[u].y += v[p][u].y; acc[u].z += v[p][u].z; acc[u].w += v[p][u].w;
                }
            }
#pragma unroll
            for (int p = 0; p < WT; ++p) {
                float4 *dst = reinterpret_cast<float4 *>(bases[p] + off);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t i = i0 + u * step;
                    if (i < hi) __stcg(dst + i, acc[u]);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t p = 0; p < W; ++p) {
                const float4 *src = reinterpret_cast<const float4 *>(bases[p] + off);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t i = i0 + u * step;
                    if (i < hi) {
                        const float4 x = __ldcg(src + i);
                        acc[u].x += x.x; acc[u].y += x.y; acc[u].z += x.z; acc[u].w += x.w;
                    }
                }
            }
            for (uint32_t p = 0; p < W; ++p) {
                float4 *dst = reinterpret_cast<float4 *>(bases[p] + off);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t i = i0 + u * step;
                    if (i < hi) __stcg(dst + i, acc[u]);
                }
            }
        }
    }
    if (dev_mode != 1) meet_peers(flag_bases, flag_off, W, rank, slot);          // every slice has landed everywhere
}

}  // namespace b2s

using namespace b2s;

extern "C" size_t b200splat_peer_flag_bytes(uint32_t world) {
    return (size_t)(1 + kArBlocks) * world * sizeof(uint32_t);
}

extern "C" int b200splat_peer_publish_cotangents(uint32_t C, uint32_t N, uint32_t cams_per_block, uint32_t hdr_floats,
                                                 const float *campos, const float *colors, const float *v_colors,
                                                 float *block, void *stream) {
    const char *where = "b200splat_peer_publish_cotangents";
    B2S_REQUIRE(cams_per_block > 0 && C <= cams_per_block, where, "C must be <= cams_per_block");
    B2S_REQUIRE(hdr_floats >= 3 * cams_per_block && hdr_floats % 4 == 0 && hdr_floats <= kThreads * kNumSMs, where,
                "hdr_floats must be a multiple of 4 that holds the camera centres");
    B2S_REQUIRE(block != nullptr && (reinterpret_cast<uintptr_t>(block) & 15) == 0, where, "block must be 16-byte aligned");
    const uint64_t n3 = 3ull * N;
    const uint64_t work = (uint64_t)cams_per_block * n3 / 4 + hdr_floats;
    const unsigned grid = (unsigned)std::min<uint64_t>(div_up(work, kThreads), 8ull * kNumSMs);
    peer_publish_kernel<<<std::max(grid, div_up(hdr_floats, kThreads)), kThreads, 0, (cudaStream_t)stream>>>(
        C, n3, cams_per_block, hdr_floats, campos, colors, v_colors, block);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_peer_barrier(uint32_t world, uint32_t rank, const void *flag_bases, uint64_t flag_offset_bytes,
                                      void *stream) {
    const char *where = "b200splat_peer_barrier";
    B2S_REQUIRE(world >= 1 && world <= 32 && rank < world && flag_bases != nullptr, where, "1 <= world <= 32, rank < world");
    if (world == 1) return 0;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long *>(flag_bases),
                                                            flag_offset_bytes, world, rank);
    B2S_CHECK_LAUNCH(where);
    return 0;
}

extern "C" int b200splat_peer_allreduce_f32(uint32_t world, uint32_t rank, const void *peer_bases, uint64_t multicast_base,
                                            uint64_t offset_bytes, uint64_t n_floats, const void *flag_bases,
                                            uint64_t flag_offset_bytes, void *stream) {
    const char *where = "b200splat_peer_allreduce_f32";
    B2S_REQUIRE(world >= 1 && world <= 32 && rank < world, where, "1 <= world <= 32, rank < world");
    B2S_REQUIRE(peer_bases != nullptr && flag_bases != nullptr, where, "peer and flag tables are required");
    B2S_REQUIRE(offset_bytes % 16 == 0 && n_floats % 4 == 0, where, "the range must be made of whole 16-byte words");
    if (world == 1 || n_floats == 0) return 0;
    const unsigned long long *b = reinterpret_cast<const unsigned long long *>(peer_bases);
    const unsigned long long *f = reinterpret_cast<const unsigned long long *>(flag_bases);
    int blocks = kArBlocks, dev_mode = 0;
#ifdef B2S_TUNING
    if (tuning_variant() >= 3000 && tuning_variant() < 6000) {
        dev_mode = (tuning_variant() - 3000) / 1000;
        blocks = tuning_variant() % 1000;
    }
#endif
#define B2S_AR(MC, WT, U)                                                                              \
    peer_allreduce_kernel<MC, WT, U><<<blocks, kArThreads, 0, (cudaStream_t)stream>>>(                 \
        b, multicast_base, offset_bytes, n_floats / 4, f, flag_offset_bytes, world, rank, dev_mode)
    if (multicast_base != 0) B2S_AR(true, 0, 4);
    else if (world == 2) B2S_AR(false, 2, 4);
    else if (world == 4) B2S_AR(false, 4, 2);
    else if (world == 8) B2S_AR(false, 8, 1);
    else B2S_AR(false, 0, 4);
#undef B2S_AR
    B2S_CHECK_LAUNCH(where);
    return 0;
}
