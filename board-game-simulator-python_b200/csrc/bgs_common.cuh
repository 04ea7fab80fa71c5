// bgs_common.cuh -- shared device / host helpers for libbgs_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/bgs_b200.h"

namespace bgs {

// ---------------------------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI
// ---------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);
int require_device();  // BGS_OK or BGS_ENODEVICE

#define BGS_CUDA_TRY(expr)                                         \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::bgs::cuda_error(_e, #expr); \
    } while (0)

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards.
struct DeviceGuard {
    int prev = -1, rc = BGS_OK;
    explicit DeviceGuard(int device) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e == cudaSuccess) e = cudaSetDevice(device);
        if (e != cudaSuccess) { rc = cuda_error(e, "cudaSetDevice"); prev = -1; }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// float[n,2] reward arrays are written as float2: BGS_EINVAL for a pointer that is not 8-byte aligned (null is fine)
inline int check_reward_alignment(const float* reward, const char* who) {
    if (reward && (reinterpret_cast<uintptr_t>(reward) & 7u) != 0)
        return set_error(BGS_EINVAL, "%s: the reward array must be 8-byte aligned", who);
    return BGS_OK;
}

// Number of SMs of the current device (cached per device).
int sm_count();

// Per-device workspace (api.cu): a zeroable claim counter for one launch, and write-only scratch.
int next_counter(unsigned int** out);
int scratch_buffer(size_t bytes, void** out);
// Stream-ordered temporaries from the library's own memory pool (api.cu).
int temp_alloc(void** out, size_t bytes, cudaStream_t stream);
void temp_free(void* ptr, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (Salmon et al., SC'11).  One call yields the 4 draws of plies
// 4b .. 4b+3 of one game: key = seed, counter = (game id lo, game id hi, b, domain).
// The key schedule is warp-uniform, so ptxas keeps it in uniform registers.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr uint32_t DOMAIN_CONNECT = 0u;
constexpr uint32_t DOMAIN_BOUNCE = 1u;

// ---------------------------------------------------------------------------------------------
// Warp-pooled claim of game indices from a global counter: the warp owns the private pool
// [pool_next, pool_next + pool_cnt); the lanes in `m` (ballot of lanes that need a game) take
// consecutive indices from it; ONE atomic per CHUNK (>= 32) indices.  Returns this lane's index
// (meaningful only for lanes in `m`); indices >= n_games mean "no more work".
// ---------------------------------------------------------------------------------------------
struct NoChunkHook {
    __device__ __forceinline__ void operator()(uint32_t) const {}
};

// `on_new_chunk(base)` runs warp-convergently right after a fresh chunk [base, base + CHUNK) was claimed.
template <int CHUNK, class Hook = NoChunkHook>
__device__ __forceinline__ uint32_t claim_index(unsigned m, unsigned int* counter, uint32_t& pool_next,
                                                uint32_t& pool_cnt, Hook on_new_chunk = Hook()) {
    static_assert(CHUNK >= 32, "one chunk must cover a whole warp");
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t want = __popc(m);
    const uint32_t rank = __popc(m & ((1u << lane) - 1u));
    uint32_t id = pool_next + rank;
    if (want > pool_cnt) {  // warp-uniform, once per CHUNK claims
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned)CHUNK);
        base = __shfl_sync(0xffffffffu, base, 0);
        on_new_chunk(base);
        if (rank >= pool_cnt) id = base + (rank - pool_cnt);
        pool_next = base + (want - pool_cnt);
        pool_cnt = CHUNK - (want - pool_cnt);
    } else {
        pool_next += want;
        pool_cnt -= want;
    }
    return id;
}

// Warp-cooperative copy of `span` contiguous bytes; 128-bit accesses when both sides are 16-byte
// aligned (`vec`).
__device__ __forceinline__ void warp_copy_bytes(uint8_t* dst, const uint8_t* src, unsigned span, unsigned lane, bool vec) {
    unsigned done = 0;
    if (vec) {
        const unsigned nvec = span >> 4;
        for (unsigned q = lane; q < nvec; q += 32)
            reinterpret_cast<uint4*>(dst)[q] = reinterpret_cast<const uint4*>(src)[q];
        done = nvec << 4;
    }
    for (unsigned i = done + lane; i < span; i += 32) dst[i] = src[i];
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// length-histogram bin of the stats vector
__device__ __forceinline__ int hist_bin(int length) {
    return length < (BGS_STATS_LEN - BGS_STAT_HIST0 - 1) ? length : (BGS_STATS_LEN - BGS_STAT_HIST0 - 1);
}

constexpr int HIST_BINS = BGS_STATS_LEN - BGS_STAT_HIST0;

}  // namespace bgs
