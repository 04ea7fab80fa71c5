// keys.cu -- canonical 128-bit keys of batched states (equality, ordering, hashing, dedup).
//
// The reference gives every Config / State / Action `== != < <= > >=` and `__hash__`
// (src/simulator/game/helper.hpp:10-25, bound at connect.cpp:56-58 / bounce.cpp:55-57).  The hash and
// ordering VALUES are not pinned by anything in the reference, so the batched side is free to choose:
// one 128-bit key per state, computed on the device, such that two states of one configuration have
// equal keys iff the reference's `==` holds (same grid, player, winner).
//
//   Connect, H*W <= 62: EXACT.  bit (r*W + c) of b0 / b1 = stone of player 0 / 1 on cell (row r from the
//     bottom, column c);  key = b0 | b1 << HW | player << 2HW | (winner + 1) << (2HW + 1)   (<= 127 bits),
//     so equal keys <=> equal states, and the key can be unpacked again.
//   Connect on larger boards, and Bounce: a 128-bit HASH of the state record
//     words w_i = the H*W grid bytes, 8 per little-endian 64-bit word (zero padded), then one tail word
//              (player & 0xFF) | (winner & 0xFF) << 8 | H << 16 | W << 24 | game << 32   (game 1 = Connect, 2 = Bounce)
//     h0 = 0x9E3779B97F4A7C15, h1 = 0xC2B2AE3D27D4EB4F;  for every word:
//              h0 = mix(h0 ^ w),  h1 = mix(h1 + w + 0x632BE59BD9B4E019)
//     mix = the splitmix64 finaliser;  key = (h0, h1).
// One warp stages the 32*H*W grid bytes of 32 consecutive states through shared memory with 128-bit
// loads (rows of 42 bytes are not 16-byte aligned, 32 of them are); lane l then works on state g0 + l.
// HBM-bound: H*W + 2 bytes read, 16 written per state.
#include "bgs_common.cuh"

namespace bgs {
namespace keys {

constexpr int KEY_THREADS = 256;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

template <int GAME>
__global__ void __launch_bounds__(KEY_THREADS)
state_keys_kernel(int H, int W, unsigned long long n, const int8_t* __restrict__ grid,
                  const int8_t* __restrict__ player, const int8_t* __restrict__ winner, uint64_t* __restrict__ keys,
                  bool vec) {
    extern __shared__ __align__(16) uint8_t s_stage[];
    const int HW = H * W;
    const int SPAN = (32 * HW + 15) & ~15;  // bytes of one warp's stage
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* st = s_stage + (size_t)warp * SPAN;
    const uint8_t* mine = st + lane * HW;
    constexpr int WARPS = KEY_THREADS / 32;
    const unsigned long long ngroups = (n + 31ull) / 32ull;
    const bool exact = GAME == 1 && HW <= 62;
    for (unsigned long long group = (unsigned long long)blockIdx.x * WARPS + warp; group < ngroups;
         group += (unsigned long long)gridDim.x * WARPS) {
        const unsigned long long g0 = group * 32ull;
        const unsigned rows = (unsigned)((n - g0) < 32ull ? (n - g0) : 32ull);
        warp_copy_bytes(st, reinterpret_cast<const uint8_t*>(grid) + g0 * (unsigned)HW, rows * (unsigned)HW, lane, vec);
        __syncwarp();
        const unsigned long long i = g0 + lane;
        if (i < n) {
            const unsigned pl = (uint8_t)player[i], wn = (uint8_t)(winner ? winner[i] : (int8_t)-1);
            uint64_t k0, k1;
            if (exact) {
                unsigned __int128 key = 0;
                for (int c = 0; c < HW; ++c) {
                    const unsigned v = mine[c];
                    if (v == 0u) key |= (unsigned __int128)1 << c;
                    else if (v == 1u) key |= (unsigned __int128)1 << (HW + c);
                }
                key |= (unsigned __int128)(pl & 1u) << (2 * HW);
                key |= (unsigned __int128)(((unsigned)(int8_t)wn + 1u) & 3u) << (2 * HW + 1);
                k0 = (uint64_t)key; k1 = (uint64_t)(key >> 64);
            } else {
                uint64_t h0 = 0x9E3779B97F4A7C15ull, h1 = 0xC2B2AE3D27D4EB4Full;
                for (int c0 = 0; c0 < HW; c0 += 8) {
                    uint64_t w = 0;
                    for (int j = 0; j < 8 && c0 + j < HW; ++j) w |= (uint64_t)mine[c0 + j] << (8 * j);
                    h0 = mix64(h0 ^ w);
                    h1 = mix64(h1 + w + 0x632BE59BD9B4E019ull);
                }
                const uint64_t tail = (uint64_t)pl | ((uint64_t)wn << 8) | ((uint64_t)H << 16) | ((uint64_t)W << 24) |
                                      ((uint64_t)GAME << 32);
                h0 = mix64(h0 ^ tail);
                h1 = mix64(h1 + tail + 0x632BE59BD9B4E019ull);
                k0 = h0; k1 = h1;
            }
            *reinterpret_cast<ulonglong2*>(keys + 2 * i) = make_ulonglong2(k0, k1);
        }
        __syncwarp();
    }
}

template <int GAME>
static int launch_keys(int H, int W, uint64_t n, const int8_t* grid, const int8_t* player, const int8_t* winner,
                       uint64_t* keys, void* stream_) {
    if (H < 1 || W < 1 || H * W > 255) return set_error(BGS_EUNSUPPORTED, "state_keys: unsupported board %dx%d", H, W);
    if (!grid || !player || !keys) return set_error(BGS_EINVAL, "state_keys: null required pointer");
    if (((uintptr_t)keys & 15u) != 0) return set_error(BGS_EINVAL, "state_keys: `keys` must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    const size_t smem = (size_t)(KEY_THREADS / 32) * ((32 * H * W + 15) & ~15);
    auto kern = state_keys_kernel<GAME>;
    if (smem > 48 * 1024) BGS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long blocks = (n + KEY_THREADS - 1) / KEY_THREADS;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, KEY_THREADS, smem, (cudaStream_t)stream_>>>(H, W, n, grid, player, winner, keys,
                                                                        ((uintptr_t)grid & 15u) == 0);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

}  // namespace keys
}  // namespace bgs

extern "C" int bgs_connect_keys(int H, int W, uint64_t n, const int8_t* grid, const int8_t* player,
                                const int8_t* winner, uint64_t* keys, void* stream) {
    return bgs::keys::launch_keys<1>(H, W, n, grid, player, winner, keys, stream);
}

extern "C" int bgs_bounce_keys(int H, int W, uint64_t n, const int8_t* grid, const int8_t* player,
                               const int8_t* winner, uint64_t* keys, void* stream) {
    return bgs::keys::launch_keys<2>(H, W, n, grid, player, winner, keys, stream);
}
