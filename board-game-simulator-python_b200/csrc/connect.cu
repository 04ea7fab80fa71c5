// connect.cu -- Connect-k kernels for sm_100a and their C-ABI entry points.
//
// Replaces, for millions of games at once, the per-object path of the reference's binding
// src/simulator/game/connect.cpp:24-54 (Config::sample_initial_state, State::get_actions,
// State::get_action_at, Action::sample_next_state, State::has_ended / get_reward / get_grid).
//
// Data layout
//   * On chip a board is two bitboards (stones of player 0 / player 1), bit index
//     = (H-1-row)*W + col, i.e. the TOP row is bits 0..W-1 (so "which columns are playable" is one
//     LOP3 on the low word), held in registers: one 64-bit word when H*W <= 64 (6x7), an
//     unsigned __int128 otherwise (8x9 = 72 bits, 10x12 = 120 bits).  There is no sentinel
//     row/column: k-in-a-row is `AND of K shifted copies` masked with the set of cells from which a
//     K-run in that direction stays on the board, so wrapped runs can never count.
//   * Generic kernel: column heights and the ascending list of playable columns are nibble-packed
//     registers.  LUT kernel (H*W <= 64, W <= 8): see connect_rollout_lut_kernel.
//   * In HBM: per-game records only (length u8, winner i8, optional actions u8[H*W], optional
//     packed final board 16/32 B); grids int8[n,H,W] are produced by connect_export_kernel.
//
// Rollout kernel structure (persistent): grid = resident CTAs only; every lane plays one game at a
// time.  One outer iteration = [retire finished games + claim new game ids, warp-convergent] ->
// [one Philox4x32-10 call = the 4 draws of plies 4b..4b+3] -> [4 predicated plies].  Games start on
// 4-ply boundaries so that the Philox call and the game-end bookkeeping are never divergent, and
// the mover of ply slot j is statically player j & 1 (no board swap).
#include <cstdlib>
#include <type_traits>

#include "bgs_common.cuh"

namespace bgs {
namespace connect {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------
// geometry policies
// ---------------------------------------------------------------------------------------------
// Cells from which a K-run in direction (dc, dr) stays on the board.  `r` here is the BIT row
// (bit = r*W + c, bit row 0 = top board row); the set of lines is symmetric under that flip.
template <typename T>
__host__ __device__ constexpr T valid_starts(int H, int W, int K, int dc, int dr) {
    T m = 0;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const int ce = c + (K - 1) * dc, re = r + (K - 1) * dr;
            if (ce >= 0 && ce < W && re >= 0 && re < H) m |= (T)1 << (r * W + c);
        }
    return m;
}

// the four line directions as (dc, dr) with a positive bit-index delta dr*W + dc
__host__ __device__ constexpr int dir_dc(int i) { return i == 0 ? 1 : (i == 1 ? 0 : (i == 2 ? 1 : -1)); }
__host__ __device__ constexpr int dir_dr(int i) { return i == 0 ? 0 : 1; }

template <int H_, int W_, int K_>
struct StaticGeo {
    static constexpr int NW = (H_ * W_ <= 64) ? 1 : 2;
    typedef typename std::conditional<NW == 1, uint64_t, u128>::type bb_t;
    typedef typename std::conditional<(W_ <= 8), uint32_t, uint64_t>::type nib_t;
    __host__ __device__ static constexpr int H() { return H_; }
    __host__ __device__ static constexpr int W() { return W_; }
    __host__ __device__ static constexpr int K() { return K_; }
    template <int I>
    __device__ static constexpr bb_t valid() {
        return valid_starts<bb_t>(H_, W_, K_, dir_dc(I), dir_dr(I));
    }
};

struct DynGeo {
    static constexpr int NW = 2;
    typedef u128 bb_t;
    typedef uint64_t nib_t;
    int h, w, k;
    uint64_t v[4][2];
    __host__ __device__ int H() const { return h; }
    __host__ __device__ int W() const { return w; }
    __host__ __device__ int K() const { return k; }
    template <int I>
    __device__ bb_t valid() const {
        return ((u128)v[I][1] << 64) | v[I][0];
    }
};

static DynGeo make_dyn_geo(int H, int W, int K) {
    DynGeo g;
    g.h = H; g.w = W; g.k = K;
    for (int i = 0; i < 4; ++i) {
        // boards beyond the bit-word limits only use H / W / K of this struct (byte-scanning kernels)
        const u128 m = (H <= 15 && W <= 16 && H * W <= 128) ? valid_starts<u128>(H, W, K, dir_dc(i), dir_dr(i)) : (u128)0;
        g.v[i][0] = (uint64_t)m;
        g.v[i][1] = (uint64_t)(m >> 64);
    }
    return g;
}

template <typename T>
__device__ __forceinline__ T shr(T x, int s) {
    return s >= (int)(8 * sizeof(T)) ? (T)0 : (x >> s);
}

// Does `me` contain K consecutive stones on any line?  Only the newest stone can have created one,
// but testing the whole board costs the same handful of shift/AND pairs and needs no coordinates.
template <class G, int I>
__device__ __forceinline__ typename G::bb_t runs_dir(const G& g, typename G::bb_t me) {
    const int K = g.K();
    const int d = dir_dr(I) * g.W() + dir_dc(I);
    typename G::bb_t m = me;
    int len = 1;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        if (2 * len <= K) {
            m &= shr(m, len * d);
            len *= 2;
        }
    }
    if (len < K) m &= shr(m, (K - len) * d);
    return m & g.template valid<I>();
}

template <class G>
__device__ __forceinline__ bool has_run(const G& g, typename G::bb_t me) {
    typename G::bb_t acc = runs_dir<G, 0>(g, me);
    acc |= runs_dir<G, 1>(g, me);
    acc |= runs_dir<G, 2>(g, me);
    acc |= runs_dir<G, 3>(g, me);
    return acc != 0;
}

// ---------------------------------------------------------------------------------------------
// rollout kernels
// ---------------------------------------------------------------------------------------------
struct RolloutParams {
    uint32_t n_games;        // <= 2^31 per launch (the host splits larger batches)
    unsigned long long game_id0;
    uint32_t seed_lo, seed_hi;
    uint8_t* actions;        // [n, H*W] pre-filled with 0xFF (ACTIONS variants only)
    uint8_t* length;         // [n]
    int8_t* winner;          // [n]
    uint64_t* final_packed;  // [n, 2*words] (PACKED variants only)
    unsigned long long* stats;  // [BGS_STATS_LEN] or null
    unsigned int* counter;      // zero-initialised claim counter
    const uint64_t* start;      // START variants: [n, start_words] records written by connect_import_kernel
    int8_t* final_grid;         // fused export (line kernel, GRID): int8[n, H, W] written once, at the end of each game
    uint32_t one;               // always 1: an IMAD multiplier ptxas cannot fold (keeps adds on the FMA pipe)
};

#ifndef BGS_ROLLOUT_THREADS
#define BGS_ROLLOUT_THREADS 256
#endif
constexpr int ROLLOUT_THREADS = BGS_ROLLOUT_THREADS;
constexpr int CLAIM_CHUNK = 64;

// Block-level statistics: every finished game bumps s_hist[length]; draws (only possible on a full
// board) are counted apart.  A game won at an odd length was won by player 0, at an even length by
// player 1, so the win counters need no per-game arithmetic at all.
template <bool PARITY_WINS = true>
__device__ __forceinline__ void flush_stats(const unsigned int* s_hist, unsigned int s_draws, int HW,
                                            unsigned long long* stats, int nbins = HIST_BINS) {
    unsigned long long games = 0, steps = 0, odd = 0, even = 0;
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
        const unsigned long long h = s_hist[i];
        if (h) {
            atomicAdd(&stats[BGS_STAT_HIST0 + i], h);
            games += h;
            steps += h * (unsigned)i;
            if (i & 1) odd += h; else even += h;
        }
    }
    games = warp_sum(games); steps = warp_sum(steps); odd = warp_sum(odd); even = warp_sum(even);
    if ((threadIdx.x & 31) == 0 && games) {
        atomicAdd(&stats[BGS_STAT_GAMES], games);
        atomicAdd(&stats[BGS_STAT_STEPS], steps);
        if (PARITY_WINS) {
            atomicAdd(&stats[BGS_STAT_WIN0], odd);
            atomicAdd(&stats[BGS_STAT_WIN1], even);
        }
    }
    if (PARITY_WINS && threadIdx.x == 0 && s_draws) {  // draws were counted as wins of the parity class of H*W
        atomicAdd(&stats[BGS_STAT_DRAWS], (unsigned long long)s_draws);
        atomicAdd(&stats[(HW & 1) ? BGS_STAT_WIN0 : BGS_STAT_WIN1], 0ull - (unsigned long long)s_draws);
    }
}

template <typename bb_t>
__device__ __forceinline__ void store_packed(uint64_t* final_packed, uint32_t idx, int HW, bb_t b0, bb_t b1) {
    // public record format: (H*W <= 64 ? 1 : 2) words per player, whatever bb_t is
    if (HW <= 64) {
        *reinterpret_cast<ulonglong2*>(final_packed + (size_t)idx * 2) = make_ulonglong2((uint64_t)b0, (uint64_t)b1);
    } else {
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(final_packed + (size_t)idx * 4);
        dst[0] = make_ulonglong2((uint64_t)b0, (uint64_t)((u128)b0 >> 64));
        dst[1] = make_ulonglong2((uint64_t)b1, (uint64_t)((u128)b1 >> 64));
    }
}

// ---- generic kernel: any supported board; legal columns / heights as nibble-packed registers ----
template <class G>
struct Lane {
    typename G::bb_t p[2];   // stones of player 0 / 1 (the mover of ply slot j is player j & 1)
    typename G::nib_t hts;   // nibble c = stones in column c
    typename G::nib_t cols;  // nibble j = j-th playable column (ascending)
    uint32_t nleg;           // number of playable columns
    uint32_t t;              // plies played
    int res;                 // winner so far (BGS_WINNER_DRAW while nobody has won)
};

// One ply of player P: pick the k-th playable column, drop, test for a win.  Returns false when the
// game ended.
// ACT: 0 = no trajectory, 1 = one byte per ply straight into the (0xFF pre-filled) row, 2 = four
// 4-bit columns per 16-bit word, stored once per 4-ply block at the start of the game's own row and
// expanded in place by connect_export_actions_kernel.
template <class G, int J, int ACT, bool CHECK = true>
__device__ __forceinline__ bool play_ply(const G& g, Lane<G>& s, uint32_t r, uint8_t* act_row, uint32_t& blk) {
    constexpr int P = J & 1;
    typedef typename G::bb_t bb_t;
    typedef typename G::nib_t nib_t;
    const uint32_t k = __umulhi(r, s.nleg);
    const uint32_t sh = 4u * k;
    const uint32_t c = (uint32_t)(s.cols >> sh) & 15u;
    const uint32_t sh2 = 4u * c;
    const uint32_t h = (uint32_t)(s.hts >> sh2) & 15u;
    s.hts += (nib_t)1 << sh2;
    if (h + 1u == (uint32_t)g.H()) {  // column is now full: delete nibble k from the list
        const nib_t low = ((nib_t)1 << sh) - 1;
        s.cols = (s.cols & low) | ((s.cols >> 4) & ~low);
        s.nleg -= 1;
    }
    if (ACT == 1) act_row[s.t] = (uint8_t)c;
    if (ACT == 2) blk |= c << (4 * J);
    s.t += 1;
    s.p[P] |= (bb_t)1 << (((uint32_t)g.H() - 1u - h) * (uint32_t)g.W() + c);
    if (!CHECK) return true;  // opening plies: nobody can own K stones yet and the board cannot be full
    const bool won = has_run(g, s.p[P]);
    if (won) s.res = P;
    return !(won || s.nleg == 0);
}

template <class G>
__device__ __forceinline__ typename G::nib_t initial_cols(const G& g) {
    typename G::nib_t v = 0;
    for (int c = g.W() - 1; c >= 0; --c) v = (v << 4) | (typename G::nib_t)c;
    return v;
}

// ---- opening phase of the generic kernel (same idea as open_games of the LUT kernel below) ----------
// Lane l of a warp plays the first 8 plies of game base+l; plies before 2K-2 skip the k-in-a-row test
// (nobody can own K stones yet), and since the whole warp is in the same phase nothing is divergent.
// The positions are parked in a per-warp shared-memory ring from which free lanes pick up games.
constexpr int GRING = 64;  // prepared games per warp (>= 31 + 32)
struct __align__(16) PreparedG {
    uint64_t b[4];      // p0 lo/hi, p1 lo/hi
    uint64_t hts, cols;
    uint32_t nleg, idx, t_res, pad;
};

template <class G, int ACT>
__device__ __forceinline__ void open_games_generic(const G& g, const RolloutParams& p, uint32_t base, PreparedG* slot) {
    typedef typename G::nib_t nib_t;
    const int HW = g.H() * g.W();
    const int first_check = 2 * g.K() - 2;  // first ply (0-based) at which the mover can own K stones
    const uint32_t id = base + (threadIdx.x & 31u);
    const bool valid = id < p.n_games;
    const unsigned long long gid = p.game_id0 + id;
    Lane<G> s;
    s.p[0] = 0; s.p[1] = 0; s.hts = 0; s.cols = initial_cols(g); s.nleg = g.W(); s.t = 0; s.res = BGS_WINNER_DRAW;
    bool alive = true;
    // invalid lanes still play (their results are discarded) but must not write trajectories
    uint8_t* act_row = (ACT && valid) ? p.actions + (size_t)id * (unsigned)HW : nullptr;
    uint32_t r[4];
#define BGS_GOPEN_PLY(J, T, BLK)                                                                      \
    if ((T) < first_check) {                                                                          \
        if (ACT == 1 && !act_row) { uint32_t dump = 0; play_ply<G, J, 0, false>(g, s, r[J], nullptr, dump); } \
        else play_ply<G, J, ACT, false>(g, s, r[J], act_row, BLK);                                    \
    } else if (alive) {                                                                               \
        if (ACT == 1 && !act_row) { uint32_t dump = 0; alive = play_ply<G, J, 0, true>(g, s, r[J], nullptr, dump); } \
        else alive = play_ply<G, J, ACT, true>(g, s, r[J], act_row, BLK);                             \
    }
    uint32_t blk0 = 0, blk1 = 0;
    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 0u, DOMAIN_CONNECT, p.seed_lo, p.seed_hi, r);
    BGS_GOPEN_PLY(0, 0, blk0)
    BGS_GOPEN_PLY(1, 1, blk0)
    BGS_GOPEN_PLY(2, 2, blk0)
    BGS_GOPEN_PLY(3, 3, blk0)
    const uint32_t t4 = s.t;
    if (alive) {
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 1u, DOMAIN_CONNECT, p.seed_lo, p.seed_hi, r);
        BGS_GOPEN_PLY(0, 4, blk1)
        BGS_GOPEN_PLY(1, 5, blk1)
        BGS_GOPEN_PLY(2, 6, blk1)
        BGS_GOPEN_PLY(3, 7, blk1)
    }
#undef BGS_GOPEN_PLY
    if (ACT == 2 && valid) {
        uint16_t* row = reinterpret_cast<uint16_t*>(p.actions + (size_t)id * (unsigned)HW);
        row[0] = (uint16_t)blk0;
        if (s.t > t4) row[1] = (uint16_t)blk1;
    }
    const u128 b0 = (u128)s.p[0], b1 = (u128)s.p[1];
    uint4* dst = reinterpret_cast<uint4*>(slot);
    dst[0] = make_uint4((uint32_t)b0, (uint32_t)(b0 >> 32), (uint32_t)(b0 >> 64), (uint32_t)(b0 >> 96));
    dst[1] = make_uint4((uint32_t)b1, (uint32_t)(b1 >> 32), (uint32_t)(b1 >> 64), (uint32_t)(b1 >> 96));
    const uint64_t hts = (uint64_t)s.hts, cols = (uint64_t)s.cols;
    dst[2] = make_uint4((uint32_t)hts, (uint32_t)(hts >> 32), (uint32_t)cols, (uint32_t)(cols >> 32));
    dst[3] = make_uint4(s.nleg, valid ? id : 0xFFFFFFFFu,
                        s.t | ((uint32_t)(s.res + 1) << 8) | (alive ? 0u : 1u << 16), 0u);
}

// Start-position record of the START variants (rollouts from caller-supplied states), written by
// connect_import_kernel: [p0 words | p1 words | nibble-packed column heights | meta], where meta bit 0
// = side to move, bit 1 = already ended, bits 8..15 = winner so far + 1.
__host__ __device__ constexpr int start_words(int HW) { return HW <= 64 ? 4 : 6; }

template <class G, int ACT, bool PACKED, bool START, bool OPEN>
__global__ void __launch_bounds__(ROLLOUT_THREADS)
connect_rollout_kernel(const G g, const RolloutParams p) {
    typedef typename G::bb_t bb_t;
    typedef typename G::nib_t nib_t;
    static_assert(!(START && OPEN), "rollouts from supplied positions have no opening phase");
    __shared__ unsigned int s_hist[HIST_BINS];
    __shared__ unsigned int s_draws;
    __shared__ PreparedG s_gring[OPEN ? ROLLOUT_THREADS / 32 : 1][OPEN ? GRING : 1];
    PreparedG* ring = s_gring[OPEN ? (threadIdx.x >> 5) : 0];
    uint32_t ring_head = 0, ring_cnt = 0;  // warp-uniform
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x == 0) s_draws = 0;
    __syncthreads();

    const int HW = g.H() * g.W();
    const nib_t cols0 = initial_cols(g);

    Lane<G> s;
    s.p[0] = 0; s.p[1] = 0; s.hts = 0; s.cols = cols0; s.nleg = g.W(); s.t = 0; s.res = BGS_WINNER_DRAW;
    bool alive = false;    // a game is in progress on this lane
    bool retired = false;  // no more game indices for this lane
    uint32_t idx = 0;      // index of the lane's game in [0, n)
    uint32_t pool_next = 0, pool_cnt = 0;
    // START only: a game is held (it may have length 0), who moves first, explicit outcome counters
    bool has_game = false;
    uint32_t first = 0, acc_w0 = 0, acc_w1 = 0, acc_dr = 0;

    for (;;) {
        // ---- warp-convergent: retire finished games, claim new ones -------------------------
        if (START ? (has_game && !alive) : (!alive && s.t != 0)) {
            // slot parity is relative to the side that moved first: undo that for the outputs
            const int win = (START && s.res >= 0) ? (int)((uint32_t)s.res ^ first) : s.res;
            p.length[idx] = (uint8_t)s.t;
            p.winner[idx] = (int8_t)win;
            if (PACKED) {
                const bool swapped = START && first != 0;  // selects, not a dynamically indexed array
                store_packed(p.final_packed, idx, HW, swapped ? s.p[1] : s.p[0], swapped ? s.p[0] : s.p[1]);
            }
            atomicAdd(&s_hist[s.t], 1u);
            if (START) {
                acc_w0 += (win == 0); acc_w1 += (win == 1); acc_dr += (win < 0);
                has_game = false;
            } else if (s.res < 0) {
                atomicAdd(&s_draws, 1u);
            }
            s.t = 0;
        }
        const bool need = START ? (!has_game && !retired) : (!alive && !retired);
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (OPEN) {
            if (m) {
                const unsigned lane = threadIdx.x & 31u;
                const uint32_t want = __popc(m);
                if (ring_cnt < want) {  // warp-uniform: prepare 32 more games (one atomic per 32 ids)
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(p.counter, 32u);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    open_games_generic<G, ACT>(g, p, base, ring + ((ring_head + ring_cnt + lane) & (GRING - 1)));
                    ring_cnt += 32;
                    __syncwarp();
                }
                if (need) {
                    const uint32_t rank = __popc(m & ((1u << lane) - 1u));
                    const uint4* src = reinterpret_cast<const uint4*>(ring + ((ring_head + rank) & (GRING - 1)));
                    const uint4 a = src[0], b = src[1], c = src[2], d = src[3];
                    if (d.y == 0xFFFFFFFFu) {
                        retired = true;
                    } else {
                        s.p[0] = (bb_t)(((u128)a.w << 96) | ((u128)a.z << 64) | ((u128)a.y << 32) | a.x);
                        s.p[1] = (bb_t)(((u128)b.w << 96) | ((u128)b.z << 64) | ((u128)b.y << 32) | b.x);
                        s.hts = (nib_t)(((uint64_t)c.y << 32) | c.x);
                        s.cols = (nib_t)(((uint64_t)c.w << 32) | c.z);
                        s.nleg = d.x;
                        idx = d.y;
                        s.t = d.z & 0xFFu;
                        s.res = (int)((d.z >> 8) & 0xFFu) - 1;
                        alive = (d.z >> 16) == 0u;
                    }
                }
                ring_head = (ring_head + want) & (GRING - 1);
                ring_cnt -= want;
                __syncwarp();
            }
        } else if (m) {
            const uint32_t id = claim_index<CLAIM_CHUNK>(m, p.counter, pool_next, pool_cnt);
            if (need) {
                if (id < p.n_games) {
                    idx = id;
                    if (START) {
                        const uint64_t* rec = p.start + (size_t)id * start_words(HW);
                        uint64_t meta;
                        bb_t b0, b1;
                        if (HW <= 64) {
                            b0 = (bb_t)rec[0]; b1 = (bb_t)rec[1];
                            s.hts = (nib_t)rec[2]; meta = rec[3];
                        } else {
                            b0 = (bb_t)(((u128)rec[1] << 64) | rec[0]);
                            b1 = (bb_t)(((u128)rec[3] << 64) | rec[2]);
                            s.hts = (nib_t)rec[4]; meta = rec[5];
                        }
                        first = (uint32_t)meta & 1u;
                        s.p[0] = first ? b1 : b0;  // p[0] = stones of the side that moves at even plies
                        s.p[1] = first ? b0 : b1;
                        const int w_in = (int)((meta >> 8) & 0xFFu) - 1;
                        s.res = w_in < 0 ? BGS_WINNER_DRAW : (int)((uint32_t)w_in ^ first);
                        // ascending list of the columns that are not full
                        s.cols = 0; s.nleg = 0;
                        for (int c = g.W() - 1; c >= 0; --c)
                            if ((uint32_t)((s.hts >> (4 * c)) & 15u) < (uint32_t)g.H()) {
                                s.cols = (s.cols << 4) | (nib_t)c;
                                s.nleg += 1;
                            }
                        alive = ((meta >> 1) & 1ull) == 0ull && s.nleg != 0;
                        has_game = true;
                    } else {
                        s.p[0] = 0; s.p[1] = 0; s.hts = 0; s.cols = cols0; s.nleg = g.W(); s.res = BGS_WINNER_DRAW;
                        alive = true;
                    }
                } else {
                    retired = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, alive || (START && has_game) || (OPEN && s.t != 0))) break;

        // ---- the 4 draws of plies t .. t+3 (t is a multiple of 4 on every live lane) ---------
        const unsigned long long gid = p.game_id0 + idx;
        uint32_t r[4];
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), s.t >> 2, DOMAIN_CONNECT, p.seed_lo, p.seed_hi, r);
        uint8_t* act_row = ACT ? p.actions + (size_t)idx * (unsigned)HW : nullptr;
        const bool started = alive;
        const uint32_t tb = s.t;
        uint32_t blk = 0;
        if (alive) alive = play_ply<G, 0, ACT>(g, s, r[0], act_row, blk);
        if (alive) alive = play_ply<G, 1, ACT>(g, s, r[1], act_row, blk);
        if (alive) alive = play_ply<G, 2, ACT>(g, s, r[2], act_row, blk);
        if (alive) alive = play_ply<G, 3, ACT>(g, s, r[3], act_row, blk);
        if (ACT == 2 && started) *reinterpret_cast<uint16_t*>(act_row + (tb >> 1)) = (uint16_t)blk;
    }
    __syncthreads();
    if (p.stats) {
        flush_stats<!START>(s_hist, s_draws, HW, p.stats);
        if (START) {
            const unsigned long long w0 = warp_sum(acc_w0), w1 = warp_sum(acc_w1), dr = warp_sum(acc_dr);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&p.stats[BGS_STAT_WIN0], w0);
                atomicAdd(&p.stats[BGS_STAT_WIN1], w1);
                atomicAdd(&p.stats[BGS_STAT_DRAWS], dr);
            }
        }
    }
}

// ---- LUT kernel: boards with H*W <= 64 and W <= 8 (the 6x7x4 headline board) ------------------
// The ALU pipe (LOP3/SHF, 64 lanes/clk/SM) is what bounds the rollout, so everything that is not
// the k-in-a-row test is moved off it: the playable-column mask is ONE LOP3 (the top row is bits
// 0..W-1), its population count runs on the XU pipe, "k-th playable column" is a shared-memory
// table lookup (LSU pipe), each column's landing cell lives in per-thread shared-memory bytes (LSU pipe) and
// the arithmetic in between is IMAD (FMA pipe).
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// Integer multiply-add on the FMA pipe (IMAD); used where an add / shift-by-constant would otherwise
// land on the saturated ALU pipe.
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t imad_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// 1 << s, or 0 when s >= 32 (PTX shl clamps the shift amount; one SHF, no compare/select)
__device__ __forceinline__ uint32_t bit_or_zero(uint32_t s) {
    uint32_t d;
    asm volatile("shl.b32 %0, 1, %1;" : "=r"(d) : "r"(s));
    return d;
}

// x << S on the FMA pipe (IMAD.SHL) instead of the ALU pipe (SHF)
template <int S>
__device__ __forceinline__ uint32_t shl_fma(uint32_t x) {
    uint32_t d;
    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "n"(1u << S));
    return d;
}

// k-in-a-row test for one-word boards with the first doubling step done by LEFT shifts: the low
// word's shift is then a plain multiply (FMA pipe) and only the high word needs a funnel shift.
// m(i) = me(i) & me(i-d) is a pair ENDING at i, i.e. the run starting at i-d, so the remaining
// right-shift doubling steps are unchanged and the start mask is shifted up by d.
template <class G, int I>
__device__ __forceinline__ uint32_t runs_dir_mixed(uint64_t me) {  // OR of the two masked halves
    constexpr int K = G::K();
    constexpr int d = dir_dr(I) * G::W() + dir_dc(I);
    constexpr uint64_t vm = (K < 2) ? G::template valid<I>() : (G::template valid<I>() << d);
    constexpr uint32_t vlo = (uint32_t)vm, vhi = (uint32_t)(vm >> 32);
    uint64_t m = me;
    if (K >= 2) {
        const uint32_t lo = (uint32_t)me, hi = (uint32_t)(me >> 32);
        const uint32_t mlo = lo & shl_fma<d>(lo);
        const uint32_t mhi = (G::H() * G::W() > 32) ? (hi & __funnelshift_l(lo, hi, d)) : 0u;
        m = ((uint64_t)mhi << 32) | mlo;
        int len = 2;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            if (2 * len <= K) {
                m &= shr(m, len * d);
                len *= 2;
            }
        }
        if (len < K) m &= shr(m, (K - len) * d);
    }
    // the halves a direction's start mask empties are dropped at compile time, and the caller ORs plain 32-bit
    // values (6x7x4: 5 half-words remain = two 3-input LOP3s; a 64-bit OR + compare took three)
    uint32_t r = 0;
    if (vlo) r = (uint32_t)m & vlo;
    if (vhi) r |= (uint32_t)(m >> 32) & vhi;
    return r;
}

template <class G>
__device__ __forceinline__ bool has_run_mixed(uint64_t me) {
    return (runs_dir_mixed<G, 0>(me) | runs_dir_mixed<G, 1>(me) | runs_dir_mixed<G, 2>(me) | runs_dir_mixed<G, 3>(me)) != 0;
}

// One ply of player P.  `lut` / `ht` / `bitlut` are 32-bit shared-memory addresses: lut[free*8 + k] is the bit
// index of the BOTTOM cell of the k-th playable column ((H-1)*W + c); ht is this thread's 8 column bytes pre-biased
// by -(H-1)*W, so that ht[that index] is the column's LANDING CELL (bottom cell at first, W less after each stone;
// the byte of a full column wraps and is never read again); bitlut[cell] is 1 << cell as two words.
//
// The lane's progress is ONE register: tr = plies played | status << 8 (status 0 = running, 1 / 2 = won by
// player 0 / 1, 4 = draw; 0x80000000 = the lane has left the game loop).  `tb` is its value at the start of the 4-ply block (a multiple of 4, status 0), so slot
// J plays ply tb + J and nothing is written on the hot path: a win is one predicated IMAD (FMA pipe), the
// board-full test exists only in the slot whose ply count can reach H*W, and the caller adds 4 after a whole block.
template <int H, int W, int K, int J, bool ACTIONS>
__device__ __forceinline__ bool lut_ply(uint64_t& me, uint32_t top_occ, uint32_t r, uint32_t tb, uint32_t& tr,
                                        uint32_t lut, uint32_t ht, uint32_t& blk, uint32_t one, uint32_t bitlut) {
    typedef StaticGeo<H, W, K> G;
    constexpr int P = J & 1;
    constexpr int HW = H * W;
    const uint32_t freem = ~top_occ & ((1u << W) - 1u);                   // ALU: one LOP3
    const uint32_t n = (uint32_t)__popc(freem);                           // XU
    const uint32_t cb = lds_u8(imad(freem, 8u * one, imad_hi(r, n, 0u)) + lut);  // FMA, FMA, LSU (base folded into the address)
    const uint32_t hp = imad(cb, one, ht);                                // FMA
    const uint32_t cell = lds_u8(hp);                                     // LSU: the column's landing cell
    sts_u8(hp, imad(one, (uint32_t)(-W), cell));                          // FMA, LSU: one row up
    if (ACTIONS) blk = imad(cb, (1u << (4 * J)) * one, blk);            // FMA (bias removed at the block end)
    // the cell is empty, so adding the bit is OR-ing it, and no carry can cross the words; the bit
    // itself comes from a 64-bit shared-memory table (LSU pipe) instead of two ALU shifts
    const uint2 bit = lds_u64(imad(cell, 8u * one, bitlut));              // FMA, LSU
    const uint32_t lo = imad(bit.x, one, (uint32_t)me);                   // FMA
    const uint32_t hi = (H * W > 32) ? imad(bit.y, one, (uint32_t)(me >> 32)) : 0u;
    me = ((uint64_t)hi << 32) | lo;
    const bool won = has_run_mixed<G>(me);
    if (won) tr = imad(one, (uint32_t)((J + 1) | ((P + 1) << 8)), tb);
    if (((J + 1) & 3) == (HW & 3)) {  // the only slot in which the board can fill up
        const bool full = tb == (uint32_t)(HW - (J + 1));
        if (full && !won) tr = imad(one, (uint32_t)((J + 1) | (4 << 8)), tb);
        return !(won || full);
    }
    return !won;
}

// A ply that cannot end the game (fewer than K stones of the mover on the board, board not full):
// the same move selection and bookkeeping as lut_ply without the k-in-a-row test.
template <int H, int W, int J, bool ACTIONS>
__device__ __forceinline__ void lut_ply_light(uint64_t& me, uint32_t top_occ, uint32_t r, uint32_t lut, uint32_t ht,
                                              uint32_t& blk, uint32_t one, uint32_t bitlut) {
    const uint32_t freem = ~top_occ & ((1u << W) - 1u);
    const uint32_t n = (uint32_t)__popc(freem);
    const uint32_t cb = lds_u8(imad(freem, 8u * one, imad_hi(r, n, 0u)) + lut);
    const uint32_t hp = imad(cb, one, ht);
    const uint32_t cell = lds_u8(hp);
    sts_u8(hp, imad(one, (uint32_t)(-W), cell));
    if (ACTIONS) blk = imad(cb, (1u << (4 * J)) * one, blk);
    const uint2 bit = lds_u64(imad(cell, 8u * one, bitlut));
    const uint32_t lo = imad(bit.x, one, (uint32_t)me);
    const uint32_t hi = (H * W > 32) ? imad(bit.y, one, (uint32_t)(me >> 32)) : 0u;
    me = ((uint64_t)hi << 32) | lo;
}

// The same for a ply before which no column can be full (fewer than H stones on the board): all W columns are
// playable, so the column is mulhi(r, W) itself -- no playable mask, no population count, no table.  `blk`
// receives the column WITHOUT the (H-1)*W bias of the other ply functions.
template <int H, int W, int J, bool ACTIONS>
__device__ __forceinline__ void lut_ply_first(uint64_t& me, uint32_t r, uint32_t ht, uint32_t& blk, uint32_t one,
                                              uint32_t bitlut) {
    const uint32_t col = imad_hi(r, (uint32_t)W, 0u);
    const uint32_t hp = imad(col, one, ht + (uint32_t)((H - 1) * W));
    const uint32_t cell = lds_u8(hp);
    sts_u8(hp, imad(one, (uint32_t)(-W), cell));
    if (ACTIONS) blk = imad(col, (1u << (4 * J)) * one, blk);
    const uint2 bit = lds_u64(imad(cell, 8u * one, bitlut));
    const uint32_t lo = imad(bit.x, one, (uint32_t)me);
    const uint32_t hi = (H * W > 32) ? imad(bit.y, one, (uint32_t)(me >> 32)) : 0u;
    me = ((uint64_t)hi << 32) | lo;
}

// the 8 column bytes of an empty board: byte c = (H-1)*W + c
template <int H, int W>
__device__ __forceinline__ uint2 landing_cells() {
    constexpr uint32_t B = (H - 1) * W;
    return make_uint2((B) | (B + 1) << 8 | (B + 2) << 16 | (B + 3) << 24, (B + 4) | (B + 5) << 8 | (B + 6) << 16 | (B + 7) << 24);
}

// A game that has been played through its opening, waiting in the warp's ring for a free lane.
struct __align__(16) Prepared {
    uint32_t p0lo, p0hi, p1lo, p1hi;
    uint32_t ht_lo, ht_hi;  // the 8 column bytes (landing cells)
    uint32_t idx;           // game index
    uint32_t tr;            // plies played | status << 8 (see lut_ply); 0x80000000 = beyond n_games
};

constexpr int RING = 64;        // prepared games per warp (>= 31 + 32)
constexpr int OPEN_PLIES = 8;   // the opening = the first two Philox blocks

// The opening phase, executed by all 32 lanes of a warp at once: lane l plays the first 8 plies of
// game base+l.  No player can have K stones before ply 2K-2, so those plies skip the k-in-a-row test
// entirely -- and because the whole warp is in the same phase, skipping it is not divergent (in the
// main loop lanes are at arbitrary plies and the test could never be skipped).
template <int H, int W, int K, bool ACTIONS>
__device__ __forceinline__ void open_games(const RolloutParams& p, uint32_t base, Prepared* ring, uint32_t slot,
                                           uint32_t lut, uint32_t ht2, uint2* ht2_row, uint32_t one, uint32_t bitlut) {
    constexpr int HW = H * W;
    static_assert(HW > OPEN_PLIES + 1, "a draw inside the opening is not handled");
    const uint32_t id = base + (threadIdx.x & 31u);
    const bool valid = id < p.n_games;
    const unsigned long long gid = p.game_id0 + id;
    *ht2_row = landing_cells<H, W>();
    uint64_t q0 = 0, q1 = 0;
    uint32_t tr = 0, blk0 = 0, blk1 = 0;
    bool go = true;
    uint32_t r[4];
    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 0u, DOMAIN_CONNECT, p.seed_lo, p.seed_hi, r);
#define BGS_OPEN_PLY(J, TB, ME, BLK)                                                                           \
    if ((TB) + (J) < 2 * K - 2 && (TB) + (J) < H) {                                                            \
        lut_ply_first<H, W, J, ACTIONS>(ME, r[J], ht2, BLK, one, bitlut);                                      \
    } else if ((TB) + (J) < 2 * K - 2) {                                                                       \
        lut_ply_light<H, W, J, ACTIONS>(ME, (uint32_t)q0 | (uint32_t)q1, r[J], lut, ht2, BLK, one, bitlut);  \
    } else if (go) {                                                                                           \
        go = lut_ply<H, W, K, J, ACTIONS>(ME, (uint32_t)q0 | (uint32_t)q1, r[J], (uint32_t)(TB), tr, lut, ht2, BLK, one, bitlut); \
    }
    BGS_OPEN_PLY(0, 0, q0, blk0)
    BGS_OPEN_PLY(1, 0, q1, blk0)
    BGS_OPEN_PLY(2, 0, q0, blk0)
    BGS_OPEN_PLY(3, 0, q1, blk0)
    if (go) {
        tr = 4u;
        if (2 * K - 2 < 8) philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 1u, DOMAIN_CONNECT, p.seed_lo, p.seed_hi, r);
        BGS_OPEN_PLY(0, 4, q0, blk1)
        BGS_OPEN_PLY(1, 4, q1, blk1)
        BGS_OPEN_PLY(2, 4, q0, blk1)
        BGS_OPEN_PLY(3, 4, q1, blk1)
        if (go) tr = 8u;
    }
#undef BGS_OPEN_PLY
    if (ACTIONS && valid) {
        // every slot played by lut_ply / lut_ply_light added (H-1)*W + column, lut_ply_first the bare column
        constexpr uint32_t B = (H - 1) * W;
        constexpr int FIRST = (2 * K - 2 < H) ? 2 * K - 2 : H;  // plies 0 .. FIRST-1 go through lut_ply_first
        uint16_t* row = reinterpret_cast<uint16_t*>(p.actions + (size_t)id * HW);
        const uint32_t t = tr & 0xFFu;
        uint32_t b0 = 0, b1 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= FIRST && (uint32_t)j < t) b0 += B << (4 * j);
            if (4 + j >= FIRST && (uint32_t)(4 + j) < t) b1 += B << (4 * j);
        }
        row[0] = (uint16_t)(blk0 - b0);
        if (t > 4) row[1] = (uint16_t)(blk1 - b1);
    }
    const uint2 hts = *ht2_row;
    uint4* dst = reinterpret_cast<uint4*>(ring + slot);
    dst[0] = make_uint4((uint32_t)q0, (uint32_t)(q0 >> 32), (uint32_t)q1, (uint32_t)(q1 >> 32));
    dst[1] = make_uint4(hts.x, hts.y, id, valid ? tr : 0x80000000u);
}

// NOTE: no minBlocksPerSM argument: with `__launch_bounds__(256, 1)` ptxas spends 72 registers (3 CTAs
// per SM, 1.02 ms), with (256, 8) it squeezes into 32 (1.04 ms); left alone it uses 40 (6 CTAs, 0.97 ms).
#ifndef BGS_LUT_MINBLOCKS
#define BGS_LUT_MINBLOCKS 0
#endif
template <int H, int W, int K, bool ACTIONS, bool PACKED>
__global__ void __launch_bounds__(ROLLOUT_THREADS, BGS_LUT_MINBLOCKS)
connect_rollout_lut_kernel(const RolloutParams p) {
    static_assert(H * W <= 64 && W <= 8, "LUT kernel: one 64-bit board word, at most 8 columns");
    constexpr int WARPS = ROLLOUT_THREADS / 32;
    __shared__ unsigned int s_hist[HIST_BINS];
    __shared__ unsigned int s_draws;
    __shared__ uint8_t s_lut[(1 << W) * 8];                       // [free mask][k] -> k-th set bit
    __shared__ __align__(8) uint8_t s_ht[ROLLOUT_THREADS * 8];    // [thread][column] -> stones (main loop)
    __shared__ __align__(8) uint8_t s_ht2[ROLLOUT_THREADS * 8];   // same, scratch of the opening phase
    __shared__ Prepared s_ring[WARPS][RING];
    __shared__ uint2 s_bit[64];                                   // [cell] -> 1 << cell as two words
    if (threadIdx.x < 64)
        s_bit[threadIdx.x] = make_uint2(threadIdx.x < 32 ? 1u << threadIdx.x : 0u, threadIdx.x >= 32 ? 1u << (threadIdx.x - 32) : 0u);
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x == 0) s_draws = 0;
    for (int i = threadIdx.x; i < (1 << W) * 8; i += blockDim.x) {
        int mask = i >> 3, k = i & 7, c = 0;
        for (; c < W; ++c)
            if ((mask >> c) & 1) {
                if (k == 0) break;
                --k;
            }
        s_lut[i] = (uint8_t)((H - 1) * W + (c < W ? c : 0));
    }
    uint2* ht_row = reinterpret_cast<uint2*>(s_ht + threadIdx.x * 8);
    uint2* ht2_row = reinterpret_cast<uint2*>(s_ht2 + threadIdx.x * 8);
    *ht_row = make_uint2(0u, 0u);
    __syncthreads();
    const uint32_t one = p.one;
    // table bases behind an opaque move: ptxas would otherwise rebuild them (S2R CgaCtaId, LEA, add) in every
    // block; like this they stay in uniform registers and the loads take the [R + UR] form
    uint32_t lut = (uint32_t)__cvta_generic_to_shared(s_lut);
    uint32_t bitlut = (uint32_t)__cvta_generic_to_shared(s_bit);
    uint32_t hist = (uint32_t)__cvta_generic_to_shared(s_hist);
    asm volatile("mov.u32 %0, %0;" : "+r"(lut));
    asm volatile("mov.u32 %0, %0;" : "+r"(bitlut));
    asm volatile("mov.u32 %0, %0;" : "+r"(hist));
    const uint32_t ht = (uint32_t)__cvta_generic_to_shared(ht_row) - (uint32_t)((H - 1) * W);
    const uint32_t ht2 = (uint32_t)__cvta_generic_to_shared(ht2_row) - (uint32_t)((H - 1) * W);
    Prepared* ring = s_ring[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;

    constexpr int HW = H * W;
    static_assert(HW + 1 < HIST_BINS, "draws are counted in bin H*W + 1");
    uint64_t p0 = 0, p1 = 0;
    uint32_t tr = 0;  // plies played | status << 8 (lut_ply); 0 = the lane holds no game, 0x80000000 = no game is left
    uint32_t idx = 0;
    uint32_t ring_head = 0, ring_cnt = 0;  // warp-uniform

    for (;;) {
        // ---- warp-convergent: retire finished games, hand out prepared ones ------------------
        if ((int)tr >= 256) {
            const uint32_t code = tr >> 8;  // 1, 2: the winner + 1; 4: draw
            p.length[idx] = (uint8_t)tr;
            p.winner[idx] = (int8_t)(code - 1u - (code & 4u));
            if (PACKED) store_packed(p.final_packed, idx, HW, p0, p1);
            // one histogram bump per game; a draw goes to bin H*W + 1 (folded back before the flush)
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist + 4u * ((tr & 0xFFu) + (tr >> 10))) : "memory");
            tr = 0;
        }
        const bool need = tr == 0u;
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m) {
            const uint32_t want = __popc(m);
            if (ring_cnt < want) {  // warp-uniform: prepare 32 more games (one atomic per 32 ids)
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(p.counter, 32u);
                base = __shfl_sync(0xffffffffu, base, 0);
                open_games<H, W, K, ACTIONS>(p, base, ring, (ring_head + ring_cnt + lane) & (RING - 1), lut, ht2,
                                             ht2_row, one, bitlut);
                ring_cnt += 32;
                __syncwarp();
            }
            if (need) {
                const uint32_t rank = __popc(m & ((1u << lane) - 1u));
                const uint4* src = reinterpret_cast<const uint4*>(ring + ((ring_head + rank) & (RING - 1)));
                const uint4 a = src[0], b = src[1];
                p0 = ((uint64_t)a.y << 32) | a.x;
                p1 = ((uint64_t)a.w << 32) | a.z;
                *ht_row = make_uint2(b.x, b.y);
                idx = b.z;
                tr = b.w;  // GONE for ids beyond n_games
            }
            ring_head = (ring_head + want) & (RING - 1);
            ring_cnt -= want;
            __syncwarp();
        }
        if (!__any_sync(0xffffffffu, (int)tr > 0)) break;

        // ---- the 4 draws of plies tb .. tb+3 (tb is a multiple of 4 on every running lane) ----
        const unsigned long long gid = p.game_id0 + idx;
        const uint32_t tb = tr;
        uint32_t r[4];
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), tb >> 2, DOMAIN_CONNECT, p.seed_lo, p.seed_hi, r);
        // trajectory: the 4 columns of this block as 4-bit fields of one 16-bit word, accumulated with
        // IMADs; each played slot adds (H-1)*W + c, so the bias of the slots played is subtracted at the end
        uint32_t blk = 0;
        if (tb - 1u < 255u) {  // running (a game that ended in its opening is retired by the next iteration)
            bool go = lut_ply<H, W, K, 0, ACTIONS>(p0, (uint32_t)p0 | (uint32_t)p1, r[0], tb, tr, lut, ht, blk, one, bitlut);
            if (go) go = lut_ply<H, W, K, 1, ACTIONS>(p1, (uint32_t)p0 | (uint32_t)p1, r[1], tb, tr, lut, ht, blk, one, bitlut);
            if (go) go = lut_ply<H, W, K, 2, ACTIONS>(p0, (uint32_t)p0 | (uint32_t)p1, r[2], tb, tr, lut, ht, blk, one, bitlut);
            if (go) go = lut_ply<H, W, K, 3, ACTIONS>(p1, (uint32_t)p0 | (uint32_t)p1, r[3], tb, tr, lut, ht, blk, one, bitlut);
            if (go) tr = imad(one, 4u, tb);
            if (ACTIONS) {
                constexpr uint32_t B = (H - 1) * W;  // bias per played slot
                const uint32_t played = (tr & 0xFFu) - tb;  // 1..4
                const uint32_t bias = played == 4 ? B * 0x1111u : (played == 3 ? B * 0x111u : (played == 2 ? B * 0x11u : B));
                *reinterpret_cast<uint16_t*>(p.actions + (size_t)idx * HW + (tb >> 1)) = (uint16_t)(blk - bias);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // draws were counted in their own bin
        s_draws = s_hist[HW + 1];
        s_hist[HW] += s_hist[HW + 1];
        s_hist[HW + 1] = 0;
    }
    __syncthreads();
    if (p.stats) flush_stats(s_hist, s_draws, HW, p.stats);
}

__device__ __forceinline__ float2 reward_of(int winner) {
    const float r0 = winner == 0 ? 1.f : (winner == 1 ? -1.f : 0.f);
    return make_float2(r0, 0.f - r0);  // 0 - 0 = +0: a draw is [0, 0], not [0, -0]
}

// ---- line kernel: boards of more than 64 cells (8x9x5, 10x12x6) -----------------------------------
// On a two-word (unsigned __int128) board the shift-and-AND run test costs ~100 ALU instructions per
// ply.  This kernel keeps no bitboard at all: every line of the board (H rows, W columns, H+W-1
// diagonals, H+W-1 anti-diagonals) is one 32-bit shared-memory word holding player 0's stones in
// bits 0..15 and player 1's in bits 16..31.  A move ORs one bit into the 4 lines through the new
// cell and tests those 4 words for K consecutive bits: the opponent cannot own a run (the game would
// be over), and bit 15 of each half is never used (lines are at most 15 long), so the whole word is
// tested without extracting the mover's half.  Lines are stored [line][thread], so every access of a
// warp is bank-conflict free whatever lines its lanes touch (measured 9 % faster than a
// [4 lines][thread][4] layout that resets a game with 128-bit stores).
constexpr int LINES_THREADS = 128;

template <int H, int W>
struct LineGeo {
    static constexpr int D = H + W - 1;
    static constexpr int NL = H + W + 2 * D;  // rows | columns | diagonals | anti-diagonals
    static constexpr int NG = (NL + 3) / 4;
    static constexpr int COL0 = H, DIA0 = H + W, ANT0 = H + W + D;
    // "k-th playable column" table: one look-up for W <= 8, else two halves of HALF_BITS bits (a half has at
    // most 8 bits, so its k-th set bit still fits the [mask][8] table)
    static constexpr int HALF_BITS = W <= 8 ? W : (W + 1) / 2;
    static constexpr int LUT_ROWS = 1 << HALF_BITS;
    // Does line li have to be cleared when a game starts?  A diagonal / anti-diagonal shorter than K cannot hold a
    // run, and its word only ever receives the bits of its own (fewer than K, consecutive) cells -- stale bits of
    // earlier games included -- so it is cleared once per kernel and never again: 33 instead of 49 stores per game
    // start on 8x9x5, 44 instead of 64 on 10x12x6, at ~4 active lanes (the reset was 13 % of all instructions and
    // half of the shared-memory wavefronts of the plain kernel).
    __host__ __device__ static constexpr bool per_game_reset(int li, int K) {
        if (li < DIA0) return true;
        const int d = li < ANT0 ? li - DIA0 : li - ANT0;  // the diagonal's length is min(d + 1, D - d, H, W)
        return d >= K - 1 && d <= D - K;
    }
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// K consecutive set bits anywhere in x?
template <int K>
__device__ __forceinline__ uint32_t run_bits(uint32_t x) {
    uint32_t m = x;
    int len = 1;
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        if (2 * len <= K) {
            m &= m >> len;
            len *= 2;
        }
    }
    if (len < K) m &= m >> (K - len);
    return m;
}

// OR `bit` into line `li` of this thread and return the updated word.
__device__ __forceinline__ uint32_t line_or(uint32_t lines, uint32_t li, uint32_t bit) {
    const uint32_t addr = lines + li * (LINES_THREADS * 4);
    const uint32_t x = lds_u32(addr) | bit;
    sts_u32(addr, x);
    return x;
}

// The lane's progress is one register, as in the LUT kernel: tr = plies played | status << 8 (0 running, 1 / 2 won
// by player 0 / 1, 4 draw); `tb` = its value at the start of the 4-ply block, so slot J plays ply tb + J, a win is
// one predicated write and the board-full test exists only in the slot whose ply count can reach H*W.
template <int H, int W, int K, int J, int ACT>
__device__ __forceinline__ bool lines_ply(uint32_t& toprow, uint32_t r, uint32_t tb, uint32_t& tr,
                                          uint32_t lut8, uint32_t lines, uint8_t* act_row, uint32_t& blk) {
    typedef LineGeo<H, W> LG;
    constexpr int P = J & 1;
    constexpr int HW = H * W;
    const uint32_t freem = ~toprow & ((1u << W) - 1u);
    const uint32_t k = __umulhi(r, (uint32_t)__popc(freem));
    uint32_t c;
    constexpr int HB = LG::HALF_BITS;
    if (W <= 8) {
        c = lds_u8(lut8 + freem * 8u + k);
    } else {  // k-th set bit of the W-bit mask from the table of its two HB-bit halves
        const uint32_t lo = freem & ((1u << HB) - 1u), nlo = (uint32_t)__popc(lo);
        const bool hi = k >= nlo;
        c = lds_u8(lut8 + (hi ? (freem >> HB) : lo) * 8u + (hi ? k - nlo : k)) + (hi ? (uint32_t)HB : 0u);
    }
    // the column's line word holds its stones: their number is the row the new stone lands on (one POPC instead
    // of 64-bit nibble arithmetic on a packed height register: 0.723 -> 0.672 ms per 4 Mi 8x9 games)
    const uint32_t caddr = lines + (LG::COL0 + c) * (LINES_THREADS * 4);
    const uint32_t xc0 = lds_u32(caddr);
    const uint32_t h = (uint32_t)__popc(xc0);
    if (h == (uint32_t)(H - 1)) toprow |= 1u << c;
    if (ACT == 1) act_row[tb + J] = (uint8_t)c;
    if (ACT == 2) blk |= c << (4 * J);
    // fused export: byte J of the block word = the column (one PRMT); unplayed slots keep their 0xFF
    if (ACT == 3) blk = __byte_perm(blk, c, J == 0 ? 0x3214 : (J == 1 ? 0x3240 : (J == 2 ? 0x3410 : 0x4210)));
    const uint32_t bc = 1u << (c + 16u * P), bh = 1u << (h + 16u * P);
    const uint32_t xr = line_or(lines, h, bc);                                // row h, position c
    const uint32_t xc = xc0 | bh;                                             // column c, position h
    sts_u32(caddr, xc);
    const uint32_t xd = line_or(lines, LG::DIA0 + c + (uint32_t)(H - 1) - h, bc);  // diagonal (c - h const)
    const uint32_t xa = line_or(lines, LG::ANT0 + c + h, bc);                 // anti-diagonal (c + h const)
    // Two lines per run test: the mover's 16-bit halves of two line words side by side (one PRMT).  Bit 15
    // of a half is never set (lines are at most 15 long), so no run can cross from one half into the other.
    constexpr uint32_t SEL = P ? 0x7632u : 0x5410u;
    const bool won = (run_bits<K>(__byte_perm(xr, xc, SEL)) | run_bits<K>(__byte_perm(xd, xa, SEL))) != 0u;
    if (won) tr = tb + (uint32_t)((J + 1) | ((P + 1) << 8));
    // the only slot in which the board can fill up; a draw leaves `tr` untouched (the caller sees tr == tb)
    if (((J + 1) & 3) == (HW & 3)) return !(won || tb == (uint32_t)(HW - (J + 1)));
    return !won;
}

// ACT: 0 no trajectory / 1 one byte per ply into the pre-filled row / 2 16-bit blocks of 4-bit columns
// (expanded in place by connect_expand_actions_kernel) / 3 FUSED: the warp pre-fills the rows of every
// chunk of 64 game ids it claims with 0xFF (coalesced 128-bit stores) and a lane stores the 4 columns of
// a 4-ply block as ONE 32-bit word in the final uint8[n, H*W] layout -- no expansion pass.
// GRID (fused final grids, int8[n,H,W] in the reference's layout, tensor.hpp:69-87): the H row lines of a
// finished game ARE its board (player 0 in bits 0..W-1, player 1 in bits 16..16+W-1), so when lanes
// retire the whole warp turns the finished lanes' row lines into grid bytes -- lane q produces 8
// consecutive cells of one finished game (two look-ups in a 256-entry table [4 bits of player 0 | 4 bits
// of player 1] -> 4 grid bytes) and writes them with one 64-bit store: every output byte is written
// exactly once, straight from the rollout kernel, and no packed board ever goes through HBM.
template <int H, int W, int K, int ACT, bool PACKED, bool GRID = false>
__global__ void __launch_bounds__(LINES_THREADS)
connect_rollout_lines_kernel(const RolloutParams p) {
    typedef LineGeo<H, W> LG;
    static_assert(H <= 15 && W <= 15, "bit 15 of each half word must stay free");
    constexpr int HW = H * W;
    constexpr bool FUSED = GRID || ACT == 3;
    static_assert(!FUSED || HW % 8 == 0, "fused export: rows are written in 8-byte units");
    static_assert(!(FUSED && PACKED), "fused export replaces the packed boards");
    __shared__ unsigned int s_hist[HW + 2];  // a game lasts at most H*W plies; bin H*W + 1 counts the draws
    __shared__ unsigned int s_draws;
    __shared__ uint8_t s_lut8[LG::LUT_ROWS * 8];  // [mask of HALF_BITS bits][k] -> index of the k-th set bit
    __shared__ __align__(16) uint32_t s_lines[LG::NL * LINES_THREADS];  // [line][thread]: conflict-free whatever lines the lanes touch
    __shared__ uint32_t s_cell4[GRID ? 256 : 1];                    // [p0 nibble | p1 nibble << 4] -> 4 grid bytes
    __shared__ uint2 s_list[GRID ? LINES_THREADS / 32 : 1][32];     // (lane, game index) of the lanes retiring now
    for (int i = threadIdx.x; i < HW + 2; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x == 0) s_draws = 0;
#pragma unroll
    for (int li = 0; li < LG::NL; ++li)  // the lines no game start clears (LineGeo::per_game_reset)
        if (!LG::per_game_reset(li, K)) s_lines[li * LINES_THREADS + threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < LG::LUT_ROWS * 8; i += blockDim.x) {
        int mask = i >> 3, k = i & 7, c = 0;
        for (; c < 8; ++c)
            if ((mask >> c) & 1) {
                if (k == 0) break;
                --k;
            }
        s_lut8[i] = (uint8_t)(c < 8 ? c : 0);
    }
    if (GRID)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            uint32_t v = 0;
            for (int j = 0; j < 4; ++j) v |= (((i >> j) & 1) ? 0u : (((i >> (4 + j)) & 1) ? 1u : 0xFFu)) << (8 * j);
            s_cell4[i] = v;
        }
    __syncthreads();
    const uint32_t lut8 = (uint32_t)__cvta_generic_to_shared(s_lut8);
    // table bases behind an opaque move, as in the LUT kernel: they stay in (uniform) registers instead of being
    // rebuilt (S2R CgaCtaId, LEA, add) by every retiring lane / flush pass
    uint32_t hist = (uint32_t)__cvta_generic_to_shared(s_hist);
    uint32_t cell4 = (uint32_t)__cvta_generic_to_shared(s_cell4);
    asm volatile("mov.u32 %0, %0;" : "+r"(hist));
    asm volatile("mov.u32 %0, %0;" : "+r"(cell4));
    uint32_t* my_lines = s_lines + threadIdx.x;  // line li at my_lines[li * LINES_THREADS]
    const uint32_t lines = (uint32_t)__cvta_generic_to_shared(my_lines);
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const uint32_t warp_lines = lines - lane * 4u;  // line 0 of lane 0
    uint2* list = s_list[GRID ? (threadIdx.x >> 5) : 0];

    // Grid flush geometry, fixed per lane: a pass serves GPP finished games with LPG lanes each.
    //   W % 4 != 0 (8x9): lane (fg, fu) writes the 8-byte unit fu (cells 8*fu .. 8*fu+7, at most two board rows).
    //   W % 4 == 0 (10x12): lane (fg, fu) writes board ROW fu -- one line word in, W/4 table look-ups, W/4 32-bit
    //   stores out; no second row, and H lanes per game instead of H*W/8 (10x12: 3 games a pass instead of 2).
    constexpr bool ROWS = W % 4 == 0;
    constexpr unsigned UPG = ROWS ? (unsigned)H : (unsigned)HW / 8u;         // lane-units per game
    constexpr unsigned LPG = UPG <= 8 ? 8u : (UPG <= 10 ? 10u : 16u);         // 8x9: 10 lanes (3 games a pass), 10x12: 10 (3)
    constexpr unsigned GPP = 32u / LPG;
    static_assert(!GRID || UPG <= 16, "fused grids: at most 16 lane-units per game");
    const unsigned fg = lane / LPG, fu = lane - fg * LPG;
    const bool f_on = fg < GPP && fu < UPG;
    const unsigned fr0 = ROWS ? fu : (8u * fu) / (unsigned)W, fc0 = ROWS ? 0u : 8u * fu - fr0 * (unsigned)W;
    const uint32_t f_x0 = fr0 * (LINES_THREADS * 4);

    constexpr uint32_t EMPTY = 0x40000000u, GONE = 0x80000000u;  // the lane needs a game / no game is left for it
    uint32_t toprow = 0;
    uint32_t tr = EMPTY;  // plies played | status << 8 (lines_ply), or EMPTY / GONE
    uint32_t idx = 0, pool_next = 0, pool_cnt = 0;

    for (;;) {
        // ---- warp-convergent: retire finished games, claim new ones -------------------------
        const bool fin = (tr & 0x700u) != 0u;
        if (fin) {
            const uint32_t code = tr >> 8;  // 1, 2: the winner + 1; 4: draw
            p.length[idx] = (uint8_t)tr;
            p.winner[idx] = (int8_t)(code - 1u - (code & 4u));
            if (PACKED) {  // rebuild the two bitboards from the row lines
                u128 b0 = 0, b1 = 0;
#pragma unroll
                for (int r = 0; r < H; ++r) {
                    const uint32_t x = lds_u32(lines + r * (LINES_THREADS * 4));
                    b0 |= (u128)(x & 0xFFFFu) << ((H - 1 - r) * W);
                    b1 |= (u128)(x >> 16) << ((H - 1 - r) * W);
                }
                store_packed(p.final_packed, idx, HW, b0, b1);
            }
            // a draw goes to bin H*W + 1 (folded back before the flush)
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist + 4u * ((tr & 0xFFu) + (tr >> 10))) : "memory");
            tr = EMPTY;
        }
        if (GRID) {
            const unsigned fm = __ballot_sync(0xffffffffu, fin);
            if (fm) {  // warp-uniform: the whole warp writes the final grids of the lanes in fm
                if (fin) list[__popc(fm & lt)] = make_uint2(lane, idx);
                __syncwarp();
                constexpr uint32_t MW = (1u << W) - 1u;
                const unsigned nfin = (unsigned)__popc(fm);
                unsigned g0 = 0;
#pragma unroll 1
                do {  // one pass in most iterations
                    const unsigned gi = g0 + fg;
                    const bool on = f_on && gi < nfin;
                    const uint2 e = list[on ? gi : 0u];
                    const uint32_t gidx = e.y;
                    const uint32_t src = warp_lines + e.x * 4u;
                    if (ROWS) {
                        if (on) {
                            const uint32_t x = lds_u32(src + f_x0);  // row fu: player 0 in bits 0..W-1, player 1 in 16..16+W-1
                            uint8_t* dst = reinterpret_cast<uint8_t*>(p.final_grid) + ((size_t)gidx * HW + (unsigned)W * fu);
#pragma unroll
                            for (int q = 0; q < W / 4; ++q) {
                                const uint32_t y = x >> (4 * q);
                                *reinterpret_cast<uint32_t*>(dst + 4 * q) = lds_u32(cell4 + 4u * ((y & 0xFu) | ((y >> 12) & 0xF0u)));
                            }
                        }
                    } else {
                        // lane fu of a game's group loads row line fu ONCE (the lanes of a group hit one bank: H
                        // wavefronts instead of 2 * UPG), the two rows a unit needs come from its neighbours by shuffle
                        const uint32_t xrow = (on && fu < (unsigned)H) ? lds_u32(src + fu * (LINES_THREADS * 4)) : 0u;
                        const unsigned gbase = fg * LPG;
                        const uint32_t x0 = __shfl_sync(0xffffffffu, xrow, gbase + fr0);
                        const uint32_t x1 = __shfl_sync(0xffffffffu, xrow, gbase + (fr0 + 1u < (unsigned)H ? fr0 + 1u : fr0));
                        if (on) {
                            const uint32_t a = ((x0 & MW) | ((x1 & MW) << W)) >> fc0;    // player 0's stones on cells 8*fu ..
                            const uint32_t b = ((x0 >> 16) | ((x1 >> 16) << W)) >> fc0;  // player 1's
                            const uint32_t lo = lds_u32(cell4 + 4u * ((a & 0xFu) | ((b & 0xFu) << 4)));
                            const uint32_t hi = lds_u32(cell4 + 4u * (((a >> 4) & 0xFu) | (b & 0xF0u)));
                            *reinterpret_cast<uint2*>(p.final_grid + ((size_t)gidx * HW + 8u * fu)) = make_uint2(lo, hi);
                        }
                    }
                    g0 += GPP;
                } while (g0 < nfin);
                __syncwarp();  // the row lines are read before they are reset below
            }
        }
        const bool need = tr == EMPTY;
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m) {
            const uint32_t id = claim_index<CLAIM_CHUNK>(m, p.counter, pool_next, pool_cnt, [&](uint32_t base) {
                if (ACT == 3 && base < p.n_games) {  // a fresh chunk of ids: 0xFF over its trajectory rows
                    const uint32_t rows = (p.n_games - base) < (uint32_t)CLAIM_CHUNK ? (p.n_games - base) : (uint32_t)CLAIM_CHUNK;
                    uint8_t* dst = p.actions + (size_t)base * HW;  // 16-byte aligned: base is a multiple of 64
                    const uint32_t bytes = rows * (uint32_t)HW, n16 = bytes >> 4;
                    for (uint32_t q = lane; q < n16; q += 32)
                        reinterpret_cast<uint4*>(dst)[q] = make_uint4(~0u, ~0u, ~0u, ~0u);
                    if ((bytes & 8u) && lane == 0) *reinterpret_cast<uint2*>(dst + (n16 << 4)) = make_uint2(~0u, ~0u);
                    __syncwarp();  // orders the fill before the block stores of the lanes that get these ids
                }
            });
            if (need) {
                if (id < p.n_games) {
                    idx = id;
                    // (a cooperative reset -- 8 lanes per starting game, NL/8 stores per pass -- was measured
                    // slower: 0.852 against 0.804 ms per 4 Mi 8x9 games; the 8-way bank conflicts stall the warp)
#pragma unroll
                    for (int li = 0; li < LG::NL; ++li)
                        if (LG::per_game_reset(li, K)) my_lines[li * LINES_THREADS] = 0u;
                    toprow = 0;
                    tr = 0;
                } else {
                    tr = GONE;
                }
            }
        }
        if (!__any_sync(0xffffffffu, tr < EMPTY)) break;

        // ---- the 4 draws of plies tb .. tb+3 (tb is a multiple of 4 on every running lane) ----
        const unsigned long long gid = p.game_id0 + idx;
        const uint32_t tb = tr;
        uint32_t r[4];
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), tb >> 2, DOMAIN_CONNECT, p.seed_lo, p.seed_hi, r);
        uint8_t* act_row = ACT ? p.actions + (size_t)idx * HW : nullptr;
        uint32_t blk = ACT == 3 ? 0xFFFFFFFFu : 0u;
        if (tb < 256u) {  // running
            bool go = lines_ply<H, W, K, 0, ACT>(toprow, r[0], tb, tr, lut8, lines, act_row, blk);
            if (go) go = lines_ply<H, W, K, 1, ACT>(toprow, r[1], tb, tr, lut8, lines, act_row, blk);
            if (go) go = lines_ply<H, W, K, 2, ACT>(toprow, r[2], tb, tr, lut8, lines, act_row, blk);
            if (go) go = lines_ply<H, W, K, 3, ACT>(toprow, r[3], tb, tr, lut8, lines, act_row, blk);
            if (go) tr = tb + 4u;
            else if (tr == tb) tr = (uint32_t)(HW | (4 << 8));  // left without a winner: the board is full
            if (ACT == 2) *reinterpret_cast<uint16_t*>(act_row + (tb >> 1)) = (uint16_t)blk;
            if (ACT == 3) *reinterpret_cast<uint32_t*>(act_row + tb) = blk;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // draws were counted in their own bin
        s_draws = s_hist[HW + 1];
        s_hist[HW] += s_hist[HW + 1];
    }
    __syncthreads();
    if (p.stats) flush_stats(s_hist, s_draws, HW, p.stats, HW + 1);
}

// ---- byte-board kernel: boards beyond the bit-word limits (up to 255 cells, 32 columns) ----------
// Correctness fallback, not a tuned path: the board is a per-thread byte array (local memory), the
// playable columns a 32-bit mask, the k-in-a-row test a scan of the <= 8*(K-1) cells around the new
// stone.  Same draws, same action order, same outputs as the bitboard kernels; `final_packed` receives
// the grid bytes themselves (row 0 bottom, 0xFF empty), (H*W+7)/8 words per game.
constexpr int BYTES_THREADS = 128;

__device__ __forceinline__ int kth_set_bit32(uint32_t w, int k) {
    int pos = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t low = w & ((1u << s) - 1u);
        const int c = __popc(low);
        if (k >= c) { k -= c; w >>= s; pos += s; } else { w = low; }
    }
    return pos;
}

__global__ void __launch_bounds__(BYTES_THREADS)
connect_rollout_bytes_kernel(int H, int W, int K, const RolloutParams p) {
    __shared__ unsigned int s_hist[HIST_BINS];
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int HW = H * W;
    const int PB = ((HW + 7) / 8) * 8;  // bytes of one final_packed record
    const uint32_t all = W >= 32 ? 0xFFFFFFFFu : ((1u << W) - 1u);
    uint8_t board[256];
    uint8_t hts[32];
    uint32_t legal = 0, t = 0, idx = 0, pool_next = 0, pool_cnt = 0;
    int res = BGS_WINNER_DRAW;
    bool alive = false, retired = false;
    uint32_t acc_w0 = 0, acc_w1 = 0, acc_dr = 0;
    unsigned long long acc_steps = 0;
    for (;;) {
        if (!alive && t != 0) {  // retire the finished game
            p.length[idx] = (uint8_t)t;
            p.winner[idx] = (int8_t)res;
            if (p.final_packed) {
                uint8_t* out = reinterpret_cast<uint8_t*>(p.final_packed) + (size_t)idx * PB;
                for (int c = 0; c < HW; ++c) out[c] = board[c];
            }
            acc_w0 += (res == 0); acc_w1 += (res == 1); acc_dr += (res < 0);
            acc_steps += t;
            atomicAdd(&s_hist[hist_bin((int)t)], 1u);
            t = 0;
        }
        const bool need = !alive && !retired;
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m) {
            const uint32_t id = claim_index<CLAIM_CHUNK>(m, p.counter, pool_next, pool_cnt);
            if (need) {
                if (id < p.n_games) {
                    idx = id;
                    for (int c = 0; c < HW; ++c) board[c] = 0xFFu;
                    for (int c = 0; c < W; ++c) hts[c] = 0;
                    legal = all; res = BGS_WINNER_DRAW; alive = true;
                } else {
                    retired = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, alive)) break;
        const unsigned long long gid = p.game_id0 + idx;
        uint32_t r[4];
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), t >> 2, DOMAIN_CONNECT, p.seed_lo, p.seed_hi, r);
        uint8_t* act_row = p.actions ? p.actions + (size_t)idx * HW : nullptr;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            if (!alive) break;
            const int col = kth_set_bit32(legal, (int)__umulhi(r[j], (uint32_t)__popc(legal)));
            const int row = hts[col];
            const uint8_t pl = (uint8_t)(t & 1u);
            board[row * W + col] = pl;
            hts[col] = (uint8_t)(row + 1);
            if (row + 1 == H) legal &= ~(1u << col);
            if (act_row) act_row[t] = (uint8_t)col;
            ++t;
            auto run = [&](int dr, int dc) {
                int cnt = 0, rr = row + dr, cc = col + dc;
                while (cnt < K - 1 && rr >= 0 && rr < H && cc >= 0 && cc < W && board[rr * W + cc] == pl) {
                    ++cnt; rr += dr; cc += dc;
                }
                return cnt;
            };
            if (1 + run(0, 1) + run(0, -1) >= K || 1 + run(-1, 0) >= K || 1 + run(1, 1) + run(-1, -1) >= K ||
                1 + run(1, -1) + run(-1, 1) >= K) {
                res = pl;
                alive = false;
            } else if (legal == 0u) {
                alive = false;  // full board: draw
            }
        }
    }
    __syncwarp();
    if (p.stats) {
        const unsigned long long w0 = warp_sum(acc_w0), w1 = warp_sum(acc_w1), dr = warp_sum(acc_dr), st = warp_sum(acc_steps);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&p.stats[BGS_STAT_GAMES], w0 + w1 + dr);
            atomicAdd(&p.stats[BGS_STAT_WIN0], w0);
            atomicAdd(&p.stats[BGS_STAT_WIN1], w1);
            atomicAdd(&p.stats[BGS_STAT_DRAWS], dr);
            atomicAdd(&p.stats[BGS_STAT_STEPS], st);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&p.stats[BGS_STAT_HIST0 + i], (unsigned long long)s_hist[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// export: per-game records -> the reference's row layouts.  HBM-bound.
//   MODE_GRID    packed boards (16 / 32 B per game)      -> int8[n,H,W] grids (-1 / 0 / 1)
//   MODE_ACTIONS 16-bit blocks of 4-bit columns, stored by the rollout kernel at the start of each
//                game's own row of `out`                 -> uint8[n,H*W] columns, 0xFF after the end
// One warp handles 32 consecutive games: lane l expands game g0+l into the warp's shared-memory
// stage, then the warp writes the 32*H*W contiguous output bytes with 128-bit stores (32*H*W is a
// multiple of 16 for every board, so the stores are aligned although a row of 42 bytes is not).
// The in-place MODE_ACTIONS is safe because every lane reads its row's blocks before the warp
// writes anything.
// ---------------------------------------------------------------------------------------------
constexpr int EXPORT_THREADS = 256;
enum { MODE_GRID = 0, MODE_ACTIONS = 1, MODE_TRAJ = 2 };

// Warp-cooperative copy of `span` contiguous bytes; 128-bit accesses when both sides are 16-byte
// aligned (`vec`), which they are for torch allocations because 32 rows of any board are a multiple
// of 32 bytes.
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, unsigned span, unsigned lane, bool vec) {
    unsigned done = 0;
    if (vec) {
        const unsigned nvec = span >> 4;
        for (unsigned q = lane; q < nvec; q += 32)
            reinterpret_cast<uint4*>(dst)[q] = reinterpret_cast<const uint4*>(src)[q];
        done = nvec << 4;
    }
    for (unsigned i = done + lane; i < span; i += 32) dst[i] = src[i];
}

// The same copy (16-byte aligned, span a multiple of 16) as asynchronous global -> shared copies (cp.async /
// LDGSTS): the warp issues them and goes on; cp_async_wait<N>() + __syncwarp() make all but the N newest groups visible.
__device__ __forceinline__ void warp_copy_async(uint8_t* dst, const uint8_t* src, unsigned span, unsigned lane) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    for (unsigned q = lane; q < (span >> 4); q += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * q), "l"(src + 16ull * q) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 4 one-bit flags (bits 0..3 of x) -> 4 bytes of 0 / 1
__device__ __forceinline__ uint32_t spread4(uint32_t x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }

// Compile-time board: output bytes 4q..4q+3 of a game (cells in the reference's row-major order)
// come from at most two runs of consecutive board bits, so each 32-bit output word costs a couple of
// constant shifts and two multiplies; the words go to the (2-byte aligned) stage row as 16-bit stores.
template <int SH, int SW>
__device__ __forceinline__ void expand_grid_static(u128 v0, u128 v1, uint8_t* mine) {
    constexpr int HW = SH * SW;
    static_assert(HW % 2 == 0, "16-bit stage stores need an even row length");
#pragma unroll
    for (int q = 0; q < (HW + 3) / 4; ++q) {
        constexpr int dummy = 0; (void)dummy;
        const int c0 = 4 * q;                                   // first cell of this word
        const int n = (HW - c0) < 4 ? (HW - c0) : 4;            // cells in this word
        const int r0 = c0 / SW, x0 = c0 % SW;
        const int n1 = (SW - x0) < n ? (SW - x0) : n;           // cells still in board row r0
        const int bitA = (SH - 1 - r0) * SW + x0;               // their first bit
        uint32_t a = (uint32_t)(v0 >> bitA) & ((1u << n1) - 1u);
        uint32_t b = (uint32_t)(v1 >> bitA) & ((1u << n1) - 1u);
        if (n1 < n) {                                           // the rest starts board row r0+1
            const int bitB = (SH - 2 - r0) * SW;
            a |= ((uint32_t)(v0 >> bitB) & ((1u << (n - n1)) - 1u)) << n1;
            b |= ((uint32_t)(v1 >> bitB) & ((1u << (n - n1)) - 1u)) << n1;
        }
        const uint32_t A = spread4(a), B = spread4(b);
        const uint32_t v = (0x01010101u - A - B) * 0xFFu + B;  // 0 / 1 / 0xFF per byte
        *reinterpret_cast<uint16_t*>(mine + c0) = (uint16_t)v;
        if (n > 2) *reinterpret_cast<uint16_t*>(mine + c0 + 2) = (uint16_t)(v >> 16);
    }
}

// Two bitboards (two words each) -> H*W grid bytes at `mine`: compile-time board if SH != 0.
template <int SH, int SW>
__device__ __forceinline__ void expand_grid(int H, int W, const uint64_t (&b0)[2], const uint64_t (&b1)[2], uint8_t* mine) {
    if constexpr (SH != 0) {
        expand_grid_static<SH, SW>(((u128)b0[1] << 64) | b0[0], ((u128)b1[1] << 64) | b1[0], mine);
    } else {
        // bit row br (bits br*W .. br*W+W-1) is board row H-1-br
        for (int br = 0; br < H; ++br) {
            const int bit = br * W, wd = bit >> 6, sh = bit & 63;
            uint64_t r0 = b0[wd] >> sh, r1 = b1[wd] >> sh;
            if (sh + W > 64) {
                r0 |= b0[1] << (64 - sh);
                r1 |= b1[1] << (64 - sh);
            }
            uint8_t* dst = mine + (H - 1 - br) * W;
            for (int c = 0; c < W; c += 4) {
                const uint32_t a = spread4((uint32_t)(r0 >> c)), b = spread4((uint32_t)(r1 >> c));
                const uint32_t v = (0x01010101u - a - b) * 0xFFu + b;  // 0 / 1 / 0xFF per byte
                for (int j = 0; j < 4 && c + j < W; ++j) dst[c + j] = (uint8_t)(v >> (8 * j));
            }
        }
    }
}

//   MODE_TRAJ    trajectories uint8[n_games, H*W] + lengths -> the board after every ply,
//                int8[n_games, H*W+1, H, W] (entry t = position after t plies; entries past the end
//                repeat the final position): observation tensors for a learner.  A "row" is one
//                (game, ply) pair; its lane replays the first t moves on bitboards (t <= H*W cheap
//                register operations) and expands the result -- output-stationary, so the stores are
//                the same coalesced 128-bit stores as for the final grids.
template <int MODE, int SH = 0, int SW = 0>
__global__ void __launch_bounds__(EXPORT_THREADS)
connect_export_rows_kernel(int H, int W, unsigned long long n, const uint64_t* __restrict__ packed,
                           const uint8_t* __restrict__ length, uint8_t* out, bool vec,
                           const uint8_t* __restrict__ actions, int rpl) {
    extern __shared__ __align__(16) uint8_t s_stage[];
    if (SH) { H = SH; W = SW; }
    const int HW = H * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // every lane produces `rpl` consecutive rows (1 except for MODE_TRAJ, where consecutive rows of a
    // game share the replayed prefix): a warp covers 32*rpl rows
    const unsigned rows_per_warp = 32u * (unsigned)rpl;
    uint8_t* st = s_stage + (size_t)warp * rows_per_warp * HW;  // multiple of 32 bytes: 16-byte aligned
    uint8_t* mine = st + (size_t)lane * rpl * HW;
    const unsigned long long ngroups = (n + rows_per_warp - 1ull) / rows_per_warp;
    constexpr int WARPS = EXPORT_THREADS / 32;
    for (unsigned long long group = (unsigned long long)blockIdx.x * WARPS + warp; group < ngroups;
         group += (unsigned long long)gridDim.x * WARPS) {
        const unsigned long long g = group * rows_per_warp + (unsigned long long)lane * rpl;
        if (g < n) {
            if (MODE == MODE_GRID) {
                uint64_t b0[2] = {0, 0}, b1[2] = {0, 0};
                if (HW <= 64) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(packed + g * 2);
                    b0[0] = v.x; b1[0] = v.y;
                } else {
                    const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(packed + g * 4);
                    const ulonglong2 v1 = *reinterpret_cast<const ulonglong2*>(packed + g * 4 + 2);
                    b0[0] = v0.x; b0[1] = v0.y; b1[0] = v1.x; b1[1] = v1.y;
                }
                expand_grid<SH, SW>(H, W, b0, b1, mine);
            } else if (MODE == MODE_TRAJ) {
                const unsigned T = (unsigned)HW + 1u;
                unsigned long long game = g / T;
                unsigned t = (unsigned)(g - game * T);
                unsigned len = length[game];
                const uint8_t* act = actions + game * (unsigned)HW;
                u128 q[2] = {0, 0};
                uint64_t hts = 0;
                auto apply = [&](unsigned j) {  // move j of the current game
                    const unsigned c = act[j];
                    const unsigned h = (unsigned)(hts >> (4u * c)) & 15u;
                    hts += 1ull << (4u * c);
                    const u128 bit = (u128)1 << ((H - 1 - h) * W + c);
                    if (j & 1) q[1] |= bit; else q[0] |= bit;
                };
                const unsigned plies = t < len ? t : len;
                for (unsigned j = 0; j < plies; ++j) apply(j);
                for (int i = 0; i < rpl && g + i < n; ++i) {
                    if (i > 0) {  // the next row: one more ply of the same game, or the next game's empty board
                        if (++t == T) {
                            ++game; t = 0;
                            len = length[game];
                            act += HW;
                            q[0] = 0; q[1] = 0; hts = 0;
                        } else if (t <= len) {
                            apply(t - 1);
                        }
                    }
                    const uint64_t b0[2] = {(uint64_t)q[0], (uint64_t)(q[0] >> 64)};
                    const uint64_t b1[2] = {(uint64_t)q[1], (uint64_t)(q[1] >> 64)};
                    expand_grid<SH, SW>(H, W, b0, b1, mine + (size_t)i * HW);
                }
            } else {
                const uint8_t* row = out + g * (unsigned)HW;
                const int len = length[g];
                for (int t0 = 0; t0 < HW; t0 += 4) {
                    uint32_t v = 0xFFFFFFFFu;
                    if (t0 < len) {
                        const uint32_t w = *reinterpret_cast<const uint16_t*>(row + (t0 >> 1));
                        v = (w & 0xFu) | ((w & 0xF0u) << 4) | ((w & 0xF00u) << 8) | ((w & 0xF000u) << 12);
                        const int live = len - t0;  // plies of this block that were played
                        if (live < 4) v |= 0xFFFFFFFFu << (8 * live);
                    }
                    *reinterpret_cast<uint16_t*>(mine + t0) = (uint16_t)v;  // rows are 2-byte aligned (H*W even)
                    if (t0 + 2 < HW) *reinterpret_cast<uint16_t*>(mine + t0 + 2) = (uint16_t)(v >> 16);
                }
            }
        }
        __syncwarp();
        const unsigned long long g0 = group * rows_per_warp;
        const unsigned rows = (unsigned)((n - g0) < rows_per_warp ? (n - g0) : rows_per_warp);
        const unsigned span = rows * (unsigned)HW;
        warp_copy(out + g0 * (unsigned)HW, st, span, lane, vec);
        __syncwarp();
    }
}

// MODE_ACTIONS for the BASELINE boards (compile-time row length HW, 16-byte aligned `out`): the
// generic kernel above reads every row's 16-bit blocks straight from global memory -- 32 lanes, 32
// different lines per load instruction -- and is LSU-bound (10x12: 0.49 ms for 0.76 GB).  Here the
// warp first stages the 32 rows' blocks in shared memory with coalesced 8 / 16-byte loads, every lane
// then expands its own row in registers (8 plies per 32-bit word: two masks, two byte permutes) and
// the warp writes the 32*HW contiguous bytes back with 128-bit stores.
template <int HW>
__global__ void __launch_bounds__(EXPORT_THREADS)
connect_expand_actions_kernel(unsigned long long n, const uint8_t* __restrict__ length, uint8_t* out) {
    static_assert(HW % 2 == 0, "4-bit blocks come in pairs");
    constexpr int WARPS = EXPORT_THREADS / 32;
    constexpr int SPAN = 32 * HW;         // bytes of one warp group
    constexpr int NB = HW / 2;            // block bytes at the start of every row
    constexpr int NIW = (HW + 7) / 8;     // 32-bit block words per row (8 plies each)
    constexpr int NOW = (HW + 3) / 4;     // 32-bit output words per row
    constexpr bool A8 = HW % 8 == 0;      // rows 8-byte aligned: row-wise staging, 32 / 64-bit shared accesses
    constexpr int N8 = (NB + 7) / 8;      // 8-byte pieces holding a row's blocks
    __shared__ __align__(16) uint8_t s_stage[WARPS * SPAN];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* st = s_stage + warp * SPAN;
    uint8_t* mine = st + lane * HW;
    const unsigned long long ngroups = (n + 31ull) / 32ull;
    for (unsigned long long group = (unsigned long long)blockIdx.x * WARPS + warp; group < ngroups;
         group += (unsigned long long)gridDim.x * WARPS) {
        const unsigned long long g0 = group * 32ull;
        const unsigned rows = (unsigned)((n - g0) < 32ull ? (n - g0) : 32ull);
        uint8_t* base = out + g0 * (unsigned)HW;
        const unsigned span = rows * (unsigned)HW;
        // ---- (a) blocks -> stage, coalesced
        if (A8) {
            for (unsigned q = lane; q < rows * N8; q += 32) {
                const unsigned r = q / N8, k = q - r * N8;
                *reinterpret_cast<uint2*>(st + r * HW + 8 * k) = *reinterpret_cast<const uint2*>(base + r * HW + 8 * k);
            }
        } else {
            for (unsigned q = lane; q < (span >> 4); q += 32)
                reinterpret_cast<uint4*>(st)[q] = reinterpret_cast<const uint4*>(base)[q];
            for (unsigned i = (span & ~15u) + lane; i < span; i += 32) st[i] = base[i];
        }
        __syncwarp();
        // ---- (b) every lane expands its own row in registers
        if (lane < rows) {
            const unsigned len = length[g0 + lane];
            uint32_t w[NIW];
#pragma unroll
            for (int i = 0; i < NIW; ++i) {
                if (A8) {
                    w[i] = *reinterpret_cast<const uint32_t*>(mine + 4 * i);
                } else {
                    w[i] = *reinterpret_cast<const uint16_t*>(mine + 4 * i);
                    if (4 * i + 2 < NB) w[i] |= (uint32_t)*reinterpret_cast<const uint16_t*>(mine + 4 * i + 2) << 16;
                }
            }
            const unsigned fw = len >> 2;                              // output words that are all moves
            const uint32_t pm = 0xFFFFFFFFu << (8u * (len & 3u));      // 0xFF padding of word fw
            uint32_t o[NOW];
#pragma unroll
            for (int ow = 0; ow < NOW; ++ow) {
                const uint32_t even = w[ow >> 1] & 0x0F0F0F0Fu, odd = (w[ow >> 1] >> 4) & 0x0F0F0F0Fu;
                const uint32_t v = __byte_perm(even, odd, (ow & 1) ? 0x7362 : 0x5140);
                const uint32_t mask = (unsigned)ow < fw ? 0u : ((unsigned)ow > fw ? 0xFFFFFFFFu : pm);
                o[ow] = v | mask;
            }
#pragma unroll
            for (int ow = 0; ow < NOW; ++ow) {
                if (A8) {
                    if ((ow & 1) == 0) *reinterpret_cast<uint2*>(mine + 4 * ow) = make_uint2(o[ow], o[ow + 1]);
                } else {
                    *reinterpret_cast<uint16_t*>(mine + 4 * ow) = (uint16_t)o[ow];
                    if (4 * ow + 2 < HW) *reinterpret_cast<uint16_t*>(mine + 4 * ow + 2) = (uint16_t)(o[ow] >> 16);
                }
            }
        }
        __syncwarp();
        // ---- (c) stage -> out
        for (unsigned q = lane; q < (span >> 4); q += 32)
            reinterpret_cast<uint4*>(base)[q] = reinterpret_cast<const uint4*>(st)[q];
        for (unsigned i = (span & ~15u) + lane; i < span; i += 32) base[i] = st[i];
        __syncwarp();
    }
}

// MODE_TRAJ for boards whose rows are a multiple of 8 bytes (8x9, 10x12): cell-stationary.
// Consecutive positions of a game differ in ONE byte, so expanding every position from bitboards
// (~5 instructions per output byte, above) is wasted work.  Here a lane owns 8 fixed cells of one game
// and holds, in registers, the ply at which each of them is filled (tm, 0x7F = never) and by whom
// (ow); position t is then  byte = tm <= t ? ow : 0xFF  for all 8 cells at once (per 32-bit word: one
// subtract whose byte-wise sign bits are the comparison, one sign-replicating PRMT, one LOP3); the
// H*W/8 lanes of a game produce H*W contiguous bytes, 32 / (H*W/8) games share a warp, and 8 rows at a
// time leave through a shared-memory stage as 128-bit stores.  tm / ow are built per game by a ply-parallel scatter: lane
// q handles ply q, its row is the number of earlier plies in the same column (__match_any_sync rank
// + a per-column counter in shared memory).
// every byte of x replaced by 0xFF if its sign bit is set, else 0x00: PRMT with the sign-replicate bit
// of each selector nibble (the __byte_perm intrinsic masks that bit off, hence the PTX)
__device__ __forceinline__ uint32_t sign_bytes(uint32_t x) {
    uint32_t m;
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(m) : "r"(x));
    return m;
}

template <int H, int W>
__global__ void __launch_bounds__(EXPORT_THREADS)
connect_traj_cells_kernel(unsigned long long n_games, const uint8_t* __restrict__ actions,
                          const uint8_t* __restrict__ length, uint8_t* out) {
    constexpr int HW = H * W, T = HW + 1, CH = HW / 8, GPW = 32 / CH, WARPS = EXPORT_THREADS / 32;
    static_assert(HW % 8 == 0 && HW <= 126 && W <= 16 && GPW >= 1, "8-byte cell chunks, ply numbers below 0x7F");
    __shared__ __align__(8) uint8_t s_tm[WARPS][GPW * HW];
    __shared__ __align__(8) uint8_t s_ow[WARPS][GPW * HW];
    __shared__ uint8_t s_cnt[WARPS][GPW * 16];
    constexpr int RB = 8;                        // rows per pass through the stage (7 for 8x9 and more resident CTAs: no gain)
    constexpr int NPASS = (T - 1) / RB;          // full passes; the remaining T - NPASS*RB rows (one) go out last
    constexpr int SPAN = RB * HW;                // bytes of one game per pass
    constexpr int UNITS = SPAN / 8;              // 8-byte units of one game per pass
    static_assert(T - NPASS * RB == 1, "one row left after the full passes");
    // GPW even (10x12: 2 games per warp): game j of a group is an even / odd game for j even / odd, whatever the
    // group, and one game is 8 mod 16 bytes long -- an odd game's bytes start 8 bytes into a 16-byte line.  Its stage
    // area is shifted by the same 8 bytes (SPAN + 16 per game), so stage and global addresses agree mod 16 and the
    // copy-out is 128-bit (8-byte head and tail for the odd game), all of it known at compile time.  GPW odd (8x9):
    // the parity changes with the group; the copy-out works on 8-byte units instead.
    constexpr bool VEC16 = GPW % 2 == 0 && (T * HW) % 16 == 8 && SPAN % 16 == 0;
    // stage bytes per game, padded so that the 64-bit row-piece stores of a half warp (lanes of two games) fall into
    // 32 different banks: 8x9: 9 lanes x 2 words per game, next game 146 words on (= 18 mod 32); 10x12: 15 lanes x 2
    // words, the odd game's bytes start at 1008 + 8 = word 254 (= 30 mod 32).  (ncu before: 16 M of 44 M shared-memory
    // wavefronts were bank conflicts and the kernel waited on the MIO queue.)
    constexpr int SPAN_ST = (H == 8 && W == 9) ? SPAN + 8 : (H == 10 && W == 12) ? SPAN + 48 : (VEC16 ? SPAN + 16 : SPAN);
    static_assert(!VEC16 || SPAN_ST % 16 == 0, "128-bit copy-out: stage areas keep the global alignment");
    __shared__ __align__(16) uint8_t s_stage[WARPS][GPW * SPAN_ST];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    const unsigned sg = lane / CH, ci = lane - sg * CH;  // game of the warp's group, 8-cell chunk
    uint8_t* tm = s_tm[warp];
    uint8_t* ow = s_ow[warp];
    uint8_t* cnt = s_cnt[warp];
    const unsigned long long ngroups = (n_games + GPW - 1ull) / GPW;
    for (unsigned long long group = (unsigned long long)blockIdx.x * WARPS + warp; group < ngroups;
         group += (unsigned long long)gridDim.x * WARPS) {
        const unsigned long long g0 = group * GPW;
        for (unsigned i = lane; i < GPW * HW / 8; i += 32) {
            reinterpret_cast<uint64_t*>(tm)[i] = 0x7F7F7F7F7F7F7F7Full;
            reinterpret_cast<uint64_t*>(ow)[i] = ~0ull;
        }
        for (unsigned i = lane; i < GPW * 16; i += 32) cnt[i] = 0;
        __syncwarp();
        // ---- the group's trajectories and lengths: ONE round trip to global memory (the scatter below
        // would otherwise pay two dependent load latencies per 32 plies), parked in the stage
        uint8_t* stg = s_stage[warp];
        {
            uint2 a8 = make_uint2(0u, 0u);
            if (lane < GPW * CH && g0 + lane / CH < n_games)
                a8 = *reinterpret_cast<const uint2*>(actions + g0 * HW + 8 * lane);
            *reinterpret_cast<uint2*>(stg + 8 * lane) = a8;
        }
        const unsigned len_mine = (lane < GPW && g0 + lane < n_games) ? length[g0 + lane] : 0u;
        __syncwarp();
        // ---- scatter: ply p of game j fills the lowest empty cell of its column
        for (unsigned q0 = 0; q0 < GPW * HW; q0 += 32) {
            const unsigned q = q0 + lane;
            const unsigned j = q / HW, pl = q - j * HW;
            const unsigned lenj = __shfl_sync(0xffffffffu, len_mine, j < GPW ? j : 0);
            const bool valid = q < GPW * HW && pl < lenj;
            const unsigned col = valid ? stg[q] : 0u;
            const unsigned mm = __match_any_sync(0xffffffffu, valid ? (j * 16u + col) : (0x100u + lane));
            const unsigned below = valid ? cnt[j * 16 + col] : 0u;  // stones already in the column
            __syncwarp();
            if (valid) {
                if ((mm >> lane) == 1u) cnt[j * 16 + col] = (uint8_t)(below + __popc(mm));  // last ply of the column
                const unsigned cell = (below + __popc(mm & lt)) * W + col;
                tm[j * HW + cell] = (uint8_t)(pl + 1);
                ow[j * HW + cell] = (uint8_t)(pl & 1);
            }
            __syncwarp();
        }
        // ---- every position of the game, 8 cells per lane, RB rows at a time through the stage: the
        // lanes' 64-bit pieces of a row would reach L2 as partial sectors (72 / 120-byte rows: 0.6 of the
        // copy peak); from the stage every game's RB*HW contiguous bytes leave as 128-bit stores.
        uint2 tm8 = make_uint2(0x7F7F7F7Fu, 0x7F7F7F7Fu), ow8 = make_uint2(0u, 0u);
        const bool mine = sg < GPW && g0 + sg < n_games;
        if (mine) {
            tm8 = *reinterpret_cast<const uint2*>(tm + sg * HW + 8 * ci);
            ow8 = *reinterpret_cast<const uint2*>(ow + sg * HW + 8 * ci);
        }
        // (round 2: both loops fully unrolled -- the thresholds 0x80 | t and every stage / global offset are
        // immediates -- and the copy-out works on 8-byte units per game with a warp-uniform base: 7 instructions per
        // 8 output bytes to build a row piece, 2 per 8 bytes to copy it out; the first form recomputed 64-bit offsets,
        // head / body / tail splits and loop bounds for every game of every pass: 3 470 instructions per group of
        // three 8x9 games, of which the row pieces were 600)
        uint8_t* mystage = stg + sg * SPAN_ST + (VEC16 ? 8u * (sg & 1u) : 0u) + 8 * ci;
        const unsigned long long GB = (unsigned long long)(T * HW);
#pragma unroll
        for (int ps = 0; ps <= NPASS; ++ps) {
            const int rows = ps < NPASS ? RB : 1;
            if (mine) {
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    if (r < rows) {
                        const uint32_t tb = 0x80808080u + (uint32_t)(ps * RB + r) * 0x01010101u;  // 0x80 | t in every byte
                        const uint32_t m0 = sign_bytes(tb - tm8.x), m1 = sign_bytes(tb - tm8.y);
                        *reinterpret_cast<uint2*>(mystage + r * HW) = make_uint2((ow8.x & m0) | ~m0, (ow8.y & m1) | ~m1);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < GPW; ++j) {
                if (g0 + j < n_games) {  // warp-uniform
                    const int bytes = rows * HW;  // of this game in this pass (a multiple of 8)
                    if (VEC16) {
                        const int head = (j & 1) ? 8 : 0;                 // bytes before the first 16-byte line
                        const int nvec = (bytes - head) / 16;
                        const int tail = bytes - head - 16 * nvec;       // 0 or 8
                        uint8_t* dst = out + (g0 + j) * GB + (unsigned)(ps * SPAN);
                        const uint8_t* src = stg + j * SPAN_ST + head;  // byte 0 of the span (the odd game's area starts 8 bytes in)
#pragma unroll
                        for (int i = 0; i < (SPAN / 16 + 31) / 32; ++i)
                            if (32 * i < nvec && (32 * (i + 1) <= nvec || lane < (unsigned)(nvec - 32 * i)))
                                *reinterpret_cast<uint4*>(dst + head + 16u * lane + 512 * i) =
                                    *reinterpret_cast<const uint4*>(src + head + 16u * lane + 512 * i);
                        if (head && lane == 30) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
                        if (tail && lane == 31)
                            *reinterpret_cast<uint2*>(dst + head + 16 * nvec) = *reinterpret_cast<const uint2*>(src + head + 16 * nvec);
                    } else {
                        uint8_t* dst = out + (g0 + j) * GB + (unsigned)(ps * SPAN) + 8u * lane;
                        const uint8_t* src = stg + j * SPAN_ST + 8u * lane;
                        const int units = bytes / 8;
#pragma unroll
                        for (int i = 0; i < (UNITS + 31) / 32; ++i)
                            if (32 * i < units && (32 * (i + 1) <= units || lane < (unsigned)(units - 32 * i)))
                                *reinterpret_cast<uint2*>(dst + 256 * i) = *reinterpret_cast<const uint2*>(src + 256 * i);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// MODE_TRAJ for boards whose cell count is even but not a multiple of 8 (6x7): word-stationary.
// Two consecutive positions of a game are 2*H*W bytes = NW = H*W/2 32-bit words, and position t+2 differs from
// position t in the bytes of (at most) two cells, so a lane owns fixed WORDS of that double row: the 4 cells of
// its word, the ply at which each is filled (minus the byte's position parity) and by whom sit in registers,
// and the word of double position s is  (ow & m) | ~m  with  m = sign_bytes((0x80 | 2s) - tm)  -- five
// instructions per 4 output bytes instead of ~5 per byte for expanding every position from bitboards.
// GPW = 3 games share a warp (3 * 21 = 63 words on 2 * 32 lane slots); their T*H*W contiguous bytes are staged
// in shared memory (16-bit stores: a game is only 2-byte aligned) and leave as 128-bit stores.  tm / ow come from
// the ply-parallel scatter of the cell kernel above.
constexpr int TRAJW_THREADS = 128;
// (round 2, second form) A warp takes PAIRS of games: one game is (H*W+1)*H*W bytes = 2 mod 4 for these boards, so a
// pair starts on a 4-byte boundary and the pair's byte stream can be cut into ALIGNED 32-bit words -- every staged
// word is one 32-bit store (the first form staged 16-bit halves, two stores per word with 2-way bank conflicts: 86 M
// shared-memory store wavefronts per 1 Mi games, LSU pipe 92 % busy).  Game 0 of the pair owns the words at offsets
// r = 0, 4, .., 2*H*W-4 of each double row, game 1 -- shifted by two bytes -- those at r = 2, 6, .., 2*H*W-2; its last
// word straddles into the next double row (two of its bytes are compared against 2(s+1) instead of 2s), and the
// two bytes after game 0's last position are game 1's first two cells of the empty board (always 0xFF).  The lane
// slots are dealt so that the 32 lanes of one store fall into 32 different banks.
template <int H, int W>
__global__ void __launch_bounds__(TRAJW_THREADS)
connect_traj_words_kernel(unsigned long long n_games, const uint8_t* __restrict__ actions,
                          const uint8_t* __restrict__ length, uint8_t* out) {
    constexpr int HW = H * W, T = HW + 1, NW = HW / 2, GPW = 2, WARPS = TRAJW_THREADS / 32;
    constexpr int GB = T * HW;                // bytes of one game
    constexpr int DR = 2 * HW;                // bytes of a double row (two consecutive positions)
    constexpr int NS = (T + 1) / 2;           // double positions (the last one holds a single position: T is odd)
    static_assert(HW % 4 == 2 && HW <= 64 && W <= 16, "H*W = 2 mod 4: a game is 2-byte, a pair of games 4-byte aligned");
    static_assert(NW <= 32, "one game's words of a double row fit one store");
    constexpr int STAGE = (GPW * GB + 15 + 15 + 4) & ~15;
    // tm / ow as DOUBLE-ROW IMAGES: byte b of a game's image is the threshold / owner of byte b of its double row
    // (b < H*W: first position of the pair, threshold = fill ply; b < 2*H*W: second position, fill ply - 1; the 4 bytes
    // beyond: the first cells of the NEXT double row, fill ply - 2, clamped at 0), so a lane's 4 thresholds are one
    // 32-bit load (game 0) or two 16-bit loads (game 1, shifted by two bytes) instead of 8 byte loads and selects
    constexpr int IMG = (DR + 4 + 3) & ~3;
    __shared__ __align__(16) uint8_t s_stage[WARPS][STAGE];
    __shared__ __align__(4) uint8_t s_tm[WARPS][GPW * IMG];
    __shared__ __align__(4) uint8_t s_ow[WARPS][GPW * IMG];
    __shared__ __align__(4) uint8_t s_act[WARPS][(GPW * HW + 3) & ~3];
    __shared__ __align__(4) uint8_t s_cnt[WARPS][GPW * 16];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    uint8_t* tm = s_tm[warp];
    uint8_t* ow = s_ow[warp];
    uint8_t* act = s_act[warp];
    uint8_t* cnt = s_cnt[warp];
    uint8_t* stg = s_stage[warp];
    // ---- this lane's two word slots, fixed for the whole kernel.  Slot 0: lanes 0..NW-1 hold game 0's words, the
    // other lanes those words of game 1 whose bank differs from all of game 0's; slot 1: the rest of game 1.
    // Word wi of game j sits at byte r = 4*wi + 2*j of its double row; game 1's word 0 is C1 words after game 0's.
    constexpr unsigned C1 = ((unsigned)GB + 2u) / 4u;
    unsigned sj[2] = {0u, 1u}, sw[2] = {0u, 0u};
    bool son[2] = {false, false};
    {
        unsigned n0 = NW, n1 = 0;  // lanes dealt so far in slot 0 / slot 1
        if (lane < (unsigned)NW) { son[0] = true; sj[0] = 0u; sw[0] = lane; }
#pragma unroll 1
        for (unsigned wi = 0; wi < (unsigned)NW; ++wi) {
            const bool clash = ((C1 + wi) & 31u) < (unsigned)NW;  // same bank as one of game 0's words
            if (!clash) {
                if (lane == n0) { son[0] = true; sj[0] = 1u; sw[0] = wi; }
                ++n0;
            } else {
                if (lane == n1) { son[1] = true; sj[1] = 1u; sw[1] = wi; }
                ++n1;
            }
        }
    }
    unsigned sr[2];          // byte offset of the word in its game's double row
    uint32_t last_or[2];     // OR-ed into the thresholds in the last (single-position) double row: 0x7F = "never shown"
    bool last_on[2];         // the word exists in the last double row
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        sr[k] = 4u * sw[k] + 2u * sj[k];
        last_on[k] = sr[k] < (unsigned)HW;
        last_or[k] = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (sr[k] + i >= (unsigned)HW) last_or[k] |= 0x7Fu << (8 * i);  // beyond the last position: the next game's empty cells
    }
    const unsigned long long ngroups = (n_games + GPW - 1ull) / GPW;
    for (unsigned long long group = (unsigned long long)blockIdx.x * WARPS + warp; group < ngroups;
         group += (unsigned long long)gridDim.x * WARPS) {
        const unsigned long long g0 = group * GPW;
        const unsigned ng = (unsigned)((n_games - g0) < (unsigned long long)GPW ? (n_games - g0) : GPW);
        // (32-bit accesses: g0 * HW is a multiple of 4 for a pair of games; the words beyond the last game's row
        // are never used -- `valid` below -- but must not be read past the end of `actions`)
        for (unsigned i = lane; i < GPW * IMG / 4; i += 32) {
            reinterpret_cast<uint32_t*>(tm)[i] = 0x7F7F7F7Fu;
            reinterpret_cast<uint32_t*>(ow)[i] = 0xFFFFFFFFu;
        }
        {
            const unsigned nb = ng * HW;  // bytes of trajectory rows of this group
            const uint8_t* src = actions + g0 * HW;
            for (unsigned i = lane; 4u * i < nb; i += 32) {
                uint32_t v;
                if (4u * i + 4u <= nb) v = *reinterpret_cast<const uint32_t*>(src + 4u * i);
                else v = (uint32_t)src[4u * i] | ((uint32_t)src[4u * i + 1u] << 8);  // nb = 2 mod 4: the last half word
                reinterpret_cast<uint32_t*>(act)[i] = v;
            }
        }
        if (lane < GPW * 4) reinterpret_cast<uint32_t*>(cnt)[lane] = 0u;
        const unsigned len_mine = lane < ng ? length[g0 + lane] : 0u;
        __syncwarp();
        // ---- scatter: ply p of game j fills the lowest empty cell of its column
        for (unsigned q0 = 0; q0 < GPW * HW; q0 += 32) {
            const unsigned q = q0 + lane;
            const unsigned j = q / HW, pl = q - j * HW;
            const unsigned lenj = __shfl_sync(0xffffffffu, len_mine, j < GPW ? j : 0);
            const bool valid = q < GPW * HW && pl < lenj;
            const unsigned col = valid ? act[q] : 0u;
            const unsigned mm = __match_any_sync(0xffffffffu, valid ? (j * 16u + col) : (0x100u + lane));
            const unsigned below = valid ? cnt[j * 16 + col] : 0u;  // stones already in the column
            __syncwarp();
            if (valid) {
                if ((mm >> lane) == 1u) cnt[j * 16 + col] = (uint8_t)(below + __popc(mm));  // last ply of the column
                const unsigned cell = (below + __popc(mm & lt)) * W + col;
                uint8_t* tj = tm + j * IMG + cell;
                uint8_t* oj = ow + j * IMG + cell;
                const uint8_t o = (uint8_t)(pl & 1);
                tj[0] = (uint8_t)(pl + 1); oj[0] = o;         // first position of a pair: shown from position pl + 1
                tj[HW] = (uint8_t)pl; oj[HW] = o;             // second position: one earlier
                if (cell < 4u) { tj[DR] = (uint8_t)(pl > 0u ? pl - 1u : 0u); oj[DR] = o; }  // seen from the previous double row
            }
            __syncwarp();
        }
        // ---- this lane's words: the 4 thresholds ("shown in double position s iff thr <= 2s") and owners
        uint32_t tm4[2], ow4[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            tm4[k] = 0x7F7F7F7Fu; ow4[k] = 0u;
            if (son[k] && sj[k] < ng) {
                const uint8_t* tp = tm + sj[k] * IMG + sr[k];
                const uint8_t* op = ow + sj[k] * IMG + sr[k];
                if (sj[k] == 0u) {  // 4-byte aligned
                    tm4[k] = *reinterpret_cast<const uint32_t*>(tp);
                    ow4[k] = *reinterpret_cast<const uint32_t*>(op);
                } else {            // 2 mod 4
                    tm4[k] = (uint32_t)*reinterpret_cast<const uint16_t*>(tp) | ((uint32_t)*reinterpret_cast<const uint16_t*>(tp + 2) << 16);
                    ow4[k] = (uint32_t)*reinterpret_cast<const uint16_t*>(op) | ((uint32_t)*reinterpret_cast<const uint16_t*>(op + 2) << 16);
                }
            }
        }
        // ---- all positions, two at a time: one aligned 32-bit store per word
        const unsigned long long G0 = g0 * (unsigned long long)GB;  // a multiple of 4
        const unsigned pad = (unsigned)(G0 & 15ull);
        // (fully unrolled: the thresholds 0x80 | 2s and the stage offsets s * DR are immediates, 4 instructions per
        // word -- subtract, sign-replicating PRMT, LOP3, predicated store)
        const bool on0 = son[0] && sj[0] < ng, on1 = son[1] && sj[1] < ng;
        uint8_t* base0 = stg + pad + sj[0] * GB + sr[0];
        uint8_t* base1 = stg + pad + sj[1] * GB + sr[1];
#pragma unroll
        for (int s2 = 0; s2 < NS - 1; ++s2) {
            const uint32_t tb = 0x80808080u + (uint32_t)s2 * 0x02020202u;
            const uint32_t m0 = sign_bytes(tb - tm4[0]), m1 = sign_bytes(tb - tm4[1]);
            if (on0) *reinterpret_cast<uint32_t*>(base0 + s2 * DR) = (ow4[0] & m0) | ~m0;
            if (on1) *reinterpret_cast<uint32_t*>(base1 + s2 * DR) = (ow4[1] & m1) | ~m1;
        }
        {
            constexpr uint32_t tb = 0x80808080u + (uint32_t)(NS - 1) * 0x02020202u;
            if (on0 && last_on[0]) {  // the last double row holds one position
                const uint32_t m = sign_bytes(tb - (tm4[0] | last_or[0]));
                *reinterpret_cast<uint32_t*>(base0 + (NS - 1) * DR) = (ow4[0] & m) | ~m;
            }
            if (on1 && last_on[1]) {
                const uint32_t m = sign_bytes(tb - (tm4[1] | last_or[1]));
                *reinterpret_cast<uint32_t*>(base1 + (NS - 1) * DR) = (ow4[1] & m) | ~m;
            }
        }
        __syncwarp();
        // ---- stage -> global: head (32-bit stores up to the first 16-byte boundary), 128-bit body, tail
        const unsigned L = ng * (unsigned)GB;  // ng == 1 (the last, odd game): L = 2 mod 4, the tail ends with a 16-bit store
        uint8_t* gdst = out + G0;
        const unsigned head = (16u - pad) & 15u;  // a multiple of 4
        if (4u * lane < head) *reinterpret_cast<uint32_t*>(gdst + 4u * lane) = *reinterpret_cast<const uint32_t*>(stg + pad + 4u * lane);
        const unsigned nvec = (L - head) >> 4;
        for (unsigned q = lane; q < nvec; q += 32)
            *reinterpret_cast<uint4*>(gdst + head + 16u * q) = *reinterpret_cast<const uint4*>(stg + pad + head + 16u * q);
        for (unsigned i = head + (nvec << 4) + 2u * lane; i < L; i += 64)
            *reinterpret_cast<uint16_t*>(gdst + i) = *reinterpret_cast<const uint16_t*>(stg + pad + i);
        __syncwarp();
    }
}

// length (<= 63) and winner of every game in one byte: bits 0..5 length, bits 6..7 winner + 1.
// Halves the device->host traffic of the per-game results (16 bytes per thread in, 16 out).
// WIDE (boards of 64..127 cells, games from the empty board): bits 0..6 length, bit 7 = draw; the winner of a
// decided game follows from the parity of its length (odd = player 0).
template <bool WIDE>
__global__ void __launch_bounds__(256)
pack_results_kernel(unsigned long long n, const uint8_t* __restrict__ length, const int8_t* __restrict__ winner,
                    uint8_t* __restrict__ packed) {
    const unsigned long long nvec = n / 16ull;
    for (unsigned long long v = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; v < nvec;
         v += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 l = reinterpret_cast<const uint4*>(length)[v];
        const uint4 w = reinterpret_cast<const uint4*>(winner)[v];
        // per byte: (winner + 1) << 6 | length
        // (a winner byte of -1 is 0xFF: reduce it to 2 bits BEFORE adding 1, or the carry crosses into the next byte)
        auto f = [](uint32_t lw, uint32_t ww) {
            if (WIDE) return (ww & 0x80808080u) | (lw & 0x7F7F7F7Fu);  // winner -1 = 0xFF: its sign bit is the draw flag
            return ((((ww & 0x03030303u) + 0x01010101u) & 0x03030303u) << 6) | (lw & 0x3F3F3F3Fu);
        };
        reinterpret_cast<uint4*>(packed)[v] = make_uint4(f(l.x, w.x), f(l.y, w.y), f(l.z, w.z), f(l.w, w.w));
    }
    if (blockIdx.x == 0)
        for (unsigned long long i = nvec * 16ull + threadIdx.x; i < n; i += blockDim.x)
            packed[i] = WIDE ? (uint8_t)(((uint8_t)winner[i] & 0x80u) | (length[i] & 0x7Fu))
                             : (uint8_t)((((uint32_t)(winner[i] + 1) & 3u) << 6) | (length[i] & 63u));
}

// Dense per-game results: a game from the empty board ends after Lmin = 2K-1 .. H*W plies with a winner given
// by the parity of its length, or in a draw -- S = H*W - Lmin + 2 symbols.  G = floor(16 / log2 S) games share
// one 16-bit word, word = s0 + S*s1 + S^2*s2 + ..  (6x7x4: S = 37, G = 3: 5.33 bits per game instead of 8).
__global__ void __launch_bounds__(256)
pack_results_dense_kernel(unsigned long long n, const uint8_t* __restrict__ length, const int8_t* __restrict__ winner,
                          uint16_t* __restrict__ packed, int lmin, int S, int G) {
    const unsigned long long nwords = (n + (unsigned)G - 1ull) / (unsigned)G;
    unsigned long long w0 = 0;
    if (G == 3 && (((uintptr_t)length | (uintptr_t)winner | (uintptr_t)packed) & 15u) == 0) {
        // 48 games -> 16 words per thread: three 128-bit loads of each input, two 128-bit stores
        const unsigned long long nblk = n / 48ull;
        for (unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b < nblk;
             b += (unsigned long long)gridDim.x * blockDim.x) {
            uint32_t lw[12], ww[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const uint4 l = reinterpret_cast<const uint4*>(length)[3 * b + q];
                const uint4 w = reinterpret_cast<const uint4*>(winner)[3 * b + q];
                lw[4 * q] = l.x; lw[4 * q + 1] = l.y; lw[4 * q + 2] = l.z; lw[4 * q + 3] = l.w;
                ww[4 * q] = w.x; ww[4 * q + 1] = w.y; ww[4 * q + 2] = w.z; ww[4 * q + 3] = w.w;
            }
            uint32_t out[8];
#pragma unroll
            for (int k = 0; k < 16; ++k) {  // word k = games 3k .. 3k+2 of the block
                uint32_t code = 0, mul = 1;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int g = 3 * k + j;
                    const uint32_t len = (lw[g >> 2] >> (8 * (g & 3))) & 0xFFu;
                    const bool draw = ((ww[g >> 2] >> (8 * (g & 3))) & 0x80u) != 0u;
                    code += (draw ? (uint32_t)(S - 1) : len - (uint32_t)lmin) * mul;
                    mul *= (uint32_t)S;
                }
                if (k & 1) out[k >> 1] |= code << 16; else out[k >> 1] = code;
            }
            uint4* dst = reinterpret_cast<uint4*>(packed) + 2 * b;
            dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
            dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
        }
        w0 = nblk * 16ull;  // the remaining words go through the generic loop below
    }
    for (unsigned long long w = w0 + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; w < nwords;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t code = 0, mul = 1;
        for (int j = 0; j < G; ++j) {
            const unsigned long long i = w * (unsigned)G + j;
            uint32_t sym = 0;
            if (i < n) sym = winner[i] < 0 ? (uint32_t)(S - 1) : (uint32_t)((int)length[i] - lmin);
            code += sym * mul;
            mul *= (uint32_t)S;
        }
        packed[w] = (uint16_t)code;
    }
}

__global__ void __launch_bounds__(256)
reward_kernel(unsigned long long n, const int8_t* __restrict__ winner, float2* __restrict__ reward) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        reward[i] = reward_of(winner[i]);
}

// ---------------------------------------------------------------------------------------------
// batched single step / query on reference-layout states (int8 grids)
// ---------------------------------------------------------------------------------------------
struct BoardBits {
    u128 p[2];
    uint32_t legal;  // bit c set <=> column c not full
    uint64_t hts;    // nibble c = stones in column c
    bool full;
};

__device__ __forceinline__ BoardBits load_grid(const uint8_t* g, int H, int W) {
    BoardBits b;
    b.p[0] = 0; b.p[1] = 0; b.legal = 0; b.hts = 0;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const int v = (int8_t)g[r * W + c];
            const int bit = (H - 1 - r) * W + c;
            if (v == 0) b.p[0] |= (u128)1 << bit;
            else if (v == 1) b.p[1] |= (u128)1 << bit;
            if (v >= 0) b.hts += 1ull << (4 * c);
        }
    for (int c = 0; c < W; ++c)
        if ((int8_t)g[(H - 1) * W + c] < 0) b.legal |= 1u << c;
    b.full = b.legal == 0;
    return b;
}

// One warp per 32 consecutive states: the 32*H*W grid bytes are staged through shared memory with
// 128-bit loads / stores (a row of 42 bytes is not 16-byte aligned, 32 rows are); lane l then works on
// state g0+l out of shared memory.
constexpr int STEP_THREADS = 256;

// Weighted action choice (bgs_connect_sample_step): the caller's `random.choices(actions, weights)` of the
// reference's agent loop (textual/examples/arena.py:64-68, agent.py:58-67) inside the transition kernel.
//   weights  w_c = probs[i, c] on the playable columns (NaN / negative / zero -> 0, +inf -> FLT_MAX)
//   integers q_c = (uint32)(w_c / max_c w_c * 65535 + 0.5)  (IEEE single, no contraction); all w_c zero -> q_c = 1
//   draw     r = Philox4x32-10(key = seed, ctr = (id_lo, id_hi, t >> 2, 0))[t & 3],  t = draw_index[i] or, if
//            that is NULL, the number of stones on the board (= plies played from the empty board)
//   choice   the first playable column j (ascending) with  (q_0 + .. + q_j) * 2^32 > r * sum(q)
// With equal weights this is exactly the uniform rule of the rollout kernels, column mulhi32(r, n_legal).
struct StepPolicy {
    const float* probs;           // [n, W] or null (actions come from `action`)
    const uint64_t* game_ids;     // [n] or null (global id = game_id0 + i)
    const int32_t* draw_index;    // [n] or null
    unsigned long long game_id0;
    uint32_t seed_lo, seed_hi;
    int32_t* action_out;          // [n] the chosen column (-1: the state had ended), or null
};

__device__ __forceinline__ uint32_t quantize_weight(float w, float wmax) {
    return (uint32_t)__fadd_rn(__fmul_rn(__fdiv_rn(w, wmax), 65535.0f), 0.5f);
}
__device__ __forceinline__ float sane_weight(float w) { return w > 0.0f ? fminf(w, 3.402823466e+38f) : 0.0f; }

__global__ void __launch_bounds__(STEP_THREADS)
connect_step_kernel(const DynGeo g, unsigned long long n, const int8_t* __restrict__ grid,
                    const int8_t* __restrict__ player, const int8_t* __restrict__ winner,
                    const int32_t* __restrict__ action, int8_t* grid_out, int8_t* player_out,
                    int8_t* winner_out, uint8_t* ended_out, float* reward_out, uint32_t* legal_out,
                    int32_t* status, bool vec, bool dbuf, const StepPolicy pol) {
    extern __shared__ __align__(16) uint8_t s_stage[];
    const int H = g.H(), W = g.W(), HW = H * W;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARPS = STEP_THREADS / 32;
    // dbuf (aligned pointers, room for two tiles per warp): the NEXT group's grids are fetched with cp.async while this
    // one is processed and written back -- the kernel was waiting on its global loads (ncu: 7 warps per issue stalled on
    // the long scoreboard, 0.62 of the copy peak) because a warp had one tile in flight at a time
    uint8_t* buf0 = s_stage + (size_t)warp * 32 * HW * (dbuf ? 2 : 1);
    uint8_t* buf1 = buf0 + (dbuf ? 32 * HW : 0);
    const unsigned long long ngroups = (n + 31ull) / 32ull;
    const unsigned long long stride = (unsigned long long)gridDim.x * WARPS;
    const unsigned long long first = (unsigned long long)blockIdx.x * WARPS + warp;
    const uint8_t* gin = reinterpret_cast<const uint8_t*>(grid);
    // a full group is 32 * HW bytes: a multiple of 16 at a multiple of 16; only the last group can be ragged
    auto full = [&](unsigned long long grp) { return (grp + 1ull) * 32ull <= n; };
    if (dbuf && first < ngroups && full(first)) warp_copy_async(buf0, gin + first * 32ull * (unsigned)HW, 32u * (unsigned)HW, lane);
    cp_async_commit();
    unsigned phase = 0;
    for (unsigned long long group = first; group < ngroups; group += stride, phase ^= 1u) {
        const unsigned long long g0 = group * 32ull;
        const unsigned rows = (unsigned)((n - g0) < 32ull ? (n - g0) : 32ull);
        const unsigned span = rows * (unsigned)HW;
        uint8_t* st_ = phase ? buf1 : buf0;
        uint8_t* mine = st_ + lane * HW;
        const unsigned long long i = g0 + lane;
        // the per-state scalars do not depend on the tile: their loads are in flight while the tile arrives
        int pl = 0, win = 0;
        if (dbuf && i < n) { pl = player[i]; win = winner[i]; }
        if (dbuf) {
            const unsigned long long nxt = group + stride;
            if (nxt < ngroups && full(nxt)) warp_copy_async(phase ? buf0 : buf1, gin + nxt * 32ull * (unsigned)HW, 32u * (unsigned)HW, lane);
            cp_async_commit();
            if (full(group)) cp_async_wait<1>();  // this group's tile (all but the newest copy group)
            else warp_copy(st_, gin + g0 * (unsigned)HW, span, lane, vec);  // the ragged last group: synchronous
        } else {
            warp_copy(st_, gin + g0 * (unsigned)HW, span, lane, vec);
        }
        __syncwarp();
        if (!dbuf && i < n) { pl = player[i]; win = winner[i]; }
        if (i < n) {
            // Work on the staged bytes directly (empty = 0xFF): the transition only needs the height
            // of one column, the top row and the <= 8*(K-1) cells around the new stone -- converting
            // the whole grid to bitboards cost ~15 instructions per cell and made the kernel
            // instruction-bound (4 Mi 6x7 states: 0.41 ms against 0.07 ms of HBM time).
            uint32_t legal_mask = 0;
            for (int c = 0; c < W; ++c)
                if (mine[(H - 1) * W + c] == 0xFFu) legal_mask |= 1u << c;
            const bool ended = win >= 0 || legal_mask == 0;
            int col = -1;
            if (pol.probs) {
                if (!ended) {
                    const float* pw = pol.probs + i * (unsigned)W;
                    float wmax = 0.0f;
                    for (int c = 0; c < W; ++c)
                        if ((legal_mask >> c) & 1u) wmax = fmaxf(wmax, sane_weight(pw[c]));
                    uint64_t total = 0;
                    for (int c = 0; c < W; ++c)
                        if ((legal_mask >> c) & 1u) total += wmax > 0.0f ? quantize_weight(sane_weight(pw[c]), wmax) : 1u;
                    uint32_t t;
                    if (pol.draw_index) {
                        t = (uint32_t)pol.draw_index[i];
                    } else {
                        t = 0;
                        for (int c = 0; c < HW; ++c) t += mine[c] != 0xFFu;
                    }
                    const unsigned long long gid = pol.game_ids ? pol.game_ids[i] : pol.game_id0 + i;
                    uint32_t r4[4];
                    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), t >> 2, DOMAIN_CONNECT, pol.seed_lo, pol.seed_hi, r4);
                    const uint32_t r = (t & 3u) == 0 ? r4[0] : ((t & 3u) == 1 ? r4[1] : ((t & 3u) == 2 ? r4[2] : r4[3]));
                    const uint64_t thresh = (uint64_t)r * total;
                    uint64_t cum = 0;
                    for (int c = 0; c < W; ++c)
                        if ((legal_mask >> c) & 1u) {
                            cum += wmax > 0.0f ? quantize_weight(sane_weight(pw[c]), wmax) : 1u;
                            if ((cum << 32) > thresh) { col = c; break; }
                        }
                }
                if (pol.action_out) pol.action_out[i] = col;
            } else {
                col = action[i];
            }
            const bool legal = !ended && col >= 0 && col < W && ((legal_mask >> col) & 1u) && (pl == 0 || pl == 1);
            if (legal) {
                int row = 0;  // lowest empty cell of the column
                while (mine[row * W + col] != 0xFFu) ++row;
                mine[row * W + col] = (uint8_t)pl;
                if (row == H - 1) legal_mask &= ~(1u << col);
                const int K = g.K();
                // a run of >= K stones of `pl` through (row, col): horizontal, vertical, two diagonals
                auto run = [&](int dr, int dc) {
                    int cnt = 0, r = row + dr, c = col + dc;
                    while (cnt < K - 1 && r >= 0 && r < H && c >= 0 && c < W && mine[r * W + c] == (uint8_t)pl) {
                        ++cnt; r += dr; c += dc;
                    }
                    return cnt;
                };
                if (1 + run(0, 1) + run(0, -1) >= K || 1 + run(-1, 0) >= K || 1 + run(1, 1) + run(-1, -1) >= K ||
                    1 + run(1, -1) + run(-1, 1) >= K)
                    win = pl;
                pl = 1 - pl;
            }
            player_out[i] = (int8_t)pl;
            winner_out[i] = (int8_t)win;
            const bool ended_new = win >= 0 || legal_mask == 0;
            if (ended_out) ended_out[i] = ended_new;
            if (legal_out) legal_out[i] = ended_new ? 0u : legal_mask;
            if (reward_out) reinterpret_cast<float2*>(reward_out)[i] = reward_of(win);
            if (status) status[i] = legal ? 0 : 1;
        }
        __syncwarp();
        warp_copy(reinterpret_cast<uint8_t*>(grid_out) + g0 * (unsigned)HW, st_, span, lane, vec);
        __syncwarp();
    }
}

// Start positions in the reference layout (int8 grids) -> start records of the START rollout.
__global__ void __launch_bounds__(STEP_THREADS)
connect_import_kernel(int H, int W, unsigned long long n, const int8_t* __restrict__ grid,
                      const int8_t* __restrict__ player, const int8_t* __restrict__ winner_in, uint64_t* rec,
                      bool vec) {
    extern __shared__ __align__(16) uint8_t s_stage[];
    const int HW = H * W;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* st_ = s_stage + (size_t)warp * 32 * HW;
    const uint8_t* mine = st_ + lane * HW;
    const unsigned long long ngroups = (n + 31ull) / 32ull;
    constexpr int WARPS = STEP_THREADS / 32;
    for (unsigned long long group = (unsigned long long)blockIdx.x * WARPS + warp; group < ngroups;
         group += (unsigned long long)gridDim.x * WARPS) {
        const unsigned long long g0 = group * 32ull;
        const unsigned rows = (unsigned)((n - g0) < 32ull ? (n - g0) : 32ull);
        warp_copy(st_, reinterpret_cast<const uint8_t*>(grid) + g0 * (unsigned)HW, rows * (unsigned)HW, lane, vec);
        __syncwarp();
        const unsigned long long i = g0 + lane;
        if (i < n) {
            const BoardBits b = load_grid(mine, H, W);
            const int w_in = winner_in ? (int)winner_in[i] : -1;
            const uint64_t meta =
                (uint64_t)(player[i] & 1) | (w_in >= 0 ? 2ull : 0ull) | ((uint64_t)((w_in + 1) & 0xFF) << 8);
            uint64_t* out = rec + i * start_words(HW);
            if (HW <= 64) {
                reinterpret_cast<ulonglong2*>(out)[0] = make_ulonglong2((uint64_t)b.p[0], (uint64_t)b.p[1]);
                reinterpret_cast<ulonglong2*>(out)[1] = make_ulonglong2(b.hts, meta);
            } else {
                reinterpret_cast<ulonglong2*>(out)[0] = make_ulonglong2((uint64_t)b.p[0], (uint64_t)(b.p[0] >> 64));
                reinterpret_cast<ulonglong2*>(out)[1] = make_ulonglong2((uint64_t)b.p[1], (uint64_t)(b.p[1] >> 64));
                reinterpret_cast<ulonglong2*>(out)[2] = make_ulonglong2(b.hts, meta);
            }
        }
        __syncwarp();
    }
}

// Reference-layout states -> packed positions: the two bitboards in the public packed-board format
// (bgs_connect_packed_words) + one meta byte (bit 0 = side to move, bits 1..2 = winner + 1).  17 bytes per
// 6x7 position instead of 44: what a host sends over PCIe for rollouts from positions.
__global__ void __launch_bounds__(STEP_THREADS)
connect_pack_kernel(int H, int W, unsigned long long n, const int8_t* __restrict__ grid,
                    const int8_t* __restrict__ player, const int8_t* __restrict__ winner, uint64_t* packed,
                    uint8_t* meta, bool vec) {
    extern __shared__ __align__(16) uint8_t s_stage[];
    const int HW = H * W;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* st_ = s_stage + (size_t)warp * 32 * HW;
    const uint8_t* mine = st_ + lane * HW;
    const unsigned long long ngroups = (n + 31ull) / 32ull;
    constexpr int WARPS = STEP_THREADS / 32;
    for (unsigned long long group = (unsigned long long)blockIdx.x * WARPS + warp; group < ngroups;
         group += (unsigned long long)gridDim.x * WARPS) {
        const unsigned long long g0 = group * 32ull;
        const unsigned rows = (unsigned)((n - g0) < 32ull ? (n - g0) : 32ull);
        warp_copy(st_, reinterpret_cast<const uint8_t*>(grid) + g0 * (unsigned)HW, rows * (unsigned)HW, lane, vec);
        __syncwarp();
        const unsigned long long i = g0 + lane;
        if (i < n) {
            const BoardBits b = load_grid(mine, H, W);
            if (HW <= 64) {
                *reinterpret_cast<ulonglong2*>(packed + i * 2) = make_ulonglong2((uint64_t)b.p[0], (uint64_t)b.p[1]);
            } else {
                ulonglong2* dst = reinterpret_cast<ulonglong2*>(packed + i * 4);
                dst[0] = make_ulonglong2((uint64_t)b.p[0], (uint64_t)(b.p[0] >> 64));
                dst[1] = make_ulonglong2((uint64_t)b.p[1], (uint64_t)(b.p[1] >> 64));
            }
            const int w_in = winner ? (int)winner[i] : -1;
            meta[i] = (uint8_t)((player[i] & 1) | (((w_in + 1) & 3) << 1));
        }
        __syncwarp();
    }
}

// Packed positions -> start records of the START rollout (column heights = stones per column).
__global__ void __launch_bounds__(256)
connect_import_packed_kernel(int H, int W, unsigned long long n, const uint64_t* __restrict__ packed,
                             const uint8_t* __restrict__ meta, uint64_t* rec) {
    const int HW = H * W;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        u128 b0, b1;
        if (HW <= 64) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(packed + i * 2);
            b0 = v.x; b1 = v.y;
        } else {
            const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(packed + i * 4);
            const ulonglong2 v1 = *reinterpret_cast<const ulonglong2*>(packed + i * 4 + 2);
            b0 = ((u128)v0.y << 64) | v0.x; b1 = ((u128)v1.y << 64) | v1.x;
        }
        const u128 board = HW >= 128 ? ~(u128)0 : (((u128)1 << HW) - 1);
        b0 &= board; b1 &= board & ~b0;  // a cell holds at most one stone
        const u128 occ = b0 | b1;
        uint64_t hts = 0;
        for (int c = 0; c < W; ++c) {  // stones in column c = occupied cells (H-1-row)*W + c over all rows
            int cnt = 0;
            for (int r = 0; r < H; ++r) cnt += (int)((occ >> (r * W + c)) & 1);
            hts |= (uint64_t)cnt << (4 * c);
        }
        const unsigned mb = meta[i];
        const int w_in = (int)((mb >> 1) & 3u) - 1;
        const uint64_t m = (uint64_t)(mb & 1u) | (w_in >= 0 ? 2ull : 0ull) | ((uint64_t)((w_in + 1) & 0xFF) << 8);
        uint64_t* out = rec + i * start_words(HW);
        if (HW <= 64) {
            reinterpret_cast<ulonglong2*>(out)[0] = make_ulonglong2((uint64_t)b0, (uint64_t)b1);
            reinterpret_cast<ulonglong2*>(out)[1] = make_ulonglong2(hts, m);
        } else {
            reinterpret_cast<ulonglong2*>(out)[0] = make_ulonglong2((uint64_t)b0, (uint64_t)(b0 >> 64));
            reinterpret_cast<ulonglong2*>(out)[1] = make_ulonglong2((uint64_t)b1, (uint64_t)(b1 >> 64));
            reinterpret_cast<ulonglong2*>(out)[2] = make_ulonglong2(hts, m);
        }
    }
}

__global__ void __launch_bounds__(128)
connect_query_kernel(int H, int W, unsigned long long n, const int8_t* __restrict__ grid,
                     const int8_t* __restrict__ winner, uint8_t* ended_out, uint32_t* legal_out,
                     float* reward_out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int8_t* gi = grid + i * (unsigned long long)(H * W);
    uint32_t legal = 0;
    for (int c = 0; c < W; ++c)
        if (gi[(H - 1) * W + c] < 0) legal |= 1u << c;
    const int win = winner[i];
    const bool ended = win >= 0 || legal == 0;
    if (ended_out) ended_out[i] = ended;
    if (legal_out) legal_out[i] = ended ? 0u : legal;
    if (reward_out) reinterpret_cast<float2*>(reward_out)[i] = reward_of(win);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// boards the bit-word kernels cover
static bool bitboard_supported(int H, int W, int K) {
    return H >= 1 && W >= 1 && K >= 1 && H <= 15 && W <= 16 && H * W <= 128;
}
// ... and the byte-board fallback beyond them (length is a uint8, the playable columns a 32-bit mask)
static bool supported(int H, int W, int K) {
    return H >= 1 && W >= 1 && K >= 1 && W <= 32 && H * W <= 255;
}

template <typename Kern, typename... Args>
static int launch_persistent(Kern kern, const RolloutParams& p, cudaStream_t stream, Args... args) {
    int per_sm = 0;
    BGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ROLLOUT_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    // persistent launch: exactly the resident CTAs, a multiple of the SM count
    unsigned long long want = ((unsigned long long)p.n_games + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS;
    unsigned long long blocks = (unsigned long long)sm_count() * per_sm;
    if (want < blocks) blocks = want ? want : 1;
    kern<<<(unsigned)blocks, ROLLOUT_THREADS, 0, stream>>>(args..., p);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

// trajectory mode for a board: packed 16-bit blocks in the game's own row need an even row length
// of at least 4 bytes; other boards store one byte per ply into the pre-filled row
static int actions_mode(int H, int W) { return ((H * W) % 2 == 0 && H * W >= 4) ? 2 : 1; }

template <class G, bool START, bool OPEN>
static int launch_rollout_s(const G& g, const RolloutParams& p, cudaStream_t stream) {
    const int act = p.actions ? actions_mode(g.H(), g.W()) : 0;
    if (p.final_packed) {
        if (act == 2) return launch_persistent(connect_rollout_kernel<G, 2, true, START, OPEN>, p, stream, g);
        if (act == 1) return launch_persistent(connect_rollout_kernel<G, 1, true, START, OPEN>, p, stream, g);
        return launch_persistent(connect_rollout_kernel<G, 0, true, START, OPEN>, p, stream, g);
    }
    if (act == 2) return launch_persistent(connect_rollout_kernel<G, 2, false, START, OPEN>, p, stream, g);
    if (act == 1) return launch_persistent(connect_rollout_kernel<G, 1, false, START, OPEN>, p, stream, g);
    return launch_persistent(connect_rollout_kernel<G, 0, false, START, OPEN>, p, stream, g);
}

template <class G>
static int launch_rollout(const G& g, const RolloutParams& p, cudaStream_t stream) {
    if (p.start) return launch_rollout_s<G, true, false>(g, p, stream);
    // the opening phase plays 8 plies unconditionally: the board must not be able to fill up in them
    if (g.H() * g.W() >= 12) return launch_rollout_s<G, false, true>(g, p, stream);
    return launch_rollout_s<G, false, false>(g, p, stream);
}

template <int H, int W, int K>
static int launch_rollout_lut(const RolloutParams& p, cudaStream_t stream) {
    static_assert((H * W) % 2 == 0 && H * W >= 4, "LUT kernel stores packed trajectory blocks");
    if (p.actions && p.final_packed) return launch_persistent(connect_rollout_lut_kernel<H, W, K, true, true>, p, stream);
    if (p.actions) return launch_persistent(connect_rollout_lut_kernel<H, W, K, true, false>, p, stream);
    if (p.final_packed) return launch_persistent(connect_rollout_lut_kernel<H, W, K, false, true>, p, stream);
    return launch_persistent(connect_rollout_lut_kernel<H, W, K, false, false>, p, stream);
}

template <typename Kern>
static int launch_persistent_n(Kern kern, int threads, const RolloutParams& p, cudaStream_t stream) {
    int per_sm = 0;
    BGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
    if (per_sm < 1) per_sm = 1;
    unsigned long long want = ((unsigned long long)p.n_games + threads - 1) / threads;
    unsigned long long blocks = (unsigned long long)sm_count() * per_sm;
    if (want < blocks) blocks = want ? want : 1;
    kern<<<(unsigned)blocks, threads, 0, stream>>>(p);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

template <int H, int W, int K>
static int launch_rollout_lines(const RolloutParams& p, cudaStream_t stream, bool fused = false) {
    if (fused) {  // trajectory rows / final grids / rewards straight from the rollout kernel
        if constexpr ((H * W) % 8 == 0) {
            if (p.actions && p.final_grid) return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 3, false, true>, LINES_THREADS, p, stream);
            if (p.actions) return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 3, false, false>, LINES_THREADS, p, stream);
            if (p.final_grid) return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 0, false, true>, LINES_THREADS, p, stream);
            return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 0, false, false>, LINES_THREADS, p, stream);
        } else {
            return set_error(BGS_EUNSUPPORTED, "connect: no fused export for this board");
        }
    }
    const int act = p.actions ? actions_mode(H, W) : 0;
    if (p.final_packed) {
        if (act == 2) return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 2, true>, LINES_THREADS, p, stream);
        if (act == 1) return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 1, true>, LINES_THREADS, p, stream);
        return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 0, true>, LINES_THREADS, p, stream);
    }
    if (act == 2) return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 2, false>, LINES_THREADS, p, stream);
    if (act == 1) return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 1, false>, LINES_THREADS, p, stream);
    return launch_persistent_n(connect_rollout_lines_kernel<H, W, K, 0, false>, LINES_THREADS, p, stream);
}

template <int MODE, int SH = 0, int SW = 0>
static int launch_export_rows(int H, int W, unsigned long long n, const uint64_t* packed, const uint8_t* length,
                              uint8_t* out, cudaStream_t stream, const uint8_t* actions = nullptr) {
    if (MODE != MODE_ACTIONS && SH == 0) {  // compile-time boards of the BASELINE configurations
        if (H == 6 && W == 7) return launch_export_rows<MODE, 6, 7>(H, W, n, packed, length, out, stream, actions);
        if (H == 8 && W == 9) return launch_export_rows<MODE, 8, 9>(H, W, n, packed, length, out, stream, actions);
        if (H == 10 && W == 12) return launch_export_rows<MODE, 10, 12>(H, W, n, packed, length, out, stream, actions);
    }
    if (MODE == MODE_ACTIONS && ((uintptr_t)out & 15u) == 0 && (H * W == 42 || H * W == 72 || H * W == 120)) {
        auto launch = [&](auto kern) {
            int per_sm = 0;
            BGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EXPORT_THREADS, 0));
            if (per_sm < 1) per_sm = 1;
            unsigned long long blocks = (n + 255ull) / 256ull;
            const unsigned long long cap = (unsigned long long)sm_count() * per_sm;
            if (blocks > cap) blocks = cap;
            kern<<<(unsigned)blocks, EXPORT_THREADS, 0, stream>>>(n, length, out);
            BGS_CUDA_TRY(cudaGetLastError());
            return (int)BGS_OK;
        };
        if (H * W == 42) return launch(connect_expand_actions_kernel<42>);
        if (H * W == 72) return launch(connect_expand_actions_kernel<72>);
        return launch(connect_expand_actions_kernel<120>);
    }
    // MODE_TRAJ: as many rows per lane (4 / 2 / 1) as fit the 48 KB of static-limit shared memory
    int rpl = 1;
    if (MODE == MODE_TRAJ)
        for (int r = 4; r >= 1; r >>= 1)
            if ((size_t)(EXPORT_THREADS / 32) * 32 * r * H * W <= 48 * 1024) { rpl = r; break; }
    const size_t smem = (size_t)(EXPORT_THREADS / 32) * 32 * rpl * H * W;
    auto kern = connect_export_rows_kernel<MODE, SH, SW>;
    int per_sm = 0;
    BGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EXPORT_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long blocks = (n + 256ull * rpl - 1ull) / (256ull * rpl);
    const unsigned long long cap = (unsigned long long)sm_count() * per_sm;  // one resident wave, grid-stride
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, EXPORT_THREADS, smem, stream>>>(H, W, n, packed, length, out, ((uintptr_t)out & 15u) == 0,
                                                            actions, rpl);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

}  // namespace connect
}  // namespace bgs

using namespace bgs;
using namespace bgs::connect;

extern "C" int bgs_connect_supported(int H, int W, int K) { return supported(H, W, K) ? 1 : 0; }

extern "C" int bgs_connect_packed_words(int H, int W) {
    if (!bitboard_supported(H, W, 1)) return (H * W + 7) / 8;  // byte boards: the grid itself, padded to 8 bytes
    return 2 * (H * W <= 64 ? 1 : 2);
}

// Boards whose rollout kernel also writes trajectories / final grids / rewards in the reference's
// layouts (fused export): the line kernel of the two larger BASELINE boards.
static bool fused_export_board(int H, int W, int K) {
    return (H == 8 && W == 9 && K == 5) || (H == 10 && W == 12 && K == 6);
}

// `fused`: final_grid / reward (and the trajectories) come straight from the rollout kernel
// (fused_export_board only; the caller checked the pointer alignment).
static int rollout_impl(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                        const uint64_t* start, uint8_t* actions, uint8_t* length, int8_t* winner,
                        uint64_t* final_packed, int64_t* stats, void* stream_, bool fused = false,
                        int8_t* final_grid = nullptr) {
    if (!supported(H, W, K)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d k=%d", H, W, K);
    if (int rc = require_device()) return rc;
    if (n_games == 0) return BGS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t HW = (size_t)H * W;
    const int PW = bgs_connect_packed_words(H, W);
    // the kernels always write length and winner: point the unwanted one at write-only scratch
    uint8_t* scratch = nullptr;
    const uint64_t max_launch = 1ull << 31;
    const uint64_t per_launch = n_games < max_launch ? n_games : max_launch;
    if (actions && !length) return set_error(BGS_EINVAL, "connect_rollout: `actions` requires `length`");
    if (!length || !winner) {
        if (int rc = scratch_buffer((size_t)(2 * per_launch), (void**)&scratch)) return rc;
    }
    unsigned int* counter = nullptr;
    int rc = next_counter(&counter);
    cudaError_t e = cudaSuccess;
    const bool bytes_board = !bitboard_supported(H, W, K);
    if (bytes_board && start) return set_error(BGS_EUNSUPPORTED, "connect: rollouts from positions need a board of at most 128 cells, 16 columns, 15 rows (%dx%d)", H, W);
    const int act_mode = actions ? (fused ? 3 : (bytes_board ? 1 : actions_mode(H, W))) : 0;
    if (act_mode == 2 && ((uintptr_t)actions & 1u) != 0)  // 4-ply blocks are stored as 16-bit words
        return set_error(BGS_EINVAL, "connect_rollout: `actions` must be 2-byte aligned");
    if (!bytes_board && ((((uintptr_t)final_packed) | (uintptr_t)start) & 15u) != 0)  // records move as 16-byte pairs
        return set_error(BGS_EINVAL, "connect_rollout: `final_packed` / the start-record workspace must be 16-byte aligned");
    if (rc == BGS_OK && act_mode == 1) {
        e = cudaMemsetAsync(actions, 0xFF, n_games * HW, stream);
        if (e != cudaSuccess) rc = cuda_error(e, "cudaMemsetAsync");
    }
    for (uint64_t off = 0; rc == BGS_OK && off < n_games; off += max_launch) {
        RolloutParams p;
        p.n_games = (uint32_t)((n_games - off) < max_launch ? (n_games - off) : max_launch);
        p.game_id0 = game_id0 + off;
        p.seed_lo = (uint32_t)seed;
        p.seed_hi = (uint32_t)(seed >> 32);
        p.actions = actions ? actions + off * HW : nullptr;
        p.length = length ? length + off : scratch;
        p.winner = winner ? winner + off : reinterpret_cast<int8_t*>(scratch + per_launch);
        p.final_packed = final_packed ? final_packed + off * PW : nullptr;
        p.stats = reinterpret_cast<unsigned long long*>(stats);
        p.counter = counter;
        p.one = 1u;
        p.start = start ? start + off * start_words((int)HW) : nullptr;
        p.final_grid = (fused && final_grid) ? final_grid + off * HW : nullptr;
        e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess) { rc = cuda_error(e, "cudaMemsetAsync"); break; }
        if (bytes_board) {
            const unsigned long long want = ((unsigned long long)p.n_games + BYTES_THREADS - 1) / BYTES_THREADS;
            unsigned long long blocks = (unsigned long long)sm_count() * 8;
            if (want < blocks) blocks = want;
            connect_rollout_bytes_kernel<<<(unsigned)blocks, BYTES_THREADS, 0, stream>>>(H, W, K, p);
            e = cudaGetLastError();
            if (e != cudaSuccess) rc = cuda_error(e, "connect_rollout_bytes_kernel");
        }
        else if (H == 6 && W == 7 && K == 4 && !start) rc = launch_rollout_lut<6, 7, 4>(p, stream);
        else if (H == 6 && W == 7 && K == 4) rc = launch_rollout(StaticGeo<6, 7, 4>(), p, stream);
        else if (H == 8 && W == 9 && K == 5 && !start) rc = launch_rollout_lines<8, 9, 5>(p, stream, fused);
        else if (H == 10 && W == 12 && K == 6 && !start) rc = launch_rollout_lines<10, 12, 6>(p, stream, fused);
        else if (H == 8 && W == 9 && K == 5) rc = launch_rollout(StaticGeo<8, 9, 5>(), p, stream);
        else if (H == 10 && W == 12 && K == 6) rc = launch_rollout(StaticGeo<10, 12, 6>(), p, stream);
        else rc = launch_rollout(make_dyn_geo(H, W, K), p, stream);
        // packed trajectory blocks -> one byte per ply, in place (needs this launch's lengths)
        if (rc == BGS_OK && act_mode == 2)
            rc = launch_export_rows<MODE_ACTIONS>(H, W, p.n_games, nullptr, p.length, p.actions, stream);
    }
    return rc;
}

extern "C" int bgs_connect_rollout(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                                   uint8_t* actions, uint8_t* length, int8_t* winner,
                                   uint64_t* final_packed, int64_t* stats, void* stream_) {
    return rollout_impl(H, W, K, n_games, game_id0, seed, nullptr, actions, length, winner, final_packed, stats, stream_);
}

extern "C" int bgs_connect_rollout_export(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                                          uint8_t* actions, uint8_t* length, int8_t* winner, int8_t* final_grid,
                                          float* reward, int64_t* stats, void* stream_) {
    if (!supported(H, W, K)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d k=%d", H, W, K);
    if (actions && !length) return set_error(BGS_EINVAL, "connect_rollout_export: `actions` requires `length`");
    if (int rc = check_reward_alignment(reward, "connect_rollout_export")) return rc;
    if (int rc = require_device()) return rc;
    if (n_games == 0) return BGS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    // fused boards: trajectory rows and final grids are written once, by the rollout kernel itself; rewards
    // (a pure function of the winner: 1 byte read, 8 written per game) come from reward_kernel -- measured
    // cheaper than a 64-bit store per retiring lane inside the rollout kernel (9 us against 13 - 27 us per 4 Mi games).
    // Other boards: rollout (packed final boards) + export, with stream-ordered temporaries.
    const bool fused = fused_export_board(H, W, K) && ((uintptr_t)actions & 15u) == 0 && ((uintptr_t)final_grid & 7u) == 0;
    uint64_t* packed = nullptr;
    int8_t* win_tmp = nullptr;
    int rc = BGS_OK;
    if (final_grid && !fused) rc = temp_alloc((void**)&packed, n_games * (size_t)bgs_connect_packed_words(H, W) * 8, stream);
    if (rc == BGS_OK && reward && !winner) rc = temp_alloc((void**)&win_tmp, n_games, stream);
    int8_t* win = winner ? winner : win_tmp;
    if (rc == BGS_OK)
        rc = rollout_impl(H, W, K, n_games, game_id0, seed, nullptr, actions, length, win, packed, stats, stream_, fused,
                          fused ? final_grid : nullptr);
    if (rc == BGS_OK && ((final_grid && !fused) || reward))
        rc = bgs_connect_export(H, W, n_games, packed, win, fused ? nullptr : final_grid, reward, stream_);
    temp_free(packed, stream);
    temp_free(win_tmp, stream);
    return rc;
}

extern "C" int bgs_connect_start_words(int H, int W) { return start_words(H * W); }

extern "C" int bgs_connect_rollout_from(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                                        const int8_t* grid, const int8_t* player, const int8_t* winner_in,
                                        uint64_t* workspace, uint8_t* actions, uint8_t* length, int8_t* winner,
                                        uint64_t* final_packed, int64_t* stats, void* stream_) {
    if (!bitboard_supported(H, W, K)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d k=%d", H, W, K);
    if (!grid || !player || !workspace) return set_error(BGS_EINVAL, "connect_rollout_from: null required pointer");
    if (int rc = require_device()) return rc;
    if (n_games == 0) return BGS_OK;
    {
        const size_t smem = (size_t)(STEP_THREADS / 32) * 32 * H * W;
        unsigned long long blocks = (n_games + STEP_THREADS - 1) / STEP_THREADS;
        const unsigned long long cap = (unsigned long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        connect_import_kernel<<<(unsigned)blocks, STEP_THREADS, smem, (cudaStream_t)stream_>>>(
            H, W, n_games, grid, player, winner_in, workspace, ((uintptr_t)grid & 15u) == 0);
        BGS_CUDA_TRY(cudaGetLastError());
    }
    return rollout_impl(H, W, K, n_games, game_id0, seed, workspace, actions, length, winner, final_packed, stats, stream_);
}

extern "C" int bgs_connect_pack(int H, int W, uint64_t n, const int8_t* grid, const int8_t* player,
                                const int8_t* winner, uint64_t* packed, uint8_t* meta, void* stream_) {
    if (!bitboard_supported(H, W, 1)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d", H, W);
    if (!grid || !player || !packed || !meta) return set_error(BGS_EINVAL, "connect_pack: null required pointer");
    if (((uintptr_t)packed & 15u) != 0) return set_error(BGS_EINVAL, "connect_pack: `packed` must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    const size_t smem = (size_t)(STEP_THREADS / 32) * 32 * H * W;
    unsigned long long blocks = (n + STEP_THREADS - 1) / STEP_THREADS;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    connect_pack_kernel<<<(unsigned)blocks, STEP_THREADS, smem, (cudaStream_t)stream_>>>(
        H, W, n, grid, player, winner, packed, meta, ((uintptr_t)grid & 15u) == 0);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

extern "C" int bgs_connect_rollout_from_packed(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                                               const uint64_t* packed, const uint8_t* meta, uint64_t* workspace,
                                               uint8_t* actions, uint8_t* length, int8_t* winner,
                                               uint64_t* final_packed, int64_t* stats, void* stream_) {
    if (!bitboard_supported(H, W, K)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d k=%d", H, W, K);
    if (!packed || !meta || !workspace) return set_error(BGS_EINVAL, "connect_rollout_from_packed: null required pointer");
    if (((uintptr_t)packed & 15u) != 0) return set_error(BGS_EINVAL, "connect_rollout_from_packed: `packed` must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    if (n_games == 0) return BGS_OK;
    {
        unsigned long long blocks = (n_games + 255) / 256;
        const unsigned long long cap = (unsigned long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        connect_import_packed_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(H, W, n_games, packed, meta, workspace);
        BGS_CUDA_TRY(cudaGetLastError());
    }
    return rollout_impl(H, W, K, n_games, game_id0, seed, workspace, actions, length, winner, final_packed, stats, stream_);
}

extern "C" int bgs_connect_export(int H, int W, uint64_t n, const uint64_t* packed, const int8_t* winner,
                                  int8_t* grid, float* reward, void* stream_) {
    if (!supported(H, W, 1)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d", H, W);
    if (int rc = check_reward_alignment(reward, "connect_export")) return rc;
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int sms = sm_count();
    if (grid) {
        if (!packed) return set_error(BGS_EINVAL, "connect_export: grid requested without packed boards");
        if (((uintptr_t)packed & 15u) != 0 && bitboard_supported(H, W, 1))
            return set_error(BGS_EINVAL, "connect_export: `packed` must be 16-byte aligned");
        if (!bitboard_supported(H, W, 1)) {  // byte boards: the record is the grid, padded to 8 bytes
            BGS_CUDA_TRY(cudaMemcpy2DAsync(grid, (size_t)H * W, packed, (size_t)bgs_connect_packed_words(H, W) * 8,
                                           (size_t)H * W, n, cudaMemcpyDeviceToDevice, stream));
        } else if (int rc = launch_export_rows<MODE_GRID>(H, W, n, packed, nullptr, reinterpret_cast<uint8_t*>(grid), stream))
            return rc;
    }
    if (reward) {
        if (!winner) return set_error(BGS_EINVAL, "connect_export: reward requested without winner");
        unsigned long long blocks = (n + 255) / 256;
        const unsigned long long cap = (unsigned long long)sms * 8 * 4;
        if (blocks > cap) blocks = cap;
        reward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(n, winner, reinterpret_cast<float2*>(reward));
        BGS_CUDA_TRY(cudaGetLastError());
    }
    return BGS_OK;
}

extern "C" int bgs_connect_trajectory_grids(int H, int W, uint64_t n_games, const uint8_t* actions,
                                           const uint8_t* length, int8_t* grids, void* stream_) {
    if (!bitboard_supported(H, W, 1)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d", H, W);
    if (!actions || !length || !grids) return set_error(BGS_EINVAL, "connect_trajectory_grids: null pointer");
    if (int rc = require_device()) return rc;
    if (n_games == 0) return BGS_OK;
    // (the cell kernels read the trajectories as 8-byte pieces, the word kernel as 4-byte words: other pointers take
    // the row kernel below)
    if (((uintptr_t)grids & 15u) == 0 && ((uintptr_t)actions & 7u) == 0 && ((H == 8 && W == 9) || (H == 10 && W == 12))) {
        cudaStream_t stream = (cudaStream_t)stream_;
        auto launch = [&](auto kern, int gpw) {
            int per_sm = 0;
            BGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EXPORT_THREADS, 0));
            if (per_sm < 1) per_sm = 1;
            // TWO resident CTAs per SM, not the five that fit: the kernel is bound by how well its writes stream into
            // HBM, and the fewer warps write at the same time, the closer together their addresses are -- resident
            // CTAs per SM 1 / 2 / 3 / 5: 8x9 0.376 / 0.280 / 0.294 / 0.313 ms per 256 Ki games, 10x12 0.414 / 0.376 /
            // 0.390 / 0.403 ms per 128 Ki games (more registers per thread, prefetching the next group's
            // trajectories or a bulk-copy copy-out changed nothing beyond that)
#ifndef BGS_TRAJ_MAX_PER_SM
#define BGS_TRAJ_MAX_PER_SM 2
#endif
            if (per_sm > BGS_TRAJ_MAX_PER_SM) per_sm = BGS_TRAJ_MAX_PER_SM;
            unsigned long long blocks = (n_games + 8ull * gpw - 1ull) / (8ull * gpw);
            const unsigned long long cap = (unsigned long long)sm_count() * per_sm;
            if (blocks > cap) blocks = cap;
            kern<<<(unsigned)blocks, EXPORT_THREADS, 0, stream>>>(n_games, actions, length, reinterpret_cast<uint8_t*>(grids));
            BGS_CUDA_TRY(cudaGetLastError());
            return (int)BGS_OK;
        };
        if (H == 8) return launch(connect_traj_cells_kernel<8, 9>, 3);
        return launch(connect_traj_cells_kernel<10, 12>, 2);
    }
    if (((uintptr_t)grids & 15u) == 0 && ((uintptr_t)actions & 3u) == 0 && H == 6 && W == 7) {  // the headline board: word-stationary kernel
        auto kern = connect_traj_words_kernel<6, 7>;
        int per_sm = 0;
        BGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TRAJW_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
        // (no cap as in the cell kernels: this one is instruction-bound -- 2 / 3 / 4 / 6 resident CTAs per SM:
        // 1.06 / 0.76 / 0.64 / 0.50 ms against 0.44 ms with all nine)
        unsigned long long blocks = (n_games + 2ull * (TRAJW_THREADS / 32) - 1ull) / (2ull * (TRAJW_THREADS / 32));
        const unsigned long long cap = (unsigned long long)sm_count() * per_sm;
        if (blocks > cap) blocks = cap;
        kern<<<(unsigned)blocks, TRAJW_THREADS, 0, (cudaStream_t)stream_>>>(n_games, actions, length, reinterpret_cast<uint8_t*>(grids));
        BGS_CUDA_TRY(cudaGetLastError());
        return BGS_OK;
    }
    return launch_export_rows<MODE_TRAJ>(H, W, n_games * (unsigned long long)(H * W + 1), nullptr, length,
                                         reinterpret_cast<uint8_t*>(grids), (cudaStream_t)stream_, actions);
}

// (Lmin, S, G) of the dense result code of a board; G = 0: no gain over one byte per game
static void dense_code(int H, int W, int K, int* lmin, int* S, int* G) {
    const int HW = H * W;
    *lmin = 2 * K - 1 < HW ? 2 * K - 1 : HW;
    *S = HW - *lmin + 2;
    int g = 0;
    for (unsigned long long p = 1; p * (unsigned)*S <= 65536ull; p *= (unsigned)*S) ++g;
    *G = (HW <= 255 && g >= 2) ? g : 0;
}

extern "C" int bgs_connect_dense_results(int H, int W, int K, int* lmin, int* symbols) {
    int l, s, g;
    dense_code(H, W, K, &l, &s, &g);
    if (lmin) *lmin = l;
    if (symbols) *symbols = s;
    return g;
}

extern "C" int bgs_connect_pack_results_dense(int H, int W, int K, uint64_t n, const uint8_t* length, const int8_t* winner,
                                              uint16_t* packed, void* stream_) {
    if (!length || !winner || !packed) return set_error(BGS_EINVAL, "connect_pack_results_dense: null pointer");
    int lmin, S, G;
    dense_code(H, W, K, &lmin, &S, &G);
    if (!supported(H, W, K) || G == 0) return set_error(BGS_EUNSUPPORTED, "connect_pack_results_dense: no dense code for %dx%d k=%d", H, W, K);
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    const unsigned long long nwords = (n + G - 1) / G;
    unsigned long long blocks = (nwords / 16ull + 255ull) / 256ull + 1ull;
    const unsigned long long cap = (unsigned long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    pack_results_dense_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(n, length, winner, packed, lmin, S, G);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

static int pack_results_impl(bool wide, uint64_t n, const uint8_t* length, const int8_t* winner, uint8_t* packed, void* stream_);

extern "C" int bgs_connect_pack_results(uint64_t n, const uint8_t* length, const int8_t* winner, uint8_t* packed,
                                        void* stream_) {
    return pack_results_impl(false, n, length, winner, packed, stream_);
}

extern "C" int bgs_connect_pack_results_wide(uint64_t n, const uint8_t* length, const int8_t* winner, uint8_t* packed,
                                             void* stream_) {
    return pack_results_impl(true, n, length, winner, packed, stream_);
}

static int pack_results_impl(bool wide, uint64_t n, const uint8_t* length, const int8_t* winner, uint8_t* packed, void* stream_) {
    if (!length || !winner || !packed) return set_error(BGS_EINVAL, "connect_pack_results: null pointer");
    if ((((uintptr_t)length | (uintptr_t)winner | (uintptr_t)packed) & 15u) != 0)
        return set_error(BGS_EINVAL, "connect_pack_results: pointers must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    unsigned long long blocks = (n / 16ull + 255ull) / 256ull;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (wide) pack_results_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(n, length, winner, packed);
    else pack_results_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(n, length, winner, packed);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

static int step_impl(int H, int W, int K, uint64_t n, const int8_t* grid, const int8_t* player,
                     const int8_t* winner, const int32_t* action, int8_t* grid_out,
                     int8_t* player_out, int8_t* winner_out, uint8_t* ended_out,
                     float* reward_out, uint32_t* legal_out, int32_t* status, void* stream_, const StepPolicy& pol) {
    if (!supported(H, W, K)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d k=%d", H, W, K);
    if (!grid || !player || !winner || !grid_out || !player_out || !winner_out)
        return set_error(BGS_EINVAL, "connect_step: null required pointer");
    if (int rc = check_reward_alignment(reward_out, "connect_step")) return rc;
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    const DynGeo g = make_dyn_geo(H, W, K);
    const bool vec = (((uintptr_t)grid | (uintptr_t)grid_out) & 15u) == 0;
    const bool dbuf = vec && H * W <= 128;  // two tiles per warp (cp.async prefetch of the next group)
    const size_t smem = (size_t)(STEP_THREADS / 32) * 32 * H * W * (dbuf ? 2 : 1);
    if (smem > 48 * 1024)  // large tiles: opt in to the large carve-out
        BGS_CUDA_TRY(cudaFuncSetAttribute(connect_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long blocks = (n + STEP_THREADS - 1) / STEP_THREADS;
    // (unlike the per-ply grid kernels this one does not gain from fewer resident CTAs: 2 / 3 / 4 per SM 0.236 / 0.181 /
    // 0.157 ms per 4 Mi 6x7 states against 0.155 ms with 8; neither do the export / expansion kernels)
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    connect_step_kernel<<<(unsigned)blocks, STEP_THREADS, smem, (cudaStream_t)stream_>>>(
        g, n, grid, player, winner, action, grid_out, player_out, winner_out, ended_out, reward_out,
        legal_out, status, vec, dbuf, pol);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

extern "C" int bgs_connect_step(int H, int W, int K, uint64_t n, const int8_t* grid, const int8_t* player,
                                const int8_t* winner, const int32_t* action, int8_t* grid_out,
                                int8_t* player_out, int8_t* winner_out, uint8_t* ended_out,
                                float* reward_out, uint32_t* legal_out, int32_t* status, void* stream_) {
    if (!action) return set_error(BGS_EINVAL, "connect_step: null required pointer");
    StepPolicy pol{};
    return step_impl(H, W, K, n, grid, player, winner, action, grid_out, player_out, winner_out, ended_out, reward_out,
                     legal_out, status, stream_, pol);
}

extern "C" int bgs_connect_sample_step(int H, int W, int K, uint64_t n, const int8_t* grid, const int8_t* player,
                                       const int8_t* winner, const float* probs, uint64_t seed, uint64_t game_id0,
                                       const uint64_t* game_ids, const int32_t* draw_index, int8_t* grid_out,
                                       int8_t* player_out, int8_t* winner_out, uint8_t* ended_out, float* reward_out,
                                       uint32_t* legal_out, int32_t* action_out, int32_t* status, void* stream_) {
    if (!probs) return set_error(BGS_EINVAL, "connect_sample_step: null `probs`");
    StepPolicy pol{};
    pol.probs = probs; pol.game_ids = game_ids; pol.draw_index = draw_index; pol.game_id0 = game_id0;
    pol.seed_lo = (uint32_t)seed; pol.seed_hi = (uint32_t)(seed >> 32); pol.action_out = action_out;
    return step_impl(H, W, K, n, grid, player, winner, nullptr, grid_out, player_out, winner_out, ended_out, reward_out,
                     legal_out, status, stream_, pol);
}

extern "C" int bgs_connect_query(int H, int W, uint64_t n, const int8_t* grid, const int8_t* winner,
                                 uint8_t* ended_out, uint32_t* legal_out, float* reward_out, void* stream_) {
    if (!supported(H, W, 1)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d", H, W);
    if (!grid || !winner) return set_error(BGS_EINVAL, "connect_query: null required pointer");
    if (int rc = check_reward_alignment(reward_out, "connect_query")) return rc;
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    const unsigned long long blocks = (n + 127) / 128;
    connect_query_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream_>>>(H, W, n, grid, winner, ended_out,
                                                                             legal_out, reward_out);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

extern "C" int bgs_connect_rollout_host(int device, int H, int W, int K, uint64_t n, uint64_t game_id0,
                                        uint64_t seed, uint8_t* actions, uint8_t* length, int8_t* winner,
                                        int8_t* final_grid, float* reward, int64_t* stats) {
    if (!supported(H, W, K)) return set_error(BGS_EUNSUPPORTED, "connect: unsupported board %dx%d k=%d", H, W, K);
    if (int rc = require_device()) return rc;
    DeviceGuard guard(device);  // the caller's current device is restored on every exit path
    if (guard.rc != BGS_OK) return guard.rc;
    if (n == 0) return BGS_OK;
    const size_t HW = (size_t)H * W;
    cudaStream_t st;
    BGS_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint8_t *d_act = nullptr, *d_len = nullptr;
    int8_t *d_win = nullptr, *d_grid = nullptr;
    float* d_rew = nullptr;
    int64_t* d_stats = nullptr;
    int rc = BGS_OK;
    auto fail = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && rc == BGS_OK) rc = cuda_error(e, what); };
    if (actions) fail(cudaMallocAsync((void**)&d_act, n * HW, st), "alloc actions");
    if (length || actions) fail(cudaMallocAsync((void**)&d_len, n, st), "alloc length");
    if (winner) fail(cudaMallocAsync((void**)&d_win, n, st), "alloc winner");
    if (final_grid) fail(cudaMallocAsync((void**)&d_grid, n * HW, st), "alloc grid");
    if (reward) fail(cudaMallocAsync((void**)&d_rew, n * 2 * sizeof(float), st), "alloc reward");
    if (stats) {
        fail(cudaMallocAsync((void**)&d_stats, BGS_STATS_LEN * sizeof(int64_t), st), "alloc stats");
        if (rc == BGS_OK) fail(cudaMemcpyAsync(d_stats, stats, BGS_STATS_LEN * sizeof(int64_t), cudaMemcpyHostToDevice, st), "h2d stats");
    }
    if (rc == BGS_OK) rc = bgs_connect_rollout_export(H, W, K, n, game_id0, seed, d_act, d_len, d_win, d_grid, d_rew, d_stats, st);
    if (rc == BGS_OK) {
        if (actions) fail(cudaMemcpyAsync(actions, d_act, n * HW, cudaMemcpyDeviceToHost, st), "d2h actions");
        if (length) fail(cudaMemcpyAsync(length, d_len, n, cudaMemcpyDeviceToHost, st), "d2h length");
        if (winner) fail(cudaMemcpyAsync(winner, d_win, n, cudaMemcpyDeviceToHost, st), "d2h winner");
        if (final_grid) fail(cudaMemcpyAsync(final_grid, d_grid, n * HW, cudaMemcpyDeviceToHost, st), "d2h grid");
        if (reward) fail(cudaMemcpyAsync(reward, d_rew, n * 2 * sizeof(float), cudaMemcpyDeviceToHost, st), "d2h reward");
        if (stats) fail(cudaMemcpyAsync(stats, d_stats, BGS_STATS_LEN * sizeof(int64_t), cudaMemcpyDeviceToHost, st), "d2h stats");
    }
    void* bufs[] = {d_act, d_len, d_win, d_grid, d_rew, d_stats};
    for (void* b : bufs)
        if (b) cudaFreeAsync(b, st);
    fail(cudaStreamSynchronize(st), "sync");
    cudaStreamDestroy(st);
    return rc;
}
