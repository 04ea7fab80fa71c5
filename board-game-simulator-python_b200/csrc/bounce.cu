// bounce.cu -- Bounce kernels for sm_100a and their C-ABI entry points.
//
// Replaces, for many games at once, the per-object path of the reference's binding
// src/simulator/game/bounce.cpp:24-53 (State::get_actions / get_actions_at / get_action_at,
// Action::sample_next_state, State::has_ended / get_reward / get_grid).  The rules are restated in
// SURVEY.md 4.4 from tests/test_bounce.py (the engine source is not part of the reference tree).
//
// Data layout: a board is held in registers as NP bit-planes of the piece values (plane b, bit cell =
// bit b of the value; NP = 2 for values <= 3, 4 for values <= 15); a plane is one 64-bit word for
// boards of up to 64 cells and 8 columns, an unsigned __int128 up to 128 cells and 16 columns.  Move
// generation is bit-parallel reachability, not a recursive search:
//   * one "segment" of u steps keeps three frontier masks keyed by the last direction
//     (forward / left / right) and advances all cells at once with a shift + mask per direction;
//   * a segment whose last step lands on pieces seeds new segments ("bounces"), grouped by the value
//     of the piece hit, until no unexpanded bounce cell is left (at most #pieces expansions).
// The per-source target masks of the mover (<= W sources, all in one row) are staged in shared
// memory so that the uniform draw can be mapped to the k-th (source, target) pair in ascending order.
//
// The rules live in ONE place, the lane state machine of bounce_lane.cuh (mover-relative orientation;
// compile-time geometry and a guard column between rows for the rollout of the default 9x6 board,
// the plain y*W + x layout for the batched moves / step kernels).
#include "bgs_common.cuh"
#include "bounce_lane.cuh"

namespace bgs {
namespace bounce {

// ---------------------------------------------------------------------------------------------
// rollout kernel
// ---------------------------------------------------------------------------------------------
struct RolloutParams {
    uint32_t n_games;  // <= 2^31 per launch
    unsigned long long game_id0;
    uint32_t seed_lo, seed_hi;
    int max_plies;
    uint64_t plane0[4];     // bit-planes of the start position (the Config grid), low 64 bits
    uint64_t plane0_hi[4];  // ... bits 64..127 (boards on 128-bit words)
    uint8_t* moves;      // [n, max_plies, 2] pre-filled 0xFF, or null
    uint16_t* length;
    int8_t* winner;
    int8_t* final_grid;  // [n, H*W]
    float* reward;       // [n, 2]
    unsigned long long* stats;
    unsigned int* counter;
    // rollouts from caller-supplied positions (all null: every game starts from plane0, player 0)
    const int8_t* start_grid;     // [n, H*W]
    const int8_t* start_player;   // [n]
    const int8_t* start_winner;   // [n] or null
    const uint8_t* start_ended;   // [n] or null
    int ply_batch, idle_batch;    // slot kernel: waiting slots / idle lanes that trigger the transition pass
    uint32_t one, k16;            // always 1 and 16 (MoveGen::one, k16)
};

constexpr int ROLLOUT_THREADS = 128;

// The tables of the table-driven segment (MoveGen::lut_segment) in shared memory: landing sets [window hash][value],
// then the power table at byte 4096.
template <class G>
__device__ __forceinline__ void build_seg_tables(const G& g, uint32_t* s_lut) {
    static_assert(SEG_LUT_WORDS * 4 == 4096, "lut_segment addresses the power table at byte 4096");
    for (int i = threadIdx.x; i < SEG_LUT_WORDS; i += blockDim.x)
        s_lut[seg_lut_slot(g, i >> 8, (uint32_t)(i & 255))] = seg_lut_entry(g.s(), i >> 8, (uint32_t)(i & 255));
    for (int i = threadIdx.x; i < SEG_POW_WORDS; i += blockDim.x) s_lut[SEG_LUT_WORDS + i] = seg_pow_entry(i >> 2, i & 3);
}

// MoveGen::one / k16: read back from shared memory (values stored from kernel parameters), so that ptxas keeps them
// in registers -- it re-loads a kernel parameter from the constant bank wherever it is used, and that load then sits
// on the critical path of every segment.
template <class MG>
__device__ __forceinline__ void load_seg_consts(MG& mg, uint32_t* s_two, uint32_t one, uint32_t k16) {  // s_two: 8-byte aligned
    if (threadIdx.x == 0) {
        s_two[0] = one;
        s_two[1] = k16;
    }
    __syncthreads();
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(mg.one), "=r"(mg.k16)
                 : "r"((uint32_t)__cvta_generic_to_shared(s_two)));
}
// Lanes that must be ready before the ply transition runs, and move-generation segments per readiness
// check (straight-line copies; a run-time loop is slower).  Frontier-propagation kernels: batch 8 / 12 /
// 16 / 24 -> 13.0 / 12.45 / 12.6 / 13.3 ms, segments 1 / 2 -> 12.8 / 11.9 ms.  Table-driven kernel
// (cheaper segments): (2,12) 11.0 ms, (3,12) 10.7, (4,12) 10.6, (3,14) 10.56, (3,16) 10.57, (4,14) 10.65.
template <bool LUT> struct Tune { static constexpr int PLY_BATCH = LUT ? 14 : 12, SEGMENTS = LUT ? 3 : 2; };

// ---------------------------------------------------------------------------------------------
// rollout kernel, lane formulation (used for boards on 128-bit words): one game per lane (Game +
// MoveGen of bounce_lane.cuh in registers).  Lanes whose move generation is complete wait until Tune::PLY_BATCH of them can run the
// ply transition together.
// ---------------------------------------------------------------------------------------------
template <int NP, class G, int RULES>
__global__ void __launch_bounds__(ROLLOUT_THREADS)
bounce_rollout_lane_kernel(const GeoRTb<typename G::bits> grt, const RolloutParams p) {
    typedef typename G::bits B;
    constexpr int MAXSRC = sizeof(B) == 8 ? 8 : 16;  // movable pieces = columns of one row
    __shared__ unsigned int s_hist[HIST_BINS];
    __shared__ __align__(16) B s_T[MAXSRC * ROLLOUT_THREADS];
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    B* T = s_T + threadIdx.x;  // T[j * ROLLOUT_THREADS] = targets of the j-th movable piece
    const G g(grt);
    constexpr bool USE_LUT = G::LUT && NP == 2;                  // compile-time: the default board
    constexpr bool MAY_LUT = NP == 2 && sizeof(B) == 8;          // run-time: any small board with a guard column
    const bool lut_on = USE_LUT || (MAY_LUT && g.lut_rt());
    __shared__ __align__(16) uint32_t s_lut[MAY_LUT ? SEG_LUT_WORDS + SEG_POW_WORDS + 4 : 4];  // landing sets of a segment, [passable neighbours][value]; the power table; MoveGen::one / k16
    if (lut_on) {
        build_seg_tables(g, s_lut);
        __syncthreads();
    }
    B plane0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) plane0[k] = make_bits<B>(p.plane0[k], p.plane0_hi[k]);
    const LaneOut out{p.moves, p.length, p.winner, p.final_grid, p.reward};
    Game<NP, G> gm;
    MoveGen<NP, G, RULES> mg;
    mg.done = false;
    mg.lut = lut_on ? s_lut : nullptr;
    mg.lut_saddr = (uint32_t)__cvta_generic_to_shared(s_lut);
    asm volatile("mov.u32 %0, %0;" : "+r"(mg.lut_saddr));  // opaque: see MoveGen::lut_saddr
    load_seg_consts(mg, s_lut + (MAY_LUT ? SEG_LUT_WORDS + SEG_POW_WORDS : 0), p.one, p.k16);
    uint32_t r[4] = {0, 0, 0, 0};
    uint32_t acc_w0 = 0, acc_w1 = 0, acc_dr = 0, acc_tr = 0;
    unsigned long long acc_steps = 0;

    uint32_t idx = atomicAdd(p.counter, 1u);
    bool active = idx < p.n_games;
    auto start_movegen = [&](bool probe, bool no_moves) {
#pragma unroll
        for (int i = 0; i < NP; ++i) mg.b[i] = gm.b[i];
        mg.begin(g, probe, no_moves);
    };
    auto begin_game = [&]() {
        bool no_moves = false;
        if (p.start_grid)
            no_moves = gm.begin_grid(g, p.start_grid + (size_t)idx * (g.h() * g.w()), p.start_player[idx],
                                     p.start_winner ? (int)p.start_winner[idx] : BGS_WINNER_DRAW,
                                     p.start_ended && p.start_ended[idx]);
        else
            gm.begin_planes(g, plane0);
        start_movegen(false, no_moves);
    };
    if (active) begin_game();

    for (;;) {
        const unsigned am = __ballot_sync(0xffffffffu, active);
        if (!am) break;
        const unsigned wm = __ballot_sync(0xffffffffu, active && mg.done);
        if (__popc(wm) >= Tune<USE_LUT>::PLY_BATCH || wm == am) {
            if (active && mg.done) {
                uint8_t* row = p.moves ? p.moves + (size_t)idx * p.max_plies * 2ull : nullptr;
                const Next nx = gm.transition(g, T, ROLLOUT_THREADS, mg.total, mg.probe, mg.found, p.max_plies, row,
                                              [&](int t) -> uint32_t {
                                                  if ((t & 3) == 0) {
                                                      const unsigned long long gid = p.game_id0 + idx;
                                                      philox_hd((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t >> 2,
                                                                DOMAIN_BOUNCE, p.seed_lo, p.seed_hi, r);
                                                  }
                                                  return (t & 3) == 0 ? r[0] : ((t & 3) == 1 ? r[1] : ((t & 3) == 2 ? r[2] : r[3]));
                                              });
                if (nx == NEXT_OVER) {
                    gm.write_result(g, out, idx);
                    acc_w0 += (gm.win == 0);
                    acc_w1 += (gm.win == 1);
                    acc_dr += (gm.win == BGS_WINNER_DRAW);
                    acc_tr += (gm.win == BGS_WINNER_TRUNCATED);
                    acc_steps += (unsigned)gm.t;
                    atomicAdd(&s_hist[hist_bin(gm.t)], 1u);
                    idx = atomicAdd(p.counter, 1u);
                    if (idx < p.n_games) begin_game();
                    else active = false;
                } else {
                    start_movegen(nx == NEXT_PROBE, false);
                }
            }
        }
        if (active && !mg.done) {
            mg.iter(g, T, ROLLOUT_THREADS);
#pragma unroll
            for (int q = 1; q < Tune<USE_LUT>::SEGMENTS; ++q)  // straight-line copies (a run-time loop is slower)
                if (!mg.done) mg.iter(g, T, ROLLOUT_THREADS);
        }
    }
    __syncwarp();

    if (p.stats) {
        const unsigned long long w0 = warp_sum(acc_w0), w1 = warp_sum(acc_w1), dr = warp_sum(acc_dr);
        const unsigned long long tr = warp_sum(acc_tr), st = warp_sum(acc_steps);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&p.stats[BGS_STAT_GAMES], w0 + w1 + dr + tr);
            atomicAdd(&p.stats[BGS_STAT_WIN0], w0);
            atomicAdd(&p.stats[BGS_STAT_WIN1], w1);
            atomicAdd(&p.stats[BGS_STAT_DRAWS], dr);
            atomicAdd(&p.stats[BGS_STAT_TRUNCATED], tr);
            atomicAdd(&p.stats[BGS_STAT_STEPS], st);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&p.stats[BGS_STAT_HIST0 + i], (unsigned long long)s_hist[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// rollout kernel, slot formulation.
//
// A warp owns M = 64 game SLOTS in shared memory.  A game alternates between two kinds of work:
//   * its move generation (MoveGen): ~9 segments, the number varies from 1 to 40 between positions;
//   * its ply transition (Game::transition): ~300 instructions, once per ply.
// Lanes are not tied to games.  A lane without work takes any slot from the warp's READY ring and runs
// that position's move generation; when it is done the slot goes to the WAITING ring and the lane
// takes the next ready slot.  When 32 slots are waiting (or lanes starve), the whole warp runs the
// transition for up to 32 of them at once and they become ready again (a finished game's slot is
// refilled from the global game counter).  With M - 32 >= the batch size no lane ever idles and the
// transition always runs on full warps; the rings are warp-private, their heads and counts are
// warp-uniform registers derived from ballots, so there are no atomics besides the game counter (one
// per 64 games) and the final statistics.
// ---------------------------------------------------------------------------------------------
template <int NP, int M, int MAXSRC>
struct WarpSlots {
    uint64_t b[NP][M];      // planes, oriented for the player whose moves are generated next
    uint64_t src[M];        // that player's movable pieces (MoveGen::sources), 0 for an ended start position
    uint64_t T[MAXSRC][M];  // target masks of the j-th movable piece
    uint32_t r[3][M];       // words 1..3 of the Philox block of plies 4*(t>>2) .. +3 (word 0 is used at once)
    uint32_t idx[M];        // game index in [0, n)
    uint32_t meta[M];       // t:16 | total:9 | player | orient | probe | found | win+2:2
    uint8_t rq[64];         // READY ring (slot ids)
    uint8_t wq[64];         // WAITING ring
};

constexpr uint32_t META_TOTAL_SHIFT = 16, META_PLAYER = 1u << 25, META_ORIENT = 1u << 26, META_PROBE = 1u << 27,
                   META_FOUND = 1u << 28, META_WIN_SHIFT = 29;

template <int NP, class G, int RULES, int M>
// resident CTAs per SM that the shared-memory footprint allows (30 / 34 / 38 KB per CTA): the register budget follows
__global__ void __launch_bounds__(ROLLOUT_THREADS, (NP == 2 ? (G::MAX_SOURCES <= 6 ? 7 : 6) : 5))
bounce_rollout_slots_kernel(const GeoRT grt, const RolloutParams p) {
    static_assert(M > 32 && M <= 64, "slot ids are ring entries of 64");
    constexpr int MAXSRC = G::MAX_SOURCES;   // movable pieces = columns of one row
    // iterations (piece boundary + segment) per trip.  Compile-time table-driven geometry (34-instruction segments):
    // 2 / 3 / 4 / 5 / 6 / 8 -> 9.06 / 8.64 / 8.56 / 8.45 / 8.60 / 8.93 ms per 4 Mi default games; with ONE extra,
    // boundary-free segment after each (for the lanes that still have a pending cell: 63 % of the pieces need more than
    // one segment, and the boundary block costs more than a segment): 2 / 3 / 4 iterations -> 8.33 / 8.25 / 8.40 ms; two
    // or three extra segments lose (8.57 - 9.11 ms).  The other variants keep 3 iterations (round 1, with the costlier
    // segments of that build: 3 / 4 / 5 / 6 / 8 -> 9.83 / 10.0 / 9.9 / 10.1 / 10.6 ms)
#ifndef BGS_BOUNCE_ITERS
#define BGS_BOUNCE_ITERS 3
#endif
#ifndef BGS_BOUNCE_EXTRA
#define BGS_BOUNCE_EXTRA 1
#endif
    constexpr int SEGMENTS = (G::LUT && NP == 2) ? BGS_BOUNCE_ITERS : 3;
    constexpr int EXTRA_SEG = NP == 2 ? BGS_BOUNCE_EXTRA : 0;  // boundary-free table-driven segments after the first one of an iteration (lut_on)
    __shared__ unsigned int s_hist[HIST_BINS];
    // outcome counters of the CTA, indexed by winner + 2 (truncated, draw, player 0, player 1), and its env-steps: one
    // shared-memory atomic each per finished game instead of six accumulator registers per lane
    __shared__ unsigned int s_outcome[4];
    __shared__ unsigned long long s_steps;
    __shared__ WarpSlots<NP, M, MAXSRC> s_slots[ROLLOUT_THREADS / 32];
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x < 4) s_outcome[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_steps = 0;
    const G g(grt);
    constexpr bool USE_LUT = G::LUT && NP == 2;
    const bool lut_on = USE_LUT || (NP == 2 && g.lut_rt());
    __shared__ __align__(16) uint32_t s_lut[NP == 2 ? SEG_LUT_WORDS + SEG_POW_WORDS + 4 : 4];
    if (lut_on) build_seg_tables(g, s_lut);
    __syncthreads();
    WarpSlots<NP, M, MAXSRC>& S = s_slots[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const LaneOut out{p.moves, p.length, p.winner, p.final_grid, p.reward};
    uint64_t plane0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) plane0[k] = p.plane0[k];
    uint32_t pool_next = 0, pool_cnt = 0;
    int rq_head = 0, rq_cnt = 0, wq_head = 0, wq_cnt = 0;  // warp-uniform
    bool more = true;                                        // the game counter may still have games

    auto pack_meta = [](const Game<NP, G>& gm, bool probe) {
        return (uint32_t)gm.t | (gm.player ? META_PLAYER : 0u) | (gm.orient ? META_ORIENT : 0u) |
               (probe ? META_PROBE : 0u) | ((uint32_t)(gm.win + 2) << META_WIN_SHIFT);
    };
    auto park = [&](int s, const Game<NP, G>& gm, bool probe, bool no_moves) {  // -> ready for its move generation
#pragma unroll
        for (int i = 0; i < NP; ++i) S.b[i][s] = gm.b[i];
        S.src[s] = MoveGen<NP, G, RULES>::sources(g, gm.b, no_moves);
        S.meta[s] = pack_meta(gm, probe);
    };
    auto new_game = [&](int s, uint32_t idx) {
        Game<NP, G> gm;
        bool no_moves = false;
        if (p.start_grid)
            no_moves = gm.begin_grid(g, p.start_grid + (size_t)idx * (g.h() * g.w()), p.start_player[idx],
                                     p.start_winner ? (int)p.start_winner[idx] : BGS_WINNER_DRAW,
                                     p.start_ended && p.start_ended[idx]);
        else
            gm.begin_planes(g, plane0);
        S.idx[s] = idx;
        park(s, gm, false, no_moves);
    };

    // ---- initial fill of the slots
    for (int s0 = 0; s0 < M; s0 += 32) {
        const int s = s0 + (int)lane;
        const bool want = s < M;
        const unsigned m = __ballot_sync(0xffffffffu, want);
        const uint32_t idx = claim_index<64>(m, p.counter, pool_next, pool_cnt);
        const bool ok = want && idx < p.n_games;
        if (ok) new_game(s, idx);
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (ok) S.rq[(rq_head + rq_cnt + __popc(okm & lt)) & 63] = (uint8_t)s;
        rq_cnt += __popc(okm);
        if (okm != m) more = false;
    }
    __syncwarp();

    MoveGen<NP, G, RULES> mg;
    mg.lut = lut_on ? s_lut : nullptr;
    mg.lut_saddr = (uint32_t)__cvta_generic_to_shared(s_lut);
    asm volatile("mov.u32 %0, %0;" : "+r"(mg.lut_saddr));  // opaque: see MoveGen::lut_saddr
    load_seg_consts(mg, s_lut + (NP == 2 ? SEG_LUT_WORDS + SEG_POW_WORDS : 0), p.one, p.k16);
    // the extra segment pays where pieces bounce often: boards of 5+ columns (8x7: 4.72 -> 4.14 ms per 2 Mi games,
    // 7x5: 2.60 -> 2.31; 6x3: 3.10 -> 3.33, so not there)
    const bool extra_on = lut_on && g.w() >= 5;
    bool has_work = false;
    int slot = 0;
    uint32_t me = 0;  // meta word of the slot in work
    for (;;) {
        // ---- (1) lanes without work take ready slots
        const unsigned need = __ballot_sync(0xffffffffu, !has_work);
#ifndef BGS_BOUNCE_COOP
#define BGS_BOUNCE_COOP 1
#endif
        // ---- (0) cooperative tail.  Once the game counter is exhausted a warp ends up with a handful of long games,
        // each a chain of plies whose move generation (6 pieces x ~2.5 segments) runs on ONE lane: ~3.5 us per ply,
        // and the launch waits for the longest chain.  With at most 32 / MAXSRC games left (and nothing in flight) the
        // movable pieces of a position are dealt to MAXSRC consecutive lanes instead -- lane j generates the moves of
        // the j-th piece (MoveGen on a one-piece source set, its targets into T[j]) -- all lanes run to completion,
        // the totals are summed over the lane group, and the ordinary transition pass follows.  Run-time geometries
        // only: 2 Mi games on 6x3 / 8x7 / 7x5 boards 3.28 / 3.95 / 2.17 -> 2.89 / 3.83 / 2.14 ms, values up to 7 on 9x6
        // 6.37 -> 6.11 ms; on the default board (compile-time geometry) the drain of a warp's last 64 games is spread
        // over 6..64 live slots most of the time and the extra code costs what the last five games gain (7.50 -> 7.60 ms).
        constexpr int COOP_SLOTS = 32 / MAXSRC;
        constexpr bool COOP = BGS_BOUNCE_COOP && !(G::LUT && NP == 2);
        if (COOP && !more && need == 0xffffffffu && wq_cnt == 0 && rq_cnt > 0 && rq_cnt <= COOP_SLOTS) {
            const int grp = (int)lane / MAXSRC, j = (int)lane - grp * MAXSRC;
            const int nsl = rq_cnt;
            bool running = false;
            int cslot = 0;
            uint32_t cme = 0;
            if (grp < nsl) {
                cslot = S.rq[(rq_head + grp) & 63];
#pragma unroll
                for (int i = 0; i < NP; ++i) mg.b[i] = S.b[i][cslot];
                cme = S.meta[cslot];
                uint64_t sj = S.src[cslot];
                for (int q = 0; q < j; ++q) sj &= sj - 1ull;
                sj &= ~sj + 1ull;  // the j-th movable piece, or nothing
                mg.begin_with(g, sj, (cme & META_PROBE) != 0u);
                mg.nsrc = j;
                running = true;
            }
            rq_head = (rq_head + nsl) & 63;
            rq_cnt = 0;
            while (__any_sync(0xffffffffu, running)) {
                if (running) {
                    mg.iter(g, &S.T[0][cslot], M);
                    running = !mg.done;
                }
            }
            int tot = grp < nsl ? mg.total : 0;
            int sum = tot;
#pragma unroll
            for (int k = 1; k < MAXSRC; ++k) {
                const int o = __shfl_down_sync(0xffffffffu, tot, k);
                if (j + k < MAXSRC) sum += o;
            }
            if (grp < nsl && j == 0) {
                S.meta[cslot] = cme | ((uint32_t)sum << META_TOTAL_SHIFT) | (sum ? META_FOUND : 0u);
                S.wq[(wq_head + wq_cnt + grp) & 63] = (uint8_t)cslot;
            }
            wq_cnt += nsl;
            __syncwarp();
            continue;
        }
        if (need != 0u && rq_cnt > 0) {
            const int rank = __popc(need & lt);
            if (!has_work && rank < rq_cnt) {
                slot = S.rq[(rq_head + rank) & 63];
#pragma unroll
                for (int i = 0; i < NP; ++i) mg.b[i] = S.b[i][slot];
                me = S.meta[slot];
                mg.begin_with(g, S.src[slot], (me & META_PROBE) != 0u);
                has_work = true;
            }
            const int take = min(__popc(need), rq_cnt);
            rq_head = (rq_head + take) & 63;
            rq_cnt -= take;
        }
        const unsigned wk = __ballot_sync(0xffffffffu, has_work);
        // ---- (2) the ply transition of up to 32 waiting slots, whole warp
        if (wq_cnt > 0 && (wq_cnt >= p.ply_batch || wk == 0u || (rq_cnt == 0 && __popc(~wk) >= p.idle_batch))) {
            const int nproc = min(32, wq_cnt);
            bool ready = false, over = false;
            int ws = 0;
            if ((int)lane < nproc) {
                ws = S.wq[(wq_head + (int)lane) & 63];
                Game<NP, G> gm;
#pragma unroll
                for (int i = 0; i < NP; ++i) gm.b[i] = S.b[i][ws];
                const uint32_t mw = S.meta[ws];
                const uint32_t gidx = S.idx[ws];
                gm.t = (int)(mw & 0xffffu);
                gm.player = (mw & META_PLAYER) ? 1 : 0;
                gm.orient = (mw & META_ORIENT) ? 1 : 0;
                gm.win = (int)((mw >> META_WIN_SHIFT) & 3u) - 2;
                uint8_t* row = p.moves ? p.moves + (size_t)gidx * p.max_plies * 2ull : nullptr;
                const Next nx = gm.transition(
                    g, &S.T[0][ws], M, (int)((mw >> META_TOTAL_SHIFT) & 0x1ffu), (mw & META_PROBE) != 0u,
                    (mw & META_FOUND) != 0u, p.max_plies, row, [&](int t) -> uint32_t {
                        if ((t & 3) == 0) {
                            const unsigned long long gid = p.game_id0 + gidx;
                            uint32_t r[4];
                            philox_hd((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t >> 2, DOMAIN_BOUNCE, p.seed_lo,
                                      p.seed_hi, r);
                            S.r[0][ws] = r[1]; S.r[1][ws] = r[2]; S.r[2][ws] = r[3];
                            return r[0];
                        }
                        return S.r[(t & 3) - 1][ws];
                    });
                if (nx == NEXT_OVER) {
                    gm.write_result(g, out, gidx);
                    atomicAdd(&s_outcome[gm.win + 2], 1u);
                    atomicAdd(&s_steps, (unsigned long long)(unsigned)gm.t);
                    atomicAdd(&s_hist[hist_bin(gm.t)], 1u);
                    over = true;
                } else {
                    park(ws, gm, nx == NEXT_PROBE, false);
                    ready = true;
                }
            }
            wq_head = (wq_head + nproc) & 63;
            wq_cnt -= nproc;
            const unsigned ov = __ballot_sync(0xffffffffu, over);
            if (ov != 0u && more) {  // refill the slots of finished games
                const uint32_t nidx = claim_index<64>(ov, p.counter, pool_next, pool_cnt);
                const bool ok = over && nidx < p.n_games;
                if (ok) {
                    new_game(ws, nidx);
                    ready = true;
                }
                if (__ballot_sync(0xffffffffu, ok) != ov) more = false;
            }
            const unsigned pr = __ballot_sync(0xffffffffu, ready);
            if (ready) S.rq[(rq_head + rq_cnt + __popc(pr & lt)) & 63] = (uint8_t)ws;
            rq_cnt += __popc(pr);
            __syncwarp();
            continue;
        }
        if (wk == 0u) break;  // no work in flight, nothing waiting, nothing ready
        // ---- (3) up to SEGMENTS move-generation segments
        bool fin = false;
        if (has_work) {
            mg.iter(g, &S.T[0][slot], M);
#pragma unroll
            for (int e = 0; e < EXTRA_SEG; ++e)
                if (extra_on && !mg.done && mg.pending != 0) mg.lut_segment(g);
#pragma unroll
            for (int q = 1; q < SEGMENTS; ++q)
                if (!mg.done) {
                    mg.iter(g, &S.T[0][slot], M);
#pragma unroll
                    for (int e = 0; e < EXTRA_SEG; ++e)
                        if (extra_on && !mg.done && mg.pending != 0) mg.lut_segment(g);
                }
            if (mg.done) {
                S.meta[slot] = me | ((uint32_t)mg.total << META_TOTAL_SHIFT) | (mg.found ? META_FOUND : 0u);
                has_work = false;
                fin = true;
            }
        }
        const unsigned fm = __ballot_sync(0xffffffffu, fin);
        if (fm != 0u) {
            if (fin) S.wq[(wq_head + wq_cnt + __popc(fm & lt)) & 63] = (uint8_t)slot;
            wq_cnt += __popc(fm);
            __syncwarp();
        }
    }
    __syncwarp();

    if (p.stats) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long tr = s_outcome[0], dr = s_outcome[1], w0 = s_outcome[2], w1 = s_outcome[3];
            atomicAdd(&p.stats[BGS_STAT_GAMES], w0 + w1 + dr + tr);
            atomicAdd(&p.stats[BGS_STAT_WIN0], w0);
            atomicAdd(&p.stats[BGS_STAT_WIN1], w1);
            atomicAdd(&p.stats[BGS_STAT_DRAWS], dr);
            atomicAdd(&p.stats[BGS_STAT_TRUNCATED], tr);
            atomicAdd(&p.stats[BGS_STAT_STEPS], s_steps);
        }
        for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&p.stats[BGS_STAT_HIST0 + i], (unsigned long long)s_hist[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// batched move generation / single step on reference-layout states (int8 grids).  Same MoveGen as
// the rollout kernel (bounce_lane.cuh), here on the plain y*W + x layout (no guard column), so that
// the relative target masks of player 0 ARE the public masks and those of player 1 are one 180-degree
// rotation away.
// ---------------------------------------------------------------------------------------------
constexpr int STEP_THREADS = 64;
template <class B>
using StepGen = MoveGen<4, GeoRTb<B>, -1, true>;

__device__ __forceinline__ float2 reward_of(int winner) {
    return make_float2(winner == 0 ? 1.f : (winner == 1 ? -1.f : 0.f),
                       winner == 1 ? 1.f : (winner == 0 ? -1.f : 0.f));
}

template <class B>
__device__ __forceinline__ B rot180(const GeoRTb<B>& g, B x) { return revb(x) >> g.rot_sh(); }

// Value planes (ABSOLUTE orientation, cell = y*W + x; g has no guard column) of a state whose grid
// bytes are staged in shared memory.  Returns false if a cell holds a value outside 0..15.
template <class B>
__device__ __forceinline__ bool planes_from_stage(const uint8_t* mine, int HW, B (&b)[4]) {
    constexpr int NW = (int)sizeof(B) / 4;  // 32-bit words of a plane
    uint32_t bad = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = 0;
#pragma unroll
    for (int wi = 0; wi < NW; ++wi) {
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
        const int c1 = HW < 32 * (wi + 1) ? HW : 32 * (wi + 1);
        for (int c = 32 * wi; c < c1; ++c) {
            const uint32_t v = mine[c], m = 1u << (c - 32 * wi);
            bad |= v;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] |= ((v >> k) & 1u) ? m : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) b[k] |= (B)acc[k] << (32 * wi);
    }
    return bad <= 15u;  // a negative int8 has bit 7 set
}

// Runs a started move generation to completion.
template <class B>
__device__ __forceinline__ void run_movegen(const GeoRTb<B>& g, StepGen<B>& mg, B* T) {
    while (!mg.done) mg.iter(g, T, STEP_THREADS);
}

// One warp per 32 consecutive states: their grids (32*H*W contiguous bytes) are staged in shared
// memory with coalesced 128-bit loads; lane l works on state g0 + l out of the stage.
template <class B>
__global__ void __launch_bounds__(STEP_THREADS)
bounce_moves_kernel(const GeoRTb<B> g, unsigned long long n, const int8_t* __restrict__ grid,
                    const int8_t* __restrict__ player, const uint8_t* __restrict__ ended,
                    int8_t* source_row, uint64_t* targets, int32_t* count, bool vec) {
    constexpr int MAXSRC = sizeof(B) == 8 ? 8 : 16, MAXCELLS = (int)sizeof(B) * 8;
    __shared__ __align__(16) B s_T[MAXSRC * STEP_THREADS];
    __shared__ __align__(16) uint8_t s_stage[STEP_THREADS / 32][32 * MAXCELLS];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long g0 = ((unsigned long long)blockIdx.x * (STEP_THREADS / 32) + warp) * 32ull;
    if (g0 >= n) return;  // warp-uniform
    const int HW = g.H * g.W;
    const unsigned rows = (unsigned)((n - g0) < 32ull ? (n - g0) : 32ull);
    warp_copy_bytes(s_stage[warp], reinterpret_cast<const uint8_t*>(grid) + g0 * (unsigned)HW, rows * (unsigned)HW, lane, vec);
    __syncwarp();
    const unsigned long long i = g0 + lane;
    if (i >= n) return;
    B* T = s_T + threadIdx.x;
    StepGen<B> mg;
    mg.lut = nullptr;
    mg.lut_saddr = 0;
    const bool ok = planes_from_stage(s_stage[warp] + lane * HW, HW, mg.b);
    const bool over = (ended && ended[i]) || !ok;
    const int pl = player[i] & 1;
    if (pl) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mg.b[k] = rot180(g, mg.b[k]);
    }
    B src = StepGen<B>::sources(g, mg.b, over);  // movable pieces, mover-relative
    mg.begin_with(g, src, false);
    run_movegen(g, mg, T);
    constexpr int TW = (int)sizeof(B) / 8;  // 64-bit words of one target mask
    for (int x = 0; x < g.W * TW; ++x) targets[i * g.W * TW + x] = 0ull;
    int row_rel = -1;
    if (src) row_rel = g.row_of(ctzb(src));
    for (int j = 0; src; ++j) {  // the j-th movable piece in ascending relative column
        const int cell = ctzb(src);
        src &= src - (B)1;
        const int x_rel = cell - row_rel * g.S;
        B t = T[j * STEP_THREADS];
        if (pl) t = rot180(g, t);
        uint64_t* dst = targets + (i * g.W + (pl ? g.W - 1 - x_rel : x_rel)) * TW;
#pragma unroll
        for (int w = 0; w < TW; ++w) dst[w] = (uint64_t)(t >> (w * 32) >> (w * 32));
    }
    if (source_row) source_row[i] = (int8_t)(mg.total > 0 ? (pl ? g.H - 1 - row_rel : row_rel) : -1);
    if (count) count[i] = ok ? mg.total : -1;
}

// Weighted action choice (bgs_bounce_sample_step): `random.choices(actions, weights)` of the reference's
// agent loop (textual/examples/arena.py:64-68) inside the transition kernel.  probs[i, sx, ty*W + tx] is the
// weight of moving the piece in column sx of the mover's source row to cell (tx, ty); quantisation, draw and
// choice exactly as for Connect (connect.cu, StepPolicy) over the legal actions in canonical order
// (ascending sx, then target cell), RNG domain 1, t = draw_index[i] (0 if NULL).  Equal weights reproduce
// the uniform choice of the rollout kernels: action mulhi32(r, n_actions).
struct SamplePolicy {
    const float* probs;           // [n, W, H*W] or null (moves come from `move`)
    const uint64_t* game_ids;     // [n] or null (global id = game_id0 + i)
    const int32_t* draw_index;    // [n] or null
    unsigned long long game_id0;
    uint32_t seed_lo, seed_hi;
    int32_t* move_out;            // [n, 4] the chosen (sx, sy, tx, ty), -1s when there was nothing to choose; or null
};

__device__ __forceinline__ uint32_t quantize_weight(float w, float wmax) {
    return (uint32_t)__fadd_rn(__fmul_rn(__fdiv_rn(w, wmax), 65535.0f), 0.5f);
}
__device__ __forceinline__ float sane_weight(float w) { return w > 0.0f ? fminf(w, 3.402823466e+38f) : 0.0f; }

template <class B>
__global__ void __launch_bounds__(STEP_THREADS)
bounce_step_kernel(const GeoRTb<B> g, unsigned long long n, const int8_t* __restrict__ grid,
                   const int8_t* __restrict__ player, const int8_t* __restrict__ winner,
                   const uint8_t* __restrict__ ended,
                   const int32_t* __restrict__ move, int8_t* grid_out, int8_t* player_out,
                   int8_t* winner_out, uint8_t* ended_out, float* reward_out, int32_t* status, bool vec,
                   const SamplePolicy pol) {
    constexpr int MAXSRC = sizeof(B) == 8 ? 8 : 16, MAXCELLS = (int)sizeof(B) * 8;
    __shared__ __align__(16) B s_T[MAXSRC * STEP_THREADS];
    __shared__ __align__(16) uint8_t s_stage[STEP_THREADS / 32][32 * MAXCELLS];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long g0 = ((unsigned long long)blockIdx.x * (STEP_THREADS / 32) + warp) * 32ull;
    if (g0 >= n) return;  // warp-uniform
    const int HW = g.H * g.W;
    const unsigned rows = (unsigned)((n - g0) < 32ull ? (n - g0) : 32ull);
    const unsigned span = rows * (unsigned)HW;
    warp_copy_bytes(s_stage[warp], reinterpret_cast<const uint8_t*>(grid) + g0 * (unsigned)HW, span, lane, vec);
    __syncwarp();
    const unsigned long long i = g0 + lane;
    if (i < n) {
        B* T = s_T + threadIdx.x;
        uint8_t* mine = s_stage[warp] + lane * HW;  // the new grid = the old one with two cells changed
        StepGen<B> mg;
    mg.lut = nullptr;
    mg.lut_saddr = 0;
        const bool ok = planes_from_stage(mine, HW, mg.b);
        int pl = player[i] & 1;
        const bool over = ended && ended[i];
        int sx = -1, sy = -1, tx = -1, ty = -1;
        if (pol.probs) {
            if (ok && !over) {  // every legal action of the mover, then the weighted draw
                B keep[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    keep[k] = mg.b[k];
                    if (pl) mg.b[k] = rot180(g, mg.b[k]);
                }
                B src = StepGen<B>::sources(g, mg.b, false);
                const int row_rel = src ? g.row_of(ctzb(src)) : 0;
                mg.begin_with(g, src, false);
                run_movegen(g, mg, T);
                if (mg.total > 0) {
                    // absolute target masks by absolute source column (rotate back for player 1), in local registers
                    B A[MAXSRC];
#pragma unroll
                    for (int x = 0; x < MAXSRC; ++x) A[x] = 0;
                    for (int j = 0; src; ++j) {
                        const int x_rel = ctzb(src) - row_rel * g.S;
                        src &= src - (B)1;
                        const B tm = T[j * STEP_THREADS];
                        const int xa = pl ? g.W - 1 - x_rel : x_rel;
#pragma unroll
                        for (int x = 0; x < MAXSRC; ++x)
                            if (x == xa) A[x] = pl ? rot180(g, tm) : tm;
                    }
                    const float* pw = pol.probs + i * (unsigned long long)(g.W * HW);
                    float wmax = 0.0f;
#pragma unroll
                    for (int x = 0; x < MAXSRC; ++x)
                        for (B m = A[x]; m; m &= m - (B)1) wmax = fmaxf(wmax, sane_weight(pw[x * HW + ctzb(m)]));
                    uint64_t total = 0;
#pragma unroll
                    for (int x = 0; x < MAXSRC; ++x)
                        for (B m = A[x]; m; m &= m - (B)1)
                            total += wmax > 0.0f ? quantize_weight(sane_weight(pw[x * HW + ctzb(m)]), wmax) : 1u;
                    const uint32_t t = pol.draw_index ? (uint32_t)pol.draw_index[i] : 0u;
                    const unsigned long long gid = pol.game_ids ? pol.game_ids[i] : pol.game_id0 + i;
                    uint32_t r4[4];
                    philox_hd((uint32_t)gid, (uint32_t)(gid >> 32), t >> 2, DOMAIN_BOUNCE, pol.seed_lo, pol.seed_hi, r4);
                    const uint32_t r = (t & 3u) == 0 ? r4[0] : ((t & 3u) == 1 ? r4[1] : ((t & 3u) == 2 ? r4[2] : r4[3]));
                    const uint64_t thresh = (uint64_t)r * total;
                    uint64_t cum = 0;
                    bool found = false;
#pragma unroll
                    for (int x = 0; x < MAXSRC; ++x)
                        for (B m = A[x]; m && !found; m &= m - (B)1) {
                            const int cell = ctzb(m);
                            cum += wmax > 0.0f ? quantize_weight(sane_weight(pw[x * HW + cell]), wmax) : 1u;
                            if ((cum << 32) > thresh) {
                                found = true;
                                sx = x; sy = pl ? g.H - 1 - row_rel : row_rel;
                                tx = cell % g.W; ty = cell / g.W;
                            }
                        }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) mg.b[k] = keep[k];
            }
            if (pol.move_out) {
                pol.move_out[4 * i + 0] = sx; pol.move_out[4 * i + 1] = sy;
                pol.move_out[4 * i + 2] = tx; pol.move_out[4 * i + 3] = ty;
            }
        } else {
            sx = move[4 * i + 0]; sy = move[4 * i + 1]; tx = move[4 * i + 2]; ty = move[4 * i + 3];
        }
        bool legal = ok && !over && sx >= 0 && sx < g.W && sy >= 0 && sy < g.H && tx >= 0 && tx < g.W && ty >= 0 && ty < g.H;
        int win = winner ? (int)winner[i] : -1;
        bool end_new = over;
        if (legal) {
            if (pl) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mg.b[k] = rot180(g, mg.b[k]);
            }
            const int scell_abs = sy * g.W + sx, tcell_abs = ty * g.W + tx;
            const int scell = pl ? HW - 1 - scell_abs : scell_abs, tcell = pl ? HW - 1 - tcell_abs : tcell_abs;
            const B smask = (B)1 << scell, tmask = (B)1 << tcell;
            legal = (StepGen<B>::sources(g, mg.b, false) & smask) != 0;  // a movable piece of the mover
            if (legal) {
                mg.begin_with(g, smask, false);  // the targets of that piece only
                run_movegen(g, mg, T);
                legal = mg.total > 0 && (T[0] & tmask) != 0;
            }
            if (legal) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool has = (mg.b[k] & smask) != 0;
                    mg.b[k] = (mg.b[k] & ~smask) | (has ? tmask : (B)0);
                }
                mine[tcell_abs] = mine[scell_abs];
                mine[scell_abs] = 0;
                if (tmask & g.m_far()) {  // reached the far goal row
                    win = pl;
                    end_new = true;
                } else {
                    B own[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        own[k] = mg.b[k];
                        mg.b[k] = rot180(g, mg.b[k]);  // the opponent's orientation
                    }
                    mg.begin(g, true, false);  // does the opponent have any action?
                    run_movegen(g, mg, T);
                    if (!mg.found) {  // blocked: the mover wins unless blocked too (draw)
                        end_new = true;
#pragma unroll
                        for (int k = 0; k < 4; ++k) mg.b[k] = own[k];
                        mg.begin(g, true, false);
                        run_movegen(g, mg, T);
                        if (mg.found) win = pl;
                    }
                }
                pl = 1 - pl;
            }
        }
        player_out[i] = (int8_t)pl;
        winner_out[i] = (int8_t)win;
        if (ended_out) ended_out[i] = end_new;
        if (reward_out) reinterpret_cast<float2*>(reward_out)[i] = reward_of(win);
        if (status) status[i] = legal ? 0 : 1;
    }
    __syncwarp();
    warp_copy_bytes(reinterpret_cast<uint8_t*>(grid_out) + g0 * (unsigned)HW, s_stage[warp], span, lane, vec);
}

// length (< 16384) and winner of every game in one 16-bit word: bits 0..13 length, bits 14..15 winner + 2
// (0 truncated, 1 draw, 2 player 0, 3 player 1): 2 instead of 3 bytes per game over PCIe.
__global__ void __launch_bounds__(256)
bounce_pack_results_kernel(unsigned long long n, const uint16_t* __restrict__ length, const int8_t* __restrict__ winner,
                           uint16_t* __restrict__ packed) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        packed[i] = (uint16_t)((length[i] & 0x3FFFu) | (((unsigned)(winner[i] + 2) & 3u) << 14));
}

// boards that need 128-bit words
static bool wide_board(int H, int W) { return W > 8 || H * W > 64; }

static bool supported(int H, int W, int max_value) {
    return H >= 1 && W >= 1 && W <= 16 && H * W <= 128 && max_value <= 15;
}

}  // namespace bounce
}  // namespace bgs

using namespace bgs;
using namespace bgs::bounce;

extern "C" int bgs_bounce_supported(int H, int W, int max_value) { return supported(H, W, max_value) ? 1 : 0; }

extern "C" int bgs_bounce_moves(int H, int W, int rules, uint64_t n, const int8_t* grid, const int8_t* player,
                                const uint8_t* ended, int8_t* source_row, uint64_t* targets, int32_t* count,
                                void* stream_) {
    if (!supported(H, W, 0)) return set_error(BGS_EUNSUPPORTED, "bounce: unsupported board %dx%d", H, W);
    if (!grid || !player || !targets) return set_error(BGS_EINVAL, "bounce_moves: null required pointer");
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    const unsigned long long blocks = (n + STEP_THREADS - 1) / STEP_THREADS;
    const bool vec = ((uintptr_t)grid & 15u) == 0;
    if (wide_board(H, W))
        bounce_moves_kernel<u128><<<(unsigned)blocks, STEP_THREADS, 0, (cudaStream_t)stream_>>>(
            make_geo_rt_b<u128>(H, W, rules, /*guard=*/false), n, grid, player, ended, source_row, targets, count, vec);
    else
        bounce_moves_kernel<uint64_t><<<(unsigned)blocks, STEP_THREADS, 0, (cudaStream_t)stream_>>>(
            make_geo_rt(H, W, rules, /*guard=*/false), n, grid, player, ended, source_row, targets, count, vec);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

static int bounce_step_impl(int H, int W, int rules, uint64_t n, const int8_t* grid, const int8_t* player,
                            const int8_t* winner, const uint8_t* ended, const int32_t* move, int8_t* grid_out, int8_t* player_out,
                            int8_t* winner_out, uint8_t* ended_out, float* reward_out, int32_t* status,
                            void* stream_, const SamplePolicy& pol) {
    if (int rc = check_reward_alignment(reward_out, "bounce_step")) return rc;
    if (!supported(H, W, 0)) return set_error(BGS_EUNSUPPORTED, "bounce: unsupported board %dx%d", H, W);
    if (!grid || !player || !grid_out || !player_out || !winner_out)
        return set_error(BGS_EINVAL, "bounce_step: null required pointer");
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    const unsigned long long blocks = (n + STEP_THREADS - 1) / STEP_THREADS;
    const bool vec = (((uintptr_t)grid | (uintptr_t)grid_out) & 15u) == 0;
    if (wide_board(H, W))
        bounce_step_kernel<u128><<<(unsigned)blocks, STEP_THREADS, 0, (cudaStream_t)stream_>>>(
            make_geo_rt_b<u128>(H, W, rules, /*guard=*/false), n, grid, player, winner, ended, move, grid_out, player_out,
            winner_out, ended_out, reward_out, status, vec, pol);
    else
        bounce_step_kernel<uint64_t><<<(unsigned)blocks, STEP_THREADS, 0, (cudaStream_t)stream_>>>(
            make_geo_rt(H, W, rules, /*guard=*/false), n, grid, player, winner, ended, move, grid_out, player_out,
            winner_out, ended_out, reward_out, status, vec, pol);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

extern "C" int bgs_bounce_step(int H, int W, int rules, uint64_t n, const int8_t* grid, const int8_t* player,
                               const int8_t* winner, const uint8_t* ended, const int32_t* move, int8_t* grid_out, int8_t* player_out,
                               int8_t* winner_out, uint8_t* ended_out, float* reward_out, int32_t* status,
                               void* stream_) {
    if (!move) return set_error(BGS_EINVAL, "bounce_step: null required pointer");
    SamplePolicy pol{};
    return bounce_step_impl(H, W, rules, n, grid, player, winner, ended, move, grid_out, player_out, winner_out, ended_out,
                            reward_out, status, stream_, pol);
}

extern "C" int bgs_bounce_sample_step(int H, int W, int rules, uint64_t n, const int8_t* grid, const int8_t* player,
                                      const int8_t* winner, const uint8_t* ended, const float* probs, uint64_t seed,
                                      uint64_t game_id0, const uint64_t* game_ids, const int32_t* draw_index,
                                      int8_t* grid_out, int8_t* player_out, int8_t* winner_out, uint8_t* ended_out,
                                      float* reward_out, int32_t* move_out, int32_t* status, void* stream_) {
    if (!probs) return set_error(BGS_EINVAL, "bounce_sample_step: null `probs`");
    SamplePolicy pol{};
    pol.probs = probs; pol.game_ids = game_ids; pol.draw_index = draw_index; pol.game_id0 = game_id0;
    pol.seed_lo = (uint32_t)seed; pol.seed_hi = (uint32_t)(seed >> 32); pol.move_out = move_out;
    return bounce_step_impl(H, W, rules, n, grid, player, winner, ended, nullptr, grid_out, player_out, winner_out,
                            ended_out, reward_out, status, stream_, pol);
}

template <class K>
static int persistent_blocks(K kern, uint32_t n_games, int* blocks_out) {
    int per_sm = 0;
    BGS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ROLLOUT_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    const unsigned long long want = ((unsigned long long)n_games + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS;
    unsigned long long blocks = (unsigned long long)sm_count() * per_sm;
    if (want < blocks) blocks = want ? want : 1;
    *blocks_out = (int)blocks;
    return BGS_OK;
}

template <int NP, class G, int RULES>
static int launch_bounce_lane(const GeoRTb<typename G::bits>& grt, const RolloutParams& p, cudaStream_t stream) {
    auto kern = bounce_rollout_lane_kernel<NP, G, RULES>;
    int blocks = 0;
    if (int rc = persistent_blocks(kern, p.n_games, &blocks)) return rc;
    kern<<<(unsigned)blocks, ROLLOUT_THREADS, 0, stream>>>(grt, p);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

// 64-bit boards run on the slot kernel (2 Mi games, slot / lane kernel: default board 5.51 / 5.89 ms, 8x7 4.47 /
// 5.08 ms, values up to 7 3.69 / 4.07 ms, 6x3 1.84 / 1.87 ms); boards on 128-bit words keep the lane kernel,
// whose per-warp state fits the 48 KB of static shared memory.
template <int NP, class G, int RULES>
static int launch_bounce_slots(const GeoRT& grt, RolloutParams p, cudaStream_t stream) {
#ifndef BGS_BOUNCE_PLY_BATCH
#define BGS_BOUNCE_PLY_BATCH 32
#endif
#ifndef BGS_BOUNCE_IDLE_BATCH
#define BGS_BOUNCE_IDLE_BATCH 8
#endif
    p.ply_batch = BGS_BOUNCE_PLY_BATCH;    // 20 / 24 / 28 / 32 waiting slots: 10.35 / 10.08 / 9.83 / 9.83 ms per 4 Mi default games
    p.idle_batch = BGS_BOUNCE_IDLE_BATCH;  // 4 / 8 / 16 idle lanes: 9.90 / 9.83 / 9.99 ms
    auto kern = bounce_rollout_slots_kernel<NP, G, RULES, 64>;
    int blocks = 0;
    if (int rc = persistent_blocks(kern, p.n_games, &blocks)) return rc;
    kern<<<(unsigned)blocks, ROLLOUT_THREADS, 0, stream>>>(grt, p);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

static int bounce_rollout_impl(const int8_t* grid0, const int8_t* start_grid, const int8_t* start_player,
                               const int8_t* start_winner, const uint8_t* start_ended, int H, int W, int rules,
                               int max_plies, uint64_t n_games, uint64_t game_id0, uint64_t seed, uint8_t* moves,
                               uint16_t* length, int8_t* winner, int8_t* final_grid, float* reward, int64_t* stats,
                               void* stream_) {
    if (max_plies < 0 || max_plies > 65535) return set_error(BGS_EINVAL, "bounce_rollout: max_plies out of range");
    int maxv = 0;
    if (grid0 && H >= 1 && W >= 1 && H * W <= 128)
        for (int c = 0; c < H * W; ++c) {
            if (grid0[c] < 0) return set_error(BGS_EINVAL, "bounce_rollout: negative cell value");
            if (grid0[c] > maxv) maxv = grid0[c];
        }
    if (!supported(H, W, maxv))
        return set_error(BGS_EUNSUPPORTED, "bounce: unsupported board %dx%d (max value %d)", H, W, maxv);
    if (n_games > (1ull << 31)) return set_error(BGS_EINVAL, "bounce_rollout: more than 2^31 games per call");
    if (int rc = require_device()) return rc;
    if (n_games == 0) return BGS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    RolloutParams p;
    p.one = 1u; p.k16 = 16u;
    p.n_games = (uint32_t)n_games; p.game_id0 = game_id0;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
    p.max_plies = max_plies;
    const bool wide = wide_board(H, W);
    const GeoRT grt = make_geo_rt(wide ? 1 : H, wide ? 1 : W, rules);
    const GeoRT128 grt128 = make_geo_rt_b<u128>(H, W, rules);
    for (int i = 0; i < 4; ++i) p.plane0[i] = p.plane0_hi[i] = 0;
    if (grid0) {
        if (wide) {
            u128 pl[4];
            planes_from_grid(grt128, grid0, pl);
            for (int i = 0; i < 4; ++i) {
                p.plane0[i] = (uint64_t)pl[i];
                p.plane0_hi[i] = (uint64_t)(pl[i] >> 64);
            }
        } else {
            planes_from_grid(grt, grid0, p.plane0);
        }
    }
    p.moves = moves; p.length = length; p.winner = winner; p.final_grid = final_grid; p.reward = reward;
    p.stats = reinterpret_cast<unsigned long long*>(stats);
    p.start_grid = start_grid; p.start_player = start_player; p.start_winner = start_winner; p.start_ended = start_ended;
    unsigned int* counter = nullptr;
    if (int rc = next_counter(&counter)) return rc;
    p.counter = counter;
    BGS_CUDA_TRY(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), stream));
    if (moves) BGS_CUDA_TRY(cudaMemsetAsync(moves, 0xFF, n_games * (size_t)max_plies * 2, stream));
    // per-game start grids may hold any value up to 15: use the 4-plane kernel
    if (wide) return launch_bounce_lane<4, GeoRT128, -1>(grt128, p, stream);
    if (maxv <= 3 && !start_grid) {
        // the table-driven segments assume that the goal rows hold no piece (a window never starts below cell 3)
        bool goal_rows_empty = true;
        for (int x = 0; x < W; ++x) goal_rows_empty = goal_rows_empty && grid0[x] == 0 && grid0[(H - 1) * W + x] == 0;
        if (H == 9 && W == 6 && rules == 0 && goal_rows_empty) return launch_bounce_slots<2, GeoCT<9, 6>, 0>(grt, p, stream);
        GeoRT g2 = grt;
        g2.lut_ok = g2.lut_ok && goal_rows_empty;
        return launch_bounce_slots<2, GeoRT, -1>(g2, p, stream);
    }
    return launch_bounce_slots<4, GeoRT, -1>(grt, p, stream);
}

extern "C" int bgs_bounce_rollout(const int8_t* grid0, int H, int W, int rules, int max_plies, uint64_t n_games,
                                  uint64_t game_id0, uint64_t seed, uint8_t* moves, uint16_t* length,
                                  int8_t* winner, int8_t* final_grid, float* reward, int64_t* stats,
                                  void* stream_) {
    if (!grid0) return set_error(BGS_EINVAL, "bounce_rollout: null grid0");
    return bounce_rollout_impl(grid0, nullptr, nullptr, nullptr, nullptr, H, W, rules, max_plies, n_games, game_id0,
                               seed, moves, length, winner, final_grid, reward, stats, stream_);
}

extern "C" int bgs_bounce_rollout_from(int H, int W, int rules, int max_plies, uint64_t n_games, uint64_t game_id0,
                                       uint64_t seed, const int8_t* grid, const int8_t* player,
                                       const int8_t* winner_in, const uint8_t* ended_in, uint8_t* moves,
                                       uint16_t* length, int8_t* winner, int8_t* final_grid, float* reward,
                                       int64_t* stats, void* stream_) {
    if (!grid || !player) return set_error(BGS_EINVAL, "bounce_rollout_from: null required pointer");
    return bounce_rollout_impl(nullptr, grid, player, winner_in, ended_in, H, W, rules, max_plies, n_games, game_id0,
                               seed, moves, length, winner, final_grid, reward, stats, stream_);
}

extern "C" int bgs_bounce_pack_results(uint64_t n, const uint16_t* length, const int8_t* winner, uint16_t* packed,
                                       void* stream_) {
    if (!length || !winner || !packed) return set_error(BGS_EINVAL, "bounce_pack_results: null pointer");
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    unsigned long long blocks = (n + 255ull) / 256ull;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    bounce_pack_results_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(n, length, winner, packed);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

extern "C" int bgs_bounce_rollout_host(int device, const int8_t* grid0, int H, int W, int rules, int max_plies,
                                       uint64_t n, uint64_t game_id0, uint64_t seed, uint8_t* moves,
                                       uint16_t* length, int8_t* winner, int8_t* final_grid, float* reward,
                                       int64_t* stats) {
    if (int rc = require_device()) return rc;
    DeviceGuard guard(device);  // the caller's current device is restored on every exit path
    if (guard.rc != BGS_OK) return guard.rc;
    if (n == 0) return BGS_OK;
    if (!supported(H, W, 0)) return set_error(BGS_EUNSUPPORTED, "bounce: unsupported board %dx%d", H, W);
    const size_t HW = (size_t)H * W;
    cudaStream_t st;
    BGS_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint8_t* d_moves = nullptr;
    uint16_t* d_len = nullptr;
    int8_t *d_win = nullptr, *d_grid = nullptr;
    float* d_rew = nullptr;
    int64_t* d_stats = nullptr;
    int rc = BGS_OK;
    auto fail = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && rc == BGS_OK) rc = cuda_error(e, what); };
    if (moves) fail(cudaMallocAsync((void**)&d_moves, n * (size_t)max_plies * 2, st), "alloc moves");
    if (length) fail(cudaMallocAsync((void**)&d_len, n * sizeof(uint16_t), st), "alloc length");
    if (winner) fail(cudaMallocAsync((void**)&d_win, n, st), "alloc winner");
    if (final_grid) fail(cudaMallocAsync((void**)&d_grid, n * HW, st), "alloc grid");
    if (reward) fail(cudaMallocAsync((void**)&d_rew, n * 2 * sizeof(float), st), "alloc reward");
    if (stats) {
        fail(cudaMallocAsync((void**)&d_stats, BGS_STATS_LEN * sizeof(int64_t), st), "alloc stats");
        if (rc == BGS_OK) fail(cudaMemcpyAsync(d_stats, stats, BGS_STATS_LEN * sizeof(int64_t), cudaMemcpyHostToDevice, st), "h2d stats");
    }
    if (rc == BGS_OK)
        rc = bgs_bounce_rollout(grid0, H, W, rules, max_plies, n, game_id0, seed, d_moves, d_len, d_win, d_grid, d_rew, d_stats, st);
    if (rc == BGS_OK) {
        if (moves) fail(cudaMemcpyAsync(moves, d_moves, n * (size_t)max_plies * 2, cudaMemcpyDeviceToHost, st), "d2h moves");
        if (length) fail(cudaMemcpyAsync(length, d_len, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, st), "d2h length");
        if (winner) fail(cudaMemcpyAsync(winner, d_win, n, cudaMemcpyDeviceToHost, st), "d2h winner");
        if (final_grid) fail(cudaMemcpyAsync(final_grid, d_grid, n * HW, cudaMemcpyDeviceToHost, st), "d2h grid");
        if (reward) fail(cudaMemcpyAsync(reward, d_rew, n * 2 * sizeof(float), cudaMemcpyDeviceToHost, st), "d2h reward");
        if (stats) fail(cudaMemcpyAsync(stats, d_stats, BGS_STATS_LEN * sizeof(int64_t), cudaMemcpyDeviceToHost, st), "d2h stats");
    }
    void* bufs[] = {d_moves, d_len, d_win, d_grid, d_rew, d_stats};
    for (void* b : bufs)
        if (b) cudaFreeAsync(b, st);
    fail(cudaStreamSynchronize(st), "sync");
    cudaStreamDestroy(st);
    return rc;
}
