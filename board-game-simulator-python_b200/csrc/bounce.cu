// bounce.cu -- Bounce kernels for sm_100a and their C-ABI entry points.
//
// Replaces, for many games at once, the per-object path of the reference's binding
// src/simulator/game/bounce.cpp:24-53 (State::get_actions / get_actions_at / get_action_at,
// Action::sample_next_state, State::has_ended / get_reward / get_grid).  The rules are restated in
// SURVEY.md 4.4 from tests/test_bounce.py (the engine source is not part of the reference tree).
//
// Data layout: a board of H*W <= 64 cells (cell = y*W + x, row 0 = bottom) is held in registers as
// NP bit-planes of the piece values (plane b, bit cell = bit b of the value; NP = 2 for values <= 3,
// 4 for values <= 15).  Move generation is bit-parallel reachability, not a recursive search:
//   * one "segment" of u steps keeps three frontier masks keyed by the last direction
//     (forward / left / right) and advances all cells at once with a shift + mask per direction;
//   * a segment whose last step lands on pieces seeds new segments ("bounces"), grouped by the value
//     of the piece hit, until no unexpanded bounce cell is left (at most #pieces expansions).
// The per-source target masks of the mover (<= W sources, all in one row) are staged in shared
// memory so that the uniform draw can be mapped to the k-th (source, target) pair in ascending order.
//
// The rollout kernel uses the lane state machine of bounce_lane.cuh (mover-relative orientation, a
// guard column between rows, compile-time geometry for the default 9x6 board); the batched
// moves / step kernels below keep the plain y*W + x layout of this file.
#include "bgs_common.cuh"
#include "bounce_lane.cuh"

namespace bgs {
namespace bounce {

struct Geo {
    int H, W, rules;
    uint64_t board;      // low H*W bits
    uint64_t not_left;   // cells with x > 0
    uint64_t not_right;  // cells with x < W-1
    uint64_t far0, far1;  // far goal row of player 0 (row H-1) / player 1 (row 0)
    uint64_t row0;        // (1 << W) - 1
    __host__ __device__ uint64_t far(int player) const { return player == 0 ? far0 : far1; }
};

static Geo make_geo(int H, int W, int rules) {
    Geo g;
    g.H = H; g.W = W; g.rules = rules;
    g.board = (H * W == 64) ? ~0ull : ((1ull << (H * W)) - 1ull);
    g.not_left = 0; g.not_right = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (x > 0) g.not_left |= 1ull << (y * W + x);
            if (x < W - 1) g.not_right |= 1ull << (y * W + x);
        }
    g.row0 = (1ull << W) - 1ull;
    g.far0 = g.row0 << ((H - 1) * W);
    g.far1 = g.row0;
    return g;
}

template <int NP>
struct Planes {
    uint64_t b[NP];
    __device__ __forceinline__ uint64_t occ() const {
        uint64_t o = b[0];
#pragma unroll
        for (int i = 1; i < NP; ++i) o |= b[i];
        return o;
    }
    __device__ __forceinline__ int value_at(int cell) const {
        int v = 0;
#pragma unroll
        for (int i = 0; i < NP; ++i) v |= (int)((b[i] >> cell) & 1ull) << i;
        return v;
    }
    __device__ __forceinline__ uint64_t cells_with_value(int u) const {
        uint64_t m = ~0ull;
#pragma unroll
        for (int i = 0; i < NP; ++i) m &= ((u >> i) & 1) ? b[i] : ~b[i];
        return m;
    }
};

// Mask of the row holding the movable pieces of `player`: the occupied row nearest to its own side
// (tests/test_bounce.py:43-48,60).  0 if the board is empty.  *row receives the row index.
__device__ __forceinline__ uint64_t source_row_mask(const Geo& g, uint64_t occ, int player, int* row) {
    if (!occ) {
        *row = -1;
        return 0;
    }
    const int cell = player == 0 ? (__ffsll((long long)occ) - 1) : (63 - __clzll((long long)occ));
    const int y = cell / g.W;
    *row = y;
    return g.row0 << (y * g.W);
}

// Move generation for `player` -- SURVEY.md 4.4 rules 2-4 -- as ONE flat loop per lane.
//
// A lane walks its own work list: for every movable piece (ascending column) a sequence of segments,
// for every segment `u` frontier steps.  Each loop iteration performs exactly one frontier step of
// whatever (piece, segment) the lane is at; the bookkeeping between segments / pieces is a short
// branch at the top.  Lanes of a warp therefore stay busy until the lane with the most steps is
// done, instead of serialising over columns and bounce rounds (the nested formulation ran at 5 of 32
// active lanes).
//
// ANY = false: writes the target mask of the piece in column x to T[x*stride] (0 where there is no
//              movable piece) and returns the number of (source, target) pairs; *row = source row.
// ANY = true : returns 1 as soon as one piece has a target, else 0 (T is not touched).
template <int NP, bool ANY>
__device__ __forceinline__ int movegen(const Geo& g, const Planes<NP>& P, int player, uint64_t* T, int stride,
                                       int* row) {
    const uint64_t occ = P.occ();
    const uint64_t rowm = source_row_mask(g, occ, player, row);
    if (!ANY)
        for (int x = 0; x < g.W; ++x) T[x * stride] = 0ull;
    const int base = *row * g.W;
    const int variant = g.rules & 3;
    const bool allow_null = (g.rules & BGS_BOUNCE_ALLOW_NULL_MOVE) != 0;
    const uint64_t farm = g.far(player);
    uint64_t src_left = rowm & occ;
    uint64_t sbit = 0, occS = 0, open = 0, inter = 0, expanded = 0, pending = 0, targets = 0;
    uint64_t Ff = 0, Fl = 0, Fr = 0, Nn = 0;  // frontier by last direction: forward / left / right / none
    int rem = 0, xs = 0, total = 0;
    bool have = false;
    for (;;) {
        if (rem == 0) {  // between segments
            int u;
            uint64_t S;
            if (pending == 0) {  // between pieces
                if (have) {
                    if (!allow_null) targets &= ~sbit;
                    if (ANY) {
                        if (targets) return 1;
                    } else {
                        T[xs * stride] = targets;
                        total += __popcll(targets);
                    }
                }
                if (src_left == 0) break;
                sbit = src_left & (~src_left + 1ull);
                src_left ^= sbit;
                have = true;
                const int cell = __ffsll((long long)sbit) - 1;
                xs = cell - base;
                occS = variant == BGS_BOUNCE_SOURCE_PIECE ? occ : (occ & ~sbit);
                open = variant == BGS_BOUNCE_SOURCE_BLOCKED ? (g.board & ~sbit) : g.board;
                inter = open & ~occS & ~farm;  // cells a path may pass through
                expanded = sbit;
                targets = 0;
                S = sbit;
                u = P.value_at(cell);
            } else {  // bounce: all unexpanded landing cells that hold a piece of the same value
                const int c = __ffsll((long long)pending) - 1;
                u = P.value_at(c);
                S = pending & P.cells_with_value(u);
                pending &= ~S;
                expanded |= S;
            }
            Ff = 0; Fl = 0; Fr = 0; Nn = S;  // no direction memory at the start of a segment
            rem = u;
        }
        // one step of the current segment, all frontier cells at once
        const uint64_t fl = Ff | Fl | Nn, fr = Ff | Fr | Nn;  // may go left / right (no reversal)
        const uint64_t all = fl | Fr;
        const uint64_t nf = player == 0 ? (all << g.W) : (all >> g.W);  // never backwards
        const uint64_t nl = (fl & g.not_left) >> 1;
        const uint64_t nr = (fr & g.not_right) << 1;
        if (rem > 1) {  // intermediate cells: empty, not the far goal row
            Ff = nf & inter; Fl = nl & inter; Fr = nr & inter; Nn = 0;
            rem = (Ff | Fl | Fr) ? rem - 1 : 0;
        } else {        // last step: rest on an empty cell, or bounce off a piece
            const uint64_t land = (nf | nl | nr) & open;
            targets |= land & ~occS;
            pending |= land & occS & ~expanded;
            rem = 0;
        }
    }
    return ANY ? 0 : total;
}

// All targets of the single piece on (cell) -- used by the batched step kernel to validate a move.
template <int NP>
__device__ __forceinline__ uint64_t targets_of(const Geo& g, const Planes<NP>& P, uint64_t occ, int player,
                                               uint64_t sbit, int v) {
    const int variant = g.rules & 3;
    const uint64_t occS = variant == BGS_BOUNCE_SOURCE_PIECE ? occ : (occ & ~sbit);
    const uint64_t open = variant == BGS_BOUNCE_SOURCE_BLOCKED ? (g.board & ~sbit) : g.board;
    const uint64_t inter = open & ~occS & ~g.far(player);
    uint64_t expanded = sbit, pending = 0, targets = 0, S = sbit;
    int u = v;
    for (;;) {
        uint64_t Ff = 0, Fl = 0, Fr = 0, Nn = S, land = 0;
        for (int step = u; step >= 1; --step) {
            const uint64_t fl = Ff | Fl | Nn, fr = Ff | Fr | Nn, all = fl | Fr;
            const uint64_t nf = player == 0 ? (all << g.W) : (all >> g.W);
            const uint64_t nl = (fl & g.not_left) >> 1, nr = (fr & g.not_right) << 1;
            if (step > 1) {
                Ff = nf & inter; Fl = nl & inter; Fr = nr & inter; Nn = 0;
                if (!(Ff | Fl | Fr)) break;
            } else {
                land = (nf | nl | nr) & open;
            }
        }
        targets |= land & ~occS;
        pending |= land & occS & ~expanded;
        if (!pending) break;
        const int c = __ffsll((long long)pending) - 1;
        u = P.value_at(c);
        S = pending & P.cells_with_value(u);
        pending &= ~S;
        expanded |= S;
    }
    if (!(g.rules & BGS_BOUNCE_ALLOW_NULL_MOVE)) targets &= ~sbit;
    return targets;
}

template <int NP>
__device__ __forceinline__ bool has_any(const Geo& g, const Planes<NP>& P, int player) {
    int row;
    return movegen<NP, true>(g, P, player, nullptr, 0, &row) != 0;
}

template <int NP>
__device__ __forceinline__ void move_piece(Planes<NP>& P, int scell, int tcell) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const uint64_t bit = (P.b[i] >> scell) & 1ull;
        P.b[i] &= ~(1ull << scell);
        P.b[i] |= bit << tcell;
    }
}

__device__ __forceinline__ float2 reward_of(int winner) {
    return make_float2(winner == 0 ? 1.f : (winner == 1 ? -1.f : 0.f),
                       winner == 1 ? 1.f : (winner == 0 ? -1.f : 0.f));
}

template <int NP>
__device__ __forceinline__ void store_grid(const Planes<NP>& P, int HW, int8_t* out) {
    for (int c = 0; c < HW; ++c) out[c] = (int8_t)P.value_at(c);
}

// ---------------------------------------------------------------------------------------------
// rollout kernel
// ---------------------------------------------------------------------------------------------
struct RolloutParams {
    uint32_t n_games;  // <= 2^31 per launch
    unsigned long long game_id0;
    uint32_t seed_lo, seed_hi;
    int max_plies;
    uint64_t plane0[4];  // bit-planes of the start position (the Config grid)
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
};

constexpr int ROLLOUT_THREADS = 128;
constexpr int PLY_BATCH = 12;  // lanes that must be ready before the ply transition runs (8: 13.0 ms, 12: 12.45, 16: 12.6, 24: 13.3)

// ---------------------------------------------------------------------------------------------
// rollout kernel, lane formulation: one game per lane (Game + MoveGen of bounce_lane.cuh in
// registers).  Lanes whose move generation is complete wait until PLY_BATCH of them can run the
// ply transition together.
// ---------------------------------------------------------------------------------------------
template <int NP, class G, int RULES>
__global__ void __launch_bounds__(ROLLOUT_THREADS)
bounce_rollout_lane_kernel(const GeoRT grt, const RolloutParams p) {
    __shared__ unsigned int s_hist[HIST_BINS];
    __shared__ uint64_t s_T[8 * ROLLOUT_THREADS];
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint64_t* T = s_T + threadIdx.x;  // T[j * ROLLOUT_THREADS] = targets of the j-th movable piece
    const G g(grt);
    const LaneOut out{p.moves, p.length, p.winner, p.final_grid, p.reward};
    Game<NP, G> gm;
    MoveGen<NP, G, RULES> mg;
    mg.done = false;
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
            gm.begin_planes(g, p.plane0);
        start_movegen(false, no_moves);
    };
    if (active) begin_game();

    for (;;) {
        const unsigned am = __ballot_sync(0xffffffffu, active);
        if (!am) break;
        const unsigned wm = __ballot_sync(0xffffffffu, active && mg.done);
        if (__popc(wm) >= PLY_BATCH || wm == am) {
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
            if (!mg.done) mg.iter(g, T, ROLLOUT_THREADS);  // two segments per readiness check: 12.8 -> 11.9 ms
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
// batched move generation / single step on reference-layout states (int8 grids)
// ---------------------------------------------------------------------------------------------
constexpr int STEP_THREADS = 64;

// Returns false if a cell holds a value outside 0..15.
__device__ __forceinline__ bool load_planes(const int8_t* __restrict__ grid, int HW, Planes<4>& P) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) P.b[i] = 0;
    for (int c = 0; c < HW; ++c) {
        const int v = grid[c];
        if (v < 0 || v > 15) ok = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) P.b[i] |= (uint64_t)((v >> i) & 1) << c;
    }
    return ok;
}

__global__ void __launch_bounds__(STEP_THREADS)
bounce_moves_kernel(const Geo g, unsigned long long n, const int8_t* __restrict__ grid,
                    const int8_t* __restrict__ player, const uint8_t* __restrict__ ended,
                    int8_t* source_row, uint64_t* targets, int32_t* count) {
    __shared__ uint64_t s_T[8 * STEP_THREADS];
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t* T = s_T + threadIdx.x;
    const int HW = g.H * g.W;
    Planes<4> P;
    const bool ok = load_planes(grid + i * HW, HW, P);
    int row = -1, total = 0;
    const bool over = (ended && ended[i]) || !ok;
    const int pl = player[i] & 1;
    if (!over) total = movegen<4, false>(g, P, pl, T, STEP_THREADS, &row);
    for (int x = 0; x < g.W; ++x) targets[i * g.W + x] = over ? 0ull : T[x * STEP_THREADS];
    if (source_row) source_row[i] = (int8_t)(total > 0 ? row : -1);
    if (count) count[i] = ok ? total : -1;
}

__global__ void __launch_bounds__(STEP_THREADS)
bounce_step_kernel(const Geo g, unsigned long long n, const int8_t* __restrict__ grid,
                   const int8_t* __restrict__ player, const int8_t* __restrict__ winner,
                   const uint8_t* __restrict__ ended,
                   const int32_t* __restrict__ move, int8_t* grid_out, int8_t* player_out,
                   int8_t* winner_out, uint8_t* ended_out, float* reward_out, int32_t* status) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int HW = g.H * g.W;
    const int8_t* gi = grid + i * HW;
    int8_t* go = grid_out + i * HW;
    Planes<4> P;
    const bool ok = load_planes(gi, HW, P);
    int pl = player[i] & 1;
    const bool over = ended && ended[i];
    const int sx = move[4 * i + 0], sy = move[4 * i + 1], tx = move[4 * i + 2], ty = move[4 * i + 3];
    bool legal = ok && !over && sx >= 0 && sx < g.W && sy >= 0 && sy < g.H && tx >= 0 && tx < g.W && ty >= 0 && ty < g.H;
    int win = winner ? (int)winner[i] : -1;
    bool end_new = over;
    if (legal) {
        const uint64_t occ = P.occ();
        int row;
        const uint64_t rowm = source_row_mask(g, occ, pl, &row);
        const int scell = sy * g.W + sx, tcell = ty * g.W + tx;
        const uint64_t sbit = 1ull << scell;
        legal = (rowm & occ & sbit) != 0;
        if (legal) legal = (targets_of<4>(g, P, occ, pl, sbit, P.value_at(scell)) >> tcell) & 1ull;
        if (legal) {
            move_piece<4>(P, scell, tcell);
            if ((1ull << tcell) & g.far(pl)) {
                win = pl;
                end_new = true;
            } else if (!has_any<4>(g, P, 1 - pl)) {
                end_new = true;
                if (has_any<4>(g, P, pl)) win = pl;
            }
            pl = 1 - pl;
        }
    }
    if (legal) store_grid<4>(P, HW, go);
    else if (go != gi)
        for (int c = 0; c < HW; ++c) go[c] = gi[c];
    player_out[i] = (int8_t)pl;
    winner_out[i] = (int8_t)win;
    if (ended_out) ended_out[i] = end_new;
    if (reward_out) reinterpret_cast<float2*>(reward_out)[i] = reward_of(win);
    if (status) status[i] = legal ? 0 : 1;
}

static bool supported(int H, int W, int max_value) {
    return H >= 1 && W >= 1 && W <= 8 && H * W <= 64 && max_value <= 15;
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
    const Geo g = make_geo(H, W, rules);
    const unsigned long long blocks = (n + STEP_THREADS - 1) / STEP_THREADS;
    bounce_moves_kernel<<<(unsigned)blocks, STEP_THREADS, 0, (cudaStream_t)stream_>>>(g, n, grid, player, ended,
                                                                                     source_row, targets, count);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
}

extern "C" int bgs_bounce_step(int H, int W, int rules, uint64_t n, const int8_t* grid, const int8_t* player,
                               const int8_t* winner, const uint8_t* ended, const int32_t* move, int8_t* grid_out, int8_t* player_out,
                               int8_t* winner_out, uint8_t* ended_out, float* reward_out, int32_t* status,
                               void* stream_) {
    if (!supported(H, W, 0)) return set_error(BGS_EUNSUPPORTED, "bounce: unsupported board %dx%d", H, W);
    if (!grid || !player || !move || !grid_out || !player_out || !winner_out)
        return set_error(BGS_EINVAL, "bounce_step: null required pointer");
    if (int rc = require_device()) return rc;
    if (n == 0) return BGS_OK;
    const Geo g = make_geo(H, W, rules);
    const unsigned long long blocks = (n + STEP_THREADS - 1) / STEP_THREADS;
    bounce_step_kernel<<<(unsigned)blocks, STEP_THREADS, 0, (cudaStream_t)stream_>>>(
        g, n, grid, player, winner, ended, move, grid_out, player_out, winner_out, ended_out, reward_out, status);
    BGS_CUDA_TRY(cudaGetLastError());
    return BGS_OK;
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
static int launch_bounce_lane(const GeoRT& grt, const RolloutParams& p, cudaStream_t stream) {
    auto kern = bounce_rollout_lane_kernel<NP, G, RULES>;
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
    if (grid0 && H >= 1 && W >= 1 && H * W <= 64)
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
    const GeoRT grt = make_geo_rt(H, W, rules);
    RolloutParams p;
    p.n_games = (uint32_t)n_games; p.game_id0 = game_id0;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
    p.max_plies = max_plies;
    for (int i = 0; i < 4; ++i) p.plane0[i] = 0;
    if (grid0) planes_from_grid(grt, grid0, p.plane0);
    p.moves = moves; p.length = length; p.winner = winner; p.final_grid = final_grid; p.reward = reward;
    p.stats = reinterpret_cast<unsigned long long*>(stats);
    p.start_grid = start_grid; p.start_player = start_player; p.start_winner = start_winner; p.start_ended = start_ended;
    unsigned int* counter = nullptr;
    if (int rc = next_counter(&counter)) return rc;
    p.counter = counter;
    BGS_CUDA_TRY(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), stream));
    if (moves) BGS_CUDA_TRY(cudaMemsetAsync(moves, 0xFF, n_games * (size_t)max_plies * 2, stream));
    // per-game start grids may hold any value up to 15: use the 4-plane kernel
    if (maxv <= 3 && !start_grid) {
        if (H == 9 && W == 6 && rules == 0) return launch_bounce_lane<2, GeoCT<9, 6>, 0>(grt, p, stream);
        return launch_bounce_lane<2, GeoRT, -1>(grt, p, stream);
    }
    return launch_bounce_lane<4, GeoRT, -1>(grt, p, stream);
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

extern "C" int bgs_bounce_rollout_host(int device, const int8_t* grid0, int H, int W, int rules, int max_plies,
                                       uint64_t n, uint64_t game_id0, uint64_t seed, uint8_t* moves,
                                       uint16_t* length, int8_t* winner, int8_t* final_grid, float* reward,
                                       int64_t* stats) {
    if (int rc = require_device()) return rc;
    BGS_CUDA_TRY(cudaSetDevice(device));
    if (n == 0) return BGS_OK;
    if (H < 1 || W < 1 || H * W > 64) return set_error(BGS_EUNSUPPORTED, "bounce: unsupported board %dx%d", H, W);
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
