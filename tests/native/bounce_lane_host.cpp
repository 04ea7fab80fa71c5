// Host build of the Bounce lane state machine (csrc/bounce_lane.cuh) -- TEST HARNESS ONLY.
//
// Plays games one lane at a time on the CPU with exactly the code the CUDA kernel inlines, so that the
// move generation, the mover-relative orientation and the action-selection map can be compared with the
// oracle in the CPU test suite (tests/test_bounce_lane_host.py).  The product never loads this.
#include <cstring>
#include <vector>

#include "../../board-game-simulator-python_b200/csrc/bounce_lane.cuh"

using namespace bgs::bounce;

template <int NP, class G, int RULES>
static void play(const G& g, const typename G::bits* plane0, const int8_t* start_grid, const int8_t* start_player,
                 const int8_t* start_winner, const uint8_t* start_ended, int max_plies, uint64_t n, uint64_t gid0,
                 uint64_t seed, const LaneOut& out, int64_t* stats) {
    std::vector<typename G::bits> T(16);
    std::vector<uint32_t> lut(SEG_LUT_WORDS);
    for (int i = 0; i < SEG_LUT_WORDS; ++i) lut[seg_lut_slot(g, i >> 8, (uint32_t)(i & 255))] = seg_lut_entry(g.s(), i >> 8, (uint32_t)(i & 255));
    for (uint64_t idx = 0; idx < n; ++idx) {
        Game<NP, G> game;
        MoveGen<NP, G, RULES> mg;
        mg.lut = lut.data();
        bool no_moves = false;
        if (start_grid)
            no_moves = game.begin_grid(g, start_grid + idx * (size_t)(g.h() * g.w()), start_player[idx],
                                       start_winner ? (int)start_winner[idx] : BGS_WINNER_DRAW,
                                       start_ended && start_ended[idx]);
        else
            game.begin_planes(g, plane0);
        Next next = NEXT_MOVEGEN;
        while (next != NEXT_OVER) {
            for (int i = 0; i < NP; ++i) mg.b[i] = game.b[i];
            mg.begin(g, next == NEXT_PROBE, no_moves);
            no_moves = false;
            while (!mg.done) mg.iter(g, T.data(), 1);
            uint8_t* row = out.moves ? out.moves + idx * (size_t)max_plies * 2 : nullptr;
            const uint64_t gid = gid0 + idx;
            next = game.transition(g, T.data(), 1, mg.total, mg.probe, mg.found, max_plies, row, [&](int t) {
                return bounce_draw(gid, (uint32_t)seed, (uint32_t)(seed >> 32), t);
            });
        }
        game.write_result(g, out, idx);
        if (stats) {
            stats[BGS_STAT_GAMES] += 1;
            stats[BGS_STAT_WIN0] += game.win == 0;
            stats[BGS_STAT_WIN1] += game.win == 1;
            stats[BGS_STAT_DRAWS] += game.win == BGS_WINNER_DRAW;
            stats[BGS_STAT_TRUNCATED] += game.win == BGS_WINNER_TRUNCATED;
            stats[BGS_STAT_STEPS] += game.t;
            const int cap = BGS_STATS_LEN - BGS_STAT_HIST0 - 1;
            stats[BGS_STAT_HIST0 + (game.t < cap ? game.t : cap)] += 1;
        }
    }
}

// mode 0: run-time geometry and rules; mode 1: compile-time 9x6 board with compile-time rules 0
extern "C" int bgs_lane_host_bounce_rollout(int mode, const int8_t* grid0, const int8_t* start_grid,
                                            const int8_t* start_player, const int8_t* start_winner,
                                            const uint8_t* start_ended, int H, int W, int rules, int max_plies,
                                            uint64_t n, uint64_t gid0, uint64_t seed, uint8_t* moves,
                                            uint16_t* length, int8_t* winner, int8_t* final_grid, float* reward,
                                            int64_t* stats) {
    if (H < 1 || W < 1 || W > 16 || H * W > 128) return -1;
    int maxv = 0;
    if (grid0)
        for (int c = 0; c < H * W; ++c)
            if (grid0[c] > maxv) maxv = grid0[c];
    if (W > 8 || H * W > 64) {  // 128-bit board words
        if (mode == 1) return -2;
        const GeoRT128 g = make_geo_rt_b<u128>(H, W, rules);
        u128 plane0[4] = {0, 0, 0, 0};
        if (grid0) planes_from_grid(g, grid0, plane0);
        if (moves) memset(moves, 0xFF, n * (size_t)max_plies * 2);
        const LaneOut out{moves, length, winner, final_grid, reward};
        play<4, GeoRT128, -1>(g, plane0, start_grid, start_player, start_winner, start_ended, max_plies, n, gid0, seed, out, stats);
        return 0;
    }
    GeoRT g = make_geo_rt(H, W, rules);
    // table-driven segments only when the goal rows hold no piece (as the kernel's host code decides)
    bool goal_rows_empty = grid0 != nullptr;
    if (grid0)
        for (int x = 0; x < W; ++x) goal_rows_empty = goal_rows_empty && grid0[x] == 0 && grid0[(H - 1) * W + x] == 0;
    g.lut_ok = g.lut_ok && goal_rows_empty;
    uint64_t plane0[4] = {0, 0, 0, 0};
    if (grid0) planes_from_grid(g, grid0, plane0);
    if (moves) memset(moves, 0xFF, n * (size_t)max_plies * 2);
    const LaneOut out{moves, length, winner, final_grid, reward};
    if (mode == 1) {
        if (H != 9 || W != 6 || rules != 0 || maxv > 3 || start_grid) return -2;
        play<2, GeoCT<9, 6>, 0>(GeoCT<9, 6>(g), plane0, nullptr, nullptr, nullptr, nullptr, max_plies, n, gid0, seed, out, stats);
    } else if (maxv <= 3 && !start_grid) {
        play<2, GeoRT, -1>(g, plane0, start_grid, start_player, start_winner, start_ended, max_plies, n, gid0, seed, out, stats);
    } else {
        play<4, GeoRT, -1>(g, plane0, start_grid, start_player, start_winner, start_ended, max_plies, n, gid0, seed, out, stats);
    }
    return 0;
}

// 1 when the multiplicative hashes that index the segment tables (seg_hash_mul: row strides 4..9, the default 9x6
// board is 7) map the 256 subsets of the 8 window bits onto 256 different slots
extern "C" int bgs_lane_host_seg_hash_is_perfect() {
    static_assert(GeoCT<9, 6>::HASH, "the default board uses the hashed table");
    int hashed = 0;
    for (int S = 3; S <= 9; ++S) {
        if (!seg_hash_is_perfect(S)) return 0;
        hashed += seg_hash_mul(S) != 0u;
    }
    return hashed == 6 ? 1 : 0;
}
