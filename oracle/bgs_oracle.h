/*
 * bgs_oracle.h -- CPU oracle for the rollout hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a scalar, grid-based restatement of the rules of the two games bound by
 * the reference (src/simulator/game/connect.cpp:24-54, src/simulator/game/bounce.cpp:24-53).
 * The arithmetic itself lives in the third-party header library
 *   github.com/jojolebarjos/board-game-simulator @ c8f8a075cc82ae91732627ca47640338736a40cb
 * (CMakeLists.txt:12-18 of the reference), which is NOT available in this environment, so the
 * rules are restated from the reference's own tests (tests/test_connect.py, tests/test_bounce.py),
 * the binding signatures and the README loop (README.md:38-73).
 *
 * Pinning: the oracle passes the reference's 9 tests run unmodified and in place
 * (tests/test_reference_in_place.py) and the golden vectors extracted from them
 * (tests/golden/).  Points the reference tests do not pin are listed in DESIGN.md
 * ("parity unpinned" items) and are explicit switches here (BGSO_BOUNCE_*).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libbgs_b200.so) never links or calls it.
 */
#ifndef BGS_ORACLE_H
#define BGS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- shared conventions (ours, not the reference's; see DESIGN.md "action-selection map") ---- */
#define BGSO_STATS_LEN 256
#define BGSO_STAT_GAMES 0
#define BGSO_STAT_WIN0 1
#define BGSO_STAT_WIN1 2
#define BGSO_STAT_DRAWS 3
#define BGSO_STAT_STEPS 4
#define BGSO_STAT_TRUNCATED 5
#define BGSO_STAT_HIST0 16 /* stats[16 + min(length, 239)] += 1 */

#define BGSO_DOMAIN_CONNECT 0u
#define BGSO_DOMAIN_BOUNCE 1u

/* Bounce rule switches that the reference tests leave unpinned. */
#define BGSO_BOUNCE_SOURCE_EMPTY 0      /* default: vacated source cell is empty during the search */
#define BGSO_BOUNCE_SOURCE_BLOCKED 1    /* source cell impassable */
#define BGSO_BOUNCE_SOURCE_PIECE 2      /* moving piece left on the grid (can bounce off itself) */
#define BGSO_BOUNCE_ALLOW_NULL_MOVE 4   /* flag: a move may end on its own source cell */

/* Philox4x32-10 (Salmon et al. 2011, Random123), the counter RNG used by both sides. */
void bgso_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The t-th draw of game `gid` under `seed` in `domain`. */
uint32_t bgso_draw(uint64_t seed, uint64_t gid, uint32_t t, uint32_t domain);

/* ---- Connect-k (rules: SURVEY.md 4.4; pins: tests/test_connect.py:24-54,63-115) ---- */
/* grid: int8[H*W], row 0 = bottom, -1 empty, 0 / 1 = player stones. */
int bgso_connect_ended(const int8_t* grid, int H, int W, int winner);
/* Legal columns in ascending order; returns the count (0 when ended). */
int bgso_connect_actions(const int8_t* grid, int H, int W, int winner, int32_t* cols);
/* Transition. Returns 0, or -1 if the move is illegal (column out of range / full / game ended). */
int bgso_connect_next(const int8_t* grid, int H, int W, int K, int player, int winner, int col,
                      int8_t* grid_out, int* player_out, int* winner_out);
/* reward[2] for a terminal (or running: zeros) state */
void bgso_reward(int winner, float* reward2);

/* Uniform-random rollouts from the empty board. Any output pointer may be NULL.
 * actions: uint8[n, H*W] padded with 0xFF; length: uint8[n]; winner: int8[n] (0,1 or -1 = draw);
 * final_grid: int8[n,H,W]; reward: float[n,2]; stats: int64[BGSO_STATS_LEN] (accumulated into). */
int bgso_connect_rollout(int H, int W, int K, uint64_t n, uint64_t gid0, uint64_t seed,
                         uint8_t* actions, uint8_t* length, int8_t* winner, int8_t* final_grid,
                         float* reward, int64_t* stats);
/* Uniform-random rollouts from caller-supplied positions: grid int8[n,H,W], player int8[n] (side to
 * move), winner_in int8[n] (-1 = nobody has won yet).  Draw t of game i is bgso_draw(seed, gid0+i, t, 0)
 * with t = plies played IN THIS ROLLOUT.  length = plies played in the rollout (0 for a position that
 * has already ended); the other outputs as bgso_connect_rollout. */
int bgso_connect_rollout_from(int H, int W, int K, uint64_t n, uint64_t gid0, uint64_t seed,
                              const int8_t* grid, const int8_t* player, const int8_t* winner_in,
                              uint8_t* actions, uint8_t* length, int8_t* winner, int8_t* final_grid,
                              float* reward, int64_t* stats);
/* Replays recorded trajectories through bgso_connect_next, checking at every ply that the move is
 * legal and the game has not ended, and at the end that it HAS ended exactly at `length`, and that
 * winner / final_grid / reward (each optional) are identical.  Returns the number of games with any
 * mismatch; *first_bad receives the index of the first one (or -1). */
int64_t bgso_connect_replay(int H, int W, int K, uint64_t n, const uint8_t* actions,
                            const uint8_t* length, const int8_t* winner, const int8_t* final_grid,
                            const float* reward, int64_t* first_bad);

/* ---- Bounce (rules: SURVEY.md 4.4; pins: tests/test_bounce.py) ---- */
/* grid: int8[H*W], row 0 = bottom, 0 empty, v>0 piece of value v.  Coordinates are (x, y). */
/* Row holding the movable pieces of `player` (-1 if the board has no piece). */
int bgso_bounce_source_row(const int8_t* grid, int H, int W, int player);
/* Targets of the piece at (sx, sy) for `player` as a 0/1 map uint8[H*W]; returns the count.
 * Does not check that the piece is movable. */
int bgso_bounce_targets(const int8_t* grid, int H, int W, int player, int sx, int sy, int rules,
                        uint8_t* target_map);
/* All legal actions in ascending (sy, sx, ty, tx) order as int32[count,4] = (sx,sy,tx,ty);
 * returns the count (0 when ended; `winner` >= 0 or draw flag). cap = capacity in actions. */
int bgso_bounce_actions(const int8_t* grid, int H, int W, int player, int ended, int rules,
                        int32_t* moves, int cap);
/* Transition. status_out: winner_out in {-1,0,1}, ended_out in {0,1}. Returns 0 or -1 if illegal. */
int bgso_bounce_next(const int8_t* grid, int H, int W, int player, int ended, int rules, int sx,
                     int sy, int tx, int ty, int8_t* grid_out, int* player_out, int* winner_out,
                     int* ended_out);
/* Uniform-random rollouts from grid0. moves: uint8[n, max_plies, 2] = (source cell, target cell),
 * cell = y*W+x, padded 0xFF; length: uint16[n]; winner: int8[n] (0, 1, -1 draw, -2 truncated). */
int bgso_bounce_rollout(const int8_t* grid0, int H, int W, int rules, int max_plies, uint64_t n,
                        uint64_t gid0, uint64_t seed, uint8_t* moves, uint16_t* length,
                        int8_t* winner, int8_t* final_grid, float* reward, int64_t* stats);
/* The same from per-game positions: grids int8[n,H,W], player int8[n], winner_in / ended_in optional. */
int bgso_bounce_rollout_from(const int8_t* grids, const int8_t* player, const int8_t* winner_in,
                             const uint8_t* ended_in, int H, int W, int rules, int max_plies, uint64_t n,
                             uint64_t gid0, uint64_t seed, uint8_t* moves, uint16_t* length, int8_t* winner,
                             int8_t* final_grid, float* reward, int64_t* stats);
int64_t bgso_bounce_replay(const int8_t* grid0, int H, int W, int rules, int max_plies, uint64_t n,
                           const uint8_t* moves, const uint16_t* length, const int8_t* winner,
                           const int8_t* final_grid, const float* reward, int64_t* first_bad);

/* ---- fast_connect.c: the same Connect rollout loop on single-word bitboards (the "best CPU" baseline) ---- */
int bgso_fast_connect_supported(int H, int W, int K);
int bgso_fast_connect_rollout(int H, int W, int K, uint64_t n, uint64_t gid0, uint64_t seed, uint8_t* length,
                              int8_t* winner, int64_t* stats);

/* ---- conventions of the B200 build's own additions (include/bgs_b200.h), restated for the tests ---- */
/* weighted action choice (bgs_connect_sample_step / bgs_bounce_sample_step) */
int bgso_connect_sample(const int8_t* grid, int H, int W, int winner, const float* probs, uint64_t seed,
                        uint64_t gid, uint32_t t);
int bgso_bounce_sample(const int8_t* grid, int H, int W, int player, int ended, int rules, const float* probs,
                       uint64_t seed, uint64_t gid, uint32_t t, int32_t* move4);
/* 128-bit state keys (bgs_connect_keys: game 1, bgs_bounce_keys: game 2) */
void bgso_state_key(int game, const int8_t* grid, int H, int W, int player, int winner, uint64_t* key);

#ifdef __cplusplus
}
#endif
#endif
