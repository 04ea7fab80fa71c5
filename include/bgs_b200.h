/*
 * bgs_b200.h -- C ABI of libbgs_b200.so, the B200 (sm_100a) batched board-game rollout engine.
 *
 * This is the drop-in boundary for the reference's one data-parallel hot path: legal-move
 * generation -> state transition -> terminal / reward evaluation inside a random-rollout loop.
 * In the reference that path is reached through two nanobind modules,
 *     src/simulator/game/connect.cpp:24-61   (Config / State / Action of Connect-k)
 *     src/simulator/game/bounce.cpp:24-60    (Config / State / Action of Bounce)
 * one Python call per property per game.  Every entry point below cites the binding(s) it replaces;
 * INTEGRATION.md shows the binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - Every function returns 0 on success and a negative BGS_E* code on failure; the message is in
 *     bgs_last_error() (thread local).  Nothing throws across the ABI.
 *   - Pointers are DEVICE pointers on the current CUDA device unless the function name ends in
 *     `_host`.  `stream` is a cudaStream_t (NULL = legacy default stream).  Device-pointer entry
 *     points are asynchronous with respect to the host; `_host` entry points synchronise.
 *   - Output pointers documented as "optional" may be NULL.
 *   - Grids use the reference's own layout: int8[H][W], row 0 = bottom row
 *     (tests/test_connect.py:25-30, tests/test_bounce.py:24-31).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     BGS_ENODEVICE.
 *   - Alignment: every pointer needs the natural alignment of its element type; on top of that float[n,2]
 *     reward arrays must be 8-byte aligned (they are written as pairs), `actions` of bgs_connect_rollout 2-byte
 *     aligned on boards with an even number of cells, and packed boards (`final_packed`, `packed`), start-record
 *     workspaces, state keys and the inputs of the pack_results calls 16-byte aligned (BGS_EINVAL otherwise).  Byte grids / trajectories may sit anywhere: 16-byte aligned ones take the vectorised
 *     kernels, others a slower path with the same results.
 */
#ifndef BGS_B200_H
#define BGS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGS_VERSION 100 /* 0.1.0 */

/* error codes */
#define BGS_OK 0
#define BGS_EINVAL (-1)       /* bad argument */
#define BGS_EUNSUPPORTED (-2) /* configuration outside what the kernels cover */
#define BGS_ECUDA (-3)        /* CUDA runtime error (message has the cudaError string) */
#define BGS_ENODEVICE (-4)    /* no CUDA device / driver */

/* statistics vector: int64[BGS_STATS_LEN], ACCUMULATED into (zero it first for a fresh count).
 * This is the buffer that is all-reduced (sum) across GPUs. */
#define BGS_STATS_LEN 256
#define BGS_STAT_GAMES 0
#define BGS_STAT_WIN0 1
#define BGS_STAT_WIN1 2
#define BGS_STAT_DRAWS 3
#define BGS_STAT_STEPS 4     /* total env-steps (plies) */
#define BGS_STAT_TRUNCATED 5 /* Bounce only: games cut at max_plies */
#define BGS_STAT_HIST0 16    /* stats[16 + min(length, 239)] = number of games of that length */

/* per-game `winner` codes in batched outputs */
#define BGS_WINNER_DRAW (-1)
#define BGS_WINNER_TRUNCATED (-2)

/* Bounce rule switches left unpinned by the reference's tests. Default 0. */
#define BGS_BOUNCE_SOURCE_EMPTY 0
#define BGS_BOUNCE_SOURCE_BLOCKED 1
#define BGS_BOUNCE_SOURCE_PIECE 2
#define BGS_BOUNCE_ALLOW_NULL_MOVE 4

int bgs_version(void);
const char* bgs_last_error(void);
/* Number of CUDA devices visible (0 if none / no driver). Never fails. */
int bgs_device_count(void);

/* ----------------------------------------------------------------------------------------------
 * Connect-k   --  replaces game::connect::{Config,State,Action} as bound in connect.cpp:24-54
 * -------------------------------------------------------------------------------------------- */

/* 1 if (H, W, K) is covered by the kernels, else 0: K >= 1 and H*W <= 255, W <= 32.  Boards with
 * H*W <= 128, H <= 15, W <= 16 run on bit-word kernels (the tuned path); larger ones on a byte-board
 * fallback that supports bgs_connect_rollout / _rollout_host / _export / _step / _query but not
 * bgs_connect_rollout_from and bgs_connect_trajectory_grids (BGS_EUNSUPPORTED). */
int bgs_connect_supported(int H, int W, int K);

/* Number of uint64 words per game in the packed board format used by *_packed / export:
 * [stones of player 0 | stones of player 1], each (H*W <= 64 ? 1 : 2) words, little-endian word order,
 * bit index of cell (row, col) = (H-1-row)*W + col (the top row occupies bits 0..W-1).
 * Byte-board fallback: the record is the int8 grid itself (row 0 = bottom, -1 empty), padded to a
 * multiple of 8 bytes: (H*W + 7) / 8 words. */
int bgs_connect_packed_words(int H, int W);

/* The whole rollout loop of README.md:49-72 for n_games independent games from the empty board:
 *   Config::sample_initial_state (connect.cpp:32) -> while !has_ended (connect.cpp:39):
 *   actions (connect.cpp:43) -> uniform choice -> sample_next_state (connect.cpp:52) -> reward (connect.cpp:41).
 * Game i uses global id game_id0 + i; the action at ply t of a game is
 *   legal_columns_ascending[ mulhi32( philox4x32_10(key=seed, ctr=(id_lo,id_hi,t>>2,0))[t&3], n_legal ) ].
 * actions      optional uint8[n_games, H*W]  column per ply, 0xFF after the end (requires `length`)
 * length       optional uint8[n_games]       plies played
 * winner       optional int8[n_games]        0 / 1 / BGS_WINNER_DRAW
 * final_packed optional uint64[n_games, bgs_connect_packed_words] final boards (feed to bgs_connect_export)
 * stats        optional int64[BGS_STATS_LEN] accumulated */
int bgs_connect_rollout(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                        uint8_t* actions, uint8_t* length, int8_t* winner, uint64_t* final_packed,
                        int64_t* stats, void* stream);

/* bgs_connect_rollout + bgs_connect_export in ONE call, with every output in the reference's own
 * layouts (State::get_grid / get_reward, connect.cpp:41-42; arrays as tensor.hpp:69-87 hands them to
 * numpy): the loop of README.md:49-72 for n_games games, then per game
 *   actions    optional uint8[n_games, H*W]  column per ply, 0xFF after the end (requires `length`)
 *   length     optional uint8[n_games]; winner optional int8[n_games]
 *   final_grid optional int8[n_games, H, W]  row 0 = bottom, -1 / 0 / 1
 *   reward     optional float[n_games, 2]
 * For Connect(8,9,5) and Connect(10,12,6) -- BASELINE.json configs[3] -- this is a single pass: the
 * rollout kernel itself writes every output byte once (needs `actions` 16-byte and `final_grid` /
 * `reward` 8-byte aligned, else the two-step path runs).  Other boards: rollout + export with
 * stream-ordered temporaries for the packed boards. */
int bgs_connect_rollout_export(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                               uint8_t* actions, uint8_t* length, int8_t* winner, int8_t* final_grid,
                               float* reward, int64_t* stats, void* stream);

/* The same loop from caller-supplied positions (the State objects of connect.cpp:36-46 as tensors:
 * grid int8[n,H,W], player int8[n] = side to move, winner_in optional int8[n], -1 = nobody has won):
 * State::from_json -> while !has_ended: actions -> uniform choice -> sample_next_state.
 * The draw index t counts the plies played in THIS rollout; `length` is that count (0 for a position
 * that has already ended) and `actions` lists only those plies.
 * workspace: uint64[n_games * bgs_connect_start_words(H, W)] device scratch (overwritten). */
int bgs_connect_start_words(int H, int W);
int bgs_connect_rollout_from(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                             const int8_t* grid, const int8_t* player, const int8_t* winner_in,
                             uint64_t* workspace, uint8_t* actions, uint8_t* length, int8_t* winner,
                             uint64_t* final_packed, int64_t* stats, void* stream);

/* Packed positions: the State objects of connect.cpp:36-46 in 17 bytes (H*W <= 64) or 33 bytes each instead
 * of H*W + 2 -- what a host sends over PCIe for rollouts from positions (State::from_json, connect.cpp:45-46).
 *   packed uint64[n, bgs_connect_packed_words(H, W)]  the two bitboards in the packed-board format above
 *   meta   uint8[n]   bit 0 = side to move, bits 1..2 = winner + 1 (0 = nobody has won)
 * bgs_connect_pack writes them from reference-layout states (grid int8[n,H,W], player int8[n], winner optional
 * int8[n]); bgs_connect_rollout_from_packed is bgs_connect_rollout_from on such positions (same draws, same
 * outputs).  Boards of the bit-word kernels only (H*W <= 128, W <= 16, H <= 15); `packed` 16-byte aligned. */
int bgs_connect_pack(int H, int W, uint64_t n, const int8_t* grid, const int8_t* player, const int8_t* winner,
                     uint64_t* packed, uint8_t* meta, void* stream);
int bgs_connect_rollout_from_packed(int H, int W, int K, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                                    const uint64_t* packed, const uint8_t* meta, uint64_t* workspace,
                                    uint8_t* actions, uint8_t* length, int8_t* winner, uint64_t* final_packed,
                                    int64_t* stats, void* stream);

/* State::get_grid (connect.cpp:42) and State::get_reward (connect.cpp:41) for n packed boards:
 * grid optional int8[n,H,W] (-1 / 0 / 1); reward optional float[n,2] from winner int8[n]. */
int bgs_connect_export(int H, int W, uint64_t n, const uint64_t* packed, const int8_t* winner,
                       int8_t* grid, float* reward, void* stream);

/* State::get_grid (connect.cpp:42) at EVERY ply of n recorded games: actions uint8[n, H*W] and length
 * uint8[n] as written by bgs_connect_rollout -> grids int8[n, H*W+1, H, W]; entry t is the position
 * after t plies (entry 0 = empty board), entries past the end of a game repeat its final position.
 * HBM-write bound: (H*W+1) * H*W bytes per game. */
int bgs_connect_trajectory_grids(int H, int W, uint64_t n_games, const uint8_t* actions, const uint8_t* length,
                                 int8_t* grids, void* stream);

/* Per-game results in one byte each, to halve the device->host traffic of State::get_reward
 * (connect.cpp:41) for a whole batch: packed[i] = length[i] | (winner[i] + 1) << 6.  Valid for boards of
 * at most 63 cells (length < 64); winner + 1 is 0 (draw), 1 (player 0), 2 (player 1).  All pointers 16-byte
 * aligned device pointers. */
int bgs_connect_pack_results(uint64_t n, const uint8_t* length, const int8_t* winner, uint8_t* packed, void* stream);
/* The same for boards of 64..127 cells, games played from the empty board: packed[i] = length[i] | draw << 7;
 * the winner of a decided game is the parity of its length (odd: player 0). */
int bgs_connect_pack_results_wide(uint64_t n, const uint8_t* length, const int8_t* winner, uint8_t* packed, void* stream);

/* Dense per-game results for games played from the empty board: such a game ends after Lmin = min(2K-1, H*W) ..
 * H*W plies with the winner given by the parity of its length (odd: player 0), or in a draw, so there are
 * S = H*W - Lmin + 2 symbols (length - Lmin, or S-1 for a draw) and G = floor(16 / log2 S) games fit one
 * uint16: packed[w] = sym[G*w] + S*sym[G*w+1] + S^2*sym[G*w+2] + ...   (6x7x4: S = 37, G = 3, 5.33 bits per game).
 * bgs_connect_dense_results returns G (0: no gain for this board) and Lmin / S through the pointers;
 * packed holds (n + G - 1) / G words. */
int bgs_connect_dense_results(int H, int W, int K, int* lmin, int* symbols);
int bgs_connect_pack_results_dense(int H, int W, int K, uint64_t n, const uint8_t* length, const int8_t* winner,
                                   uint16_t* packed, void* stream);

/* Batched single transition on reference-layout states.  Replaces, for n states at once,
 *   State::get_action_at (connect.cpp:44)  -> status[i] = 0 ok / 1 illegal (state left unchanged)
 *   Action::sample_next_state (connect.cpp:52) -> grid_out, player_out, winner_out
 *   State::has_ended / get_reward / get_actions on the NEW state (connect.cpp:39,41,43)
 * grid int8[n,H,W], player int8[n], winner int8[n] (-1 none), action int32[n] (column).
 * ended_out uint8[n]; reward_out float[n,2]; legal_out uint32[n] (bit c = column c playable in the
 * new state, 0 when ended).  grid_out may alias grid.  All outputs optional except grid_out,
 * player_out, winner_out. */
int bgs_connect_step(int H, int W, int K, uint64_t n, const int8_t* grid, const int8_t* player,
                     const int8_t* winner, const int32_t* action, int8_t* grid_out,
                     int8_t* player_out, int8_t* winner_out, uint8_t* ended_out, float* reward_out,
                     uint32_t* legal_out, int32_t* status, void* stream);

/* bgs_connect_step with the action CHOSEN in the kernel from per-state weights: the agent loop of
 * textual/examples/arena.py:64-68 (`random.choices(actions, weights)` then `sample_next_state`) for n states.
 *   probs float[n, W]: weight of each column; columns that are not playable are ignored.  NaN / negative /
 *   zero count as 0, +inf as FLT_MAX; if every playable column has weight 0 the choice is uniform.
 *   q_c = (uint32)(w_c / max_c w_c * 65535 + 0.5) in IEEE single; r = the Philox draw of DESIGN.md 2 for
 *   global id (game_ids ? game_ids[i] : game_id0 + i) and draw index t = draw_index ? draw_index[i] : number of
 *   stones on the board; the action is the first playable column j with (q_0 + .. + q_j) * 2^32 > r * sum q.
 *   Equal weights give exactly the uniform choice of bgs_connect_rollout (column mulhi32(r, n_legal)), so
 *   stepping n boards from empty with constant probs replays the rollout kernel's games ply by ply.
 * action_out optional int32[n] = chosen column (-1 where the state had ended; status[i] = 1 there). */
int bgs_connect_sample_step(int H, int W, int K, uint64_t n, const int8_t* grid, const int8_t* player,
                            const int8_t* winner, const float* probs, uint64_t seed, uint64_t game_id0,
                            const uint64_t* game_ids, const int32_t* draw_index, int8_t* grid_out,
                            int8_t* player_out, int8_t* winner_out, uint8_t* ended_out, float* reward_out,
                            uint32_t* legal_out, int32_t* action_out, int32_t* status, void* stream);

/* ==, <, hash of n states at once (helper.hpp:10-25 gives every reference object == != < <= > >= and
 * __hash__; their values are not pinned): keys uint64[n, 2] (16-byte aligned), equal for two states of one
 * configuration iff grid, player and winner are equal.  H*W <= 62: an exact, invertible packing
 *   b0 | b1 << HW | player << 2HW | (winner + 1) << (2HW + 1),  bit r*W + c of b0 / b1 = stone of player 0 / 1
 * on row r (from the bottom), column c.  Larger boards: a 128-bit hash (csrc/keys.cu states the function). */
int bgs_connect_keys(int H, int W, uint64_t n, const int8_t* grid, const int8_t* player, const int8_t* winner,
                     uint64_t* keys, void* stream);

/* State::has_ended / get_actions / get_reward (connect.cpp:39,41,43) of n existing states. */
int bgs_connect_query(int H, int W, uint64_t n, const int8_t* grid, const int8_t* winner,
                      uint8_t* ended_out, uint32_t* legal_out, float* reward_out, void* stream);

/* Host-buffer convenience wrapper of bgs_connect_rollout + bgs_connect_export: allocates device
 * buffers on `device`, runs, copies the requested outputs back and synchronises.
 * final_grid optional int8[n,H,W]; reward optional float[n,2]; others as bgs_connect_rollout. */
int bgs_connect_rollout_host(int device, int H, int W, int K, uint64_t n_games, uint64_t game_id0,
                             uint64_t seed, uint8_t* actions, uint8_t* length, int8_t* winner,
                             int8_t* final_grid, float* reward, int64_t* stats);

/* ----------------------------------------------------------------------------------------------
 * Bounce   --  replaces game::bounce::{Config,State,Action} as bound in bounce.cpp:24-53
 * Boards with H*W <= 128 cells, W <= 16, piece values 1..15 (one 64-bit board word per value plane up to
 * 64 cells and 8 columns, an unsigned __int128 beyond).
 * -------------------------------------------------------------------------------------------- */

int bgs_bounce_supported(int H, int W, int max_value);

/* State::get_actions / get_actions_at (bounce.cpp:40-41) for n states:
 * source_row int8[n]  row of the mover's movable pieces (-1: none / ended)
 * targets uint64[n,W] bit (y*W+x) of targets[i][sx] set <=> (sx, source_row) -> (x, y) is legal;
 *         boards with W > 8 or H*W > 64 use two words per mask: uint64[n,W,2] (bits 0..63, bits 64..127)
 * count   optional int32[n] total number of legal actions. `ended` optional uint8[n]. */
int bgs_bounce_moves(int H, int W, int rules, uint64_t n, const int8_t* grid, const int8_t* player,
                     const uint8_t* ended, int8_t* source_row, uint64_t* targets, int32_t* count,
                     void* stream);

/* Batched single transition (Action::sample_next_state bounce.cpp:51, State::get_action_at :42).
 * move int32[n,4] = (sx, sy, tx, ty). status[i] = 0 ok / 1 illegal (state copied unchanged).
 * winner (optional int8[n], -1 none) and ended (optional uint8[n]) describe the CURRENT states.
 * Outputs describe the NEW state; reward_out float[n,2], ended_out uint8[n]. */
int bgs_bounce_step(int H, int W, int rules, uint64_t n, const int8_t* grid, const int8_t* player,
                    const int8_t* winner, const uint8_t* ended, const int32_t* move, int8_t* grid_out, int8_t* player_out,
                    int8_t* winner_out, uint8_t* ended_out, float* reward_out, int32_t* status,
                    void* stream);

/* bgs_bounce_step with the move chosen in the kernel from per-state weights (arena.py:64-68):
 * probs float[n, W, H*W], probs[i, sx, ty*W + tx] = weight of moving the piece in column sx of the mover's
 * source row to (tx, ty); quantisation, draw and choice as bgs_connect_sample_step over the legal actions in
 * canonical order (ascending sx, then target cell), RNG domain 1, t = draw_index ? draw_index[i] : 0.
 * Equal weights give the uniform choice of bgs_bounce_rollout.  move_out optional int32[n,4] (-1s when the
 * state had ended or the mover is blocked; status[i] = 1 there). */
int bgs_bounce_sample_step(int H, int W, int rules, uint64_t n, const int8_t* grid, const int8_t* player,
                           const int8_t* winner, const uint8_t* ended, const float* probs, uint64_t seed,
                           uint64_t game_id0, const uint64_t* game_ids, const int32_t* draw_index,
                           int8_t* grid_out, int8_t* player_out, int8_t* winner_out, uint8_t* ended_out,
                           float* reward_out, int32_t* move_out, int32_t* status, void* stream);

/* 128-bit hash keys of n Bounce states (see bgs_connect_keys; always the hash form). */
int bgs_bounce_keys(int H, int W, uint64_t n, const int8_t* grid, const int8_t* player, const int8_t* winner,
                    uint64_t* keys, void* stream);

/* Uniform-random rollouts from grid0 (HOST pointer, int8[H,W]; it is the Config, bounce.cpp:26).
 * Action order for the uniform choice: ascending (sy, sx, ty, tx); RNG as for Connect with ctr[3]=1.
 * moves       optional uint8[n, max_plies, 2] (source cell, target cell), cell = y*W+x, 0xFF padded
 * length      optional uint16[n]; winner optional int8[n] (0/1/DRAW/TRUNCATED)
 * final_grid  optional int8[n,H,W]; reward optional float[n,2]; stats optional, accumulated. */
int bgs_bounce_rollout(const int8_t* grid0_host, int H, int W, int rules, int max_plies,
                       uint64_t n_games, uint64_t game_id0, uint64_t seed, uint8_t* moves,
                       uint16_t* length, int8_t* winner, int8_t* final_grid, float* reward,
                       int64_t* stats, void* stream);

/* The same loop from per-game positions (BounceBatch tensors): grid int8[n,H,W] (values 0..15), player
 * int8[n], winner_in optional int8[n] (-1 none), ended_in optional uint8[n]; all DEVICE pointers.  Draws
 * and `length` count the plies played in this rollout.
 * A start position whose side to move has no action but is not flagged (winner_in / ended_in) ends at
 * once with length 0 and NO winner (README.md:60 promises an action whenever has_ended is false; the
 * blocked rule of Action::sample_next_state needs the move that led here).  Callers that want the blocked
 * rule applied pass ended_in / winner_in from bgs_bounce_step, which evaluates it when the move is made. */
int bgs_bounce_rollout_from(int H, int W, int rules, int max_plies, uint64_t n_games, uint64_t game_id0,
                            uint64_t seed, const int8_t* grid, const int8_t* player, const int8_t* winner_in,
                            const uint8_t* ended_in, uint8_t* moves, uint16_t* length, int8_t* winner,
                            int8_t* final_grid, float* reward, int64_t* stats, void* stream);

/* Per-game Bounce results in two bytes instead of three: packed[i] = length[i] | (winner[i] + 2) << 14
 * (lengths below 16384; winner + 2: 0 truncated, 1 draw, 2 player 0, 3 player 1). */
int bgs_bounce_pack_results(uint64_t n, const uint16_t* length, const int8_t* winner, uint16_t* packed, void* stream);

int bgs_bounce_rollout_host(int device, const int8_t* grid0_host, int H, int W, int rules,
                            int max_plies, uint64_t n_games, uint64_t game_id0, uint64_t seed,
                            uint8_t* moves, uint16_t* length, int8_t* winner, int8_t* final_grid,
                            float* reward, int64_t* stats);

#ifdef __cplusplus
}
#endif
#endif /* BGS_B200_H */
