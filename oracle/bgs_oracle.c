/*
 * bgs_oracle.c -- CPU oracle for the rollout hot path.  TEST INFRASTRUCTURE ONLY (see bgs_oracle.h).
 *
 * Deliberately a different algorithm from the CUDA kernels: int8 grids, nested loops, a brute-force
 * scan of the whole board for k-in-a-row, and an explicit memoised depth-first search for Bounce.
 * No bitboards anywhere.  Each function cites the reference binding / test it restates.
 */
#include "bgs_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10.  Not part of the reference (its games are deterministic; randomness lives in the
 * caller, README.md:61-62 `random.choice(actions)`).  The counter RNG + mulhi map is OUR definition
 * of "uniform choice among legal actions" so that CPU and GPU pick the same action for (seed,
 * game id, ply).  See DESIGN.md "action-selection map".
 * ---------------------------------------------------------------------------------------------- */
void bgso_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t bgso_draw(uint64_t seed, uint64_t gid, uint32_t t, uint32_t domain) {
    uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), t >> 2, domain};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    bgso_philox4x32_10(ctr, key, out);
    return out[t & 3u];
}

static inline uint32_t pick(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

void bgso_reward(int winner, float* reward2) {
    /* tests/test_connect.py:115 ([1,-1]); tests/test_bounce.py:152,278,362 ([-1,1], [1,-1], [0,0]) */
    reward2[0] = winner == 0 ? 1.0f : (winner == 1 ? -1.0f : 0.0f);
    reward2[1] = winner == 1 ? 1.0f : (winner == 0 ? -1.0f : 0.0f);
}

static void stats_game(int64_t* stats, int winner, int length, int truncated) {
    if (!stats) return;
    stats[BGSO_STAT_GAMES] += 1;
    if (truncated) stats[BGSO_STAT_TRUNCATED] += 1;
    else if (winner == 0) stats[BGSO_STAT_WIN0] += 1;
    else if (winner == 1) stats[BGSO_STAT_WIN1] += 1;
    else stats[BGSO_STAT_DRAWS] += 1;
    stats[BGSO_STAT_STEPS] += length;
    int bin = length < (BGSO_STATS_LEN - BGSO_STAT_HIST0 - 1) ? length : (BGSO_STATS_LEN - BGSO_STAT_HIST0 - 1);
    stats[BGSO_STAT_HIST0 + bin] += 1;
}

/* ================================================================================================
 * Connect-k.   State = grid int8[H,W] (row 0 bottom; -1 / 0 / 1; tests/test_connect.py:24-30),
 * player (0 first, alternates; :33-38), winner (-1 none; :136).
 * ============================================================================================== */

static int connect_full(const int8_t* g, int H, int W) {
    for (int c = 0; c < W; ++c)
        if (g[(H - 1) * W + c] < 0) return 0;
    return 1;
}

/* has_ended (connect.cpp:39): a winner exists (test_connect.py:107-115) or the board is full
 * (draw; not pinned by the reference tests -- standard rule, by analogy with test_bounce.py:360). */
int bgso_connect_ended(const int8_t* grid, int H, int W, int winner) {
    return winner >= 0 || connect_full(grid, H, W);
}

/* state.actions (connect.cpp:43): one action per non-full column, ascending; none when ended. */
int bgso_connect_actions(const int8_t* grid, int H, int W, int winner, int32_t* cols) {
    if (bgso_connect_ended(grid, H, W, winner)) return 0;
    int n = 0;
    for (int c = 0; c < W; ++c)
        if (grid[(H - 1) * W + c] < 0) cols[n++] = c;
    return n;
}

/* Brute force: does `who` own K (or more) consecutive cells anywhere, in any of the 4 directions? */
static int connect_has_run(const int8_t* g, int H, int W, int K, int who) {
    static const int DR[4] = {0, 1, 1, 1};
    static const int DC[4] = {1, 0, 1, -1};
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c)
            for (int d = 0; d < 4; ++d) {
                int k = 0;
                int rr = r, cc = c;
                while (k < K && rr >= 0 && rr < H && cc >= 0 && cc < W && g[rr * W + cc] == who) {
                    ++k;
                    rr += DR[d];
                    cc += DC[d];
                }
                if (k >= K) return 1;
            }
    return 0;
}

/* action.sample_next_state (connect.cpp:52): the stone lands on the lowest empty cell of the column
 * (test_connect.py:85-104), K in a row wins (:107-115), the side to move flips (:53-54; also at
 * terminal -- unpinned). Illegal moves are errors (textual/connect.py:115-118 expects RuntimeError). */
int bgso_connect_next(const int8_t* grid, int H, int W, int K, int player, int winner, int col,
                      int8_t* grid_out, int* player_out, int* winner_out) {
    if (bgso_connect_ended(grid, H, W, winner)) return -1;
    if (col < 0 || col >= W) return -1;
    int row = 0;
    while (row < H && grid[row * W + col] >= 0) ++row;
    if (row >= H) return -1;
    if (grid_out != grid) memcpy(grid_out, grid, (size_t)H * W);
    grid_out[row * W + col] = (int8_t)player;
    *winner_out = connect_has_run(grid_out, H, W, K, player) ? player : -1;
    *player_out = 1 - player;
    return 0;
}

int bgso_connect_rollout(int H, int W, int K, uint64_t n, uint64_t gid0, uint64_t seed,
                         uint8_t* actions, uint8_t* length, int8_t* winner, int8_t* final_grid,
                         float* reward, int64_t* stats) {
    if (H < 1 || W < 1 || K < 1 || H * W > 255) return -1;
    const int HW = H * W;
    int8_t* g = (int8_t*)malloc((size_t)HW);
    int32_t* cols = (int32_t*)malloc(sizeof(int32_t) * (size_t)W);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t gid = gid0 + i;
        memset(g, -1, (size_t)HW);
        int player = 0, win = -1, t = 0;
        if (actions) memset(actions + i * HW, 0xFF, (size_t)HW);
        for (;;) {
            int nl = bgso_connect_actions(g, H, W, win, cols);
            if (nl == 0) break;
            uint32_t r = bgso_draw(seed, gid, (uint32_t)t, BGSO_DOMAIN_CONNECT);
            int col = cols[pick(r, (uint32_t)nl)];
            bgso_connect_next(g, H, W, K, player, win, col, g, &player, &win);
            if (actions) actions[i * HW + t] = (uint8_t)col;
            ++t;
        }
        if (length) length[i] = (uint8_t)t;
        if (winner) winner[i] = (int8_t)win;
        if (final_grid) memcpy(final_grid + i * HW, g, (size_t)HW);
        if (reward) bgso_reward(win, reward + 2 * i);
        stats_game(stats, win, t, 0);
    }
    free(g);
    free(cols);
    return 0;
}

int bgso_connect_rollout_from(int H, int W, int K, uint64_t n, uint64_t gid0, uint64_t seed,
                              const int8_t* grid, const int8_t* player_in, const int8_t* winner_in,
                              uint8_t* actions, uint8_t* length, int8_t* winner, int8_t* final_grid,
                              float* reward, int64_t* stats) {
    if (H < 1 || W < 1 || K < 1 || H * W > 255) return -1;
    const int HW = H * W;
    int8_t* g = (int8_t*)malloc((size_t)HW);
    int32_t* cols = (int32_t*)malloc(sizeof(int32_t) * (size_t)W);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t gid = gid0 + i;
        memcpy(g, grid + i * HW, (size_t)HW);
        int player = player_in[i], win = winner_in ? winner_in[i] : -1, t = 0;
        if (actions) memset(actions + i * HW, 0xFF, (size_t)HW);
        for (;;) {
            int nl = bgso_connect_actions(g, H, W, win, cols);
            if (nl == 0) break;
            uint32_t r = bgso_draw(seed, gid, (uint32_t)t, BGSO_DOMAIN_CONNECT);
            int col = cols[pick(r, (uint32_t)nl)];
            bgso_connect_next(g, H, W, K, player, win, col, g, &player, &win);
            if (actions) actions[i * HW + t] = (uint8_t)col;
            ++t;
        }
        if (length) length[i] = (uint8_t)t;
        if (winner) winner[i] = (int8_t)win;
        if (final_grid) memcpy(final_grid + i * HW, g, (size_t)HW);
        if (reward) bgso_reward(win, reward + 2 * i);
        stats_game(stats, win, t, 0);
    }
    free(g);
    free(cols);
    return 0;
}

int64_t bgso_connect_replay(int H, int W, int K, uint64_t n, const uint8_t* actions,
                            const uint8_t* length, const int8_t* winner, const int8_t* final_grid,
                            const float* reward, int64_t* first_bad) {
    const int HW = H * W;
    int8_t* g = (int8_t*)malloc((size_t)HW);
    int64_t bad = 0;
    if (first_bad) *first_bad = -1;
    for (uint64_t i = 0; i < n; ++i) {
        memset(g, -1, (size_t)HW);
        int player = 0, win = -1, ok = 1;
        const int len = length[i];
        if (len > HW) ok = 0;
        for (int t = 0; ok && t < len; ++t) {
            if (player != (t & 1)) ok = 0; /* state.player alternates 0,1,0,... (test_connect.py:33-38) */
            if (bgso_connect_ended(g, H, W, win)) ok = 0;
            else if (bgso_connect_next(g, H, W, K, player, win, actions[i * HW + t], g, &player, &win)) ok = 0;
        }
        if (ok && !bgso_connect_ended(g, H, W, win)) ok = 0;
        for (int t = len; ok && t < HW; ++t)
            if (actions[i * HW + t] != 0xFF) ok = 0;
        if (ok && winner && winner[i] != win) ok = 0;
        if (ok && final_grid && memcmp(final_grid + i * HW, g, (size_t)HW)) ok = 0;
        if (ok && reward) {
            float rw[2];
            bgso_reward(win, rw);
            if (rw[0] != reward[2 * i] || rw[1] != reward[2 * i + 1]) ok = 0;
        }
        if (!ok) {
            if (first_bad && *first_bad < 0) *first_bad = (int64_t)i;
            ++bad;
        }
    }
    free(g);
    return bad;
}

/* ================================================================================================
 * Bounce.   State = grid int8[H,W] (row 0 bottom, 0 empty, v>0 a piece of value v, no owner;
 * tests/test_bounce.py:24-31), player (0 first, plays "up"; player 1 plays "down"; :43-48), winner.
 * ============================================================================================== */

/* Movable pieces of player p = every piece in the occupied row nearest to p's own side
 * (test_bounce.py:43-48,60; SURVEY 4.4 rule 2). */
int bgso_bounce_source_row(const int8_t* grid, int H, int W, int player) {
    if (player == 0) {
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                if (grid[y * W + x] > 0) return y;
    } else {
        for (int y = H - 1; y >= 0; --y)
            for (int x = 0; x < W; ++x)
                if (grid[y * W + x] > 0) return y;
    }
    return -1;
}

typedef struct {
    const int8_t* g; /* grid during the search (source cell already adjusted for the rule variant) */
    int H, W, fwd, far_row, maxv;
    int bx, by; /* impassable cell (SOURCE_BLOCKED) or -1 */
    uint8_t* seen; /* [H*W][maxv+1][4] */
    uint8_t* out;  /* [H*W] */
} bsearch_t;

/* Depth-first search over (x, y, steps left in this segment, last direction) -- SURVEY 4.4 rule 3.
 * Directions: 0 = forward, 1 = left, 2 = right, 3 = none (start of a segment). */
static void bounce_dfs(bsearch_t* s, int x, int y, int rem, int last) {
    uint8_t* seen = &s->seen[((y * s->W + x) * (s->maxv + 1) + rem) * 4 + last];
    if (*seen) return;
    *seen = 1;
    static const int DX[3] = {0, -1, 1};
    for (int d = 0; d < 3; ++d) {
        /* no immediate left<->right reversal inside a segment (test_bounce.py:175-186) */
        if ((last == 1 && d == 2) || (last == 2 && d == 1)) continue;
        const int nx = x + DX[d];
        const int ny = y + (d == 0 ? s->fwd : 0); /* never backwards (test_bounce.py:96-105) */
        if (nx < 0 || nx >= s->W || ny < 0 || ny >= s->H) continue;
        if (nx == s->bx && ny == s->by) continue;
        const int v = s->g[ny * s->W + nx];
        if (rem > 1) {
            /* intermediate cells must be empty (test_bounce.py:109-117) and the far goal row can
             * only be entered by the final step (test_bounce.py:109-117,282-294) */
            if (v != 0 || ny == s->far_row) continue;
            bounce_dfs(s, nx, ny, rem - 1, d);
        } else if (v > 0) {
            /* landing exactly on a piece: bounce with that piece's value, direction memory resets
             * (test_bounce.py:137-145,209-220) */
            bounce_dfs(s, nx, ny, v, 3);
        } else {
            s->out[ny * s->W + nx] = 1; /* final resting cell must be empty */
        }
    }
}

int bgso_bounce_targets(const int8_t* grid, int H, int W, int player, int sx, int sy, int rules,
                        uint8_t* target_map) {
    const int HW = H * W;
    memset(target_map, 0, (size_t)HW);
    if (sx < 0 || sx >= W || sy < 0 || sy >= H) return 0;
    const int v0 = grid[sy * W + sx];
    if (v0 <= 0) return 0;
    int maxv = 0;
    for (int i = 0; i < HW; ++i)
        if (grid[i] > maxv) maxv = grid[i];
    int8_t* g = (int8_t*)malloc((size_t)HW);
    memcpy(g, grid, (size_t)HW);
    bsearch_t s;
    s.g = g; s.H = H; s.W = W; s.maxv = maxv;
    s.fwd = player == 0 ? 1 : -1;
    s.far_row = player == 0 ? H - 1 : 0;
    s.bx = s.by = -1;
    const int variant = rules & 3;
    if (variant == BGSO_BOUNCE_SOURCE_EMPTY) g[sy * W + sx] = 0;
    else if (variant == BGSO_BOUNCE_SOURCE_BLOCKED) { g[sy * W + sx] = 0; s.bx = sx; s.by = sy; }
    s.seen = (uint8_t*)calloc((size_t)HW * (maxv + 1) * 4, 1);
    s.out = target_map;
    bounce_dfs(&s, sx, sy, v0, 3);
    if (!(rules & BGSO_BOUNCE_ALLOW_NULL_MOVE)) target_map[sy * W + sx] = 0;
    int cnt = 0;
    for (int i = 0; i < HW; ++i) cnt += target_map[i];
    free(s.seen);
    free(g);
    return cnt;
}

/* state.actions (bounce.cpp:40): every (source, target) of the mover's movable pieces, none when
 * ended (test_bounce.py:151).  Order (ours): ascending (sy, sx, ty, tx). */
int bgso_bounce_actions(const int8_t* grid, int H, int W, int player, int ended, int rules,
                        int32_t* moves, int cap) {
    if (ended) return 0;
    const int sy = bgso_bounce_source_row(grid, H, W, player);
    if (sy < 0) return 0;
    uint8_t* map = (uint8_t*)malloc((size_t)H * W);
    int n = 0;
    for (int sx = 0; sx < W; ++sx) {
        if (grid[sy * W + sx] <= 0) continue;
        bgso_bounce_targets(grid, H, W, player, sx, sy, rules, map);
        for (int ty = 0; ty < H; ++ty)
            for (int tx = 0; tx < W; ++tx)
                if (map[ty * W + tx]) {
                    if (moves && n < cap) {
                        moves[4 * n + 0] = sx; moves[4 * n + 1] = sy;
                        moves[4 * n + 2] = tx; moves[4 * n + 3] = ty;
                    }
                    ++n;
                }
    }
    free(map);
    return n;
}

/* action.sample_next_state (bounce.cpp:51): move the piece; far goal row wins (test_bounce.py:
 * 148-152,274-278); a blocked next player loses (:323-341) unless the mover would be blocked too,
 * which is a draw (:344-362); the side to move flips. */
int bgso_bounce_next(const int8_t* grid, int H, int W, int player, int ended, int rules, int sx,
                     int sy, int tx, int ty, int8_t* grid_out, int* player_out, int* winner_out,
                     int* ended_out) {
    if (ended) return -1;
    if (sx < 0 || sx >= W || sy < 0 || sy >= H || tx < 0 || tx >= W || ty < 0 || ty >= H) return -1;
    if (sy != bgso_bounce_source_row(grid, H, W, player) || grid[sy * W + sx] <= 0) return -1;
    uint8_t* map = (uint8_t*)malloc((size_t)H * W);
    bgso_bounce_targets(grid, H, W, player, sx, sy, rules, map);
    const int legal = map[ty * W + tx];
    free(map);
    if (!legal) return -1;
    if (grid_out != grid) memcpy(grid_out, grid, (size_t)H * W);
    const int8_t v = grid_out[sy * W + sx];
    grid_out[sy * W + sx] = 0;
    grid_out[ty * W + tx] = v;
    int win = -1, end = 0;
    const int far_row = player == 0 ? H - 1 : 0;
    if (ty == far_row) {
        win = player;
        end = 1;
    } else if (bgso_bounce_actions(grid_out, H, W, 1 - player, 0, rules, 0, 0) == 0) {
        end = 1;
        if (bgso_bounce_actions(grid_out, H, W, player, 0, rules, 0, 0) > 0) win = player;
    }
    *player_out = 1 - player;
    *winner_out = win;
    *ended_out = end;
    return 0;
}

/* shared body: per_game == 0 -> every game starts from grid0 with player 0; otherwise game i starts
 * from grid0 + i*H*W with player_in[i], winner_in[i] (optional) and ended_in[i] (optional). */
static int bounce_rollout_impl(const int8_t* grid0, int per_game, const int8_t* player_in,
                               const int8_t* winner_in, const uint8_t* ended_in, int H, int W, int rules,
                               int max_plies, uint64_t n, uint64_t gid0, uint64_t seed, uint8_t* moves,
                               uint16_t* length, int8_t* winner, int8_t* final_grid, float* reward,
                               int64_t* stats) {
    const int HW = H * W;
    if (HW > 255 || max_plies < 0 || max_plies > 65535) return -1;
    const int cap = W * HW;
    int8_t* g = (int8_t*)malloc((size_t)HW);
    int32_t* acts = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)cap);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t gid = gid0 + i;
        memcpy(g, per_game ? grid0 + i * HW : grid0, (size_t)HW);
        int player = per_game ? player_in[i] : 0;
        int win = (per_game && winner_in) ? winner_in[i] : -1;
        int end = win >= 0 || (per_game && ended_in && ended_in[i]);
        int t = 0;
        if (moves) memset(moves + i * 2 * (size_t)max_plies, 0xFF, 2 * (size_t)max_plies);
        /* a start position in which the side to move is already blocked has no actions: treat as
         * ended-with-no-winner (README.md:60 promises an action whenever has_ended is false) */
        while (!end && t < max_plies) {
            int na = bgso_bounce_actions(g, H, W, player, end, rules, acts, cap);
            if (na == 0) { end = 1; break; }
            uint32_t r = bgso_draw(seed, gid, (uint32_t)t, BGSO_DOMAIN_BOUNCE);
            const int32_t* a = acts + 4 * pick(r, (uint32_t)na);
            if (moves) {
                moves[(i * (size_t)max_plies + t) * 2 + 0] = (uint8_t)(a[1] * W + a[0]);
                moves[(i * (size_t)max_plies + t) * 2 + 1] = (uint8_t)(a[3] * W + a[2]);
            }
            bgso_bounce_next(g, H, W, player, end, rules, a[0], a[1], a[2], a[3], g, &player, &win, &end);
            ++t;
        }
        const int truncated = !end;
        if (length) length[i] = (uint16_t)t;
        if (winner) winner[i] = (int8_t)(truncated ? -2 : win);
        if (final_grid) memcpy(final_grid + i * HW, g, (size_t)HW);
        if (reward) bgso_reward(win, reward + 2 * i);
        stats_game(stats, win, t, truncated);
    }
    free(g);
    free(acts);
    return 0;
}

int bgso_bounce_rollout(const int8_t* grid0, int H, int W, int rules, int max_plies, uint64_t n,
                        uint64_t gid0, uint64_t seed, uint8_t* moves, uint16_t* length,
                        int8_t* winner, int8_t* final_grid, float* reward, int64_t* stats) {
    return bounce_rollout_impl(grid0, 0, 0, 0, 0, H, W, rules, max_plies, n, gid0, seed, moves, length, winner,
                               final_grid, reward, stats);
}

int bgso_bounce_rollout_from(const int8_t* grids, const int8_t* player, const int8_t* winner_in,
                             const uint8_t* ended_in, int H, int W, int rules, int max_plies, uint64_t n,
                             uint64_t gid0, uint64_t seed, uint8_t* moves, uint16_t* length, int8_t* winner,
                             int8_t* final_grid, float* reward, int64_t* stats) {
    return bounce_rollout_impl(grids, 1, player, winner_in, ended_in, H, W, rules, max_plies, n, gid0, seed, moves,
                               length, winner, final_grid, reward, stats);
}

int64_t bgso_bounce_replay(const int8_t* grid0, int H, int W, int rules, int max_plies, uint64_t n,
                           const uint8_t* moves, const uint16_t* length, const int8_t* winner,
                           const int8_t* final_grid, const float* reward, int64_t* first_bad) {
    const int HW = H * W;
    int8_t* g = (int8_t*)malloc((size_t)HW);
    int64_t bad = 0;
    if (first_bad) *first_bad = -1;
    for (uint64_t i = 0; i < n; ++i) {
        memcpy(g, grid0, (size_t)HW);
        int player = 0, win = -1, end = 0, ok = 1;
        const int len = length[i];
        if (len > max_plies) ok = 0;
        for (int t = 0; ok && t < len; ++t) {
            const uint8_t* m = moves + (i * (size_t)max_plies + t) * 2;
            if (player != (t & 1) || end) { ok = 0; break; }
            if (m[0] >= HW || m[1] >= HW) { ok = 0; break; }
            if (bgso_bounce_next(g, H, W, player, end, rules, m[0] % W, m[0] / W, m[1] % W, m[1] / W,
                                 g, &player, &win, &end))
                ok = 0;
        }
        int expect_w = win;
        if (ok && !end) {
            /* not ended after `len` plies: legal only as a truncation at the cap, or a start
             * position without any action */
            if (len == max_plies && bgso_bounce_actions(g, H, W, player, 0, rules, 0, 0) > 0) expect_w = -2;
            else if (!(len == 0 && bgso_bounce_actions(g, H, W, player, 0, rules, 0, 0) == 0)) ok = 0;
        }
        if (ok && winner && winner[i] != expect_w) ok = 0;
        if (ok && final_grid && memcmp(final_grid + i * HW, g, (size_t)HW)) ok = 0;
        if (ok && reward) {
            float rw[2];
            bgso_reward(win, rw);
            if (rw[0] != reward[2 * i] || rw[1] != reward[2 * i + 1]) ok = 0;
        }
        if (!ok) {
            if (first_bad && *first_bad < 0) *first_bad = (int64_t)i;
            ++bad;
        }
    }
    free(g);
    return bad;
}

/* ------------------------------------------------------------------------------------------------
 * Weighted action choice and state keys: conventions of the B200 build itself (include/bgs_b200.h,
 * bgs_connect_sample_step / bgs_bounce_sample_step / bgs_*_keys), restated independently so that the
 * tests can check the kernels.  The reference's counterpart of the first is the caller's
 * random.choices(actions, weights) (textual/examples/arena.py:64-68), of the second helper.hpp:10-25
 * (== / < / hash on every object; values unpinned).
 * ---------------------------------------------------------------------------------------------- */
#pragma STDC FP_CONTRACT OFF
static float sane_weight(float w) { return w > 0.0f ? (w < 3.402823466e+38f ? w : 3.402823466e+38f) : 0.0f; }
static uint32_t quantize_weight(float w, float wmax) {
    volatile float t = w / wmax;       /* volatile: separate IEEE single operations, no contraction */
    volatile float u = t * 65535.0f;
    volatile float v = u + 0.5f;
    return (uint32_t)v;
}

/* index (into the `count` candidate weights) chosen by draw r: first j with (q_0+..+q_j)*2^32 > r*sum q */
static int weighted_pick(const float* w, int count, uint32_t r) {
    float wmax = 0.0f;
    for (int j = 0; j < count; ++j) { const float s = sane_weight(w[j]); if (s > wmax) wmax = s; }
    uint64_t total = 0;
    for (int j = 0; j < count; ++j) total += wmax > 0.0f ? quantize_weight(sane_weight(w[j]), wmax) : 1u;
    const uint64_t thresh = (uint64_t)r * total;
    uint64_t cum = 0;
    for (int j = 0; j < count; ++j) {
        cum += wmax > 0.0f ? quantize_weight(sane_weight(w[j]), wmax) : 1u;
        if ((cum << 32) > thresh) return j;
    }
    return -1;
}

/* Column chosen for one state from probs[W] (weights of the columns; illegal ones ignored), draw index t;
 * -1 if the state has ended. */
int bgso_connect_sample(const int8_t* grid, int H, int W, int winner, const float* probs, uint64_t seed,
                        uint64_t gid, uint32_t t) {
    int32_t cols[64];
    float w[64];
    const int n = bgso_connect_actions(grid, H, W, winner, cols);
    if (n == 0) return -1;
    for (int j = 0; j < n; ++j) w[j] = probs[cols[j]];
    const int j = weighted_pick(w, n, bgso_draw(seed, gid, t, BGSO_DOMAIN_CONNECT));
    return j < 0 ? -1 : cols[j];
}

/* Move chosen for one Bounce state from probs[W][H*W]; returns 0 and (sx,sy,tx,ty) in move4, or -1. */
int bgso_bounce_sample(const int8_t* grid, int H, int W, int player, int ended, int rules, const float* probs,
                       uint64_t seed, uint64_t gid, uint32_t t, int32_t* move4) {
    const int cap = W * H * W;
    int32_t* acts = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)cap);
    float* w = (float*)malloc(sizeof(float) * (size_t)cap);
    const int n = bgso_bounce_actions(grid, H, W, player, ended, rules, acts, cap);
    int rc = -1;
    if (n > 0) {
        for (int j = 0; j < n; ++j) w[j] = probs[acts[4 * j] * (H * W) + acts[4 * j + 3] * W + acts[4 * j + 2]];
        const int j = weighted_pick(w, n, bgso_draw(seed, gid, t, BGSO_DOMAIN_BOUNCE));
        if (j >= 0) { for (int k = 0; k < 4; ++k) move4[k] = acts[4 * j + k]; rc = 0; }
    }
    free(acts);
    free(w);
    return rc;
}

static uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

/* key[2] of one state; game 1 = Connect (exact packing when H*W <= 62), 2 = Bounce (hash). */
void bgso_state_key(int game, const int8_t* grid, int H, int W, int player, int winner, uint64_t* key) {
    const int HW = H * W;
    if (game == 1 && HW <= 62) {
        unsigned __int128 k = 0;
        for (int c = 0; c < HW; ++c) {
            if (grid[c] == 0) k |= (unsigned __int128)1 << c;
            else if (grid[c] == 1) k |= (unsigned __int128)1 << (HW + c);
        }
        k |= (unsigned __int128)(player & 1) << (2 * HW);
        k |= (unsigned __int128)((winner + 1) & 3) << (2 * HW + 1);
        key[0] = (uint64_t)k; key[1] = (uint64_t)(k >> 64);
        return;
    }
    uint64_t h0 = 0x9E3779B97F4A7C15ull, h1 = 0xC2B2AE3D27D4EB4Full;
    for (int c0 = 0; c0 < HW; c0 += 8) {
        uint64_t w = 0;
        for (int j = 0; j < 8 && c0 + j < HW; ++j) w |= (uint64_t)(uint8_t)grid[c0 + j] << (8 * j);
        h0 = mix64(h0 ^ w);
        h1 = mix64(h1 + w + 0x632BE59BD9B4E019ull);
    }
    const uint64_t tail = (uint64_t)(uint8_t)player | ((uint64_t)(uint8_t)winner << 8) | ((uint64_t)H << 16) |
                          ((uint64_t)W << 24) | ((uint64_t)game << 32);
    key[0] = mix64(h0 ^ tail);
    key[1] = mix64(h1 + tail + 0x632BE59BD9B4E019ull);
}
