/*
 * fast_connect.c -- a tight single-word BITBOARD rollout loop for Connect-k on the host CPU.
 * TEST / BASELINE INFRASTRUCTURE ONLY (same rules as bgs_oracle.h: only tests/ and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product never does).
 *
 * bgs_oracle.c is deliberately brute force (whole-board scan per move), which makes it a slow yardstick.
 * This file is the "best CPU" figure of SURVEY.md 8d (iii): the same loop -- legal columns, uniform choice
 * with the SAME Philox draws, drop, k-in-a-row, terminal -- written the way a CPU engine would: column-major
 * bitboards with a sentinel row (bit = col*(H+1) + row), shift-and-AND run test, -O3.  It is checked against
 * bgs_oracle.c game by game (tests/test_oracle_golden.py), so it restates the same path
 * (README.md:49-72 of the reference; connect.cpp:32-52).  Boards with (H+1)*W <= 64.
 */
#include <stdint.h>
#include <string.h>

#include "bgs_oracle.h"

int bgso_fast_connect_supported(int H, int W, int K) { return H >= 1 && W >= 1 && K >= 1 && (H + 1) * W <= 64 && W <= 16; }

static inline int has_run(uint64_t b, int H1, int K) {
    const int dirs[4] = {1, H1, H1 - 1, H1 + 1};  /* vertical, horizontal, two diagonals */
    for (int i = 0; i < 4; ++i) {
        uint64_t m = b;
        int len = 1;
        while (2 * len <= K) { m &= m >> (len * dirs[i]); len *= 2; }
        if (len < K) m &= m >> ((K - len) * dirs[i]);
        if (m) return 1;
    }
    return 0;
}

/* length / winner optional; stats accumulated as bgso_connect_rollout does. Returns 0, or -1 if unsupported. */
int bgso_fast_connect_rollout(int H, int W, int K, uint64_t n, uint64_t gid0, uint64_t seed, uint8_t* length,
                              int8_t* winner, int64_t* stats) {
    if (!bgso_fast_connect_supported(H, W, K)) return -1;
    const int H1 = H + 1, HW = H * W;
    int64_t games = 0, w0 = 0, w1 = 0, dr = 0, steps = 0;
    int64_t hist[BGSO_STATS_LEN - BGSO_STAT_HIST0];
    memset(hist, 0, sizeof(hist));
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t gid = gid0 + i;
        uint64_t b[2] = {0, 0};
        uint8_t h[16] = {0};
        uint8_t cols[16];
        int nleg = W, t = 0, win = -1;
        for (int c = 0; c < W; ++c) cols[c] = (uint8_t)c;
        uint32_t r[4];
        while (nleg > 0) {
            if ((t & 3) == 0) {
                const uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t >> 2, BGSO_DOMAIN_CONNECT};
                bgso_philox4x32_10(ctr, key, r);
            }
            const int k = (int)(((uint64_t)r[t & 3] * (uint32_t)nleg) >> 32);
            const int c = cols[k];
            const int p = t & 1;
            b[p] |= 1ull << (c * H1 + h[c]);
            ++t;
            if (++h[c] == H) {  /* column full: remove it from the ascending list */
                for (int j = k; j + 1 < nleg; ++j) cols[j] = cols[j + 1];
                --nleg;
            }
            if (has_run(b[p], H1, K)) { win = p; break; }
        }
        if (length) length[i] = (uint8_t)t;
        if (winner) winner[i] = (int8_t)win;
        ++games; steps += t;
        if (win == 0) ++w0; else if (win == 1) ++w1; else ++dr;
        ++hist[t < BGSO_STATS_LEN - BGSO_STAT_HIST0 - 1 ? t : BGSO_STATS_LEN - BGSO_STAT_HIST0 - 1];
        (void)HW;
    }
    if (stats) {
        stats[BGSO_STAT_GAMES] += games; stats[BGSO_STAT_WIN0] += w0; stats[BGSO_STAT_WIN1] += w1;
        stats[BGSO_STAT_DRAWS] += dr; stats[BGSO_STAT_STEPS] += steps;
        for (int j = 0; j < BGSO_STATS_LEN - BGSO_STAT_HIST0; ++j) stats[BGSO_STAT_HIST0 + j] += hist[j];
    }
    return 0;
}
