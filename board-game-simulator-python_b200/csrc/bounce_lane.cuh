// bounce_lane.cuh -- the per-lane state machine of the Bounce rollout kernel (one game per lane).
//
// Everything a lane does between claiming a game and writing its result lives here, as
// __host__ __device__ code with no warp intrinsics, so the same source is (a) inlined into
// bounce_rollout_kernel (bounce.cu) and (b) compiled by g++ into a host harness
// (tests/native/bounce_lane_host.cpp) that plays games lane by lane on the CPU and is compared with
// the oracle without a GPU.  The product never runs the host build.
//
// Replaces, inside the rollout loop, the reference's State::get_actions + the caller's uniform choice
// + Action::sample_next_state + has_ended / reward (src/simulator/game/bounce.cpp:36-51,
// README.md:52-69).  Rules: SURVEY.md 4.4, pinned by tests/test_bounce.py.
//
// Orientation.  The planes are kept MOVER-RELATIVE: for player 1 the board is rotated by 180 degrees
// (cell c -> H*W-1-c, one BREV per word), which maps "towards row 0" to "towards row H-1" and swaps
// left / right -- the move set {forward, left, right} is symmetric under that swap, so one code
// path with compile-time shift directions (forward = << W) serves both players.  The canonical
// action order (ascending source, then target cell) is exactly reversed by the rotation: the k-th
// action of player 1 is the (n-1-k)-th in relative order.
//
// Move generation = bit-parallel reachability.  One "segment" = the u steps a piece of value u
// travels; all frontier cells advance at once, three masks keyed by the last direction (no immediate
// reversal).  A segment that lands on pieces seeds further segments ("bounces"), grouped by piece
// value.  The first segment of a piece is treated as a bounce off the piece itself (pending = its
// cell), so every iteration of the lane's loop has the same shape:
//     [piece boundary, short, only when no bounce is pending]  ->  [segment setup]  ->  [u steps].
#pragma once

#include <stdint.h>

#include "../../include/bgs_b200.h"

#if defined(__CUDACC__)
#define BGS_HD __host__ __device__ __forceinline__
#else
#define BGS_HD inline
#endif

namespace bgs {
namespace bounce {

BGS_HD int popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
BGS_HD int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
BGS_HD int ctz64(uint64_t x) {  // x != 0
#ifdef __CUDA_ARCH__
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
BGS_HD int ctz32(uint32_t x) {  // x != 0
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
BGS_HD uint64_t brev64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    uint64_t r = 0;
    for (int i = 0; i < 64; ++i) r |= ((x >> i) & 1ull) << (63 - i);
    return r;
#endif
}
BGS_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Philox4x32-10, counter (c0..c3), key (k0, k1) -- same function as bgs_common.cuh's device copy.
BGS_HD void philox_hd(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                      uint32_t (&out)[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---- geometry: run-time (any board) or compile-time (the BASELINE default 9x6) -------------------
struct GeoRT {
    int H, W, HW, rules;
    uint64_t board;      // low H*W bits
    uint64_t not_left;   // cells with x > 0
    uint64_t not_right;  // cells with x < W-1
    uint64_t far;        // the mover's far goal row in mover-relative orientation = row H-1
    uint32_t row0;       // (1 << W) - 1
    uint32_t inv_w;      // ceil(2^16 / W): cell / W == (cell * inv_w) >> 16 for cell < 64, W <= 8
    BGS_HD int h() const { return H; }
    BGS_HD int w() const { return W; }
    BGS_HD int hw() const { return HW; }
    BGS_HD uint64_t m_board() const { return board; }
    BGS_HD uint64_t m_not_left() const { return not_left; }
    BGS_HD uint64_t m_not_right() const { return not_right; }
    BGS_HD uint64_t m_far() const { return far; }
    BGS_HD uint32_t m_row0() const { return row0; }
    BGS_HD int row_of(int cell) const { return (int)(((uint32_t)cell * inv_w) >> 16); }
};

inline GeoRT make_geo_rt(int H, int W, int rules) {
    GeoRT g;
    g.H = H; g.W = W; g.HW = H * W; g.rules = rules;
    g.board = (H * W == 64) ? ~0ull : ((1ull << (H * W)) - 1ull);
    g.not_left = 0; g.not_right = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (x > 0) g.not_left |= 1ull << (y * W + x);
            if (x < W - 1) g.not_right |= 1ull << (y * W + x);
        }
    g.row0 = (uint32_t)((1ull << W) - 1ull);
    g.far = (uint64_t)g.row0 << ((H - 1) * W);
    g.inv_w = (65536u + (uint32_t)W - 1u) / (uint32_t)W;
    return g;
}

template <int H_, int W_>
struct GeoCT {
    static_assert(H_ * W_ <= 64 && W_ <= 8 && W_ >= 1 && H_ >= 1, "board must fit one 64-bit word");
    static constexpr uint64_t kBoard = (H_ * W_ == 64) ? ~0ull : ((1ull << (H_ * W_ % 64)) - 1ull);
    static constexpr uint64_t col_mask(int x0, int x1) {
        uint64_t m = 0;
        for (int y = 0; y < H_; ++y)
            for (int x = x0; x < x1; ++x) m |= 1ull << (y * W_ + x);
        return m;
    }
    static constexpr uint64_t kNotLeft = col_mask(1, W_);
    static constexpr uint64_t kNotRight = col_mask(0, W_ - 1);
    static constexpr uint32_t kRow0 = (uint32_t)((1ull << W_) - 1ull);
    static constexpr uint64_t kFar = (uint64_t)kRow0 << ((H_ - 1) * W_);
    int rules;
    BGS_HD explicit GeoCT(const GeoRT& g) : rules(g.rules) {}
    BGS_HD int h() const { return H_; }
    BGS_HD int w() const { return W_; }
    BGS_HD int hw() const { return H_ * W_; }
    BGS_HD uint64_t m_board() const { return kBoard; }
    BGS_HD uint64_t m_not_left() const { return kNotLeft; }
    BGS_HD uint64_t m_not_right() const { return kNotRight; }
    BGS_HD uint64_t m_far() const { return kFar; }
    BGS_HD uint32_t m_row0() const { return kRow0; }
    BGS_HD int row_of(int cell) const { return cell / W_; }
};

// index of the k-th (0-based) set bit of m (k < popcount(m))
BGS_HD int kth_set_bit64(uint64_t m, int k) {
    uint32_t w = (uint32_t)m;
    int pos = 0;
    const int c = popc32(w);
    if (k >= c) {
        k -= c;
        w = (uint32_t)(m >> 32);
        pos = 32;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t low = w & ((1u << s) - 1u);
        const int cl = popc32(low);
        if (k >= cl) {
            k -= cl;
            w >>= s;
            pos += s;
        } else {
            w = low;
        }
    }
    return pos;
}

// Outputs of one launch (any may be null) -- the per-game part of RolloutParams.
struct LaneOut {
    uint8_t* moves;      // [n, max_plies, 2] pre-filled 0xFF
    uint16_t* length;
    int8_t* winner;
    int8_t* final_grid;  // [n, H*W]
    float* reward;       // [n, 2]
};

// RULES_ >= 0: compile-time rule set; -1: read g.rules.
template <int NP, class G, int RULES_>
struct Lane {
    // ---- game --------------------------------------------------------------------------------
    uint64_t b[NP];  // value bit-planes, oriented for `orient`
    int t;           // plies played in this rollout
    int player;      // side to move
    int orient;      // whose orientation b[] is in
    int win;
    uint32_t r[4];   // the Philox block of plies 4*(t>>2) .. +3
    // ---- move generation ---------------------------------------------------------------------
    uint64_t occ, src_left, sbit, occS, inter, open, expanded, pending, targets;
    int total, nsrc;
    bool probe, found, have, waiting;

    BGS_HD int rules(const G& g) const { return RULES_ >= 0 ? RULES_ : g.rules; }
    BGS_HD uint64_t rot(const G& g, uint64_t x) const { return brev64(x) >> (64 - g.hw()); }

    // Start (or restart, for the blocked test) a move generation for `pl`.
    BGS_HD void begin_movegen(const G& g, int pl, bool prb) {
        if (orient != pl) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < NP; ++i) b[i] = rot(g, b[i]);
            orient = pl;
        }
        uint64_t o = b[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 1; i < NP; ++i) o |= b[i];
        occ = o;
        src_left = 0;
        if (o) {  // movable pieces: the occupied row nearest to the mover (tests/test_bounce.py:43-48,60)
            const int row = g.row_of(ctz64(o));
            src_left = o & ((uint64_t)g.m_row0() << (row * g.w()));
        }
        probe = prb; found = false; have = false; waiting = false;
        pending = 0; total = 0; nsrc = 0;
    }

    BGS_HD void begin_game_planes(const G& g, const uint64_t* plane0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < NP; ++i) b[i] = plane0[i];
        t = 0; player = 0; orient = 0; win = BGS_WINNER_DRAW;
        begin_movegen(g, 0, false);
    }

    // Per-game start position in the reference's layout (int8 grid, row 0 = bottom).
    BGS_HD void begin_game_grid(const G& g, const int8_t* grid, int pl, int winner_in, bool ended) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < NP; ++i) b[i] = 0;
        for (int c = 0; c < g.hw(); ++c) {
            const int v = grid[c];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < NP; ++i) b[i] |= (uint64_t)((v >> i) & 1) << c;
        }
        t = 0; player = pl & 1; orient = 0; win = winner_in;
        begin_movegen(g, player, false);
        if (ended || winner_in >= 0) src_left = 0;  // no move generation: ends at t == 0 with total == 0
    }

    // One iteration of the move generation: at most one piece boundary, then one whole segment.
    // T[j * stride] receives the target mask (mover-relative) of the j-th movable piece.
    BGS_HD void movegen_iter(const G& g, uint64_t* T, int stride) {
        const int rl = rules(g);
        if (pending == 0) {  // piece boundary
            if (have) {
                const uint64_t tg = (rl & BGS_BOUNCE_ALLOW_NULL_MOVE) ? targets : (targets & ~sbit);
                if (probe) {
                    found = tg != 0;
                } else {
                    T[nsrc * stride] = tg;
                    total += popc64(tg);
                    ++nsrc;
                }
                have = false;
            }
            if (src_left == 0 || found) {
                waiting = true;
                return;
            }
            sbit = src_left & (~src_left + 1ull);  // next movable piece, ascending relative column
            src_left ^= sbit;
            have = true;
            const int variant = rl & 3;
            occS = variant == BGS_BOUNCE_SOURCE_PIECE ? occ : (occ & ~sbit);
            open = variant == BGS_BOUNCE_SOURCE_BLOCKED ? (g.m_board() & ~sbit) : g.m_board();
            inter = open & ~occS & ~g.m_far();  // cells a path may pass through
            expanded = 0;
            targets = 0;
            pending = sbit;  // the first segment = a "bounce" off the piece itself
        }
        // ---- segment setup: all unexpanded landing cells holding a piece of the same value as the
        // lowest one travel together
        const uint64_t low = pending & (~pending + 1ull);
        uint64_t S = pending;
        int u = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < NP; ++i) {
            const bool in = (b[i] & low) != 0;
            S &= in ? b[i] : ~b[i];
            u |= in ? (1 << i) : 0;
        }
        pending ^= S;
        expanded |= S;
        // ---- u steps, every frontier cell at once; x* = cells entered by a forward / left / right
        // step (a left step may not follow a right step and vice versa; never backwards)
        const int W = g.w();
        uint64_t xf = S << W;
        uint64_t xl = (S & g.m_not_left()) >> 1;
        uint64_t xr = (S & g.m_not_right()) << 1;
        // intermediate cells: empty, not the far goal row
        auto advance = [&]() {
            const uint64_t af = (xf | xl | xr) & inter;
            const uint64_t al = (xf | xl) & inter & g.m_not_left();
            const uint64_t ar = (xf | xr) & inter & g.m_not_right();
            xf = af << W;
            xl = al >> 1;
            xr = ar << 1;
            return (af | al | ar) != 0;
        };
        if (NP == 2) {  // values <= 3: at most two intermediate steps, no loop
            if (u >= 2) advance();
            if (u == 3) advance();
        } else {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int i = 1; i < u; ++i)
                if (!advance()) break;
        }
        // last step: rest on an empty cell, or bounce off a piece
        const uint64_t land = (xf | xl | xr) & open;
        targets |= land & ~occS;
        pending |= land & occS & ~expanded;
    }

    // The ply transition of a lane whose move generation is complete (waiting == true).
    // Returns true when the game is over (win / t final; the caller writes the result).
    BGS_HD bool transition(const G& g, const uint64_t* T, int stride, uint64_t gid, uint32_t seed_lo,
                           uint32_t seed_hi, int max_plies, uint8_t* moves_row) {
        waiting = false;
        if (probe) {  // `player` is blocked; the previous mover wins unless blocked too (draw)
            win = found ? 1 - player : BGS_WINNER_DRAW;
            return true;
        }
        if (total == 0) {
            if (t == 0) return true;  // a blocked / ended start position: no winner
            begin_movegen(g, 1 - player, true);
            return false;
        }
        if (t >= max_plies) {
            win = BGS_WINNER_TRUNCATED;
            return true;
        }
        if ((t & 3) == 0) philox_hd((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t >> 2, 1u, seed_lo, seed_hi, r);
        const uint32_t rr = (t & 3) == 0 ? r[0] : ((t & 3) == 1 ? r[1] : ((t & 3) == 2 ? r[2] : r[3]));
        int k = (int)mulhi32(rr, (uint32_t)total);
        if (player) k = total - 1 - k;  // canonical (absolute) order is the reverse of the rotated one
        // k-th action in ascending relative (source, target) order
        const int base = g.row_of(ctz64(occ)) * g.w();
        uint32_t sm = (uint32_t)(occ >> base) & g.m_row0();
        uint64_t tm = T[0];
        int j = 0;
        for (;;) {
            const int c = popc64(tm);
            if (k < c) break;
            k -= c;
            ++j;
            sm &= sm - 1u;
            tm = T[j * stride];
        }
        const int scell = base + ctz32(sm);
        const int tcell = kth_set_bit64(tm, k);
        if (moves_row) {
            const int HW1 = g.hw() - 1;
            const int sa = player ? HW1 - scell : scell, ta = player ? HW1 - tcell : tcell;
            moves_row[2 * t] = (uint8_t)sa;
            moves_row[2 * t + 1] = (uint8_t)ta;
        }
        const uint64_t smask = 1ull << scell, tmask = 1ull << tcell;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < NP; ++i) {
            const bool has = (b[i] & smask) != 0;
            b[i] = (b[i] & ~smask) | (has ? tmask : 0ull);
        }
        ++t;
        const bool goal = (tmask & g.m_far()) != 0;
        if (goal) win = player;
        player ^= 1;
        if (goal) return true;
        begin_movegen(g, player, false);
        return false;
    }

    BGS_HD int value_abs(const G& g, int cell) const {
        const int c = orient ? g.hw() - 1 - cell : cell;
        int v = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < NP; ++i) v |= (int)((b[i] >> c) & 1ull) << i;
        return v;
    }

    // Result of a finished game (lengths / winners / final grid / reward; statistics are the caller's).
    BGS_HD void write_result(const G& g, const LaneOut& o, size_t idx) const {
        if (o.length) o.length[idx] = (uint16_t)t;
        if (o.winner) o.winner[idx] = (int8_t)win;
        if (o.final_grid) {
            int8_t* out = o.final_grid + idx * (size_t)g.hw();
            for (int c = 0; c < g.hw(); ++c) out[c] = (int8_t)value_abs(g, c);
        }
        if (o.reward) {
            o.reward[2 * idx] = win == 0 ? 1.f : (win == 1 ? -1.f : 0.f);
            o.reward[2 * idx + 1] = win == 1 ? 1.f : (win == 0 ? -1.f : 0.f);
        }
    }
};

}  // namespace bounce
}  // namespace bgs
