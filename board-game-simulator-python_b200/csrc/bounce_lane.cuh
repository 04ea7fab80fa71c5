// bounce_lane.cuh -- the per-lane state machine of the Bounce rollout kernel (one game per lane).
//
// Everything a lane does between claiming a game and writing its result lives here, as
// __host__ __device__ code with no warp intrinsics, so the same source is (a) inlined into
// bounce_rollout_kernel (bounce.cu) and (b) compiled by g++ into a host harness
// (tests/native/bounce_lane_host.cpp) that plays games lane by lane on the CPU and is checked by
// the CPU test suite without a GPU.  The product never runs the host build.
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
// index of the only set bit of x
BGS_HD int bit_index64(uint64_t x) {
#ifdef __CUDA_ARCH__
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t pos;  // bfind = FLO: the bit's position in whichever half holds it, no select between the halves
    asm("bfind.u32 %0, %1;" : "=r"(pos) : "r"(lo | hi));
    return (int)(hi ? pos + 32u : pos);
#else
    return __builtin_ctzll(x);
#endif
}
// index of the highest set bit of x (x != 0), minus 3: FLO finds the highest bit directly (no isolation of the lowest one)
BGS_HD int top_bit_minus3(uint64_t x) {
#ifdef __CUDA_ARCH__
    return 60 - __clzll((long long)x);
#else
    return 60 - __builtin_clzll(x);
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

// ---- board words: one 64-bit word (H*W <= 64, W <= 8) or an unsigned __int128 (H*W <= 128, W <= 16) ----
typedef unsigned __int128 u128;
BGS_HD int popcb(uint64_t x) { return popc64(x); }
BGS_HD int popcb(u128 x) { return popc64((uint64_t)x) + popc64((uint64_t)(x >> 64)); }
BGS_HD int ctzb(uint64_t x) { return ctz64(x); }
BGS_HD int ctzb(u128 x) {
    const uint64_t lo = (uint64_t)x;
    return lo ? ctz64(lo) : 64 + ctz64((uint64_t)(x >> 64));
}
BGS_HD uint64_t revb(uint64_t x) { return brev64(x); }
BGS_HD u128 revb(u128 x) { return ((u128)brev64((uint64_t)x) << 64) | (u128)brev64((uint64_t)(x >> 64)); }

template <class B> BGS_HD B make_bits(uint64_t lo, uint64_t hi);
template <> BGS_HD uint64_t make_bits<uint64_t>(uint64_t lo, uint64_t) { return lo; }
template <> BGS_HD u128 make_bits<u128>(uint64_t lo, uint64_t hi) { return ((u128)hi << 64) | lo; }

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

// ---- segment-table hash ---------------------------------------------------------------------------
// The table-driven segment (MoveGen::lut_segment) indexes its table with 8 bits of the 32-bit window around the
// pending cell: window bits {1, 2, 4, 5, S+2, S+3, S+4, 2S+3}.  Gathering them costs 11 instructions; for the row
// strides S = 4..9 a multiplier M exists (random search, checked by seg_hash_is_perfect) such that
// (window & mask) * M >> 24 maps the 256 subsets of those bits onto 256 different values -- 3 instructions -- and
// the kernels build their tables THROUGH that map (seg_lut_slot).  S = 3 (two columns) has no 8 distinct bits.
BGS_HD constexpr uint32_t seg_hash_mask(int S) {
    return (1u << 1) | (1u << 2) | (1u << 4) | (1u << 5) | (1u << (S + 2)) | (1u << (S + 3)) | (1u << (S + 4)) | (1u << (2 * S + 3));
}
BGS_HD constexpr uint32_t seg_hash_mul(int S) {
    return S == 4 ? 0x0C808055u : S == 5 ? 0x68401173u : S == 6 ? 0x4420061Du : S == 7 ? 0x9410021Fu
         : S == 8 ? 0x1C00C409u : S == 9 ? 0xB4012103u : 0u;  // 0: no hash, gather the bits
}

// ---- geometry: run-time (any board) or compile-time (the BASELINE default 9x6) -------------------
// Cell (x, y) is bit y*S + x.  S = W + 1 whenever H*(W+1) <= 64: the spare GUARD column (never part
// of `board`) absorbs horizontal steps off the left / right edge, so that the frontier steps need no
// edge masks (not_left / not_right are all-ones then); S = W otherwise.
template <class B>
struct GeoRTb {
    typedef B bits;
    static constexpr int BITS = (int)sizeof(B) * 8;
    static constexpr bool LUT = false;  // no compile-time guarantee; lut_ok decides at run time
    static constexpr bool HASH = false; // no compile-time hash; hash_mul decides at run time (0 = gather the 8 bits)
    static constexpr int MAX_SOURCES = sizeof(B) == 8 ? 8 : 16;  // columns of one row
    bool lut_ok;         // table-driven segments possible: guard column, landing window within 32 bits
    uint32_t hash_mask, hash_mul;  // segment-table hash of this row stride (seg_hash_mul), 0 = none
    BGS_HD uint32_t hmask() const { return hash_mask; }
    BGS_HD uint32_t hmul() const { return hash_mul; }
    int H, W, S, rules;
    int rot_shift;       // BITS-1 - (index of the last cell): rot180(x) = bit-reverse(x) >> rot_shift
    B board;             // the H*W valid cells
    B not_left;          // cells with x > 0      (all-ones with a guard column)
    B not_right;         // cells with x < W-1    (all-ones with a guard column)
    B far;               // the mover's far goal row in mover-relative orientation = row H-1
    uint32_t row0;       // (1 << W) - 1
    uint32_t inv_s;      // ceil(2^16 / S): cell / S == (cell * inv_s) >> 16 for cell < 128, S <= 17
    BGS_HD int h() const { return H; }
    BGS_HD int w() const { return W; }
    BGS_HD int s() const { return S; }
    BGS_HD int rot_sh() const { return rot_shift; }
    BGS_HD B m_board() const { return board; }
    BGS_HD B m_not_left() const { return not_left; }
    BGS_HD B m_not_right() const { return not_right; }
    BGS_HD B m_far() const { return far; }
    BGS_HD uint32_t m_row0() const { return row0; }
    BGS_HD int row_of(int cell) const { return (int)(((uint32_t)cell * inv_s) >> 16); }
    BGS_HD bool lut_rt() const { return lut_ok; }
};
typedef GeoRTb<uint64_t> GeoRT;
typedef GeoRTb<u128> GeoRT128;

template <class B>
inline GeoRTb<B> make_geo_rt_b(int H, int W, int rules, bool guard = true) {
    GeoRTb<B> g;
    g.H = H; g.W = W; g.rules = rules;
    g.S = (guard && H * (W + 1) <= GeoRTb<B>::BITS) ? W + 1 : W;
    g.rot_shift = GeoRTb<B>::BITS - 1 - ((H - 1) * g.S + W - 1);
    g.board = 0; g.not_left = 0; g.not_right = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            g.board |= (B)1 << (y * g.S + x);
            if (x > 0) g.not_left |= (B)1 << (y * g.S + x);
            if (x < W - 1) g.not_right |= (B)1 << (y * g.S + x);
        }
    if (g.S > W) g.not_left = g.not_right = ~(B)0;
    g.row0 = (uint32_t)((1ull << W) - 1ull);
    g.far = (B)g.row0 << ((H - 1) * g.S);
    g.inv_s = (65536u + (uint32_t)g.S - 1u) / (uint32_t)g.S;
    // table-driven segments (seg_lut_entry): 64-bit words, a guard column, the landing window of a 3-step
    // segment within 32 bits.  The caller clears the flag when the goal rows hold pieces or values exceed 3.
    // (S >= 3: the lowest piece cell is S, and the window starts 3 cells below the piece.)
    g.lut_ok = GeoRTb<B>::BITS == 64 && g.S > W && g.S >= 3 && 3 * g.S + 3 <= 31;
    g.hash_mul = g.lut_ok ? seg_hash_mul(g.S) : 0u;
    g.hash_mask = g.hash_mul ? seg_hash_mask(g.S) : 0u;
    return g;
}
inline GeoRT make_geo_rt(int H, int W, int rules, bool guard = true) { return make_geo_rt_b<uint64_t>(H, W, rules, guard); }

// Bit-planes of a reference-layout grid (int8[H*W], row 0 = bottom) in the layout of g.
template <class B>
inline void planes_from_grid(const GeoRTb<B>& g, const int8_t* grid, B plane[4]) {
    for (int i = 0; i < 4; ++i) plane[i] = 0;
    for (int y = 0; y < g.H; ++y)
        for (int x = 0; x < g.W; ++x)
            for (int i = 0; i < 4; ++i) plane[i] |= (B)((grid[y * g.W + x] >> i) & 1) << (y * g.S + x);
}

template <int H_, int W_>
struct GeoCT {
    static_assert(H_ * W_ <= 64 && W_ <= 8 && W_ >= 1 && H_ >= 1, "board must fit one 64-bit word");
    typedef uint64_t bits;
    static constexpr int S_ = (H_ * (W_ + 1) <= 64) ? W_ + 1 : W_;
    static constexpr int MAX_SOURCES = W_;
    // segments by table look-up (seg_lut_entry): needs the guard column and a 32-bit landing window
    static constexpr bool LUT = S_ > W_ && S_ >= 3 && 3 * S_ + 3 <= 31;
    // the table is indexed by a multiplicative hash of the window instead of the 8 gathered bits (seg_hash_mul)
    static constexpr bool HASH = LUT && seg_hash_mul(S_) != 0u;
    static constexpr uint32_t kHashMask = seg_hash_mask(S_);
    static constexpr uint32_t kHashMul = seg_hash_mul(S_);
    BGS_HD uint32_t hmask() const { return kHashMask; }
    BGS_HD uint32_t hmul() const { return kHashMul; }
    static constexpr uint64_t col_mask(int x0, int x1) {
        uint64_t m = 0;
        for (int y = 0; y < H_; ++y)
            for (int x = x0; x < x1; ++x) m |= 1ull << (y * S_ + x);
        return m;
    }
    static constexpr uint64_t kBoard = col_mask(0, W_);
    static constexpr uint64_t kNotLeft = S_ > W_ ? ~0ull : col_mask(1, W_);
    static constexpr uint64_t kNotRight = S_ > W_ ? ~0ull : col_mask(0, W_ - 1);
    static constexpr uint32_t kRow0 = (uint32_t)((1ull << W_) - 1ull);
    static constexpr uint64_t kFar = (uint64_t)kRow0 << ((H_ - 1) * S_);
    int rules;
    BGS_HD explicit GeoCT(const GeoRT& g) : rules(g.rules) {}
    BGS_HD int h() const { return H_; }
    BGS_HD int w() const { return W_; }
    BGS_HD int s() const { return S_; }
    BGS_HD int rot_sh() const { return 63 - ((H_ - 1) * S_ + W_ - 1); }
    BGS_HD uint64_t m_board() const { return kBoard; }
    BGS_HD uint64_t m_not_left() const { return kNotLeft; }
    BGS_HD uint64_t m_not_right() const { return kNotRight; }
    BGS_HD uint64_t m_far() const { return kFar; }
    BGS_HD uint32_t m_row0() const { return kRow0; }
    BGS_HD int row_of(int cell) const { return cell / S_; }
    BGS_HD bool lut_rt() const { return LUT; }
};

// public cell index y*W + x of internal cell y*S + x
template <class G>
BGS_HD int pub_cell(const G& g, int cell) { return cell - g.row_of(cell) * (g.s() - g.w()); }

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

BGS_HD int kth_set_bit(uint64_t m, int k) { return kth_set_bit64(m, k); }
BGS_HD int kth_set_bit(u128 m, int k) {
    const uint64_t lo = (uint64_t)m;
    const int c = popc64(lo);
    return k < c ? kth_set_bit64(lo, k) : 64 + kth_set_bit64((uint64_t)(m >> 64), k - c);
}

// Outputs of one launch (any may be null) -- the per-game part of RolloutParams.
struct LaneOut {
    uint8_t* moves;      // [n, max_plies, 2] pre-filled 0xFF
    uint16_t* length;
    int8_t* winner;
    int8_t* final_grid;  // [n, H*W]
    float* reward;       // [n, 2]
};

// ---------------------------------------------------------------------------------------------
// Segment look-up table for small piece values (<= 3) on a layout with a guard column.
// The cells a segment of u <= 3 steps from cell c can LAND on depend only on which of the 8 cells at
// distance 1 and 2 ahead / beside c may be passed through: offsets {-2, -1, +1, +2, S-1, S, S+1, 2S}
// <-> bits 0..7 of idx (1 = passable).  The entry is the landing set as a window mask, bit 3 = cell c
// (offsets -3 .. 3S -> bits 0 .. 3S+3).  Same rules as the frontier propagation in MoveGen::iter:
// forward / left / right, no left<->right reversal inside a segment, never backwards; cells off the
// left / right edge are guard cells (never passable, never a landing cell once masked with the board).
// ---------------------------------------------------------------------------------------------
constexpr int SEG_LUT_WORDS = 4 * 256;  // [idx][u], u = 0 unused (seg_lut_slot)
BGS_HD uint32_t seg_lut_entry(int S, int u, uint32_t idx) {
    if (u < 1 || u > 3) return 0u;
    const int off[8] = {-2, -1, 1, 2, S - 1, S, S + 1, 2 * S};
    const int dir[3] = {S, -1, 1};  // forward, left, right
    uint32_t mask = 0;
    for (int d1 = 0; d1 < 3; ++d1) {
        const int p1 = dir[d1];
        if (u == 1) { mask |= 1u << (3 + p1); continue; }
        bool ok1 = false;
        for (int k = 0; k < 8; ++k) ok1 = ok1 || (off[k] == p1 && ((idx >> k) & 1u));
        if (!ok1) continue;
        for (int d2 = 0; d2 < 3; ++d2) {
            if ((d1 == 1 && d2 == 2) || (d1 == 2 && d2 == 1)) continue;
            const int p2 = p1 + dir[d2];
            if (u == 2) { mask |= 1u << (3 + p2); continue; }
            bool ok2 = false;
            for (int k = 0; k < 8; ++k) ok2 = ok2 || (off[k] == p2 && ((idx >> k) & 1u));
            if (!ok2) continue;
            for (int d3 = 0; d3 < 3; ++d3) {
                if ((d2 == 1 && d3 == 2) || (d2 == 2 && d3 == 1)) continue;
                mask |= 1u << (3 + p2 + dir[d3]);
            }
        }
    }
    return mask;
}

// Position of entry (u, idx) in the table a kernel builds: [idx][u] in general, [hash of the window bits][u] for
// geometries with G::HASH (16 bytes per window: the value bits of the piece are address bits 2 and 3).
// `idx` bit k <-> window bit {1, 2, 4, 5, S+2, S+3, S+4, 2S+3}[k].
BGS_HD uint32_t seg_window_of_idx(int S, uint32_t idx) {
    const int wb[8] = {1, 2, 4, 5, S + 2, S + 3, S + 4, 2 * S + 3};
    uint32_t x = 0;
    for (int k = 0; k < 8; ++k) x |= ((idx >> k) & 1u) << wb[k];
    return x;
}
template <class G>
BGS_HD uint32_t seg_hash(uint32_t window) {  // G::HASH only
    return ((window & G::kHashMask) * G::kHashMul) >> 24;
}
// table index of a window for geometry g: compile-time hash, run-time hash, or -- no multiplier for this stride --
// the 8 gathered bits
template <class G>
BGS_HD uint32_t seg_index(const G& g, uint32_t x) {
    if constexpr (G::HASH) return seg_hash<G>(x);
    else {
        const uint32_t mul = g.hmul();
        if (mul) return ((x & g.hmask()) * mul) >> 24;
        const int S = g.s();
        return ((x >> 1) & 3u) | ((x >> 2) & 0xCu) | ((x >> (S - 2)) & 0x70u) | ((x >> (2 * S - 4)) & 0x80u);
    }
}
template <class G>
BGS_HD uint32_t seg_lut_slot(const G& g, int u, uint32_t idx) {
    return seg_index(g, seg_window_of_idx(g.s(), idx)) * 4u + (uint32_t)u;
}
inline bool seg_hash_is_perfect(int S) {  // host-side check of seg_hash_mul(S)
    const uint32_t mul = seg_hash_mul(S), mask = seg_hash_mask(S);
    if (!mul) return true;
    bool seen[256] = {false};
    for (uint32_t i = 0; i < 256; ++i) {
        const uint32_t h = ((seg_window_of_idx(S, i) & mask) * mul) >> 24;
        if (h > 255u || seen[h]) return false;
        seen[h] = true;
    }
    return true;
}

// Power table of the table-driven segment: entry i (16 bytes) = {1 << i, 8 << i} as two 64-bit words, i = the
// pending cell's index minus 3; word k of the entry (0..3) is what this returns.
constexpr int SEG_POW_WORDS = 64 * 4;
BGS_HD uint32_t seg_pow_entry(int i, int k) {
    const uint64_t v = (k < 2 ? 1ull : 8ull) << i;  // (8 << i loses bits for i > 60: no such cell)
    return (uint32_t)(v >> (32 * (k & 1)));
}

#if defined(__CUDA_ARCH__)
// a * b + c on the FMA pipe (IMAD); with an opaque b ptxas cannot turn it into an ALU-pipe add / LEA
__device__ __forceinline__ uint32_t fma_mad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#endif

#if defined(__CUDA_ARCH__)
#define BGS_UNROLL _Pragma("unroll")
#else
#define BGS_UNROLL
#endif

// ---------------------------------------------------------------------------------------------
// MoveGen: the move generation of ONE position for its mover -- SURVEY.md 4.4 rules 2-4.
// RULES_ >= 0: compile-time rule set; -1: read g.rules.
// FASTPROBE: a probe stops at the first piece that can move and writes no target masks (the single-step kernels,
// which probe the opponent after EVERY move); the rollout kernels leave it off -- they probe once in 3000
// generations, and their piece boundary runs at ~8 of 32 lanes, where each instruction of a special case costs four.
// ---------------------------------------------------------------------------------------------
template <int NP, class G, int RULES_, bool FASTPROBE = false>
struct MoveGen {
    typedef typename G::bits B;
    B b[NP];  // value bit-planes in the mover's orientation (read-only here)
    B occ, src_left, sbit, occS, inter, open, unexp, pending, targets;  // unexp: pieces (of occS) not expanded yet; pending is a subset of it, except for the source cell itself
    int total, nsrc;
    bool probe;  // only "does the mover have any action?" (the blocked test): `found` is the answer, T / total are not used
    bool found, have, done;
    const uint32_t* lut;  // seg_lut_entry table [256][4] (seg_lut_slot) when G::LUT (shared memory in the kernel)
    uint32_t lut_saddr;   // device: the same table as a 32-bit shared-memory address (kept opaque by the kernel so
                          // that the base stays in a register instead of being rebuilt for every look-up); the
                          // power table (seg_pow_entry) follows it at byte 4096
    uint32_t one, k16;    // device: always 1 and 16, from kernel parameters -- multipliers ptxas cannot fold, which
                          // keeps adds and shifts by constants on the FMA pipe (IMAD)
    B b0r;                // b[0] >> 1 (table-driven segments: the value's bit 0 where the table address wants it)

    BGS_HD int rules(const G& g) const { return RULES_ >= 0 ? RULES_ : g.rules; }

    // The movable pieces of the mover: the occupied row nearest to it (tests/test_bounce.py:43-48,60).
    // no_moves: an ended start position (no generation at all).
    BGS_HD static B sources(const G& g, const B* planes, bool no_moves) {
        B o = planes[0];
        BGS_UNROLL
        for (int i = 1; i < NP; ++i) o |= planes[i];
        if (!o || no_moves) return (B)0;
        const int row = g.row_of(ctzb(o));
        return o & ((B)g.m_row0() << (row * g.s()));
    }

    // planes b[] already set; src = sources(g, b, no_moves), possibly computed earlier.
    BGS_HD void begin_with(const G& g, B src, bool prb) {
        B o = b[0];
        BGS_UNROLL
        for (int i = 1; i < NP; ++i) o |= b[i];
        occ = o;
        src_left = src;
        probe = prb; found = false; have = false; done = false;
        pending = 0; total = 0; nsrc = 0;
        if (NP == 2 && sizeof(B) == 8) b0r = b[0] >> 1;
    }

    BGS_HD void begin(const G& g, bool prb, bool no_moves) { begin_with(g, sources(g, b, no_moves), prb); }

    // One table-driven segment (pending != 0): one pending cell, its landing set from the table -- no step loop, no
    // dependence on the piece value, so every lane of the warp executes the same instructions.  The rollout kernel
    // also calls it on its own: a second segment for the lanes that still have a pending cell, without paying for
    // another piece-boundary block (37 % of the pieces need one segment, 23 % two, the rest up to twelve).
    BGS_HD void lut_segment(const G& g) {
        // The rollout kernel was bound by the ALU pipe (logic, shifts, compares, adds: 80 % busy where 81 % is its
        // ceiling) while the FMA pipe (integer multiply-add) idled at 15 %, so the segment is written for few
        // instructions on the former: 15 of its 28, where the first form had 28 of 34.
        //  * The HIGHEST pending cell is expanded (the closure does not depend on the order): FLO finds that bit
        //    directly -- no isolation of the lowest bit (a 64-bit negate and two ANDs).
        //  * Everything is taken from words shifted by c - 3: the window of passable cells (bit 3 = c) and the value
        //    bits of c, which land on bits 2 and 3 -- the table is laid out [window][value], so they ARE address bits
        //    and the value itself is never extracted.
        //  * 1 << (c - 3) and the bit of c itself come from a 64-entry table (one 128-bit load), and the landing set
        //    is entry * 2^(c-3) as a wide multiply instead of a 64-bit shift.
        //  * unexp (pieces not expanded yet) replaces the set of expanded cells: c leaves pending because it is no
        //    longer in unexp, one three-input logic operation per half.
        //  * The adds and the table strides are multiply-adds with opaque multipliers (one, k16).
#if defined(__CUDA_ARCH__)
        const uint32_t plo = (uint32_t)pending, phi = (uint32_t)(pending >> 32);
        uint32_t c3;  // c - 3, c >= S >= 3: no piece in row 0
        asm("{\n\t.reg .pred p;\n\t.reg .u32 f;\n\t"
            "setp.ne.u32 p, %2, 0;\n\t"
            "bfind.u32 f, %1;\n\t"
            "@p bfind.u32 f, %2;\n\t"
            "mad.lo.u32 %0, %3, 0xfffffffd, f;\n\t"
            "@p mad.lo.u32 %0, %3, 29, f;\n\t}"
            : "=r"(c3) : "r"(plo), "r"(phi), "r"(one));
        uint32_t p1lo, p1hi, lowlo, lowhi;  // 1 << c3 and 8 << c3
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4 + 4096];"
                     : "=r"(p1lo), "=r"(p1hi), "=r"(lowlo), "=r"(lowhi) : "r"(fma_mad(c3, k16, lut_saddr)));
        const uint32_t x = (uint32_t)(inter >> c3);
        // bit 0 of the value -> address bit 2 (b0r = b[0] >> 1), bit 1 -> address bit 3
        const uint32_t vv = ((uint32_t)(b0r >> c3) & 4u) | ((uint32_t)(b[1] >> c3) & 8u);
        uint32_t entry;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(entry) : "r"(fma_mad(seg_index(g, x), k16, vv) + lut_saddr));
        uint32_t ldlo, ldhi;  // the landing set = entry << c3 (not masked with `open`: occS lies inside it, targets are
                              // masked at the piece boundary)
        asm("{\n\t.reg .u64 w;\n\t"
            "mul.wide.u32 w, %2, %3;\n\t"
            "mov.b64 {%0, %1}, w;\n\t"
            "mad.lo.u32 %1, %2, %4, %1;\n\t}"
            : "=r"(ldlo), "=&r"(ldhi) : "r"(entry), "r"(p1lo), "r"(p1hi));
        const B land = (B)(((uint64_t)ldhi << 32) | ldlo);
        const B low = (B)(((uint64_t)lowhi << 32) | lowlo);
#else
        const int c3 = top_bit_minus3((uint64_t)pending);
        const uint32_t x = (uint32_t)(inter >> c3);
        const uint32_t u = ((uint32_t)(b[0] >> c3) >> 3 & 1u) | ((uint32_t)(b[1] >> c3) >> 2 & 2u);
        const B low = (B)8 << c3;
        const B land = (B)lut[seg_index(g, x) * 4u + u] << c3;
#endif
        unexp &= ~low;
        targets |= land & ~occS;
        pending = (pending | land) & unexp;  // c leaves (it is no longer in unexp), unexpanded pieces landed on enter
    }

    // One iteration: at most one piece boundary, then one whole segment.  T[j * stride] receives the
    // target mask (mover-relative) of the j-th movable piece.  Sets done when nothing is left.
    BGS_HD void iter(const G& g, B* T, int stride) {
        const int rl = rules(g);
        if (pending == 0) {  // piece boundary
            if (have) {
                // (`& open`: the table-driven segments leave the landing sets unmasked -- guard cells, cells above
                // the board -- and the mask is applied once per piece here)
                const B tg = ((rl & BGS_BOUNCE_ALLOW_NULL_MOVE) ? targets : (targets & ~sbit)) & open;
                if (FASTPROBE && probe) {
                    found = tg != 0;
                } else {  // (without FASTPROBE a probe takes this path too: its target masks are simply not used)
                    T[nsrc * stride] = tg;
                    total += popcb(tg);
                    ++nsrc;
                }
                have = false;
            }
            if (src_left == 0 || (FASTPROBE && found)) {
                done = true;
                if (!(FASTPROBE && probe)) found = total != 0;
                return;
            }
            sbit = src_left & (~src_left + (B)1);  // next movable piece, ascending relative column
            src_left ^= sbit;
            have = true;
            const int variant = rl & 3;
            occS = variant == BGS_BOUNCE_SOURCE_PIECE ? occ : (occ & ~sbit);
            open = variant == BGS_BOUNCE_SOURCE_BLOCKED ? (g.m_board() & ~sbit) : g.m_board();
            inter = open & ~occS & ~g.m_far();  // cells a path may pass through
            unexp = occS;
            targets = 0;
            pending = sbit;  // the first segment = a "bounce" off the piece itself
        }
        if (NP == 2 && sizeof(B) == 8 && (G::LUT || (g.lut_rt() && lut != nullptr))) {
            lut_segment(g);
            return;
        }
        const B low = pending & (~pending + (B)1);
        // ---- segment setup: all unexpanded landing cells holding a piece of the same value as the
        // lowest one travel together
        B S = pending;
        int u = 0;
        BGS_UNROLL
        for (int i = 0; i < NP; ++i) {
            const bool in = (b[i] & low) != 0;
            S &= in ? b[i] : ~b[i];
            u |= in ? (1 << i) : 0;
        }
        unexp &= ~S;
        // ---- u steps, every frontier cell at once; x* = cells entered by a forward / left / right
        // step (a left step may not follow a right step and vice versa; never backwards)
        const int W = g.s();  // one row up
        B xf = S << W;
        B xl = (S & g.m_not_left()) >> 1;
        B xr = (S & g.m_not_right()) << 1;
        // intermediate cells: empty, not the far goal row
        auto advance = [&]() {
            const B af = (xf | xl | xr) & inter;
            const B al = (xf | xl) & inter & g.m_not_left();
            const B ar = (xf | xr) & inter & g.m_not_right();
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
        const B land = (xf | xl | xr) & open;
        targets |= land & ~occS;
        pending = (pending | land) & unexp;  // S leaves, unexpanded pieces landed on enter
    }
};

// What the game needs next after a transition.
enum Next { NEXT_OVER = 0, NEXT_MOVEGEN = 1, NEXT_PROBE = 2 };

// ---------------------------------------------------------------------------------------------
// Game: one rollout between two move generations -- the caller's uniform choice (README.md:61-62)
// + Action::sample_next_state + has_ended / reward (SURVEY.md 4.4 rules 5-6).
// ---------------------------------------------------------------------------------------------
template <int NP, class G>
struct Game {
    typedef typename G::bits B;
    B b[NP];  // value bit-planes, oriented for `orient`
    int t;           // plies played in this rollout
    int player;      // side to move
    int orient;      // whose orientation b[] is in
    int win;

    BGS_HD B rot(const G& g, B x) const { return revb(x) >> g.rot_sh(); }
    BGS_HD void orient_for(const G& g, int pl) {
        if (orient != pl) {
            BGS_UNROLL
            for (int i = 0; i < NP; ++i) b[i] = rot(g, b[i]);
            orient = pl;
        }
    }

    BGS_HD void begin_planes(const G& g, const B* plane0) {
        BGS_UNROLL
        for (int i = 0; i < NP; ++i) b[i] = plane0[i];
        t = 0; player = 0; orient = 0; win = BGS_WINNER_DRAW;
    }

    // Per-game start position in the reference's layout (int8 grid, row 0 = bottom).  Returns true if
    // the position has already ended (the first move generation must then be started with no_moves).
    BGS_HD bool begin_grid(const G& g, const int8_t* grid, int pl, int winner_in, bool ended) {
        BGS_UNROLL
        for (int i = 0; i < NP; ++i) b[i] = 0;
        for (int y = 0; y < g.h(); ++y)
            for (int x = 0; x < g.w(); ++x) {
                const int v = grid[y * g.w() + x];
                BGS_UNROLL
                for (int i = 0; i < NP; ++i) b[i] |= (B)((v >> i) & 1) << (y * g.s() + x);
            }
        t = 0; player = pl & 1; orient = 0; win = winner_in;
        orient_for(g, player);
        return ended || winner_in >= 0;
    }

    // The ply transition after a complete move generation (total / found / probe are its results, T
    // its target masks, rr the draw of ply t -- read only when a move is played).  On NEXT_MOVEGEN /
    // NEXT_PROBE the planes are oriented for the player whose moves must be generated next.
    template <class Draw>
    BGS_HD Next transition(const G& g, const B* T, int stride, int total, bool probe, bool found,
                           int max_plies, uint8_t* moves_row, Draw draw) {
        if (probe) {  // `player` is blocked; the previous mover wins unless blocked too (draw)
            win = found ? 1 - player : BGS_WINNER_DRAW;
            return NEXT_OVER;
        }
        if (total == 0) {
            if (t == 0) return NEXT_OVER;  // a blocked / ended start position: no winner
            orient_for(g, 1 - player);
            return NEXT_PROBE;
        }
        if (t >= max_plies) {
            win = BGS_WINNER_TRUNCATED;
            return NEXT_OVER;
        }
        int k = (int)mulhi32(draw(t), (uint32_t)total);
        if (player) k = total - 1 - k;  // canonical (absolute) order is the reverse of the rotated one
        // k-th action in ascending relative (source, target) order
        B occ = b[0];
        BGS_UNROLL
        for (int i = 1; i < NP; ++i) occ |= b[i];
        const int base = g.row_of(ctzb(occ)) * g.s();
        uint32_t sm = (uint32_t)(occ >> base) & g.m_row0();
        B tm = T[0];
        int j = 0;
        for (;;) {
            const int c = popcb(tm);
            if (k < c) break;
            k -= c;
            ++j;
            sm &= sm - 1u;
            tm = T[j * stride];
        }
        const int scell = base + ctz32(sm);
        const int tcell = kth_set_bit(tm, k);
        if (moves_row) {
            const int HW1 = g.h() * g.w() - 1, sp = pub_cell(g, scell), tp = pub_cell(g, tcell);
            moves_row[2 * t] = (uint8_t)(player ? HW1 - sp : sp);
            moves_row[2 * t + 1] = (uint8_t)(player ? HW1 - tp : tp);
        }
        const B smask = (B)1 << scell, tmask = (B)1 << tcell;
        BGS_UNROLL
        for (int i = 0; i < NP; ++i) {
            const bool has = (b[i] & smask) != 0;
            b[i] = (b[i] & ~smask) | (has ? tmask : (B)0);
        }
        ++t;
        const bool goal = (tmask & g.m_far()) != 0;
        if (goal) win = player;
        player ^= 1;
        if (goal) return NEXT_OVER;
        orient_for(g, player);
        return NEXT_MOVEGEN;
    }

    BGS_HD int value_abs(const G& g, int x, int y) const {  // absolute (x, y)
        const int c = orient ? (g.h() - 1 - y) * g.s() + (g.w() - 1 - x) : y * g.s() + x;
        int v = 0;
        BGS_UNROLL
        for (int i = 0; i < NP; ++i) v |= (int)((b[i] >> c) & (B)1) << i;
        return v;
    }

    // Result of a finished game (lengths / winners / final grid / reward; statistics are the caller's).
    BGS_HD void write_result(const G& g, const LaneOut& o, size_t idx) const {
        if (o.length) o.length[idx] = (uint16_t)t;
        if (o.winner) o.winner[idx] = (int8_t)win;
        if (o.final_grid) {
            int8_t* out = o.final_grid + idx * (size_t)(g.h() * g.w());
            for (int y = 0; y < g.h(); ++y)
                for (int x = 0; x < g.w(); ++x) out[y * g.w() + x] = (int8_t)value_abs(g, x, y);
        }
        if (o.reward) {
            o.reward[2 * idx] = win == 0 ? 1.f : (win == 1 ? -1.f : 0.f);
            o.reward[2 * idx + 1] = win == 1 ? 1.f : (win == 0 ? -1.f : 0.f);
        }
    }
};

// The t-th draw of game gid: word t & 3 of Philox(key = seed, ctr = (gid, t >> 2, 1)).
BGS_HD uint32_t bounce_draw(uint64_t gid, uint32_t seed_lo, uint32_t seed_hi, int t) {
    uint32_t r[4];
    philox_hd((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t >> 2, 1u, seed_lo, seed_hi, r);
    return (t & 3) == 0 ? r[0] : ((t & 3) == 1 ? r[1] : ((t & 3) == 2 ? r[2] : r[3]));
}

}  // namespace bounce
}  // namespace bgs
