// bitboard.cuh -- REF-rules Othello bit operations, one game per thread, branch-free.
//
// Replaces the data-dependent ray walks of the reference
// (src/cython/bitboard.pyx:71-158, tables at :20-38) with fixed-length
// shift-and-mask directional fills.  "REF rules" = the reference's rule set, in
// which the file mask is ANDed in AFTER the shift (bitboard.pyx:91-92,104-105);
// this differs from standard Othello on the A/H files and is reproduced here
// bit for bit (parity target: tests/golden/bitboard.json, ref_games.npz).
//
// Six fill steps per ray are sufficient: a ray can cross at most six opponent
// discs before it must meet an own disc (vertical rays: 8 rows; horizontal and
// diagonal rays, including the wrapped ones the after-shift mask permits, are
// cut by the masked file within seven squares).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OTH_HD __host__ __device__ __forceinline__
#else
#define OTH_HD inline
#endif

namespace oth {

constexpr uint64_t kNotA = 0xFEFEFEFEFEFEFEFEULL;
constexpr uint64_t kNotH = 0x7F7F7F7F7F7F7F7FULL;
constexpr uint64_t kAll = 0xFFFFFFFFFFFFFFFFULL;
constexpr uint64_t kStartSelf = (1ULL << 28) | (1ULL << 35);   // bitboard.pyx:60-66
constexpr uint64_t kStartOpp = (1ULL << 27) | (1ULL << 36);
constexpr int kPass = 64;

// One ray = (shift amount, direction of shift, keep-mask applied after the shift).
template <int SH, bool LEFT, uint64_t KEEP>
struct Ray {
    // forward step of the reference walk: shift, then mask
    static OTH_HD uint64_t fwd(uint64_t x) { return (LEFT ? (x << SH) : (x >> SH)) & KEEP; }
    // exact pre-image of fwd on the board
    static OTH_HD uint64_t bwd(uint64_t y) { return LEFT ? ((y & KEEP) >> SH) : ((y & KEEP) << SH); }
};

// the eight rays in the reference's order: -8,+8,-1,+1,-9,-7,+7,+9 (bitboard.pyx:20,29-38)
using RayN = Ray<8, false, kAll>;
using RayS = Ray<8, true, kAll>;
using RayW = Ray<1, false, kNotA>;
using RayE = Ray<1, true, kNotH>;
using RayNW = Ray<9, false, kNotA>;
using RayNE = Ray<7, false, kNotH>;
using RaySW = Ray<7, true, kNotA>;
using RaySE = Ray<9, true, kNotH>;

template <class R>
OTH_HD uint64_t ray_legal(uint64_t me, uint64_t you, uint64_t vacant)
{
    uint64_t t = R::bwd(me) & you;
    t |= R::bwd(t) & you;
    t |= R::bwd(t) & you;
    t |= R::bwd(t) & you;
    t |= R::bwd(t) & you;
    t |= R::bwd(t) & you;
    return R::bwd(t) & vacant;
}

template <class R>
OTH_HD uint64_t ray_flips(uint64_t placed, uint64_t me, uint64_t you)
{
    uint64_t f = R::fwd(placed) & you;
    f |= R::fwd(f) & you;
    f |= R::fwd(f) & you;
    f |= R::fwd(f) & you;
    f |= R::fwd(f) & you;
    f |= R::fwd(f) & you;
    return (R::fwd(f) & me) ? f : 0ULL;
}

// == _compute_legal_moves (bitboard.pyx:135-158)
OTH_HD uint64_t legal_moves(uint64_t me, uint64_t you)
{
    const uint64_t vacant = ~(me | you);
    return ray_legal<RayN>(me, you, vacant) | ray_legal<RayS>(me, you, vacant) |
           ray_legal<RayW>(me, you, vacant) | ray_legal<RayE>(me, you, vacant) |
           ray_legal<RayNW>(me, you, vacant) | ray_legal<RayNE>(me, you, vacant) |
           ray_legal<RaySW>(me, you, vacant) | ray_legal<RaySE>(me, you, vacant);
}

// == _get_flip_bits (bitboard.pyx:116-133); pos in 0..63
OTH_HD uint64_t flip_bits(int pos, uint64_t me, uint64_t you)
{
    const uint64_t p = 1ULL << pos;
    return ray_flips<RayN>(p, me, you) | ray_flips<RayS>(p, me, you) |
           ray_flips<RayW>(p, me, you) | ray_flips<RayE>(p, me, you) |
           ray_flips<RayNW>(p, me, you) | ray_flips<RayNE>(p, me, you) |
           ray_flips<RaySW>(p, me, you) | ray_flips<RaySE>(p, me, you);
}

OTH_HD int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

OTH_HD int ctz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

// index of the k-th (0-based) set bit of m; m must have more than k bits set
OTH_HD int nth_set_bit(uint64_t m, int k)
{
    // branch-free binary search on popcounts (6 halvings)
    int pos = 0;
    int c = popc64(m & 0xFFFFFFFFULL);
    if (k >= c) { k -= c; pos += 32; m >>= 32; }
    c = popc64(m & 0xFFFFULL);
    if (k >= c) { k -= c; pos += 16; m >>= 16; }
    c = popc64(m & 0xFFULL);
    if (k >= c) { k -= c; pos += 8; m >>= 8; }
    c = popc64(m & 0xFULL);
    if (k >= c) { k -= c; pos += 4; m >>= 4; }
    c = popc64(m & 0x3ULL);
    if (k >= c) { k -= c; pos += 2; m >>= 2; }
    c = (int)(m & 1ULL);
    if (k >= c) { pos += 1; }
    return pos;
}

// == make_move (bitboard.pyx:195-247) including every reject path; returns ok.
// State is untouched when the move is rejected.
OTH_HD bool make_move(uint64_t& me, uint64_t& you, int& move_count, int pos)
{
    if (pos == kPass) {                                   // :208-219
        if (legal_moves(me, you) != 0) return false;
        const uint64_t t = me; me = you; you = t;
        ++move_count;
        return true;
    }
    if (pos < 0 || pos > 63) return false;                // :222-223
    const uint64_t bit = 1ULL << pos;
    if ((me | you) & bit) return false;                   // :228-229
    const uint64_t turned = flip_bits(pos, me, you);
    if (turned == 0) return false;                        // :235-236
    const uint64_t mine = me | bit | turned;              // :239
    const uint64_t theirs = you & ~turned;                // :240
    me = theirs; you = mine;                              // :243 (swap)
    ++move_count;                                         // :244
    return true;
}

// make_move for an action already known to be legal (or a forced pass): no checks.
OTH_HD void apply_known_legal(uint64_t& me, uint64_t& you, int pos)
{
    if (pos == kPass) { const uint64_t t = me; me = you; you = t; return; }
    const uint64_t turned = flip_bits(pos, me, you);
    const uint64_t mine = me | (1ULL << pos) | turned;
    const uint64_t theirs = you & ~turned;
    me = theirs; you = mine;
}

// == is_terminal (bitboard.pyx:249-264)
OTH_HD bool is_terminal(uint64_t me, uint64_t you)
{
    if (legal_moves(me, you) != 0) return false;
    return legal_moves(you, me) == 0;
}

// == get_winner (bitboard.pyx:266-282): side-to-move perspective
OTH_HD int winner(uint64_t me, uint64_t you)
{
    const int a = popc64(me), b = popc64(you);
    return (a > b) - (a < b);
}

// counter-based RNG (same definition as oracle/ref_rules.c mix64, so playouts can be
// compared bit for bit with the CPU checker)
OTH_HD uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // namespace oth
