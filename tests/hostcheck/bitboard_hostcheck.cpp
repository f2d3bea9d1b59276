// Test-only: compiles the DEVICE header csrc/bitboard.cuh for the host so the branch-free
// REF-rules bit operations can be checked against the oracle without a GPU.
// Not part of the product library.
#include "../../othello_reinforcement_learning_test_b200/csrc/bitboard.cuh"

extern "C" {

void hc_legal(const uint64_t* s, const uint64_t* o, uint64_t* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = oth::legal_moves(s[i], o[i]);
}

void hc_flips(const uint64_t* s, const uint64_t* o, const int32_t* pos, uint64_t* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = (pos[i] >= 0 && pos[i] < 64) ? oth::flip_bits(pos[i], s[i], o[i]) : 0;
}

void hc_make_move(uint64_t* s, uint64_t* o, int32_t* mc, const int32_t* a, uint8_t* ok, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        int m = mc[i];
        ok[i] = oth::make_move(s[i], o[i], m, a[i]) ? 1 : 0;
        mc[i] = m;
    }
}

void hc_terminal_winner(const uint64_t* s, const uint64_t* o, uint8_t* t, int8_t* w, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) { t[i] = oth::is_terminal(s[i], o[i]); w[i] = (int8_t)oth::winner(s[i], o[i]); }
}

int hc_nth_set_bit(uint64_t m, int k) { return oth::nth_set_bit(m, k); }

// same walk as k_playouts, one game
int hc_playout(uint64_t seed, int64_t g, uint64_t* fs, uint64_t* fo)
{
    uint64_t me = oth::kStartSelf, you = oth::kStartOpp;
    const uint64_t key = oth::mix64(seed ^ oth::mix64((uint64_t)g));
    uint64_t ctr = 0;
    int plies = 0;
    for (;;) {
        const uint64_t lg = oth::legal_moves(me, you);
        if (lg == 0) {
            if (oth::legal_moves(you, me) == 0) break;
            const uint64_t t = me; me = you; you = t; ++plies;
            continue;
        }
        const int n = oth::popc64(lg);
        const uint64_t r = oth::mix64(key + (ctr++) * 0xD1342543DE82EF95ULL);
        const int pick = (int)(((r >> 32) * (uint64_t)n) >> 32);
        oth::apply_known_legal(me, you, oth::nth_set_bit(lg, pick));
        ++plies;
    }
    *fs = me; *fo = you;
    return plies;
}
}
