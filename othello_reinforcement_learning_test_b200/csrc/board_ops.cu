// board_ops.cu -- batched bitboard kernels: thousands of games as SoA uint64 x2.
//
// One game per thread, branch-free move generation (bitboard.cuh), coalesced 8-byte
// loads from the SoA arrays, grid sized as a multiple of the SM count.  These kernels
// stand behind OthelloBitboard's methods (src/cython/bitboard.pyx) for batched callers
// and behind benchmark.py's random playout.
#include "bitboard.cuh"
#include "common.cuh"

namespace oth {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) k_legal(const uint64_t* __restrict__ me, const uint64_t* __restrict__ you,
                                                  uint64_t* __restrict__ out, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = legal_moves(me[i], you[i]);
}

__global__ void __launch_bounds__(kBlock) k_flips(const uint64_t* __restrict__ me, const uint64_t* __restrict__ you,
                                                  const int32_t* __restrict__ pos, uint64_t* __restrict__ out, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = pos[i];
        out[i] = (p >= 0 && p < 64) ? flip_bits(p, me[i], you[i]) : 0ULL;
    }
}

__global__ void __launch_bounds__(kBlock) k_make_move(uint64_t* __restrict__ me, uint64_t* __restrict__ you,
                                                      int32_t* __restrict__ move_count, const int32_t* __restrict__ action,
                                                      uint8_t* __restrict__ ok, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t a = me[i], b = you[i];
        int mc = move_count ? move_count[i] : 0;
        const bool good = make_move(a, b, mc, action[i]);
        if (good) {
            me[i] = a; you[i] = b;
            if (move_count) move_count[i] = mc;
        }
        if (ok) ok[i] = good ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kBlock) k_terminal_winner(const uint64_t* __restrict__ me, const uint64_t* __restrict__ you,
                                                            uint8_t* __restrict__ term, int8_t* __restrict__ win,
                                                            int32_t* __restrict__ counts, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t a = me[i], b = you[i];
        if (term) term[i] = is_terminal(a, b) ? 1 : 0;
        if (win) win[i] = (int8_t)winner(a, b);
        if (counts) { counts[2 * i] = popc64(a); counts[2 * i + 1] = popc64(b); }
    }
}

// The single-board path: make_move (optional) + everything the callers ask about the position reached, one thread.
// Arguments are kernel parameters; the result goes to a mapped page-locked mailbox, sequence word last.
__global__ void k_board_step(uint64_t me, uint64_t you, int move_count, int action, uint64_t seq, oth_board_state* __restrict__ box)
{
    if (threadIdx.x != 0) return;
    int ok = 1;
    if (action != OTH_ACTION_NONE) ok = make_move(me, you, move_count, action) ? 1 : 0;     // bitboard.pyx:195-247
    const uint64_t lg = legal_moves(me, you);                                               // :135-158
    box->self_b = me; box->opp_b = you; box->legal = lg;
    box->move_count = move_count; box->ok = ok;
    box->terminal = (lg == 0 && legal_moves(you, me) == 0) ? 1 : 0;                          // :249-264
    box->winner = winner(me, you);                                                           // :266-282
    box->self_count = popc64(me); box->opp_count = popc64(you);                              // :292-298
    __threadfence_system();
    *reinterpret_cast<volatile uint64_t*>(&box->seq) = seq;
}

// float32 [n,3,8,8]; one thread writes 4 consecutive squares of one plane (16-byte store)
__global__ void __launch_bounds__(kBlock) k_tensor_input(const uint64_t* __restrict__ me, const uint64_t* __restrict__ you,
                                                         float4* __restrict__ out, int64_t n)
{
    const int64_t total = n * 48;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = t / 48;
        const int c = (int)(t % 48);
        const int plane = c >> 4, q = (c & 15) * 4;
        const uint64_t a = me[g], b = you[g];
        const uint64_t bits = plane == 0 ? a : (plane == 1 ? b : legal_moves(a, b));
        const uint32_t nib = (uint32_t)(bits >> q) & 0xFu;
        out[t] = make_float4((float)(nib & 1u), (float)((nib >> 1) & 1u), (float)((nib >> 2) & 1u), (float)((nib >> 3) & 1u));
    }
}

// ---- baseline players of the evaluation arena (src/eval/players.py) -------------------------------------
// RandomPlayer.get_action (players.py:60-67): uniform over get_legal_moves() (64 when the side must pass).
__global__ void __launch_bounds__(kBlock) k_choose_random(const uint64_t* __restrict__ me, const uint64_t* __restrict__ you,
                                                          const uint64_t* __restrict__ salt, uint64_t seed,
                                                          int32_t* __restrict__ action, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t lg = legal_moves(me[i], you[i]);
        if (lg == 0) { action[i] = kPass; continue; }
        const uint64_t r = mix64(seed ^ mix64(salt ? salt[i] : (uint64_t)i));
        action[i] = nth_set_bit(lg, (int)(((r >> 32) * (uint64_t)popc64(lg)) >> 32));
    }
}

// GreedyPlayer.get_action (players.py:79-113), including its scoring rule as written: after the trial move the
// board is seen from the other side; on even move counts the score is the mover's discs (get_stone_counts()[1]),
// on odd move counts it is get_stone_counts()[0] -- the OPPONENT's discs.  First maximum in ascending order.
__global__ void __launch_bounds__(kBlock) k_choose_greedy(const uint64_t* __restrict__ me, const uint64_t* __restrict__ you,
                                                          const int32_t* __restrict__ move_count,
                                                          int32_t* __restrict__ action, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t a = me[i], b = you[i];
        uint64_t lg = legal_moves(a, b);
        if (lg == 0) { action[i] = kPass; continue; }
        const bool even = (move_count[i] & 1) == 0;
        int best = -1, best_score = -1;
        while (lg) {
            const int sq = ctz64(lg);
            lg &= lg - 1;
            uint64_t x = a, y = b;
            apply_known_legal(x, y, sq);                 // x = new side to move, y = the mover
            const int score = even ? popc64(y) : popc64(x);
            if (score > best_score) { best_score = score; best = sq; }
        }
        action[i] = best;
    }
}

// ---- random playouts ---------------------------------------------------------------
// One game per thread, state in registers from reset to the terminal position; the only
// global traffic is the optional 20 B/game result record.  Warp ballots decide whether
// anyone needs the "can the opponent move?" test (pass / game over) and when the whole
// warp is finished.
__global__ void __launch_bounds__(kBlock) k_playouts(int64_t n_games, uint64_t seed,
                                                     unsigned long long* __restrict__ total_plies,
                                                     unsigned long long* __restrict__ hist,   // [3]
                                                     uint64_t* __restrict__ final_me, uint64_t* __restrict__ final_you,
                                                     int32_t* __restrict__ plies_out)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const unsigned full = 0xFFFFFFFFu;
    bool done = g >= n_games;
    const bool real = !done;
    uint64_t me = kStartSelf, you = kStartOpp;
    const uint64_t key = mix64(seed ^ mix64((uint64_t)g));
    uint64_t ctr = 0;
    int plies = 0;
    while (!__all_sync(full, done)) {
        const uint64_t lg = done ? 0ULL : legal_moves(me, you);
        const bool stuck = !done && lg == 0;
        if (__any_sync(full, stuck)) {
            const uint64_t other = legal_moves(you, me);
            if (stuck) {
                if (other == 0) {
                    done = true;                         // is_terminal (bitboard.pyx:249-264)
                } else {
                    const uint64_t t = me; me = you; you = t;   // forced pass (bitboard.pyx:209-219)
                    ++plies;
                }
            }
        }
        if (!done && !stuck) {
            const int n = popc64(lg);
            const uint64_t r = mix64(key + (ctr++) * 0xD1342543DE82EF95ULL);
            const int pick = (int)(((r >> 32) * (uint64_t)n) >> 32);
            apply_known_legal(me, you, nth_set_bit(lg, pick));
            ++plies;
        }
    }
    const int w = winner(me, you);
    if (real) {
        if (final_me) final_me[g] = me;
        if (final_you) final_you[g] = you;
        if (plies_out) plies_out[g] = plies;
    }
    // warp-aggregated statistics
    unsigned long long p = real ? (unsigned long long)plies : 0ULL;
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(full, p, o);
    const unsigned m0 = __ballot_sync(full, real && w < 0), m1 = __ballot_sync(full, real && w == 0),
                   m2 = __ballot_sync(full, real && w > 0);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(total_plies, p);
        if (m0) atomicAdd(&hist[0], (unsigned long long)__popc(m0));
        if (m1) atomicAdd(&hist[1], (unsigned long long)__popc(m1));
        if (m2) atomicAdd(&hist[2], (unsigned long long)__popc(m2));
    }
}

// ---- perft --------------------------------------------------------------------------
// Level-synchronous expansion until the frontier is wide enough, then one depth-first
// walk per frontier position.

__global__ void __launch_bounds__(kBlock) k_perft_expand(const uint64_t* __restrict__ me, const uint64_t* __restrict__ you,
                                                         int64_t n, uint64_t* __restrict__ out_me, uint64_t* __restrict__ out_you,
                                                         unsigned long long* __restrict__ out_count,
                                                         unsigned long long* __restrict__ leaf_nodes, int64_t out_cap,
                                                         int* __restrict__ overflow)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t a = me[i], b = you[i];
        uint64_t lg = legal_moves(a, b);
        if (lg == 0) {
            if (legal_moves(b, a) == 0) { atomicAdd(leaf_nodes, 1ULL); continue; }   // terminal = leaf
            const unsigned long long at = atomicAdd(out_count, 1ULL);
            if ((int64_t)at >= out_cap) { *overflow = 1; continue; }
            out_me[at] = b; out_you[at] = a;                                         // pass = one ply
            continue;
        }
        const int c = popc64(lg);
        const unsigned long long at = atomicAdd(out_count, (unsigned long long)c);
        if ((int64_t)(at + c) > out_cap) { *overflow = 1; continue; }
        for (int k = 0; k < c; ++k) {
            const int sq = ctz64(lg);
            lg &= lg - 1;
            uint64_t x = a, y = b;
            apply_known_legal(x, y, sq);
            out_me[at + k] = x; out_you[at + k] = y;
        }
    }
}

constexpr int kPerftMaxDfs = 8;

__global__ void __launch_bounds__(kBlock) k_perft_dfs(const uint64_t* __restrict__ me, const uint64_t* __restrict__ you,
                                                      int64_t n, int depth, unsigned long long* __restrict__ nodes)
{
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (depth == 0) { ++local; continue; }
        uint64_t sa[kPerftMaxDfs], sb[kPerftMaxDfs], todo[kPerftMaxDfs];
        bool pass_pending[kPerftMaxDfs];
        int lvl = 0;
        sa[0] = me[i]; sb[0] = you[i];
        // prepare level 0
        bool fresh = true;
        while (lvl >= 0) {
            if (fresh) {
                const uint64_t lg = legal_moves(sa[lvl], sb[lvl]);
                const int remaining = depth - lvl;
                if (lg == 0) {
                    const bool term = legal_moves(sb[lvl], sa[lvl]) == 0;
                    if (term || remaining == 1) { ++local; --lvl; fresh = false; continue; }   // leaf, or a single pass child
                    todo[lvl] = 0; pass_pending[lvl] = true;
                } else {
                    if (remaining == 1) { local += (unsigned long long)popc64(lg); --lvl; fresh = false; continue; }
                    todo[lvl] = lg; pass_pending[lvl] = false;
                }
                fresh = false;
            }
            if (pass_pending[lvl]) {
                pass_pending[lvl] = false;
                sa[lvl + 1] = sb[lvl]; sb[lvl + 1] = sa[lvl];
                ++lvl; fresh = true;
            } else if (todo[lvl]) {
                const int sq = ctz64(todo[lvl]);
                todo[lvl] &= todo[lvl] - 1;
                uint64_t x = sa[lvl], y = sb[lvl];
                apply_known_legal(x, y, sq);
                sa[lvl + 1] = x; sb[lvl + 1] = y;
                ++lvl; fresh = true;
            } else {
                --lvl;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(nodes, local);
}

}  // namespace oth

using namespace oth;

#define LAUNCH_CHECK(ctx)                          \
    do {                                           \
        (ctx)->launches++;                         \
        OTH_CHECK_CUDA(cudaGetLastError());        \
    } while (0)

extern "C" {

int oth_legal_moves(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, uint64_t* legal_out, int64_t n, int mem)
{
    OTH_REQUIRE(ctx && (n == 0 || (self_b && opp_b && legal_out)), OTH_ERR_ARG, "oth_legal_moves: NULL argument");
    OTH_REQUIRE(n >= 0, OTH_ERR_ARG, "oth_legal_moves: n < 0");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    Staged st(ctx, mem);
    const uint64_t* a = st.in(self_b, n); const uint64_t* b = st.in(opp_b, n); uint64_t* o = st.out(legal_out, n);
    if (st.failed) return OTH_ERR_CUDA;
    k_legal<<<grid_for(n, kBlock, ctx->sm_count), kBlock, 0, ctx->stream>>>(a, b, o, n);
    LAUNCH_CHECK(ctx);
    return st.finish();
}

int oth_flips(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, const int32_t* pos, uint64_t* flips_out,
              int64_t n, int mem)
{
    OTH_REQUIRE(ctx && (n == 0 || (self_b && opp_b && pos && flips_out)), OTH_ERR_ARG, "oth_flips: NULL argument");
    OTH_REQUIRE(n >= 0, OTH_ERR_ARG, "oth_flips: n < 0");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    Staged st(ctx, mem);
    const uint64_t* a = st.in(self_b, n); const uint64_t* b = st.in(opp_b, n); const int32_t* p = st.in(pos, n);
    uint64_t* o = st.out(flips_out, n);
    if (st.failed) return OTH_ERR_CUDA;
    k_flips<<<grid_for(n, kBlock, ctx->sm_count), kBlock, 0, ctx->stream>>>(a, b, p, o, n);
    LAUNCH_CHECK(ctx);
    return st.finish();
}

int oth_make_move(oth_ctx* ctx, uint64_t* self_b, uint64_t* opp_b, int32_t* move_count, const int32_t* action,
                  uint8_t* ok_out, int64_t n, int mem)
{
    OTH_REQUIRE(ctx && (n == 0 || (self_b && opp_b && action)), OTH_ERR_ARG, "oth_make_move: NULL argument");
    OTH_REQUIRE(n >= 0, OTH_ERR_ARG, "oth_make_move: n < 0");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    Staged st(ctx, mem);
    uint64_t* a = st.inout(self_b, n); uint64_t* b = st.inout(opp_b, n); int32_t* mc = st.inout(move_count, n);
    const int32_t* act = st.in(action, n); uint8_t* ok = st.out(ok_out, n);
    if (st.failed) return OTH_ERR_CUDA;
    k_make_move<<<grid_for(n, kBlock, ctx->sm_count), kBlock, 0, ctx->stream>>>(a, b, mc, act, ok, n);
    LAUNCH_CHECK(ctx);
    return st.finish();
}

int oth_terminal_winner(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, uint8_t* terminal_out,
                        int8_t* winner_out, int32_t* counts_out, int64_t n, int mem)
{
    OTH_REQUIRE(ctx && (n == 0 || (self_b && opp_b)), OTH_ERR_ARG, "oth_terminal_winner: NULL argument");
    OTH_REQUIRE(n >= 0, OTH_ERR_ARG, "oth_terminal_winner: n < 0");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    Staged st(ctx, mem);
    const uint64_t* a = st.in(self_b, n); const uint64_t* b = st.in(opp_b, n);
    uint8_t* t = st.out(terminal_out, n); int8_t* w = st.out(winner_out, n); int32_t* c = st.out(counts_out, 2 * n);
    if (st.failed) return OTH_ERR_CUDA;
    k_terminal_winner<<<grid_for(n, kBlock, ctx->sm_count), kBlock, 0, ctx->stream>>>(a, b, t, w, c, n);
    LAUNCH_CHECK(ctx);
    return st.finish();
}

int oth_board_step(oth_ctx* ctx, uint64_t self_b, uint64_t opp_b, int32_t move_count, int32_t action, oth_board_state* out)
{
    OTH_REQUIRE(ctx && out, OTH_ERR_ARG, "oth_board_step: NULL argument");
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    const uint64_t seq = ++ctx->mailbox_seq;
    k_board_step<<<1, 32, 0, ctx->stream>>>(self_b, opp_b, move_count, action, seq, ctx->mailbox_dev);
    LAUNCH_CHECK(ctx);
    // wait on the mailbox itself: a few microseconds sooner than cudaStreamSynchronize's wake-up; the stream sync
    // remains the fallback so that a failed launch surfaces as an error instead of a spin
    volatile uint64_t* flag = &ctx->mailbox->seq;
    for (int spin = 0; spin < 200000 && *flag != seq; ++spin) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    if (*flag != seq) OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    OTH_REQUIRE(*flag == seq, OTH_ERR_CUDA, "oth_board_step: the kernel did not deliver a result");
    *out = *ctx->mailbox;
    return OTH_OK;
}

int oth_tensor_input(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, float* out, int64_t n, int mem)
{
    OTH_REQUIRE(ctx && (n == 0 || (self_b && opp_b && out)), OTH_ERR_ARG, "oth_tensor_input: NULL argument");
    OTH_REQUIRE(n >= 0, OTH_ERR_ARG, "oth_tensor_input: n < 0");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    Staged st(ctx, mem);
    const uint64_t* a = st.in(self_b, n); const uint64_t* b = st.in(opp_b, n); float* o = st.out(out, n * 192);
    if (st.failed) return OTH_ERR_CUDA;
    OTH_REQUIRE(((uintptr_t)o & 15) == 0, OTH_ERR_ARG, "oth_tensor_input: output must be 16-byte aligned");
    k_tensor_input<<<grid_for(n * 48, kBlock, ctx->sm_count), kBlock, 0, ctx->stream>>>(a, b, (float4*)o, n);
    LAUNCH_CHECK(ctx);
    return st.finish();
}

int oth_choose_random(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, const uint64_t* salt, uint64_t seed,
                      int32_t* action_out, int64_t n, int mem)
{
    OTH_REQUIRE(ctx && (n == 0 || (self_b && opp_b && action_out)), OTH_ERR_ARG, "oth_choose_random: NULL argument");
    OTH_REQUIRE(n >= 0, OTH_ERR_ARG, "oth_choose_random: n < 0");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    Staged st(ctx, mem);
    const uint64_t* a = st.in(self_b, n); const uint64_t* b = st.in(opp_b, n); const uint64_t* sl = st.in(salt, n);
    int32_t* act = st.out(action_out, n);
    if (st.failed) return OTH_ERR_CUDA;
    k_choose_random<<<grid_for(n, kBlock, ctx->sm_count), kBlock, 0, ctx->stream>>>(a, b, sl, seed, act, n);
    LAUNCH_CHECK(ctx);
    return st.finish();
}

int oth_choose_greedy(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, const int32_t* move_count,
                      int32_t* action_out, int64_t n, int mem)
{
    OTH_REQUIRE(ctx && (n == 0 || (self_b && opp_b && move_count && action_out)), OTH_ERR_ARG, "oth_choose_greedy: NULL argument");
    OTH_REQUIRE(n >= 0, OTH_ERR_ARG, "oth_choose_greedy: n < 0");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    Staged st(ctx, mem);
    const uint64_t* a = st.in(self_b, n); const uint64_t* b = st.in(opp_b, n); const int32_t* mc = st.in(move_count, n);
    int32_t* act = st.out(action_out, n);
    if (st.failed) return OTH_ERR_CUDA;
    k_choose_greedy<<<grid_for(n, kBlock, ctx->sm_count), kBlock, 0, ctx->stream>>>(a, b, mc, act, n);
    LAUNCH_CHECK(ctx);
    return st.finish();
}

int oth_random_playouts(oth_ctx* ctx, int64_t n_games, uint64_t seed, int64_t* total_plies_out, int64_t* winner_hist_out,
                        uint64_t* final_self, uint64_t* final_opp, int32_t* plies, int mem)
{
    OTH_REQUIRE(ctx, OTH_ERR_ARG, "oth_random_playouts: ctx is NULL");
    OTH_REQUIRE(n_games >= 0, OTH_ERR_ARG, "oth_random_playouts: n_games < 0");
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    unsigned long long* stats = nullptr;   // [0]=plies, [1..3]=hist
    OTH_CHECK_CUDA(cudaMallocAsync((void**)&stats, 4 * sizeof(unsigned long long), ctx->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
    Staged st(ctx, mem);
    uint64_t* fs = st.out(final_self, n_games); uint64_t* fo = st.out(final_opp, n_games); int32_t* pl = st.out(plies, n_games);
    if (st.failed) return OTH_ERR_CUDA;
    if (n_games > 0) {
        const int64_t blocks = (n_games + kBlock - 1) / kBlock;
        k_playouts<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(n_games, seed, stats, stats + 1, fs, fo, pl);
        LAUNCH_CHECK(ctx);
    }
    unsigned long long h[4];
    OTH_CHECK_CUDA(cudaMemcpyAsync(h, stats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    OTH_CHECK_CUDA(cudaFreeAsync(stats, ctx->stream));
    int rc = st.finish();
    if (rc) return rc;
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (total_plies_out) *total_plies_out = (int64_t)h[0];
    if (winner_hist_out) { winner_hist_out[0] = (int64_t)h[1]; winner_hist_out[1] = (int64_t)h[2]; winner_hist_out[2] = (int64_t)h[3]; }
    return OTH_OK;
}

int oth_perft(oth_ctx* ctx, uint64_t self_b, uint64_t opp_b, int depth, uint64_t* nodes_out)
{
    OTH_REQUIRE(ctx && nodes_out, OTH_ERR_ARG, "oth_perft: NULL argument");
    OTH_REQUIRE(depth >= 0 && depth <= 16, OTH_ERR_ARG, "oth_perft: depth %d out of range 0..16", depth);
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int64_t cap = (int64_t)1 << 24;                       // frontier capacity per buffer (256 MiB total)
    const int64_t wide_enough = (int64_t)ctx->sm_count * 2048;  // switch to DFS once the frontier fills the GPU
    uint64_t *fa[2] = {nullptr, nullptr}, *fb[2] = {nullptr, nullptr};
    unsigned long long* counters = nullptr;                      // [0]=out_count, [1]=nodes
    int* overflow = nullptr;
    for (int k = 0; k < 2; ++k) {
        OTH_CHECK_CUDA(cudaMallocAsync((void**)&fa[k], cap * 8, ctx->stream));
        OTH_CHECK_CUDA(cudaMallocAsync((void**)&fb[k], cap * 8, ctx->stream));
    }
    OTH_CHECK_CUDA(cudaMallocAsync((void**)&counters, 2 * sizeof(unsigned long long), ctx->stream));
    OTH_CHECK_CUDA(cudaMallocAsync((void**)&overflow, sizeof(int), ctx->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), ctx->stream));
    OTH_CHECK_CUDA(cudaMemcpyAsync(fa[0], &self_b, 8, cudaMemcpyHostToDevice, ctx->stream));
    OTH_CHECK_CUDA(cudaMemcpyAsync(fb[0], &opp_b, 8, cudaMemcpyHostToDevice, ctx->stream));
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));          // self_b/opp_b are stack arguments
    int64_t width = 1;
    int cur = 0, remaining = depth;
    int rc = OTH_OK;
    while (remaining > 0 && width > 0 && (width < wide_enough || remaining > kPerftMaxDfs - 1) ) {
        OTH_CHECK_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned long long), ctx->stream));
        k_perft_expand<<<grid_for(width, kBlock, ctx->sm_count), kBlock, 0, ctx->stream>>>(
            fa[cur], fb[cur], width, fa[cur ^ 1], fb[cur ^ 1], counters, counters + 1, cap, overflow);
        LAUNCH_CHECK(ctx);
        unsigned long long w = 0; int ov = 0;
        OTH_CHECK_CUDA(cudaMemcpyAsync(&w, counters, sizeof w, cudaMemcpyDeviceToHost, ctx->stream));
        OTH_CHECK_CUDA(cudaMemcpyAsync(&ov, overflow, sizeof ov, cudaMemcpyDeviceToHost, ctx->stream));
        OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ov) { set_error("oth_perft: frontier exceeded %lld positions at remaining depth %d", (long long)cap, remaining); rc = OTH_ERR_CAPACITY; break; }
        width = (int64_t)w; cur ^= 1; --remaining;
    }
    if (rc == OTH_OK && width > 0) {
        k_perft_dfs<<<grid_for(width, kBlock, ctx->sm_count, 16), kBlock, 0, ctx->stream>>>(fa[cur], fb[cur], width, remaining, counters + 1);
        LAUNCH_CHECK(ctx);
    }
    unsigned long long total = 0;
    OTH_CHECK_CUDA(cudaMemcpyAsync(&total, counters + 1, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    for (int k = 0; k < 2; ++k) { cudaFreeAsync(fa[k], ctx->stream); cudaFreeAsync(fb[k], ctx->stream); }
    cudaFreeAsync(counters, ctx->stream); cudaFreeAsync(overflow, ctx->stream);
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *nodes_out = (uint64_t)total;
    return rc;
}

}  // extern "C"
