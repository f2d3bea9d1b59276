/*
 * othello_b200.h -- C ABI of libothello_b200.so (sm_100a).
 *
 * Drop-in boundary for the reference's self-play hot path.  The reference has no
 * FFI of its own: the boundary is Python duck typing at four call sites
 * (SURVEY.md section 8(b)).  Each entry point below names the reference interface
 * it stands behind (paths relative to the reference repository); the Python
 * classes in othello_reinforcement_learning_test_b200/ bind these with ctypes.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (0 = OTH_OK, <0 =
 *     error, text via oth_last_error()); nothing throws across the boundary;
 *   - opaque handles own device memory and are released by *_destroy;
 *   - batched arrays are caller-allocated structure-of-arrays; `mem` says where
 *     they live: OTH_MEM_DEVICE (device pointers, asynchronous on the context
 *     stream) or OTH_MEM_HOST (host pointers; the call copies in/out and returns
 *     when the results are in the caller's buffers);
 *   - one context is used by one host thread at a time (the reference is
 *     single-threaded under the GIL); every context owns one CUDA stream;
 *   - bit i of a board word is square i = row*8+col, A1 = bit 0
 *     (src/cython/bitboard.pxd:18-22); `self` is always the side to move.
 *   - there is no CPU fallback: without a CUDA device oth_ctx_create fails.
 */
#ifndef OTHELLO_B200_H
#define OTHELLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTH_OK 0
#define OTH_ERR_CUDA (-1)
#define OTH_ERR_ARG (-2)
#define OTH_ERR_STATE (-3)
#define OTH_ERR_CAPACITY (-4)
#define OTH_ERR_UNSUPPORTED (-5)

#define OTH_MEM_DEVICE 0
#define OTH_MEM_HOST 1

#define OTH_ACTIONS 65 /* 64 squares + pass (64) */

typedef struct oth_ctx oth_ctx;
typedef struct oth_net oth_net;
typedef struct oth_search oth_search;
typedef struct oth_selfplay oth_selfplay;

/* ---- library / context -------------------------------------------------- */
const char* oth_last_error(void);
const char* oth_version(void);
int oth_device_count(int* count);
int oth_ctx_create(int device, oth_ctx** out);
int oth_ctx_destroy(oth_ctx* ctx);
int oth_ctx_sync(oth_ctx* ctx);
/* cudaStream_t of the context, as an integer, for callers that enqueue their own work */
uint64_t oth_ctx_stream(oth_ctx* ctx);
/* kernels launched through this context so far (bench.py: gpu_launches) */
uint64_t oth_ctx_launch_count(oth_ctx* ctx);
/* Per-category kernel timing with CUDA events on the context stream (off by default).
 * Categories: 0 = network forward, 1 = tree kernels, 2 = self-play move kernels.
 * oth_ctx_timing_read synchronises, returns the summed milliseconds / launch counts since the last
 * read (arrays of 3) and resets the accumulators. */
int oth_ctx_timing_enable(oth_ctx* ctx, int on);
int oth_ctx_timing_read(oth_ctx* ctx, double* ms_out, uint64_t* count_out);

/* ---- bitboard: src/cython/bitboard.pyx --------------------------------------- */
/* OthelloBitboard.get_legal_moves_bits / _compute_legal_moves (bitboard.pyx:135-158,187-193) */
int oth_legal_moves(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, uint64_t* legal_out,
                    int64_t n, int mem);
/* _get_flip_bits (bitboard.pyx:116-133); pos[i] in 0..63, anything else yields 0 */
int oth_flips(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, const int32_t* pos,
              uint64_t* flips_out, int64_t n, int mem);
/* OthelloBitboard.make_move (bitboard.pyx:195-247) in place, every reject path included;
 * ok_out[i] = 1 if applied, 0 if rejected (state untouched).  move_count may be NULL. */
int oth_make_move(oth_ctx* ctx, uint64_t* self_b, uint64_t* opp_b, int32_t* move_count,
                  const int32_t* action, uint8_t* ok_out, int64_t n, int mem);
/* is_terminal + get_winner + get_stone_counts (bitboard.pyx:249-298).  Any output may be NULL. */
int oth_terminal_winner(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b,
                        uint8_t* terminal_out, int8_t* winner_out, int32_t* counts_out /* n*2 */,
                        int64_t n, int mem);
/* get_tensor_input (bitboard.pyx:300-323): float32 [n,3,8,8] = self, opp, legal planes */
int oth_tensor_input(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, float* out,
                     int64_t n, int mem);
/* ---- the single-board path: one launch per call, no allocation, no copies ---------------------------------
 * What src/eval/arena.py:106-119 needs per move from OthelloBitboard: make_move(action) (bitboard.pyx:195-247, every
 * reject path), then get_legal_moves / is_terminal / get_winner / get_stone_counts of the position reached
 * (:166-193, :249-298).  One one-warp kernel does all of it; arguments travel as kernel parameters, the result lands in a
 * page-locked mailbox mapped into the device (no cudaMalloc, no memcpy), the host waits on the mailbox's sequence word.
 * action == OTH_ACTION_NONE: no move, just the state of (self_b, opp_b). */
#define OTH_ACTION_NONE (-1000)
typedef struct {
    uint64_t self_b, opp_b;      /* position after the move (unchanged when it was rejected) */
    uint64_t legal;              /* get_legal_moves_bits of that position */
    int32_t move_count;
    int32_t ok;                  /* make_move's return value (1 for OTH_ACTION_NONE) */
    int32_t terminal, winner;    /* is_terminal, get_winner (side to move's view) */
    int32_t self_count, opp_count; /* get_stone_counts */
    uint64_t seq;                /* mailbox sequence word (internal) */
} oth_board_state;
int oth_board_step(oth_ctx* ctx, uint64_t self_b, uint64_t opp_b, int32_t move_count, int32_t action, oth_board_state* out);

/* perft under REF rules (pass = one ply, terminal = leaf); golden values in SURVEY.md 8(c) */
int oth_perft(oth_ctx* ctx, uint64_t self_b, uint64_t opp_b, int depth, uint64_t* nodes_out);
/* benchmark.py:18-40 play_random_game x n_games, one game per thread, from the start position.
 * total_plies_out / winner_hist_out[3] (get_winner == -1,0,+1) are HOST scalars; the per-game
 * arrays (final boards, plies) are optional and live where `mem` says. */
int oth_random_playouts(oth_ctx* ctx, int64_t n_games, uint64_t seed, int64_t* total_plies_out,
                        int64_t* winner_hist_out, uint64_t* final_self, uint64_t* final_opp,
                        int32_t* plies, int mem);

/* baseline players of the arena, batched: RandomPlayer.get_action (src/eval/players.py:60-67; salt[i] individualises
 * the draw, NULL = index) and GreedyPlayer.get_action (players.py:79-113, scoring rule exactly as written there) */
int oth_choose_random(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, const uint64_t* salt, uint64_t seed,
                      int32_t* action_out, int64_t n, int mem);
int oth_choose_greedy(oth_ctx* ctx, const uint64_t* self_b, const uint64_t* opp_b, const int32_t* move_count,
                      int32_t* action_out, int64_t n, int mem);

/* ---- network: src/model/net.py OthelloResNet ------------------------------------ */
#define OTH_NET_ENGINE_TCGEN05 0 /* bf16 tcgen05/TMEM implicit-GEMM trunk (product path) */
#define OTH_NET_ENGINE_SIMT 1    /* CUDA-core validation kernel, same rounding points */
#define OTH_NET_ENGINE_TCGEN05_PAIR 2 /* same trunk on CTA pairs (tcgen05 cta_group::2): half the weight traffic per SM */

#define OTH_NET_OUT_LOGPROBS 0   /* log_softmax, what model(x) returns (net.py:94) */
#define OTH_NET_OUT_PROBS 1      /* exp(log_softmax) as mcts.py:191 */
#define OTH_NET_OUT_PRIORS 2     /* masked to legal moves and renormalised as node.py:71-80 */

int oth_net_create(oth_ctx* ctx, int num_blocks, int num_filters, oth_net** out);
int oth_net_destroy(oth_net* net);
/* number of float32 values oth_net_load_weights expects */
int64_t oth_net_param_count(const oth_net* net);
/* `flat` = the reference state_dict's floating-point tensors concatenated in key order
 * (SURVEY.md 8(a) R-NN; num_batches_tracked skipped), HOST float32.  BatchNorm (eval mode,
 * eps 1e-5) is folded into bf16 conv weights + fp32 bias inside. */
int oth_net_load_weights(oth_net* net, const float* flat, int64_t count);
int oth_net_set_engine(oth_net* net, int engine);
/* the engine oth_net_forward / the searches will run: OTH_NET_ENGINE_* (tcgen05 for 64 or 128 filters, else the
 * validation engine), or OTH_ERR_ARG for a NULL handle */
int oth_net_engine(const oth_net* net);
/* OthelloResNet.forward on n positions given as bitboards (the (3,8,8) planes are built
 * in-kernel).  policy_out float32 [n,65] in `out_kind`; value_out float32 [n]. */
int oth_net_forward(oth_net* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n,
                    float* policy_out, float* value_out, int out_kind, int mem);

/* ---- tree search: src/mcts/mcts.py, src/mcts/node.py, BatchMCTS ------------------ */
#define OTH_FLAG_ROOT_N_SUM 1u   /* root N = completed sims (canonical); default: stays 0 (mcts.py:152-172) */
#define OTH_FLAG_Q_CANONICAL 2u  /* negate child Q in select; default: un-negated (node.py:113,119) */
#define OTH_FLAG_WINNER_BLACK 4u /* self-play labels from black's view; default: parallel_self_play.py:397-404 */
#define OTH_FLAG_EVAL_HASHNET 8u /* built-in integer test evaluator instead of the network */
#define OTH_FLAG_NO_SEARCH_SHARING 32u /* self-play: do NOT let slots with identical root positions share one search
                                          (sharing is result-transparent: the search is a deterministic function of the
                                          root position; it is switched off by itself when per-game Dirichlet noise
                                          enters the search, i.e. add_dirichlet_noise with OTH_FLAG_ROOT_N_SUM) */
#define OTH_FLAG_EVAL_CACHE 16u  /* position-keyed evaluation cache + same-step dedup in HBM (result-transparent:
                                    the network's output for a position does not depend on its batch slot) */

int oth_search_create(oth_ctx* ctx, int64_t max_games, int max_simulations, oth_search** out);
int oth_search_destroy(oth_search* s);
/* MCTS(model, device, c_puct, dirichlet_alpha, dirichlet_epsilon) (mcts.py:27-47) */
int oth_search_configure(oth_search* s, double c_puct, double dirichlet_alpha, double dirichlet_epsilon,
                         uint32_t flags);
/* fresh roots (mcts.py:71: no tree reuse) for n <= max_games positions */
int oth_search_begin(oth_search* s, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, int mem);
/* External-evaluator mode (parity with identical network outputs):
 *   collect: one selection per game (the very first call after begin returns the roots);
 *            need_eval[i] = 0 where the leaf was terminal and has been backed up already.
 *   apply:   expand the collected leaves with probs[n,65] / value[n] and back up. */
int oth_search_collect(oth_search* s, uint64_t* leaf_self, uint64_t* leaf_opp, uint8_t* need_eval, int mem);
int oth_search_apply(oth_search* s, const float* probs, const float* value, int mem);
/* MCTS.search / BatchMCTS.search_batch (mcts.py:49-98, parallel_self_play.py:80-170) fully on the
 * device with `net` (or the hash-net when OTH_FLAG_EVAL_HASHNET).  Call after oth_search_begin. */
int oth_search_run(oth_search* s, oth_net* net, int num_simulations, int add_dirichlet_noise, uint64_t seed);
/* Opt-in throughput mode (INFLIGHT_K): up to inflight_k simulations per game per network launch, kept apart by
 * virtual loss; deterministic; inflight_k == 1 is identical to oth_search_run, larger values deviate from the
 * reference's visit counts by construction (tests report the total-variation distance). */
int oth_search_run_waves(oth_search* s, oth_net* net, int num_simulations, int inflight_k, int add_dirichlet_noise,
                         uint64_t seed);
/* root child statistics: visits int32 [n,65], q float64 [n,65] (0 for non-children),
 * n_evals int32 [n].  Any may be NULL. */
int oth_search_results(oth_search* s, int32_t* visits, double* q, int32_t* n_evals, int mem);
/* evaluation statistics since the last call (HOST uint64[4]): network positions actually evaluated, cache hits,
 * same-step duplicates, hash collisions.  n_evals in oth_search_results keeps counting expansions, like the reference. */
int oth_search_stats(oth_search* s, uint64_t* out4);
/* forget every cached evaluation (call when the network's weights changed) */
int oth_search_invalidate_cache(oth_search* s);
/* get_policy_distribution (node.py:147-182): float32 [n,65] for temperature 0 or 1 (others: powf) */
int oth_search_policy(oth_search* s, double temperature, float* policy_out, int mem);

/* ---- self-play: src/train/parallel_self_play.py ---------------------------------- */
typedef struct {
    int32_t num_simulations;        /* mcts.num_simulations */
    int32_t temperature_threshold;  /* self_play.temperature_threshold */
    int32_t add_dirichlet_noise;
    int32_t concurrent_games;       /* device-resident game slots (num_parallel_games) */
    double c_puct, dirichlet_alpha, dirichlet_epsilon;
    uint32_t flags;                 /* OTH_FLAG_* */
    uint32_t schedule;              /* OTH_SCHEDULE_* (0 = auto) */
    uint64_t seed;                  /* move-sampling seed of the first campaign; later campaigns on the handle derive their own */
} oth_selfplay_config;

/* How a campaign is scheduled on the device; the records produced are identical.
 *   LOCKSTEP: one search per ply for every slot, 1 + num_simulations network launches per ply
 *             (BatchMCTS.search_batch, parallel_self_play.py:80-170); slots with identical roots share one search.
 *   ASYNC:    run-until-miss -- every slot keeps simulating, moving and starting its next search while its leaves hit
 *             the evaluation cache, and only stops when it needs the network: one launch per cache miss of the slowest
 *             slot instead of one per simulation.  Needs OTH_FLAG_EVAL_CACHE to pay off.
 *   AUTO:     ASYNC up to 8192 slots when the evaluation cache is on, LOCKSTEP otherwise. */
#define OTH_SCHEDULE_AUTO 0u
#define OTH_SCHEDULE_LOCKSTEP 1u
#define OTH_SCHEDULE_ASYNC 2u

/* one training sample, packed (expanded to (f32[3,8,8], f32[65], float) by the Python shim) */
typedef struct {
    uint64_t self_b, opp_b, legal;  /* get_tensor_input planes */
    int32_t game;                   /* episode index within the run */
    int16_t ply;
    int8_t value;                   /* winner * player (parallel_self_play.py:404) */
    uint8_t n_children;
    uint16_t visits[OTH_ACTIONS];   /* root child visit counts; policy = visits / sum */
    uint16_t pad[3];                /* sizeof(oth_sample) == 168 */
} oth_sample;

int oth_selfplay_create(oth_ctx* ctx, const oth_selfplay_config* cfg, oth_selfplay** out);
int oth_selfplay_destroy(oth_selfplay* sp);
/* ParallelSelfPlayWorker.execute_episodes(num_episodes) (parallel_self_play.py:282-322): plays
 * num_episodes games to completion with `net` (NULL + OTH_FLAG_EVAL_HASHNET = test evaluator).
 * n_samples_out / n_evals_out are HOST scalars. */
int oth_selfplay_run(oth_selfplay* sp, oth_net* net, int64_t num_episodes, int64_t* n_samples_out,
                     int64_t* n_evals_out);
/* statistics of the last run (HOST uint64[5]): [0..3] as oth_search_stats, [4] = searches actually run (slots with
 * identical root positions share one) */
int oth_selfplay_stats(oth_selfplay* sp, uint64_t* out5);
/* timing of the last run (HOST double[4]): [0] device milliseconds between the first and the last kernel of the campaign
 * (CUDA events on the context stream), [1] network launches (lock-steps / ticks), [2] OTH_SCHEDULE_* really used,
 * [3] kernels launched */
int oth_selfplay_timing(oth_selfplay* sp, double* out4);
/* new move-sampling seed for the next campaign on this handle (no re-allocation) */
int oth_selfplay_set_seed(oth_selfplay* sp, uint64_t seed);
/* copy the samples of the last run into a caller buffer (HOST or DEVICE) */
int oth_selfplay_fetch(oth_selfplay* sp, oth_sample* out, int64_t capacity, int mem);
/* device pointer + count of the last run's samples (for NCCL all-gather without a host hop) */
int oth_selfplay_samples_device(oth_selfplay* sp, const oth_sample** dev_ptr_out, int64_t* count_out);

/* ---- replay buffer: src/train/buffer.py ReplayBuffer ------------------------------------------------ */
typedef struct oth_replay oth_replay;
int oth_replay_create(oth_ctx* ctx, int64_t max_size, oth_replay** out);       /* ReplayBuffer(max_size), buffer.py:23-31 */
int oth_replay_destroy(oth_replay* r);
int64_t oth_replay_size(const oth_replay* r);                                  /* __len__, buffer.py:86-88 */
int oth_replay_clear(oth_replay* r);                                           /* clear, buffer.py:90-92 */
/* add / deque(maxlen).append of n packed samples (HOST or DEVICE, e.g. oth_selfplay_samples_device), buffer.py:33-45 */
int oth_replay_add(oth_replay* r, const oth_sample* samples, int64_t n, int mem);
/* sample (buffer.py:58-84) for caller-drawn logical indices (0 = oldest): states f32 [n,3,8,8], policies f32 [n,65],
 * values f32 [n] (= [n,1]); idx and outputs live where `mem` says */
int oth_replay_gather(oth_replay* r, const int64_t* idx, int64_t n, float* states, float* policies, float* values, int mem);
/* same with symmetry augmentation on the packed records: sym[i] in 0..7 selects the dihedral image 2*k + flip of sample i in
 * the order of OthelloBitboard.get_symmetries (src/cython/bitboard.pyx:338-370: np.rot90 k times, then np.flip of the
 * columns); the three bit-planes and the 64 square counts are permuted, the pass count is carried.  This is the
 * augmentation the reference declares (src/train/self_play.py:166-212) but never wired in. */
int oth_replay_gather_sym(oth_replay* r, const int64_t* idx, const uint8_t* sym, int64_t n, float* states, float* policies,
                          float* values, int mem);
/* OTH_MEM_DEVICE gathers cannot validate their indices on the host: an index outside [0, size) reads entry 0 instead of a
 * stale ring slot and raises a flag; this call synchronises and returns OTH_ERR_ARG once if the flag is up */
int oth_replay_check(oth_replay* r);
int oth_replay_value_stats(oth_replay* r, double* mean_out, double* std_out);  /* get_statistics, buffer.py:102-123 */

/* ---- diagnostics ---------------------------------------------------------------------- */
/* One accumulation chain of tcgen05.mma (M=128, N=n, K=16*k_steps, bf16 -> fp32) over a caller
 * supplied shared-memory image; descriptor fields (byte offsets into the image, LBO/SBO in bytes,
 * per-k-step advance) are the caller's.  d_out: HOST float32 [128][n].  Used by the tests to pin
 * the no-swizzle K-major descriptor semantics the convolution kernel relies on. */
int oth_debug_umma_probe(oth_ctx* ctx, const void* smem_image, int image_bytes, int n, int k_steps,
                         uint32_t a_off, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep,
                         uint32_t b_off, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep, float* d_out);

/* Per-layer clock64 stamps of CTA 0 of the tcgen05 network kernel for one forward pass over n positions
 * (HOST pointers).  trace_out[layer*8 + k]: k=0 MMA issue starts, 1 MMA issue done, 2/3 tile-0 epilogue
 * starts/ends, 4/5 tile-1 epilogue starts/ends.  Tooling for tools/net_trace.py. */
int oth_debug_net_trace(oth_net* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n,
                        uint64_t* trace_out, int trace_len);

#ifdef __cplusplus
}
#endif
#endif /* OTHELLO_B200_H */
