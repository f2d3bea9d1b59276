/*
 * oracle/ref_rules.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's Othello rule set ("REF rules")
 * and of its MCTS, used as the parity checker for the CUDA path.  Nothing in
 * the product package may link, load or call this; only tests/, the smoke
 * check and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Every function cites the reference lines it restates
 * (paths relative to the reference repository root).
 *
 * Pinned against: the compiled reference itself (oracle/_ref, built by
 * oracle/build_ref.py) -- see oracle/gen_golden.py and tests/golden/.
 */
#ifndef ORACLE_REF_RULES_H
#define ORACLE_REF_RULES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- bitboard (src/cython/bitboard.pyx) -------------------------------- */

/* flips for placing at pos; bitboard.pyx:71-133 */
uint64_t ref_flips(int pos, uint64_t self_b, uint64_t opp_b);
/* legal-move mask; bitboard.pyx:135-158 */
uint64_t ref_legal(uint64_t self_b, uint64_t opp_b);
/* make_move incl. pass (64) and all reject paths; bitboard.pyx:195-247.
 * returns 1 on success, 0 on reject (state untouched). */
int ref_make_move(uint64_t *self_b, uint64_t *opp_b, int *move_count, int *passed, int pos);
/* bitboard.pyx:249-264 */
int ref_is_terminal(uint64_t self_b, uint64_t opp_b);
/* bitboard.pyx:266-282 (side-to-move perspective) */
int ref_winner(uint64_t self_b, uint64_t opp_b);
/* bitboard.pyx:284-290 */
int ref_popcount(uint64_t x);
/* bitboard.pyx:300-323; out is float[3*64], channel-major */
void ref_tensor_input(uint64_t self_b, uint64_t opp_b, float *out);
/* bitboard.pyx:166-185: ascending list, or the single entry 64; returns count */
int ref_legal_list(uint64_t self_b, uint64_t opp_b, int *out65);

/* perft under REF rules: pass = one ply, terminal = leaf, children in
 * get_legal_moves order. */
uint64_t ref_perft(uint64_t self_b, uint64_t opp_b, int depth);

/* batched helpers for numpy/ctypes */
void ref_legal_batch(const uint64_t *self_b, const uint64_t *opp_b, uint64_t *out, int64_t n);
void ref_flips_batch(const uint64_t *self_b, const uint64_t *opp_b, const int32_t *pos, uint64_t *out, int64_t n);
void ref_make_move_batch(uint64_t *self_b, uint64_t *opp_b, int32_t *move_count,
                         const int32_t *action, uint8_t *ok, int64_t n);
void ref_terminal_winner_batch(const uint64_t *self_b, const uint64_t *opp_b,
                               uint8_t *terminal, int8_t *winner, int64_t n);
void ref_tensor_input_batch(const uint64_t *self_b, const uint64_t *opp_b, float *out, int64_t n);

/* random playouts in the shape of benchmark.py:18-40 (uniform legal move,
 * pass when the list is [64], stop at terminal).  Own RNG (splitmix64 keyed by
 * seed, game index): only RNG-free statistics are comparable to other
 * implementations.  winner_hist[0..2] = counts of get_winner()==-1,0,+1.
 * threads<=0 -> all hardware threads.  returns total plies. */
int64_t ref_random_playouts(int64_t n_games, uint64_t seed, int threads,
                            int64_t winner_hist[3], uint64_t *final_self, uint64_t *final_opp,
                            int32_t *plies_out);

/* ---- test evaluator ("hash-net") --------------------------------------- */
/* Deterministic integer-only stand-in for the network: 65 dyadic
 * pseudo-probabilities (w+1)*2^-22, w in [0,65535], and a value k*2^-19 in
 * [-1,1).  No transcendental maths, so every implementation (C, Python, CUDA)
 * agrees bit for bit. */
void ref_hashnet(uint64_t self_b, uint64_t opp_b, float *probs65, float *value);

/* ---- MCTS (src/mcts/node.py, src/mcts/mcts.py) -------------------------- */

typedef void (*ref_eval_fn)(uint64_t self_b, uint64_t opp_b, float *probs65, float *value, void *user);

typedef struct {
    double c_puct;
    int    num_simulations;
    /* flags replicate the reference's quirks by default (all 0 = REF) */
    int    root_n_sum;      /* 0: root visit_count stays 0 (mcts.py:152-172) */
    int    q_canonical;     /* 0: maximise child's own-perspective Q (node.py:113,119) */
} ref_mcts_cfg;

typedef struct {
    int32_t visits[65];     /* child visit counts at the root (0 for non-children) */
    double  q[65];          /* child.get_value() at the root */
    int     n_children;
    int     n_evals;        /* evaluator calls (root + non-terminal leaves) */
    int     max_depth;
    int     is_child[65];
} ref_mcts_result;

/* one MCTS.search (mcts.py:49-98) without the final policy shaping */
int ref_mcts_search(uint64_t self_b, uint64_t opp_b, const ref_mcts_cfg *cfg,
                    ref_eval_fn eval, void *user, ref_mcts_result *out);

/* lock-step batch of searches with an external, batched evaluator: the shape of
 * BatchMCTS.search_batch (src/train/parallel_self_play.py:80-197).  collect -> evaluate the
 * leaves with need[i] != 0 -> apply; the first collect after begin returns the roots. */
typedef struct ref_batch ref_batch;
ref_batch *ref_batch_create(int n_games, double c_puct, int max_simulations, int root_n_sum, int q_canonical);
void ref_batch_destroy(ref_batch *b);
int ref_batch_begin(ref_batch *b, const uint64_t *self_b, const uint64_t *opp_b, int n);
int ref_batch_collect(ref_batch *b, int n, uint64_t *leaf_self, uint64_t *leaf_opp, uint8_t *need);
int ref_batch_apply(ref_batch *b, int n, const float *probs, const float *value);
void ref_batch_visits(const ref_batch *b, int n, int32_t *visits, int32_t *n_evals);

/* masked renormalisation used by MCTSNode.expand (node.py:62-89) in numpy's
 * float32 arithmetic (pairwise 8-lane summation order); priors[65] out. */
void ref_expand_priors(const float *probs65, const int *legal, int n_legal, float *priors65);

/* root.get_policy_distribution (node.py:147-182) for T==0 and T==1 */
void ref_policy_from_visits(const int32_t *visits, const int *is_child, double temperature, float *policy65);

/* convenience: search a batch of roots with the built-in hash-net */
int ref_mcts_search_hashnet_batch(const uint64_t *self_b, const uint64_t *opp_b, int64_t n,
                                  double c_puct, int num_simulations, int threads,
                                  int32_t *visits_out /* n*65 */, int32_t *n_evals_out /* n */);

#ifdef __cplusplus
}
#endif
#endif
