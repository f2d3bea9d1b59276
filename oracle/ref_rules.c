/*
 * oracle/ref_rules.c -- TEST INFRASTRUCTURE ONLY (see ref_rules.h).
 *
 * Plain-C restatement of the reference's rule set and tree search.  It keeps
 * the reference's *observable* behaviour, including the behaviours the survey
 * lists as quirks:
 *   - the per-direction file mask is applied AFTER the shift
 *     (src/cython/bitboard.pyx:29-38 applied at :91-92,97,104-105,109), which
 *     is not standard Othello on the A/H files ("REF rules");
 *   - get_winner is from the side-to-move's perspective (bitboard.pyx:266-282);
 *   - the MCTS root is never updated, child Q is maximised un-negated
 *     (src/mcts/mcts.py:152-172, src/mcts/node.py:91-126).
 *
 * Parity status: PINNED -- checked against the compiled reference by
 * oracle/gen_golden.py (perft 0..9, edge vectors, 2,000 replayed random games,
 * MCTS visit vectors); the fixtures live in tests/golden/.
 */
#include "ref_rules.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* bitboard                                                                 */
/* ------------------------------------------------------------------------ */

#define FILE_A_OUT 0xFEFEFEFEFEFEFEFEULL /* bitboard.pyx:24 */
#define FILE_H_OUT 0x7F7F7F7F7F7F7F7FULL /* bitboard.pyx:26 */
#define EVERYTHING 0xFFFFFFFFFFFFFFFFULL

/* one entry per ray: signed shift and the mask that is ANDed in after the
 * shift (bitboard.pyx:20 and :29-38, same order) */
typedef struct { int shift; uint64_t keep; } ray_t;
static const ray_t RAYS[8] = {
    {-8, EVERYTHING}, {8, EVERYTHING}, {-1, FILE_A_OUT}, {1, FILE_H_OUT},
    {-9, FILE_A_OUT}, {-7, FILE_H_OUT}, {7, FILE_A_OUT}, {9, FILE_H_OUT},
};

static inline uint64_t advance(uint64_t bits, const ray_t *r)
{
    /* shift then mask: bitboard.pyx:90-92 / :103-105 */
    uint64_t moved = r->shift > 0 ? (bits << r->shift) : (bits >> (-r->shift));
    return moved & r->keep;
}

/* bitboard.pyx:71-114, one ray */
static uint64_t ray_flips(int pos, const ray_t *r, uint64_t mine, uint64_t theirs)
{
    uint64_t taken = 0;
    uint64_t at = advance(1ULL << pos, r);
    while (at & theirs) {          /* :95-97 / :107-109 */
        taken |= at;
        at = advance(at, r);
    }
    return (at & mine) ? taken : 0; /* :100-101 / :111-112 */
}

uint64_t ref_flips(int pos, uint64_t self_b, uint64_t opp_b)
{
    /* bitboard.pyx:116-133 */
    uint64_t all = 0;
    for (int d = 0; d < 8; ++d) all |= ray_flips(pos, &RAYS[d], self_b, opp_b);
    return all;
}

uint64_t ref_legal(uint64_t self_b, uint64_t opp_b)
{
    /* bitboard.pyx:135-158: every empty square with a non-empty flip set */
    uint64_t vacant = ~(self_b | opp_b);
    uint64_t ok = 0;
    for (int sq = 0; sq < 64; ++sq) {
        uint64_t bit = 1ULL << sq;
        if ((vacant & bit) && ref_flips(sq, self_b, opp_b)) ok |= bit;
    }
    return ok;
}

int ref_popcount(uint64_t x)
{
    /* bitboard.pyx:284-290 */
    int n = 0;
    for (; x; x &= x - 1) ++n;
    return n;
}

int ref_make_move(uint64_t *self_b, uint64_t *opp_b, int *move_count, int *passed, int pos)
{
    uint64_t me = *self_b, you = *opp_b;
    if (pos == 64) {                       /* bitboard.pyx:208-219 */
        if (ref_legal(me, you) != 0) return 0;
        *self_b = you; *opp_b = me;
        *move_count += 1;
        if (passed) *passed = 1;
        return 1;
    }
    if (pos < 0 || pos > 63) return 0;     /* :222-223 */
    uint64_t bit = 1ULL << pos;
    if ((me | you) & bit) return 0;        /* :228-229 */
    uint64_t turned = ref_flips(pos, me, you);
    if (!turned) return 0;                 /* :235-236 */
    me |= bit | turned;                    /* :239 */
    you &= ~turned;                        /* :240 */
    *self_b = you; *opp_b = me;            /* :243 swap */
    *move_count += 1;                      /* :244 */
    if (passed) *passed = 0;               /* :245 */
    return 1;
}

int ref_is_terminal(uint64_t self_b, uint64_t opp_b)
{
    /* bitboard.pyx:249-264 */
    if (ref_legal(self_b, opp_b)) return 0;
    return ref_legal(opp_b, self_b) == 0;
}

int ref_winner(uint64_t self_b, uint64_t opp_b)
{
    /* bitboard.pyx:266-282 */
    int a = ref_popcount(self_b), b = ref_popcount(opp_b);
    return (a > b) - (a < b);
}

void ref_tensor_input(uint64_t self_b, uint64_t opp_b, float *out)
{
    /* bitboard.pyx:300-323: planes = mine, theirs, legal */
    uint64_t planes[3] = { self_b, opp_b, ref_legal(self_b, opp_b) };
    for (int c = 0; c < 3; ++c)
        for (int sq = 0; sq < 64; ++sq)
            out[c * 64 + sq] = (float)((planes[c] >> sq) & 1ULL);
}

int ref_legal_list(uint64_t self_b, uint64_t opp_b, int *out65)
{
    /* bitboard.pyx:166-185: never empty, [64] means "must pass" */
    uint64_t ok = ref_legal(self_b, opp_b);
    if (!ok) { out65[0] = 64; return 1; }
    int n = 0;
    for (int sq = 0; sq < 64; ++sq)
        if (ok & (1ULL << sq)) out65[n++] = sq;
    return n;
}

uint64_t ref_perft(uint64_t self_b, uint64_t opp_b, int depth)
{
    if (depth == 0) return 1;
    if (ref_is_terminal(self_b, opp_b)) return 1;
    int moves[65];
    int n = ref_legal_list(self_b, opp_b, moves);
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) {
        uint64_t a = self_b, b = opp_b; int mc = 0;
        ref_make_move(&a, &b, &mc, 0, moves[i]);
        total += ref_perft(a, b, depth - 1);
    }
    return total;
}

void ref_legal_batch(const uint64_t *self_b, const uint64_t *opp_b, uint64_t *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = ref_legal(self_b[i], opp_b[i]);
}

void ref_flips_batch(const uint64_t *self_b, const uint64_t *opp_b, const int32_t *pos, uint64_t *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = ref_flips(pos[i], self_b[i], opp_b[i]);
}

void ref_make_move_batch(uint64_t *self_b, uint64_t *opp_b, int32_t *move_count,
                         const int32_t *action, uint8_t *ok, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        int mc = move_count[i];
        ok[i] = (uint8_t)ref_make_move(&self_b[i], &opp_b[i], &mc, 0, action[i]);
        move_count[i] = mc;
    }
}

void ref_terminal_winner_batch(const uint64_t *self_b, const uint64_t *opp_b,
                               uint8_t *terminal, int8_t *winner, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        terminal[i] = (uint8_t)ref_is_terminal(self_b[i], opp_b[i]);
        winner[i] = (int8_t)ref_winner(self_b[i], opp_b[i]);
    }
}

void ref_tensor_input_batch(const uint64_t *self_b, const uint64_t *opp_b, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) ref_tensor_input(self_b[i], opp_b[i], out + i * 192);
}

/* ------------------------------------------------------------------------ */
/* counter-based RNG shared (by definition) with the CUDA playout kernel      */
/* ------------------------------------------------------------------------ */

static inline uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

static inline int nth_set_bit(uint64_t m, int k)
{
    for (; k > 0; --k) m &= m - 1;
    return __builtin_ctzll(m);
}

static int one_playout(uint64_t key, uint64_t *fs, uint64_t *fo, int *winner)
{
    /* shape of benchmark.py:18-40 */
    uint64_t me = (1ULL << 28) | (1ULL << 35);     /* bitboard.pyx:60-66 */
    uint64_t you = (1ULL << 27) | (1ULL << 36);
    int plies = 0;
    uint64_t ctr = 0;
    for (;;) {
        uint64_t ok = ref_legal(me, you);
        if (!ok) {
            if (!ref_legal(you, me)) break;        /* terminal */
            uint64_t t = me; me = you; you = t;    /* pass */
            ++plies;
            continue;
        }
        int n = ref_popcount(ok);
        uint64_t r = mix64(key + (ctr++) * 0xD1342543DE82EF95ULL);
        int pick = (int)(((r >> 32) * (uint64_t)n) >> 32);
        int sq = nth_set_bit(ok, pick);
        int mc = 0;
        ref_make_move(&me, &you, &mc, 0, sq);
        ++plies;
    }
    *fs = me; *fo = you; *winner = ref_winner(me, you);
    return plies;
}

int64_t ref_random_playouts(int64_t n_games, uint64_t seed, int threads,
                            int64_t winner_hist[3], uint64_t *final_self, uint64_t *final_opp,
                            int32_t *plies_out)
{
    int64_t total = 0, h0 = 0, h1 = 0, h2 = 0;
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(static) reduction(+ : total, h0, h1, h2)
    for (int64_t g = 0; g < n_games; ++g) {
        uint64_t fs, fo; int w;
        uint64_t key = mix64(seed ^ mix64((uint64_t)g));
        int p = one_playout(key, &fs, &fo, &w);
        total += p;
        if (w < 0) ++h0; else if (w == 0) ++h1; else ++h2;
        if (final_self) final_self[g] = fs;
        if (final_opp) final_opp[g] = fo;
        if (plies_out) plies_out[g] = p;
    }
    winner_hist[0] = h0; winner_hist[1] = h1; winner_hist[2] = h2;
    return total;
}

/* ------------------------------------------------------------------------ */
/* hash-net                                                                  */
/* ------------------------------------------------------------------------ */

void ref_hashnet(uint64_t self_b, uint64_t opp_b, float *probs65, float *value)
{
    uint64_t h = mix64(self_b ^ mix64(opp_b + 0x632BE59BD9B4E019ULL));
    for (int i = 0; i < 65; ++i) {
        uint64_t x = mix64(h + (uint64_t)(i + 1) * 0xD1342543DE82EF95ULL);
        uint32_t w = (uint32_t)((x >> 24) & 0xFFFFu);
        probs65[i] = (float)(w + 1u) * (1.0f / 4194304.0f);      /* (w+1)*2^-22 */
    }
    uint64_t xv = mix64(h ^ 0xA5A5A5A5A5A5A5A5ULL);
    int32_t k = (int32_t)((xv >> 16) & 0xFFFFFu) - (1 << 19);
    *value = (float)k * (1.0f / 524288.0f);                      /* k*2^-19 */
}

static void hashnet_eval(uint64_t s, uint64_t o, float *p, float *v, void *user)
{
    (void)user;
    ref_hashnet(s, o, p, v);
}

/* ------------------------------------------------------------------------ */
/* numpy float32 add-reduce order (pairwise, 8 accumulators, n <= 128)       */
/* ------------------------------------------------------------------------ */

static float np_sum_f32(const float *a, int n)
{
    if (n < 8) {
        float acc = 0.0f;
        for (int i = 0; i < n; ++i) acc += a[i];
        return acc;
    }
    float lane[8];
    for (int j = 0; j < 8; ++j) lane[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) lane[j] += a[i + j];
    float acc = ((lane[0] + lane[1]) + (lane[2] + lane[3])) + ((lane[4] + lane[5]) + (lane[6] + lane[7]));
    for (; i < n; ++i) acc += a[i];
    return acc;
}

void ref_expand_priors(const float *probs65, const int *legal, int n_legal, float *priors65)
{
    /* node.py:71-80 */
    float kept[65];
    memset(kept, 0, sizeof kept);
    for (int i = 0; i < n_legal; ++i) kept[legal[i]] = probs65[legal[i]];
    float total = np_sum_f32(kept, 65);
    if (total > 0.0f) {
        for (int i = 0; i < 65; ++i) kept[i] = kept[i] / total;
    } else {
        float u = (float)(1.0 / (double)n_legal);
        for (int i = 0; i < n_legal; ++i) kept[legal[i]] = u;
    }
    memcpy(priors65, kept, sizeof kept);
}

void ref_policy_from_visits(const int32_t *visits, const int *is_child, double temperature, float *policy65)
{
    /* node.py:147-182 */
    memset(policy65, 0, 65 * sizeof(float));
    float cnt[65]; int act[65]; int n = 0;
    for (int a = 0; a < 65; ++a)
        if (is_child[a]) { act[n] = a; cnt[n] = (float)visits[a]; ++n; }
    if (n == 0) return;
    if (temperature == 0.0) {
        int best = 0;                               /* np.argmax: first maximum */
        for (int i = 1; i < n; ++i) if (cnt[i] > cnt[best]) best = i;
        policy65[act[best]] = 1.0f;
        return;
    }
    if (temperature != 1.0) {
        float e = (float)(1.0 / temperature);
        for (int i = 0; i < n; ++i) cnt[i] = powf(cnt[i], e);
    }
    float total = np_sum_f32(cnt, n);
    for (int i = 0; i < n; ++i) policy65[act[i]] = cnt[i] / total;
}

/* ------------------------------------------------------------------------ */
/* MCTS                                                                      */
/* ------------------------------------------------------------------------ */

typedef struct {
    double  value_sum;      /* node.py:40 (Python float) */
    float   prior;          /* node.py:35, np.float32 out of expand */
    int32_t visit_count;    /* node.py:39 */
    int32_t first_child;    /* index of first child node, -1 when a leaf */
    int16_t n_children;
    int16_t action;         /* action that leads here */
} tnode;

typedef struct { tnode *v; int len, cap; } tpool;

static int pool_grow(tpool *p, int extra)
{
    if (p->len + extra <= p->cap) return 0;
    int nc = p->cap ? p->cap * 2 : 1024;
    while (nc < p->len + extra) nc *= 2;
    tnode *nv = (tnode *)realloc(p->v, (size_t)nc * sizeof(tnode));
    if (!nv) return -1;
    p->v = nv; p->cap = nc;
    return 0;
}

/* node.py:62-89 */
static int expand_node(tpool *p, int idx, const float *probs, uint64_t me, uint64_t you)
{
    int legal[65];
    int n = ref_legal_list(me, you, legal);
    float pri[65];
    ref_expand_priors(probs, legal, n, pri);
    if (pool_grow(p, n)) return -1;
    p->v[idx].first_child = p->len;
    p->v[idx].n_children = (int16_t)n;
    for (int i = 0; i < n; ++i) {
        tnode *c = &p->v[p->len++];
        c->value_sum = 0.0; c->prior = pri[legal[i]]; c->visit_count = 0;
        c->first_child = -1; c->n_children = 0; c->action = (int16_t)legal[i];
    }
    return 0;
}

/* node.py:91-126 */
static int pick_child(const tpool *p, int idx, const ref_mcts_cfg *cfg, int parent_n)
{
    const tnode *nd = &p->v[idx];
    double best = -INFINITY; int best_i = -1;
    double root_of_n = sqrt((double)parent_n);
    float c32 = (float)cfg->c_puct;
    for (int i = 0; i < nd->n_children; ++i) {
        const tnode *c = &p->v[nd->first_child + i];
        double q = c->visit_count ? c->value_sum / (double)c->visit_count : 0.0;   /* node.py:51-60 */
        if (cfg->q_canonical) q = -q;
        float cp = c32 * c->prior;                       /* weak-scalar promotion: float32 product */
        double u = (double)cp * root_of_n / (double)(1 + c->visit_count);
        double s = q + u;
        if (s > best) { best = s; best_i = nd->first_child + i; }
    }
    return best_i;
}

/* One game's search state, steppable: select a leaf -> (evaluate elsewhere) -> apply.  This is
 * the lock-step shape of BatchMCTS.search_batch (parallel_self_play.py:121-161). */
typedef struct {
    tpool pool;
    int *trail; int trail_cap;
    uint64_t root_me, root_you;
    uint64_t leaf_me, leaf_you;
    int depth, cur, pending, sims_done, n_evals, deepest;
} gtree;

static int gtree_begin(gtree *g, uint64_t me, uint64_t you, int max_sims)
{
    g->pool.len = 0;
    if (pool_grow(&g->pool, 1)) return -1;
    tnode *root = &g->pool.v[g->pool.len++];
    root->value_sum = 0.0; root->prior = 1.0f; root->visit_count = 0;
    root->first_child = -1; root->n_children = 0; root->action = -1;
    if (g->trail_cap < max_sims + 8) {
        free(g->trail);
        g->trail = (int *)malloc(sizeof(int) * (size_t)(max_sims + 8));
        g->trail_cap = max_sims + 8;
    }
    g->root_me = me; g->root_you = you;
    g->leaf_me = me; g->leaf_you = you;          /* first request: the root itself (mcts.py:74-75) */
    g->depth = 0; g->cur = 0; g->pending = 1; g->sims_done = 0; g->n_evals = 0; g->deepest = 0;
    return 0;
}

/* mcts.py:100-130: descend; a terminal leaf is scored and backed up at once (pending = 0) */
static void gtree_select(gtree *g, const ref_mcts_cfg *cfg)
{
    uint64_t me = g->root_me, you = g->root_you; int mc = 0;
    int cur = 0, depth = 0;
    tpool *pool = &g->pool;
    while (pool->v[cur].n_children > 0) {                        /* mcts.py:117-123 */
        int parent_n = pool->v[cur].visit_count;
        if (cur == 0 && cfg->root_n_sum) parent_n = g->sims_done; /* opt-in: root N = completed sims */
        int nxt = pick_child(pool, cur, cfg, parent_n);
        ref_make_move(&me, &you, &mc, 0, pool->v[nxt].action);
        g->trail[depth++] = nxt;   /* depth <= simulations: every level below the root is an expanded node */
        cur = nxt;
    }
    if (depth > g->deepest) g->deepest = depth;
    g->depth = depth; g->cur = cur; g->leaf_me = me; g->leaf_you = you;
    if (ref_is_terminal(me, you)) {                              /* mcts.py:127-130 */
        double value = (double)ref_winner(me, you);
        for (int i = depth - 1; i >= 0; --i) {                   /* mcts.py:152-168 */
            tnode *c = &pool->v[g->trail[i]];
            c->visit_count += 1; c->value_sum += value; value = -value;
        }
        g->sims_done += 1;
        g->pending = 0;
    } else {
        g->pending = 1;
    }
}

/* mcts.py:133-148: expand the pending leaf and back its value up */
static int gtree_apply(gtree *g, const float *probs, float val)
{
    if (!g->pending) return 0;
    if (expand_node(&g->pool, g->cur, probs, g->leaf_me, g->leaf_you)) return -1;
    g->n_evals += 1;
    if (g->cur != 0) {
        double value = (double)val;
        for (int i = g->depth - 1; i >= 0; --i) {
            tnode *c = &g->pool.v[g->trail[i]];
            c->visit_count += 1; c->value_sum += value; value = -value;
        }
        g->sims_done += 1;
    }
    g->pending = 0;
    return 0;
}

static void gtree_root_stats(const gtree *g, int32_t *visits65, double *q65, int *is_child65, int *n_children)
{
    const tnode *root = &g->pool.v[0];
    for (int a = 0; a < 65; ++a) { if (visits65) visits65[a] = 0; if (q65) q65[a] = 0.0; if (is_child65) is_child65[a] = 0; }
    if (n_children) *n_children = root->n_children;
    for (int i = 0; i < root->n_children; ++i) {
        const tnode *c = &g->pool.v[root->first_child + i];
        if (visits65) visits65[c->action] = c->visit_count;
        if (q65) q65[c->action] = c->visit_count ? c->value_sum / (double)c->visit_count : 0.0;
        if (is_child65) is_child65[c->action] = 1;
    }
}

static void gtree_free(gtree *g) { free(g->pool.v); free(g->trail); memset(g, 0, sizeof *g); }

int ref_mcts_search(uint64_t self_b, uint64_t opp_b, const ref_mcts_cfg *cfg,
                    ref_eval_fn eval, void *user, ref_mcts_result *out)
{
    gtree g; memset(&g, 0, sizeof g);
    float probs[65], val;
    if (gtree_begin(&g, self_b, opp_b, cfg->num_simulations)) return -1;
    eval(self_b, opp_b, probs, &val, user);                      /* mcts.py:74-75 */
    if (gtree_apply(&g, probs, val)) { gtree_free(&g); return -1; }
    for (int s = 0; s < cfg->num_simulations; ++s) {             /* mcts.py:89-92 */
        gtree_select(&g, cfg);
        if (g.pending) {
            eval(g.leaf_me, g.leaf_you, probs, &val, user);
            if (gtree_apply(&g, probs, val)) { gtree_free(&g); return -1; }
        }
    }
    memset(out, 0, sizeof *out);
    gtree_root_stats(&g, out->visits, out->q, out->is_child, &out->n_children);
    out->n_evals = g.n_evals;
    out->max_depth = g.deepest;
    gtree_free(&g);
    return 0;
}

/* ---- lock-step batch of searches with an external (batched) evaluator ---------------------- */
struct ref_batch {
    int n; ref_mcts_cfg cfg; gtree *g; int root_phase;
};

ref_batch *ref_batch_create(int n_games, double c_puct, int max_simulations, int root_n_sum, int q_canonical)
{
    ref_batch *b = (ref_batch *)calloc(1, sizeof *b);
    if (!b) return 0;
    b->n = n_games;
    b->cfg.c_puct = c_puct; b->cfg.num_simulations = max_simulations;
    b->cfg.root_n_sum = root_n_sum; b->cfg.q_canonical = q_canonical;
    b->g = (gtree *)calloc((size_t)n_games, sizeof(gtree));
    return b;
}

void ref_batch_destroy(ref_batch *b)
{
    if (!b) return;
    for (int i = 0; i < b->n; ++i) gtree_free(&b->g[i]);
    free(b->g); free(b);
}

int ref_batch_begin(ref_batch *b, const uint64_t *self_b, const uint64_t *opp_b, int n)
{
    if (n > b->n) return -1;
    for (int i = 0; i < n; ++i)
        if (gtree_begin(&b->g[i], self_b[i], opp_b[i], b->cfg.num_simulations)) return -1;
    for (int i = n; i < b->n; ++i) b->g[i].pending = 0;
    b->root_phase = 1;
    return 0;
}

/* one lock-step: every game selects a leaf (first call after begin: the roots).  Returns the
 * number of leaves that need the evaluator. */
int ref_batch_collect(ref_batch *b, int n, uint64_t *leaf_self, uint64_t *leaf_opp, uint8_t *need)
{
    int cnt = 0;
    if (!b->root_phase) {
#pragma omp parallel for schedule(static) if (n >= 64)
        for (int i = 0; i < n; ++i) gtree_select(&b->g[i], &b->cfg);
    }
    b->root_phase = 0;
    for (int i = 0; i < n; ++i) {
        leaf_self[i] = b->g[i].leaf_me; leaf_opp[i] = b->g[i].leaf_you;
        need[i] = (uint8_t)b->g[i].pending;
        cnt += b->g[i].pending;
    }
    return cnt;
}

int ref_batch_apply(ref_batch *b, int n, const float *probs, const float *value)
{
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (n >= 64)
    for (int i = 0; i < n; ++i) bad |= gtree_apply(&b->g[i], probs + (size_t)i * 65, value[i]) ? 1 : 0;
    return bad ? -1 : 0;
}

void ref_batch_visits(const ref_batch *b, int n, int32_t *visits, int32_t *n_evals)
{
    for (int i = 0; i < n; ++i) {
        gtree_root_stats(&b->g[i], visits + (size_t)i * 65, 0, 0, 0);
        if (n_evals) n_evals[i] = b->g[i].n_evals;
    }
}

int ref_mcts_search_hashnet_batch(const uint64_t *self_b, const uint64_t *opp_b, int64_t n,
                                  double c_puct, int num_simulations, int threads,
                                  int32_t *visits_out, int32_t *n_evals_out)
{
    int failed = 0;
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(| : failed)
    for (int64_t i = 0; i < n; ++i) {
        ref_mcts_cfg cfg = { c_puct, num_simulations, 0, 0 };
        ref_mcts_result r;
        if (ref_mcts_search(self_b[i], opp_b[i], &cfg, hashnet_eval, 0, &r)) { failed |= 1; continue; }
        memcpy(visits_out + i * 65, r.visits, sizeof r.visits);
        if (n_evals_out) n_evals_out[i] = r.n_evals;
    }
    return failed ? -1 : 0;
}
