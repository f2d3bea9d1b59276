// search.cu -- GPU-resident Monte-Carlo tree search, a group of 8 lanes per game (four games per warp).
//
// Restates src/mcts/node.py and src/mcts/mcts.py (and the lock-step BatchMCTS of
// src/train/parallel_self_play.py:80-197) on a structure-of-arrays tree in HBM:
//   select  (node.py:91-126, mcts.py:117-123): lanes score the children in parallel
//           (float64 PUCT with a float32 c_puct*P product, exactly the promotion NumPy >= 2
//           applies to `c_puct * child.prior * np.sqrt(N) / (1 + n)`), group arg-max with the
//           reference's tie rule (first child in ascending action order);
//   expand  (node.py:62-89): masked, renormalised priors in numpy's float32 summation order;
//   backup  (mcts.py:152-168): sign flip per level, root never updated.
// One simulation per game is in flight (K = 1), so results are identical to the reference's
// serial search; games are independent, so there are no atomics on the tree itself.
// Defaults reproduce the reference's behaviour (root N stays 0, child Q un-negated); the
// canonical AlphaZero variants are opt-in flags.
#include <math.h>

#include "bitboard.cuh"
#include "search.cuh"

namespace oth {

constexpr int kWarpsPerBlock = 8;
constexpr int kSearchBlock = kWarpsPerBlock * 32;
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kLanesPerGame = 8;                       // k_tree_select / k_tree_expand: lanes that walk one game
constexpr int kGamesPerWarp = 32 / kLanesPerGame;
constexpr int kGamesPerBlock = kWarpsPerBlock * kGamesPerWarp;

__global__ void __launch_bounds__(kSearchBlock)
k_tree_begin(TreeDev t, const uint64_t* __restrict__ self_b, const uint64_t* __restrict__ opp_b,
             const uint8_t* __restrict__ active, int64_t n, int32_t* __restrict__ build_list)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= t.games) return;
    const bool live = g < n && (active == nullptr || active[g]);
    t.root_self[g] = live ? self_b[g] : 0ULL;
    t.root_opp[g] = live ? opp_b[g] : 0ULL;
    t.active[g] = live ? 1 : 0;
    t.n_nodes[g] = 0; t.n_edges[g] = 0; t.n_evals[g] = 0; t.sims_done[g] = 0; t.path_len[g] = 0; t.root_count[g] = 0;
    t.pending[g] = 0; t.eval_slot[g] = -1;
    t.leaf_self[g] = 0ULL; t.leaf_opp[g] = 0ULL; t.leaf_legal[g] = 0ULL;
    if (g == 0) *t.batch_count = 0;
    if (live && build_list) build_list[atomicAdd(t.act_count, 1)] = (int32_t)g;   // order is irrelevant: games are independent
}

// A pending leaf asks for its evaluation: table hit, or a request that k_tree_assign resolves, or (cache off)
// a slot in the compacted batch right away.  Called by one thread per game.
__device__ __forceinline__ void request_evaluation(const TreeDev& t, int64_t g, uint64_t me, uint64_t you, uint32_t epoch,
                                                   uint32_t gen)
{
    if (t.cache_mask == 0) {
        const int slot = atomicAdd(t.batch_count, 1);          // order is irrelevant: the network is slot-independent
        t.batch_self[slot] = me; t.batch_opp[slot] = you; t.eval_slot[g] = slot;
        t.leaf_src[g] = kSrcSlot;
        atomicAdd(&t.stats[0], 1ULL);
        return;
    }
    const uint32_t h = cache_index(t, me, you);
    t.leaf_h[g] = h;
    const ulonglong2 key = t.c_key[h];
    if (key.x == me && key.y == you && t.c_gen[h] == gen) {
        t.leaf_src[g] = kSrcCache;
        t.c_hit_epoch[h] = epoch;                               // pins the entry for this step
        atomicAdd(&t.stats[1], 1ULL);
    } else {
        t.leaf_src[g] = kSrcMiss;
        atomicMin(&t.c_owner[h], ((unsigned long long)(~epoch) << 32) | (unsigned long long)(uint32_t)g);
    }
}

// First step after begin: every live game asks for its root evaluation (mcts.py:74-75).
__global__ void __launch_bounds__(kSearchBlock) k_tree_root(TreeDev t, int64_t n, uint32_t epoch, uint32_t gen)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    if (!t.active[g]) { t.pending[g] = 0; return; }
    const uint64_t a = t.root_self[g], b = t.root_opp[g];
    t.leaf_self[g] = a; t.leaf_opp[g] = b; t.leaf_legal[g] = legal_moves(a, b);
    t.path_len[g] = 0;
    t.pending[g] = 1;
    request_evaluation(t, g, a, b, epoch, gen);
}

// One descent per searching game.  A game is walked by a GROUP of 8 lanes (4 games per warp): a node has ~9 children
// on average (26 at most in the observed trees), so 8 lanes score them in one or two trips, and four independent
// pointer chases per warp hide each other's latency -- the kernel is bound by dependent 24-byte loads, not by lanes.
__global__ void __launch_bounds__(kSearchBlock)
k_tree_select(TreeDev t, int64_t n, float c32, uint32_t flags, uint32_t epoch, uint32_t gen)
{
    const int lane = threadIdx.x & 31, sub = lane & (kLanesPerGame - 1);
    const int64_t i = (blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5)) * kGamesPerWarp + (lane / kLanesPerGame);
    bool live = i < n;
    const int64_t g = live ? (t.act_list ? (int64_t)t.act_list[i] : i) : 0;
    if (live && !t.active[g]) { if (sub == 0) t.pending[g] = 0; live = false; }
    if (!__any_sync(kFull, live)) return;
    Edge* E = t.edges + g * (int64_t)t.edge_cap;
    int32_t* path = t.path + g * t.path_cap;

    uint64_t me = live ? t.root_self[g] : 0ULL, you = live ? t.root_opp[g] : 0ULL;
    int first = 0, cnt = live ? t.root_count[g] : 0, depth = 0;
    int parent_n = (live && (flags & OTH_FLAG_ROOT_N_SUM)) ? t.sims_done[g] : 0;   // mcts.py:152-172: the root is never updated
    bool descending = live;
    while (__any_sync(kFull, descending)) {
        const double root_of_n = sqrt((double)parent_n);
        double best = -INFINITY;
        int best_e = 0x7FFFFFFF, b_n = 0, b_first = kEdgeLeaf, b_cnt = 0, b_act = 0;
        if (descending) {
            for (int k = sub; k < cnt; k += kLanesPerGame) {
                const Edge ed = E[first + k];                                   // one 24-byte record per child
                double q = ed.n ? ed.w / (double)ed.n : 0.0;                    // node.py:51-60
                if (flags & OTH_FLAG_Q_CANONICAL) q = -q;
                const float cp = __fmul_rn(c32, ed.p);                          // float32 product (weak Python scalar)
                const double u = __ddiv_rn(__dmul_rn((double)cp, root_of_n), (double)(1 + ed.n));
                const double sc = __dadd_rn(q, u);
                if (sc > best) {                                                // strict >: first maximum wins
                    best = sc; best_e = first + k; b_n = ed.n; b_first = ed.child_first; b_cnt = ed.child_count; b_act = ed.action;
                }
            }
        }
        int win_lane = sub;
#pragma unroll
        for (int o = kLanesPerGame / 2; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(kFull, best, o, kLanesPerGame);
            const int oe = __shfl_xor_sync(kFull, best_e, o, kLanesPerGame);
            const int ol = __shfl_xor_sync(kFull, win_lane, o, kLanesPerGame);
            if (os > best || (os == best && oe < best_e)) { best = os; best_e = oe; win_lane = ol; }
        }
        const int w_n = __shfl_sync(kFull, b_n, win_lane, kLanesPerGame);
        const int w_first = __shfl_sync(kFull, b_first, win_lane, kLanesPerGame);
        const int w_cnt = __shfl_sync(kFull, b_cnt, win_lane, kLanesPerGame);
        const int w_act = __shfl_sync(kFull, b_act, win_lane, kLanesPerGame);
        if (descending) {
            parent_n = w_n;
            cnt = w_cnt;
            if (sub == 0) path[depth] = best_e;
            ++depth;
            apply_known_legal(me, you, w_act);                                  // mcts.py:122
            if (cnt == 0 || depth >= t.path_cap) descending = false;            // child not expanded: this is the leaf
            else first = w_first;
        }
    }
    if (!live) return;
    const uint64_t lg = legal_moves(me, you);
    const bool terminal = lg == 0 && legal_moves(you, me) == 0;           // mcts.py:127
    if (sub == 0) {
        if (terminal) {
            double v = (double)winner(me, you);                            // mcts.py:129-130
            for (int d = depth - 1; d >= 0; --d) {                         // mcts.py:152-168
                Edge* ed = E + path[d];
                ed->n += 1;
                ed->w += v;
                v = -v;
            }
            t.sims_done[g] += 1;
            t.pending[g] = 0;
        } else {
            t.leaf_self[g] = me; t.leaf_opp[g] = you; t.leaf_legal[g] = lg;
            t.path_len[g] = depth;
            t.pending[g] = 1;
            request_evaluation(t, g, me, you, epoch, gen);
        }
    }
}

// Resolve the misses of this step: the elected game of every table entry gets a batch slot (and will insert),
// games with the same position share it, colliding positions get a slot of their own.
__global__ void __launch_bounds__(kSearchBlock) k_tree_assign(TreeDev t, int64_t n)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t g = t.act_list ? (int64_t)t.act_list[i] : i;
    if (!t.pending[g] || t.leaf_src[g] != kSrcMiss) return;
    const uint32_t h = t.leaf_h[g];
    const uint32_t o = (uint32_t)(t.c_owner[h] & 0xFFFFFFFFULL);
    const uint64_t me = t.leaf_self[g], you = t.leaf_opp[g];
    if (o != (uint32_t)g && t.leaf_self[o] == me && t.leaf_opp[o] == you) {
        t.leaf_src[g] = kSrcDedup;
        t.dedup_of[g] = (int32_t)o;
        atomicAdd(&t.stats[2], 1ULL);
        return;
    }
    const int slot = atomicAdd(t.batch_count, 1);
    t.batch_self[slot] = me; t.batch_opp[slot] = you; t.eval_slot[g] = slot;
    t.leaf_src[g] = (o == (uint32_t)g) ? kSrcOwner : kSrcSlot;
    atomicAdd(&t.stats[0], 1ULL);
    if (o != (uint32_t)g) atomicAdd(&t.stats[3], 1ULL);        // different position on the same entry
}

// Expand the pending leaf of every game with the evaluator's output and back the value up (8 lanes per game).
__global__ void __launch_bounds__(kSearchBlock)
k_tree_expand(TreeDev t, int64_t n, const float* __restrict__ policy, const float* __restrict__ value, int policy_is_raw,
              int by_slot, uint32_t epoch, uint32_t gen)
{
    __shared__ float s_pri[kWarpsPerBlock * kGamesPerWarp][68];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & (kLanesPerGame - 1), grp = lane / kLanesPerGame;
    const int64_t i = (blockIdx.x * (int64_t)kWarpsPerBlock + w) * kGamesPerWarp + grp;
    if (blockIdx.x == 0 && threadIdx.x == 0) *t.batch_count = 0;      // the evaluator has consumed the batch
    bool live = i < n;
    const int64_t g = live ? (t.act_list ? (int64_t)t.act_list[i] : i) : 0;
    if (live && !t.pending[g]) live = false;
    if (!__any_sync(kFull, live)) return;
    const uint64_t lg = live ? t.leaf_legal[g] : 0ULL;
    // where this leaf's evaluation is: the batch slot (own or shared), the table, or (external / hash-net) slot = game
    const uint8_t how = (live && by_slot) ? t.leaf_src[g] : kSrcSlot;
    int64_t src = g;
    if (live && by_slot) src = how == kSrcDedup ? (int64_t)t.eval_slot[t.dedup_of[g]] : (int64_t)t.eval_slot[g];
    const float* prow = how == kSrcCache ? t.c_priors + (size_t)t.leaf_h[g] * 68 : policy + src * 65;
    float leaf_value = 0.f;
    float* pri = s_pri[w * kGamesPerWarp + grp];
    if (live) {
        leaf_value = how == kSrcCache ? t.c_value[t.leaf_h[g]] : value[src];
        for (int j = sub; j < 65; j += kLanesPerGame) pri[j] = prow[j];
    }
    __syncwarp();
    if (live && policy_is_raw && sub == 0) mask_and_renormalise(pri, lg);  // node.py:71-80
    __syncwarp();
    const int depth = live ? t.path_len[g] : 0;
    const int cnt = lg ? popc64(lg) : 1;                                   // [64] = forced pass (bitboard.pyx:176-178)
    const int first = live ? t.n_edges[g] : 0;
    if (live && first + cnt > t.edge_cap) {
        if (sub == 0) { atomicExch(t.error_flag, 1); t.pending[g] = 0; }
        live = false;
    }
    if (!live) return;                                                     // no warp-wide synchronisation below
    Edge* E = t.edges + g * (int64_t)t.edge_cap;
    for (int k = sub; k < cnt; k += kLanesPerGame) {
        Edge ed;
        ed.w = 0.0; ed.n = 0; ed.child_first = kEdgeLeaf; ed.child_count = 0; ed.pad = 0;
        const int action = lg ? nth_set_bit(lg, k) : kPass;
        ed.p = pri[action]; ed.action = (uint8_t)action;
        E[first + k] = ed;
    }
    if (sub == 0) {
        t.n_nodes[g] += 1;
        t.n_edges[g] = first + cnt;
        t.n_evals[g] += 1;
        if (depth == 0) {
            t.root_count[g] = cnt;                                         // the root's children are edges [0, cnt)
        } else {
            const int32_t* path = t.path + g * t.path_cap;
            Edge* leaf = E + path[depth - 1];
            leaf->child_first = first; leaf->child_count = (uint8_t)cnt;
            double v = (double)leaf_value;                                 // value.item(), mcts.py:144
            for (int d = depth - 1; d >= 0; --d) {
                Edge* ed = E + path[d];
                ed->n += 1;
                ed->w += v;
                v = -v;
            }
            t.sims_done[g] += 1;
        }
        t.pending[g] = 0;
    }
    // the elected evaluator publishes the result, unless somebody read the old entry in this very step
    if (how == kSrcOwner) {
        const uint32_t h = t.leaf_h[g];
        if (t.c_hit_epoch[h] != epoch) {
            float* dst = t.c_priors + (size_t)h * 68;
            for (int j = sub; j < 65; j += kLanesPerGame) dst[j] = pri[j];
            if (sub == 0) {
                t.c_key[h] = make_ulonglong2(t.leaf_self[g], t.leaf_opp[g]);
                t.c_value[h] = leaf_value;
                t.c_gen[h] = gen;
            }
        }
    }
}

// Built-in integer test evaluator (same definition as oracle/ref_rules.c ref_hashnet).
__global__ void __launch_bounds__(kSearchBlock) k_hashnet(TreeDev t, int64_t n)
{
    const int64_t g = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= n || !t.pending[g]) return;
    const uint64_t a = t.leaf_self[g], b = t.leaf_opp[g];
    const uint64_t h = mix64(a ^ mix64(b + 0x632BE59BD9B4E019ULL));
    for (int i = lane; i < 65; i += 32) {
        const uint64_t x = mix64(h + (uint64_t)(i + 1) * 0xD1342543DE82EF95ULL);
        const uint32_t wv = (uint32_t)((x >> 24) & 0xFFFFu);
        t.eval_policy[g * 65 + i] = (float)(wv + 1u) * (1.0f / 4194304.0f);
    }
    if (lane == 0) {
        const uint64_t xv = mix64(h ^ 0xA5A5A5A5A5A5A5A5ULL);
        const int32_t k = (int32_t)((xv >> 16) & 0xFFFFFu) - (1 << 19);
        t.eval_value[g] = (float)k * (1.0f / 524288.0f);
    }
}

// ---- Dirichlet noise on the root priors (mcts.py:210-228) -----------------------------------
// Only has an effect when the root's visit count takes part in PUCT (OTH_FLAG_ROOT_N_SUM); with
// the reference's default (root N == 0) the exploration term is identically zero and the noise
// cannot change any result, so the host skips this kernel in that mode.
__device__ __forceinline__ double u01(uint64_t& s)
{
    s = mix64(s);
    return ((double)(s >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ double gamma_sample(double alpha, uint64_t& s)
{
    // Marsaglia-Tsang; alpha < 1 via the boost Gamma(alpha+1) * U^(1/alpha)
    const double boost = alpha < 1.0 ? pow(u01(s), 1.0 / alpha) : 1.0;
    const double d = (alpha < 1.0 ? alpha + 1.0 : alpha) - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 64; ++it) {
        const double u1 = u01(s), u2 = u01(s);
        const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = u01(s);
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
    }
    return boost * d;
}
__global__ void __launch_bounds__(kSearchBlock)
k_root_noise(TreeDev t, int64_t n, double alpha, double eps, uint64_t seed, const int32_t* __restrict__ salt)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n || !t.active[g] || t.root_count[g] < 1) return;
    Edge* E = t.edges + g * (int64_t)t.edge_cap;
    const int cnt = t.root_count[g];
    uint64_t s = mix64(seed ^ mix64((uint64_t)g * 0x9E3779B97F4A7C15ULL + (salt ? (uint64_t)salt[g] : 0ULL)));
    double noise[64];
    double total = 0.0;
    for (int k = 0; k < cnt && k < 64; ++k) { noise[k] = gamma_sample(alpha, s); total += noise[k]; }
    for (int k = 0; k < cnt && k < 64; ++k)
        E[k].p = (float)((1.0 - eps) * (double)E[k].p + eps * (noise[k] / total));
}

__global__ void __launch_bounds__(kSearchBlock)
k_tree_results(TreeDev t, int64_t n, int32_t* __restrict__ visits, double* __restrict__ q, int32_t* __restrict__ n_evals)
{
    const int64_t g = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= n) return;
    for (int j = lane; j < 65; j += 32) {
        if (visits) visits[g * 65 + j] = 0;
        if (q) q[g * 65 + j] = 0.0;
    }
    __syncwarp();
    if (lane == 0 && n_evals) n_evals[g] = t.n_evals[g];
    if (!t.active[g] || t.root_count[g] < 1) return;
    const Edge* E = t.edges + g * (int64_t)t.edge_cap;
    const int cnt = t.root_count[g];
    for (int k = lane; k < cnt; k += 32) {
        const Edge ed = E[k];
        if (visits) visits[g * 65 + ed.action] = ed.n;
        if (q) q[g * 65 + ed.action] = ed.n ? ed.w / (double)ed.n : 0.0;
    }
}

// get_policy_distribution (node.py:147-182)
__global__ void __launch_bounds__(kSearchBlock) k_tree_policy(TreeDev t, int64_t n, double temperature, float* __restrict__ out)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    float* p = out + g * 65;
    for (int j = 0; j < 65; ++j) p[j] = 0.f;
    if (!t.active[g] || t.root_count[g] < 1) return;
    const Edge* E = t.edges + g * (int64_t)t.edge_cap;
    const int cnt = t.root_count[g];
    if (temperature == 0.0) {
        int best = 0;                                    // np.argmax: first maximum (node.py:171-174)
        for (int k = 1; k < cnt; ++k) if (E[k].n > E[best].n) best = k;
        p[E[best].action] = 1.0f;
        return;
    }
    // counts ** (1/T) / sum, float32 (node.py:177-180).  For T == 1 the counts are small integers,
    // every partial sum is exact, so numpy's pairwise order cannot matter.
    const float ex = (float)(1.0 / temperature);
    float total = 0.f;
    for (int k = 0; k < cnt; ++k) {
        const float c = (float)E[k].n;
        total += (temperature == 1.0) ? c : powf(c, ex);
    }
    for (int k = 0; k < cnt; ++k) {
        const float c = (float)E[k].n;
        p[E[k].action] = __fdiv_rn((temperature == 1.0) ? c : powf(c, ex), total);
    }
}

// ---- host side ----------------------------------------------------------------------------------

template <class T>
static int dev_alloc(std::vector<void*>& owned, T** p, size_t count)
{
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, count * sizeof(T) + 16);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e)); return OTH_ERR_CUDA; }
    owned.push_back(d);
    *p = (T*)d;
    return OTH_OK;
}

int SearchHost::allocate(oth_ctx* c, int64_t games, int sims)
{
    ctx = c; max_games = games; max_sims = sims;
    OTH_CHECK_CUDA(cudaSetDevice(c->device));
    t.games = games;
    t.node_cap = sims + 2;
    t.edge_cap = t.node_cap * kEdgesPerNodeBudget;
    if (t.edge_cap < 128) t.edge_cap = 128;
    t.path_cap = sims + 2;
    const size_t G = (size_t)games;
    int rc = 0;
#define A(ptr, cnt) if ((rc = dev_alloc(allocs, &ptr, (cnt)))) return rc
    A(t.root_self, G); A(t.root_opp, G);
    A(t.n_nodes, G); A(t.n_edges, G); A(t.n_evals, G); A(t.sims_done, G); A(t.path_len, G);
    A(t.pending, G); A(t.active, G);
    A(t.leaf_self, G); A(t.leaf_opp, G); A(t.leaf_legal, G);
    A(t.batch_self, G); A(t.batch_opp, G); A(t.eval_slot, G); A(t.batch_count, 1);
    A(t.path, G * t.path_cap);
    A(t.root_count, G); A(t.edges, G * t.edge_cap);
    A(t.eval_policy, G * 65); A(t.eval_value, G);
    A(t.error_flag, 1);
    A(t.leaf_h, G); A(t.leaf_src, G); A(t.dedup_of, G); A(t.stats, 4);
    A(act_list_buf, G); A(t.act_count, 1);
#undef A
    t.cache_mask = 0;
    t.act_list = nullptr;
    OTH_CHECK_CUDA(cudaMemsetAsync(t.stats, 0, 4 * sizeof(unsigned long long), c->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(t.leaf_src, 0, G, c->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(t.error_flag, 0, sizeof(int32_t), c->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(t.eval_policy, 0, G * 65 * sizeof(float), c->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(t.eval_value, 0, G * sizeof(float), c->stream));
    return OTH_OK;
}

void SearchHost::release()
{
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
}

static inline int warp_grid(int64_t n) { return n > 0 ? (int)((n + kWarpsPerBlock - 1) / kWarpsPerBlock) : 1; }
static inline int group_grid(int64_t n) { return n > 0 ? (int)((n + kGamesPerBlock - 1) / kGamesPerBlock) : 1; }
static inline int thread_grid(int64_t n) { return n > 0 ? (int)((n + kSearchBlock - 1) / kSearchBlock) : 1; }

#define K_CHECK()                                 \
    do {                                          \
        ctx->launches++;                          \
        OTH_CHECK_CUDA(cudaGetLastError());       \
    } while (0)

int SearchHost::begin(const uint64_t* d_self, const uint64_t* d_opp, const uint8_t* d_active, int64_t n_games)
{
    OTH_REQUIRE(n_games >= 0 && n_games <= max_games, OTH_ERR_ARG, "search: %lld games exceed the capacity %lld",
                (long long)n_games, (long long)max_games);
    n = n_games;
    // with an activity mask (self-play: one search per group of identical roots) the searching games are compacted
    const bool listed = d_active != nullptr;
    if (listed) OTH_CHECK_CUDA(cudaMemsetAsync(t.act_count, 0, sizeof(int32_t), ctx->stream));
    k_tree_begin<<<thread_grid(max_games), kSearchBlock, 0, ctx->stream>>>(t, d_self, d_opp, d_active, n, listed ? act_list_buf : nullptr);
    K_CHECK();
    t.act_list = nullptr; n_act = n;
    if (listed) {
        int32_t cnt = 0;
        OTH_CHECK_CUDA(cudaMemcpyAsync(&cnt, t.act_count, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
        OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        t.act_list = act_list_buf; n_act = cnt;
    }
    begun = true; awaiting_apply = false; root_pending = true;
    return OTH_OK;
}

// The table is only consulted when the built-in network evaluates the leaves (use_cache).
static inline TreeDev tree_view(const TreeDev& t, bool use_cache)
{
    TreeDev v = t;
    if (!use_cache) v.cache_mask = 0;
    return v;
}

int SearchHost::root(bool use_cache)
{
    TimedLaunch timed(ctx, 1);
    ++epoch;
    k_tree_root<<<thread_grid(n), kSearchBlock, 0, ctx->stream>>>(tree_view(t, use_cache), n, epoch, generation);
    K_CHECK();
    return OTH_OK;
}

int SearchHost::select(bool use_cache)
{
    TimedLaunch timed(ctx, 1);
    ++epoch;
    k_tree_select<<<group_grid(n_act), kSearchBlock, 0, ctx->stream>>>(tree_view(t, use_cache), n_act, (float)c_puct, flags, epoch, generation);
    K_CHECK();
    return OTH_OK;
}

int SearchHost::assign()
{
    if (!t.cache_mask) return OTH_OK;
    TimedLaunch timed(ctx, 1);
    k_tree_assign<<<thread_grid(n_act), kSearchBlock, 0, ctx->stream>>>(t, n_act);
    K_CHECK();
    return OTH_OK;
}

int SearchHost::enable_cache(uint64_t capacity)
{
    if (cache_on) return OTH_OK;
    OTH_REQUIRE(capacity >= 1024 && (capacity & (capacity - 1)) == 0, OTH_ERR_ARG, "cache capacity must be a power of two >= 1024");
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    int rc = 0;
    if ((rc = dev_alloc(allocs, &t.c_key, capacity))) return rc;
    if ((rc = dev_alloc(allocs, &t.c_gen, capacity))) return rc;
    if ((rc = dev_alloc(allocs, &t.c_value, capacity))) return rc;
    if ((rc = dev_alloc(allocs, &t.c_priors, capacity * 68))) return rc;
    if ((rc = dev_alloc(allocs, &t.c_owner, capacity))) return rc;
    if ((rc = dev_alloc(allocs, &t.c_hit_epoch, capacity))) return rc;
    OTH_CHECK_CUDA(cudaMemsetAsync(t.c_gen, 0, capacity * sizeof(uint32_t), ctx->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(t.c_hit_epoch, 0, capacity * sizeof(uint32_t), ctx->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(t.c_owner, 0xFF, capacity * sizeof(unsigned long long), ctx->stream));
    OTH_CHECK_CUDA(cudaMemsetAsync(t.c_key, 0, capacity * sizeof(ulonglong2), ctx->stream));
    t.cache_mask = capacity - 1;
    cache_on = true;
    return OTH_OK;
}

int SearchHost::read_stats(unsigned long long out[4], bool reset)
{
    OTH_CHECK_CUDA(cudaMemcpyAsync(out, t.stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) OTH_CHECK_CUDA(cudaMemsetAsync(t.stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return OTH_OK;
}

int SearchHost::expand(const float* d_policy, const float* d_value, bool policy_is_raw, bool by_slot)
{
    TimedLaunch timed(ctx, 1);
    k_tree_expand<<<group_grid(n_act), kSearchBlock, 0, ctx->stream>>>(t, n_act, d_policy, d_value, policy_is_raw ? 1 : 0, by_slot ? 1 : 0,
                                                                   epoch, generation);
    K_CHECK();
    return OTH_OK;
}

int SearchHost::evaluate(NetHost* net)
{
    if (flags & OTH_FLAG_EVAL_HASHNET) {
        TimedLaunch timed(ctx, 1);
        k_hashnet<<<warp_grid(n), kSearchBlock, 0, ctx->stream>>>(t, n);
        K_CHECK();
        return OTH_OK;
    }
    OTH_REQUIRE(net && net->loaded, OTH_ERR_STATE, "search: no network (or weights not loaded) and OTH_FLAG_EVAL_HASHNET not set");
    OTH_REQUIRE(net->ctx == ctx, OTH_ERR_ARG, "search: network belongs to a different context");
    net->evals += (uint64_t)n_act;
    // upper bound of the compacted batch = games that search this move (the device count in t.batch_count can only be
    // smaller); a small bound sends the launch to the latency shape of the network without a read-back
    return net_forward_device(net, t.batch_self, t.batch_opp, n_act, t.eval_policy, t.eval_value, kOutPriors, t.batch_count);
}

int SearchHost::run(NetHost* net, int sims, bool add_noise, uint64_t seed)
{
    OTH_REQUIRE(begun, OTH_ERR_STATE, "oth_search_run: call oth_search_begin first");
    OTH_REQUIRE(sims >= 0 && sims <= max_sims, OTH_ERR_ARG, "oth_search_run: %d simulations exceed max_simulations %d", sims, max_sims);
    if (n == 0) return OTH_OK;
    const bool raw = (flags & OTH_FLAG_EVAL_HASHNET) != 0;   // the hash-net emits unmasked pseudo-probabilities
    int rc;
    const bool use_cache = !raw && cache_on && (flags & OTH_FLAG_EVAL_CACHE);
    if ((rc = root(use_cache))) return rc;
    if (use_cache && (rc = assign())) return rc;
    if ((rc = evaluate(net))) return rc;
    if ((rc = expand(t.eval_policy, t.eval_value, raw, !raw))) return rc;
    if (add_noise && (flags & OTH_FLAG_ROOT_N_SUM)) {
        k_root_noise<<<thread_grid(n), kSearchBlock, 0, ctx->stream>>>(t, n, dir_alpha, dir_eps, seed, nullptr);
        K_CHECK();
    }
    for (int s = 0; s < sims; ++s) {
        if ((rc = select(use_cache))) return rc;
        if (use_cache && (rc = assign())) return rc;
        if ((rc = evaluate(net))) return rc;
        if ((rc = expand(t.eval_policy, t.eval_value, raw, !raw))) return rc;
    }
    awaiting_apply = false;
    return OTH_OK;
}

int SearchHost::check_overflow()
{
    int32_t flag = 0;
    OTH_CHECK_CUDA(cudaMemcpyAsync(&flag, t.error_flag, sizeof flag, cudaMemcpyDeviceToHost, ctx->stream));
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    OTH_REQUIRE(flag == 0, OTH_ERR_CAPACITY, "search: a tree pool overflowed (edge_cap %d per game)", t.edge_cap);
    return OTH_OK;
}

}  // namespace oth

using namespace oth;


extern "C" {

int oth_search_create(oth_ctx* ctx, int64_t max_games, int max_simulations, oth_search** out)
{
    OTH_REQUIRE(ctx && out, OTH_ERR_ARG, "oth_search_create: NULL argument");
    OTH_REQUIRE(max_games > 0 && max_games <= (1 << 22), OTH_ERR_ARG, "oth_search_create: max_games %lld out of range", (long long)max_games);
    OTH_REQUIRE(max_simulations >= 0 && max_simulations <= 4096, OTH_ERR_ARG, "oth_search_create: max_simulations %d out of range 0..4096", max_simulations);
    oth_search* s = new oth_search();
    int rc = s->allocate(ctx, max_games, max_simulations);
    if (rc) { s->release(); delete s; return rc; }
    *out = s;
    return OTH_OK;
}

int oth_search_destroy(oth_search* s)
{
    if (!s) return OTH_OK;
    s->release();
    delete s;
    return OTH_OK;
}

int oth_search_configure(oth_search* s, double c_puct, double dirichlet_alpha, double dirichlet_epsilon, uint32_t flags)
{
    OTH_REQUIRE(s, OTH_ERR_ARG, "oth_search_configure: NULL handle");
    s->c_puct = c_puct; s->dir_alpha = dirichlet_alpha; s->dir_eps = dirichlet_epsilon; s->flags = flags;
    if ((flags & OTH_FLAG_EVAL_CACHE) && !s->cache_on) {
        uint64_t want = (uint64_t)s->max_games * (uint64_t)(s->max_sims + 1) * 2, cap = 1 << 16;
        while (cap < want && cap < (1ULL << 24)) cap <<= 1;
        int rc = s->enable_cache(cap);
        if (rc) return rc;
    }
    return OTH_OK;
}

int oth_search_stats(oth_search* s, uint64_t* out4)
{
    OTH_REQUIRE(s && out4, OTH_ERR_ARG, "oth_search_stats: NULL argument");
    OTH_CHECK_CUDA(cudaSetDevice(s->ctx->device));
    unsigned long long v[4];
    int rc = s->read_stats(v, true);
    for (int i = 0; i < 4; ++i) out4[i] = v[i];
    return rc;
}

int oth_search_invalidate_cache(oth_search* s)
{
    OTH_REQUIRE(s, OTH_ERR_ARG, "oth_search_invalidate_cache: NULL handle");
    s->invalidate_cache();
    return OTH_OK;
}

int oth_search_begin(oth_search* s, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, int mem)
{
    OTH_REQUIRE(s && (n == 0 || (self_b && opp_b)), OTH_ERR_ARG, "oth_search_begin: NULL argument");
    OTH_CHECK_CUDA(cudaSetDevice(s->ctx->device));
    Staged st(s->ctx, mem);
    const uint64_t* a = st.in(self_b, n); const uint64_t* b = st.in(opp_b, n);
    if (st.failed) return OTH_ERR_CUDA;
    int rc = s->begin(a, b, nullptr, n);
    if (rc) return rc;
    return st.finish();
}

int oth_search_collect(oth_search* s, uint64_t* leaf_self, uint64_t* leaf_opp, uint8_t* need_eval, int mem)
{
    OTH_REQUIRE(s && leaf_self && leaf_opp && need_eval, OTH_ERR_ARG, "oth_search_collect: NULL argument");
    OTH_REQUIRE(s->begun, OTH_ERR_STATE, "oth_search_collect: call oth_search_begin first");
    OTH_REQUIRE(!s->awaiting_apply, OTH_ERR_STATE, "oth_search_collect: previous leaves were not applied");
    oth_ctx* ctx = s->ctx;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = s->n;
    if (n == 0) return OTH_OK;
    if (s->root_pending) {                     // the first collect after begin asks for the roots (mcts.py:74-75)
        int rc = s->root(false);
        if (rc) return rc;
        s->root_pending = false;
    } else {
        int rc = s->select(false);
        if (rc) return rc;
    }
    const cudaMemcpyKind kind = mem == OTH_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    OTH_CHECK_CUDA(cudaMemcpyAsync(leaf_self, s->t.leaf_self, n * 8, kind, ctx->stream));
    OTH_CHECK_CUDA(cudaMemcpyAsync(leaf_opp, s->t.leaf_opp, n * 8, kind, ctx->stream));
    OTH_CHECK_CUDA(cudaMemcpyAsync(need_eval, s->t.pending, n, kind, ctx->stream));
    if (mem == OTH_MEM_HOST) OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    s->awaiting_apply = true;
    return OTH_OK;
}

int oth_search_apply(oth_search* s, const float* probs, const float* value, int mem)
{
    OTH_REQUIRE(s && probs && value, OTH_ERR_ARG, "oth_search_apply: NULL argument");
    OTH_REQUIRE(s->awaiting_apply, OTH_ERR_STATE, "oth_search_apply: nothing collected");
    OTH_CHECK_CUDA(cudaSetDevice(s->ctx->device));
    const int64_t n = s->n;
    Staged st(s->ctx, mem);
    const float* p = st.in(probs, n * 65); const float* v = st.in(value, n);
    if (st.failed) return OTH_ERR_CUDA;
    int rc = s->expand(p, v, true, false);
    if (rc) return rc;
    s->awaiting_apply = false;
    return st.finish();
}

int oth_search_run(oth_search* s, oth_net* net, int num_simulations, int add_dirichlet_noise, uint64_t seed)
{
    OTH_REQUIRE(s, OTH_ERR_ARG, "oth_search_run: NULL handle");
    OTH_CHECK_CUDA(cudaSetDevice(s->ctx->device));
    s->root_pending = false;
    int rc = s->run(net, num_simulations, add_dirichlet_noise != 0, seed);
    if (rc) return rc;
    return s->check_overflow();
}

int oth_search_results(oth_search* s, int32_t* visits, double* q, int32_t* n_evals, int mem)
{
    OTH_REQUIRE(s, OTH_ERR_ARG, "oth_search_results: NULL handle");
    OTH_REQUIRE(s->begun, OTH_ERR_STATE, "oth_search_results: no search");
    oth_ctx* ctx = s->ctx;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = s->n;
    if (n == 0) return OTH_OK;
    Staged st(ctx, mem);
    int32_t* dv = st.out(visits, n * 65); double* dq = st.out(q, n * 65); int32_t* de = st.out(n_evals, n);
    if (st.failed) return OTH_ERR_CUDA;
    k_tree_results<<<warp_grid(n), kSearchBlock, 0, ctx->stream>>>(s->t, n, dv, dq, de);
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    int rc = st.finish();
    if (rc) return rc;
    return mem == OTH_MEM_HOST ? s->check_overflow() : OTH_OK;
}

int oth_search_policy(oth_search* s, double temperature, float* policy_out, int mem)
{
    OTH_REQUIRE(s && policy_out, OTH_ERR_ARG, "oth_search_policy: NULL argument");
    OTH_REQUIRE(s->begun, OTH_ERR_STATE, "oth_search_policy: no search");
    OTH_REQUIRE(temperature >= 0.0, OTH_ERR_ARG, "oth_search_policy: negative temperature");
    oth_ctx* ctx = s->ctx;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = s->n;
    if (n == 0) return OTH_OK;
    Staged st(ctx, mem);
    float* dp = st.out(policy_out, n * 65);
    if (st.failed) return OTH_ERR_CUDA;
    k_tree_policy<<<thread_grid(n), kSearchBlock, 0, ctx->stream>>>(s->t, n, temperature, dp);
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    return st.finish();
}

}  // extern "C"
