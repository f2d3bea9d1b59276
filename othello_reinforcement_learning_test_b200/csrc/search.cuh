// search.cuh -- device-resident SoA Monte-Carlo tree, shared by search.cu and selfplay.cu.
#pragma once
#include "common.cuh"
#include "bitboard.cuh"
#include "net_host.cuh"

namespace oth {

constexpr int kEdgesPerNodeBudget = 40;   // pool sizing only; a node may have up to 64 children

// An "edge" carries what the reference keeps on the child MCTSNode (node.py:28-45): prior P (float32),
// visit_count N, value_sum W (float64) -- plus where that child's own edges live, so a descent needs ONE
// dependent load per level (the 24-byte records of a node's children are contiguous).
struct __align__(8) Edge {
    double w;             // value_sum
    int32_t n;            // visit_count
    float p;              // prior
    int32_t child_first;  // first edge of the child node; kEdgeLeaf = not expanded, kEdgePending = being expanded (waves)
    uint8_t child_count;  // number of edges of the child node (0 while not expanded)
    uint8_t action;       // 0..63 square, 64 pass
    uint16_t pad;
};
static_assert(sizeof(Edge) == 24, "Edge record layout");
constexpr int32_t kEdgeLeaf = -1, kEdgePending = -2;

// All arrays are device memory.  Per game g the edges [g*edge_cap .. +edge_cap) are allocated CSR-style in
// expansion order; the root's children are edges [0, root_count).
struct TreeDev {
    int64_t games;            // capacity (slots)
    int node_cap, edge_cap, path_cap;
    // per game
    uint64_t *root_self, *root_opp;
    int32_t *n_nodes, *n_edges, *n_evals, *sims_done, *path_len;
    uint8_t *pending, *active;
    uint64_t *leaf_self, *leaf_opp, *leaf_legal;   // pending leaf of every game
    // compacted evaluation batch: only leaves that really need the network (no terminal leaves, no idle slots)
    uint64_t *batch_self, *batch_opp;
    int32_t *eval_slot;       // [games] index into the batch / eval_policy / eval_value, -1 = none
    int32_t *batch_count;     // device-side counter, reset by k_tree_expand
    // position-keyed evaluation cache + same-step dedup (optional; result-transparent because the
    // network's output does not depend on where in a batch a position sits).  Direct-mapped on a hash
    // of (self, opp); entries carry the masked priors and the value exactly as the network wrote them.
    uint64_t cache_mask;             // capacity - 1; 0 = cache off
    ulonglong2* c_key;               // (self, opp)
    uint32_t* c_gen;                 // entry valid iff == current generation (bumped when the weights change)
    float* c_value;
    float* c_priors;                 // [capacity][68]
    unsigned long long* c_owner;     // (~epoch << 32 | game): atomicMin elects one evaluator per entry and step
    uint32_t* c_hit_epoch;           // step in which the entry was last read (no overwrite in that step)
    uint32_t* leaf_h;                // [games] table index of the pending leaf
    uint8_t* leaf_src;               // [games] kSrc*
    int32_t* dedup_of;               // [games] game whose evaluation this leaf shares
    unsigned long long* stats;       // [0] network positions, [1] cache hits, [2] same-step duplicates, [3] hash collisions
    int32_t* path;            // [games][path_cap] edge indices of the pending simulation
    int32_t* root_count;      // [games] number of root children (0 = root not expanded yet)
    Edge* edges;              // [games][edge_cap]
    // evaluator outputs for the batch
    float* eval_policy;       // [games][65]
    float* eval_value;        // [games]
    int32_t* error_flag;      // != 0: a pool overflowed
    // compact list of the games that really search this move (self-play shares searches among identical roots, so
    // most slots idle): the per-simulation kernels walk it instead of all slots.  nullptr = every game (identity).
    const int32_t* act_list;
    int32_t* act_count;
};

// table index of a position in the evaluation cache
__device__ __forceinline__ uint32_t cache_index(const TreeDev& t, uint64_t me, uint64_t you)
{
    return (uint32_t)(mix64(me ^ mix64(you + 0x9FB21C651E98DF25ULL)) & t.cache_mask);
}

constexpr uint8_t kSrcSlot = 0;      // own slot in the evaluation batch (cache off, or colliding entry: no insert)
constexpr uint8_t kSrcOwner = 1;     // own slot, and inserts the result into the table
constexpr uint8_t kSrcCache = 2;     // table hit
constexpr uint8_t kSrcDedup = 3;     // same position as another game's leaf in this step
constexpr uint8_t kSrcMiss = 4;      // (between select and assign) wants an evaluation

struct SearchHost {
    oth_ctx* ctx = nullptr;
    int64_t max_games = 0, n = 0;
    int64_t n_act = 0;                    // entries of t.act_list (== n when the list is off)
    int32_t* act_list_buf = nullptr;
    int max_sims = 0;
    double c_puct = 1.0, dir_alpha = 0.3, dir_eps = 0.25;
    uint32_t flags = 0;
    bool begun = false, awaiting_apply = false, root_pending = false;
    TreeDev t{};
    std::vector<void*> allocs;
    uint64_t total_evals = 0;
    uint32_t epoch = 0, generation = 1;
    bool cache_on = false;

    int allocate(oth_ctx* c, int64_t games, int sims);
    void release();
    // device-pointer, asynchronous building blocks (n = live games, all on ctx->stream)
    int begin(const uint64_t* d_self, const uint64_t* d_opp, const uint8_t* d_active, int64_t n_games);
    int root(bool use_cache);                                  // every live game requests its root evaluation
    int select(bool use_cache);                                // one descent per game -> leaf batch / terminal backup
    int expand(const float* d_policy, const float* d_value, bool policy_is_raw, bool by_slot);   // expand pending leaves + backup
    int evaluate(NetHost* net);                                // leaf batch -> t.eval_policy (priors) / t.eval_value
    int run(NetHost* net, int sims, bool add_noise, uint64_t seed);
    int check_overflow();
    int enable_cache(uint64_t capacity_pow2);                  // allocate the evaluation cache (once)
    void invalidate_cache() { ++generation; }                  // weights changed / new campaign
    int assign();                                              // resolve misses into batch slots / duplicates
    int read_stats(unsigned long long out[4], bool reset);
};

}  // namespace oth

struct oth_search : public oth::SearchHost {};
