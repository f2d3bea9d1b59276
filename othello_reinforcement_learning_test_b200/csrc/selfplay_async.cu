// selfplay_async.cu -- run-until-miss schedule of the self-play campaign (OTH_SCHEDULE_ASYNC).
//
// The lock-step schedule (selfplay.cu) launches the network once per simulation of a ply: 1 + sims launches per ply,
// whatever the evaluation cache answers.  When few games run (BASELINE config 3: 100 games per iteration) a launch of
// the 21-layer network costs its LATENCY, not its throughput, and most leaves are cache hits (the previous ply's
// subtree is searched again after every move, mcts.py:71: no tree reuse).  Here every slot is its own state machine:
//
//     root request -> simulations -> move (record, sample / arg-max, play, label + flush at game end, refill) -> next root ...
//
// and one launch of k_as_advance drives each slot through as many of those steps as the cache can answer (up to a small
// cap, see async_steps_per_tick); a slot stops when its leaf MISSES.  One tick = advance -> assign (elect one evaluator per
// distinct position) -> network on the compacted misses -> expand (consume + insert).  Network launches per campaign =
// misses of the slowest slot (~0.3-0.5 x its expansions) instead of (1 + sims) x plies: 2,365 instead of 3,332 for 100
// games of 10x128 / 50 simulations.  Every search is still the reference's serial search (one simulation of a game in flight, K = 1),
// so the records equal the lock-step schedule's byte for byte: select/expand/backup restate node.py:62-136 and
// mcts.py:100-172 exactly as search.cu does, the move step restates parallel_self_play.py:354-405 as k_sp_move does.
//
// The table is only READ in k_as_advance and only WRITTEN in k_tree_expand, which are different kernels: no torn entries.
#include <math.h>
#include <stdlib.h>

#include "selfplay.cuh"

namespace oth {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kLanes = 8;                       // lanes per slot (as k_tree_select / k_tree_expand)
constexpr int kAsBlock = 64;                    // 2 warps per block: small campaigns spread over many SMs

// Simulations / moves one slot may complete per launch.  A launch lasts as long as its longest chain of cache hits and
// the network launch behind it waits; measured on B200 (10x128, 50 simulations): 100 slots -- cap 4: 186 games/s, 2: 177,
// 12: 175, unbounded: 154 (lock-step: 178); 4,096 slots -- cap 2: 3,730, 4: 3,530, 12: 3,383 (lock-step: 3,407).
__host__ inline int async_steps_per_tick(int64_t slots) { return slots < 2048 ? 4 : 2; }

constexpr int kEvalPerGame = 0;                 // evaluator indexed by game (hash-net): no batch slot
constexpr int kEvalDirect = 1;                  // cache off: a batch slot straight away
constexpr int kEvalCached = 2;                  // probe the table; misses are resolved by k_tree_assign

struct AsyncParams {
    int64_t num_episodes;
    int sims, threshold, max_steps, eval_mode;
    float c32;
    uint32_t flags, epoch, gen;
    uint64_t seed;
};

// Fresh trees for every slot (mcts.py:71) at the start of a campaign.
__global__ void __launch_bounds__(256) k_as_begin(SelfPlayDev d, TreeDev t)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g == 0) *t.batch_count = 0;
    if (g >= t.games) return;
    const bool live = g < d.slots && d.active[g];
    t.root_self[g] = live ? d.self_b[g] : 0ULL;
    t.root_opp[g] = live ? d.opp_b[g] : 0ULL;
    t.active[g] = live ? 1 : 0;
    t.n_nodes[g] = 0; t.n_edges[g] = 0; t.n_evals[g] = 0; t.sims_done[g] = 0; t.path_len[g] = 0; t.root_count[g] = 0;
    t.pending[g] = 0; t.eval_slot[g] = -1;
}

// One step of the path a descent took: the edge and the statistics it had when it was chosen.  Kept in shared memory so
// that the backup is a set of independent stores (no re-load of the edges, lanes take one level each).
struct PathStep {
    double w;
    int32_t idx, n;
};
constexpr int kPathSmem = 64;                   // levels kept in shared memory per slot (tree depth stays far below: <= 7 measured)

// mcts.py:152-168: value v at the leaf's edge, sign flip per level going up, the root untouched.  Level d is handled by
// lane d (mod 8): E[idx].n = n + 1, E[idx].w = w + (+-v) -- the same two operations `child.update(value)` performs.
__device__ __forceinline__ void backup_from_smem(Edge* E, const PathStep* ps, int depth, double v, int sub)
{
    for (int dd = sub; dd < depth; dd += kLanes) {
        const PathStep st = ps[dd];
        Edge* ed = E + st.idx;
        ed->n = st.n + 1;
        ed->w = st.w + (((depth - 1 - dd) & 1) ? -v : v);
    }
}

// Per-slot state lives in registers for the whole launch (all 8 lanes hold the same values) and is written back once;
// a step then costs the descent's dependent loads (one 24-byte record per level), the table probe and, on a hit, the
// priors row -- not the two dozen dependent global round trips of a load-modify-store per counter and per backup level.
// SPW = slots per warp.  4: every lane works (large campaigns: throughput).  1: one slot per warp, 8 of 32 lanes work -- the
// slots of a warp advance in lock-step through the descent / leaf / move phases, so with four of them every step costs the
// slowest of four and the union of their phases; a small campaign has warps to spare and wants the shortest chain instead.
template <int SPW>
__global__ void __launch_bounds__(kAsBlock) k_as_advance(SelfPlayDev d, TreeDev t, AsyncParams p)
{
    constexpr int kSlotsPerBlock = (kAsBlock / 32) * SPW;
    __shared__ PathStep s_path[kSlotsPerBlock][kPathSmem];
    const int lane = threadIdx.x & 31, sub = lane & (kLanes - 1);
    const int grp_in_block = (threadIdx.x >> 5) * SPW + (lane / kLanes) % SPW;
    const bool lane_works = lane < SPW * kLanes;
    const int64_t g = blockIdx.x * (int64_t)kSlotsPerBlock + grp_in_block;
    const int64_t gs = (lane_works && g < d.slots) ? g : 0;
    bool run = lane_works && g < d.slots && d.active[gs] && !t.pending[gs];
    Edge* E = t.edges + gs * (int64_t)t.edge_cap;
    PathStep* ps = s_path[grp_in_block];
    const int path_cap = t.path_cap < kPathSmem ? t.path_cap : kPathSmem;

    // slot state
    int rc = 0, sd = 0, n_edges = 0, n_nodes = 0, n_evals = 0, ply = 0, game = 0;
    uint64_t root_me = 0ULL, root_you = 0ULL;
    unsigned hits = 0;
    bool dirty = false;
    if (run) {
        rc = t.root_count[gs]; sd = t.sims_done[gs]; n_edges = t.n_edges[gs]; n_nodes = t.n_nodes[gs]; n_evals = t.n_evals[gs];
        root_me = t.root_self[gs]; root_you = t.root_opp[gs];
        ply = d.move_count[gs]; game = d.game_id[gs];
    }

    for (int step = 0; step < p.max_steps; ++step) {
        if (!__any_sync(kFull, run)) break;
        uint64_t me = root_me, you = root_you;
        const bool do_move = run && rc > 0 && sd >= p.sims;          // the search of this ply is complete
        const bool do_sim = run && !do_move;                         // root request (rc == 0) or one simulation

        // ---------------- descent (node.py:91-126, mcts.py:117-123), as k_tree_select ----------------
        int depth = 0;
        {
            int first = 0, cnt = (do_sim && rc > 0) ? rc : 0;
            int parent_n = (p.flags & OTH_FLAG_ROOT_N_SUM) ? sd : 0;   // mcts.py:152-172: the root is never updated
            bool descending = cnt > 0;
            while (__any_sync(kFull, descending)) {
                const double root_of_n = sqrt((double)parent_n);
                double best = -INFINITY, b_w = 0.0;
                int best_e = 0x7FFFFFFF, b_n = 0, b_first = kEdgeLeaf, b_cnt = 0, b_act = 0;
                if (descending) {
                    for (int k = sub; k < cnt; k += kLanes) {
                        const Edge ed = E[first + k];
                        double q = ed.n ? ed.w / (double)ed.n : 0.0;                    // node.py:51-60
                        if (p.flags & OTH_FLAG_Q_CANONICAL) q = -q;
                        const float cp = __fmul_rn(p.c32, ed.p);                        // float32 product (weak Python scalar)
                        const double u = __ddiv_rn(__dmul_rn((double)cp, root_of_n), (double)(1 + ed.n));
                        const double sc = __dadd_rn(q, u);
                        if (sc > best) {                                                // strict >: first maximum wins
                            best = sc; best_e = first + k; b_n = ed.n; b_w = ed.w; b_first = ed.child_first; b_cnt = ed.child_count; b_act = ed.action;
                        }
                    }
                }
                int win_lane = sub;
#pragma unroll
                for (int o = kLanes / 2; o > 0; o >>= 1) {
                    const double os = __shfl_xor_sync(kFull, best, o, kLanes);
                    const int oe = __shfl_xor_sync(kFull, best_e, o, kLanes);
                    const int ol = __shfl_xor_sync(kFull, win_lane, o, kLanes);
                    if (os > best || (os == best && oe < best_e)) { best = os; best_e = oe; win_lane = ol; }
                }
                const int w_n = __shfl_sync(kFull, b_n, win_lane, kLanes);
                const double w_w = __shfl_sync(kFull, b_w, win_lane, kLanes);
                const int w_first = __shfl_sync(kFull, b_first, win_lane, kLanes);
                const int w_cnt = __shfl_sync(kFull, b_cnt, win_lane, kLanes);
                const int w_act = __shfl_sync(kFull, b_act, win_lane, kLanes);
                if (descending) {
                    parent_n = w_n;
                    cnt = w_cnt;
                    if (sub == 0) { PathStep st; st.w = w_w; st.idx = best_e; st.n = w_n; ps[depth] = st; }
                    ++depth;
                    apply_known_legal(me, you, w_act);                                  // mcts.py:122
                    if (cnt == 0 || depth >= path_cap) descending = false;              // child not expanded: this is the leaf
                    else first = w_first;
                }
            }
        }
        __syncwarp();                                                                   // the path is visible to every lane of its slot

        // ---------------- leaf: terminal backup, table hit (expand now), or a request for the network ----------------
        if (__any_sync(kFull, do_sim)) {
            uint64_t lg = 0ULL;
            bool terminal = false;
            if (do_sim) {
                lg = legal_moves(me, you);
                terminal = depth > 0 && lg == 0 && legal_moves(you, me) == 0;           // mcts.py:127 (a root is never terminal here)
            }
            if (do_sim && terminal) {
                backup_from_smem(E, ps, depth, (double)winner(me, you), sub);           // mcts.py:129-130: re-scored, never expanded
                ++sd; dirty = true;
            } else if (do_sim) {
                uint32_t h = 0;
                bool hit = false;
                if (p.eval_mode == kEvalCached) {
                    h = cache_index(t, me, you);
                    const ulonglong2 key = t.c_key[h];
                    const uint32_t gen = t.c_gen[h];
                    hit = key.x == me && key.y == you && gen == p.gen;
                }
                if (hit) {
                    const float* prow = t.c_priors + (size_t)h * 68;                    // masked, renormalised priors as the network wrote them
                    const double value = (double)t.c_value[h];                          // value.item(), mcts.py:144
                    const int cnt = lg ? popc64(lg) : 1;                                // [64] = forced pass (bitboard.pyx:176-178)
                    const int first = n_edges;
                    if (first + cnt > t.edge_cap) {
                        if (sub == 0) atomicExch(t.error_flag, 1);
                        run = false;
                    } else {
                        for (int k = sub; k < cnt; k += kLanes) {                       // node.py:62-89
                            Edge ed;
                            ed.w = 0.0; ed.n = 0; ed.child_first = kEdgeLeaf; ed.child_count = 0; ed.pad = 0;
                            const int action = lg ? nth_set_bit(lg, k) : kPass;
                            ed.p = prow[action]; ed.action = (uint8_t)action;
                            E[first + k] = ed;
                        }
                        ++n_nodes; n_edges = first + cnt; ++n_evals; ++hits; dirty = true;  // n_evals: what the reference would have evaluated
                        if (depth == 0) {
                            rc = cnt;
                        } else {
                            if (sub == 0) {
                                Edge* leaf = E + ps[depth - 1].idx;
                                leaf->child_first = first; leaf->child_count = (uint8_t)cnt;
                            }
                            backup_from_smem(E, ps, depth, value, sub);
                            ++sd;
                        }
                    }
                } else {
                    for (int dd = sub; dd < depth; dd += kLanes) t.path[gs * t.path_cap + dd] = ps[dd].idx;   // k_tree_expand backs up from here
                    if (sub == 0) {
                        t.leaf_self[gs] = me; t.leaf_opp[gs] = you; t.leaf_legal[gs] = lg;
                        t.path_len[gs] = depth;
                        t.pending[gs] = 1;
                        if (p.eval_mode == kEvalCached) {
                            t.leaf_h[gs] = h;
                            t.leaf_src[gs] = kSrcMiss;
                            atomicMin(&t.c_owner[h], ((unsigned long long)(~p.epoch) << 32) | (unsigned long long)(uint32_t)gs);
                        } else if (p.eval_mode == kEvalDirect) {
                            const int slot = atomicAdd(t.batch_count, 1);
                            t.batch_self[slot] = me; t.batch_opp[slot] = you; t.eval_slot[gs] = slot;
                            t.leaf_src[gs] = kSrcSlot;
                            atomicAdd(&t.stats[0], 1ULL);
                        }
                    }
                    run = false;                                                        // until the evaluation arrives
                }
            }
        }

        // ---------------- move (parallel_self_play.py:354-405), as k_sp_move ----------------
        if (__any_sync(kFull, do_move)) {
            const int cnt = do_move ? rc : 0;
            int total = 0, best_n = -1, best_k = 0x7FFFFFFF;
            for (int k = sub; k < cnt; k += kLanes) {
                const int nv = E[k].n;
                total += nv;
                if (nv > best_n) { best_n = nv; best_k = k; }                           // np.argmax: first maximum (:380)
            }
#pragma unroll
            for (int o = kLanes / 2; o > 0; o >>= 1) {
                total += __shfl_xor_sync(kFull, total, o, kLanes);
                const int on = __shfl_xor_sync(kFull, best_n, o, kLanes);
                const int ok = __shfl_xor_sync(kFull, best_k, o, kLanes);
                if (on > best_n || (on == best_n && ok < best_k)) { best_n = on; best_k = ok; }
            }
            oth_sample* smp = d.staging + gs * kMaxPlies + (ply < kMaxPlies ? ply : kMaxPlies - 1);
            // ---- record (state, visit distribution, player): :364,385-388
            if (do_move) for (int j = sub; j < OTH_ACTIONS; j += kLanes) smp->visits[j] = 0;
            __syncwarp();
            if (do_move) for (int k = sub; k < cnt; k += kLanes) smp->visits[E[k].action] = (uint16_t)E[k].n;
            int pick = best_k;
            if (do_move && sub == 0) {
                smp->self_b = me; smp->opp_b = you; smp->legal = legal_moves(me, you);
                smp->game = game; smp->ply = (int16_t)ply; smp->value = 0; smp->n_children = (uint8_t)cnt;
                smp->pad[0] = smp->pad[1] = smp->pad[2] = 0;
                atomicAdd(&d.counters[4], (unsigned long long)n_evals);
                atomicAdd(&d.counters[6], 1ULL);
                if (ply >= kMaxPlies) atomicExch(&d.counters[5], 1ULL);
                // ---- choose the move (:379-382)
                if (ply < p.threshold && total > 0) {
                    const uint64_t r = move_draw(p.seed, game, ply);
                    int target = (int)(((r >> 32) * (uint64_t)total) >> 32);            // uniform in [0,total)
                    pick = cnt - 1;
                    for (int k = 0; k < cnt; ++k) {
                        const int nv = E[k].n;
                        if (target < nv) { pick = k; break; }
                        target -= nv;
                    }
                }
            }
            pick = __shfl_sync(kFull, pick, 0, kLanes);
            bool over = false;
            int plies = 0;
            if (do_move) {
                apply_known_legal(me, you, (int)E[pick].action);                        // game.board.make_move(action) (:391)
                plies = ply + 1;
                over = legal_moves(me, you) == 0 && legal_moves(you, me) == 0;          // :395
                if (!over && sub == 0) { d.self_b[gs] = me; d.opp_b[gs] = you; d.move_count[gs] = plies; }
            }
            // ---- game over: label and flush the trajectory (:397-404)
            int wv = 0, n_rec = 0;
            oth_sample* rec = d.staging + gs * kMaxPlies;
            __syncwarp();
            if (over) {
                wv = winner(me, you);                                                   // perspective of the side to move at the end
                if ((p.flags & OTH_FLAG_WINNER_BLACK) && (plies & 1)) wv = -wv;
                n_rec = plies < kMaxPlies ? plies : kMaxPlies;
                for (int i = sub; i < n_rec; i += kLanes) rec[i].value = (int8_t)(wv * ((i & 1) ? -1 : 1));
            }
            __syncwarp();
            unsigned long long base = 0;
            int next_game = -1;
            if (over && sub == 0) {
                base = atomicAdd(&d.counters[2], (unsigned long long)n_rec);
                atomicAdd(&d.counters[3], (unsigned long long)plies);
                const unsigned long long id = atomicAdd(&d.counters[0], 1ULL);
                next_game = id < (unsigned long long)p.num_episodes ? (int)id : -1;
            }
            base = __shfl_sync(kFull, base, 0, kLanes);
            next_game = __shfl_sync(kFull, next_game, 0, kLanes);
            if (over) {
                if ((int64_t)(base + n_rec) <= d.out_cap) {
                    const uint2* src = reinterpret_cast<const uint2*>(rec);
                    uint2* dst = reinterpret_cast<uint2*>(d.out + base);
                    const int words = n_rec * (int)(sizeof(oth_sample) / 8);
                    for (int i = sub; i < words; i += kLanes) dst[i] = src[i];
                } else if (sub == 0) {
                    atomicExch(&d.counters[5], 1ULL);
                }
            }
            __syncwarp();
            if (do_move) {
                ply = plies;
                if (over) {
                    if (sub == 0) {
                        __threadfence();
                        atomicAdd(&d.counters[1], 1ULL);
                        d.self_b[gs] = kStartSelf; d.opp_b[gs] = kStartOpp; d.move_count[gs] = 0;
                        d.game_id[gs] = next_game;
                        d.active[gs] = next_game >= 0 ? 1 : 0;
                    }
                    me = kStartSelf; you = kStartOpp;                                   // board_class(); board.reset() (:338-341)
                    ply = 0; game = next_game;
                    if (next_game < 0) run = false;                                     // the campaign has no episode left for this slot
                }
                // a fresh tree for the next search (mcts.py:71)
                root_me = me; root_you = you;
                rc = 0; sd = 0; n_edges = 0; n_nodes = 0; n_evals = 0; dirty = true;
                if (sub == 0) { t.root_self[gs] = me; t.root_opp[gs] = you; t.path_len[gs] = 0; t.eval_slot[gs] = -1; }
            }
        }
        __syncwarp();
    }
    // write the slot state back
    if (dirty && sub == 0) {
        t.root_count[gs] = rc; t.sims_done[gs] = sd; t.n_edges[gs] = n_edges; t.n_nodes[gs] = n_nodes; t.n_evals[gs] = n_evals;
        if (hits) atomicAdd(&t.stats[1], (unsigned long long)hits);
    }
}

int SelfPlayHost::run_async(NetHost* net, int64_t num_episodes)
{
    SearchHost& s = search;
    const bool hash = (cfg.flags & OTH_FLAG_EVAL_HASHNET) != 0;
    const bool use_cache = !hash && s.cache_on && (cfg.flags & OTH_FLAG_EVAL_CACHE);
    s.n = d.slots; s.n_act = d.slots; s.t.act_list = nullptr;
    s.begun = true; s.awaiting_apply = false; s.root_pending = false;
    k_as_begin<<<(int)((s.t.games + 255) / 256), 256, 0, ctx->stream>>>(d, s.t);
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());

    AsyncParams p{};
    p.num_episodes = num_episodes;
    p.sims = cfg.num_simulations; p.threshold = cfg.temperature_threshold;
    // A launch lasts as long as its longest chain of cache hits; the network launch that follows waits for it.  Chains are
    // short (about half the leaves miss), so a small cap costs few extra ticks and keeps the tail bounded.
    p.max_steps = async_steps_per_tick(d.slots);
    if (const char* e = getenv("OTH_ASYNC_MAX_STEPS")) { const int v = atoi(e); if (v > 0) p.max_steps = v; }
    p.eval_mode = hash ? kEvalPerGame : (use_cache ? kEvalCached : kEvalDirect);
    p.c32 = (float)cfg.c_puct; p.flags = cfg.flags; p.seed = run_seed;
    TreeDev view = s.t;
    if (!use_cache) view.cache_mask = 0;

    const bool one_per_warp = d.slots <= 512;             // small campaign: one slot per warp (measured: +4.5 % at 100 and 256 slots, -3 % at 1,024)
    const int slots_per_block = (kAsBlock / 32) * (one_per_warp ? 1 : 4);
    const int grid = (int)((d.slots + slots_per_block - 1) / slots_per_block);
    const int check_every = 16;                           // termination is polled, not awaited: the launch queue stays full
                                                          // (a poll drains it once; ticks after the last game ends are empty launches)
    // every tick resolves at least one simulation of every unfinished slot
    const int64_t rounds = (num_episodes + d.slots - 1) / d.slots + 1;
    const int64_t max_ticks = rounds * kMaxPlies * (int64_t)(cfg.num_simulations + 3) + 64;
    for (int64_t tick = 0;; ++tick) {
        OTH_REQUIRE(tick <= max_ticks, OTH_ERR_STATE, "oth_selfplay_run: games did not terminate");
        ++s.epoch;
        p.epoch = s.epoch; p.gen = s.generation;
        {
            TimedLaunch timed(ctx, 1);
            if (one_per_warp) k_as_advance<1><<<grid, kAsBlock, 0, ctx->stream>>>(d, view, p);
            else k_as_advance<4><<<grid, kAsBlock, 0, ctx->stream>>>(d, view, p);
        }
        ctx->launches++;
        OTH_CHECK_CUDA(cudaGetLastError());
        int rc;
        if (use_cache && (rc = s.assign())) return rc;
        if ((rc = s.evaluate(net))) return rc;
        if ((rc = s.expand(s.t.eval_policy, s.t.eval_value, hash, !hash))) return rc;
        ++last_ticks;
        if ((tick + 1) % check_every == 0) {
            OTH_CHECK_CUDA(cudaMemcpyAsync(h_counters, d.counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
            OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
            OTH_REQUIRE(h_counters[5] == 0, OTH_ERR_CAPACITY, "oth_selfplay_run: trajectory buffer overflow");
            if ((int64_t)h_counters[1] >= num_episodes) break;
            if ((tick + 1) % (8 * check_every) == 0 && (rc = s.check_overflow())) return rc;   // a full edge pool would stall its slot forever
        }
    }
    moves_played += (uint64_t)h_counters[6];
    return OTH_OK;
}

}  // namespace oth
