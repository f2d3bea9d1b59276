// search_wave.cu -- opt-in throughput mode of the tree search: K simulations per game in flight per step,
// kept apart by virtual loss (the "INFLIGHT_K" switch of SURVEY.md 8(b)).
//
// The reference runs one simulation at a time (src/mcts/mcts.py:89-92); that is the default everywhere in
// this library and the mode every parity claim is made in.  With few games (an arena match, the GUI) one leaf
// per game per network launch wastes the GPU, so this file adds waves: the K descents of a game are made one
// after the other by the game's warp, each leaving a virtual visit (a per-edge counter scored as a lost result) on its path so that the
// next descent is pushed elsewhere; the K leaves are evaluated in ONE network launch; expansion and backup then
// replace the virtual loss by the real value, again in descent order.  Everything a game does is sequential
// inside its warp, so results are deterministic and the tree needs no atomics; K = 1 reproduces the reference
// search exactly (tested).  For K > 1 visit counts deviate from the reference by construction -- the tests
// report the total-variation distance to the K = 1 result instead of asserting equality.
#include <math.h>

#include "bitboard.cuh"
#include "search.cuh"

namespace oth {

constexpr int kWaveWarps = 8;
constexpr int kWaveBlock = kWaveWarps * 32;

struct WaveDev {
    int K;
    int32_t* path;        // [G][K][path_cap]
    int32_t* path_len;    // [G][K]
    uint8_t* state;       // [G][K] 0 = unused, 1 = waiting for the evaluator
    uint64_t *leaf_self, *leaf_opp, *leaf_legal;   // [G*K]
    int32_t* slot;        // [G*K] index into the compacted batch
    uint64_t *batch_self, *batch_opp;              // [G*K]
    float* policy;        // [G*K][65]
    float* value;         // [G*K]
    int32_t* count;       // compacted batch size (device)
    int32_t* edge_vn;     // [G][edge_cap] virtual visits currently on each edge (kept apart from N/W so that K = 1 is exact)
};

// One level of the descent: returns the chosen edge index and hands back what the next level needs.
__device__ __forceinline__ int wave_pick_child(const TreeDev& t, const WaveDev& w, int64_t g, int first, int cnt, int parent_n,
                                                float c32, uint32_t flags, int lane, int& n_eff, int& child_first, int& child_cnt,
                                                int& action)
{
    const int64_t eb = g * (int64_t)t.edge_cap;
    const double root_of_n = sqrt((double)parent_n);
    double best = -INFINITY;
    int best_e = 0x7FFFFFFF, b_n = 0, b_first = kEdgeLeaf, b_cnt = 0, b_act = 0;
    for (int k = lane; k < cnt; k += 32) {
        const Edge ed = t.edges[eb + first + k];
        const int vn = w.edge_vn[eb + first + k];
        const int nv = ed.n + vn;                                               // real + virtual visits
        // a virtual visit counts as a result that makes the edge look worse to its parent
        const double wsum = vn ? ed.w + ((flags & OTH_FLAG_Q_CANONICAL) ? 1.0 : -1.0) * (double)vn : ed.w;
        double q = nv ? wsum / (double)nv : 0.0;
        if (flags & OTH_FLAG_Q_CANONICAL) q = -q;
        const float cp = __fmul_rn(c32, ed.p);
        const double s = __dadd_rn(q, __ddiv_rn(__dmul_rn((double)cp, root_of_n), (double)(1 + nv)));
        if (s > best) { best = s; best_e = first + k; b_n = nv; b_first = ed.child_first; b_cnt = ed.child_count; b_act = ed.action; }
    }
    int win_lane = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const int oe = __shfl_xor_sync(0xFFFFFFFFu, best_e, o);
        const int ol = __shfl_xor_sync(0xFFFFFFFFu, win_lane, o);
        if (os > best || (os == best && oe < best_e)) { best = os; best_e = oe; win_lane = ol; }
    }
    n_eff = __shfl_sync(0xFFFFFFFFu, b_n, win_lane);
    child_first = __shfl_sync(0xFFFFFFFFu, b_first, win_lane);
    child_cnt = __shfl_sync(0xFFFFFFFFu, b_cnt, win_lane);
    action = __shfl_sync(0xFFFFFFFFu, b_act, win_lane);
    return best_e;
}

// K descents per game, one after the other, each leaving virtual loss on its path.
__global__ void __launch_bounds__(kWaveBlock)
k_wave_select(TreeDev t, WaveDev w, int64_t n, int sims_target, float c32, uint32_t flags)
{
    const int64_t g = blockIdx.x * (int64_t)kWaveWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= n || !t.active[g]) return;
    const int64_t eb = g * (int64_t)t.edge_cap;
    Edge* E = t.edges + eb;
    for (int k = 0; k < w.K; ++k) {
        if (lane == 0) w.state[g * w.K + k] = 0;
    }
    __syncwarp();
    int in_flight = 0;
    for (int k = 0; k < w.K; ++k) {
        if (t.sims_done[g] + in_flight >= sims_target) break;
        int32_t* path = w.path + (g * w.K + k) * (int64_t)t.path_cap;
        uint64_t me = t.root_self[g], you = t.root_opp[g];
        int first = 0, cnt = t.root_count[g], depth = 0;
        int parent_n = (flags & OTH_FLAG_ROOT_N_SUM) ? t.sims_done[g] + in_flight : 0;
        int child_first = kEdgeLeaf;
        for (;;) {
            int n_eff, c_cnt, action;
            const int e = wave_pick_child(t, w, g, first, cnt, parent_n, c32, flags, lane, n_eff, child_first, c_cnt, action);
            parent_n = n_eff;
            if (lane == 0) path[depth] = e;
            ++depth;
            apply_known_legal(me, you, action);
            if (c_cnt == 0 || depth >= t.path_cap) break;
            first = child_first; cnt = c_cnt;
        }
        __syncwarp();
        if (child_first == kEdgePending) break;    // an earlier descent of this wave already owns this leaf: end the wave here
        const uint64_t lg = legal_moves(me, you);
        const bool terminal = lg == 0 && legal_moves(you, me) == 0;
        if (lane == 0) {
            if (terminal) {                          // scored at once, like mcts.py:127-130,152-168
                double v = (double)winner(me, you);
                for (int i = depth - 1; i >= 0; --i) { Edge* ed = E + path[i]; ed->n += 1; ed->w += v; v = -v; }
                t.sims_done[g] += 1;
            } else {
                for (int i = 0; i < depth; ++i) w.edge_vn[eb + path[i]] += 1;                       // virtual loss
                E[path[depth - 1]].child_first = kEdgePending;
                const int64_t j = g * w.K + k;
                w.leaf_self[j] = me; w.leaf_opp[j] = you; w.leaf_legal[j] = lg;
                w.path_len[j] = depth;
                w.state[j] = 1;
                const int s = atomicAdd(w.count, 1);
                w.batch_self[s] = me; w.batch_opp[s] = you; w.slot[j] = s;
            }
        }
        __syncwarp();
        if (!terminal) ++in_flight;
    }
}

__global__ void __launch_bounds__(kWaveBlock) k_wave_hashnet(WaveDev w, int64_t total)
{
    const int64_t s = blockIdx.x * (int64_t)kWaveWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= total || s >= *w.count) return;
    const uint64_t a = w.batch_self[s], b = w.batch_opp[s];
    const uint64_t h = mix64(a ^ mix64(b + 0x632BE59BD9B4E019ULL));
    for (int i = lane; i < 65; i += 32) {
        const uint64_t x = mix64(h + (uint64_t)(i + 1) * 0xD1342543DE82EF95ULL);
        w.policy[s * 65 + i] = (float)((uint32_t)((x >> 24) & 0xFFFFu) + 1u) * (1.0f / 4194304.0f);
    }
    if (lane == 0) {
        const uint64_t xv = mix64(h ^ 0xA5A5A5A5A5A5A5A5ULL);
        w.value[s] = (float)((int32_t)((xv >> 16) & 0xFFFFFu) - (1 << 19)) * (1.0f / 524288.0f);
    }
}

// Expand the wave's leaves in descent order and swap the virtual loss for the real value.
__global__ void __launch_bounds__(kWaveBlock) k_wave_expand(TreeDev t, WaveDev w, int64_t n, int policy_is_raw)
{
    __shared__ float s_pri[kWaveWarps][68];
    const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = blockIdx.x * (int64_t)kWaveWarps + wp;
    if (blockIdx.x == 0 && threadIdx.x == 0) *w.count = 0;
    if (g >= n || !t.active[g]) return;
    const int64_t eb = g * (int64_t)t.edge_cap;
    float* pri = s_pri[wp];
    for (int k = 0; k < w.K; ++k) {
        const int64_t j = g * w.K + k;
        if (!w.state[j]) continue;
        const int64_t src = w.slot[j];
        const uint64_t lg = w.leaf_legal[j];
        for (int a = lane; a < 65; a += 32) pri[a] = w.policy[src * 65 + a];
        __syncwarp();
        if (policy_is_raw && lane == 0) mask_and_renormalise(pri, lg);
        __syncwarp();
        const int depth = w.path_len[j];
        const int32_t* path = w.path + j * (int64_t)t.path_cap;
        const int cnt = lg ? popc64(lg) : 1;
        const int first = t.n_edges[g];
        if (first + cnt > t.edge_cap) {
            if (lane == 0) atomicExch(t.error_flag, 1);
            return;
        }
        Edge* E = t.edges + eb;
        for (int c = lane; c < cnt; c += 32) {
            Edge ed;
            ed.w = 0.0; ed.n = 0; ed.child_first = kEdgeLeaf; ed.child_count = 0; ed.pad = 0;
            const int action = lg ? nth_set_bit(lg, c) : kPass;
            ed.p = pri[action]; ed.action = (uint8_t)action;
            E[first + c] = ed;
        }
        __syncwarp();
        if (lane == 0) {
            t.n_nodes[g] += 1;
            t.n_edges[g] = first + cnt;
            t.n_evals[g] += 1;
            Edge* leaf = E + path[depth - 1];
            leaf->child_first = first; leaf->child_count = (uint8_t)cnt;
            double v = (double)w.value[src];
            for (int i = depth - 1; i >= 0; --i) {                  // the virtual visit becomes a real one (mcts.py:152-168)
                Edge* ed = E + path[i];
                w.edge_vn[eb + path[i]] -= 1; ed->n += 1; ed->w += v; v = -v;
            }
            t.sims_done[g] += 1;
            w.state[j] = 0;
        }
        __syncwarp();
    }
}

struct WaveHost {
    WaveDev w{};
    std::vector<void*> allocs;
    int64_t games = 0;
    int K = 0;
};

}  // namespace oth

using namespace oth;

extern "C" int oth_search_run_waves(oth_search* s, oth_net* net, int num_simulations, int inflight_k, int add_dirichlet_noise,
                                    uint64_t seed)
{
    OTH_REQUIRE(s, OTH_ERR_ARG, "oth_search_run_waves: NULL handle");
    OTH_REQUIRE(inflight_k >= 1 && inflight_k <= 64, OTH_ERR_ARG, "oth_search_run_waves: inflight_k %d out of range 1..64", inflight_k);
    OTH_REQUIRE(s->begun, OTH_ERR_STATE, "oth_search_run_waves: call oth_search_begin first");
    OTH_REQUIRE(num_simulations >= 0 && num_simulations <= s->max_sims, OTH_ERR_ARG, "oth_search_run_waves: too many simulations");
    oth_ctx* ctx = s->ctx;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = s->n;
    if (n == 0) return OTH_OK;
    const bool raw = (s->flags & OTH_FLAG_EVAL_HASHNET) != 0;
    OTH_REQUIRE(raw || (net && net->loaded && net->ctx == ctx), OTH_ERR_STATE, "oth_search_run_waves: no usable network");
    // root evaluation + expansion through the ordinary path (mcts.py:74-82)
    s->root_pending = false;
    int rc = s->run(net, 0, add_dirichlet_noise != 0, seed);
    if (rc) return rc;
    // wave scratch (allocated per call: this is the latency mode, a handful of games)
    WaveDev w{};
    w.K = inflight_k;
    const size_t GK = (size_t)n * inflight_k;
    std::vector<void*> owned;
    auto grab = [&](void** p, size_t bytes) {
        cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
        if (e != cudaSuccess) { set_error("oth_search_run_waves: alloc %zu failed: %s", bytes, cudaGetErrorString(e)); return OTH_ERR_CUDA; }
        owned.push_back(*p);
        return OTH_OK;
    };
    auto release = [&] { for (void* p : owned) cudaFreeAsync(p, ctx->stream); };
#define G_(ptr, bytes) if ((rc = grab((void**)&ptr, (bytes)))) { release(); return rc; }
    G_(w.path, GK * s->t.path_cap * 4) G_(w.path_len, GK * 4) G_(w.state, GK) G_(w.leaf_self, GK * 8) G_(w.leaf_opp, GK * 8)
    G_(w.leaf_legal, GK * 8) G_(w.slot, GK * 4) G_(w.batch_self, GK * 8) G_(w.batch_opp, GK * 8) G_(w.policy, GK * 65 * 4)
    G_(w.value, GK * 4) G_(w.count, 4) G_(w.edge_vn, (size_t)n * s->t.edge_cap * 4)
#undef G_
    cudaMemsetAsync(w.count, 0, 4, ctx->stream);
    cudaMemsetAsync(w.edge_vn, 0, (size_t)n * s->t.edge_cap * 4, ctx->stream);
    cudaMemsetAsync(w.state, 0, GK, ctx->stream);
    const int grid = (int)((n + kWaveWarps - 1) / kWaveWarps);
    const int max_waves = num_simulations + 2;                    // every wave completes at least one simulation per live game
    int32_t* h_done = nullptr;
    cudaMallocHost((void**)&h_done, sizeof(int32_t) * (size_t)n);
    for (int wave = 0; wave < max_waves; ++wave) {
        {
            TimedLaunch timed(ctx, 1);
            k_wave_select<<<grid, kWaveBlock, 0, ctx->stream>>>(s->t, w, n, num_simulations, (float)s->c_puct, s->flags);
        }
        ctx->launches++;
        if (raw) {
            TimedLaunch timed(ctx, 1);
            k_wave_hashnet<<<(unsigned)((GK + kWaveWarps - 1) / kWaveWarps), kWaveBlock, 0, ctx->stream>>>(w, (int64_t)GK);
            ctx->launches++;
        } else {
            rc = net_forward_device(net, w.batch_self, w.batch_opp, (int64_t)GK, w.policy, w.value, kOutPriors, w.count);
            if (rc) break;
        }
        {
            TimedLaunch timed(ctx, 1);
            k_wave_expand<<<grid, kWaveBlock, 0, ctx->stream>>>(s->t, w, n, raw ? 1 : 0);
        }
        ctx->launches++;
        // finished when every live game has all its simulations (one small read-back per wave)
        cudaMemcpyAsync(h_done, s->t.sims_done, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("oth_search_run_waves: %s", cudaGetErrorString(e)); rc = OTH_ERR_CUDA; break; }
        bool all = true;
        for (int64_t i = 0; i < n; ++i) if (h_done[i] < num_simulations) { all = false; break; }
        if (all) break;
    }
    cudaFreeHost(h_done);
    release();
    if (rc) return rc;
    OTH_CHECK_CUDA(cudaGetLastError());
    return s->check_overflow();
}
