// selfplay.cu -- device-resident self-play campaign.
//
// Restates ParallelSelfPlayWorker._execute_batch / execute_episodes
// (src/train/parallel_self_play.py:282-407): `concurrent_games` slots advance in lock-step, one
// BatchMCTS.search_batch-equivalent search per ply (always recorded at temperature 1, :367-372),
// argmax after `temperature_threshold` plies and visit-proportional sampling before (:379-382),
// player sign +1 on even plies (:385), label = get_winner() at the terminal position x player
// (:397-404; the reference takes the winner from whoever is to move at the end -- replicated by
// default, OTH_FLAG_WINNER_BLACK gives the black-relative label).  Unlike the reference a finished
// slot is refilled with the next episode immediately instead of idling until the slowest game of
// its batch ends.  Boards, trees, trajectories and labels never leave the GPU until fetched.
#include "selfplay.cuh"

namespace oth {

__global__ void k_sp_reset(SelfPlayDev d, int64_t num_episodes)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s == 0) {
        d.counters[0] = (unsigned long long)(num_episodes < d.slots ? num_episodes : d.slots);
        d.counters[1] = d.counters[2] = d.counters[3] = d.counters[4] = d.counters[5] = d.counters[6] = 0ULL;
    }
    if (s >= d.slots) return;
    const bool live = s < num_episodes;
    d.self_b[s] = kStartSelf; d.opp_b[s] = kStartOpp;      // board_class(); board.reset()  (:338-341)
    d.move_count[s] = 0;
    d.game_id[s] = live ? (int32_t)s : -1;
    d.active[s] = live ? 1 : 0;
}

// Elect one slot per distinct root position (phase 1) and point everybody else at it (phase 2).
__global__ void __launch_bounds__(256) k_sp_group_elect(SelfPlayDev d, uint32_t ply_epoch, int share)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= d.slots) return;
    d.leader[s] = (int32_t)s;
    d.search_active[s] = d.active[s];
    if (!d.active[s]) return;
    if (!share) { atomicAdd(&d.counters[6], 1ULL); return; }      // every live slot runs its own search
    const uint32_t h = (uint32_t)(mix64(d.self_b[s] ^ mix64(d.opp_b[s] + 0x51ED270B27B4F3CFULL)) & d.r_mask);
    d.root_h[s] = h;
    atomicMin(&d.r_owner[h], ((unsigned long long)(~ply_epoch) << 32) | (unsigned long long)(uint32_t)s);
}

__global__ void __launch_bounds__(256) k_sp_group_follow(SelfPlayDev d)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= d.slots || !d.active[s]) return;
    const uint32_t o = (uint32_t)(d.r_owner[d.root_h[s]] & 0xFFFFFFFFULL);
    if (o != (uint32_t)s && d.self_b[o] == d.self_b[s] && d.opp_b[o] == d.opp_b[s]) {
        d.leader[s] = (int32_t)o;           // same position: the leader's search is this slot's search
        d.search_active[s] = 0;
    } else {
        atomicAdd(&d.counters[6], 1ULL);     // elected, or a different position on the same table entry
    }
}

// One warp per slot: record the sample, choose and play the move, finish / refill the slot.
__global__ void __launch_bounds__(256)
k_sp_move(SelfPlayDev d, TreeDev t, int64_t num_episodes, int threshold, uint64_t seed, uint32_t flags)
{
    const int64_t s = blockIdx.x * 8LL + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= d.slots || !d.active[s]) return;
    uint64_t me = d.self_b[s], you = d.opp_b[s];
    const int ply = d.move_count[s];
    const int game = d.game_id[s];
    const int64_t ld = d.leader[s];                      // whose tree holds this slot's search (itself unless shared)
    const int cnt = t.root_count[ld];
    const Edge* E = t.edges + ld * (int64_t)t.edge_cap;
    oth_sample* smp = d.staging + s * kMaxPlies + (ply < kMaxPlies ? ply : kMaxPlies - 1);
    // ---- record (state, visit distribution, player) : parallel_self_play.py:364,385-388
    for (int j = lane; j < OTH_ACTIONS; j += 32) smp->visits[j] = 0;
    __syncwarp();
    int total = 0, best_k = 0, best_n = -1;
    for (int k = 0; k < cnt; ++k) {                      // <= 33 children: uniform loop, every lane keeps the stats
        const int nv = E[k].n;
        total += nv;
        if (nv > best_n) { best_n = nv; best_k = k; }    // np.argmax: first maximum (:380)
    }
    for (int k = lane; k < cnt; k += 32) smp->visits[E[k].action] = (uint16_t)E[k].n;
    if (lane == 0) {
        smp->self_b = me; smp->opp_b = you; smp->legal = legal_moves(me, you);
        smp->game = game; smp->ply = (int16_t)ply; smp->value = 0; smp->n_children = (uint8_t)cnt;
        smp->pad[0] = smp->pad[1] = smp->pad[2] = 0;
        atomicAdd(&d.counters[4], (unsigned long long)t.n_evals[ld]);   // what the reference would have evaluated for this game
        if (ply >= kMaxPlies) atomicExch(&d.counters[5], 1ULL);
    }
    // ---- choose the move (:379-382)
    int pick = best_k;
    if (ply < threshold && total > 0) {
        const uint64_t r = move_draw(seed, game, ply);
        int target = (int)(((r >> 32) * (uint64_t)total) >> 32);   // uniform in [0,total)
        pick = cnt - 1;
        for (int k = 0; k < cnt; ++k) {
            const int nv = E[k].n;
            if (target < nv) { pick = k; break; }
            target -= nv;
        }
    }
    const int action = E[pick].action;
    apply_known_legal(me, you, action);                   // game.board.make_move(action) (:391)
    const int plies = ply + 1;
    const uint64_t lg = legal_moves(me, you);
    const bool terminal = lg == 0 && legal_moves(you, me) == 0;   // :395
    if (!terminal) {
        if (lane == 0) { d.self_b[s] = me; d.opp_b[s] = you; d.move_count[s] = plies; }
        return;
    }
    // ---- game over: label and flush the trajectory (:397-404)
    int w = winner(me, you);                              // perspective of the side to move at the end
    if ((flags & OTH_FLAG_WINNER_BLACK) && (plies & 1)) w = -w;
    const int n_rec = plies < kMaxPlies ? plies : kMaxPlies;
    oth_sample* rec = d.staging + s * kMaxPlies;
    __syncwarp();
    for (int i = lane; i < n_rec; i += 32) rec[i].value = (int8_t)(w * ((i & 1) ? -1 : 1));
    __syncwarp();
    unsigned long long base = 0;
    int next_game = -1;
    if (lane == 0) {
        base = atomicAdd(&d.counters[2], (unsigned long long)n_rec);
        atomicAdd(&d.counters[3], (unsigned long long)plies);
        const unsigned long long id = atomicAdd(&d.counters[0], 1ULL);
        next_game = id < (unsigned long long)num_episodes ? (int)id : -1;
    }
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    next_game = __shfl_sync(0xFFFFFFFFu, next_game, 0);
    if ((int64_t)(base + n_rec) <= d.out_cap) {
        const uint2* src = reinterpret_cast<const uint2*>(rec);
        uint2* dst = reinterpret_cast<uint2*>(d.out + base);
        const int words = n_rec * (int)(sizeof(oth_sample) / 8);
        for (int i = lane; i < words; i += 32) dst[i] = src[i];
    } else if (lane == 0) {
        atomicExch(&d.counters[5], 1ULL);
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        atomicAdd(&d.counters[1], 1ULL);
        d.self_b[s] = kStartSelf; d.opp_b[s] = kStartOpp; d.move_count[s] = 0;
        d.game_id[s] = next_game;
        d.active[s] = next_game >= 0 ? 1 : 0;
    }
}

int SelfPlayHost::create(oth_ctx* c, const oth_selfplay_config* cf)
{
    ctx = c; cfg = *cf;
    OTH_CHECK_CUDA(cudaSetDevice(c->device));
    int rc = search.allocate(c, cfg.concurrent_games, cfg.num_simulations);
    if (rc) return rc;
    search.c_puct = cfg.c_puct; search.dir_alpha = cfg.dirichlet_alpha; search.dir_eps = cfg.dirichlet_epsilon;
    search.flags = cfg.flags;
    if (cfg.flags & OTH_FLAG_EVAL_CACHE) {
        uint64_t want = (uint64_t)cfg.concurrent_games * 2048ULL, cap = 1 << 16;
        while (cap < want && cap < (1ULL << 27)) cap <<= 1;
        if ((rc = search.enable_cache(cap))) return rc;
    }
    d.slots = cfg.concurrent_games;
    const size_t S = (size_t)d.slots;
    auto grab = [&](void** p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) { set_error("selfplay: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return OTH_ERR_CUDA; }
        allocs.push_back(*p);
        return OTH_OK;
    };
    if ((rc = grab((void**)&d.self_b, S * 8))) return rc;
    if ((rc = grab((void**)&d.opp_b, S * 8))) return rc;
    if ((rc = grab((void**)&d.move_count, S * 4))) return rc;
    if ((rc = grab((void**)&d.game_id, S * 4))) return rc;
    if ((rc = grab((void**)&d.active, S))) return rc;
    if ((rc = grab((void**)&d.staging, S * kMaxPlies * sizeof(oth_sample)))) return rc;
    if ((rc = grab((void**)&d.counters, 8 * sizeof(unsigned long long)))) return rc;
    if ((rc = grab((void**)&d.leader, S * 4))) return rc;
    if ((rc = grab((void**)&d.search_active, S))) return rc;
    if ((rc = grab((void**)&d.root_h, S * 4))) return rc;
    uint64_t rcap = 1024;
    while (rcap < 4 * (uint64_t)S) rcap <<= 1;
    d.r_mask = rcap - 1;
    if ((rc = grab((void**)&d.r_owner, rcap * sizeof(unsigned long long)))) return rc;
    OTH_CHECK_CUDA(cudaMemsetAsync(d.r_owner, 0xFF, rcap * sizeof(unsigned long long), c->stream));
    OTH_CHECK_CUDA(cudaMallocHost((void**)&h_counters, 8 * sizeof(unsigned long long)));
    OTH_CHECK_CUDA(cudaEventCreate(&ev_begin));
    OTH_CHECK_CUDA(cudaEventCreate(&ev_end));
    // trajectories of one full house of games; grown by run() when a campaign plays more episodes than there are slots
    d.out = nullptr; d.out_cap = 0;
    OTH_CHECK_CUDA(cudaMalloc((void**)&d.out, S * kMaxPlies * sizeof(oth_sample)));
    d.out_cap = (int64_t)S * kMaxPlies;
    return OTH_OK;
}

void SelfPlayHost::release()
{
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    search.release();
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    if (d.out) cudaFree(d.out);
    d.out = nullptr;
    if (h_counters) cudaFreeHost(h_counters);
    h_counters = nullptr;
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    ev_begin = ev_end = nullptr;
}

// Which schedule plays the campaign.  Both produce the same records (the search is a deterministic function of the
// root position and the move draw is keyed on (seed, episode, ply)); they differ in how often the network is launched:
// lock-step needs (1 + sims) launches per ply but lets identical roots share one search -- the winner when hundreds of
// thousands of games start together; run-until-miss needs one launch per cache MISS of the slowest slot and wins as
// long as the network launch, not its throughput, is what a step costs.
int SelfPlayHost::pick_schedule() const
{
    const bool noisy = cfg.add_dirichlet_noise && (cfg.flags & OTH_FLAG_ROOT_N_SUM);   // per-game noise enters the search
    if (noisy) return OTH_SCHEDULE_LOCKSTEP;
    if (cfg.schedule == OTH_SCHEDULE_LOCKSTEP || cfg.schedule == OTH_SCHEDULE_ASYNC) return (int)cfg.schedule;
    const bool cached = (cfg.flags & OTH_FLAG_EVAL_CACHE) && !(cfg.flags & OTH_FLAG_EVAL_HASHNET);
    return (cached && d.slots <= kAsyncAutoMaxSlots) ? OTH_SCHEDULE_ASYNC : OTH_SCHEDULE_LOCKSTEP;
}

int SelfPlayHost::run_lockstep(NetHost* net, int64_t num_episodes)
{
    const int grid_t = (int)((d.slots + 255) / 256), grid_w = (int)((d.slots + 7) / 8);
    const int64_t max_moves = (num_episodes + d.slots - 1) / d.slots * kMaxPlies + kMaxPlies;
    for (int64_t mv = 0;; ++mv) {
        OTH_REQUIRE(mv <= max_moves, OTH_ERR_STATE, "oth_selfplay_run: games did not terminate");
        // identical root positions share one search -- unless per-game Dirichlet noise really enters the search
        const bool noisy = cfg.add_dirichlet_noise && (cfg.flags & OTH_FLAG_ROOT_N_SUM);
        const int share = (!noisy && !(cfg.flags & OTH_FLAG_NO_SEARCH_SHARING)) ? 1 : 0;
        ++ply_epoch;
        k_sp_group_elect<<<grid_t, 256, 0, ctx->stream>>>(d, ply_epoch, share);
        ctx->launches++;
        if (share) {
            k_sp_group_follow<<<grid_t, 256, 0, ctx->stream>>>(d);
            ctx->launches++;
        }
        OTH_CHECK_CUDA(cudaGetLastError());
        int rc = search.begin(d.self_b, d.opp_b, d.search_active, d.slots);
        if (rc) return rc;
        const uint64_t step_seed = mix64(run_seed + (uint64_t)mv * 0x9E3779B97F4A7C15ULL);
        if ((rc = search.run(net, cfg.num_simulations, cfg.add_dirichlet_noise != 0, step_seed))) return rc;
        last_ticks += (uint64_t)cfg.num_simulations + 1;
        {
            TimedLaunch timed(ctx, 2);
            k_sp_move<<<grid_w, 256, 0, ctx->stream>>>(d, search.t, num_episodes, cfg.temperature_threshold, run_seed, cfg.flags);
        }
        ctx->launches++;
        OTH_CHECK_CUDA(cudaGetLastError());
        ++moves_played;
        OTH_CHECK_CUDA(cudaMemcpyAsync(h_counters, d.counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        OTH_REQUIRE(h_counters[5] == 0, OTH_ERR_CAPACITY, "oth_selfplay_run: trajectory buffer overflow");
        if ((int64_t)h_counters[1] >= num_episodes) break;
    }
    return OTH_OK;
}

int SelfPlayHost::run(NetHost* net, int64_t num_episodes, int64_t* n_samples, int64_t* n_evals)
{
    OTH_REQUIRE(num_episodes >= 0, OTH_ERR_ARG, "oth_selfplay_run: num_episodes < 0");
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    last_samples = 0;
    if (n_samples) *n_samples = 0;
    if (n_evals) *n_evals = 0;
    if (num_episodes == 0) return OTH_OK;
    const int64_t need = num_episodes * kMaxPlies;
    if (need > d.out_cap) {
        if (d.out) cudaFree(d.out);
        d.out = nullptr; d.out_cap = 0;
        OTH_CHECK_CUDA(cudaMalloc((void**)&d.out, (size_t)need * sizeof(oth_sample)));
        d.out_cap = need;
    }
    // every campaign on this handle draws its moves from its own stream (run 0 uses the configured seed itself)
    run_seed = runs == 0 ? cfg.seed : mix64(cfg.seed ^ (runs * 0xD6E8FEB86659FD93ULL));
    ++runs;
    search.invalidate_cache();          // a campaign starts cold: the weights usually changed since the last one
    {
        unsigned long long dummy[4];
        int rc0 = search.read_stats(dummy, true);
        if (rc0) return rc0;
    }
    const uint64_t launches0 = ctx->launches;
    last_ticks = 0;
    OTH_CHECK_CUDA(cudaEventRecord(ev_begin, ctx->stream));
    k_sp_reset<<<(int)((d.slots + 255) / 256), 256, 0, ctx->stream>>>(d, num_episodes);
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    last_schedule = pick_schedule();
    int rc = last_schedule == OTH_SCHEDULE_ASYNC ? run_async(net, num_episodes) : run_lockstep(net, num_episodes);
    if (rc) return rc;
    OTH_CHECK_CUDA(cudaEventRecord(ev_end, ctx->stream));
    if ((rc = search.check_overflow())) return rc;        // synchronises the stream
    if ((rc = search.read_stats(last_stats, false))) return rc;
    float ms = 0.f;
    OTH_CHECK_CUDA(cudaEventElapsedTime(&ms, ev_begin, ev_end));
    last_run_ms = (double)ms;
    last_launches = ctx->launches - launches0;
    last_searches = (uint64_t)h_counters[6];
    last_samples = (int64_t)h_counters[2];
    if (n_samples) *n_samples = last_samples;
    if (n_evals) *n_evals = (int64_t)h_counters[4];
    return OTH_OK;
}

}  // namespace oth

using namespace oth;

extern "C" {

int oth_selfplay_create(oth_ctx* ctx, const oth_selfplay_config* cfg, oth_selfplay** out)
{
    OTH_REQUIRE(ctx && cfg && out, OTH_ERR_ARG, "oth_selfplay_create: NULL argument");
    OTH_REQUIRE(cfg->concurrent_games > 0 && cfg->concurrent_games <= (1 << 20), OTH_ERR_ARG,
                "oth_selfplay_create: concurrent_games %d out of range", cfg->concurrent_games);
    OTH_REQUIRE(cfg->num_simulations >= 0 && cfg->num_simulations <= 4096, OTH_ERR_ARG,
                "oth_selfplay_create: num_simulations %d out of range", cfg->num_simulations);
    oth_selfplay* sp = new oth_selfplay();
    int rc = sp->create(ctx, cfg);
    if (rc) { sp->release(); delete sp; return rc; }
    *out = sp;
    return OTH_OK;
}

int oth_selfplay_destroy(oth_selfplay* sp)
{
    if (!sp) return OTH_OK;
    sp->release();
    delete sp;
    return OTH_OK;
}

int oth_selfplay_run(oth_selfplay* sp, oth_net* net, int64_t num_episodes, int64_t* n_samples_out, int64_t* n_evals_out)
{
    OTH_REQUIRE(sp, OTH_ERR_ARG, "oth_selfplay_run: NULL handle");
    return sp->run(net, num_episodes, n_samples_out, n_evals_out);
}

int oth_selfplay_stats(oth_selfplay* sp, uint64_t* out5)
{
    OTH_REQUIRE(sp && out5, OTH_ERR_ARG, "oth_selfplay_stats: NULL argument");
    for (int i = 0; i < 4; ++i) out5[i] = sp->last_stats[i];
    out5[4] = sp->last_searches;
    return OTH_OK;
}

int oth_selfplay_timing(oth_selfplay* sp, double* out4)
{
    OTH_REQUIRE(sp && out4, OTH_ERR_ARG, "oth_selfplay_timing: NULL argument");
    out4[0] = sp->last_run_ms;
    out4[1] = (double)sp->last_ticks;
    out4[2] = (double)sp->last_schedule;
    out4[3] = (double)sp->last_launches;
    return OTH_OK;
}

int oth_selfplay_set_seed(oth_selfplay* sp, uint64_t seed)
{
    OTH_REQUIRE(sp, OTH_ERR_ARG, "oth_selfplay_set_seed: NULL handle");
    sp->cfg.seed = seed;
    sp->runs = 0;
    return OTH_OK;
}

int oth_selfplay_fetch(oth_selfplay* sp, oth_sample* out, int64_t capacity, int mem)
{
    OTH_REQUIRE(sp && (out || sp->last_samples == 0), OTH_ERR_ARG, "oth_selfplay_fetch: NULL argument");
    OTH_REQUIRE(capacity >= sp->last_samples, OTH_ERR_CAPACITY, "oth_selfplay_fetch: buffer holds %lld samples, need %lld",
                (long long)capacity, (long long)sp->last_samples);
    if (sp->last_samples == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(sp->ctx->device));
    OTH_CHECK_CUDA(cudaMemcpyAsync(out, sp->d.out, (size_t)sp->last_samples * sizeof(oth_sample),
                                   mem == OTH_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, sp->ctx->stream));
    if (mem == OTH_MEM_HOST) OTH_CHECK_CUDA(cudaStreamSynchronize(sp->ctx->stream));
    return OTH_OK;
}

int oth_selfplay_samples_device(oth_selfplay* sp, const oth_sample** dev_ptr_out, int64_t* count_out)
{
    OTH_REQUIRE(sp && dev_ptr_out && count_out, OTH_ERR_ARG, "oth_selfplay_samples_device: NULL argument");
    *dev_ptr_out = sp->d.out;
    *count_out = sp->last_samples;
    return OTH_OK;
}

}  // extern "C"
