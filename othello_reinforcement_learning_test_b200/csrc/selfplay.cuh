// selfplay.cuh -- device/host state of a self-play campaign, shared by the two schedules:
//   selfplay.cu        lock-step: one search per ply for every slot (BatchMCTS.search_batch shape), identical roots share a search;
//   selfplay_async.cu  run-until-miss: every slot is its own state machine that keeps simulating / moving / starting the
//                      next search while its leaves hit the evaluation cache and only stops when it needs the network.
#pragma once
#include "bitboard.cuh"
#include "search.cuh"

namespace oth {

constexpr int64_t kAsyncAutoMaxSlots = 8192;    // OTH_SCHEDULE_AUTO: run-until-miss up to this many slots (measured: +5..14 % up to 4,096 slots, even at 18,944)
constexpr int kMaxPlies = 128;   // <= 60 placements + at most one pass between/around them
static_assert(sizeof(oth_sample) == 168, "oth_sample layout is part of the C ABI");

struct SelfPlayDev {
    int64_t slots;
    uint64_t *self_b, *opp_b;
    int32_t *move_count, *game_id;
    uint8_t* active;
    oth_sample* staging;            // [slots][kMaxPlies]
    oth_sample* out;
    int64_t out_cap;
    unsigned long long* counters;   // 0 started, 1 finished, 2 samples, 3 plies, 4 evals, 5 overflow, 6 searches run
    // search-level sharing: slots whose root position is identical run ONE search (the search is a deterministic
    // function of the root position), the others read the leader's root statistics
    int32_t* leader;                // [slots] slot whose tree holds this slot's search
    uint8_t* search_active;         // [slots] 1 = this slot runs a search this ply
    uint32_t* root_h;               // [slots] index into the election table
    unsigned long long* r_owner;    // [r_mask+1] (~ply << 32 | slot), atomicMin elects the leader
    uint64_t r_mask;
};

// Move sampling: one 64-bit draw per (run seed, episode, ply) -- parallel_self_play.py:379-382 draws from numpy's
// global stream instead; the distribution (visit-proportional) is what is reproduced, not the stream.
__device__ __forceinline__ uint64_t move_draw(uint64_t seed, int game, int ply)
{
    return mix64(seed ^ mix64(((uint64_t)(uint32_t)game << 16) ^ (uint64_t)ply));
}

struct SelfPlayHost {
    oth_ctx* ctx = nullptr;
    oth_selfplay_config cfg{};
    SearchHost search;
    SelfPlayDev d{};
    std::vector<void*> allocs;
    unsigned long long* h_counters = nullptr;   // pinned
    int64_t last_samples = 0;
    uint64_t moves_played = 0;
    uint32_t ply_epoch = 0;
    uint64_t last_searches = 0;
    unsigned long long last_stats[4] = {0, 0, 0, 0};
    // per-run bookkeeping
    uint64_t runs = 0;              // campaigns played on this handle: mixed into the sampling seed
    uint64_t run_seed = 0;          // seed of the campaign in flight
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    double last_run_ms = 0.0;       // device time of the last campaign (events on the context stream)
    uint64_t last_ticks = 0;        // network launches (lock-steps / ticks) of the last campaign
    int last_schedule = 0;          // OTH_SCHEDULE_* the last campaign really used
    uint64_t last_launches = 0;     // kernels launched by the last campaign

    int create(oth_ctx* c, const oth_selfplay_config* cf);
    void release();
    int run(NetHost* net, int64_t num_episodes, int64_t* n_samples, int64_t* n_evals);
    int run_lockstep(NetHost* net, int64_t num_episodes);
    int run_async(NetHost* net, int64_t num_episodes);
    int pick_schedule() const;
};

}  // namespace oth

struct oth_selfplay : public oth::SelfPlayHost {};
