// net_tc2.cu -- OthelloResNet trunk on CTA PAIRS: tcgen05 cta_group::2 implicit-GEMM convolutions (sm_100a).
//
// Same network, same shared-memory activation layout, same epilogue and head warps as net_tc.cu; what changes is
// who feeds the tensor cores.  Two CTAs of a cluster (the two SMs of a TPC) hold 4 boards each and execute every
// MMA together: M = 256 (CTA 0's tile rows 0..127, CTA 1's rows 128..255), N = F, K = 16.  Each CTA supplies its
// own activation rows (A) and only HALF of the weight rows (B: cout [64 r, 64 r + 64) from CTA r), so per SM
//   - the B operand read from shared memory per MMA halves (A 4 KB + B 2 KB instead of 4 + 4),
//   - the weight bytes streamed from L2 into shared memory halve (each CTA loads its half of every stage),
// which is what a power-capped B200 needs: the arithmetic is unchanged, the data movement per flop is not.
//
// Roles per CTA (384 threads, as in net_tc.cu): warps 0-7 epilogue, warp 8 weight producer, warps 10-11 heads;
// warp 9 is the MMA issuer in CTA 0 (the leader issues for the pair) and a relay in CTA 1 (forwards "my half of
// this weight stage has landed" to the leader's barrier -- plain bulk copies cannot signal a peer's mbarrier).
//   bar_full[slot]  leader: own expect_tx arrive + own bytes + the relay's remote arrive; peer: local only
//   bar_empty[slot] tcgen05.commit multicast: one arrival in each CTA when the pair's MMAs have read the slot
//   bar_acc[tile]   tcgen05.commit multicast: accumulators of the tile complete (each CTA reads its own TMEM)
//   bar_act[..]     leader only: one arrival per epilogue warp of the tile, 4 local + 4 remote, per 32-channel split
// Both CTAs of a pair walk the same number of items; an index past the end is a dummy item (zero boards, no stores).
//
// Round 2 looked at this engine again as a LATENCY shape (one tile per CTA: half the weight bytes per SM for a small batch):
// a layer takes 12.5-13.4 k cycles with one tile or two (the single-CTA kernel: 10.2 k), i.e. ~350 cycles per ring stage
// whatever the MMA count, also with one relay lane per ring slot -- the pair's per-stage hand-shake, not bytes or MMAs, sets
// its pace.  Not used for small batches either.
// STATUS (measured on B200, 10x128, 18,944 positions/launch): bit-identical to net_tc.cu on every batch size tried, but
// ~10 % slower kernel-only (1.46 vs 1.60-1.65 PFLOP/s): the pair's M=256 x N=128 x K=16 MMAs retire every ~78 cycles
// instead of 64 (independent of ring depth 6/12 and of how the peer's "weights landed" reaches the leader), and the
// network's N = F = 128 cannot be widened to hide it.  Kept as an opt-in engine (OTH_NET_ENGINE_TCGEN05_PAIR); the
// single-CTA kernel stays the product path.
//
// Restates src/model/net.py:15-61,139-205 (eval mode, BN folded) -- numerics identical to net_tc.cu (bf16
// operands, fp32 accumulation in TMEM, bf16 activations between layers, fp32 heads).
#include "common.cuh"
#include "net_common.cuh"
#include "net_host.cuh"
#include "net_tc_common.cuh"
#include "tc_ptx.cuh"

namespace oth {
namespace tc {

template <int F>
struct Cfg2 {
    static constexpr int KC = F / 8;
    static constexpr int kSplits = F / kSplitChannels;
    static constexpr int kPlanesPerSplit = kSplitChannels / 8;
    static constexpr int kStagesPerConv = 9 * kSplits;
    static constexpr int kMmasPerStage = kSplitChannels / 16;
    static constexpr int kNHalf = F / 2;                            // weight rows (cout) held by one CTA
    static constexpr int kStages = F == 128 ? 12 : 6;               // ring slots
    static constexpr int kTripStages = 3 * kStages;                 // issue loop: one trip = three ring rounds
    static constexpr int kTileBytes = tile_buffer_bytes(KC);
    static constexpr int kStageBytes = kPlanesPerSplit * kNHalf * 16;   // per CTA: [planes][F/2 rows][8 ch]
    static constexpr int kStemTapBytes = 2 * kNHalf * 16;
    static constexpr int kTmemCols = 4 * F;
    static constexpr int offA = 0;
    static constexpr int offB = 2 * kTileBytes;
    static constexpr int offRing = 4 * kTileBytes;
    static constexpr int kRingSlotBytes = kStageBytes;
    static constexpr int offHeads = offRing + kStages * kRingSlotBytes;
    static constexpr int offBars = offHeads + 2 * (int)sizeof(HeadScratch);
    static constexpr int kNumBars = 2 * kStages + 2 + 2 * kSplits;
    static constexpr int offMisc = offBars + kNumBars * 8;
    static constexpr int offBias = offMisc + 144;
    static constexpr int offHeadW = offBias + 4 * F * 4;
    static constexpr int kSmemBytes = offHeadW + 3 * F * 4;
    static_assert(2 * kStemTapBytes == kStageBytes, "stem stages reuse the trunk's ring slots");
    static_assert(kStagesPerConv % kTripStages == 0, "a conv is a whole number of issue trips");
    static_assert(kStages >= 6, "the stem needs six real stages in its ring round");
    static_assert(offBias % 16 == 0 && offHeadW % 16 == 0, "bias / head-weight staging must be 16-byte aligned");
    static_assert(kRingSlotBytes % 128 == 0, "ring slots must stay 128-byte aligned");
    static_assert(kSmemBytes <= 232448, "shared-memory budget (227 KB) exceeded");
};

// Leader CTA, MMA warp: issue one convolution for both tiles of both CTAs.  Shape as in net_tc.cu (issue_layer):
// warp-uniform, elected lane issues, descriptors are "base + immediate", one trip = three ring rounds unrolled.
// The stem occupies one whole ring round: six real stages (tap row x {taps dx -1,0 | tap dx +1}) and
// kStages - 6 empty ones, so that every layer starts at ring slot 0.
template <int F, bool STEM>
__device__ __forceinline__ void issue_layer2(uint32_t smem_base, uint32_t in_off, uint32_t d_col, uint64_t* bar_full,
                                             uint64_t* bar_empty, uint64_t* bar_acc, uint64_t* bar_act, uint32_t act_phase,
                                             uint32_t& round)
{
    using C = Cfg2<F>;

    constexpr uint32_t idesc = umma_idesc_m256(F);
    constexpr uint32_t kAHi = (uint32_t)(kGroupUnits) | (1u << 14);                 // SBO = 144 B, version 1
    constexpr uint32_t kBHi = (uint32_t)(128 >> 4) | (1u << 14);                    // SBO = 128 B
    constexpr uint32_t kALboField = (uint32_t)kPlaneUnits << 16;                    // LBO = plane stride
    constexpr uint32_t kBLboField = (uint32_t)C::kNHalf << 16;                      // LBO = F/2 rows x 16 B
    constexpr uint32_t kSlotUnits = (uint32_t)(C::kRingSlotBytes >> 4);
    constexpr uint32_t kTileUnits = (uint32_t)(C::kTileBytes >> 4);
    const uint32_t a_row0 = (((smem_base + in_off) >> 4) + kGuardUnits + kHaloUnits) | kALboField;
    const uint32_t ring0 = ((smem_base + (uint32_t)C::offRing) >> 4) | kBLboField;
    if (STEM) {
#pragma unroll
        for (int i = 0; i < 2 * C::kSplits; ++i) mbar_wait_cluster(&bar_act[i], act_phase);
#pragma unroll
        for (int s = 0; s < C::kStages; ++s) {
            constexpr int kTapUnits = C::kStemTapBytes >> 4;
            const int dy = s / 2, part = s % 2;
            mbar_wait_cluster(&bar_full[s], round & 1);
            tc_fence_after();
            if (elect_one()) {
                if (s < 6) {
#pragma unroll
                    for (int tile = 0; tile < 2; ++tile) {
#pragma unroll
                        for (int k = 0; k < (part == 0 ? 2 : 1); ++k) {
                            const int tx = part == 0 ? k : 2;
                            const uint32_t a_u = a_row0 + (uint32_t)(tile * kTileUnits + (dy - 1) * 2 * kGroupUnits + (tx - 1));
                            const uint32_t b_u = ring0 + (uint32_t)(s * kSlotUnits + k * kTapUnits);
                            umma_bf16_2cta(d_col + (uint32_t)(tile * F), ((uint64_t)kAHi << 32) | a_u, ((uint64_t)kBHi << 32) | b_u, idesc,
                                           (s > 0 || k > 0) ? 1u : 0u);
                        }
                        if (s == 5) umma_commit_2cta(&bar_acc[tile]);
                    }
                }
                umma_commit_2cta(&bar_empty[s]);
            }
            __syncwarp();
        }
        ++round;
    } else {
#pragma unroll 1
        for (int trip = 0; trip < C::kStagesPerConv / C::kTripStages; ++trip) {
            constexpr int kSplitsPerTrip = C::kTripStages / 9;
            const uint32_t a_trip = a_row0 + (uint32_t)(trip * kSplitsPerTrip * C::kPlanesPerSplit * kPlaneUnits);
#pragma unroll
            for (int t = 0; t < C::kTripStages; ++t) {
                const int ql = t / 9, tap = t % 9, sl = t % C::kStages;
                if (tap == 0) {
                    mbar_wait_cluster(&bar_act[0 * C::kSplits + trip * kSplitsPerTrip + ql], act_phase);
                    mbar_wait_cluster(&bar_act[1 * C::kSplits + trip * kSplitsPerTrip + ql], act_phase);
                }
                mbar_wait_cluster(&bar_full[sl], (round + t / C::kStages) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const int shift = (tap / 3 - 1) * 2 * kGroupUnits + (tap % 3 - 1);
#pragma unroll
                    for (int tile = 0; tile < 2; ++tile) {
#pragma unroll
                        for (int j = 0; j < C::kMmasPerStage; ++j) {
                            const uint32_t a_u = a_trip + (uint32_t)(tile * kTileUnits + (ql * C::kPlanesPerSplit + 2 * j) * kPlaneUnits + shift);
                            const uint32_t b_u = ring0 + (uint32_t)(sl * kSlotUnits + 2 * j * C::kNHalf);
                            umma_bf16_2cta(d_col + (uint32_t)(tile * F), ((uint64_t)kAHi << 32) | a_u, ((uint64_t)kBHi << 32) | b_u, idesc,
                                           (t > 0 || j > 0) ? 1u : (trip > 0 ? 1u : 0u));
                        }
                        if (t == C::kTripStages - 1 && trip == C::kStagesPerConv / C::kTripStages - 1) umma_commit_2cta(&bar_acc[tile]);
                    }
                    umma_commit_2cta(&bar_empty[sl]);
                }
                __syncwarp();
            }
            round += 3;
        }
    }
}

template <int F>
__global__ void __launch_bounds__(kThreads, 1)
k_net_tc2(const NetDev net, const uint64_t* __restrict__ self_b, const uint64_t* __restrict__ opp_b, int64_t n,
          float* __restrict__ policy_out, float* __restrict__ value_out, int out_kind, const int32_t* __restrict__ n_dev)
{
    if (n_dev) { const int64_t nd = *n_dev; n = nd < n ? nd : n; }
    using C = Cfg2<F>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::offBars);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + C::kStages;
    uint64_t* bar_acc = bars + 2 * C::kStages;
    uint64_t* bar_act = bars + 2 * C::kStages + 2;
    float* head_w = reinterpret_cast<float*>(smem + C::offHeadW);
    Misc* misc = reinterpret_cast<Misc*>(smem + C::offMisc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_layers = 1 + 2 * net.blocks;
    const int64_t n_items = (n + 3) / 4;
    const int64_t n_rounds = (n_items + gridDim.x - 1) / gridDim.x;     // both CTAs of a pair run all of them

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::kStages; ++i) { mbar_init(&bar_full[i], rank == 0 ? 2 : 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; ++i) mbar_init(&bar_acc[i], 1);
        for (int i = 0; i < 2 * C::kSplits; ++i) mbar_init(&bar_act[i], 8);     // one arrival per epilogue warp of the tile, both CTAs
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 3 * F; i += kThreads)
        head_w[i] = i < 2 * F ? __ldg(net.ph_w + i) : __ldg(net.vh_w + (i - 2 * F));
    if (warp == kComputeWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc->tmem_base)),
                     "r"((uint32_t)C::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // barriers of both CTAs are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = misc->tmem_base;

    if (warp < kComputeWarps) {
        // ===================== epilogue / input =====================
        const int tile = warp >> 2;
        const int m = ((warp & 3) << 5) | lane;
        const int tt = threadIdx.x & 127;
        uint4* bufA = reinterpret_cast<uint4*>(smem + C::offA + tile * C::kTileBytes);
        uint4* bufB = reinterpret_cast<uint4*>(smem + C::offB + tile * C::kTileBytes);
        zero_tile_buffer(bufA, C::KC, tt, 128);
        zero_tile_buffer(bufB, C::KC, tt, 128);
        // activations of BOTH CTAs are announced on the leader's barriers
        const uint32_t act0 = mapa_u32(smem_u32(&bar_act[tile * C::kSplits]), 0);
        uint32_t acc_phase = 0;
        uint32_t layer_count = 0;
        const float ph_b0 = __ldg(net.ph_b), ph_b1 = __ldg(net.ph_b + 1), vh_b = __ldg(net.vh_b);
        for (int64_t it = 0; it < n_rounds; ++it) {
            const int64_t item = blockIdx.x + it * gridDim.x;
            named_bar_sync(kBarAll, kComputeWarps * 32);
            if (threadIdx.x < 4) {
                const int64_t b = item * 4 + threadIdx.x;
                const uint64_t a = b < n ? self_b[b] : 0ULL, o = b < n ? opp_b[b] : 0ULL;
                misc->s_self[threadIdx.x] = a; misc->s_opp[threadIdx.x] = o; misc->s_legal[it & 1][threadIdx.x] = legal_moves(a, o);
            }
            named_bar_sync(kBarAll, kComputeWarps * 32);
            build_input_row(bufA, m, misc->s_self + 2 * tile, misc->s_opp + 2 * tile, misc->s_legal[it & 1] + 2 * tile);
            bufA[unit_of_row(1, m)] = make_uint4(0, 0, 0, 0);
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < C::kSplits; ++q) mbar_arrive_cluster(act0 + 8u * q);
            }
            for (int layer = 0; layer < n_layers; ++layer, ++layer_count) {
                const bool into_b = (layer == 0) || ((layer & 1) == 0);
                const bool skip = layer > 0 && (layer & 1) == 0;
                const bool last = layer + 1 == n_layers;
                uint4* out = into_b ? bufB : bufA;
                float* bias_s = reinterpret_cast<float*>(smem + C::offBias) + (tile * 2 + (layer & 1)) * F;
                if (tt < F) bias_s[tt] = __ldg(net.bias + (size_t)layer * F + tt);
                if ((warp & 3) == 0) mbar_wait(&bar_acc[tile], acc_phase);          // one mbarrier watcher per tile (net_tc.cu)
                named_bar_sync(kBarTile + tile, 128);
                acc_phase ^= 1;
                tc_fence_after();
                if (net.trace && blockIdx.x == 0 && tt == 0) net.trace[layer * 8 + 2 + 2 * tile] = clock64();
                const uint32_t tcol = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((layer_count & 1) * 2 * F + tile * F);
                float hp[3] = {0.f, 0.f, 0.f};
                uint32_t r[2][32];
                tmem_ld32_nowait(tcol, r[0]);
                tmem_wait_ld(r[0]);
#pragma unroll
                for (int q = 0; q < C::kSplits; ++q) {
                    if (q + 1 < C::kSplits) tmem_ld32_nowait(tcol + (uint32_t)((q + 1) * 32), r[(q + 1) & 1]);
                    if (last) {
                        if (skip) epilogue_chunk<true, true>(r[q & 1], q, m, bias_s, bufB, out, head_w, F, hp);
                        else epilogue_chunk<false, true>(r[q & 1], q, m, bias_s, bufB, out, head_w, F, hp);
                    } else if (skip) epilogue_chunk<true, false>(r[q & 1], q, m, bias_s, bufB, out, head_w, F, hp);
                    else epilogue_chunk<false, false>(r[q & 1], q, m, bias_s, bufB, out, head_w, F, hp);
                    if (!last) {
                        fence_async_proxy();                  // generic-proxy stores -> visible to the tensor core
                        __syncwarp();                         // one (remote) arrival per warp: arrivals serialise on the barrier
                        if (lane == 0) mbar_arrive_cluster(act0 + 8u * q);
                    }
                    if (q + 1 < C::kSplits) tmem_wait_ld(r[(q + 1) & 1]);
                }
                tc_fence_before();
                if (last) {
                    HeadScratch* hs = reinterpret_cast<HeadScratch*>(smem + C::offHeads) + tile;
                    const int b = (m >> 3) & 1, sq = ((m >> 4) << 3) | (m & 7);
                    named_bar_sync(kBarHeadFree + tile, 160);
                    hs->pol_in[b][sq] = fmaxf(hp[0] + ph_b0, 0.f);
                    hs->pol_in[b][64 + sq] = fmaxf(hp[1] + ph_b1, 0.f);
                    hs->val_in[b][sq] = fmaxf(hp[2] + vh_b, 0.f);
                    named_bar_arrive(kBarHeadFull + tile, 160);
                }
                if (net.trace && blockIdx.x == 0 && tt == 0) net.trace[layer * 8 + 3 + 2 * tile] = clock64();
            }
        }
    } else if (warp == kComputeWarps) {
        // ===================== weight producer: this CTA's half of every stage =====================
        if (lane == 0) {
            unsigned char* ring = smem + C::offRing;
            uint32_t cnt = 0;
            for (int64_t it = 0; it < n_rounds; ++it) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(net.w_tc2);
                for (int layer = 0; layer < n_layers; ++layer) {
                    const int stages = layer == 0 ? C::kStages : C::kStagesPerConv;
                    for (int s = 0; s < stages; ++s, ++cnt) {
                        // global layout per stage: [CTA 0's bytes][CTA 1's bytes]; stem stages alternate two taps / one tap,
                        // the padding stages of the stem's ring round carry nothing
                        uint32_t bytes = C::kStageBytes;
                        if (layer == 0) bytes = s >= 6 ? 0u : ((s & 1) ? C::kStemTapBytes : 2 * C::kStemTapBytes);
                        const uint32_t slot = cnt % C::kStages, round = cnt / C::kStages;
                        mbar_wait(&bar_empty[slot], (round & 1) ^ 1);
                        if (bytes) {
                            mbar_expect_tx(&bar_full[slot], bytes);
                            bulk_g2s(ring + slot * C::kRingSlotBytes, src + rank * bytes, bytes, &bar_full[slot]);
                        } else {
                            mbar_arrive(&bar_full[slot]);
                        }
                        src += 2 * bytes;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= kHeadWarp0) {
        // ===================== head warps =====================
        const int tile = warp - kHeadWarp0;
        HeadScratch* hs = reinterpret_cast<HeadScratch*>(smem + C::offHeads) + tile;
        named_bar_arrive(kBarHeadFree + tile, 160);
        for (int64_t it = 0; it < n_rounds; ++it) {
            const int64_t item = blockIdx.x + it * gridDim.x;
            named_bar_sync(kBarHeadFull + tile, 160);
            heads_tail_warp(net, hs, misc->s_legal[it & 1] + 2 * tile, item * 4 + 2 * tile, n, policy_out, value_out, out_kind, lane);
            __syncwarp();
            if (it + 1 < n_rounds) named_bar_arrive(kBarHeadFree + tile, 160);
        }
    } else if (rank == 0) {
        // ===================== MMA issuer (leader CTA, for the pair) =====================
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        uint32_t round = 0, act_phase = 0, layer_count = 0;
        for (int64_t it = 0; it < n_rounds; ++it) {
            for (int layer = 0; layer < n_layers; ++layer, ++layer_count) {
                const bool from_a = (layer == 0) || ((layer & 1) == 0);
                const uint32_t in_off = from_a ? (uint32_t)C::offA : (uint32_t)C::offB;
                const uint32_t d_col = tmem_u + (layer_count & 1) * 2 * F;
                if (net.trace && blockIdx.x == 0 && lane == 0) net.trace[layer * 8 + 0] = clock64();
                if (layer == 0) issue_layer2<F, true>(smem_base, in_off, d_col, bar_full, bar_empty, bar_acc, bar_act, act_phase, round);
                else issue_layer2<F, false>(smem_base, in_off, d_col, bar_full, bar_empty, bar_acc, bar_act, act_phase, round);
                act_phase ^= 1;
                if (net.trace && blockIdx.x == 0 && lane == 0) net.trace[layer * 8 + 1] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ===================== relay (peer CTA): my half of the stage has landed -> tell the leader =====================
        const uint32_t full0 = mapa_u32(smem_u32(&bar_full[0]), 0);
        uint32_t cnt = 0;
        for (int64_t it = 0; it < n_rounds; ++it) {
            for (int layer = 0; layer < n_layers; ++layer) {
                const int stages = layer == 0 ? C::kStages : C::kStagesPerConv;
                for (int s = 0; s < stages; ++s, ++cnt) {
                    const uint32_t slot = cnt % C::kStages, round = cnt / C::kStages;
                    mbar_wait(&bar_full[slot], round & 1);
                    if (lane == 0) mbar_arrive_cluster(full0 + 8u * slot);
                    __syncwarp();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // the peer's shared memory / TMEM stay alive until the pair is done
    if (warp == kComputeWarps + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
    }
}

}  // namespace tc

template <int F>
static int launch_tc2(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value, int out_kind,
                      const int32_t* n_dev, int grid)
{
    using C = tc::Cfg2<F>;
    oth_ctx* ctx = net->ctx;
    OTH_CHECK_CUDA(cudaFuncSetAttribute(tc::k_net_tc2<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(tc::kThreads); cfg.dynamicSmemBytes = C::kSmemBytes; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    OTH_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc::k_net_tc2<F>, net->dev, self_b, opp_b, n, policy, value, out_kind, n_dev));
    return OTH_OK;
}

int net_forward_tc2(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                    int out_kind, const int32_t* n_dev)
{
    oth_ctx* ctx = net->ctx;
    OTH_REQUIRE(net_tc_supported(net->F), OTH_ERR_UNSUPPORTED, "tcgen05 engine supports num_filters 64 or 128 (got %d)", net->F);
    const int64_t items = (n + 3) / 4;
    int grid = (int)(items < ctx->sm_count ? items : ctx->sm_count);
    grid = (grid + 1) & ~1;                                  // whole pairs
    if (grid > ctx->sm_count) grid = ctx->sm_count & ~1;
    if (grid < 2) grid = 2;
    int rc = net->F == 128 ? launch_tc2<128>(net, self_b, opp_b, n, policy, value, out_kind, n_dev, grid)
                           : launch_tc2<64>(net, self_b, opp_b, n, policy, value, out_kind, n_dev, grid);
    if (rc != OTH_OK) return rc;
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    return OTH_OK;
}

}  // namespace oth
