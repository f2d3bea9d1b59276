// net_tc.cu -- OthelloResNet trunk as bf16 tcgen05 implicit-GEMM 3x3 convolutions (sm_100a).
//
// One persistent CTA per SM walks over "items" of 4 boards (2 tiles x 128 GEMM rows).  The
// whole network runs inside the kernel: activations never leave shared memory, accumulators
// live in TMEM, and only the BN-folded bf16 weights stream in (L2-resident, 1-D bulk-TMA
// copies of two 8 KB stages each into a 6-stage ring, each stage consumed by both tiles).
//
//   warps 0-3  : tile 0 epilogue (TMEM -> +bias/+skip/ReLU -> bf16 -> shared), input planes; the last layer's
//                epilogue also applies the three 1x1 head convolutions to the row it holds in registers
//   warps 4-7  : tile 1 epilogue, same
//   warp  8    : weight producer (one lane issues cp.async.bulk + mbarrier expect_tx)
//   warp  9    : MMA issuer (one lane issues tcgen05.mma / tcgen05.commit), owns the TMEM allocation
//   warps 10,11: head warps (one per tile): policy fc + log-softmax + mask, value fc1/fc2 + tanh, output stores --
//                they run while the tensor core is already busy with the NEXT item's trunk
//
// A 3x3 tap is not materialised (no im2col): the activation tile is stored K-major without
// swizzle (net_common.cuh) so that tap (dy,dx) is the SAME UMMA shared-memory descriptor with
// its start address moved by (dy*18+dx) 16-byte units; zero pad units / halo groups supply the
// padding=1 zeros.  GEMM per conv and tile: M=128 (2 boards), N=F, K=9*F in steps of 16.
//
// Restates src/model/net.py:15-61,139-205 (eval mode, BN folded) -- numerics: bf16 operands,
// fp32 accumulation in TMEM, bf16 activations between layers, fp32 heads.
#include <stdlib.h>

#include "common.cuh"
#include "net_common.cuh"
#include "net_host.cuh"
#include "tc_ptx.cuh"
#include "net_tc_common.cuh"

namespace oth {
namespace tc {

#ifndef OTH_TC_STAGES
#define OTH_TC_STAGES 6
#endif
#ifndef OTH_TC_COPY_SPLIT
#define OTH_TC_COPY_SPLIT 1
#endif
constexpr int kStages = OTH_TC_STAGES;          // weight ring depth (8 KB slots); must divide 6 (stem) and 18 (one issue trip)
constexpr int kCopySplit = OTH_TC_COPY_SPLIT;   // bulk copies per ring stage (experiment knob: requests in flight per byte)
#ifndef OTH_TC_STAGE_GROUP
#define OTH_TC_STAGE_GROUP 2
#endif
// Consecutive ring stages fetched by ONE bulk copy.  A bulk copy costs ~52 cycles + bytes / 35.4 B/clk on one SM (measured:
// 2 KB copies 110 cycles, 8 KB copies 283), and with 8 KB per copy the weight stream (283 cycles per stage), not the tensor
// core (256 cycles per stage for two tiles), sets the layer time.  Weights lie in streaming order, so a group is contiguous.
// Two stages (16 KB) per copy: +1.7 % on the kernel alone, +1.0 % on the full self-play step (A/B on one box, alternating).
constexpr int kGroup = OTH_TC_STAGE_GROUP;
constexpr int kGroups = kStages / kGroup;       // barrier pairs: one bar_full / bar_empty per group
static_assert(kStages % kGroup == 0 && 18 % kGroup == 0 && (kGroup == 1 || kGroup == 2), "stage groups: 1 or 2 stages per copy");
static_assert(6 % kStages == 0 && 18 % kStages == 0, "ring depth must divide the stem's 6 stages and a trip's 18");

template <int F>
struct Cfg {
    static constexpr int KC = F / 8;                        // 8-channel planes
    static constexpr int kSplits = F / kSplitChannels;      // channel splits per conv (4 for F=128, 2 for F=64)
    static constexpr int kPlanesPerSplit = kSplitChannels / 8;
    static constexpr int kStagesPerConv = 9 * kSplits;      // one ring stage = one (split, tap)
    static constexpr int kMmasPerStage = kSplitChannels / 16;
    static constexpr int kTileBytes = tile_buffer_bytes(KC);
    static constexpr int kStageBytes = kPlanesPerSplit * F * 16;  // one tap, one split: [planes][F rows][8 ch]
    static constexpr int kStemTapBytes = 2 * F * 16;        // one stem tap: 2 planes (K padded to 16)
    static constexpr int kTmemCols = 4 * F;                 // 2 tiles x 2 accumulator buffers (layer parity)
    // shared-memory map
    static constexpr int offA = 0;                          // A[2] : conv1 output h / network input
    static constexpr int offB = 2 * kTileBytes;             // B[2] : residual stream x
    static constexpr int offRing = 4 * kTileBytes;
    static constexpr int kRingSlotBytes = kStageBytes;      // a stem stage is two taps or one: 2 * kStemTapBytes == kStageBytes
    static constexpr int offHeads = offRing + kStages * kRingSlotBytes;
    static constexpr int offBars = offHeads + 2 * (int)sizeof(HeadScratch);
    static constexpr int kNumBars = 2 * kStages + 2 + 2 * kSplits;
    static constexpr int offMisc = offBars + kNumBars * 8;
    static constexpr int offBias = offMisc + 144;           // [tile][layer parity][F] fp32 bias of the layer in flight
    static constexpr int offHeadW = offBias + 4 * F * 4;    // [3][F] fp32: policy 1x1 (2 channels) and value 1x1 weights
    static constexpr int kSmemBytes = offHeadW + 3 * F * 4;
    static_assert(2 * kStemTapBytes == kStageBytes, "stem stages reuse the trunk's ring slots");
    static_assert(kStagesPerConv % 18 == 0 && kSplits % 2 == 0, "issue loop: one trip = two splits = 18 stages = whole ring rounds");
    static_assert(offBias % 16 == 0 && offHeadW % 16 == 0, "bias / head-weight staging must be 16-byte aligned");
    static_assert(sizeof(HeadScratch) % 16 == 0, "HeadScratch must keep 16-byte alignment");
    static_assert(kRingSlotBytes % 128 == 0, "ring slots must stay 128-byte aligned");
    static_assert(kSmemBytes <= 232448, "shared-memory budget (227 KB) exceeded");
    static_assert(kTmemCols <= 512, "TMEM has 512 columns");
};

// Issue every MMA of one convolution for both tiles.  Called by all 32 lanes of the MMA warp: control flow and
// operands are warp-uniform (they live in uniform registers), only the elected lane issues.  Each operand
// descriptor is "base + immediate" in 16-byte units (lo word = start>>4 | LBO>>4 << 16, hi word constant; every
// shared-memory address is below 2^18, so the 14-bit start field needs no masking).
// K order of a trunk conv: 32-channel split 0 for all nine taps, then split 1, ...: the MMAs over a split start
// as soon as the previous layer's epilogue has written those 32 channels (bar_act[tile * kSplits + split]).
// Loop shape: one trip = two splits = 18 stages = three ring rounds, unrolled inside (ring slots, tap shifts and
// barrier phases are immediates), rolled outside.
template <int F, bool STEM, int TILES>
__device__ __forceinline__ void issue_layer(uint32_t smem_base, uint32_t in_off, uint32_t d_col, uint64_t* bar_full,
                                            uint64_t* bar_empty, uint64_t* bar_acc, uint64_t* bar_act, uint32_t act_phase,
                                            uint32_t& round)
{
    using C = Cfg<F>;
    constexpr uint32_t idesc = umma_idesc(F);
    constexpr uint32_t kAHi = (uint32_t)(kGroupUnits) | (1u << 14);                 // SBO = 144 B, version 1
    constexpr uint32_t kBHi = (uint32_t)(128 >> 4) | (1u << 14);                    // SBO = 128 B
    constexpr uint32_t kALboField = (uint32_t)kPlaneUnits << 16;                    // LBO = plane stride
    constexpr uint32_t kBLboField = (uint32_t)F << 16;                              // LBO = F rows x 16 B
    constexpr uint32_t kSlotUnits = (uint32_t)(C::kRingSlotBytes >> 4);
    constexpr uint32_t kTileUnits = (uint32_t)(C::kTileBytes >> 4);
    // unit index of row 0, plane 0, tile 0, with the LBO field folded in
    const uint32_t a_row0 = (((smem_base + in_off) >> 4) + kGuardUnits + kHaloUnits) | kALboField;
    const uint32_t ring0 = ((smem_base + (uint32_t)C::offRing) >> 4) | kBLboField;
    if (STEM) {
#pragma unroll
        for (int i = 0; i < TILES * C::kSplits; ++i) mbar_wait(&bar_act[i], act_phase);
#pragma unroll
        for (int s = 0; s < 6; ++s) {                       // stage = (tap row dy = s/2 - 1, part): part 0 = taps dx -1,0; part 1 = dx +1
            constexpr int kTapUnits = C::kStemTapBytes >> 4;
            const int dy = s / 2, part = s % 2;
            if ((s % kStages) % kGroup == 0) mbar_wait(&bar_full[(s % kStages) / kGroup], (round + s / kStages) & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int tile = 0; tile < TILES; ++tile) {
#pragma unroll
                    for (int k = 0; k < (part == 0 ? 2 : 1); ++k) {
                        const int tx = part == 0 ? k : 2;
                        const uint32_t a_u = a_row0 + (uint32_t)(tile * kTileUnits + (dy - 1) * 2 * kGroupUnits + (tx - 1));
                        const uint32_t b_u = ring0 + (uint32_t)((s % kStages) * kSlotUnits + k * kTapUnits);
                        umma_bf16(d_col + (uint32_t)(tile * F), ((uint64_t)kAHi << 32) | a_u, ((uint64_t)kBHi << 32) | b_u, idesc,
                                  (s > 0 || k > 0) ? 1u : 0u);
                    }
                    if (s == 5) umma_commit(&bar_acc[tile]);
                }
                if ((s % kStages) % kGroup == kGroup - 1) umma_commit(&bar_empty[(s % kStages) / kGroup]);
            }
            __syncwarp();
        }
        round += 6 / kStages;
    } else {
#pragma unroll 1
        for (int trip = 0; trip < C::kSplits / 2; ++trip) {
            const uint32_t a_trip = a_row0 + (uint32_t)(trip * 2 * C::kPlanesPerSplit * kPlaneUnits);
#pragma unroll
            for (int t = 0; t < 18; ++t) {
                const int ql = t / 9, tap = t % 9, sl = t % kStages;
                if (tap == 0) {                                                     // first tap of a channel split
                    mbar_wait(&bar_act[0 * C::kSplits + trip * 2 + ql], act_phase); // tile 0
                    if (TILES > 1) mbar_wait(&bar_act[1 * C::kSplits + trip * 2 + ql], act_phase); // tile 1
                }
                if (sl % kGroup == 0) mbar_wait(&bar_full[sl / kGroup], (round + t / kStages) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const int shift = (tap / 3 - 1) * 2 * kGroupUnits + (tap % 3 - 1);
#pragma unroll
                    for (int tile = 0; tile < TILES; ++tile) {
#pragma unroll
                        for (int j = 0; j < C::kMmasPerStage; ++j) {
                            const uint32_t a_u = a_trip + (uint32_t)(tile * kTileUnits + (ql * C::kPlanesPerSplit + 2 * j) * kPlaneUnits + shift);
                            const uint32_t b_u = ring0 + (uint32_t)(sl * kSlotUnits + 2 * j * F);
                            umma_bf16(d_col + (uint32_t)(tile * F), ((uint64_t)kAHi << 32) | a_u, ((uint64_t)kBHi << 32) | b_u, idesc,
                                      (t > 0 || j > 0) ? 1u : (trip > 0 ? 1u : 0u));
                        }
                        if (t == 17 && trip == C::kSplits / 2 - 1) umma_commit(&bar_acc[tile]);   // this tile's accumulator is complete
                    }
                    if (sl % kGroup == kGroup - 1) umma_commit(&bar_empty[sl / kGroup]);    // slot(s) free once both tiles have read them
                }
                __syncwarp();
            }
            round += 18 / kStages;
        }
    }
}

// TILES = 2: the throughput shape (4 boards per item, every weight stage feeds two tiles).  TILES = 1: the latency shape
// for small batches (2 boards per item, twice as many CTAs busy, half the MMAs per layer on each): what a tick of a
// 100-game campaign needs.  Same instructions per board in the same order, so outputs are identical bit for bit.
// Tried and removed (B200, 10x128, 18,944 positions): thread-block clusters whose CTAs each fetch 1/CL of every ring stage
// and multicast it to the cluster (cp.async.bulk ... multicast::cluster).  CL = 2: +2.4 % (+7.5 % with 16 KB copies), CL = 4:
// half speed (cluster placement); a lone tile was not faster (the ~30 B/clk an SM's shared memory takes from bulk copies is
// on the receiving side), and the rank-1 CTA of every cluster returned wrong, run-to-run different outputs (a position-level
// diff showed exactly the odd items off): its ring barriers let it run ahead of the data, so even the gain was not real.
template <int F, int TILES>
__global__ void __launch_bounds__(kThreads, 1)
k_net_tc(const NetDev net, const uint64_t* __restrict__ self_b, const uint64_t* __restrict__ opp_b, int64_t n,
         float* __restrict__ policy_out, float* __restrict__ value_out, int out_kind, const int32_t* __restrict__ n_dev,
         const __grid_constant__ CUtensorMap tmap_w, const int use_tmap, const int64_t n_min)
{
    if (n_dev) { const int64_t nd = *n_dev; n = nd < n ? nd : n; }      // batch size decided on the device
    if (n < n_min) return;                                               // the latency shape (launched before) has taken this batch
    using C = Cfg<F>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::offBars);
    uint64_t* bar_full = bars;                       // [kStages] weights landed
    uint64_t* bar_empty = bars + kStages;            // [kStages] slot consumed by the tensor core
    uint64_t* bar_acc = bars + 2 * kStages;          // [2] accumulator tile complete
    uint64_t* bar_act = bars + 2 * kStages + 2;      // [tile*kSplits + split] 32 channels of the tile written (4 arrivals, one per warp)
    float* head_w = reinterpret_cast<float*>(smem + C::offHeadW);
    Misc* misc = reinterpret_cast<Misc*>(smem + C::offMisc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_layers = 1 + 2 * net.blocks;
    constexpr int kItemBoards = 2 * TILES;
    const int64_t n_items = (n + kItemBoards - 1) / kItemBoards;
    // items this CTA walks: its own share, or (clusters) the same count for everybody
    const int64_t my_rounds = n_items > blockIdx.x ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; ++i) mbar_init(&bar_acc[i], 1);
        for (int i = 0; i < 2 * C::kSplits; ++i) mbar_init(&bar_act[i], 4);       // one arrival per epilogue warp of the tile
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 3 * F; i += kThreads)
        head_w[i] = i < 2 * F ? __ldg(net.ph_w + i) : __ldg(net.vh_w + (i - 2 * F));
    if (warp == kComputeWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc->tmem_base)),
                     "r"((uint32_t)C::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = misc->tmem_base;

    if (warp < kComputeWarps && (warp >> 2) < TILES) {
        // ===================== epilogue / input / heads =====================
        const int tile = warp >> 2;
        const int m = ((warp & 3) << 5) | lane;          // GEMM row == TMEM lane
        const int tt = threadIdx.x & 127;                 // thread index within the tile's 128 threads
        uint4* bufA = reinterpret_cast<uint4*>(smem + C::offA + tile * C::kTileBytes);
        uint4* bufB = reinterpret_cast<uint4*>(smem + C::offB + tile * C::kTileBytes);
        zero_tile_buffer(bufA, C::KC, tt, 128);
        zero_tile_buffer(bufB, C::KC, tt, 128);
        uint32_t acc_phase = 0;
        uint32_t layer_count = 0;                         // accumulator buffer parity runs across items
        uint32_t it = 0;                                  // items done by this CTA
        const float ph_b0 = __ldg(net.ph_b), ph_b1 = __ldg(net.ph_b + 1), vh_b = __ldg(net.vh_b);
        for (; it < my_rounds; ++it) {
            const int64_t item = blockIdx.x + (int64_t)it * gridDim.x;
            named_bar_sync(kBarAll, TILES * 128);               // everybody is done reading the previous item's misc->s_self/opp
            if (threadIdx.x < kItemBoards) {
                const int64_t b = item * kItemBoards + threadIdx.x;
                const uint64_t a = b < n ? self_b[b] : 0ULL, o = b < n ? opp_b[b] : 0ULL;
                misc->s_self[threadIdx.x] = a; misc->s_opp[threadIdx.x] = o; misc->s_legal[it & 1][threadIdx.x] = legal_moves(a, o);
            }
            named_bar_sync(kBarAll, TILES * 128);
            build_input_row(bufA, m, misc->s_self + 2 * tile, misc->s_opp + 2 * tile, misc->s_legal[it & 1] + 2 * tile);
            bufA[unit_of_row(1, m)] = make_uint4(0, 0, 0, 0);   // K padding plane of the stem
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < C::kSplits; ++q) mbar_arrive(&bar_act[tile * C::kSplits + q]);
            }
            for (int layer = 0; layer < n_layers; ++layer, ++layer_count) {
                const bool into_b = (layer == 0) || ((layer & 1) == 0);   // stem and conv2 write the residual stream
                const bool skip = layer > 0 && (layer & 1) == 0;          // conv2: add the block input
                const bool last = layer + 1 == n_layers;
                uint4* out = into_b ? bufB : bufA;
                // stage this layer's bias in shared memory while the tensor core is still busy
                float* bias_s = reinterpret_cast<float*>(smem + C::offBias) + (tile * 2 + (layer & 1)) * F;
                if (tt < F) bias_s[tt] = __ldg(net.bias + (size_t)layer * F + tt);
                // One warp per tile watches the mbarrier and releases the other three through a named barrier: warps
                // parked in mbarrier.try_wait are not free -- every additional long-term waiter measurably slows the
                // MMA / weight pipeline of the whole SM, whereas bar.sync waiters cost nothing.
                if ((warp & 3) == 0) mbar_wait(&bar_acc[tile], acc_phase);
                named_bar_sync(kBarTile + tile, 128);
                acc_phase ^= 1;
                tc_fence_after();
                if (net.trace && blockIdx.x == 0 && tt == 0) net.trace[layer * 8 + 2 + 2 * tile] = clock64();
                const uint32_t tcol = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((layer_count & 1) * 2 * F + tile * F);
                float hp[3] = {0.f, 0.f, 0.f};
                // 32 channels at a time: the load of the next split is in flight while this one is converted, and the next
                // layer's MMAs over a split are released as soon as it is in shared memory
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
                        __syncwarp();                         // one arrival per warp: arrivals serialise on the barrier
                        if (lane == 0) mbar_arrive(&bar_act[tile * C::kSplits + q]);
                    }
                    if (q + 1 < C::kSplits) tmem_wait_ld(r[(q + 1) & 1]);
                }
                tc_fence_before();
                if (last) {
                    // hand the 1x1-conv outputs (ReLU'd, flattened channel-major, net.py:90) to this tile's head warp
                    HeadScratch* hs = reinterpret_cast<HeadScratch*>(smem + C::offHeads) + tile;
                    const int b = (m >> 3) & 1, sq = ((m >> 4) << 3) | (m & 7);
                    named_bar_sync(kBarHeadFree + tile, 160);                // previous item's heads are done with the scratch
                    hs->pol_in[b][sq] = fmaxf(hp[0] + ph_b0, 0.f);
                    hs->pol_in[b][64 + sq] = fmaxf(hp[1] + ph_b1, 0.f);
                    hs->val_in[b][sq] = fmaxf(hp[2] + vh_b, 0.f);
                    named_bar_arrive(kBarHeadFull + tile, 160);
                }
                if (net.trace && blockIdx.x == 0 && tt == 0) net.trace[layer * 8 + 3 + 2 * tile] = clock64();
            }
        }
    } else if (warp < kComputeWarps) {
        // tile 1's epilogue warps have nothing to do in the one-tile shape
    } else if (warp == kComputeWarps) {
        // ===================== weight producer =====================
        if (lane == 0) {
            unsigned char* ring = smem + C::offRing;
            uint32_t cnt = 0;
            for (int64_t r = 0; r < my_rounds; ++r) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(net.w_tc);
                for (int layer = 0; layer < n_layers; ++layer) {
                    const int groups = (layer == 0 ? 6 : C::kStagesPerConv) / kGroup;
                    for (int s = 0; s < groups; ++s, ++cnt) {
                        // stem stages alternate two taps / one tap (taps dx = -1,0 then dx = +1 of a tap row): a stem stage
                        // pair is 8 KB + 4 KB, contiguous, and the 4 KB land at the second slot's base like any second stage
                        uint32_t bytes = kGroup * C::kStageBytes;
                        if (layer == 0) bytes = kGroup == 2 ? 3 * C::kStemTapBytes : ((s & 1) ? C::kStemTapBytes : 2 * C::kStemTapBytes);
                        const uint32_t slot = cnt % kGroups, round = cnt / kGroups;
                        mbar_wait(&bar_empty[slot], (round & 1) ^ 1);
                        mbar_expect_tx(&bar_full[slot], bytes);
                        if (use_tmap && layer > 0) {
                            // trunk stages through the tensor map: one box = this stage (group), rows of 256 bytes
                            tma_load_2d(ring + slot * kGroup * C::kRingSlotBytes, &tmap_w, 0,
                                        (int32_t)((src - reinterpret_cast<const unsigned char*>(net.w_tc)) >> 8), &bar_full[slot]);
                        } else {
#pragma unroll
                            for (int part = 0; part < kCopySplit; ++part)
                                bulk_g2s(ring + slot * kGroup * C::kRingSlotBytes + part * (bytes / kCopySplit), src + part * (bytes / kCopySplit),
                                         bytes / kCopySplit, &bar_full[slot]);
                        }
                        src += bytes;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= kHeadWarp0 + TILES) {
        // no second tile, no second head warp
    } else if (warp >= kHeadWarp0) {
        // ===================== head warps =====================
        const int tile = warp - kHeadWarp0;
        HeadScratch* hs = reinterpret_cast<HeadScratch*>(smem + C::offHeads) + tile;
        uint32_t it = 0;
        // named barriers, not mbarriers (see the note at the accumulator wait): 128 epilogue threads + this warp
        named_bar_arrive(kBarHeadFree + tile, 160);                           // the scratch starts out free
        for (; it < my_rounds; ++it) {
            const int64_t item = blockIdx.x + (int64_t)it * gridDim.x;
            named_bar_sync(kBarHeadFull + tile, 160);
            heads_tail_warp(net, hs, misc->s_legal[it & 1] + 2 * tile, item * kItemBoards + 2 * tile, n, policy_out, value_out, out_kind, lane);
            __syncwarp();
            if ((int64_t)it + 1 < my_rounds) named_bar_arrive(kBarHeadFree + tile, 160);
        }
    } else {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop (warp-uniform control flow and operands, so descriptors live in
        // uniform registers); one elected lane issues tcgen05.mma / tcgen05.commit.
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        uint32_t round = 0, act_phase = 0, layer_count = 0;      // ring round: every layer uses whole ring rounds
        for (int64_t r = 0; r < my_rounds; ++r) {
            for (int layer = 0; layer < n_layers; ++layer, ++layer_count) {
                const bool from_a = (layer == 0) || ((layer & 1) == 0);   // stem reads the input, conv2 reads h: both in A
                const uint32_t in_off = from_a ? (uint32_t)C::offA : (uint32_t)C::offB;
                const uint32_t d_col = tmem_u + (layer_count & 1) * 2 * F;  // accumulator buffer of this layer
                if (net.trace && blockIdx.x == 0 && lane == 0) net.trace[layer * 8 + 0] = clock64();
                if (layer == 0) issue_layer<F, true, TILES>(smem_base, in_off, d_col, bar_full, bar_empty, bar_acc, bar_act, act_phase, round);
                else issue_layer<F, false, TILES>(smem_base, in_off, d_col, bar_full, bar_empty, bar_acc, bar_act, act_phase, round);
                act_phase ^= 1;
                if (net.trace && blockIdx.x == 0 && lane == 0) net.trace[layer * 8 + 1] = clock64();
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kComputeWarps + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
    }
}

}  // namespace tc

bool net_tc_supported(int F) { return F == 64 || F == 128; }

int net_tc_stage_rows(int F) { return tc::kGroup * (F == 128 ? tc::Cfg<128>::kStageBytes : tc::Cfg<64>::kStageBytes) / 256; }

template <int F, int TILES>
static int launch_tc(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value, int out_kind,
                     const int32_t* n_dev, int64_t n_min)
{
    using C = tc::Cfg<F>;
    oth_ctx* ctx = net->ctx;
    const int64_t items = (n + 2 * TILES - 1) / (2 * TILES);
    int grid = (int)(items < ctx->sm_count ? items : ctx->sm_count);
    if (grid < 1) grid = 1;
    OTH_CHECK_CUDA(cudaFuncSetAttribute(tc::k_net_tc<F, TILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    static const bool env_tmap = getenv("OTH_TC_TMAP") != nullptr;
    const int use_tmap = (env_tmap && net->tmap_ok) ? 1 : 0;
    tc::k_net_tc<F, TILES><<<grid, tc::kThreads, C::kSmemBytes, ctx->stream>>>(net->dev, self_b, opp_b, n, policy, value, out_kind, n_dev,
                                                                            net->tmap_w, use_tmap, n_min);
    return OTH_OK;
}

int net_forward_tc(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                   int out_kind, const int32_t* n_dev, int64_t n_min)
{
    oth_ctx* ctx = net->ctx;
    OTH_REQUIRE(net_tc_supported(net->F), OTH_ERR_UNSUPPORTED, "tcgen05 engine supports num_filters 64 or 128 (got %d)", net->F);
    // Tiles per CTA: two (4 boards per item) is the shape everything runs on.  The one-tile shape does not shorten a layer
    // (measured: 10.2 k cycles per layer with one tile or two -- a layer lasts as long as its 288 KB of weights take to enter
    // one SM's shared memory, ~30 B/clk); it is kept for experiments (OTH_TC_ONE_TILE=1, batches that fit one item per SM).
    static const bool env_one_tile = getenv("OTH_TC_ONE_TILE") != nullptr;
    const bool one_tile = env_one_tile && (n + 1) / 2 <= ctx->sm_count;
    int rc;
    if (net->F == 128) rc = one_tile ? launch_tc<128, 1>(net, self_b, opp_b, n, policy, value, out_kind, n_dev, n_min)
                                     : launch_tc<128, 2>(net, self_b, opp_b, n, policy, value, out_kind, n_dev, n_min);
    else rc = one_tile ? launch_tc<64, 1>(net, self_b, opp_b, n, policy, value, out_kind, n_dev, n_min)
                       : launch_tc<64, 2>(net, self_b, opp_b, n, policy, value, out_kind, n_dev, n_min);
    if (rc) return rc;
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    return OTH_OK;
}

}  // namespace oth
