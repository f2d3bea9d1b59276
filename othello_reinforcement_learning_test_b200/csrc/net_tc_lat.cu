// net_tc_lat.cu -- the LATENCY shape of the tcgen05 trunk: what a launch over a few hundred positions runs on.
//
// A launch of k_net_tc (net_tc.cu) over a small batch costs the same ~131 us whether it evaluates 2 positions or 592:
// every layer lasts as long as its 288 KB of weights take to enter one SM's shared memory through 8 KB bulk copies
// (~10.2 k cycles, measured), and with few positions nothing else is in the way.  That latency is what a tick of a small
// self-play campaign (BASELINE config 3: 100 games), a single search (MCTSPlayer.get_action) and the arena pay per step.
// This kernel is the same network for the case that the whole batch fits one two-board item per SM:
//   * ONE tile (2 boards) per CTA: twice as many SMs share the batch, and the shared memory the second tile would occupy
//     holds a deeper, coarser weight ring;
//   * trunk weights arrive through a TENSOR-MAP TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) in boxes of 32 KB
//     (four (split, tap) stages of one conv; 12 KB for 64 filters), issued by two lanes in turn.  Measured on B200: a
//     request costs ~260 cycles + bytes / 130 B/clk on the tensor path against ~50 + bytes / 35 B/clk for a 1-D bulk copy,
//     so big boxes are what makes the stream faster than the tensor core needs it (72 MMAs = 4.6 k cycles per layer);
//   * the stem's weights (36 KB) are loaded once per CTA and stay resident, so the ring carries whole convs only.
// MMA order per board, epilogue and heads are those of k_net_tc (shared code in net_tc_common.cuh): outputs are identical
// bit for bit, which the position-keyed evaluation cache relies on (tests/test_gpu_f_net_tc.py compares batch sizes
// across both kernels).
//
// Restates src/model/net.py:15-61,139-205 (eval mode, BN folded) -- numerics as net_tc.cu.
#include <stdlib.h>

#include "common.cuh"
#include "net_common.cuh"
#include "net_host.cuh"
#include "tc_ptx.cuh"
#include "net_tc_common.cuh"

namespace oth {
namespace tc {

constexpr int kLatEpiWarps = 4;                       // warps 0-3: epilogue of the tile (TMEM lane quarter = warp)
constexpr int kLatProducerWarp = 4;                   // weight producer (lanes 0 and 1 issue in turn)
constexpr int kLatMmaWarp = 5;                        // MMA issuer, owns the TMEM allocation
constexpr int kLatHeadWarp = 6;                       // policy / value heads of the previous item
constexpr int kLatThreads = 7 * 32;
constexpr int kLatSlots = 3;                          // ring slots (requests in flight)
constexpr int kLatIssuers = 2;                        // producer lanes that take requests in turn

template <int F>
struct CfgLat {
    static constexpr int KC = F / 8;
    static constexpr int kSplits = F / kSplitChannels;
    static constexpr int kPlanesPerSplit = kSplitChannels / 8;
    static constexpr int kStagesPerConv = 9 * kSplits;                 // (split, tap) stages, 36 or 18
    static constexpr int kMmasPerStage = kSplitChannels / 16;
    static constexpr int kStageBytes = kPlanesPerSplit * F * 16;        // 8 KB / 4 KB
    static constexpr int kGroup = F == 128 ? 4 : 3;                     // stages per TMA request
    static constexpr int kGroupBytes = kGroup * kStageBytes;            // 32 KB / 12 KB
    static constexpr int kReqPerConv = kStagesPerConv / kGroup;         // 9 / 6
    static constexpr int kStemTapBytes = 2 * F * 16;                    // one stem tap: 2 planes (K padded to 16)
    static constexpr int kStemBytes = 9 * kStemTapBytes;
    static constexpr int kTileBytes = tile_buffer_bytes(KC);
    static constexpr int kTmemCols = 2 * F;                             // one tile x two accumulator buffers (layer parity)
    static constexpr int offA = 0;
    static constexpr int offB = kTileBytes;
    static constexpr int offStem = 2 * kTileBytes;
    static constexpr int offRing = offStem + kStemBytes;
    static constexpr int offHeads = offRing + kLatSlots * kGroupBytes;
    static constexpr int offBars = offHeads + (int)sizeof(HeadScratch);
    static constexpr int kNumBars = 2 * kLatSlots + 2 + kSplits;        // full[3], empty[3], acc, stem, act[kSplits]
    static constexpr int offMisc = offBars + ((kNumBars * 8 + 15) / 16) * 16;
    static constexpr int offBias = offMisc + 144;
    static constexpr int offHeadW = offBias + 2 * F * 4;
    static constexpr int kSmemBytes = offHeadW + 3 * F * 4;
    static_assert(kStagesPerConv % kGroup == 0 && kReqPerConv % kLatSlots == 0, "a conv is a whole number of ring rounds");
    static_assert(kGroupBytes % 256 == 0 && kStemBytes % 256 == 0 && kGroupBytes / 256 <= 256, "tensor-map boxes are rows of 256 bytes");
    static_assert(offStem % 128 == 0 && offRing % 128 == 0 && kGroupBytes % 128 == 0, "TMA destinations stay 128-byte aligned");
    static_assert(offBias % 16 == 0 && offHeadW % 16 == 0 && sizeof(HeadScratch) % 16 == 0, "16-byte alignment of the small staging areas");
    static_assert(kSmemBytes <= 232448, "shared-memory budget (227 KB) exceeded");
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation: power of two <= 512 columns");
};

template <int F>
__global__ void __launch_bounds__(kLatThreads, 1)
k_net_lat(const NetDev net, const uint64_t* __restrict__ self_b, const uint64_t* __restrict__ opp_b, int64_t n,
          float* __restrict__ policy_out, float* __restrict__ value_out, int out_kind, const int32_t* __restrict__ n_dev,
          const __grid_constant__ CUtensorMap tmap_w, const int64_t n_max)
{
    if (n_dev) { const int64_t nd = *n_dev; n = nd < n ? nd : n; }      // batch size decided on the device
    if (n > n_max) return;                                               // this batch belongs to the throughput kernel (launched next)
    using C = CfgLat<F>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::offBars);
    uint64_t* bar_full = bars;                          // [kLatSlots] request landed
    uint64_t* bar_empty = bars + kLatSlots;             // [kLatSlots] slot read by the tensor core
    uint64_t* bar_acc = bars + 2 * kLatSlots;           // accumulator of the layer complete
    uint64_t* bar_stem = bars + 2 * kLatSlots + 1;      // resident stem weights landed (once)
    uint64_t* bar_act = bars + 2 * kLatSlots + 2;       // [split] 32 channels of the tile written (4 arrivals, one per warp)
    float* head_w = reinterpret_cast<float*>(smem + C::offHeadW);
    Misc* misc = reinterpret_cast<Misc*>(smem + C::offMisc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_layers = 1 + 2 * net.blocks;
    const int64_t n_items = (n + 1) / 2;
    const int64_t my_items = n_items > blockIdx.x ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kLatSlots; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(bar_acc, 1);
        mbar_init(bar_stem, 1);
        for (int i = 0; i < C::kSplits; ++i) mbar_init(&bar_act[i], 4);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 3 * F; i += kLatThreads)
        head_w[i] = i < 2 * F ? __ldg(net.ph_w + i) : __ldg(net.vh_w + (i - 2 * F));
    if (warp == kLatMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc->tmem_base)),
                     "r"((uint32_t)C::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = misc->tmem_base;

    if (warp < kLatEpiWarps) {
        // ===================== epilogue / input (as k_net_tc, tile 0) =====================
        const int m = (warp << 5) | lane;                 // GEMM row == TMEM lane
        const int tt = threadIdx.x;
        uint4* bufA = reinterpret_cast<uint4*>(smem + C::offA);
        uint4* bufB = reinterpret_cast<uint4*>(smem + C::offB);
        zero_tile_buffer(bufA, C::KC, tt, 128);
        zero_tile_buffer(bufB, C::KC, tt, 128);
        uint32_t acc_phase = 0, layer_count = 0;
        const float ph_b0 = __ldg(net.ph_b), ph_b1 = __ldg(net.ph_b + 1), vh_b = __ldg(net.vh_b);
        for (uint32_t it = 0; it < my_items; ++it) {
            const int64_t item = blockIdx.x + (int64_t)it * gridDim.x;
            named_bar_sync(kBarAll, 128);                 // everybody is done reading the previous item's misc->s_self/opp
            if (threadIdx.x < 2) {
                const int64_t b = item * 2 + threadIdx.x;
                const uint64_t a = b < n ? self_b[b] : 0ULL, o = b < n ? opp_b[b] : 0ULL;
                misc->s_self[threadIdx.x] = a; misc->s_opp[threadIdx.x] = o; misc->s_legal[it & 1][threadIdx.x] = legal_moves(a, o);
            }
            named_bar_sync(kBarAll, 128);
            build_input_row(bufA, m, misc->s_self, misc->s_opp, misc->s_legal[it & 1]);
            bufA[unit_of_row(1, m)] = make_uint4(0, 0, 0, 0);   // K padding plane of the stem
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < C::kSplits; ++q) mbar_arrive(&bar_act[q]);
            }
            for (int layer = 0; layer < n_layers; ++layer, ++layer_count) {
                const bool into_b = (layer == 0) || ((layer & 1) == 0);   // stem and conv2 write the residual stream
                const bool skip = layer > 0 && (layer & 1) == 0;          // conv2: add the block input
                const bool last = layer + 1 == n_layers;
                uint4* out = into_b ? bufB : bufA;
                float* bias_s = reinterpret_cast<float*>(smem + C::offBias) + (layer & 1) * F;
                if (tt < F) bias_s[tt] = __ldg(net.bias + (size_t)layer * F + tt);
                if (warp == 0) mbar_wait(bar_acc, acc_phase);             // one mbarrier watcher, the rest on a named barrier
                named_bar_sync(kBarTile, 128);
                acc_phase ^= 1;
                tc_fence_after();
                if (net.trace && blockIdx.x == 0 && tt == 0) net.trace[layer * 8 + 2] = clock64();
                const uint32_t tcol = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((layer_count & 1) * F);
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
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_act[q]);
                    }
                    if (q + 1 < C::kSplits) tmem_wait_ld(r[(q + 1) & 1]);
                }
                tc_fence_before();
                if (last) {
                    HeadScratch* hs = reinterpret_cast<HeadScratch*>(smem + C::offHeads);
                    const int b = (m >> 3) & 1, sq = ((m >> 4) << 3) | (m & 7);
                    named_bar_sync(kBarHeadFree, 160);                    // previous item's heads are done with the scratch
                    hs->pol_in[b][sq] = fmaxf(hp[0] + ph_b0, 0.f);
                    hs->pol_in[b][64 + sq] = fmaxf(hp[1] + ph_b1, 0.f);
                    hs->val_in[b][sq] = fmaxf(hp[2] + vh_b, 0.f);
                    named_bar_arrive(kBarHeadFull, 160);
                }
                if (net.trace && blockIdx.x == 0 && tt == 0) net.trace[layer * 8 + 3] = clock64();
            }
        }
    } else if (warp == kLatProducerWarp) {
        // ===================== weight producer =====================
        // Lane 0 first brings the stem's weights in (one 1-D bulk copy, resident for the whole kernel).  Requests of the
        // trunk are numbered across items; lanes 0 and 1 take them in turn (the fixed part of a tensor-map request is paid
        // by the issuing thread), each waiting for its own slot.
        if (lane == 0 && my_items > 0) {
            mbar_expect_tx(bar_stem, C::kStemBytes);
            bulk_g2s(smem + C::offStem, net.w_tc, C::kStemBytes, bar_stem);
        }
        if (lane < kLatIssuers) {
            unsigned char* ring = smem + C::offRing;
            const int32_t row0 = C::kStemBytes / 256, rows_per_req = C::kGroupBytes / 256;
            const uint32_t req_per_item = (uint32_t)(n_layers - 1) * C::kReqPerConv;
            for (int64_t it = 0; it < my_items; ++it) {
                for (uint32_t rq = lane; rq < req_per_item; rq += kLatIssuers) {
                    const uint32_t cnt = (uint32_t)it * req_per_item + rq;        // req_per_item is a multiple of kLatSlots
                    const uint32_t slot = cnt % kLatSlots, round = cnt / kLatSlots;
                    mbar_wait(&bar_empty[slot], (round & 1) ^ 1);
                    mbar_expect_tx(&bar_full[slot], C::kGroupBytes);
                    tma_load_2d(ring + slot * C::kGroupBytes, &tmap_w, 0, row0 + (int32_t)rq * rows_per_req, &bar_full[slot]);
                }
            }
        }
        __syncwarp();
    } else if (warp == kLatHeadWarp) {
        // ===================== head warp =====================
        HeadScratch* hs = reinterpret_cast<HeadScratch*>(smem + C::offHeads);
        named_bar_arrive(kBarHeadFree, 160);                              // the scratch starts out free
        for (uint32_t it = 0; it < my_items; ++it) {
            const int64_t item = blockIdx.x + (int64_t)it * gridDim.x;
            named_bar_sync(kBarHeadFull, 160);
            heads_tail_warp(net, hs, misc->s_legal[it & 1], item * 2, n, policy_out, value_out, out_kind, lane);
            __syncwarp();
            if ((int64_t)it + 1 < my_items) named_bar_arrive(kBarHeadFree, 160);
        }
    } else {
        // ===================== MMA issuer =====================
        // Whole warp, warp-uniform control flow and operands, one elected lane issues (as k_net_tc).  K order of a conv:
        // 32-channel split 0 for all nine taps, then split 1, ...; stage st = split * 9 + tap lies in request st / kGroup.
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        constexpr uint32_t idesc = umma_idesc(F);
        constexpr uint32_t kAHi = (uint32_t)(kGroupUnits) | (1u << 14);                 // SBO = 144 B, version 1
        constexpr uint32_t kBHi = (uint32_t)(128 >> 4) | (1u << 14);                    // SBO = 128 B
        constexpr uint32_t kALboField = (uint32_t)kPlaneUnits << 16;                    // LBO = plane stride
        constexpr uint32_t kBLboField = (uint32_t)F << 16;                              // LBO = F rows x 16 B
        constexpr uint32_t kGroupSlotUnits = (uint32_t)(C::kGroupBytes >> 4);
        constexpr uint32_t kStageUnits = (uint32_t)(C::kStageBytes >> 4);
        constexpr uint32_t kTapUnits = (uint32_t)(C::kStemTapBytes >> 4);
        const uint32_t ring0 = ((smem_base + (uint32_t)C::offRing) >> 4) | kBLboField;
        const uint32_t stem0 = ((smem_base + (uint32_t)C::offStem) >> 4) | kBLboField;
        uint32_t round = 0, act_phase = 0, layer_count = 0;
        if (my_items > 0) mbar_wait(bar_stem, 0);
        for (int64_t it = 0; it < my_items; ++it) {
            for (int layer = 0; layer < n_layers; ++layer, ++layer_count) {
                const bool from_a = (layer == 0) || ((layer & 1) == 0);   // stem reads the input, conv2 reads h: both in A
                const uint32_t a_row0 = (((smem_base + (from_a ? (uint32_t)C::offA : (uint32_t)C::offB)) >> 4) + kGuardUnits + kHaloUnits) | kALboField;
                const uint32_t d_col = tmem_u + (layer_count & 1) * F;
                if (net.trace && blockIdx.x == 0 && lane == 0) net.trace[layer * 8 + 0] = clock64();
                if (layer == 0) {
#pragma unroll
                    for (int q = 0; q < C::kSplits; ++q) mbar_wait(&bar_act[q], act_phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {                 // taps ascending: the order k_net_tc accumulates them in
                            const uint32_t a_u = a_row0 + (uint32_t)((tap / 3 - 1) * 2 * kGroupUnits + (tap % 3 - 1));
                            const uint32_t b_u = stem0 + (uint32_t)tap * kTapUnits;
                            umma_bf16(d_col, ((uint64_t)kAHi << 32) | a_u, ((uint64_t)kBHi << 32) | b_u, idesc, tap > 0 ? 1u : 0u);
                        }
                        umma_commit(bar_acc);
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int st = 0; st < C::kStagesPerConv; ++st) {
                        const int q = st / 9, tap = st % 9, rq = st / C::kGroup, in_rq = st % C::kGroup, sl = rq % kLatSlots;
                        if (tap == 0) mbar_wait(&bar_act[q], act_phase);                    // the split's 32 input channels are written
                        if (in_rq == 0) mbar_wait(&bar_full[sl], (round + rq / kLatSlots) & 1);
                        tc_fence_after();
                        if (elect_one()) {
                            const int shift = (tap / 3 - 1) * 2 * kGroupUnits + (tap % 3 - 1);
#pragma unroll
                            for (int j = 0; j < C::kMmasPerStage; ++j) {
                                const uint32_t a_u = a_row0 + (uint32_t)((q * C::kPlanesPerSplit + 2 * j) * kPlaneUnits + shift);
                                const uint32_t b_u = ring0 + (uint32_t)(sl * kGroupSlotUnits + in_rq * kStageUnits + 2 * j * F);
                                umma_bf16(d_col, ((uint64_t)kAHi << 32) | a_u, ((uint64_t)kBHi << 32) | b_u, idesc, (st > 0 || j > 0) ? 1u : 0u);
                            }
                            if (st == C::kStagesPerConv - 1) umma_commit(bar_acc);          // the layer's accumulator is complete
                            if (in_rq == C::kGroup - 1) umma_commit(&bar_empty[sl]);        // the request's slot is free again
                        }
                        __syncwarp();
                    }
                    round += C::kReqPerConv / kLatSlots;
                }
                act_phase ^= 1;
                if (net.trace && blockIdx.x == 0 && lane == 0) net.trace[layer * 8 + 1] = clock64();
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kLatMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
    }
}

}  // namespace tc

int net_tc_lat_box_rows(int F) { return F == 128 ? tc::CfgLat<128>::kGroupBytes / 256 : tc::CfgLat<64>::kGroupBytes / 256; }

template <int F>
static int launch_lat(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value, int out_kind,
                      const int32_t* n_dev)
{
    using C = tc::CfgLat<F>;
    oth_ctx* ctx = net->ctx;
    const int64_t n_max = net_tc_lat_max_positions(net);
    const int64_t items = ((n < n_max ? n : n_max) + 1) / 2;
    int grid = (int)(items < ctx->sm_count ? items : ctx->sm_count);
    if (grid < 1) grid = 1;
    OTH_CHECK_CUDA(cudaFuncSetAttribute(tc::k_net_lat<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    tc::k_net_lat<F><<<grid, tc::kLatThreads, C::kSmemBytes, ctx->stream>>>(net->dev, self_b, opp_b, n, policy, value, out_kind, n_dev,
                                                                        net->tmap_lat, n_max);
    return OTH_OK;
}

// Largest batch the latency shape takes.  One tile per CTA means every CTA pulls the whole weight stream: beyond ~100 busy
// SMs the L2 cannot feed them all at the rate the shape was built for (measured: 296 positions on 148 SMs take 225 us, the
// throughput kernel 138 us), so the shape stops at 128 items = 256 positions (84-87 us measured for 160..256).  0 = not available.
int64_t net_tc_lat_max_positions(const NetHost* net)
{
    static const bool off = getenv("OTH_NO_LATENCY_SHAPE") != nullptr;
    static const int64_t env_max = getenv("OTH_LATENCY_SHAPE_MAX") ? atoll(getenv("OTH_LATENCY_SHAPE_MAX")) : 0;
    if (off || !net->tmap_lat_ok || !net_tc_supported(net->F)) return 0;
    return env_max > 0 ? env_max : 256;
}

int net_forward_tc_lat(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                       int out_kind, const int32_t* n_dev)
{
    oth_ctx* ctx = net->ctx;
    int rc = net->F == 128 ? launch_lat<128>(net, self_b, opp_b, n, policy, value, out_kind, n_dev)
                           : launch_lat<64>(net, self_b, opp_b, n, policy, value, out_kind, n_dev);
    if (rc) return rc;
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    return OTH_OK;
}

}  // namespace oth
