// net_tc.cu -- OthelloResNet trunk as bf16 tcgen05 implicit-GEMM 3x3 convolutions (sm_100a).
//
// One persistent CTA per SM walks over "items" of 4 boards (2 tiles x 128 GEMM rows).  The
// whole network runs inside the kernel: activations never leave shared memory, accumulators
// live in TMEM, and only the BN-folded bf16 weights stream in (L2-resident, 1-D bulk-TMA
// copies into a 3-slot ring, each slot consumed by both tiles).
//
//   warps 0-3 : tile 0 epilogue (TMEM -> +bias/+skip/ReLU -> bf16 -> shared), input planes, heads
//   warps 4-7 : tile 1 epilogue, same
//   warp  8   : weight producer (one lane issues cp.async.bulk + mbarrier expect_tx)
//   warp  9   : MMA issuer (one lane issues tcgen05.mma / tcgen05.commit), owns the TMEM allocation
//
// A 3x3 tap is not materialised (no im2col): the activation tile is stored K-major without
// swizzle (net_common.cuh) so that tap (dy,dx) is the SAME UMMA shared-memory descriptor with
// its start address moved by (dy*18+dx) 16-byte units; zero pad units / halo groups supply the
// padding=1 zeros.  GEMM per conv and tile: M=128 (2 boards), N=F, K=9*F in steps of 16.
//
// Restates src/model/net.py:15-61,139-205 (eval mode, BN folded) -- numerics: bf16 operands,
// fp32 accumulation in TMEM, bf16 activations between layers, fp32 heads.
#include "common.cuh"
#include "net_common.cuh"
#include "net_host.cuh"
#include "tc_ptx.cuh"

namespace oth {
namespace tc {

constexpr int kComputeWarps = 8;
constexpr int kThreads = (kComputeWarps + 2) * 32;   // 320
constexpr int kStages = 3;
constexpr int kStagesPerConv = 18;                   // 2 channel halves x 9 taps

template <int F>
struct Cfg {
    static constexpr int KC = F / 8;                        // 8-channel planes
    static constexpr int kPlanesPerHalf = KC / 2;
    static constexpr int kMmasPerStage = F / 32;            // K = F/2 per (half, tap), 16 per MMA
    static constexpr int kTileBytes = tile_buffer_bytes(KC);
    static constexpr int kStageBytes = kPlanesPerHalf * F * 16;   // one tap, one channel half: [planes][F rows][8 ch]
    static constexpr int kStemStageBytes = 3 * 2 * F * 16;  // 3 taps x 2 planes (K padded to 16)
    static constexpr int kTmemCols = 4 * F;                 // 2 tiles x 2 accumulator buffers (layer parity)
    // shared-memory map
    static constexpr int offA = 0;                          // A[2] : conv1 output h / network input
    static constexpr int offB = 2 * kTileBytes;             // B[2] : residual stream x
    static constexpr int offRing = 4 * kTileBytes;
    static constexpr int kRingSlotBytes = kStageBytes > kStemStageBytes ? kStageBytes : kStemStageBytes;
    static constexpr int offHeads = offRing + kStages * kRingSlotBytes;
    static constexpr int offBars = offHeads + 2 * (int)sizeof(HeadScratch);
    static constexpr int kNumBars = 2 * kStages + 2 + 4;
    static constexpr int offMisc = offBars + kNumBars * 8;
    static constexpr int offBias = offMisc + 128;           // [tile][layer parity][F] fp32 bias of the layer in flight
    static constexpr int kSmemBytes = offBias + 4 * F * 4;
    static_assert(offBias % 16 == 0, "bias staging must be 16-byte aligned");
    static_assert(sizeof(HeadScratch) % 16 == 0, "HeadScratch must keep 16-byte alignment");
    static_assert(kRingSlotBytes % 128 == 0, "ring slots must stay 128-byte aligned");
    static_assert(kSmemBytes <= 232448, "shared-memory budget (227 KB) exceeded");
    static_assert(kTmemCols <= 512, "TMEM has 512 columns");
};

struct Misc {
    uint32_t tmem_base;
    uint32_t pad;
    uint64_t s_self[4], s_opp[4], s_legal[4];
};

// Issue every MMA of one convolution for both tiles.  Fully unrolled: each operand descriptor is
// "base + immediate" in 16-byte units (lo word = start>>4 | LBO>>4 << 16, hi word constant).
// Called by all 32 lanes of the MMA warp (warp-uniform operands live in uniform registers); only the
// elected lane issues.  K order of a trunk conv: channel half 0 for all nine taps, then half 1, so the
// MMAs of half 0 can start as soon as the previous layer's epilogue has written channels [0, F/2).
template <int F, bool STEM>
__device__ __forceinline__ void issue_layer(uint32_t smem_base, uint32_t in_off, uint32_t d_col, uint64_t* bar_full,
                                            uint64_t* bar_empty, uint64_t* bar_acc, uint64_t* bar_act, uint32_t act_phase,
                                            uint32_t& round)
{
    using C = Cfg<F>;
    constexpr uint32_t idesc = umma_idesc(F);
    constexpr int kStagesHere = STEM ? 3 : kStagesPerConv;
    static_assert(kStagesHere % kStages == 0, "a layer must use whole ring rounds");
    constexpr uint32_t kAHi = (uint32_t)(kGroupUnits) | (1u << 14);                 // SBO = 144 B, version 1
    constexpr uint32_t kBHi = (uint32_t)(128 >> 4) | (1u << 14);                    // SBO = 128 B
    constexpr uint32_t kALboField = (uint32_t)kPlaneUnits << 16;                    // LBO = plane stride
    constexpr uint32_t kBLboField = (uint32_t)F << 16;                              // LBO = F rows x 16 B
    const uint32_t a_row0 = ((smem_base + in_off) >> 4) + kGuardUnits + kHaloUnits; // unit index of row 0, plane 0, tile 0
    const uint32_t ring0 = (smem_base + (uint32_t)C::offRing) >> 4;
#pragma unroll
    for (int s = 0; s < kStagesHere; ++s) {
        const int slot = s % kStages;
        if (s > 0 && slot == 0) ++round;
        if (STEM) {
            if (s == 0) {
                mbar_wait(&bar_act[0], act_phase); mbar_wait(&bar_act[1], act_phase);
                mbar_wait(&bar_act[2], act_phase); mbar_wait(&bar_act[3], act_phase);
            }
        } else if (s % 9 == 0) {                                                    // first tap of a channel half
            mbar_wait(&bar_act[0 + s / 9], act_phase);                              // tile 0, this half
            mbar_wait(&bar_act[2 + s / 9], act_phase);                              // tile 1, this half
        }
        mbar_wait(&bar_full[slot], round & 1);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t wb = ring0 + (uint32_t)slot * (C::kRingSlotBytes >> 4);
#pragma unroll
            for (int tile = 0; tile < 2; ++tile) {
                const uint32_t d = d_col + (uint32_t)(tile * F);
                const uint32_t a_tile = a_row0 + (uint32_t)tile * (C::kTileBytes >> 4);
                if (STEM) {
#pragma unroll
                    for (int t3 = 0; t3 < 3; ++t3) {
                        const int tap = s * 3 + t3;
                        const int shift = (tap / 3 - 1) * 2 * kGroupUnits + (tap % 3 - 1);
                        const uint64_t ad = ((uint64_t)kAHi << 32) | (uint64_t)(((a_tile + shift) & 0x3FFFu) | kALboField);
                        const uint64_t bd = ((uint64_t)kBHi << 32) | (uint64_t)(((wb + t3 * 2 * F) & 0x3FFFu) | kBLboField);
                        umma_bf16(d, ad, bd, idesc, tap > 0 ? 1u : 0u);
                    }
                } else {
                    const int half = s / 9, tap = s % 9;
                    const int shift = (tap / 3 - 1) * 2 * kGroupUnits + (tap % 3 - 1);
#pragma unroll
                    for (int j = 0; j < C::kMmasPerStage; ++j) {
                        const int kc = half * C::kPlanesPerHalf + 2 * j;
                        const uint64_t ad = ((uint64_t)kAHi << 32) | (uint64_t)(((a_tile + kc * kPlaneUnits + shift) & 0x3FFFu) | kALboField);
                        const uint64_t bd = ((uint64_t)kBHi << 32) | (uint64_t)(((wb + 2 * j * F) & 0x3FFFu) | kBLboField);
                        umma_bf16(d, ad, bd, idesc, (s > 0 || j > 0) ? 1u : 0u);
                    }
                }
                if (s == kStagesHere - 1) umma_commit(&bar_acc[tile]);   // this tile's accumulator is complete
            }
            umma_commit(&bar_empty[slot]);                                // slot free once both tiles have read it
        }
        __syncwarp();
    }
    ++round;
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// 32 accumulator columns of one GEMM row -> +bias (+skip) -> ReLU -> bf16 -> four 16-byte stores
template <bool SKIP>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], int chunk, int m, const float* bias_s,
                                               const uint4* __restrict__ resid, uint4* __restrict__ out)
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int kc = chunk * 4 + q;
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + kc * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + kc * 8 + 4);
        v[0] = __uint_as_float(r[q * 8 + 0]) + b0.x; v[1] = __uint_as_float(r[q * 8 + 1]) + b0.y;
        v[2] = __uint_as_float(r[q * 8 + 2]) + b0.z; v[3] = __uint_as_float(r[q * 8 + 3]) + b0.w;
        v[4] = __uint_as_float(r[q * 8 + 4]) + b1.x; v[5] = __uint_as_float(r[q * 8 + 5]) + b1.y;
        v[6] = __uint_as_float(r[q * 8 + 6]) + b1.z; v[7] = __uint_as_float(r[q * 8 + 7]) + b1.w;
        const int u = unit_of_row(kc, m);
        if (SKIP) {
            const uint4 x = resid[u];
            v[0] += bf16_lo(x.x); v[1] += bf16_hi(x.x); v[2] += bf16_lo(x.y); v[3] += bf16_hi(x.y);
            v[4] += bf16_lo(x.z); v[5] += bf16_hi(x.z); v[6] += bf16_lo(x.w); v[7] += bf16_hi(x.w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        out[u] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

template <int F>
__global__ void __launch_bounds__(kThreads, 1)
k_net_tc(const NetDev net, const uint64_t* __restrict__ self_b, const uint64_t* __restrict__ opp_b, int64_t n,
         float* __restrict__ policy_out, float* __restrict__ value_out, int out_kind, const int32_t* __restrict__ n_dev)
{
    if (n_dev) { const int64_t nd = *n_dev; n = nd < n ? nd : n; }      // batch size decided on the device
    using C = Cfg<F>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::offBars);
    uint64_t* bar_full = bars;                       // [kStages] weights landed
    uint64_t* bar_empty = bars + kStages;            // [kStages] slot consumed by the tensor core
    uint64_t* bar_acc = bars + 2 * kStages;          // [2] accumulator tile complete
    uint64_t* bar_act = bars + 2 * kStages + 2;      // [tile*2 + half] activation half-tile written (128 arrivals)
    Misc* misc = reinterpret_cast<Misc*>(smem + C::offMisc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_layers = 1 + 2 * net.blocks;
    const int64_t n_items = (n + 3) / 4;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; ++i) mbar_init(&bar_acc[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&bar_act[i], 128);
        fence_barrier_init();
    }
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

    if (warp < kComputeWarps) {
        // ===================== epilogue / input / heads =====================
        const int tile = warp >> 2;
        const int m = ((warp & 3) << 5) | lane;          // GEMM row == TMEM lane
        const int tt = threadIdx.x & 127;                 // thread index within the tile's 128 threads
        uint4* bufA = reinterpret_cast<uint4*>(smem + C::offA + tile * C::kTileBytes);
        uint4* bufB = reinterpret_cast<uint4*>(smem + C::offB + tile * C::kTileBytes);
        HeadScratch* hs = reinterpret_cast<HeadScratch*>(smem + C::offHeads) + tile;
        zero_tile_buffer(bufA, C::KC, tt, 128);
        zero_tile_buffer(bufB, C::KC, tt, 128);
        uint32_t acc_phase = 0;
        uint32_t layer_count = 0;                         // accumulator buffer parity runs across items
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            named_bar_sync(1, kComputeWarps * 32);        // previous item's heads are done with misc->s_*
            if (threadIdx.x < 4) {
                const int64_t b = item * 4 + threadIdx.x;
                const uint64_t a = b < n ? self_b[b] : 0ULL, o = b < n ? opp_b[b] : 0ULL;
                misc->s_self[threadIdx.x] = a; misc->s_opp[threadIdx.x] = o; misc->s_legal[threadIdx.x] = legal_moves(a, o);
            }
            named_bar_sync(1, kComputeWarps * 32);
            build_input_row(bufA, m, misc->s_self + 2 * tile, misc->s_opp + 2 * tile, misc->s_legal + 2 * tile);
            bufA[unit_of_row(1, m)] = make_uint4(0, 0, 0, 0);   // K padding plane of the stem
            fence_async_proxy();
            mbar_arrive(&bar_act[tile * 2 + 0]);
            mbar_arrive(&bar_act[tile * 2 + 1]);
            for (int layer = 0; layer < n_layers; ++layer, ++layer_count) {
                const bool into_b = (layer == 0) || ((layer & 1) == 0);   // stem and conv2 write the residual stream
                const bool skip = layer > 0 && (layer & 1) == 0;          // conv2: add the block input
                uint4* out = into_b ? bufB : bufA;
                // stage this layer's bias in shared memory while the tensor core is still busy
                float* bias_s = reinterpret_cast<float*>(smem + C::offBias) + (tile * 2 + (layer & 1)) * F;
                if (tt < F) bias_s[tt] = __ldg(net.bias + (size_t)layer * F + tt);
                named_bar_sync(2 + tile, 128);
                mbar_wait(&bar_acc[tile], acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                if (net.trace && blockIdx.x == 0 && tt == 0) net.trace[layer * 8 + 2 + 2 * tile] = clock64();
                const uint32_t tcol = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((layer_count & 1) * 2 * F + tile * F);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    // channels [half*F/2, (half+1)*F/2): load them all, then convert; the next layer's MMAs over this
                    // channel half are released as soon as it is in shared memory
                    constexpr int kChunks = F / 64;            // 32-column chunks per half
                    uint32_t r[kChunks][32];
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) tmem_ld32_nowait(tcol + (uint32_t)((half * kChunks + c) * 32), r[c]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) {
                        if (skip) epilogue_chunk<true>(r[c], half * kChunks + c, m, bias_s, bufB, out);
                        else epilogue_chunk<false>(r[c], half * kChunks + c, m, bias_s, bufB, out);
                    }
                    tc_fence_before();
                    if (layer + 1 < n_layers) {
                        fence_async_proxy();                  // generic-proxy stores -> visible to the tensor core
                        mbar_arrive(&bar_act[tile * 2 + half]);
                    }
                }
                if (net.trace && blockIdx.x == 0 && tt == 0) net.trace[layer * 8 + 3 + 2 * tile] = clock64();
            }
            named_bar_sync(2 + tile, 128);                // final activations of this tile are complete
            heads_for_tile(net, bufB, hs, misc->s_legal + 2 * tile, item * 4 + 2 * tile, n, policy_out, value_out, out_kind, tt,
                           128, [tile] { named_bar_sync(2 + tile, 128); });
        }
    } else if (warp == kComputeWarps) {
        // ===================== weight producer =====================
        if (lane == 0) {
            unsigned char* ring = smem + C::offRing;
            uint32_t cnt = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(net.w_tc);
                for (int layer = 0; layer < n_layers; ++layer) {
                    const int stages = layer == 0 ? 3 : kStagesPerConv;
                    const uint32_t bytes = layer == 0 ? C::kStemStageBytes : C::kStageBytes;
                    for (int s = 0; s < stages; ++s, ++cnt) {
                        const uint32_t slot = cnt % kStages, round = cnt / kStages;
                        mbar_wait(&bar_empty[slot], (round & 1) ^ 1);
                        mbar_expect_tx(&bar_full[slot], bytes);
                        bulk_g2s(ring + slot * C::kRingSlotBytes, src, bytes, &bar_full[slot]);
                        src += bytes;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop (warp-uniform control flow and operands, so descriptors live in
        // uniform registers); one elected lane issues tcgen05.mma / tcgen05.commit.
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        uint32_t round = 0, act_phase = 0, layer_count = 0;      // ring round: every layer uses a multiple of kStages stages
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            for (int layer = 0; layer < n_layers; ++layer, ++layer_count) {
                const bool from_a = (layer == 0) || ((layer & 1) == 0);   // stem reads the input, conv2 reads h: both in A
                const uint32_t in_off = from_a ? (uint32_t)C::offA : (uint32_t)C::offB;
                const uint32_t d_col = tmem_u + (layer_count & 1) * 2 * F;  // accumulator buffer of this layer
                if (net.trace && blockIdx.x == 0 && lane == 0) net.trace[layer * 8 + 0] = clock64();
                if (layer == 0) issue_layer<F, true>(smem_base, in_off, d_col, bar_full, bar_empty, bar_acc, bar_act, act_phase, round);
                else issue_layer<F, false>(smem_base, in_off, d_col, bar_full, bar_empty, bar_acc, bar_act, act_phase, round);
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

int net_forward_tc(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                   int out_kind, const int32_t* n_dev)
{
    oth_ctx* ctx = net->ctx;
    OTH_REQUIRE(net_tc_supported(net->F), OTH_ERR_UNSUPPORTED, "tcgen05 engine supports num_filters 64 or 128 (got %d)", net->F);
    const int64_t items = (n + 3) / 4;
    int grid = (int)(items < ctx->sm_count ? items : ctx->sm_count);
    if (grid < 1) grid = 1;
    if (net->F == 128) {
        using C = tc::Cfg<128>;
        OTH_CHECK_CUDA(cudaFuncSetAttribute(tc::k_net_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
        tc::k_net_tc<128><<<grid, tc::kThreads, C::kSmemBytes, ctx->stream>>>(net->dev, self_b, opp_b, n, policy, value, out_kind, n_dev);
    } else {
        using C = tc::Cfg<64>;
        OTH_CHECK_CUDA(cudaFuncSetAttribute(tc::k_net_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
        tc::k_net_tc<64><<<grid, tc::kThreads, C::kSmemBytes, ctx->stream>>>(net->dev, self_b, opp_b, n, policy, value, out_kind, n_dev);
    }
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    return OTH_OK;
}

}  // namespace oth
