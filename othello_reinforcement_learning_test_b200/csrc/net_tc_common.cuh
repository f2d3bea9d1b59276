// net_tc_common.cuh -- pieces shared by the tcgen05 trunk kernels (net_tc.cu: one CTA per SM;
// net_tc2.cu: CTA pairs, cta_group::2): warp roles, named barriers, the epilogue conversion of one
// 32-channel accumulator chunk and the head warp (policy fc / softmax / mask, value fc / tanh).
#pragma once
#include "net_common.cuh"
#include "tc_ptx.cuh"

namespace oth {
namespace tc {

constexpr int kComputeWarps = 8;
constexpr int kHeadWarp0 = kComputeWarps + 2;        // first head warp
constexpr int kThreads = (kComputeWarps + 4) * 32;   // 384
// named (hardware) barriers: 0 is __syncthreads
constexpr int kBarAll = 1;        // the 8 epilogue warps
constexpr int kBarTile = 2;       // +tile: the 4 epilogue warps of a tile
constexpr int kBarHeadFull = 4;   // +tile: epilogue warps arrive, head warp syncs -- 1x1-conv outputs are in HeadScratch
constexpr int kBarHeadFree = 6;   // +tile: head warp arrives, epilogue warps sync -- HeadScratch may be overwritten
constexpr int kSplitChannels = 32;                   // K walked as (32-channel split, tap): the next layer starts on a split
                                                     // as soon as the epilogue has written those 32 channels


struct Misc {
    uint32_t tmem_base;
    uint32_t pad;
    uint64_t s_self[4], s_opp[4];
    uint64_t s_legal[2][4];      // by item parity: the head warps still need the previous item's masks
};
static_assert(sizeof(Misc) <= 144, "Misc slot");

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

// tcgen05.wait::ld, tied to the registers of the load it completes so that no use can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// 32 accumulator columns of one GEMM row -> +bias (+skip) -> ReLU -> bf16 -> four 16-byte stores.
// LAST (final trunk layer): instead of storing, feed the bf16-rounded activations to the three 1x1 head
// convolutions (net.py:85,121), channels in ascending order, accumulators carried in hp[3].
template <bool SKIP, bool LAST>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], int chunk, int m, const float* bias_s,
                                               const uint4* __restrict__ resid, uint4* __restrict__ out,
                                               const float* __restrict__ head_w, int F, float (&hp)[3])
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
        const uint4 packed = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        if (LAST) {
            const float x[8] = {bf16_lo(packed.x), bf16_hi(packed.x), bf16_lo(packed.y), bf16_hi(packed.y),
                                bf16_lo(packed.z), bf16_hi(packed.z), bf16_lo(packed.w), bf16_hi(packed.w)};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = kc * 8 + j;
                hp[0] = fmaf(x[j], head_w[c], hp[0]);
                hp[1] = fmaf(x[j], head_w[F + c], hp[1]);
                hp[2] = fmaf(x[j], head_w[2 * F + c], hp[2]);
            }
        } else {
            out[u] = packed;
        }
    }
}

// The part of the heads after the 1x1 convolutions (net.py:86-94,122-134) for the two boards of one tile, run by
// ONE warp: policy fc (128 -> 65), log-softmax (+exp, +action mask and renormalisation), value fc1 (64 -> 256, ReLU),
// fc2 (256 -> 1), tanh.  Same operation order per output as heads_for_tile (net_common.cuh), so both engines
// return identical bits.  Lane l owns policy outputs l, l+32 (and 64, computed by every lane) and hidden units l+32r.
__device__ __forceinline__ void heads_tail_warp(const NetDev& net, HeadScratch* hs, const uint64_t* s_legal, int64_t board0,
                                                int64_t n_boards, float* __restrict__ policy_out,
                                                float* __restrict__ value_out, int out_kind, int lane)
{
    float pa[kBoardsPerTile][3];
    {
        const float b0 = __ldg(net.pfc_b + lane), b1 = __ldg(net.pfc_b + 32 + lane), b2 = __ldg(net.pfc_b + 64);
#pragma unroll
        for (int b = 0; b < kBoardsPerTile; ++b) { pa[b][0] = b0; pa[b][1] = b1; pa[b][2] = b2; }
    }
#pragma unroll 8
    for (int i = 0; i < 128; ++i) {
        const float w0 = __ldg(net.pfc_t + i * 65 + lane), w1 = __ldg(net.pfc_t + i * 65 + 32 + lane), w2 = __ldg(net.pfc_t + i * 65 + 64);
#pragma unroll
        for (int b = 0; b < kBoardsPerTile; ++b) {
            const float x = hs->pol_in[b][i];
            pa[b][0] = fmaf(x, w0, pa[b][0]); pa[b][1] = fmaf(x, w1, pa[b][1]); pa[b][2] = fmaf(x, w2, pa[b][2]);
        }
    }
    float ha[kBoardsPerTile][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const float bb = __ldg(net.v1_b + 32 * r + lane);
#pragma unroll
        for (int b = 0; b < kBoardsPerTile; ++b) ha[b][r] = bb;
    }
#pragma unroll 4
    for (int i = 0; i < 64; ++i) {
        float w[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) w[r] = __ldg(net.v1_t + i * 256 + 32 * r + lane);
#pragma unroll
        for (int b = 0; b < kBoardsPerTile; ++b) {
            const float x = hs->val_in[b][i];
#pragma unroll
            for (int r = 0; r < 8; ++r) ha[b][r] = fmaf(x, w[r], ha[b][r]);
        }
    }
    float v2w[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v2w[r] = __ldg(net.v2_w + 32 * r + lane);
    const float v2b = __ldg(net.v2_b);
#pragma unroll
    for (int b = 0; b < kBoardsPerTile; ++b) {
        const int64_t board = board0 + b;
        // log_softmax over the 65 logits (net.py:94)
        float mx = fmaxf(fmaxf(-INFINITY, pa[b][0]), pa[b][1]);
        if (lane == 0) mx = fmaxf(mx, pa[b][2]);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float se = 0.f;
        se += expf(pa[b][0] - mx);
        se += expf(pa[b][1] - mx);
        if (lane == 0) se += expf(pa[b][2] - mx);
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xFFFFFFFFu, se, o);
        const float lse = logf(se);
        float out3[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float logp = (pa[b][q] - mx) - lse;
            out3[q] = (out_kind == kOutLogProbs) ? logp : expf(logp);
        }
        // value head: fc1 ReLU, fc2, tanh (net.py:128-134)
        float part = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) part = fmaf(fmaxf(ha[b][r], 0.f), v2w[r], part);
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        if (out_kind == kOutPriors) {
            // MCTSNode.expand's masking (node.py:71-80) with numpy's float32 summation order, spread over the warp
            const uint64_t legal = s_legal[b];
            const bool pass_only = legal == 0;
            if (pass_only) { out3[0] = 0.f; out3[1] = 0.f; }
            else {
                if (!((legal >> lane) & 1ULL)) out3[0] = 0.f;
                if (!((legal >> (32 + lane)) & 1ULL)) out3[1] = 0.f;
                out3[2] = 0.f;
            }
            // np_sum65: accumulator k (k = 0..7) adds p[k], p[8+k], ..., p[56+k]; p[j] lives in lane j&31, slot j>>5
            float racc = 0.f;
#pragma unroll
            for (int i = 0; i < 64; i += 8) {
                const float v = __shfl_sync(0xFFFFFFFFu, (i & 32) ? out3[1] : out3[0], (i & 31) + (lane & 7));
                racc = (i == 0) ? v : __fadd_rn(racc, v);
            }
            float rr[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) rr[k] = __shfl_sync(0xFFFFFFFFu, racc, k);
            float total = __fadd_rn(__fadd_rn(__fadd_rn(rr[0], rr[1]), __fadd_rn(rr[2], rr[3])),
                                    __fadd_rn(__fadd_rn(rr[4], rr[5]), __fadd_rn(rr[6], rr[7])));
            total = __fadd_rn(total, __shfl_sync(0xFFFFFFFFu, out3[2], 0));
            if (total > 0.f) {
#pragma unroll
                for (int q = 0; q < 3; ++q) out3[q] = __fdiv_rn(out3[q], total);
            } else {
                const float u = (float)(1.0 / (double)(pass_only ? 1 : popc64(legal)));
                if (pass_only) out3[2] = u;
                else {
                    if ((legal >> lane) & 1ULL) out3[0] = u;
                    if ((legal >> (32 + lane)) & 1ULL) out3[1] = u;
                }
            }
        }
        if (board < n_boards) {
            policy_out[board * 65 + lane] = out3[0];
            policy_out[board * 65 + 32 + lane] = out3[1];
            if (lane == 0) { policy_out[board * 65 + 64] = out3[2]; value_out[board] = tanhf(part + v2b); }
        }
    }
}

}  // namespace tc
}  // namespace oth
