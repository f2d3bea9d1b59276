// net_common.cuh -- shared-memory activation layout, input-plane builder and the fused
// policy/value heads used by both trunk engines (tcgen05 in net_tc.cu, CUDA-core
// validation kernel in net.cu).  Restates src/model/net.py:64-136 (heads) and
// src/cython/bitboard.pyx:300-323 (input planes) for eval-mode inference.
//
// Activation layout ("tile" = 2 boards = 128 GEMM rows)
// ------------------------------------------------------
// unit   = 16 bytes = 8 consecutive channels (bf16) of one square
// group  = one board row of one board: 8 units + 1 zero pad unit (so a +-1 shift in x
//          lands on a zero instead of wrapping into the neighbouring row)
// plane  = all squares of the tile for one 8-channel slice kc:
//          [2 halo groups (zero)] [16 groups, g = y*2 + board] ; the next plane's leading
//          halo doubles as this plane's trailing halo, so a +-1 shift in y (= +-2 groups)
//          also lands on zeros.
// A GEMM row m (0..127) is square (y = m>>4, x = m&7) of board (m>>3)&1 and lives at unit
//   kGuardUnits + kc*kPlaneUnits + kHaloUnits + (m>>3)*9 + (m&7).
// Rows are 16 bytes apart inside a group and groups 144 bytes apart, which is exactly a
// K-major, no-swizzle UMMA operand with SBO = 144 B and LBO = plane stride; a 3x3 tap
// (dy,dx) is the same descriptor with its start address moved by (dy*18 + dx) units.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "bitboard.cuh"

namespace oth {

constexpr int kGroupUnits = 9;
constexpr int kTileGroups = 16;
constexpr int kHaloUnits = 2 * kGroupUnits;                        // 18
constexpr int kPlaneUnits = kHaloUnits + kTileGroups * kGroupUnits; // 162
constexpr int kGuardUnits = 8;
constexpr int kTileRows = 128;
constexpr int kBoardsPerTile = 2;

__host__ __device__ constexpr int tile_buffer_units(int kc_planes)
{
    return kGuardUnits + kc_planes * kPlaneUnits + kHaloUnits + 6;   // +6: round to a multiple of 8 units below
}
__host__ __device__ constexpr int tile_buffer_bytes(int kc_planes)
{
    return ((tile_buffer_units(kc_planes) + 7) / 8) * 8 * 16;
}
__host__ __device__ __forceinline__ int unit_of_row(int kc, int m)
{
    return kGuardUnits + kc * kPlaneUnits + kHaloUnits + (m >> 3) * kGroupUnits + (m & 7);
}

struct NetDev {
    int blocks, F, KC;
    const __nv_bfloat16* w_tc;   // UMMA B tiles: stem [9][2][F][8], then per conv [9][KC][F][8] (bf16, BN folded)
    const __nv_bfloat16* w_tc2;  // the same tiles split by cout half for CTA pairs (net_tc2.cu): per stage [CTA 0's rows][CTA 1's rows]
    const float* w_simt;         // same weights widened to fp32: stem [9][8][F], then per conv [9][F(cin)][F(cout)]
    const float* bias;           // [1 + 2*blocks][F] folded BN bias
    const float* ph_w;           // [2][F]  policy 1x1 conv, BN folded (fp32)
    const float* ph_b;           // [2]
    const float* pfc_t;          // [128][65] policy fc weight, transposed
    const float* pfc_b;          // [65]
    const float* vh_w;           // [F]     value 1x1 conv, BN folded
    const float* vh_b;           // [1]
    const float* v1_t;           // [64][256] value fc1 weight, transposed
    const float* v1_b;           // [256]
    const float* v2_w;           // [256]
    const float* v2_b;           // [1]
    unsigned long long* trace;   // diagnostics: per-layer clock64 stamps of CTA 0 (NULL = off)
};

// out_kind values mirror OTH_NET_OUT_* in include/othello_b200.h
constexpr int kOutLogProbs = 0, kOutProbs = 1, kOutPriors = 2;

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Zero a tile buffer (halo, pads and all) -- call once; afterwards only real squares are written.
__device__ __forceinline__ void zero_tile_buffer(uint4* buf, int kc_planes, int t, int nthreads)
{
    const int n = tile_buffer_bytes(kc_planes) / 16;
    for (int i = t; i < n; i += nthreads) buf[i] = make_uint4(0, 0, 0, 0);
}

// get_tensor_input (bitboard.pyx:300-323) written straight into plane 0 of a tile buffer:
// channels 0,1,2 = self, opp, REF-legal; channels 3..7 of the unit stay zero.
// One thread per GEMM row.  `legal_tile[2]` (shared) receives the legal masks for the heads.
__device__ __forceinline__ void build_input_row(uint4* buf, int m, const uint64_t* s_self, const uint64_t* s_opp,
                                                const uint64_t* s_legal)
{
    const int b = (m >> 3) & 1, sq = ((m >> 4) << 3) | (m & 7);
    const float a = (float)((s_self[b] >> sq) & 1ULL), o = (float)((s_opp[b] >> sq) & 1ULL),
                l = (float)((s_legal[b] >> sq) & 1ULL);
    buf[unit_of_row(0, m)] = make_uint4(pack_bf16x2(a, o), pack_bf16x2(l, 0.f), 0u, 0u);
}

// numpy's float32 add-reduce order for 65 contiguous values (pairwise, 8 accumulators):
// what `masked_probs.sum()` does in node.py:75.
__device__ __forceinline__ float np_sum65(const float* a)
{
    float r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
#pragma unroll
    for (int i = 8; i < 64; i += 8) {
        r0 = __fadd_rn(r0, a[i]); r1 = __fadd_rn(r1, a[i + 1]); r2 = __fadd_rn(r2, a[i + 2]); r3 = __fadd_rn(r3, a[i + 3]);
        r4 = __fadd_rn(r4, a[i + 4]); r5 = __fadd_rn(r5, a[i + 5]); r6 = __fadd_rn(r6, a[i + 6]); r7 = __fadd_rn(r7, a[i + 7]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r0, r1), __fadd_rn(r2, r3)), __fadd_rn(__fadd_rn(r4, r5), __fadd_rn(r6, r7)));
    return __fadd_rn(res, a[64]);
}

// MCTSNode.expand's action masking (node.py:71-80): probs[65] -> priors[65] in place.
// legal == 0 means the single legal action is the pass (bitboard.pyx:176-178).
__device__ __forceinline__ void mask_and_renormalise(float* p, uint64_t legal)
{
    int n_legal;
    if (legal == 0) {
        for (int i = 0; i < 64; ++i) p[i] = 0.f;
        n_legal = 1;
    } else {
        for (int i = 0; i < 64; ++i) if (!((legal >> i) & 1ULL)) p[i] = 0.f;
        p[64] = 0.f;
        n_legal = popc64(legal);
    }
    const float total = np_sum65(p);
    if (total > 0.f) {
        for (int i = 0; i < 65; ++i) p[i] = __fdiv_rn(p[i], total);
    } else {
        const float u = (float)(1.0 / (double)n_legal);
        if (legal == 0) p[64] = u;
        else for (int i = 0; i < 64; ++i) if ((legal >> i) & 1ULL) p[i] = u;
    }
}

// Scratch the heads need per tile (lives in whatever shared memory is free at that point).
struct HeadScratch {
    float pol_in[kBoardsPerTile][128];   // relu(bn(conv1x1)) flattened channel-major (net.py:90)
    float val_in[kBoardsPerTile][64];
    float hidden[kBoardsPerTile][256];
    float logits[kBoardsPerTile][68];
};

// PolicyHead + ValueHead (net.py:64-136) for one tile, cooperative over `nthreads`
// threads (t = 0..nthreads-1, nthreads a multiple of 32 and >= 64).  `sync()` must be a
// barrier over exactly those threads.  act = final trunk activations (bf16 tile buffer).
template <class SyncFn>
__device__ __forceinline__ void heads_for_tile(const NetDev& net, const uint4* act, HeadScratch* hs,
                                               const uint64_t* s_legal, int64_t board0, int64_t n_boards,
                                               float* __restrict__ policy_out, float* __restrict__ value_out,
                                               int out_kind, int t, int nthreads, SyncFn sync)
{
    const int F = net.F, KC = net.KC;
    // 1) the three 1x1 convolutions, one GEMM row per loop trip
    for (int m = t; m < kTileRows; m += nthreads) {
        float p0 = 0.f, p1 = 0.f, v0 = 0.f;
        for (int kc = 0; kc < KC; ++kc) {
            const uint4 u = act[unit_of_row(kc, m)];
            const float x[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = kc * 8 + j;
                p0 = fmaf(x[j], __ldg(net.ph_w + c), p0);
                p1 = fmaf(x[j], __ldg(net.ph_w + F + c), p1);
                v0 = fmaf(x[j], __ldg(net.vh_w + c), v0);
            }
        }
        const int b = (m >> 3) & 1, sq = ((m >> 4) << 3) | (m & 7);
        hs->pol_in[b][sq] = fmaxf(p0 + __ldg(net.ph_b), 0.f);
        hs->pol_in[b][64 + sq] = fmaxf(p1 + __ldg(net.ph_b + 1), 0.f);
        hs->val_in[b][sq] = fmaxf(v0 + __ldg(net.vh_b), 0.f);
    }
    sync();
    // 2) policy fc (128 -> 65) and value fc1 (64 -> 256, relu)
    for (int idx = t; idx < kBoardsPerTile * 65; idx += nthreads) {
        const int b = idx / 65, j = idx % 65;
        float acc = __ldg(net.pfc_b + j);
        for (int i = 0; i < 128; ++i) acc = fmaf(hs->pol_in[b][i], __ldg(net.pfc_t + i * 65 + j), acc);
        hs->logits[b][j] = acc;
    }
    for (int idx = t; idx < kBoardsPerTile * 256; idx += nthreads) {
        const int b = idx >> 8, k = idx & 255;
        float acc = __ldg(net.v1_b + k);
        for (int i = 0; i < 64; ++i) acc = fmaf(hs->val_in[b][i], __ldg(net.v1_t + i * 256 + k), acc);
        hs->hidden[b][k] = fmaxf(acc, 0.f);
    }
    sync();
    // 3) one warp per board: log-softmax (+exp, +mask) and value fc2 + tanh
    const int warp = t >> 5, lane = t & 31;
    if (warp < kBoardsPerTile) {
        const int b = warp;
        const int64_t board = board0 + b;
        float* lg = hs->logits[b];
        float mx = -INFINITY;
        for (int j = lane; j < 65; j += 32) mx = fmaxf(mx, lg[j]);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float se = 0.f;
        for (int j = lane; j < 65; j += 32) se += expf(lg[j] - mx);
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xFFFFFFFFu, se, o);
        const float lse = logf(se);
        for (int j = lane; j < 65; j += 32) {
            const float logp = (lg[j] - mx) - lse;
            lg[j] = (out_kind == kOutLogProbs) ? logp : expf(logp);
        }
        float part = 0.f;
        for (int k = lane; k < 256; k += 32) part = fmaf(hs->hidden[b][k], __ldg(net.v2_w + k), part);
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        __syncwarp();
        if (lane == 0 && out_kind == kOutPriors) mask_and_renormalise(lg, s_legal[b]);
        __syncwarp();
        if (board < n_boards) {
            for (int j = lane; j < 65; j += 32) policy_out[board * 65 + j] = lg[j];
            if (lane == 0) value_out[board] = tanhf(part + __ldg(net.v2_b));
        }
    }
}

}  // namespace oth
