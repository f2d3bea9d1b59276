// net.cu -- OthelloResNet inference (src/model/net.py:139-205, eval mode): weight import
// with BatchNorm folding, engine dispatch, and the CUDA-core validation engine.
//
// The product path is the tcgen05/TMEM implicit-GEMM engine in net_tc.cu; the engine here
// (OTH_NET_ENGINE_SIMT) runs the same layer sequence with the same rounding points
// (bf16 weights, bf16 activations between layers, fp32 accumulation, fp32 heads) on CUDA
// cores and exists to cross-check the tensor-core kernel and to serve filter counts the
// tensor-core tiling does not cover.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "net_common.cuh"
#include "net_host.cuh"

namespace oth {

// ---- weight import ---------------------------------------------------------------------

static inline uint16_t f32_to_bf16_rne(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    const uint32_t lsb = (u >> 16) & 1u;
    u += 0x7FFFu + lsb;
    return (uint16_t)(u >> 16);
}
static inline float bf16_to_f32(uint16_t h)
{
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int64_t net_param_count(int blocks, int F)
{
    int64_t n = 0;
    n += (int64_t)F * 27 + 4 * F;                                  // conv_block
    n += (int64_t)blocks * 2 * ((int64_t)F * F * 9 + 4 * F);       // res blocks
    n += 2 * F + 8 + 65 * 128 + 65;                                // policy head
    n += F + 4 + 256 * 64 + 256 + 256 + 1;                         // value head
    return n;
}

// fold eval-mode BN into (scale per out channel, bias per out channel); same fp32 operation
// order as oracle/net_oracle.py forward_bf16_emulated.fold
static void bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, int c,
                    std::vector<float>& scale, std::vector<float>& bias)
{
    scale.resize(c); bias.resize(c);
    for (int i = 0; i < c; ++i) {
        const float s = gamma[i] / sqrtf(var[i] + 1e-5f);
        scale[i] = s;
        bias[i] = beta[i] - mean[i] * s;
    }
}

int NetHost::load(const float* flat, int64_t count)
{
    OTH_REQUIRE(count == net_param_count(blocks, F), OTH_ERR_ARG,
                "oth_net_load_weights: got %lld values, a %dx%d network needs %lld", (long long)count, blocks, F,
                (long long)net_param_count(blocks, F));
    const int KC = F / 8, n_conv = 1 + 2 * blocks;
    const float* p = flat;
    std::vector<uint16_t> wtc(w_tc_elems()), wtc2(w_tc_elems());
    std::vector<float> wsimt(w_simt_elems());
    std::vector<float> bias((size_t)n_conv * F);
    std::vector<float> scale, b;
    size_t tc_off = 0, simt_off = 0;
    for (int conv = 0; conv < n_conv; ++conv) {
        const int cin = conv == 0 ? 3 : F;
        const int cin_pad = conv == 0 ? 16 : F;        // tensor-core K padding for the stem
        const int cin_simt = conv == 0 ? 8 : F;
        const float* w = p; p += (int64_t)F * cin * 9;  // [cout][cin][3][3]
        const float *g = p, *be = p + F, *mu = p + 2 * F, *var = p + 3 * F; p += 4 * F;
        bn_fold(g, be, mu, var, F, scale, b);
        for (int c = 0; c < F; ++c) bias[(size_t)conv * F + c] = b[c];
        const int kcp = cin_pad / 8;
        for (int tap = 0; tap < 9; ++tap)
            for (int ci = 0; ci < cin_pad; ++ci)
                for (int co = 0; co < F; ++co) {
                    float v = 0.f;
                    if (ci < cin) v = w[((int64_t)co * cin + ci) * 9 + tap] * scale[co];
                    const uint16_t h = f32_to_bf16_rne(v);
                    // UMMA B tiles, K-major, no swizzle.  Stem: [tap][kc][cout][8].  Trunk: the kernel walks K as
                    // (32-channel split, tap, plane) so a split can start early: [split][tap][plane in split][cout][8].
                    size_t idx;
                    if (conv == 0) {
                        idx = (((size_t)tap * kcp + ci / 8) * F + co) * 8 + (ci % 8);
                    } else {
                        const int pps = 4, plane = ci / 8, split = plane / pps, pl = plane % pps;
                        idx = ((((size_t)split * 9 + tap) * pps + pl) * F + co) * 8 + (ci % 8);
                    }
                    wtc[tc_off + idx] = h;
                    // CTA-pair engine (net_tc2.cu): every ring stage is stored as [CTA 0's block][CTA 1's block], CTA r
                    // holding cout rows [r*F/2, (r+1)*F/2).  Trunk stage = (split, tap): block [4 planes][F/2][8].
                    // Stem stages per tap row dy: taps dx 0,1 (block [2 taps][2 planes][F/2][8]) then tap dx 2.
                    {
                        const int nhalf = F / 2, nh = co / nhalf, cr = co % nhalf;
                        size_t idx2;
                        if (conv == 0) {
                            const int dy = tap / 3, dx = tap % 3, plane = ci / 8;
                            const size_t tap_elems = (size_t)2 * nhalf * 8;
                            const size_t stage_base = (size_t)dy * 6 * tap_elems + (dx < 2 ? 0 : 4 * tap_elems);
                            const size_t in_block = (dx < 2 ? (size_t)dx : 0) * tap_elems + ((size_t)plane * nhalf + cr) * 8 + (ci % 8);
                            idx2 = stage_base + (size_t)nh * (dx < 2 ? 2 : 1) * tap_elems + in_block;
                        } else {
                            const int pps = 4, plane = ci / 8, split = plane / pps, pl = plane % pps;
                            idx2 = (((((size_t)split * 9 + tap) * 2 + nh) * pps + pl) * nhalf + cr) * 8 + (ci % 8);
                        }
                        wtc2[tc_off + idx2] = h;
                    }
                    if (ci < cin_simt) wsimt[simt_off + ((size_t)tap * cin_simt + ci) * F + co] = bf16_to_f32(h);
                }
        tc_off += (size_t)9 * cin_pad * F;
        simt_off += (size_t)9 * cin_simt * F;
    }
    // policy head
    std::vector<float> ph_w(2 * F), ph_b(2), pfc_t(128 * 65), pfc_b(65), vh_w(F), vh_b(1), v1_t(64 * 256), v1_b(256), v2_w(256), v2_b(1);
    {
        const float* w = p; p += 2 * F;
        bn_fold(p, p + 2, p + 4, p + 6, 2, scale, b); p += 8;
        for (int o = 0; o < 2; ++o) { for (int c = 0; c < F; ++c) ph_w[o * F + c] = w[o * F + c] * scale[o]; ph_b[o] = b[o]; }
        const float* fw = p; p += 65 * 128;
        for (int j = 0; j < 65; ++j) for (int i = 0; i < 128; ++i) pfc_t[i * 65 + j] = fw[j * 128 + i];
        for (int j = 0; j < 65; ++j) pfc_b[j] = p[j];
        p += 65;
    }
    {
        const float* w = p; p += F;
        bn_fold(p, p + 1, p + 2, p + 3, 1, scale, b); p += 4;
        for (int c = 0; c < F; ++c) vh_w[c] = w[c] * scale[0];
        vh_b[0] = b[0];
        const float* f1 = p; p += 256 * 64;
        for (int k = 0; k < 256; ++k) for (int i = 0; i < 64; ++i) v1_t[i * 256 + k] = f1[k * 64 + i];
        for (int k = 0; k < 256; ++k) v1_b[k] = p[k];
        p += 256;
        for (int k = 0; k < 256; ++k) v2_w[k] = p[k];
        p += 256;
        v2_b[0] = p[0]; p += 1;
    }
    OTH_REQUIRE(p - flat == count, OTH_ERR_STATE, "internal: weight cursor mismatch");

    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream); };
    OTH_CHECK_CUDA(up(d_w_tc, wtc.data(), wtc.size() * 2));
    OTH_CHECK_CUDA(up(d_w_tc2, wtc2.data(), wtc2.size() * 2));
    OTH_CHECK_CUDA(up(d_w_simt, wsimt.data(), wsimt.size() * 4));
    float* f = d_small;
    auto put = [&](const std::vector<float>& v, const float** slot) {
        *slot = f;
        cudaError_t e = up(f, v.data(), v.size() * 4);
        f += (v.size() + 3) / 4 * 4;
        return e;
    };
    OTH_CHECK_CUDA(put(bias, &dev.bias));
    OTH_CHECK_CUDA(put(ph_w, &dev.ph_w)); OTH_CHECK_CUDA(put(ph_b, &dev.ph_b));
    OTH_CHECK_CUDA(put(pfc_t, &dev.pfc_t)); OTH_CHECK_CUDA(put(pfc_b, &dev.pfc_b));
    OTH_CHECK_CUDA(put(vh_w, &dev.vh_w)); OTH_CHECK_CUDA(put(vh_b, &dev.vh_b));
    OTH_CHECK_CUDA(put(v1_t, &dev.v1_t)); OTH_CHECK_CUDA(put(v1_b, &dev.v1_b));
    OTH_CHECK_CUDA(put(v2_w, &dev.v2_w)); OTH_CHECK_CUDA(put(v2_b, &dev.v2_b));
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
    dev.blocks = blocks; dev.F = F; dev.KC = KC;
    dev.w_tc = (const __nv_bfloat16*)d_w_tc; dev.w_tc2 = (const __nv_bfloat16*)d_w_tc2; dev.w_simt = d_w_simt;
    loaded = true;
    return OTH_OK;
}

int NetHost::allocate()
{
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    OTH_CHECK_CUDA(cudaMalloc(&d_w_tc, w_tc_elems() * 2));
    OTH_CHECK_CUDA(cudaMalloc(&d_w_tc2, w_tc_elems() * 2));
    OTH_CHECK_CUDA(cudaMalloc((void**)&d_w_simt, w_simt_elems() * 4));
    const size_t small = (size_t)(1 + 2 * blocks) * F + 2 * F + 2 + 128 * 65 + 65 + F + 1 + 64 * 256 + 256 + 256 + 1 + 64;
    OTH_CHECK_CUDA(cudaMalloc((void**)&d_small, small * 4));
    return make_weight_tensor_map();
}

// 2-D tensor maps over d_w_tc: rows of 256 bytes (128 bf16); one box = one weight request of the kernel that uses the map
// (k_net_tc: a ring stage group; k_net_lat: four / three stages of a conv).
static bool encode_weight_map(CUtensorMap* map, void* base, size_t bytes, int box_rows)
{
    // the driver entry point is looked up at run time: the library must load on a box without libcuda (build check)
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t gdim[2] = {128, (cuuint64_t)(bytes / 256)};
    const cuuint64_t gstride[1] = {256};
    const cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return ((EncodeFn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int NetHost::make_weight_tensor_map()
{
    tmap_ok = tmap_lat_ok = false;
    if (!net_tc_supported(F)) return OTH_OK;
    const size_t bytes = w_tc_elems() * 2;
    if (bytes % 256) return OTH_OK;
    tmap_ok = encode_weight_map(&tmap_w, d_w_tc, bytes, net_tc_stage_rows(F));          // not fatal: the kernels fall back to
    tmap_lat_ok = encode_weight_map(&tmap_lat, d_w_tc, bytes, net_tc_lat_box_rows(F));  // 1-D bulk copies / the throughput shape
    return OTH_OK;
}

void NetHost::release()
{
    cudaSetDevice(ctx->device);
    cudaFree(d_w_tc); cudaFree(d_w_tc2); cudaFree(d_w_simt); cudaFree(d_small);
    d_w_tc2 = nullptr; d_w_tc = nullptr; d_w_simt = nullptr; d_small = nullptr;
}

// ---- CUDA-core validation engine -----------------------------------------------------------
// One CTA = one tile (2 boards), 256 threads: thread (m = t & 127, half = t >> 7) owns GEMM row
// m and output channels [half*F/2, (half+1)*F/2).

constexpr int kSimtThreads = 256;

template <int CPT>   // channels per thread = F/2
__device__ __forceinline__ void simt_conv(const NetDev& net, const uint4* __restrict__ in, uint4* __restrict__ out,
                                          const uint4* __restrict__ resid, const float* __restrict__ w,
                                          const float* __restrict__ bias, int cin_units, int m, int half)
{
    const int F = net.F;
    float acc[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) acc[i] = 0.f;
    const int row_unit = kGuardUnits + kHaloUnits + (m >> 3) * kGroupUnits + (m & 7);
    for (int tap = 0; tap < 9; ++tap) {
        const int shift = (tap / 3 - 1) * 2 * kGroupUnits + (tap % 3 - 1);
        for (int kc = 0; kc < cin_units; ++kc) {
            const uint4 u = in[kc * kPlaneUnits + row_unit + shift];
            const float x[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
            const float* wrow = w + ((size_t)tap * cin_units * 8 + kc * 8) * F + half * CPT;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4* w4 = reinterpret_cast<const float4*>(wrow + (size_t)j * F);
#pragma unroll
                for (int c = 0; c < CPT / 4; ++c) {
                    const float4 ww = __ldg(w4 + c);
                    acc[4 * c + 0] = fmaf(x[j], ww.x, acc[4 * c + 0]);
                    acc[4 * c + 1] = fmaf(x[j], ww.y, acc[4 * c + 1]);
                    acc[4 * c + 2] = fmaf(x[j], ww.z, acc[4 * c + 2]);
                    acc[4 * c + 3] = fmaf(x[j], ww.w, acc[4 * c + 3]);
                }
            }
        }
    }
    // epilogue: + bias (+ residual), relu, round to bf16, store 8 channels per unit
#pragma unroll
    for (int c8 = 0; c8 < CPT / 8; ++c8) {
        const int kc = (half * CPT) / 8 + c8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = acc[c8 * 8 + j] + __ldg(bias + kc * 8 + j);
        if (resid) {
            const uint4 r = resid[unit_of_row(kc, m)];
            v[0] += bf16_lo(r.x); v[1] += bf16_hi(r.x); v[2] += bf16_lo(r.y); v[3] += bf16_hi(r.y);
            v[4] += bf16_lo(r.z); v[5] += bf16_hi(r.z); v[6] += bf16_lo(r.w); v[7] += bf16_hi(r.w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        out[unit_of_row(kc, m)] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                             pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

template <int CPT>
__global__ void __launch_bounds__(kSimtThreads) k_net_simt(NetDev net, const uint64_t* __restrict__ self_b,
                                                           const uint64_t* __restrict__ opp_b, int64_t n,
                                                           float* __restrict__ policy_out, float* __restrict__ value_out,
                                                           int out_kind, const int32_t* __restrict__ n_dev)
{
    if (n_dev) { const int64_t nd = *n_dev; n = nd < n ? nd : n; }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int KC = net.KC, F = net.F;
    const int bytes = tile_buffer_bytes(KC);
    uint4* bufA = reinterpret_cast<uint4*>(smem_raw);
    uint4* bufB = reinterpret_cast<uint4*>(smem_raw + bytes);
    __shared__ uint64_t s_self[2], s_opp[2], s_legal[2];
    const int t = threadIdx.x, m = t & 127, half = t >> 7;
    const int64_t n_tiles = (n + 1) / 2;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t board0 = tile * 2;
        __syncthreads();
        zero_tile_buffer(bufA, KC, t, kSimtThreads);
        zero_tile_buffer(bufB, KC, t, kSimtThreads);
        if (t < 2) {
            const int64_t b = board0 + t;
            const uint64_t a = b < n ? self_b[b] : 0ULL, o = b < n ? opp_b[b] : 0ULL;
            s_self[t] = a; s_opp[t] = o; s_legal[t] = legal_moves(a, o);
        }
        __syncthreads();
        if (t < kTileRows) build_input_row(bufA, m, s_self, s_opp, s_legal);
        __syncthreads();
        const float* w = net.w_simt;
        // stem: A(input) -> B
        simt_conv<CPT>(net, bufA, bufB, nullptr, w, net.bias, 1, m, half);
        w += (size_t)9 * 8 * F;
        __syncthreads();
        for (int blk = 0; blk < net.blocks; ++blk) {
            simt_conv<CPT>(net, bufB, bufA, nullptr, w, net.bias + (size_t)(1 + 2 * blk) * F, KC, m, half);   // conv1: x -> h
            w += (size_t)9 * F * F;
            __syncthreads();
            simt_conv<CPT>(net, bufA, bufB, bufB, w, net.bias + (size_t)(2 + 2 * blk) * F, KC, m, half);      // conv2 + skip: in place on x
            w += (size_t)9 * F * F;
            __syncthreads();
        }
        heads_for_tile(net, bufB, reinterpret_cast<HeadScratch*>(bufA), s_legal, board0, n, policy_out, value_out, out_kind,
                       t, kSimtThreads, [] { __syncthreads(); });
    }
}

int net_forward_simt(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                     int out_kind, const int32_t* n_dev)
{
    oth_ctx* ctx = net->ctx;
    const int F = net->F;
    OTH_REQUIRE(F % 16 == 0 && F >= 16 && F <= 128, OTH_ERR_UNSUPPORTED, "SIMT engine: num_filters must be a multiple of 16 in 16..128 (got %d)", F);
    const int smem = 2 * tile_buffer_bytes(F / 8);
    const int64_t tiles = (n + 1) / 2;
    int grid = (int)(tiles < (int64_t)ctx->sm_count * 2 ? tiles : (int64_t)ctx->sm_count * 2);
    if (grid < 1) grid = 1;
#define OTH_SIMT_CASE(cpt)                                                                                          \
    case cpt: {                                                                                                     \
        OTH_CHECK_CUDA(cudaFuncSetAttribute(k_net_simt<cpt>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   \
        k_net_simt<cpt><<<grid, kSimtThreads, smem, ctx->stream>>>(net->dev, self_b, opp_b, n, policy, value, out_kind, n_dev); \
    } break;
    switch (F / 2) {
        OTH_SIMT_CASE(8) OTH_SIMT_CASE(16) OTH_SIMT_CASE(24) OTH_SIMT_CASE(32) OTH_SIMT_CASE(40) OTH_SIMT_CASE(48)
        OTH_SIMT_CASE(56) OTH_SIMT_CASE(64)
        default: set_error("SIMT engine: unsupported filter count %d", F); return OTH_ERR_UNSUPPORTED;
    }
#undef OTH_SIMT_CASE
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    return OTH_OK;
}

}  // namespace oth

using namespace oth;

extern "C" {

int oth_net_create(oth_ctx* ctx, int num_blocks, int num_filters, oth_net** out)
{
    OTH_REQUIRE(ctx && out, OTH_ERR_ARG, "oth_net_create: NULL argument");
    OTH_REQUIRE(num_blocks >= 0 && num_blocks <= 64, OTH_ERR_ARG, "oth_net_create: num_blocks %d out of range", num_blocks);
    OTH_REQUIRE(num_filters % 16 == 0 && num_filters >= 16 && num_filters <= 128, OTH_ERR_UNSUPPORTED,
                "oth_net_create: num_filters must be a multiple of 16 in 16..128 (got %d)", num_filters);
    oth_net* n = new oth_net();
    n->ctx = ctx; n->blocks = num_blocks; n->F = num_filters;
    n->engine = net_tc_supported(num_filters) ? OTH_NET_ENGINE_TCGEN05 : OTH_NET_ENGINE_SIMT;
    int rc = n->allocate();
    if (rc) { n->release(); delete n; return rc; }
    *out = n;
    return OTH_OK;
}

int oth_net_destroy(oth_net* net)
{
    if (!net) return OTH_OK;
    cudaStreamSynchronize(net->ctx->stream);
    net->release();
    delete net;
    return OTH_OK;
}

int64_t oth_net_param_count(const oth_net* net) { return net ? net_param_count(net->blocks, net->F) : -1; }

int oth_net_load_weights(oth_net* net, const float* flat, int64_t count)
{
    OTH_REQUIRE(net && flat, OTH_ERR_ARG, "oth_net_load_weights: NULL argument");
    return net->load(flat, count);
}

int oth_net_engine(const oth_net* net) { return net ? net->engine : OTH_ERR_ARG; }

int oth_net_set_engine(oth_net* net, int engine)
{
    OTH_REQUIRE(net, OTH_ERR_ARG, "oth_net_set_engine: net is NULL");
    OTH_REQUIRE(engine == OTH_NET_ENGINE_TCGEN05 || engine == OTH_NET_ENGINE_SIMT || engine == OTH_NET_ENGINE_TCGEN05_PAIR, OTH_ERR_ARG,
                "unknown engine %d", engine);
    OTH_REQUIRE(engine == OTH_NET_ENGINE_SIMT || net_tc_supported(net->F), OTH_ERR_UNSUPPORTED,
                "tcgen05 engine supports num_filters 64 or 128 (got %d)", net->F);
    net->engine = engine;
    return OTH_OK;
}

int oth_debug_net_trace(oth_net* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, uint64_t* trace_out, int trace_len)
{
    OTH_REQUIRE(net && self_b && opp_b && trace_out && trace_len > 0 && n > 0, OTH_ERR_ARG, "oth_debug_net_trace: bad argument");
    OTH_REQUIRE(net->loaded, OTH_ERR_STATE, "oth_debug_net_trace: weights not loaded");
    oth_ctx* ctx = net->ctx;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    unsigned long long* d_trace = nullptr; uint64_t *ds = nullptr, *dop = nullptr; float *dp = nullptr, *dv = nullptr;
    OTH_CHECK_CUDA(cudaMalloc((void**)&d_trace, (size_t)trace_len * 8));
    OTH_CHECK_CUDA(cudaMalloc((void**)&ds, n * 8)); OTH_CHECK_CUDA(cudaMalloc((void**)&dop, n * 8));
    OTH_CHECK_CUDA(cudaMalloc((void**)&dp, n * 65 * 4)); OTH_CHECK_CUDA(cudaMalloc((void**)&dv, n * 4));
    OTH_CHECK_CUDA(cudaMemsetAsync(d_trace, 0, (size_t)trace_len * 8, ctx->stream));
    OTH_CHECK_CUDA(cudaMemcpyAsync(ds, self_b, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    OTH_CHECK_CUDA(cudaMemcpyAsync(dop, opp_b, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = OTH_OK;
    for (int rep = 0; rep < 3 && rc == OTH_OK; ++rep) {          // warm (L2-resident weights), keep the last trace
        net->dev.trace = rep == 2 ? d_trace : nullptr;
        rc = net_forward_device(net, ds, dop, n, dp, dv, kOutPriors);
    }
    net->dev.trace = nullptr;
    if (rc == OTH_OK) {
        cudaError_t e = cudaMemcpyAsync(trace_out, d_trace, (size_t)trace_len * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("oth_debug_net_trace: %s", cudaGetErrorString(e)); rc = OTH_ERR_CUDA; }
    }
    cudaFree(d_trace); cudaFree(ds); cudaFree(dop); cudaFree(dp); cudaFree(dv);
    return rc;
}

int oth_net_forward(oth_net* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy_out,
                    float* value_out, int out_kind, int mem)
{
    OTH_REQUIRE(net && (n == 0 || (self_b && opp_b && policy_out && value_out)), OTH_ERR_ARG, "oth_net_forward: NULL argument");
    OTH_REQUIRE(n >= 0, OTH_ERR_ARG, "oth_net_forward: n < 0");
    OTH_REQUIRE(out_kind >= 0 && out_kind <= 2, OTH_ERR_ARG, "oth_net_forward: bad out_kind %d", out_kind);
    OTH_REQUIRE(net->loaded, OTH_ERR_STATE, "oth_net_forward: weights not loaded");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(net->ctx->device));
    Staged st(net->ctx, mem);
    const uint64_t* a = st.in(self_b, n); const uint64_t* b = st.in(opp_b, n);
    float* p = st.out(policy_out, n * 65); float* v = st.out(value_out, n);
    if (st.failed) return OTH_ERR_CUDA;
    int rc = net_forward_device(net, a, b, n, p, v, out_kind);
    if (rc) return rc;
    return st.finish();
}

}  // extern "C"

namespace oth {
int net_forward_device(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                       int out_kind, const int32_t* n_dev)
{
    TimedLaunch timed(net->ctx, 0);
    if (net->engine == OTH_NET_ENGINE_TCGEN05) {
        // A small batch costs the network's LATENCY and runs on the latency shape (net_tc_lat.cu: one tile per CTA, tensor-map
        // TMA weight stream, ~80 us instead of ~137 us for 10x128; identical bits), everything else on the throughput
        // kernel.  When the batch size is only known on the device (n_dev: the compacted leaf batch of a search step) and the
        // host's upper bound n exceeds the latency shape's range, BOTH kernels are launched and each one returns at once if
        // the device count is not in its range -- an empty launch costs ~3 us, a device->host read-back would cost more.
        const int64_t lat_max = net_tc_lat_max_positions(net);
        if (lat_max > 0 && n > 0) {
            if (n <= lat_max) return net_forward_tc_lat(net, self_b, opp_b, n, policy, value, out_kind, n_dev);
            if (n_dev) {
                int rc = net_forward_tc_lat(net, self_b, opp_b, n, policy, value, out_kind, n_dev);
                if (rc) return rc;
                return net_forward_tc(net, self_b, opp_b, n, policy, value, out_kind, n_dev, lat_max + 1);
            }
        }
        return net_forward_tc(net, self_b, opp_b, n, policy, value, out_kind, n_dev);
    }
    if (net->engine == OTH_NET_ENGINE_TCGEN05_PAIR) return net_forward_tc2(net, self_b, opp_b, n, policy, value, out_kind, n_dev);
    return net_forward_simt(net, self_b, opp_b, n, policy, value, out_kind, n_dev);
}
}  // namespace oth
