// net_host.cuh -- host-side network object shared by net.cu, net_tc.cu, search.cu, selfplay.cu
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "net_common.cuh"

namespace oth {

struct NetHost {
    oth_ctx* ctx = nullptr;
    int blocks = 0, F = 0;
    int engine = 0;
    bool loaded = false;
    void* d_w_tc = nullptr;
    void* d_w_tc2 = nullptr;
    float* d_w_simt = nullptr;
    float* d_small = nullptr;
    NetDev dev{};
    // tensor map over the tcgen05 weight buffer: 2-D view [rows][256 B], box = one ring stage (group), no swizzle --
    // lets the trunk's weight stream go through cp.async.bulk.tensor (UTMALDG) instead of 1-D bulk copies
    alignas(64) CUtensorMap tmap_w;       // box = one ring stage group of k_net_tc (experiment knob OTH_TC_TMAP)
    alignas(64) CUtensorMap tmap_lat;     // box = one request of the latency shape k_net_lat (32 KB / 12 KB)
    bool tmap_ok = false, tmap_lat_ok = false;
    uint64_t evals = 0;     // positions evaluated so far (bench bookkeeping)

    size_t w_tc_elems() const { return (size_t)9 * 16 * F + (size_t)2 * blocks * 9 * F * F; }
    size_t w_simt_elems() const { return (size_t)9 * 8 * F + (size_t)2 * blocks * 9 * F * F; }
    int allocate();
    int make_weight_tensor_map();
    void release();
    int load(const float* flat, int64_t count);
};

int64_t net_param_count(int blocks, int F);
bool net_tc_supported(int F);
int net_tc_stage_rows(int F);      // 256-byte rows of one weight ring stage group of the tcgen05 trunk
// all pointers are DEVICE pointers; asynchronous on net->ctx->stream
// `n_dev` (optional, device): actual batch size decided on the GPU (<= n); the kernels read it themselves,
// so a compacted leaf batch needs no host round trip.
int net_forward_device(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                       int out_kind, const int32_t* n_dev = nullptr);
int net_forward_simt(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                     int out_kind, const int32_t* n_dev);
// n_min: the kernel leaves batches smaller than this (device count) to the latency shape launched before it
int net_forward_tc(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                   int out_kind, const int32_t* n_dev, int64_t n_min = 0);
// the latency shape (net_tc_lat.cu): one tile per CTA, tensor-map TMA weight stream; for batches of at most 2 x SM count
int64_t net_tc_lat_max_positions(const NetHost* net);
int net_tc_lat_box_rows(int F);
int net_forward_tc_lat(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                       int out_kind, const int32_t* n_dev);
int net_forward_tc2(NetHost* net, const uint64_t* self_b, const uint64_t* opp_b, int64_t n, float* policy, float* value,
                    int out_kind, const int32_t* n_dev);

}  // namespace oth

struct oth_net : public oth::NetHost {};
