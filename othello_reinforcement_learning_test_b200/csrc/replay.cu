// replay.cu -- device-resident replay buffer (the sink of the trajectory all-gather).
//
// Restates ReplayBuffer (src/train/buffer.py:15-123): a FIFO of at most `max_size` samples
// (deque(maxlen), :31,44-45), uniform sampling without replacement of a minibatch (:78) returned as
// states float32 [B,3,8,8], policies float32 [B,65], values float32 [B,1] (:80-84), statistics (:102-123).
// Samples stay packed (oth_sample, 168 B: three bit-planes + 65 visit counts + label) in a ring in HBM;
// a minibatch is gathered and expanded by one kernel straight into the tensors the trainer consumes
// (trainer.py:264-269), so neither the 1 KB/sample fp32 form nor a host hop is needed.
// Which indices are drawn is the caller's business (the Python class uses `random.sample`, as the
// reference does), so a seeded reference run and this buffer return identical minibatches.
#include "bitboard.cuh"
#include "common.cuh"

namespace oth {

struct ReplayHost {
    oth_ctx* ctx = nullptr;
    int64_t capacity = 0, size = 0, head = 0;     // logical index i lives at ring[(head + i) % capacity]
    oth_sample* ring = nullptr;
    int64_t* d_idx = nullptr; int64_t idx_cap = 0;
    double* d_stats = nullptr;
    int32_t* d_bad = nullptr;                     // set by the gather kernel when a device-side index was out of range
};

// Dihedral image t = 2*k + flip of the board, in the order of OthelloBitboard.get_symmetries (bitboard.pyx:338-370):
// np.rot90(planes, k, axes=(1, 2)) then np.flip(axis=2) when flip.  Returns the SOURCE square whose content lands on
// destination square d: rot90 once is out[i][j] = in[j][7 - i]; the flip is out[i][j] = in[i][7 - j].
__host__ __device__ __forceinline__ int sym_source_square(int d, int t)
{
    int i = d >> 3, j = d & 7;
    if (t & 1) j = 7 - j;
    for (int k = t >> 1; k > 0; --k) { const int ni = j, nj = 7 - i; i = ni; j = nj; }
    return (i << 3) | j;
}

// one warp per sample: 168 B in, 192 + 65 + 1 floats out
__global__ void __launch_bounds__(256)
k_replay_gather(const oth_sample* __restrict__ ring, int64_t capacity, int64_t head, int64_t size, const int64_t* __restrict__ idx,
                const uint8_t* __restrict__ sym, int64_t n, float* __restrict__ states, float* __restrict__ policies,
                float* __restrict__ values, int32_t* __restrict__ bad)
{
    const int64_t i = blockIdx.x * 8LL + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    int64_t li = idx[i];
    if (li < 0 || li >= size) {                                        // never read a stale ring entry: flag it, use entry 0
        if (lane == 0) atomicExch(bad, 1);
        li = 0;
    }
    const oth_sample* s = ring + (head + li) % capacity;
    const uint64_t planes[3] = {s->self_b, s->opp_b, s->legal};
    const int t = sym ? (sym[i] & 7) : 0;                              // dihedral image (0 = identity)
    const int src0 = sym_source_square(lane, t), src1 = sym_source_square(lane + 32, t);
    float* st = states + i * 192;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int e = k * 32 + lane;                                   // element of the [3][64] planes
        const int sq = (k & 1) ? src1 : src0;                          // the 8 images are bit permutations of the packed planes
        st[e] = (float)((planes[e >> 6] >> sq) & 1ULL);                // get_tensor_input layout (bitboard.pyx:300-323)
    }
    // policy = counts / counts.sum() in float32 (node.py:177-180); counts are small integers, the sum is exact
    int part = 0;
    for (int j = lane; j < OTH_ACTIONS; j += 32) part += s->visits[j];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    const float total = (float)part;
    float* po = policies + i * OTH_ACTIONS;
    po[lane] = part > 0 ? __fdiv_rn((float)s->visits[src0], total) : 0.f;
    po[lane + 32] = part > 0 ? __fdiv_rn((float)s->visits[src1], total) : 0.f;
    if (lane == 0) {
        po[64] = part > 0 ? __fdiv_rn((float)s->visits[64], total) : 0.f;     // the pass probability is carried (bitboard.pyx:365)
        values[i] = (float)s->value;
    }
}

__global__ void k_replay_stats(const oth_sample* __restrict__ ring, int64_t n, double* __restrict__ out)
{
    // sum and sum of squares of the labels (values are -1/0/+1: integer arithmetic is exact)
    long long s1 = 0, s2 = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = ring[i].value;
        s1 += v; s2 += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], (double)s1); atomicAdd(&out[1], (double)s2); }
}

}  // namespace oth

using namespace oth;

struct oth_replay : public oth::ReplayHost {};

extern "C" {

int oth_replay_create(oth_ctx* ctx, int64_t max_size, oth_replay** out)
{
    OTH_REQUIRE(ctx && out, OTH_ERR_ARG, "oth_replay_create: NULL argument");
    OTH_REQUIRE(max_size > 0 && max_size <= (1LL << 30), OTH_ERR_ARG, "oth_replay_create: max_size %lld out of range", (long long)max_size);
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    oth_replay* r = new oth_replay();
    r->ctx = ctx; r->capacity = max_size;
    cudaError_t e = cudaMalloc((void**)&r->ring, (size_t)max_size * sizeof(oth_sample));
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_stats, 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_bad, sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(r->d_bad, 0, sizeof(int32_t), ctx->stream);
    if (e != cudaSuccess) { set_error("oth_replay_create: %s", cudaGetErrorString(e)); cudaFree(r->ring); delete r; return OTH_ERR_CUDA; }
    *out = r;
    return OTH_OK;
}

int oth_replay_destroy(oth_replay* r)
{
    if (!r) return OTH_OK;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    cudaFree(r->ring); cudaFree(r->d_idx); cudaFree(r->d_stats); cudaFree(r->d_bad);
    delete r;
    return OTH_OK;
}

int64_t oth_replay_size(const oth_replay* r) { return r ? r->size : -1; }

int oth_replay_clear(oth_replay* r)
{
    OTH_REQUIRE(r, OTH_ERR_ARG, "oth_replay_clear: NULL handle");
    r->size = 0; r->head = 0;
    return OTH_OK;
}

// deque(maxlen).append for n samples (buffer.py:44-45): the oldest entries fall out when full
int oth_replay_add(oth_replay* r, const oth_sample* samples, int64_t n, int mem)
{
    OTH_REQUIRE(r && (samples || n == 0) && n >= 0, OTH_ERR_ARG, "oth_replay_add: bad argument");
    if (n == 0) return OTH_OK;
    OTH_CHECK_CUDA(cudaSetDevice(r->ctx->device));
    const cudaMemcpyKind kind = mem == OTH_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (n > r->capacity) { samples += n - r->capacity; n = r->capacity; }          // only the newest `capacity` survive
    int64_t tail = (r->head + r->size) % r->capacity;
    int64_t first = n < r->capacity - tail ? n : r->capacity - tail;
    OTH_CHECK_CUDA(cudaMemcpyAsync(r->ring + tail, samples, (size_t)first * sizeof(oth_sample), kind, r->ctx->stream));
    if (n > first)
        OTH_CHECK_CUDA(cudaMemcpyAsync(r->ring, samples + first, (size_t)(n - first) * sizeof(oth_sample), kind, r->ctx->stream));
    const int64_t overflow = r->size + n - r->capacity;
    if (overflow > 0) { r->head = (r->head + overflow) % r->capacity; r->size = r->capacity; }
    else r->size += n;
    if (mem == OTH_MEM_HOST) OTH_CHECK_CUDA(cudaStreamSynchronize(r->ctx->stream));
    return OTH_OK;
}

// gather + expand the samples at logical indices idx[0..n) (0 = oldest) into the trainer's tensors; sym (optional, one
// byte per sample, 0..7) picks a dihedral image of the sample in get_symmetries' order (bitboard.pyx:338-370)
static int replay_gather(oth_replay* r, const int64_t* idx, const uint8_t* sym, int64_t n, float* states, float* policies,
                         float* values, int mem)
{
    OTH_REQUIRE(r && idx && states && policies && values && n > 0, OTH_ERR_ARG, "oth_replay_gather: bad argument");
    oth_ctx* ctx = r->ctx;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    Staged st(ctx, mem);
    const int64_t* di = st.in(idx, n);
    const uint8_t* dsym = sym ? st.in(sym, n) : nullptr;
    float* ds = st.out(states, n * 192); float* dp = st.out(policies, n * OTH_ACTIONS); float* dv = st.out(values, n);
    if (st.failed) return OTH_ERR_CUDA;
    if (mem == OTH_MEM_HOST) for (int64_t i = 0; i < n; ++i) OTH_REQUIRE(idx[i] >= 0 && idx[i] < r->size, OTH_ERR_ARG, "oth_replay_gather: index %lld out of range (size %lld)", (long long)idx[i], (long long)r->size);
    k_replay_gather<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(r->ring, r->capacity, r->head, r->size, di, dsym, n, ds, dp, dv, r->d_bad);
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    return st.finish();
}

int oth_replay_gather(oth_replay* r, const int64_t* idx, int64_t n, float* states, float* policies, float* values, int mem)
{
    return replay_gather(r, idx, nullptr, n, states, policies, values, mem);
}

int oth_replay_gather_sym(oth_replay* r, const int64_t* idx, const uint8_t* sym, int64_t n, float* states, float* policies,
                          float* values, int mem)
{
    OTH_REQUIRE(sym, OTH_ERR_ARG, "oth_replay_gather_sym: sym is NULL");
    return replay_gather(r, idx, sym, n, states, policies, values, mem);
}

// Device-side index check of the gathers issued with OTH_MEM_DEVICE so far (their indices cannot be validated on the
// host): synchronises, returns OTH_ERR_ARG once if any index was outside [0, size) and clears the flag.
int oth_replay_check(oth_replay* r)
{
    OTH_REQUIRE(r, OTH_ERR_ARG, "oth_replay_check: NULL handle");
    OTH_CHECK_CUDA(cudaSetDevice(r->ctx->device));
    int32_t bad = 0;
    OTH_CHECK_CUDA(cudaMemcpyAsync(&bad, r->d_bad, sizeof bad, cudaMemcpyDeviceToHost, r->ctx->stream));
    OTH_CHECK_CUDA(cudaStreamSynchronize(r->ctx->stream));
    if (bad) {
        OTH_CHECK_CUDA(cudaMemsetAsync(r->d_bad, 0, sizeof bad, r->ctx->stream));
        set_error("oth_replay_gather: a device-side index was outside [0, %lld)", (long long)r->size);
        return OTH_ERR_ARG;
    }
    return OTH_OK;
}

// get_statistics (buffer.py:102-123): mean and (population) std of the value labels
int oth_replay_value_stats(oth_replay* r, double* mean_out, double* std_out)
{
    OTH_REQUIRE(r && mean_out && std_out, OTH_ERR_ARG, "oth_replay_value_stats: NULL argument");
    *mean_out = 0.0; *std_out = 0.0;
    if (r->size == 0) return OTH_OK;
    oth_ctx* ctx = r->ctx;
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    OTH_CHECK_CUDA(cudaMemsetAsync(r->d_stats, 0, 2 * sizeof(double), ctx->stream));
    // the ring is full or starts at 0 unless it wrapped; statistics do not depend on order, so scan the occupied part
    const int64_t n = r->size;   // occupied entries are [0,size) when not wrapped, the whole ring when full
    k_replay_stats<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(r->ring + (r->size == r->capacity ? 0 : r->head), n, r->d_stats);
    ctx->launches++;
    OTH_CHECK_CUDA(cudaGetLastError());
    double h[2];
    OTH_CHECK_CUDA(cudaMemcpyAsync(h, r->d_stats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    const double mean = h[0] / (double)n, var = h[1] / (double)n - mean * mean;
    *mean_out = mean; *std_out = var > 0 ? sqrt(var) : 0.0;
    return OTH_OK;
}

}  // extern "C"
