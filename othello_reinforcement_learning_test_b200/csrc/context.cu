// context.cu -- library context, error text, host staging.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace oth {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void* Staged::stage(void* p, size_t bytes, bool copy_in, bool copy_out)
{
    if (mem != OTH_MEM_HOST || p == nullptr || failed) return failed ? nullptr : p;
    void* d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, bytes ? bytes : 1, ctx->stream);
    if (e != cudaSuccess) { set_error("staging alloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); failed = true; return nullptr; }
    owned.push_back(d);
    if (copy_in) {
        e = cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { set_error("H2D copy failed: %s", cudaGetErrorString(e)); failed = true; }
    }
    if (copy_out) outs.push_back({p, d, bytes});
    return d;
}

Staged::~Staged()
{
    for (void* d : owned) cudaFreeAsync(d, ctx->stream);
}

int Staged::finish()
{
    if (failed) return OTH_ERR_CUDA;
    if (mem != OTH_MEM_HOST) return OTH_OK;
    for (auto& o : outs) OTH_CHECK_CUDA(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return OTH_OK;
}

}  // namespace oth

extern "C" {

const char* oth_last_error(void) { return oth::g_err; }

const char* oth_version(void) { return "othello_b200 0.1 (sm_100a)"; }

int oth_device_count(int* count)
{
    OTH_REQUIRE(count, OTH_ERR_ARG, "count is NULL");
    *count = 0;
    OTH_CHECK_CUDA(cudaGetDeviceCount(count));
    return OTH_OK;
}

int oth_ctx_create(int device, oth_ctx** out)
{
    OTH_REQUIRE(out, OTH_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    OTH_CHECK_CUDA(cudaGetDeviceCount(&n));
    OTH_REQUIRE(n > 0, OTH_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    OTH_REQUIRE(device >= 0 && device < n, OTH_ERR_ARG, "device %d out of range (have %d)", device, n);
    OTH_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OTH_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    OTH_REQUIRE(prop.major == 10, OTH_ERR_UNSUPPORTED,
                "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    oth_ctx* c = new oth_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; oth::set_error("stream create: %s", cudaGetErrorString(e)); return OTH_ERR_CUDA; }
    e = cudaHostAlloc((void**)&c->mailbox, sizeof(oth_board_state), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&c->mailbox_dev, c->mailbox, 0);
    if (e != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; oth::set_error("mailbox alloc: %s", cudaGetErrorString(e)); return OTH_ERR_CUDA; }
    memset(c->mailbox, 0, sizeof(oth_board_state));
    *out = c;
    return OTH_OK;
}

int oth_ctx_destroy(oth_ctx* ctx)
{
    if (!ctx) return OTH_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int c = 0; c < oth::kTimerCats; ++c)
        for (cudaEvent_t e : ctx->timer.pool[c]) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
    delete ctx;
    return OTH_OK;
}

int oth_ctx_sync(oth_ctx* ctx)
{
    OTH_REQUIRE(ctx, OTH_ERR_ARG, "ctx is NULL");
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return OTH_OK;
}

uint64_t oth_ctx_stream(oth_ctx* ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }

uint64_t oth_ctx_launch_count(oth_ctx* ctx) { return ctx ? ctx->launches : 0; }

int oth_ctx_timing_enable(oth_ctx* ctx, int on)
{
    OTH_REQUIRE(ctx, OTH_ERR_ARG, "ctx is NULL");
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->timer.on = on != 0;
    for (int c = 0; c < oth::kTimerCats; ++c) ctx->timer.used[c] = 0;
    return OTH_OK;
}

int oth_ctx_timing_read(oth_ctx* ctx, double* ms_out, uint64_t* count_out)
{
    OTH_REQUIRE(ctx && ms_out && count_out, OTH_ERR_ARG, "oth_ctx_timing_read: NULL argument");
    OTH_CHECK_CUDA(cudaSetDevice(ctx->device));
    OTH_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < oth::kTimerCats; ++c) {
        double total = 0.0;
        for (size_t i = 0; i + 1 < ctx->timer.used[c]; i += 2) {
            float ms = 0.f;
            OTH_CHECK_CUDA(cudaEventElapsedTime(&ms, ctx->timer.pool[c][i], ctx->timer.pool[c][i + 1]));
            total += ms;
        }
        ms_out[c] = total;
        count_out[c] = ctx->timer.used[c] / 2;
        ctx->timer.used[c] = 0;
    }
    return OTH_OK;
}

}  // extern "C"
