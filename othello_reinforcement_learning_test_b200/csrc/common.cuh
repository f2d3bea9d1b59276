// common.cuh -- context, error plumbing and host<->device staging shared by all
// translation units of libothello_b200.so.  No torch types anywhere: the library
// only sees plain pointers and sizes (include/othello_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/othello_b200.h"

namespace oth {

void set_error(const char* fmt, ...);

#define OTH_CHECK_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::oth::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return OTH_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define OTH_REQUIRE(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            ::oth::set_error(__VA_ARGS__);    \
            return (code);                    \
        }                                     \
    } while (0)

}  // namespace oth

namespace oth {
// Optional per-category kernel timing with CUDA events on the context stream (bench.py's roofline
// numerator).  Categories: 0 = network forward, 1 = tree kernels, 2 = self-play move kernels.
constexpr int kTimerCats = 3;
struct KernelTimer {
    bool on = false;
    std::vector<cudaEvent_t> pool[kTimerCats];   // pairs: [2i] start, [2i+1] stop
    size_t used[kTimerCats] = {0, 0, 0};
};
}  // namespace oth

struct oth_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;       // kernels launched through this context (bench: gpu_launches)
    oth::KernelTimer timer;
    // single-board mailbox (oth_board_step): page-locked, mapped into the device; the kernel writes the result there
    oth_board_state* mailbox = nullptr;       // host address
    oth_board_state* mailbox_dev = nullptr;   // device alias of the same memory
    uint64_t mailbox_seq = 0;
};

namespace oth {
// RAII: records a start event now and a stop event at scope exit when timing is enabled.
struct TimedLaunch {
    oth_ctx* ctx;
    cudaEvent_t stop = nullptr;
    TimedLaunch(oth_ctx* c, int cat) : ctx(c)
    {
        KernelTimer& t = c->timer;
        if (!t.on) return;
        if (t.used[cat] + 2 > t.pool[cat].size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            t.pool[cat].push_back(a); t.pool[cat].push_back(b);
        }
        cudaEventRecord(t.pool[cat][t.used[cat]], c->stream);
        stop = t.pool[cat][t.used[cat] + 1];
        t.used[cat] += 2;
    }
    ~TimedLaunch() { if (stop) cudaEventRecord(stop, ctx->stream); }
};
}  // namespace oth

namespace oth {

// Brings a caller buffer onto the device when mem == OTH_MEM_HOST (stream-ordered
// allocation, copy-in and/or copy-out on the context stream); passes device
// pointers through untouched.
struct Staged {
    oth_ctx* ctx;
    int mem;
    struct Out { void* host; void* dev; size_t bytes; };
    std::vector<Out> outs;
    std::vector<void*> owned;
    bool failed = false;
    Staged(oth_ctx* c, int m) : ctx(c), mem(m) {}
    ~Staged();
    template <class T> const T* in(const T* p, size_t count) { return (const T*)stage((void*)p, count * sizeof(T), true, false); }
    template <class T> T* out(T* p, size_t count) { return (T*)stage((void*)p, count * sizeof(T), false, true); }
    template <class T> T* inout(T* p, size_t count) { return (T*)stage((void*)p, count * sizeof(T), true, true); }
    void* stage(void* p, size_t bytes, bool copy_in, bool copy_out);
    int finish();   // D2H of staged outputs + stream sync for host calls
};

inline int grid_for(int64_t n, int block, int sm_count, int waves = 8)
{
    int64_t g = (n + block - 1) / block;
    int64_t cap = (int64_t)sm_count * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace oth
