// context.cu -- process-wide context, options, timing and error text of libdrops_b200.
#include "common.cuh"

#include <atomic>
#include <cstring>

#include <cstdarg>
#include <cstring>
#include <mutex>

namespace db200 {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();  // clear the sticky-less error state
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? DROPS_B200_E_NODEVICE
                                                                         : DROPS_B200_E_CUDA;
}

Ctx& ctx()
{
    static Ctx c;
    return c;
}

int require_ctx()
{
    if (!ctx().inited) {
        set_error("drops_b200_init() has not been called (or failed): there is no CPU fallback");
        return DROPS_B200_E_NOTINIT;
    }
    return 0;
}

double opt(const char* key, double dflt)
{
    auto& o = ctx().options;
    auto it = o.find(key);
    return it == o.end() ? dflt : it->second;
}

int vec_grid(size_t n, int threads, int per_thread)
{
    const Ctx& c = ctx();
    const size_t per_block = (size_t)threads * per_thread;
    size_t blocks = (n + per_block - 1) / per_block;
    const size_t cap = (size_t)c.sm_count * 8;   // 8 x 256 threads = full occupancy
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (blocks > (size_t)c.max_grid) blocks = c.max_grid;
    return (int)blocks;
}

static cudaEvent_t get_event()
{
    Ctx& c = ctx();
    if (!c.event_pool.empty()) {
        cudaEvent_t e = c.event_pool.back();
        c.event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ScopedTiming::ScopedTiming(const char* key, cudaStream_t st)
{
    Ctx& c = ctx();
    if (opt("timing", 0.0) == 0.0) return;
    slot = &c.timing[key];
    stream = st ? st : c.stream;
    e0 = get_event();
    e1 = get_event();
    cudaEventRecord(e0, stream);
}

ScopedTiming::~ScopedTiming()
{
    if (!slot) return;
    cudaEventRecord(e1, stream);
    slot->pending.emplace_back(e0, e1);
    slot->launches += 1;
}

// payload first, system-wide fence, then the sequence word: the host reads the payload only after it saw the word
__global__ void k_publish(const double* __restrict__ src, int count, double* __restrict__ pub, unsigned seq)
{
    for (int i = threadIdx.x; i < count; i += blockDim.x) pub[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile unsigned*>(pub + kPubDoubles) = seq;
        __threadfence_system();
    }
}

int fetch_small(const double* dsrc, double* hdst, size_t count)
{
    Ctx& c = ctx();
    if (c.h_pub == nullptr || count > (size_t)kPubDoubles || opt("fetch.poll", 1.0) == 0.0) {
        DB_CUDA(cudaMemcpyAsync(hdst, dsrc, count * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        DB_CUDA(cudaStreamSynchronize(c.stream));
        return 0;
    }
    const unsigned seq = ++c.pub_seq;
    k_publish<<<1, count > 32 ? 128 : 32, 0, c.stream>>>(dsrc, (int)count, c.d_pub, seq);
    count_launch();
    DB_CUDA(cudaGetLastError());
    volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(c.h_pub + kPubDoubles);
    unsigned long long spins = 0;
    while (*flag != seq) {
        if ((++spins & 0x3fffu) == 0) {                       // every 16 k polls: is the stream still alive?
            const cudaError_t q = cudaStreamQuery(c.stream);
            if (q == cudaSuccess) break;                      // drained: the word is there (or the kernel never ran)
            if (q != cudaErrorNotReady) return cuda_fail(q, "fetch_small", __FILE__, __LINE__);
        }
    }
    if (*flag != seq) {                                       // stream drained without the word: report whatever went wrong
        DB_CUDA(cudaStreamSynchronize(c.stream));
        if (*flag != seq) { set_error("fetch_small: the publishing kernel did not run"); return DROPS_B200_E_CUDA; }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    std::memcpy(hdst, c.h_pub, count * sizeof(double));
    return 0;
}

int timing_flush()
{
    Ctx& c = ctx();
    for (auto& kv : c.timing) {
        for (auto& pr : kv.second.pending) {
            float ms = 0.f;
            cudaEventSynchronize(pr.second);
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) kv.second.ms += ms;
            c.event_pool.push_back(pr.first);
            c.event_pool.push_back(pr.second);
        }
        kv.second.pending.clear();
    }
    return 0;
}

}  // namespace db200

using namespace db200;

extern "C" {

const char* drops_b200_last_error(void) { return g_err; }

int drops_b200_abi_version(void) { return 1; }

int drops_b200_compiled_arch(void) { return 100; }

int drops_b200_init(int device)
{
    Ctx& c = ctx();
    if (c.inited) {
        if (c.device == device) return 0;
        set_error("drops_b200_init: already bound to device %d (one process drives one GPU)", c.device);
        return DROPS_B200_E_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible (%s): libdrops_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return DROPS_B200_E_NODEVICE;
    }
    DB_ARG(device >= 0 && device < ndev, "device index out of range");
    DB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    DB_CUDA(cudaGetDeviceProperties(&p, device));
    c.cc = p.major * 10 + p.minor;
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; the kernels in this library are compiled for sm_100a only",
                  device, p.major, p.minor);
        return DROPS_B200_E_NODEVICE;
    }
    c.device = device;
    c.sm_count = p.multiProcessorCount;
    DB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    DB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    c.max_grid = c.sm_count * 8;
    DB_CHECK(dev_alloc(&c.partials, (size_t)kPartialStride * kPartialSlots));
    DB_CHECK(dev_alloc(&c.tickets, (size_t)kPartialSlots * kTicketStride));
    DB_CUDA(cudaMemset(c.tickets, 0, (size_t)kPartialSlots * kTicketStride * sizeof(unsigned)));
    DB_CHECK(dev_alloc(&c.scratch_scalars, 64));
    DB_CUDA(cudaMallocHost((void**)&c.h_scalars, 64 * sizeof(double)));
    c.h_pub = nullptr; c.d_pub = nullptr; c.pub_seq = 0;
    if (cudaHostAlloc((void**)&c.h_pub, (kPubDoubles + 2) * sizeof(double), cudaHostAllocMapped) == cudaSuccess) {
        std::memset(c.h_pub, 0, (kPubDoubles + 2) * sizeof(double));
        if (cudaHostGetDevicePointer((void**)&c.d_pub, c.h_pub, 0) != cudaSuccess) { cudaFreeHost(c.h_pub); c.h_pub = nullptr; c.d_pub = nullptr; }
    } else {
        cudaGetLastError();
        c.h_pub = nullptr;
    }
    DB_CUDA(cudaDeviceSynchronize());   // the memsets above ran on the legacy stream
    c.inited = true;
    return 0;
}

int drops_b200_shutdown(void)
{
    Ctx& c = ctx();
    if (!c.inited) return 0;
    cudaStreamSynchronize(c.stream);
    timing_flush();
    for (cudaEvent_t e : c.event_pool) cudaEventDestroy(e);
    c.event_pool.clear();
    dev_free(c.partials); c.partials = nullptr;
    dev_free(c.tickets); c.tickets = nullptr;
    dev_free(c.scratch_scalars); c.scratch_scalars = nullptr;
    if (c.h_scalars) cudaFreeHost(c.h_scalars);
    c.h_scalars = nullptr;
    if (c.h_pub) cudaFreeHost(c.h_pub);
    c.h_pub = nullptr; c.d_pub = nullptr;
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.copy_stream);
    c.stream = c.copy_stream = nullptr;
    c.inited = false;
    c.device = -1;
    return 0;
}

int drops_b200_device_info(int* sm_count, int* cc, size_t* free_bytes, size_t* total_bytes)
{
    DB_CHECK(require_ctx());
    size_t f = 0, t = 0;
    DB_CUDA(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = ctx().sm_count;
    if (cc) *cc = ctx().cc;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return 0;
}

int drops_b200_set_option(const char* key, double value)
{
    DB_ARG(key != nullptr, "null key");
    ctx().options[key] = value;
    ctx().options_epoch += 1;
    return 0;
}

int drops_b200_get_option(const char* key, double* value)
{
    DB_ARG(key != nullptr && value != nullptr, "null argument");
    auto& o = ctx().options;
    auto it = o.find(key);
    if (it == o.end()) { set_error("option %s is not set", key); return DROPS_B200_E_ARG; }
    *value = it->second;
    return 0;
}

int drops_b200_timing_get(const char* key, double* ms, long long* launches)
{
    DB_CHECK(require_ctx());
    DB_ARG(key != nullptr, "null key");
    DB_CUDA(cudaStreamSynchronize(ctx().stream));
    timing_flush();
    auto it = ctx().timing.find(key);
    if (ms) *ms = it == ctx().timing.end() ? 0.0 : it->second.ms;
    if (launches) *launches = it == ctx().timing.end() ? 0 : it->second.launches;
    return 0;
}

int drops_b200_timing_reset(void)
{
    DB_CHECK(require_ctx());
    DB_CUDA(cudaStreamSynchronize(ctx().stream));
    timing_flush();
    ctx().timing.clear();
    return 0;
}

static cudaEvent_t g_timer[2] = {nullptr, nullptr};

int drops_b200_timer_start(void)
{
    DB_CHECK(require_ctx());
    for (int i = 0; i < 2; ++i)
        if (!g_timer[i]) DB_CUDA(cudaEventCreate(&g_timer[i]));
    DB_CUDA(cudaEventRecord(g_timer[0], ctx().stream));
    return 0;
}

int drops_b200_timer_stop(double* ms)
{
    DB_CHECK(require_ctx());
    DB_ARG(ms != nullptr && g_timer[0] != nullptr, "timer_stop without timer_start");
    DB_CUDA(cudaEventRecord(g_timer[1], ctx().stream));
    DB_CUDA(cudaEventSynchronize(g_timer[1]));
    float f = 0.f;
    DB_CUDA(cudaEventElapsedTime(&f, g_timer[0], g_timer[1]));
    *ms = (double)f;
    return 0;
}

long long drops_b200_launch_count(void) { return ctx().launches; }

int drops_b200_synchronize(void)
{
    DB_CHECK(require_ctx());
    DB_CUDA(cudaStreamSynchronize(ctx().stream));
    DB_CUDA(cudaStreamSynchronize(ctx().copy_stream));
    return 0;
}

}  // extern "C"
