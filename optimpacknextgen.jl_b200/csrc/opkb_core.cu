// opkb_core.cu -- Tier 0 (lifecycle) and Tier 1 (LazyAlgebra / SimpleBounds
// primitives, one kernel per reference call) of libopkb200.so.
//
// Reference semantics: doc/algebra.md:84-116, test/vectors-tests.jl:15-81,
// src/bounds.jl, src/quasinewton.jl:45-69.  sm_100a only; no CPU fallback.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <time.h>

#include "opkb_internal.cuh"

// ---------------------------------------------------------------------------
// errors
static thread_local char g_errbuf[512] = "";

int opkb_set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_errbuf, sizeof(g_errbuf), fmt, ap);
    va_end(ap);
    return code;
}

int opkb_check_cuda(cudaError_t err, const char* what)
{
    if (err == cudaSuccess) return 0;
    return opkb_set_error(OPKB_ECUDA, "CUDA error %d (%s) in %s", (int)err,
                          cudaGetErrorString(err), what);
}

extern "C" const char* opkb_last_error(void) { return g_errbuf; }
extern "C" int opkb_version(void) { return OPKB_VERSION; }

// ---------------------------------------------------------------------------
// context
static int context_init(opkb_context* ctx)
{
    OPKB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    OPKB_CUDA(cudaMalloc(&ctx->d_part, sizeof(double) * OPKB_MAX_GRID * (size_t)OPKB_NRED_MAX));
    OPKB_CUDA(cudaMalloc(&ctx->d_ticket, sizeof(unsigned int)));
    OPKB_CUDA(cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned int), ctx->stream));
    OPKB_CUDA(cudaMalloc(&ctx->d_res_dev, sizeof(double) * OPKB_NRED_MAX));
    OPKB_CUDA(cudaMallocHost(&ctx->h_res, sizeof(double) * OPKB_NRED_MAX));
    ctx->d_res = ctx->d_res_dev;
    {
        // one rank: the kernels write their results directly into the pinned host buffer
        const char* e = getenv("OPKB_MAPPED_RES");
        if (!e || atoi(e) != 0) ctx->d_res = ctx->h_res;
        // ... and a completion flag next to them, polled by the host (OPKB_SPIN=0: stream sync)
        OPKB_CUDA(cudaMallocHost(&ctx->h_flag, 64));
        *ctx->h_flag = 0u;
        e = getenv("OPKB_SPIN");
        ctx->spin = (!e || atoi(e) != 0) ? 1 : 0;
    }
    OPKB_CUDA(cudaEventCreate(&ctx->ev0));
    OPKB_CUDA(cudaEventCreate(&ctx->ev1));
    OPKB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int opkb_context_destroy(opkb_context* ctx);

extern "C" int opkb_context_create(opkb_context** out, int device)
{
    if (!out) return opkb_set_error(OPKB_EINVAL, "NULL context pointer");
    *out = NULL;
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev <= 0)
        return opkb_set_error(OPKB_ECUDA,
                              "no CUDA device available (%s): libopkb200 has no CPU fallback",
                              cudaGetErrorString(err));
    if (device < 0) OPKB_CUDA(cudaGetDevice(&device));
    if (device >= ndev) return opkb_set_error(OPKB_EINVAL, "device %d out of range", device);
    OPKB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OPKB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return opkb_set_error(OPKB_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a",
                              device, prop.major, prop.minor);
    opkb_context* ctx = (opkb_context*)calloc(1, sizeof(opkb_context));
    if (!ctx) return opkb_set_error(OPKB_ENOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->rank = 0;
    ctx->nranks = 1;
    int rc = context_init(ctx);
    if (rc != 0) {
        opkb_context_destroy(ctx);   // releases whatever was created (every member is NULL-safe)
        return rc;
    }
    *out = ctx;
    return 0;
}

extern "C" int opkb_context_destroy(opkb_context* ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    opkb_mailbox_close(ctx);
    opkb_context_trim(ctx);
    free(ctx->pool);
    if (ctx->d_part) cudaFree(ctx->d_part);
    if (ctx->d_ticket) cudaFree(ctx->d_ticket);
    if (ctx->d_res_dev) cudaFree(ctx->d_res_dev);
    if (ctx->d_all) cudaFree(ctx->d_all);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    if (ctx->halo_stream) {
        cudaStreamDestroy(ctx->halo_stream);
        cudaEventDestroy(ctx->ev_ready);
        cudaEventDestroy(ctx->ev_halo);
    }
    if (ctx->prof) {
        for (int i = 0; i < OPKB_PROF_RING; ++i) {
            cudaEventDestroy(ctx->prof[i].a);
            cudaEventDestroy(ctx->prof[i].b);
        }
        free(ctx->prof);
    }
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_user) cudaEventDestroy(ctx->ev_user);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    free(ctx);
    return 0;
}

// ---------------------------------------------------------------------------
// Device memory of the vectors: freed blocks are kept by the context and handed out again to
// requests of the same size (a solver workspace is 2m + 6 blocks of one size: the second solve
// in a process pays no cudaMalloc -- 2.7 ms per GiB on a B200, 0.15 s for the workspace of C4).
// All work of a context is ordered on its one stream, so a recycled block needs no synchronisation.
// A failed cudaMalloc releases the cache and retries; opkb_context_trim() releases it on demand;
// OPKB_POOL=0 disables it.
struct opkb_pool_block { void* ptr; size_t bytes; };

static int pool_enabled()
{
    static const int on = getenv("OPKB_POOL") ? atoi(getenv("OPKB_POOL")) : 1;
    return on;
}

extern "C" int opkb_context_trim(opkb_context* ctx)
{
    if (!ctx) return opkb_set_error(OPKB_EINVAL, "NULL context");
    if (ctx->pool_n > 0) {
        opkb_bind_device(ctx);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        for (int i = 0; i < ctx->pool_n; ++i) cudaFree(ctx->pool[i].ptr);
    }
    ctx->pool_n = 0;
    ctx->pool_bytes = 0;
    return 0;
}

extern "C" int64_t opkb_context_cached_bytes(const opkb_context* ctx) { return ctx ? (int64_t)ctx->pool_bytes : 0; }

int opkb_dev_alloc(opkb_context* ctx, size_t bytes, void** out)
{
    opkb_bind_device(ctx);
    for (int i = ctx->pool_n - 1; i >= 0; --i) {
        if (ctx->pool[i].bytes == bytes) {
            *out = ctx->pool[i].ptr;
            ctx->pool_bytes -= bytes;
            ctx->pool[i] = ctx->pool[--ctx->pool_n];
            return 0;
        }
    }
    cudaError_t err = cudaMalloc(out, bytes);
    if (err == cudaErrorMemoryAllocation && ctx->pool_n > 0) {
        cudaGetLastError();
        opkb_context_trim(ctx);
        err = cudaMalloc(out, bytes);
    }
    if (err != cudaSuccess) {
        *out = NULL;
        return opkb_check_cuda(err, "cudaMalloc(vector)");
    }
    return 0;
}

void opkb_dev_free(opkb_context* ctx, void* ptr, size_t bytes)
{
    if (!ptr) return;
    if (pool_enabled()) {
        if (ctx->pool_n == ctx->pool_cap) {
            const int cap = ctx->pool_cap ? 2 * ctx->pool_cap : 64;
            opkb_pool_block* p = (opkb_pool_block*)realloc(ctx->pool, sizeof(opkb_pool_block) * cap);
            if (p) { ctx->pool = p; ctx->pool_cap = cap; }
        }
        if (ctx->pool_n < ctx->pool_cap) {
            ctx->pool[ctx->pool_n].ptr = ptr;
            ctx->pool[ctx->pool_n].bytes = bytes;
            ctx->pool_n += 1;
            ctx->pool_bytes += bytes;
            return;
        }
    }
    opkb_bind_device(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ptr);
}

// The library's stream (a cudaStream_t), for callers that run their own kernels on the vectors:
// work submitted to THIS stream is ordered with the library's.  opkb_context_wait_stream makes
// the library's stream wait for everything submitted so far to another stream of the caller
// (e.g. the stream a user objective wrote the gradient on).
extern "C" void* opkb_context_stream(const opkb_context* ctx) { return ctx ? (void*)ctx->stream : NULL; }

extern "C" int opkb_context_wait_stream(opkb_context* ctx, void* user_stream)
{
    if (!ctx) return opkb_set_error(OPKB_EINVAL, "NULL context");
    opkb_bind_device(ctx);
    if (!ctx->ev_user) OPKB_CUDA(cudaEventCreateWithFlags(&ctx->ev_user, cudaEventDisableTiming));
    OPKB_CUDA(cudaEventRecord(ctx->ev_user, (cudaStream_t)user_stream));
    OPKB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_user, 0));
    return 0;
}

extern "C" int opkb_context_synchronize(opkb_context* ctx)
{
    if (!ctx) return opkb_set_error(OPKB_EINVAL, "NULL context");
    OPKB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int opkb_context_device(const opkb_context* ctx) { return ctx ? ctx->device : -1; }
extern "C" int opkb_context_sm_count(const opkb_context* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int64_t opkb_context_launch_count(const opkb_context* ctx)
{
    return ctx ? ctx->launches : 0;
}

extern "C" int opkb_timer_start(opkb_context* ctx)
{
    if (!ctx) return opkb_set_error(OPKB_EINVAL, "NULL context");
    opkb_bind_device(ctx);
    OPKB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return 0;
}

extern "C" int opkb_timer_stop(opkb_context* ctx, double* ms)
{
    if (!ctx || !ms) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    float t = 0;
    OPKB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    OPKB_CUDA(cudaEventSynchronize(ctx->ev1));
    OPKB_CUDA(cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
    *ms = (double)t;
    return 0;
}

// ---------------------------------------------------------------------------
// per-kernel profile: one CUDA-event pair per launch on the launching stream
static void prof_drain(opkb_context* ctx)
{
    if (ctx->prof_n == 0) return;
    cudaEventSynchronize(ctx->prof[ctx->prof_n - 1].b);
    for (int i = 0; i < ctx->prof_n; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->prof[i].a, ctx->prof[i].b) == cudaSuccess) {
            ctx->prof_ms[ctx->prof[i].cls] += ms;
            ctx->prof_cnt[ctx->prof[i].cls] += 1;
        }
    }
    ctx->prof_n = 0;
}

void opkb_prof_begin(opkb_context* ctx, int cls)
{
    opkb_bind_device(ctx);   // first thing every launch macro does: launches go to the context's device
    if (!ctx->prof_on) return;
    if (ctx->prof_n == OPKB_PROF_RING) prof_drain(ctx);
    ctx->prof[ctx->prof_n].cls = cls;
    cudaEventRecord(ctx->prof[ctx->prof_n].a, ctx->stream);
}

void opkb_prof_end(opkb_context* ctx)
{
    if (!ctx->prof_on) return;
    cudaEventRecord(ctx->prof[ctx->prof_n].b, ctx->stream);
    ctx->prof_n += 1;
}

extern "C" int opkb_profile_enable(opkb_context* ctx, int on)
{
    if (!ctx) return opkb_set_error(OPKB_EINVAL, "NULL context");
    if (on && !ctx->prof) {
        ctx->prof = (opkb_prof_rec*)calloc(OPKB_PROF_RING, sizeof(opkb_prof_rec));
        if (!ctx->prof) return opkb_set_error(OPKB_ENOMEM, "out of host memory");
        for (int i = 0; i < OPKB_PROF_RING; ++i) {
            OPKB_CUDA(cudaEventCreate(&ctx->prof[i].a));
            OPKB_CUDA(cudaEventCreate(&ctx->prof[i].b));
        }
    }
    if (!on && ctx->prof_on) prof_drain(ctx);
    ctx->prof_on = on ? 1 : 0;
    return 0;
}

extern "C" int opkb_profile_reset(opkb_context* ctx)
{
    if (!ctx) return opkb_set_error(OPKB_EINVAL, "NULL context");
    prof_drain(ctx);
    for (int k = 0; k < OPKB_K_NCLASS; ++k) {
        ctx->prof_ms[k] = 0;
        ctx->prof_cnt[k] = 0;
    }
    return 0;
}

extern "C" int opkb_profile_get(opkb_context* ctx, int kernel_class, double* ms_total, int64_t* count)
{
    if (!ctx || !ms_total || !count) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (kernel_class < 0 || kernel_class >= OPKB_K_NCLASS)
        return opkb_set_error(OPKB_EINVAL, "invalid kernel class");
    prof_drain(ctx);
    *ms_total = ctx->prof_ms[kernel_class];
    *count = ctx->prof_cnt[kernel_class];
    return 0;
}

extern "C" int opkb_flush_l2(opkb_context* ctx)
{
    if (!ctx) return opkb_set_error(OPKB_EINVAL, "NULL context");
    if (!ctx->flush_buf) {
        ctx->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB of L2
        OPKB_CUDA(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    OPKB_CUDA(cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
    return 0;
}

int opkb_grid_for(const opkb_context* ctx, int64_t nitems, int ctas_per_sm)
{
    // a multiple of the SM count (persistent grid-stride CTAs), shrunk for
    // small inputs; the grid only depends on (n, device) => deterministic
    int64_t want = (nitems + OPKB_BLOCK - 1) / OPKB_BLOCK;
    int64_t full = (int64_t)ctx->sm_count * ctas_per_sm;
    if (full > OPKB_MAX_GRID) full = OPKB_MAX_GRID;
    if (want < 1) want = 1;
    return (int)(want < full ? want : full);
}

// Wait until the last CTA of the latest reducing launch has stored its sequence number in
// `flag` (after its results, system-scope fence); a stream query every few microseconds catches
// launches that publish nothing, and errors.  Returns true if the flag was seen.
static inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();      // spin-wait hint: leaves the core's resources to its sibling thread
#elif defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#endif
}

// Polling with a budget.  Waits up to OPKB_SPIN_US microseconds (default 20000: every kernel of the
// BASELINE configurations finishes well within it) are pure spinning -- that is what the flag is
// for, and a nap costs at least the ~50 us timer slack of the kernel.  A longer wait (a slow user
// objective, a peer that lags or died, a far larger problem) goes on with naps of elapsed/64
// (2 us .. 200 us): the host core is then not pinned at 100 % while the wake-up delay stays below
// ~1.5 % of the wait.  OPKB_SPIN_US=200 trades that 1.5 % for an idle core on large problems too.
struct opkb_backoff {
    struct timespec t0;
    unsigned int it;
    long spin_ns;
};

static inline void backoff_begin(opkb_backoff* b)
{
    static const long spin_us = getenv("OPKB_SPIN_US") ? atol(getenv("OPKB_SPIN_US")) : 20000;
    b->it = 0;
    b->spin_ns = spin_us * 1000L;
    b->t0.tv_sec = 0;
    b->t0.tv_nsec = 0;
}

static inline void backoff_pause(opkb_backoff* b)
{
    cpu_relax();
    if ((++b->it & 255u) != 0u) return;
    struct timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
    if (b->t0.tv_sec == 0 && b->t0.tv_nsec == 0) {
        b->t0 = now;
        return;
    }
    const long long el = (long long)(now.tv_sec - b->t0.tv_sec) * 1000000000LL + (now.tv_nsec - b->t0.tv_nsec);
    if (el < b->spin_ns) return;
    long long nap = el / 64;
    if (nap < 2000) nap = 2000;
    if (nap > 200000) nap = 200000;
    struct timespec ts = {0, (long)nap};
    nanosleep(&ts, NULL);
}

// The flag is read with ACQUIRE semantics: the loads of the results that follow (h_res, the mailbox
// buffers) cannot be reordered before it by a weakly ordered host CPU (aarch64: Grace).
static bool wait_flag(opkb_context* ctx, volatile unsigned int* flag)
{
    const unsigned int want = ctx->seq;
    opkb_backoff bo;
    backoff_begin(&bo);
    for (unsigned int it = 1;; ++it) {
        if (__atomic_load_n((const unsigned int*)flag, __ATOMIC_ACQUIRE) == want) return true;
        if ((it & 4095u) == 0u && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) {
            // the stream has drained: either the launch published nothing, or its flag store has
            // landed by now (one last look, still with acquire semantics)
            return __atomic_load_n((const unsigned int*)flag, __ATOMIC_ACQUIRE) == want;
        }
        backoff_pause(&bo);
    }
}

// Several ranks, shared-memory mailbox (see opkb_context_): publish this rank's phase, wait for
// the others, combine in rank order.
static int opkb_mailbox_reduce(opkb_context* ctx, int nred, const int* kinds, double* out)
{
    unsigned char* mine = ctx->mbox + (size_t)ctx->rank * OPKB_MBOX_SLOT;
    if (!wait_flag(ctx, (volatile unsigned int*)mine)) OPKB_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned long long phase = ++ctx->phase;
    __atomic_store_n((unsigned long long*)(mine + 64), phase, __ATOMIC_RELEASE);
    const size_t buf = (size_t)(phase & 1ull) * OPKB_NRED_MAX;
    for (int r = 0; r < ctx->nranks; ++r) {
        unsigned char* slot = ctx->mbox + (size_t)r * OPKB_MBOX_SLOT;
        if (r != ctx->rank) {
            unsigned long long* ready = (unsigned long long*)(slot + 64);
            unsigned long long spins = 0;
            time_t t_start = 0;
            opkb_backoff bo;
            backoff_begin(&bo);
            while (__atomic_load_n(ready, __ATOMIC_ACQUIRE) < phase) {
                backoff_pause(&bo);
                if ((++spins & 0xfffffull) == 0ull) {   // (a nap every 256 spins: about once per second)
                    // a peer that died would hang every other rank for ever: give up after
                    // OPKB_MAILBOX_TIMEOUT seconds (default: half an hour; ranks may legitimately
                    // be far apart, e.g. a slow user objective on one shard)
                    static const long limit = getenv("OPKB_MAILBOX_TIMEOUT") ? atol(getenv("OPKB_MAILBOX_TIMEOUT")) : 1800;
                    const time_t now = time(NULL);
                    if (t_start == 0) t_start = now;
                    else if (now - t_start > limit)
                        return opkb_set_error(OPKB_ENCCL, "rank %d did not publish phase %llu within %ld s", r,
                                              phase, limit);
                }
            }
        }
        const double* res = (const double*)(slot + OPKB_MBOX_HDR) + buf;
        for (int k = 0; k < nred; ++k) {
            const double b = res[k];
            if (r == 0) { out[k] = b; continue; }
            const int kind = kinds ? kinds[k] : OPKB_RSUM;
            const double a = out[k];
            if (kind == OPKB_RSUM) out[k] = a + b;
            else if (kind == OPKB_RMAX) out[k] = (b > a ? b : a);
            else out[k] = (b < a ? b : a);
        }
    }
    return 0;
}

int opkb_finish_reduction(opkb_context* ctx, int nred, const int* kinds, double* out)
{
    if (nred > OPKB_NRED_MAX) return opkb_set_error(OPKB_EINVAL, "too many reduced scalars");
    if (ctx->nranks <= 1) {
        if (ctx->d_res != ctx->h_res)
            OPKB_CUDA(cudaMemcpyAsync(ctx->h_res, ctx->d_res, sizeof(double) * nred,
                                      cudaMemcpyDeviceToHost, ctx->stream));
        // poll the completion flag (sub-microsecond) instead of waking up through the driver
        bool seen = false;
        if (ctx->spin && ctx->d_res == ctx->h_res) seen = wait_flag(ctx, ctx->h_flag);
        if (!seen) OPKB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < nred; ++k) out[k] = ctx->h_res[k];
        return 0;
    }
    if (ctx->mbox) return opkb_mailbox_reduce(ctx, nred, kinds, out);
    // one packed all-gather per phase, then a fixed rank-order combination
    int rc = opkb_nccl_allgather_f64(ctx, ctx->d_res, ctx->d_all, nred);
    if (rc != 0) return rc;
    OPKB_CUDA(cudaMemcpyAsync(ctx->h_res, ctx->d_all, sizeof(double) * nred * ctx->nranks,
                              cudaMemcpyDeviceToHost, ctx->stream));
    OPKB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < nred; ++k) {
        int kind = kinds ? kinds[k] : OPKB_RSUM;
        double a = ctx->h_res[k];
        for (int r = 1; r < ctx->nranks; ++r) {
            double b = ctx->h_res[(size_t)r * nred + k];
            if (kind == OPKB_RSUM) a = a + b;
            else if (kind == OPKB_RMAX) a = (b > a ? b : a);
            else a = (b < a ? b : a);
        }
        out[k] = a;
    }
    return 0;
}

int opkb_chain_supported(const opkb_context* ctx)
{
    return ctx && ctx->nranks <= 1 && !ctx->mbox && ctx->spin && ctx->h_flag && ctx->d_res == ctx->h_res;
}

int opkb_finish_reduction_seq(opkb_context* ctx, unsigned int seq, int res_offset, int nred, double* out)
{
    if (!opkb_chain_supported(ctx)) return opkb_set_error(OPKB_EINVAL, "chained launches need one rank and the polled flag");
    if (res_offset < 0 || res_offset + nred > OPKB_NRED_MAX) return opkb_set_error(OPKB_EINVAL, "too many reduced scalars");
    volatile unsigned int* flag = ctx->h_flag;
    opkb_backoff bo;
    backoff_begin(&bo);
    bool seen = false;
    for (unsigned int it = 1;; ++it) {
        // (sequence numbers only grow; the difference is taken modulo 2^32)
        if ((int)(__atomic_load_n((const unsigned int*)flag, __ATOMIC_ACQUIRE) - seq) >= 0) { seen = true; break; }
        if ((it & 4095u) == 0u && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) {
            seen = (int)(__atomic_load_n((const unsigned int*)flag, __ATOMIC_ACQUIRE) - seq) >= 0;
            break;
        }
        backoff_pause(&bo);
    }
    if (!seen) {
        OPKB_CUDA(cudaStreamSynchronize(ctx->stream));
        return opkb_set_error(OPKB_EINVAL, "internal error: the chained launch %u published nothing", seq);
    }
    for (int k = 0; k < nred; ++k) out[k] = ctx->h_res[res_offset + k];
    return 0;
}

// ---------------------------------------------------------------------------
// vectors
static size_t elsize(int eltype) { return eltype == OPKB_DOUBLE ? 8 : 4; }

extern "C" int opkb_vector_create(opkb_context* ctx, int eltype, int64_t n, opkb_vector** out)
{
    if (!ctx || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (eltype != OPKB_FLOAT && eltype != OPKB_DOUBLE)
        return opkb_set_error(OPKB_EINVAL, "invalid element type %d", eltype);
    if (n < 0) return opkb_set_error(OPKB_EINVAL, "negative length");
    opkb_vector* v = (opkb_vector*)calloc(1, sizeof(opkb_vector));
    if (!v) return opkb_set_error(OPKB_ENOMEM, "out of host memory");
    v->ctx = ctx;
    v->eltype = eltype;
    v->n = n;
    v->owned = 1;
    size_t bytes = (size_t)(n > 0 ? n : 1) * elsize(eltype);
    bytes = (bytes + 255) & ~(size_t)255;
    int rc = opkb_dev_alloc(ctx, bytes, &v->data);
    if (rc != 0) {
        free(v);
        return rc;
    }
    *out = v;
    return 0;
}

extern "C" int opkb_vector_destroy(opkb_vector* v)
{
    if (!v) return 0;
    if (v->owned && v->data) {
        size_t bytes = (size_t)(v->n > 0 ? v->n : 1) * elsize(v->eltype);
        bytes = (bytes + 255) & ~(size_t)255;
        opkb_dev_free(v->ctx, v->data, bytes);
    }
    free(v);
    return 0;
}

extern "C" int64_t opkb_vector_length(const opkb_vector* v) { return v ? v->n : -1; }
extern "C" int opkb_vector_eltype(const opkb_vector* v) { return v ? v->eltype : -1; }
extern "C" void* opkb_vector_data(const opkb_vector* v) { return v ? v->data : NULL; }

extern "C" int opkb_vector_upload(opkb_vector* v, const void* host, int64_t offset, int64_t count)
{
    if (!v || (!host && count > 0)) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (offset < 0 || count < 0 || offset + count > v->n)
        return opkb_set_error(OPKB_EINVAL, "upload range out of bounds");
    size_t es = elsize(v->eltype);
    opkb_bind_device(v->ctx);
    OPKB_CUDA(cudaMemcpyAsync((char*)v->data + offset * es, host, count * es,
                              cudaMemcpyHostToDevice, v->ctx->stream));
    OPKB_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return 0;
}

extern "C" int opkb_vector_download(const opkb_vector* v, void* host, int64_t offset, int64_t count)
{
    if (!v || (!host && count > 0)) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (offset < 0 || count < 0 || offset + count > v->n)
        return opkb_set_error(OPKB_EINVAL, "download range out of bounds");
    size_t es = elsize(v->eltype);
    opkb_bind_device(v->ctx);
    OPKB_CUDA(cudaMemcpyAsync(host, (const char*)v->data + offset * es, count * es,
                              cudaMemcpyDeviceToHost, v->ctx->stream));
    OPKB_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// helpers for the argument checks of Tier 1
static int same_space(const opkb_vector* a, const opkb_vector* b)
{
    return a && b && a->ctx == b->ctx && a->eltype == b->eltype && a->n == b->n;
}

#define CHECK_CTX(ctx) \
    if (!(ctx)) return opkb_set_error(OPKB_EINVAL, "NULL context")
#define CHECK_SAME(a, b) \
    if (!same_space((a), (b))) return opkb_set_error(OPKB_EINVAL, "vectors of different spaces")

#define LAUNCH(ctx, cls, kernel, grid, ...)                                  \
    do {                                                                     \
        opkb_prof_begin((ctx), (cls));                                       \
        kernel<<<(grid), OPKB_BLOCK, 0, (ctx)->stream>>>(__VA_ARGS__);       \
        opkb_prof_end((ctx));                                                \
        (ctx)->launches += 1;                                                \
        OPKB_CUDA(cudaGetLastError());                                       \
    } while (0)

static ReduceScratch scratch(opkb_context* ctx)
{
    return opkb_make_scratch(ctx);
}

template <typename T>
static Bound<T> make_bound(const opkb_vector* arr, double val, bool lower)
{
    Bound<T> b;
    b.arr = arr ? (const T*)arr->data : NULL;
    b.val = (T)val;
    b.has = arr ? 1 : (lower ? (b.val > -(T)INFINITY) : (b.val < (T)INFINITY));
    return b;
}

static int check_bounds_args(const opkb_vector* x, const opkb_vector* lo, const opkb_vector* hi)
{
    if (lo && !same_space(x, lo)) return opkb_set_error(OPKB_EINVAL, "lower bound of a different space");
    if (hi && !same_space(x, hi)) return opkb_set_error(OPKB_EINVAL, "upper bound of a different space");
    return 0;
}

// ---------------------------------------------------------------------------
// fill kernels
template <typename T>
__global__ void k_fill(T* x, int64_t n, T value)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = value;
}

extern "C" int opkb_vector_fill(opkb_vector* v, double value)
{
    if (!v) return opkb_set_error(OPKB_EINVAL, "NULL vector");
    opkb_context* ctx = v->ctx;
    int grid = opkb_grid_for(ctx, v->n, 8);
    if (v->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_OTHER, k_fill<double>, grid, (double*)v->data, v->n, value);
    else
        LAUNCH(ctx, OPKB_K_OTHER, k_fill<float>, grid, (float*)v->data, v->n, (float)value);
    return 0;
}

// splitmix64 of (seed, index): the counter-based generator of SURVEY 8(d);
// the package's host generator computes the same integers.
__host__ __device__ __forceinline__ uint64_t opkb_splitmix64(uint64_t seed, uint64_t i)
{
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
__global__ void k_fill_random(T* x, int64_t n, uint64_t seed, int64_t first, double lo, double hi,
                              int log_scale)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t z = opkb_splitmix64(seed, (uint64_t)(first + i));
        double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);  // [0,1), 53 bits
        double r = log_scale ? lo * exp2(u * log2(hi / lo)) : lo + (hi - lo) * u;
        x[i] = (T)r;
    }
}

extern "C" int opkb_vector_fill_random(opkb_vector* v, uint64_t seed, int64_t first, double lo,
                                       double hi, int log_scale)
{
    if (!v) return opkb_set_error(OPKB_EINVAL, "NULL vector");
    opkb_context* ctx = v->ctx;
    int grid = opkb_grid_for(ctx, v->n, 8);
    if (v->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_OTHER, k_fill_random<double>, grid, (double*)v->data, v->n, seed, first, lo, hi,
               log_scale);
    else
        LAUNCH(ctx, OPKB_K_OTHER, k_fill_random<float>, grid, (float*)v->data, v->n, seed, first, lo, hi,
               log_scale);
    return 0;
}

// ---------------------------------------------------------------------------
// reductions: vdot, vdot_masked, vnorm2, vnorminf
// MODE 0: sum x*y; 1: sum x*y over w != 0; 2: max |x|
template <typename T, int MODE>
__global__ void __launch_bounds__(OPKB_BLOCK)
k_reduce(const T* x, const T* y, const T* w, int64_t n, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc[1];
    acc[0] = 0.0;
    // two packs per trip: 2-4 independent 16-byte loads in flight per stream
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += 2 * stride) {
        const int64_t iv2 = iv + stride;
        const bool two = iv2 < nvec;
        Pack<T> a0 = ld_pack<T>(x, iv), a1, b0, b1, w0, w1;
        if (two) a1 = ld_pack<T>(x, iv2);
        if (MODE < 2) {
            b0 = ld_pack<T>(y, iv);
            if (two) b1 = ld_pack<T>(y, iv2);
        }
        if (MODE == 1) {
            w0 = ld_pack<T>(w, iv);
            if (two) w1 = ld_pack<T>(w, iv2);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (MODE == 0) acc[0] = dacc(a0.v[e], b0.v[e], acc[0]);
            if (MODE == 1 && w0.v[e] != T(0)) acc[0] = dacc(a0.v[e], b0.v[e], acc[0]);
            if (MODE == 2) { double t = fabs((double)a0.v[e]); acc[0] = (t > acc[0] ? t : acc[0]); }
            if (MODE == 3) acc[0] += fabs((double)a0.v[e]);
        }
        if (two) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (MODE == 0) acc[0] = dacc(a1.v[e], b1.v[e], acc[0]);
                if (MODE == 1 && w1.v[e] != T(0)) acc[0] = dacc(a1.v[e], b1.v[e], acc[0]);
                if (MODE == 2) { double t = fabs((double)a1.v[e]); acc[0] = (t > acc[0] ? t : acc[0]); }
                if (MODE == 3) acc[0] += fabs((double)a1.v[e]);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < n; ++i) {
            if (MODE == 0) acc[0] = dacc(x[i], y[i], acc[0]);
            if (MODE == 1 && w[i] != T(0)) acc[0] = dacc(x[i], y[i], acc[0]);
            if (MODE == 2) { double t = fabs((double)x[i]); acc[0] = (t > acc[0] ? t : acc[0]); }
            if (MODE == 3) acc[0] += fabs((double)x[i]);
        }
    }
    if (MODE == 2)
        block_reduce_publish<1, 1u, 0u>(acc, rs);
    else
        block_reduce_publish<1, 0u, 0u>(acc, rs);
}

template <int MODE>
static int reduce_dispatch(opkb_context* ctx, const opkb_vector* x, const opkb_vector* y,
                           const opkb_vector* w, double* result)
{
    int grid = opkb_grid_for(ctx, x->n / (x->eltype == OPKB_DOUBLE ? 2 : 4), 4);
    if (x->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_REDUCE, (k_reduce<double, MODE>), grid, (const double*)x->data,
               y ? (const double*)y->data : NULL, w ? (const double*)w->data : NULL, x->n,
               scratch(ctx));
    else
        LAUNCH(ctx, OPKB_K_REDUCE, (k_reduce<float, MODE>), grid, (const float*)x->data,
               y ? (const float*)y->data : NULL, w ? (const float*)w->data : NULL, x->n,
               scratch(ctx));
    int kind = (MODE == 2 ? OPKB_RMAX : OPKB_RSUM);
    return opkb_finish_reduction(ctx, 1, &kind, result);
}

extern "C" int opkb_vdot(opkb_context* ctx, const opkb_vector* x, const opkb_vector* y, double* result)
{
    CHECK_CTX(ctx);
    CHECK_SAME(x, y);
    if (!result) return opkb_set_error(OPKB_EINVAL, "NULL result");
    return reduce_dispatch<0>(ctx, x, y, NULL, result);
}

extern "C" int opkb_vdot_masked(opkb_context* ctx, const opkb_vector* w, const opkb_vector* x,
                                const opkb_vector* y, double* result)
{
    CHECK_CTX(ctx);
    CHECK_SAME(x, y);
    CHECK_SAME(x, w);
    if (!result) return opkb_set_error(OPKB_EINVAL, "NULL result");
    return reduce_dispatch<1>(ctx, x, y, w, result);
}

extern "C" int opkb_vnorm2(opkb_context* ctx, const opkb_vector* x, double* result)
{
    CHECK_CTX(ctx);
    if (!x || !result) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    double s = 0;
    int rc = reduce_dispatch<0>(ctx, x, x, NULL, &s);
    if (rc == 0) *result = sqrt(s);
    return rc;
}

extern "C" int opkb_vnorminf(opkb_context* ctx, const opkb_vector* x, double* result)
{
    CHECK_CTX(ctx);
    if (!x || !result) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    return reduce_dispatch<2>(ctx, x, NULL, NULL, result);
}

// ---------------------------------------------------------------------------
// element-wise primitives.  OP: 0 copy, 1 scale, 2 update (y += a x),
// 3 masked update, 4 combine (dst = a x + b y), 5 product.
// CASE selects the special multipliers of test/vectors-tests.jl:26-81.
struct EwArgs {
    void* dst;
    const void* x;
    const void* y;
    const void* w;
    double alpha, beta;
    int64_t n;
    int acase, bcase;  // 0: zero, 1: one, 2: minus one, 3: general
};

template <typename T, int OP>
__device__ __forceinline__ T ew_apply(T dst, T x, T y, T w, T a, T b, int acase, int bcase)
{
    if (OP == 0) return x;
    if (OP == 1) {  // vscale!(x, alpha): dst aliases x
        if (acase == 0) return T(0);
        if (acase == 2) return -x;
        return a * x;
    }
    if (OP == 2 || OP == 3) {  // y += alpha*x (dst aliases y, passed as `dst`)
        if (OP == 3 && !(w != T(0))) return dst;
        if (acase == 1) return dst + x;
        if (acase == 2) return dst - x;
        return dst + a * x;
    }
    if (OP == 4) {
        if (acase == 0 && bcase == 0) return T(0);
        if (acase == 0) return b * y;
        if (bcase == 0) return a * x;
        return a * x + b * y;  // +-1 multipliers are exact: same bits as x +- y
    }
    return x * y;  // OP == 5
}

template <typename T, int OP>
__global__ void __launch_bounds__(OPKB_BLOCK) k_elementwise(EwArgs A)
{
    constexpr int VEC = VecOf<T>::N;
    T* dst = (T*)A.dst;
    const T* x = (const T*)A.x;
    const T* y = (const T*)A.y;
    const T* w = (const T*)A.w;
    const T a = (T)A.alpha, b = (T)A.beta;
    const int64_t nvec = A.n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool need_dst = (OP == 2 || OP == 3);
    const bool need_x = (OP != 4) || A.acase != 0;
    const bool need_y = (OP == 5) || (OP == 4 && A.bcase != 0);
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> pd = splat_pack<T>(T(0)), px = pd, py = pd, pw = pd, out;
        if (need_dst) pd = ld_pack<T>(dst, iv);
        if (need_x) px = ld_pack<T>(x, iv);
        if (need_y) py = ld_pack<T>(y, iv);
        if (OP == 3) pw = ld_pack<T>(w, iv);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            out.v[e] = ew_apply<T, OP>(pd.v[e], px.v[e], py.v[e], pw.v[e], a, b, A.acase, A.bcase);
        st_pack<T>(dst, iv, out);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < A.n; ++i) {
            T vd = need_dst ? dst[i] : T(0), vx = need_x ? x[i] : T(0), vy = need_y ? y[i] : T(0);
            T vw = (OP == 3) ? w[i] : T(0);
            dst[i] = ew_apply<T, OP>(vd, vx, vy, vw, a, b, A.acase, A.bcase);
        }
    }
}

static int mult_case(double a) { return a == 0 ? 0 : (a == 1 ? 1 : (a == -1 ? 2 : 3)); }

template <int OP>
static int ew_dispatch(opkb_context* ctx, opkb_vector* dst, const opkb_vector* x, const opkb_vector* y,
                       const opkb_vector* w, double alpha, double beta)
{
    EwArgs A;
    A.dst = dst->data;
    A.x = x ? x->data : NULL;
    A.y = y ? y->data : NULL;
    A.w = w ? w->data : NULL;
    A.alpha = alpha;
    A.beta = beta;
    A.n = dst->n;
    A.acase = mult_case(alpha);
    A.bcase = mult_case(beta);
    int grid = opkb_grid_for(ctx, dst->n / (dst->eltype == OPKB_DOUBLE ? 2 : 4), 8);
    if (dst->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_ELEMENTWISE, (k_elementwise<double, OP>), grid, A);
    else
        LAUNCH(ctx, OPKB_K_ELEMENTWISE, (k_elementwise<float, OP>), grid, A);
    return 0;
}

extern "C" int opkb_vcopy(opkb_context* ctx, opkb_vector* dst, const opkb_vector* src)
{
    CHECK_CTX(ctx);
    CHECK_SAME(dst, src);
    if (dst->data == src->data) return 0;
    return ew_dispatch<0>(ctx, dst, src, NULL, NULL, 1, 0);
}

extern "C" int opkb_vscale(opkb_context* ctx, opkb_vector* x, double alpha)
{
    CHECK_CTX(ctx);
    if (!x) return opkb_set_error(OPKB_EINVAL, "NULL vector");
    if (alpha == 1) return 0;
    return ew_dispatch<1>(ctx, x, x, NULL, NULL, alpha, 0);
}

extern "C" int opkb_vupdate(opkb_context* ctx, opkb_vector* y, double alpha, const opkb_vector* x)
{
    CHECK_CTX(ctx);
    CHECK_SAME(y, x);
    if (alpha == 0) return 0;
    return ew_dispatch<2>(ctx, y, x, NULL, NULL, alpha, 0);
}

extern "C" int opkb_vupdate_masked(opkb_context* ctx, opkb_vector* y, const opkb_vector* w,
                                   double alpha, const opkb_vector* x)
{
    CHECK_CTX(ctx);
    CHECK_SAME(y, x);
    CHECK_SAME(y, w);
    if (alpha == 0) return 0;
    // the reference's vupdate!(y, sel, alpha, x) has no +-1 special case that
    // would change bits: y + 1*x == y + x
    return ew_dispatch<3>(ctx, y, x, NULL, w, alpha, 0);
}

extern "C" int opkb_vcombine(opkb_context* ctx, opkb_vector* dst, double alpha, const opkb_vector* x,
                             double beta, const opkb_vector* y)
{
    CHECK_CTX(ctx);
    CHECK_SAME(dst, x);
    CHECK_SAME(dst, y);
    return ew_dispatch<4>(ctx, dst, x, y, NULL, alpha, beta);
}

extern "C" int opkb_vproduct(opkb_context* ctx, opkb_vector* dst, const opkb_vector* x,
                             const opkb_vector* y)
{
    CHECK_CTX(ctx);
    CHECK_SAME(dst, x);
    CHECK_SAME(dst, y);
    return ew_dispatch<5>(ctx, dst, x, y, NULL, 1, 1);
}

// ---------------------------------------------------------------------------
// the remaining operators of the vector-space contract (doc/algebra.md:31, :50, :60-67, :116):
// vswap!, vnorm1, vcombine! with three terms, vproduct! restricted to a free set.  None is
// called by the loops of vmlmb / spg; they complete the plug-in interface.
template <typename T, int OP>
__global__ void __launch_bounds__(OPKB_BLOCK)
k_contract(T* a, T* b, const T* x, const T* y, const T* z, T ca, T cb, T cc, int64_t n)
{
    // OP 0: swap(a, b); 1: a = (ca*x + cb*y) + cc*z; 2: a = x.*y where z != 0
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        if (OP == 0) {
            const Pack<T> pa = ld_pack<T>(a, iv), pb = ld_pack<T>(b, iv);
            st_pack<T>(a, iv, pb);
            st_pack<T>(b, iv, pa);
        } else if (OP == 1) {
            const Pack<T> px = ld_pack<T>(x, iv), py = ld_pack<T>(y, iv), pz = ld_pack<T>(z, iv);
            Pack<T> o;
#pragma unroll
            for (int e = 0; e < VEC; ++e) o.v[e] = (ca * px.v[e] + cb * py.v[e]) + cc * pz.v[e];
            st_pack<T>(a, iv, o);
        } else {
            const Pack<T> px = ld_pack<T>(x, iv), py = ld_pack<T>(y, iv), pz = ld_pack<T>(z, iv);
            Pack<T> o = ld_pack<T>(a, iv);
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                if (pz.v[e] != T(0)) o.v[e] = px.v[e] * py.v[e];
            st_pack<T>(a, iv, o);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < n; ++i) {
            if (OP == 0) {
                const T t = a[i];
                a[i] = b[i];
                b[i] = t;
            } else if (OP == 1) {
                a[i] = (ca * x[i] + cb * y[i]) + cc * z[i];
            } else if (z[i] != T(0)) {
                a[i] = x[i] * y[i];
            }
        }
    }
}

template <int OP>
static int contract_dispatch(opkb_context* ctx, opkb_vector* a, opkb_vector* b, const opkb_vector* x,
                             const opkb_vector* y, const opkb_vector* z, double ca, double cb, double cc)
{
    int grid = opkb_grid_for(ctx, a->n / (a->eltype == OPKB_DOUBLE ? 2 : 4), 8);
    if (a->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_ELEMENTWISE, (k_contract<double, OP>), grid, (double*)a->data,
               b ? (double*)b->data : NULL, x ? (const double*)x->data : NULL,
               y ? (const double*)y->data : NULL, z ? (const double*)z->data : NULL, ca, cb, cc, a->n);
    else
        LAUNCH(ctx, OPKB_K_ELEMENTWISE, (k_contract<float, OP>), grid, (float*)a->data,
               b ? (float*)b->data : NULL, x ? (const float*)x->data : NULL, y ? (const float*)y->data : NULL,
               z ? (const float*)z->data : NULL, (float)ca, (float)cb, (float)cc, a->n);
    return 0;
}

extern "C" int opkb_vswap(opkb_context* ctx, opkb_vector* x, opkb_vector* y)
{
    CHECK_CTX(ctx);
    CHECK_SAME(x, y);
    if (x == y || x->data == y->data) return 0;
    if (x->owned && y->owned) {          // both own their storage: exchanging the buffers is the swap
        void* t = x->data;
        x->data = y->data;
        y->data = t;
        return 0;
    }
    return contract_dispatch<0>(ctx, x, y, NULL, NULL, NULL, 0, 0, 0);
}

extern "C" int opkb_vcombine3(opkb_context* ctx, opkb_vector* dst, double alpha, const opkb_vector* x,
                              double beta, const opkb_vector* y, double gamma, const opkb_vector* z)
{
    CHECK_CTX(ctx);
    CHECK_SAME(dst, x);
    CHECK_SAME(dst, y);
    CHECK_SAME(dst, z);
    return contract_dispatch<1>(ctx, dst, NULL, x, y, z, alpha, beta, gamma);
}

extern "C" int opkb_vproduct_masked(opkb_context* ctx, opkb_vector* dst, const opkb_vector* w,
                                    const opkb_vector* x, const opkb_vector* y)
{
    CHECK_CTX(ctx);
    CHECK_SAME(dst, x);
    CHECK_SAME(dst, y);
    CHECK_SAME(dst, w);
    return contract_dispatch<2>(ctx, dst, NULL, x, y, w, 0, 0, 0);
}

extern "C" int opkb_vnorm1(opkb_context* ctx, const opkb_vector* x, double* result)
{
    CHECK_CTX(ctx);
    if (!x || !result) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    return reduce_dispatch<3>(ctx, x, x, NULL, result);
}

// ---------------------------------------------------------------------------
// bounds primitives
template <typename T>
__global__ void __launch_bounds__(OPKB_BLOCK)
k_project_variables(T* dst, const T* src, Bound<T> lo, Bound<T> hi, int64_t n)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> s = ld_pack<T>(src, iv), l = lo.pack(iv), h = hi.pack(iv), o;
#pragma unroll
        for (int e = 0; e < VEC; ++e) o.v[e] = fastclamp(s.v[e], l.v[e], h.v[e]);
        st_pack<T>(dst, iv, o);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = nvec * VEC; i < n; ++i) dst[i] = fastclamp(src[i], lo.at(i), hi.at(i));
}

extern "C" int opkb_project_variables(opkb_context* ctx, opkb_vector* dst, const opkb_vector* src,
                                      const opkb_vector* lo, double lo_val, const opkb_vector* hi,
                                      double hi_val)
{
    CHECK_CTX(ctx);
    CHECK_SAME(dst, src);
    int rc = check_bounds_args(dst, lo, hi);
    if (rc) return rc;
    int grid = opkb_grid_for(ctx, dst->n / (dst->eltype == OPKB_DOUBLE ? 2 : 4), 8);
    if (dst->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_BOUNDS, k_project_variables<double>, grid, (double*)dst->data, (const double*)src->data,
               make_bound<double>(lo, lo_val, true), make_bound<double>(hi, hi_val, false), dst->n);
    else
        LAUNCH(ctx, OPKB_K_BOUNDS, k_project_variables<float>, grid, (float*)dst->data, (const float*)src->data,
               make_bound<float>(lo, lo_val, true), make_bound<float>(hi, hi_val, false), dst->n);
    return 0;
}

template <typename T>
__global__ void __launch_bounds__(OPKB_BLOCK)
k_project_direction(T* dst, const T* x, Bound<T> lo, Bound<T> hi, int orient, const T* d, int64_t n)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> px = ld_pack<T>(x, iv), pd = ld_pack<T>(d, iv), l = lo.pack(iv), h = hi.pack(iv), o;
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            o.v[e] = projdir(px.v[e], l.v[e], h.v[e], lo.has, hi.has, orient, pd.v[e]);
        st_pack<T>(dst, iv, o);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = nvec * VEC; i < n; ++i)
            dst[i] = projdir(x[i], lo.at(i), hi.at(i), lo.has, hi.has, orient, d[i]);
}

static int scalar_bounds_valid(const opkb_vector* lo, double lo_val, const opkb_vector* hi,
                               double hi_val, int eltype)
{
    // `lo <= hi || argument_error("invalid bounds")` (also catches NaN):
    // only the scalar/scalar methods check, src/bounds.jl:339, :450
    if (lo || hi) return 1;
    if (eltype == OPKB_DOUBLE) return lo_val <= hi_val;
    return (float)lo_val <= (float)hi_val;
}

extern "C" int opkb_project_direction(opkb_context* ctx, opkb_vector* dst, const opkb_vector* x,
                                      const opkb_vector* lo, double lo_val, const opkb_vector* hi,
                                      double hi_val, int orient, const opkb_vector* d)
{
    CHECK_CTX(ctx);
    CHECK_SAME(dst, x);
    CHECK_SAME(dst, d);
    int rc = check_bounds_args(dst, lo, hi);
    if (rc) return rc;
    if (orient == 0) return opkb_set_error(OPKB_EORIENT, "invalid orientation");
    if (!scalar_bounds_valid(lo, lo_val, hi, hi_val, dst->eltype))
        return opkb_set_error(OPKB_EBOUNDS, "invalid bounds");
    int grid = opkb_grid_for(ctx, dst->n / (dst->eltype == OPKB_DOUBLE ? 2 : 4), 8);
    if (dst->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_BOUNDS, k_project_direction<double>, grid, (double*)dst->data, (const double*)x->data,
               make_bound<double>(lo, lo_val, true), make_bound<double>(hi, hi_val, false), orient,
               (const double*)d->data, dst->n);
    else
        LAUNCH(ctx, OPKB_K_BOUNDS, k_project_direction<float>, grid, (float*)dst->data, (const float*)x->data,
               make_bound<float>(lo, lo_val, true), make_bound<float>(hi, hi_val, false), orient,
               (const float*)d->data, dst->n);
    return 0;
}

template <typename T>
__global__ void __launch_bounds__(OPKB_BLOCK)
k_step_limits(const T* x, Bound<T> lo, Bound<T> hi, T s, const T* d, int64_t n, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    T smin = T(INFINITY), smax = T(0);
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> px = ld_pack<T>(x, iv), pd = ld_pack<T>(d, iv), l = lo.pack(iv), h = hi.pack(iv);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            step_limit_update(px.v[e], l.v[e], h.v[e], lo.has, hi.has, s * pd.v[e], smin, smax);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = nvec * VEC; i < n; ++i)
            step_limit_update(x[i], lo.at(i), hi.at(i), lo.has, hi.has, s * d[i], smin, smax);
    double acc[2] = {(double)smin, (double)smax};
    block_reduce_publish<2, 2u, 1u>(acc, rs);  // slot 0: min, slot 1: max
}

extern "C" int opkb_step_limits(opkb_context* ctx, const opkb_vector* x, const opkb_vector* lo,
                                double lo_val, const opkb_vector* hi, double hi_val, int orient,
                                const opkb_vector* d, double* smin, double* smax)
{
    CHECK_CTX(ctx);
    CHECK_SAME(x, d);
    int rc = check_bounds_args(x, lo, hi);
    if (rc) return rc;
    if (!smin || !smax) return opkb_set_error(OPKB_EINVAL, "NULL result");
    if (orient == 0) return opkb_set_error(OPKB_EORIENT, "invalid orientation");
    if (!scalar_bounds_valid(lo, lo_val, hi, hi_val, x->eltype))
        return opkb_set_error(OPKB_EBOUNDS, "invalid bounds");
    bool has_lo = lo || (x->eltype == OPKB_DOUBLE ? lo_val > -INFINITY : (float)lo_val > -INFINITY);
    bool has_hi = hi || (x->eltype == OPKB_DOUBLE ? hi_val < INFINITY : (float)hi_val < INFINITY);
    if (!has_lo && !has_hi) {  // src/bounds.jl:503-505
        *smin = INFINITY;
        *smax = INFINITY;
        return 0;
    }
    int grid = opkb_grid_for(ctx, x->n / (x->eltype == OPKB_DOUBLE ? 2 : 4), 4);
    if (x->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_BOUNDS, k_step_limits<double>, grid, (const double*)x->data,
               make_bound<double>(lo, lo_val, true), make_bound<double>(hi, hi_val, false),
               (double)(orient > 0 ? 1 : -1), (const double*)d->data, x->n, scratch(ctx));
    else
        LAUNCH(ctx, OPKB_K_BOUNDS, k_step_limits<float>, grid, (const float*)x->data,
               make_bound<float>(lo, lo_val, true), make_bound<float>(hi, hi_val, false),
               (float)(orient > 0 ? 1 : -1), (const float*)d->data, x->n, scratch(ctx));
    int kinds[2] = {OPKB_RMIN, OPKB_RMAX};
    double out[2];
    rc = opkb_finish_reduction(ctx, 2, kinds, out);
    if (rc == 0) {
        *smin = out[0];
        *smax = out[1];
    }
    return rc;
}

// free-variable mask: count (and optionally the byte mask) of gp != 0
template <typename T>
__global__ void __launch_bounds__(OPKB_BLOCK)
k_free_variables(const T* gp, int64_t n, uint8_t* mask, ReduceScratch rs)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool f = (gp[i] != T(0));
        if (mask) mask[i] = f ? 1 : 0;
        acc[0] += f ? 1.0 : 0.0;
    }
    block_reduce_publish<1, 0u, 0u>(acc, rs);
}

extern "C" int opkb_free_variables(opkb_context* ctx, const opkb_vector* gp, int64_t* count,
                                   uint8_t* mask_bytes)
{
    CHECK_CTX(ctx);
    if (!gp || !count) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    uint8_t* dmask = NULL;
    if (mask_bytes && gp->n > 0) OPKB_CUDA(cudaMalloc(&dmask, (size_t)gp->n));
    int grid = opkb_grid_for(ctx, gp->n, 4);
    if (gp->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_BOUNDS, k_free_variables<double>, grid, (const double*)gp->data, gp->n, dmask, scratch(ctx));
    else
        LAUNCH(ctx, OPKB_K_BOUNDS, k_free_variables<float>, grid, (const float*)gp->data, gp->n, dmask, scratch(ctx));
    double c = 0;
    int rc = opkb_finish_reduction(ctx, 1, NULL, &c);
    if (rc == 0) {
        *count = (int64_t)c;
        if (dmask)
            rc = opkb_check_cuda(cudaMemcpy(mask_bytes, dmask, (size_t)gp->n, cudaMemcpyDeviceToHost),
                                 "cudaMemcpy(mask)");
    }
    if (dmask) cudaFree(dmask);
    return rc;
}

// ---------------------------------------------------------------------------
// stand-alone objectives: fg!(x, g) -> f
template <typename T, int OBJ>
__global__ void __launch_bounds__(OPKB_BLOCK)
k_objective(const T* x, T* g, const T* qa, const T* qc, int64_t n, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc[2] = {0.0, 0.0};
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> px = ld_pack<T>(x, iv), pg;
        if (OBJ == OPKB_OBJ_ROSENBROCK) {
#pragma unroll
            for (int e = 0; e < VEC; e += 2)
                rosenbrock_pair(px.v[e], px.v[e + 1], pg.v[e], pg.v[e + 1], acc[0], acc[1]);
        } else {
            Pack<T> pa = ld_pack<T>(qa, iv), pc = ld_pack<T>(qc, iv);
#pragma unroll
            for (int e = 0; e < VEC; ++e) quadratic_elem(px.v[e], pa.v[e], pc.v[e], pg.v[e], acc[0]);
        }
        st_pack<T>(g, iv, pg);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (OBJ == OPKB_OBJ_ROSENBROCK) {
            for (int64_t i = nvec * VEC; i + 1 < n; i += 2)
                rosenbrock_pair(x[i], x[i + 1], g[i], g[i + 1], acc[0], acc[1]);
        } else {
            for (int64_t i = nvec * VEC; i < n; ++i) quadratic_elem(x[i], qa[i], qc[i], g[i], acc[0]);
        }
    }
    block_reduce_publish<2, 0u, 0u>(acc, rs);
}

extern "C" int opkb_rosenbrock_fg(opkb_context* ctx, const opkb_vector* x, opkb_vector* g, double* f)
{
    CHECK_CTX(ctx);
    CHECK_SAME(x, g);
    if (!f) return opkb_set_error(OPKB_EINVAL, "NULL result");
    if (x->n % 2) return opkb_set_error(OPKB_EINVAL, "Rosenbrock needs an even number of variables");
    int grid = opkb_grid_for(ctx, x->n / (x->eltype == OPKB_DOUBLE ? 2 : 4), 4);
    if (x->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_OBJECTIVE, (k_objective<double, OPKB_OBJ_ROSENBROCK>), grid, (const double*)x->data,
               (double*)g->data, (const double*)NULL, (const double*)NULL, x->n, scratch(ctx));
    else
        LAUNCH(ctx, OPKB_K_OBJECTIVE, (k_objective<float, OPKB_OBJ_ROSENBROCK>), grid, (const float*)x->data,
               (float*)g->data, (const float*)NULL, (const float*)NULL, x->n, scratch(ctx));
    double out[2];
    int rc = opkb_finish_reduction(ctx, 2, NULL, out);
    if (rc == 0) *f = out[0] + out[1];
    return rc;
}

extern "C" int opkb_quadratic_fg(opkb_context* ctx, const opkb_vector* a, const opkb_vector* c,
                                 const opkb_vector* x, opkb_vector* g, double* f)
{
    CHECK_CTX(ctx);
    CHECK_SAME(x, g);
    CHECK_SAME(x, a);
    CHECK_SAME(x, c);
    if (!f) return opkb_set_error(OPKB_EINVAL, "NULL result");
    int grid = opkb_grid_for(ctx, x->n / (x->eltype == OPKB_DOUBLE ? 2 : 4), 4);
    if (x->eltype == OPKB_DOUBLE)
        LAUNCH(ctx, OPKB_K_OBJECTIVE, (k_objective<double, OPKB_OBJ_QUADRATIC>), grid, (const double*)x->data,
               (double*)g->data, (const double*)a->data, (const double*)c->data, x->n, scratch(ctx));
    else
        LAUNCH(ctx, OPKB_K_OBJECTIVE, (k_objective<float, OPKB_OBJ_QUADRATIC>), grid, (const float*)x->data,
               (float*)g->data, (const float*)a->data, (const float*)c->data, x->n, scratch(ctx));
    double out[2];
    int rc = opkb_finish_reduction(ctx, 2, NULL, out);
    if (rc == 0) *f = 0.5 * out[0];
    return rc;
}
