// opkb_internal.cuh -- shared internals of libopkb200.so (sm_100a only).
//
// Built with -fmad=false: `a*x + b*y` written in plain C++ is two roundings,
// as in the reference's Julia loops (no contraction); the reductions use
// explicit fma() on exact double products instead.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "opkb200.h"

#define OPKB_BLOCK 256          // threads per CTA of the streaming kernels
#define OPKB_MAX_GRID 2048      // upper bound of any grid (partials buffer)
#define OPKB_NRED_MAX 1024      // scalars one launch may reduce
#define OPKB_MEM_MAX 20         // pairs the Gram-form kernels (Gram sweep, fused direction) support
#define OPKB_SLOTS_MAX 256      // pairs a workspace can store (beyond OPKB_MEM_MAX: sequential two-loop only)

// kernel classes of the per-kernel profile (opkb_profile_*)
enum { OPKB_K_OTHER = 0, OPKB_K_TRIAL = 1, OPKB_K_GRAM = 2, OPKB_K_DIRECTION = 3, OPKB_K_SPG = 4,
       OPKB_K_REDUCE = 5, OPKB_K_ELEMENTWISE = 6, OPKB_K_BOUNDS = 7, OPKB_K_OBJECTIVE = 8,
       OPKB_K_NCLASS = 9 };
#define OPKB_PROF_RING 4096

struct opkb_prof_rec { cudaEvent_t a, b; int cls; };

// reduction kinds of one scalar slot
enum { OPKB_RSUM = 0, OPKB_RMAX = 1, OPKB_RMIN = 2 };

struct opkb_context_ {
    int device;
    int sm_count;
    cudaStream_t stream;
    int rank, nranks;
    void* nccl_comm;
    double* d_part;          // [OPKB_MAX_GRID * OPKB_NRED_MAX] per-CTA partials
    unsigned int* d_ticket;  // last-CTA-done ticket
    double* d_res;           // [OPKB_NRED_MAX] this rank's reduced scalars; on one rank this IS h_res
                             // (pinned host memory is device-accessible under UVA: the last CTA
                             // writes the few hundred bytes straight to the host, no D2H copy)
    double* d_res_dev;       // the device allocation (used with several ranks: NCCL send buffer)
    double* d_all;           // [nranks * OPKB_NRED_MAX] after the all-gather
    double* h_res;           // pinned mirror of d_res / d_all
    // Several ranks on one node: a POSIX shared-memory segment, registered with CUDA in every
    // process (opkb_comm.cu).  Slot r = { completion flag of rank r's kernels (GPU-written) |
    // `ready` = last phase rank r has published (host-written) | two result buffers, by phase
    // parity }.  The last CTA of a reducing kernel writes its partials and the flag straight
    // into its rank's slot, every process polls the slots of all ranks and combines them in
    // rank order: no NCCL call, no D2H copy, no stream synchronisation per phase.
    unsigned char* mbox;     // host address of the segment (NULL: packed ncclAllGather instead)
    unsigned char* mbox_dev; // device address of the same memory
    size_t mbox_bytes;
    unsigned long long phase;  // phases (opkb_finish_reduction calls) completed so far
    unsigned int* h_flag;    // completion flag of the reducing kernels (pinned, see ReduceScratch)
    unsigned int seq;        // sequence number of the last reducing launch
    int spin;                // poll h_flag instead of cudaStreamSynchronize (one rank)
    cudaEvent_t ev0, ev1;
    int64_t launches;
    void* flush_buf;
    size_t flush_bytes;
    // optional per-kernel CUDA-event profile (bench.py roofline)
    int prof_on, prof_n;
    opkb_prof_rec* prof;
    double prof_ms[OPKB_K_NCLASS];
    int64_t prof_cnt[OPKB_K_NCLASS];
    // neighbour (halo) exchange of the stencil objective: its own stream and events,
    // created on first use (opkb_stencil.cu)
    cudaStream_t halo_stream;
    cudaEvent_t ev_ready, ev_halo;
    cudaEvent_t ev_user;     // opkb_context_wait_stream
    // freed vector blocks, handed out again to requests of the same size (opkb_core.cu)
    struct opkb_pool_block* pool;
    int pool_n, pool_cap;
    size_t pool_bytes;
};

int opkb_dev_alloc(opkb_context* ctx, size_t bytes, void** out);
void opkb_dev_free(opkb_context* ctx, void* ptr, size_t bytes);

void opkb_prof_begin(opkb_context* ctx, int cls);
void opkb_prof_end(opkb_context* ctx);

struct opkb_vector_ {
    opkb_context* ctx;
    int eltype;
    int64_t n;
    void* data;
    int owned;
};

// error plumbing (opkb_core.cu)
int opkb_set_error(int code, const char* fmt, ...);
int opkb_check_cuda(cudaError_t err, const char* what);
#define OPKB_CUDA(call)                                            \
    do {                                                           \
        int rc_ = opkb_check_cuda((call), #call);                  \
        if (rc_ != 0) return rc_;                                  \
    } while (0)

// Collect `nred` reduced scalars of this rank (already in ctx->d_res), combine
// them over the ranks in fixed rank order and return them in out[].  `kinds`
// gives OPKB_RSUM/RMAX/RMIN per slot (NULL = all sums).
int opkb_finish_reduction(opkb_context* ctx, int nred, const int* kinds, double* out);
int opkb_grid_for(const opkb_context* ctx, int64_t nwork_items, int ctas_per_sm);
// Chained launches (opkb_chain.cuh): one rank whose reducing kernels write their results straight
// into pinned host memory and are awaited by polling the completion flag.
int opkb_chain_supported(const opkb_context* ctx);
// Results of the reducing launch number `seq` (ReduceScratch::seq), written at opkb_res_ptr() +
// res_offset: waits until the flag has REACHED seq (later launches may already be in the stream).
int opkb_finish_reduction_seq(opkb_context* ctx, unsigned int seq, int res_offset, int nred, double* out);

void opkb_mailbox_close(opkb_context* ctx);
// NCCL (opkb_comm.cu, loaded with dlopen)
int opkb_nccl_allgather_f64(opkb_context* ctx, const double* send, double* recv, int count);
// rows to/from the previous and the next rank in one NCCL group on `stream` (NULL: nothing)
int opkb_nccl_halo(opkb_context* ctx, int eltype, size_t count, const void* send_prev, void* recv_prev,
                   const void* send_next, void* recv_next, cudaStream_t stream);

// ---------------------------------------------------------------------------
// 16-byte packs: 2 doubles or 4 floats per thread and request.
template <typename T> struct VecOf;
template <> struct VecOf<double> { static constexpr int N = 2; typedef double2 type; };
template <> struct VecOf<float>  { static constexpr int N = 4; typedef float4 type; };

template <typename T> struct Pack { T v[VecOf<T>::N]; };

template <typename T>
__device__ __forceinline__ Pack<T> ld_pack(const T* p, int64_t iv)
{
    typedef typename VecOf<T>::type V;
    union { V q; Pack<T> pk; } u;
    u.q = reinterpret_cast<const V*>(p)[iv];
    return u.pk;
}

template <typename T>
__device__ __forceinline__ void st_pack(T* p, int64_t iv, const Pack<T>& pk)
{
    typedef typename VecOf<T>::type V;
    union { V q; Pack<T> pk; } u;
    u.pk = pk;
    reinterpret_cast<V*>(p)[iv] = u.q;
}

template <typename T>
__device__ __forceinline__ Pack<T> splat_pack(T x)
{
    Pack<T> pk;
#pragma unroll
    for (int e = 0; e < VecOf<T>::N; ++e) pk.v[e] = x;
    return pk;
}

// A bound: an array (arr != NULL) or a scalar; `has` tells whether it counts
// (arrays always do, scalars only when finite: src/quasinewton.jl:248-269,
// src/bounds.jl:340-341).
template <typename T> struct Bound {
    const T* arr;
    T val;
    int has;
    __device__ __forceinline__ T at(int64_t i) const { return arr ? arr[i] : val; }
    __device__ __forceinline__ Pack<T> pack(int64_t iv) const
    {
        return arr ? ld_pack<T>(arr, iv) : splat_pack<T>(val);
    }
};

// fastclamp(x, lo, hi) = fastmax(fastmin(x, hi), lo): src/bounds.jl:41, :53, :65-67.
// NaN in x is kept; an infinite bound never triggers.
template <typename T>
__device__ __forceinline__ T fastclamp(T x, T lo, T hi)
{
    T t = (x > hi ? hi : x);
    return (t < lo ? lo : t);
}

// projdir / can_change: src/bounds.jl:306-311, :638-648.
template <typename T>
__device__ __forceinline__ T projdir(T x, T lo, T hi, int has_lo, int has_hi, int orient, T d)
{
    bool neg = (orient > 0 ? d < T(0) : d > T(0));  // is_negative(o, d)
    bool pos = (orient > 0 ? d > T(0) : d < T(0));  // is_positive(o, d)
    bool ok = (neg && (!has_lo || x > lo)) || (pos && (!has_hi || x < hi));
    return ok ? d : T(0);
}

// projdir with orientation -1 (the one vmlmb uses) from the two "can move" flags of
// x, flo = !has_lo || x > lo and fhi = !has_hi || x < hi: same result as projdir(),
// branch-free, and the flags are shared by the calls that concern the same x.
template <typename T>
__device__ __forceinline__ T projdir_m1(bool flo, bool fhi, T d)
{
    const bool ok = ((d > T(0)) & flo) | ((d < T(0)) & fhi);
    return ok ? d : T(0);
}

// step_limit_update() without the division when it cannot change smin or smax:
// with an = |bound - x| signed like the step and sdn = |sd|, a = an / sdn, and
// a > smax needs an > smax*sdn, a < smin needs an < smin*sdn; the filter leaves a
// margin of 2^-40 for the roundings, then the exact quotient decides as in the
// reference.  (Record-breaking quotients are rare: ~log(n) per thread.)
template <typename T>
__device__ __forceinline__ void step_limit_update_fast(T x, T lo, T hi, int has_lo, int has_hi, T sd,
                                                       T& smin, T& smax)
{
    const T inf = T(INFINITY);
    const bool up = sd > T(0), dn = sd < T(0);
    if (!(up | dn)) return;
    const bool has = up ? (has_hi != 0) : (has_lo != 0);
    if (!has) {
        smax = inf;
        return;
    }
    const T num = (up ? hi : lo) - x;
    const T an = up ? num : -num, sdn = up ? sd : -sd;
    const T m = sizeof(T) == 8 ? T(9.094947017729282e-13) : T(1e-4);
    if ((an > smax * sdn * (T(1) - m)) | ((an > T(0)) & (an < smin * sdn * (T(1) + m)))) {
        const T a = num / sd;
        if (T(0) < a && a < smin) smin = a;
        if (a > smax) smax = a;
    }
}

// one element of _step_limits: src/bounds.jl:456-506, :521-551, :566-596, :608-622.
// `sd` = s*d[i] with s = +-1; smin/smax in the element type.
template <typename T>
__device__ __forceinline__ void step_limit_update(T x, T lo, T hi, int has_lo, int has_hi, T sd,
                                                  T& smin, T& smax)
{
    const T inf = T(INFINITY);
    if (sd > T(0)) {
        if (has_hi) {
            T a = (hi - x) / sd;
            if (T(0) < a && a < smin) smin = a;
            if (a > smax) smax = a;
        } else {
            smax = inf;
        }
    } else if (sd < T(0)) {
        if (has_lo) {
            T a = (lo - x) / sd;
            if (T(0) < a && a < smin) smin = a;
            if (a > smax) smax = a;
        } else {
            smax = inf;
        }
    }
}

// exact product accumulated in double whatever T
template <typename T>
__device__ __forceinline__ double dacc(T a, T b, double acc)
{
    return fma((double)a, (double)b, acc);
}

// Rosenbrock pair, test/rosenbrock.jl:16-29 (arithmetic in T, no contraction).
template <typename T>
__device__ __forceinline__ void rosenbrock_pair(T x1, T x2, T& g1, T& g2, double& f1, double& f2)
{
    const T c1 = T(1), c2 = T(2), c10 = T(10), c200 = T(200);
    T t1 = c1 - x1;
    T u = x2 - x1 * x1;
    T t2 = c10 * u;
    g2 = c200 * u;
    g1 = -c2 * (x1 * g2 + t1);
    f1 += (double)(T)(t1 * t1);
    f2 += (double)(T)(t2 * t2);
}

// Quadratic element: r = x - c; g = a*r; term = g*r (f = sum(term)/2).
template <typename T>
__device__ __forceinline__ void quadratic_elem(T x, T a, T c, T& g, double& f)
{
    T r = x - c;
    g = a * r;
    f += (double)(T)(g * r);
}

// ---------------------------------------------------------------------------
// Deterministic reduction of NR scalars per thread: warp butterfly, fixed
// order over the warps of the CTA, per-CTA partial in global memory, and the
// last CTA (ticket) sums the partials of all CTAs in fixed order.
struct ReduceScratch {
    double* part;
    unsigned int* ticket;
    double* res;
    // completion flag in pinned host memory (one rank): the last CTA stores `seq` there after
    // its results (system-scope fence), and the host polls it instead of synchronising the stream
    unsigned int* flag;
    unsigned int seq;
};

// last thing the last CTA of a reducing kernel does (all its threads call it, after their
// writes to rs.res)
__device__ __forceinline__ void reduce_done(const ReduceScratch& rs)
{
    if (rs.flag) __threadfence_system();   // every writer's results are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {
        *rs.ticket = 0u;
        if (rs.flag) {
            // ... and the flag store is ordered after them by a fence of the storing thread itself
            // (release pattern: the barrier alone does not order thread 0's store against the other
            // threads' fenced writes at system scope)
            __threadfence_system();
            *(volatile unsigned int*)rs.flag = rs.seq;
        }
    }
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size) instead of once
// per launch (a microsecond or two of driver time per iteration of the small problems).  The
// attribute is per device; the table is shared by the host threads of the process (one context per
// thread and GPU in the single-process multi-GPU mode): a tiny spin lock guards it.
static inline cudaError_t opkb_ensure_smem(const void* kern, size_t smem)
{
    static const void* seen_k[256];
    static size_t seen_s[256];
    static int seen_d[256];
    static int nseen = 0;
    static volatile int lock = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    while (__atomic_exchange_n(&lock, 1, __ATOMIC_ACQUIRE)) { }
    for (int i = 0; i < nseen; ++i)
        if (seen_k[i] == kern && seen_d[i] == dev && seen_s[i] >= smem) {
            __atomic_store_n(&lock, 0, __ATOMIC_RELEASE);
            return cudaSuccess;
        }
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        int slot = -1;
        for (int i = 0; i < nseen; ++i) if (seen_k[i] == kern && seen_d[i] == dev) slot = i;
        if (slot < 0 && nseen < 256) slot = nseen++;
        if (slot >= 0) { seen_k[slot] = kern; seen_s[slot] = smem; seen_d[slot] = dev; }
    }
    __atomic_store_n(&lock, 0, __ATOMIC_RELEASE);
    return e;
}

// Every public entry point works on the device of its context, whatever device the host has made
// current in between (e.g. `torch.cuda.device(...)`): allocations and launches bind it first.
static inline void opkb_bind_device(const opkb_context* ctx)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != ctx->device) cudaSetDevice(ctx->device);
}

#define OPKB_MBOX_HDR 128                                              // flag | ready
#define OPKB_MBOX_SLOT (OPKB_MBOX_HDR + 2 * OPKB_NRED_MAX * sizeof(double))
// where the reducing launches of the current phase write their results (device address)
static inline double* opkb_res_ptr(opkb_context* ctx)
{
    if (ctx->mbox)
        return (double*)(ctx->mbox_dev + (size_t)ctx->rank * OPKB_MBOX_SLOT + OPKB_MBOX_HDR) +
               ((ctx->phase + 1) & 1ull) * OPKB_NRED_MAX;
    return ctx->d_res;
}
// the scratch of the next reducing launch (results at opkb_res_ptr(ctx) + res_offset)
static inline ReduceScratch opkb_make_scratch(opkb_context* ctx, int res_offset = 0)
{
    opkb_bind_device(ctx);
    ReduceScratch rs;
    rs.part = ctx->d_part;
    rs.ticket = ctx->d_ticket;
    rs.res = opkb_res_ptr(ctx) + res_offset;
    if (ctx->mbox) rs.flag = (unsigned int*)(ctx->mbox_dev + (size_t)ctx->rank * OPKB_MBOX_SLOT);
    else rs.flag = (ctx->spin && ctx->nranks <= 1 && ctx->d_res == ctx->h_res) ? ctx->h_flag : NULL;
    rs.seq = ++ctx->seq;
    return rs;
}

__device__ __forceinline__ double rcombine_dyn(int kind, double a, double b)
{
    if (kind == OPKB_RSUM) return a + b;
    if (kind == OPKB_RMAX) return (b > a ? b : a);
    return (b < a ? b : a);
}

__device__ __forceinline__ double rneutral_dyn(int kind)
{
    if (kind == OPKB_RSUM) return 0.0;
    if (kind == OPKB_RMAX) return -INFINITY;
    return INFINITY;
}

// kind of slot k from the two bit masks (max slots / min slots)
__host__ __device__ __forceinline__ int rkind(unsigned maxmask, unsigned minmask, int k)
{
    return ((maxmask >> k) & 1u) ? OPKB_RMAX : (((minmask >> k) & 1u) ? OPKB_RMIN : OPKB_RSUM);
}

// All threads of the CTA must call this (blockDim.x = OPKB_BLOCK or less, a
// multiple of 32).  On return of the LAST CTA, rs.res[0..NR) holds the result.
template <int NR, unsigned MAXM, unsigned MINM>
__device__ void block_reduce_publish(double (&v)[NR], const ReduceScratch& rs)
{
    __shared__ double sm[NR][32];
    __shared__ unsigned int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const int kind = rkind(MAXM, MINM, k);
        double a = v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double b = __shfl_xor_sync(0xffffffffu, a, off);
            a = rcombine_dyn(kind, a, b);
        }
        if (lane == 0) sm[k][warp] = a;
    }
    __syncthreads();
    if (threadIdx.x < NR) {
        const int k = threadIdx.x;
        const int kind = rkind(MAXM, MINM, k);
        double a = sm[k][0];
        for (int w = 1; w < nwarps; ++w) a = rcombine_dyn(kind, a, sm[k][w]);
        rs.part[(size_t)blockIdx.x * NR + k] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(rs.ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: for each slot, strided partial over the CTAs then the same tree.  All
    // slots are fetched before the first barrier (independent loads in flight together).
    double t[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) t[k] = rneutral_dyn(rkind(MAXM, MINM, k));
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NR; ++k)
            t[k] = rcombine_dyn(rkind(MAXM, MINM, k), t[k], __ldcg(&rs.part[(size_t)b * NR + k]));
    }
    __syncthreads();   // sm[][] of the first phase has been consumed by every thread
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const int kind = rkind(MAXM, MINM, k);
        double a = t[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double c = __shfl_xor_sync(0xffffffffu, a, off);
            a = rcombine_dyn(kind, a, c);
        }
        if (lane == 0) sm[k][warp] = a;
    }
    __syncthreads();
    if (threadIdx.x < NR) {
        const int k = threadIdx.x;
        const int kind = rkind(MAXM, MINM, k);
        double r = sm[k][0];
        for (int w = 1; w < nwarps; ++w) r = rcombine_dyn(kind, r, sm[k][w]);
        rs.res[k] = r;
    }
    reduce_done(rs);
}

// Last-CTA combination of per-CTA partials for kernels that reduce many slots
// (Gram sweep, fused direction kernel): res[k] = combine_b part[b * N + k].  Warp w
// takes the slots w, w + nwarps, ... two at a time; lane l combines the CTAs l, l + 32,
// ... in that order, then the butterfly: every load of a warp is in flight at once
// (one thread walking the CTAs serially costs ~500 cycles per CTA), fixed order.
template <typename KindOf>
__device__ __forceinline__ void last_cta_combine(const double* part, double* res, int N, unsigned nblk,
                                                 KindOf kind_of)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k0 = 2 * warp; k0 < N; k0 += 2 * nw) {
        const int k1 = (k0 + 1 < N) ? k0 + 1 : k0;
        const int kind0 = kind_of(k0), kind1 = kind_of(k1);
        double a0 = rneutral_dyn(kind0), a1 = rneutral_dyn(kind1);
        for (unsigned b = lane; b < nblk; b += 32) {
            const double v0 = __ldcg(&part[(size_t)b * N + k0]);
            const double v1 = __ldcg(&part[(size_t)b * N + k1]);
            a0 = rcombine_dyn(kind0, a0, v0);
            a1 = rcombine_dyn(kind1, a1, v1);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double c0 = __shfl_xor_sync(0xffffffffu, a0, off);
            const double c1 = __shfl_xor_sync(0xffffffffu, a1, off);
            a0 = rcombine_dyn(kind0, a0, c0);
            a1 = rcombine_dyn(kind1, a1, c1);
        }
        if (lane == 0) {
            res[k0] = a0;
            if (k1 != k0) res[k1] = a1;
        }
    }
}

// ---------------------------------------------------------------------------
// mbarrier / bulk-async-copy (TMA, SASS UBLKCP) primitives of the shared-memory
// rings of the Gram sweep (opkb_gram2.cuh) and of the direction kernel.
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// The same on the 32-bit shared-memory address of the barrier (taken once, outside the tile loop:
// the generic -> shared conversion and the polling loop were 14 % of the instructions of the fused
// direction kernel), with a suspend-time hint: the warp sleeps in hardware until the phase completes
// or the hint expires instead of spinning through the issue slots.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_A:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE_A;\n"
        "bra WAIT_LOOP_A;\n"
        "WAIT_DONE_A:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(0x989680u)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive_a(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

