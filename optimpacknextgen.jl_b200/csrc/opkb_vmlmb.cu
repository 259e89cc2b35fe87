// opkb_vmlmb.cu -- Tier 2: fused phases of `_vmlmb!` (src/quasinewton.jl:381-601)
// and of `apply_lbfgs!` (src/quasinewton.jl:703-779).  Every n-vector crosses
// HBM once per phase:
//   P1 trial      : x = prj(x0 - stp d); f, g = fg(x); p = projdir(g); 6 reductions
//   P3 gram       : in-place pair update + all inner products of the two-loop
//                   recursion in one tall-skinny sweep (opkb_gram2.cuh)
//   P4 direction  : d = H v0 from the host-solved coefficients, projection,
//                   g.d, |d|^2, step limits, save (x, g) for the next search
//   P4' steepest, P0 revert
#include <math.h>
#include <string.h>

#include "opkb_internal.cuh"
#include "opkb_chain.cuh"

struct opkb_vmlmb_ {
    opkb_context* ctx;
    int eltype;
    int64_t n;
    int mem;
    int method;
    const opkb_vector* lo;
    double lo_val;
    const opkb_vector* hi;
    double hi_val;
    int obj;
    const opkb_vector* qa;
    const opkb_vector* qc;
    opkb_vector *x, *g, *p, *d;
    opkb_vector* S[OPKB_SLOTS_MAX];
    opkb_vector* Y[OPKB_SLOTS_MAX];
    // The saves vcopy!(S[mark], x); vcopy!(Y[mark], g|p) (src/quasinewton.jl:562-563)
    // are done by aliasing: after a direction phase S[mark'] / Y[mark'] share the
    // buffers of x / g|p, and the buffers of the recycled slot become the targets
    // of the next trial point (no copy, no HBM traffic).
    void* spare_x;
    void* spare_y;
    int aliased_slot;      // -1: no aliasing active
    // Incremental Gram (opkb_direction2.cuh, CARRY).  `Gc` is the Gram block the last
    // opkb_vmlmb_gram call returned (for the pairs Gc_slots, oldest first).  After a
    // fused direction + trial launch, `carry_*` hold what that launch reduced for the
    // iteration the trial point would start; any other write to x, g, p or a pair
    // invalidates them.  carry_s / carry_y: two spare n-vectors that receive the new pair.
    double Gc[2 * OPKB_MEM_MAX * (OPKB_MEM_MAX + 1)];
    int Gc_m, Gc_valid;
    int Gc_slots[OPKB_MEM_MAX];
    opkb_vector* carry_s;
    opkb_vector* carry_y;
    int carry_valid, carry_m, carry_mb, carry_mark, carry_off;   // carry_m: pairs of the launch
    int carry_slots[OPKB_MEM_MAX];
    double carry_dense[4 * OPKB_MEM_MAX + 4];
    double carry_corr[OPKB_MEM_MAX * (OPKB_MEM_MAX + 1)];
    int64_t carry_hits, carry_misses, carry_skips;
    // VMLMB (method 2): the fused direction + trial launch does NOT write p' (the next
    // launch recomputes it from x, g, lo, hi and nothing else reads it in steady state);
    // p_stale != 0 means p does not belong to the current (x, g) and is recomputed on
    // demand (vmlmb_ensure_p) by whoever reads it: one pass less per iteration.
    int p_stale;
    // packed free-set mask of the sequential two-loop recursion (k_twoloop_step): valid from
    // the first step of a recursion until anything may change p
    void* maskb;
    int maskb_valid, maskb_failed;
    // user objective: the direction kernel already wrote x' = project(x - d) into the buffer
    // that becomes x when the next line search starts (slot xspec_mark); a trial_begin with
    // stp = 1 for that slot is then a pointer swap.  -1: nothing speculated.
    int xspec_mark;
    // switches read ONCE, when the workspace is created (every rank of a job sees the same
    // environment: the decisions they drive are then identical on all ranks)
    int opt_carry;      // OPKB_CARRY: -1 default (Float64; Float32 up to 8 pairs), 0 off, 2 always
    int opt_lazy_p;     // OPKB_LAZY_P (default 1)
    int opt_xspec;      // OPKB_XSPEC (default 1)
    int opt_maskb;      // OPKB_MASKB (default 1)
    int opt_carry_check;
    int opt_history;    // OPKB_HISTORY: 1 default (Float64, mem <= 8), 0 off, 2 always
    // Representation of the pairs on the Gram path.  hist = 1: slot k keeps the ITERATE (x_k, g_k)
    // saved when line search k started (src/quasinewton.jl:562-563) and is never turned into the
    // difference of :512-513: pair k is S[k] - S[succ(k)] (succ: the slot saved next, (x, g) for
    // the newest pair), formed on the fly by the ring kernels -- the fused kernel then writes no
    // s', y' (2m + 9 passes instead of 2m + 11).  hist = 0: the reference's stored differences
    // (sequential two-loop, BLMVM, OPKB_HISTORY=0).  -1: not decided yet (no pair formed).
    int hist;
    // Chained launches (opkb_chain.cuh): the fused launch of the NEXT iteration, enqueued before the
    // results of the current one are known.  valid: it is in the stream; go: what the current
    // launch's last CTA decided for it (-1 until its results have been collected; a launch with
    // go = 0 is a no-op and is forgotten at once, so outside direction_run valid implies go = 1).
    struct {
        int valid, go;
        int fuse, mb, nd, ncorr, m, mark_next, ran_lazy_p, resoff;
        unsigned int seq;
        int slots[OPKB_CHAIN_MB];
        const void* x;                    // the (x, g) it reads: the workspace's when it is adopted
        const void* g;
        double rec[OPKB_CHAIN_NREC];      // decision record of its predecessor: the coefficients it runs with
    } pend;
    int64_t chain_adopted, chain_heads, chain_nogo;
    // OPKB_CHAIN_CHECK=1 (development aid): the decision record of the last chained launch, compared
    // with what the host asks for next when nothing else happened in between
    int opt_chain_check, opt_chain_dbg, last_rec_valid, last_rec_m;
    double last_rec[OPKB_CHAIN_NREC];
    int64_t chain_checked, chain_check_failed;
};

// A chained launch that runs (go = 1) and is not adopted by the host would have overwritten d and
// the buffers of the recycled slot behind the host's back: every other entry point that touches
// the workspace checks that none is outstanding.  Cannot happen unless chain_decide() and the
// host's stage machine disagree -- an internal error, reported loudly.
static int chain_guard(opkb_vmlmb* ws)
{
    ws->last_rec_valid = 0;
    if (!ws->pend.valid) return 0;
    ws->pend.valid = 0;
    cudaStreamSynchronize(ws->ctx->stream);
    return opkb_set_error(OPKB_EINVAL, "internal error: a chained launch (device go = %d) was not adopted by the host",
                          ws->pend.go);
}

static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int vmlmb_ensure_p(opkb_vmlmb* ws)
{
    if (!ws->p_stale || !ws->p) return 0;
    ws->p_stale = 0;
    // project_direction!(p, x, lo, hi, -, g) src/quasinewton.jl:396: the bits the kernel
    // would have written
    return opkb_project_direction(ws->ctx, ws->p, ws->x, ws->lo, ws->lo_val, ws->hi, ws->hi_val,
                                  OPKB_BACKWARD, ws->g);
}

// x (and g|p) get their own buffers back before they are written again
static void vmlmb_unalias(opkb_vmlmb* ws)
{
    if (ws->aliased_slot < 0) return;
    ws->x->data = ws->spare_x;
    (ws->method == 1 ? ws->p : ws->g)->data = ws->spare_y;
    ws->aliased_slot = -1;
}

static void vmlmb_alias(opkb_vmlmb* ws, int slot)
{
    if (ws->aliased_slot == slot) return;   // second direction of the same iteration (:542-556)
    vmlmb_unalias(ws);
    opkb_vector* ysrc = (ws->method == 1 ? ws->p : ws->g);
    ws->spare_x = ws->S[slot]->data;
    ws->spare_y = ws->Y[slot]->data;
    ws->S[slot]->data = ws->x->data;
    ws->Y[slot]->data = ysrc->data;
    ws->aliased_slot = slot;
}

#define LAUNCHB(ctx, cls, kernel, grid, block, smem, ...)                    \
    do {                                                                     \
        opkb_prof_begin((ctx), (cls));                                       \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);     \
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

// ---------------------------------------------------------------------------
// workspace
extern "C" int opkb_vmlmb_create(opkb_context* ctx, int eltype, int64_t n, int mem, int method,
                                 const opkb_vector* lo, double lo_val, const opkb_vector* hi,
                                 double hi_val, opkb_vmlmb** out)
{
    if (!ctx || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    *out = NULL;
    if (eltype != OPKB_FLOAT && eltype != OPKB_DOUBLE)
        return opkb_set_error(OPKB_EINVAL, "invalid element type");
    if (mem < 1 || mem > OPKB_SLOTS_MAX)
        return opkb_set_error(OPKB_EINVAL, "mem must be in 1..%d", OPKB_SLOTS_MAX);
    if (method < 0 || method > 2) return opkb_set_error(OPKB_EINVAL, "invalid method");
    if ((lo && (lo->n != n || lo->eltype != eltype)) || (hi && (hi->n != n || hi->eltype != eltype)))
        return opkb_set_error(OPKB_EINVAL, "bounds of a different space");
    opkb_vmlmb* ws = (opkb_vmlmb*)calloc(1, sizeof(opkb_vmlmb));
    if (!ws) return opkb_set_error(OPKB_ENOMEM, "out of host memory");
    ws->ctx = ctx;
    ws->xspec_mark = -1;
    ws->eltype = eltype;
    ws->n = n;
    ws->mem = mem;
    ws->method = method;
    ws->lo = lo;
    ws->lo_val = lo_val;
    ws->hi = hi;
    ws->hi_val = hi_val;
    ws->obj = OPKB_OBJ_USER;
    ws->aliased_slot = -1;
    ws->opt_carry = env_int("OPKB_CARRY", -1);
    ws->opt_lazy_p = env_int("OPKB_LAZY_P", 1);
    ws->opt_xspec = env_int("OPKB_XSPEC", 1);
    ws->opt_maskb = env_int("OPKB_MASKB", 1);
    ws->opt_carry_check = getenv("OPKB_CARRY_CHECK") ? 1 : 0;
    ws->opt_history = env_int("OPKB_HISTORY", 1);
    ws->opt_chain_check = env_int("OPKB_CHAIN_CHECK", 0);
    ws->opt_chain_dbg = env_int("OPKB_CHAIN_DBG", 0);
    // The history saves two passes whatever m, forming the differences costs 2m subtractions and
    // shared-memory stores per element: measured with the SPEC instances on a power-capped box, it wins
    // 5 % at m = 5 (C2: 3417 -> 3588 it/s) and loses 1 % at m = 10 (C4: 96.0 -> 94.9 it/s); the Float32
    // instances, issue-bound, always lose (C5 shard: 145 -> 132 it/s).  Default: Float64 with up to 8
    // pairs; OPKB_HISTORY=2 forces it, =0 disables it.
    ws->hist = (method == 1 || !ws->opt_history ||
                (ws->opt_history != 2 && (eltype == OPKB_FLOAT || mem > 8))) ? 0 : -1;
    int rc = 0;
    rc = rc ? rc : opkb_vector_create(ctx, eltype, n, &ws->x);
    rc = rc ? rc : opkb_vector_create(ctx, eltype, n, &ws->g);
    if (method > 0) rc = rc ? rc : opkb_vector_create(ctx, eltype, n, &ws->p);
    rc = rc ? rc : opkb_vector_create(ctx, eltype, n, &ws->d);
    for (int k = 0; k < mem && !rc; ++k) {
        rc = rc ? rc : opkb_vector_create(ctx, eltype, n, &ws->S[k]);
        rc = rc ? rc : opkb_vector_create(ctx, eltype, n, &ws->Y[k]);
    }
    // the two spare buffers of the incremental Gram (the fused direction kernel writes the pair
    // the speculated trial point would form): allocated HERE, not lazily at the first launch that
    // wants them, so that whether a rank carries never depends on a rank-local allocation
    // Float64: always (up to 12 pairs); Float32: up to 8 pairs, where the SPLIT / 11-warp instances
    // run (the 7-warp Float32 instances of 9..12 pairs are issue-bound: only with OPKB_CARRY=2)
    // (with the iterate history the new pair is never materialised: no spare buffers)
    if (!rc && ws->hist == 0 && method != 1 && mem <= 12 && ws->opt_carry != 0 &&
        (eltype == OPKB_DOUBLE || ws->opt_carry == 2 || mem <= 8)) {
        rc = rc ? rc : opkb_vector_create(ctx, eltype, n, &ws->carry_s);
        rc = rc ? rc : opkb_vector_create(ctx, eltype, n, &ws->carry_y);
    }
    if (rc) {
        opkb_vmlmb_destroy(ws);
        return rc;
    }
    *out = ws;
    return 0;
}

extern "C" int opkb_vmlmb_destroy(opkb_vmlmb* ws)
{
    if (!ws) return 0;
    if (ws->pend.valid) cudaStreamSynchronize(ws->ctx->stream);   // a launch enqueued ahead may still run
    if (ws->x && ws->g) vmlmb_unalias(ws);
    opkb_vector_destroy(ws->x);
    opkb_vector_destroy(ws->g);
    opkb_vector_destroy(ws->p);
    if (ws->maskb) cudaFree(ws->maskb);
    opkb_vector_destroy(ws->d);
    opkb_vector_destroy(ws->carry_s);
    opkb_vector_destroy(ws->carry_y);
    for (int k = 0; k < OPKB_SLOTS_MAX; ++k) {
        opkb_vector_destroy(ws->S[k]);
        opkb_vector_destroy(ws->Y[k]);
    }
    free(ws);
    return 0;
}

extern "C" int opkb_vmlmb_set_objective(opkb_vmlmb* ws, int kind, const opkb_vector* a,
                                        const opkb_vector* c)
{
    if (!ws) return opkb_set_error(OPKB_EINVAL, "NULL workspace");
    if (kind == OPKB_OBJ_QUADRATIC) {
        if (!a || !c || a->n != ws->n || c->n != ws->n || a->eltype != ws->eltype ||
            c->eltype != ws->eltype)
            return opkb_set_error(OPKB_EINVAL, "quadratic objective needs a, c of the same space");
    } else if (kind == OPKB_OBJ_ROSENBROCK) {
        if (ws->n % 2)
            return opkb_set_error(OPKB_EINVAL, "Rosenbrock needs an even number of (local) variables");
    } else if (kind != OPKB_OBJ_USER) {
        return opkb_set_error(OPKB_EINVAL, "unknown objective");
    }
    ws->obj = kind;
    ws->qa = a;
    ws->qc = c;
    return 0;
}

// used by the native driver (opkb_driver.cu)
int opkb_vmlmb_info(const opkb_vmlmb* ws, int* mem, int* method, int* eltype, int* obj)
{
    *mem = ws->mem;
    *method = ws->method;
    *eltype = ws->eltype;
    *obj = ws->obj;
    return 0;
}

extern "C" opkb_vector* opkb_vmlmb_vector(opkb_vmlmb* ws, int which, int slot)
{
    if (!ws) return NULL;
    switch (which) {
    case 'x': return ws->x;
    case 'g': return ws->g;
    case 'p':
        ws->maskb_valid = 0;   // the caller may write to it
        return (vmlmb_ensure_p(ws) == 0) ? ws->p : NULL;
    case 'd': return ws->d;
    case 'S': return (slot >= 0 && slot < ws->mem) ? ws->S[slot] : NULL;
    case 'Y': return (slot >= 0 && slot < ws->mem) ? ws->Y[slot] : NULL;
    }
    return NULL;
}

// ---------------------------------------------------------------------------
// P1: trial point, objective, projected gradient, reductions.
// STAGE 0: everything (built-in objective); 1: x only (user objective, begin);
// 2: p + reductions from the caller's g (user objective, end).
template <typename T> struct TrialArgs {
    const T* x0;   // S[mark]
    const T* d;
    T* x;
    T* g;
    T* p;          // NULL for L-BFGS
    const T* qa;
    const T* qc;
    Bound<T> lo, hi;
    T mstp;        // -stp converted to the element type
    int64_t n;
    int initial;
    int bounded;
    // STAGE 2 (user objective): this rank's share of a separable objective, published in slot 0
    // so that the ranks' values are summed in rank order with the other partials
    double f_partial;
};

template <typename T>
__device__ __forceinline__ T trial_point(const TrialArgs<T>& A, T xin, T x0, T d, T lo, T hi)
{
    // vcombine!(x, 1, S[mark], -stp, d) :599 then project_variables! :386
    T xi = A.initial ? xin : x0 + A.mstp * d;
    if (A.bounded) xi = fastclamp(xi, lo, hi);
    return xi;
}

template <typename T>
__device__ __forceinline__ void trial_reduce(const TrialArgs<T>& A, T xi, T gi, T x0, T d, T lo, T hi,
                                             T& pi, double (&acc)[8])
{
    // project_direction!(p, x, lo, hi, -, g) :396
    pi = A.bounded ? projdir(xi, lo, hi, A.lo.has, A.hi.has, -1, gi) : gi;
    acc[2] = dacc(pi, pi, acc[2]);              // |p|^2 :399
    acc[7] += (pi != T(0)) ? 1.0 : 0.0;         // |sel| (bounds.jl:701-714)
    if (!A.initial) {
        T si = xi - x0;                         // vcombine!(s, 1, x, -1, S[mark]) :427
        acc[3] = dacc(gi, si, acc[3]);          // vdot(g, s) :432
        acc[4] = dacc(gi, d, acc[4]);           // vdot(g, d) :434
        acc[5] = dacc(xi, xi, acc[5]);          // vnorm2(x) :446
        acc[6] = dacc(si, si, acc[6]);          // vnorm2(s) :448
    } else {
        acc[5] = dacc(xi, xi, acc[5]);
    }
}

template <typename T, int OBJ, int STAGE>
__global__ void __launch_bounds__(OPKB_BLOCK) k_trial(TrialArgs<T> A, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = A.n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // acc: 0,1 objective partial sums; 2 |p|^2; 3 g.s; 4 g.d; 5 |x|^2; 6 |s|^2; 7 nfree
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> px0, pd, pxin, pl, ph, px, pg, pp, pa, pc;
        if (A.initial) {
            pxin = ld_pack<T>(A.x, iv);
            px0 = pxin;
            pd = splat_pack<T>(T(0));
        } else {
            px0 = ld_pack<T>(A.x0, iv);
            pd = ld_pack<T>(A.d, iv);
            pxin = px0;
        }
        if (A.bounded) {
            pl = A.lo.pack(iv);
            ph = A.hi.pack(iv);
        } else {
            pl = splat_pack<T>(T(0));
            ph = pl;
        }
        if (STAGE == 2) {
            px = ld_pack<T>(A.x, iv);
            pg = ld_pack<T>(A.g, iv);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                px.v[e] = trial_point(A, pxin.v[e], px0.v[e], pd.v[e], pl.v[e], ph.v[e]);
            st_pack<T>(A.x, iv, px);
        }
        if (STAGE == 0) {
            if (OBJ == OPKB_OBJ_ROSENBROCK) {
#pragma unroll
                for (int e = 0; e < VEC; e += 2)
                    rosenbrock_pair(px.v[e], px.v[e + 1], pg.v[e], pg.v[e + 1], acc[0], acc[1]);
            } else {
                pa = ld_pack<T>(A.qa, iv);
                pc = ld_pack<T>(A.qc, iv);
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    quadratic_elem(px.v[e], pa.v[e], pc.v[e], pg.v[e], acc[0]);
            }
            st_pack<T>(A.g, iv, pg);
        }
        if (STAGE != 1) {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                trial_reduce(A, px.v[e], pg.v[e], px0.v[e], pd.v[e], pl.v[e], ph.v[e], pp.v[e], acc);
            if (A.p) st_pack<T>(A.p, iv, pp);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // scalar tail (at most VEC-1 elements; an even count for Rosenbrock)
        for (int64_t i = nvec * VEC; i < A.n; ++i) {
            T x0 = A.initial ? A.x[i] : A.x0[i];
            T d = A.initial ? T(0) : A.d[i];
            T lo = A.bounded ? A.lo.at(i) : T(0), hi = A.bounded ? A.hi.at(i) : T(0);
            if (STAGE != 2) A.x[i] = trial_point(A, A.x[i], x0, d, lo, hi);
        }
        if (STAGE == 0) {
            if (OBJ == OPKB_OBJ_ROSENBROCK) {
                for (int64_t i = nvec * VEC; i + 1 < A.n; i += 2)
                    rosenbrock_pair(A.x[i], A.x[i + 1], A.g[i], A.g[i + 1], acc[0], acc[1]);
            } else {
                for (int64_t i = nvec * VEC; i < A.n; ++i)
                    quadratic_elem(A.x[i], A.qa[i], A.qc[i], A.g[i], acc[0]);
            }
        }
        if (STAGE != 1) {
            for (int64_t i = nvec * VEC; i < A.n; ++i) {
                T x0 = A.initial ? A.x[i] : A.x0[i];
                T d = A.initial ? T(0) : A.d[i];
                T lo = A.bounded ? A.lo.at(i) : T(0), hi = A.bounded ? A.hi.at(i) : T(0);
                T pi;
                trial_reduce(A, A.x[i], A.g[i], x0, d, lo, hi, pi, acc);
                if (A.p) A.p[i] = pi;
            }
        }
    }
    if (STAGE == 2 && blockIdx.x == 0 && threadIdx.x == 0) acc[0] = A.f_partial;
    if (STAGE != 1) block_reduce_publish<8, 0u, 0u>(acc, rs);
}

template <typename T>
static int trial_launch(opkb_vmlmb* ws, int mark, double stp, int initial, int stage, double f_partial = 0.0)
{
    {
        int rcg = chain_guard(ws);
        if (rcg) return rcg;
    }
    opkb_context* ctx = ws->ctx;
    ws->carry_valid = 0;                 // another point than the speculated one
    if (stage != 2) vmlmb_unalias(ws);   // x, g|p are about to be overwritten
    if (stage == 1 && !initial && ws->xspec_mark == mark && stp == 1.0) {
        ws->xspec_mark = -1;             // x' is already there (written by the direction kernel)
        ws->maskb_valid = 0;
        return 0;
    }
    if (stage != 2) ws->xspec_mark = -1;
    if (stage != 1) ws->p_stale = 0;     // this launch writes p
    ws->maskb_valid = 0;
    TrialArgs<T> A;
    A.x0 = (const T*)ws->S[mark]->data;
    A.d = (const T*)ws->d->data;
    A.x = (T*)ws->x->data;
    A.g = (T*)ws->g->data;
    A.p = ws->p ? (T*)ws->p->data : NULL;
    A.qa = ws->qa ? (const T*)ws->qa->data : NULL;
    A.qc = ws->qc ? (const T*)ws->qc->data : NULL;
    A.lo = make_bound<T>(ws->lo, ws->lo_val, true);
    A.hi = make_bound<T>(ws->hi, ws->hi_val, false);
    A.mstp = (T)(-stp);
    A.n = ws->n;
    A.initial = initial;
    A.bounded = ws->method > 0;
    A.f_partial = f_partial;
    int grid = opkb_grid_for(ctx, ws->n / VecOf<T>::N, 4);
    ReduceScratch rs = scratch(ctx);
    if (stage == 1)
        LAUNCHB(ctx, OPKB_K_TRIAL, (k_trial<T, OPKB_OBJ_USER, 1>), grid, OPKB_BLOCK, 0, A, rs);
    else if (stage == 2)
        LAUNCHB(ctx, OPKB_K_TRIAL, (k_trial<T, OPKB_OBJ_USER, 2>), grid, OPKB_BLOCK, 0, A, rs);
    else if (ws->obj == OPKB_OBJ_ROSENBROCK)
        LAUNCHB(ctx, OPKB_K_TRIAL, (k_trial<T, OPKB_OBJ_ROSENBROCK, 0>), grid, OPKB_BLOCK, 0, A, rs);
    else
        LAUNCHB(ctx, OPKB_K_TRIAL, (k_trial<T, OPKB_OBJ_QUADRATIC, 0>), grid, OPKB_BLOCK, 0, A, rs);
    return 0;
}

static int trial_collect(opkb_vmlmb* ws, double f_user, int user, double out[8])
{
    double r[8];
    int rc = opkb_finish_reduction(ws->ctx, 8, NULL, r);
    if (rc) return rc;
    // user: 1 = f_user is the global value; 2 = the ranks' shares were summed in slot 0
    double f = user ? (user == 2 ? r[0] : f_user)
                    : (ws->obj == OPKB_OBJ_ROSENBROCK ? r[0] + r[1] : 0.5 * r[0]);
    out[0] = f;
    out[1] = r[2];
    out[2] = r[3];
    out[3] = r[4];
    out[4] = r[5];
    out[5] = r[6];
    out[6] = r[7];
    out[7] = 0;
    return 0;
}

static int check_mark(opkb_vmlmb* ws, int mark)
{
    if (!ws) return opkb_set_error(OPKB_EINVAL, "NULL workspace");
    if (mark < 0 || mark >= ws->mem) return opkb_set_error(OPKB_EINVAL, "mark out of range");
    return 0;
}

extern "C" int opkb_vmlmb_trial(opkb_vmlmb* ws, int mark, double stp, int initial, double out[8])
{
    int rc = check_mark(ws, mark);
    if (rc) return rc;
    if (ws->obj == OPKB_OBJ_USER)
        return opkb_set_error(OPKB_EINVAL, "no built-in objective set: use trial_begin/trial_end");
    rc = (ws->eltype == OPKB_DOUBLE) ? trial_launch<double>(ws, mark, stp, initial, 0)
                                     : trial_launch<float>(ws, mark, stp, initial, 0);
    if (rc) return rc;
    return trial_collect(ws, 0.0, 0, out);
}

extern "C" int opkb_vmlmb_trial_begin(opkb_vmlmb* ws, int mark, double stp, int initial)
{
    int rc = check_mark(ws, mark);
    if (rc) return rc;
    if (!(initial && ws->method == 0)) {   // (x unchanged otherwise)
        rc = (ws->eltype == OPKB_DOUBLE) ? trial_launch<double>(ws, mark, stp, initial, 1)
                                         : trial_launch<float>(ws, mark, stp, initial, 1);
        if (rc) return rc;
    }
    // A user objective may run on a stream of its own (CUDA.jl, torch, ...): x is complete in
    // device memory when this call returns (include/opkb200.h, "raw device pointers").
    if (ws->obj == OPKB_OBJ_USER) OPKB_CUDA(cudaStreamSynchronize(ws->ctx->stream));
    return 0;
}

// f_mode 1: `f` is the global value; 2: this rank's share of a separable objective
int opkb_vmlmb_trial_end_mode(opkb_vmlmb* ws, int mark, double f, int initial, int f_mode, double out[8])
{
    int rc = check_mark(ws, mark);
    if (rc) return rc;
    rc = (ws->eltype == OPKB_DOUBLE) ? trial_launch<double>(ws, mark, 0.0, initial, 2, f)
                                     : trial_launch<float>(ws, mark, 0.0, initial, 2, f);
    if (rc) return rc;
    return trial_collect(ws, f, f_mode, out);
}

extern "C" int opkb_vmlmb_trial_end(opkb_vmlmb* ws, int mark, double f, int initial, double out[8])
{
    return opkb_vmlmb_trial_end_mode(ws, mark, f, initial, 1, out);
}

// ---------------------------------------------------------------------------
// P3: Gram sweep
#include "opkb_gram2.cuh"
#include "opkb_direction2.cuh"
#include <stdlib.h>

// The Gram block of the pairs `slots` (the newest one still to be formed from the
// accepted trial point) assembled from the previous block and what the fused
// direction + trial launch reduced (Dir2Carry): no kernel, no HBM traffic.
static int gram_from_carry(opkb_vmlmb* ws, int mark, int m, const int* slots, double* G)
{
    const int mo = ws->carry_m;                 // pairs the launch worked on
    const int drop = (mo == ws->mem) ? 1 : 0;   // the oldest pair was recycled for the trial point
    const int MB = ws->carry_mb;
    const int W = m + 1, Wo = mo + 1;
    const int half = MB * (MB + 1) / 2;
    // index of entry (a <= b) inside one triangular half (Dir2Carry enumeration)
    auto tri = [MB](int a, int b) { return a * MB - a * (a - 1) / 2 + (b - a); };
    for (int i = 0; i < 2 * m * W; ++i) G[i] = 0.0;
    const double* D = ws->carry_dense;
    const double* Cq = ws->carry_corr;
    const double* Go = ws->Gc;
    const int nn = m - 1;
    for (int a2 = 0; a2 < nn; ++a2) {
        const int a = a2 + drop;
        G[(2 * a2) * W] = D[4 * a + 0];             // s_a . v0'
        G[(2 * a2 + 1) * W] = D[4 * a + 1];         // y_a . v0'
        for (int b2 = a2; b2 < nn; ++b2) {
            const int b = b2 + drop;
            G[(2 * a2) * W + 1 + b2] = Go[(2 * a) * Wo + 1 + b] + Cq[tri(a, b)];
            G[(2 * a2 + 1) * W + 1 + b2] = Go[(2 * a + 1) * Wo + 1 + b] + Cq[half + tri(a, b)];
        }
        G[(2 * a2) * W + 1 + nn] = D[4 * a + 2];    // s_a . y'
        G[(2 * a2 + 1) * W + 1 + nn] = D[4 * a + 3];
    }
    G[(2 * nn) * W] = D[4 * MB + 0];
    G[(2 * nn + 1) * W] = D[4 * MB + 1];
    G[(2 * nn) * W + 1 + nn] = D[4 * MB + 2];
    G[(2 * nn + 1) * W + 1 + nn] = D[4 * MB + 3];
    if (ws->hist != 1) {
        // the pair itself was written to the spare buffers: swap them in (the buffers
        // that held x0, g0 become the spares of the next launch)
        void* t = ws->S[mark]->data;
        ws->S[mark]->data = ws->carry_s->data;
        ws->carry_s->data = t;
        t = ws->Y[mark]->data;
        ws->Y[mark]->data = ws->carry_y->data;
        ws->carry_y->data = t;
    }   // (iterate history: slot `mark` keeps (x0, g0); the pair is x0 - x, g0 - g on the fly)
    ws->carry_valid = 0;
    ws->carry_hits += 1;
    return 0;
}

extern "C" int opkb_vmlmb_gram(opkb_vmlmb* ws, int mark, int m, const int* slots, double* G)
{
    int rc = check_mark(ws, mark);
    if (rc) return rc;
    if (m < 1 || m > ws->mem || !slots || !G) return opkb_set_error(OPKB_EINVAL, "invalid m/slots/G");
    if (m > OPKB_MEM_MAX)
        return opkb_set_error(OPKB_EINVAL, "the Gram-form two-loop supports m <= %d pairs: use the sequential "
                              "form (opkb_vmlmb_twoloop_step) beyond", OPKB_MEM_MAX);
    if (slots[m - 1] != mark)
        return opkb_set_error(OPKB_EINVAL, "the newest pair (slots[m-1]) must be `mark`");
    for (int j = 0; j < m; ++j)
        if (slots[j] < 0 || slots[j] >= ws->mem) return opkb_set_error(OPKB_EINVAL, "slot out of range");
    if (ws->hist < 0) ws->hist = 1;     // the Gram path keeps the iterate history (see opkb_vmlmb_)
    bool carried = false;
    if (ws->carry_valid && ws->Gc_valid && mark == ws->carry_mark && m >= 2) {
        // the caller took the speculated trial point (nothing else touched the
        // workspace since) and asks for the pairs that launch foresaw
        const int mo = ws->carry_m, drop = (mo == ws->mem) ? 1 : 0;
        bool same = (m == mo + 1 - drop);
        for (int j = 0; same && j < m - 1; ++j) same = (slots[j] == ws->carry_slots[j + drop]);
        if (same) {
            rc = gram_from_carry(ws, mark, m, slots, G);
            if (rc) return rc;
            carried = true;
            if (ws->opt_carry_check && ws->hist != 1 && vmlmb_ensure_p(ws) == 0) {
                // development aid: the same block by a sweep over the same state (the
                // pair is already formed: x = g = 0 makes the in-place update the identity)
                opkb_vector* z = NULL;
                static thread_local double G2[2 * OPKB_MEM_MAX * (OPKB_MEM_MAX + 1)];
                if (opkb_vector_create(ws->ctx, ws->eltype, ws->n, &z) == 0) {
                    cudaMemsetAsync(z->data, 0, (size_t)ws->n * (ws->eltype == OPKB_DOUBLE ? 8 : 4), ws->ctx->stream);
                    const void* Sp[OPKB_MEM_MAX];
                    const void* Yp[OPKB_MEM_MAX];
                    for (int j = 0; j < m; ++j) {
                        Sp[j] = ws->S[slots[j]]->data;
                        Yp[j] = ws->Y[slots[j]]->data;
                    }
                    const void* v0 = (ws->method == 2 ? ws->p->data : ws->g->data);
                    if (ws->eltype == OPKB_DOUBLE)
                        gram2_run<double>(ws->ctx, ws->n, m, Sp, Yp, v0, z->data, z->data, ws->method == 2, G2);
                    else
                        gram2_run<float>(ws->ctx, ws->n, m, Sp, Yp, v0, z->data, z->data, ws->method == 2, G2);
                    opkb_vector_destroy(z);
                    double worst = 0, scale = 0;
                    int wr = -1, wc = -1;
                    const int W = m + 1;
                    for (int i = 0; i < 2 * m * W; ++i) scale = fmax(scale, fabs(G2[i]));
                    for (int r = 0; r < 2 * m; ++r)
                        for (int c = 0; c < W; ++c) {
                            if (c > 0 && (c - 1) < r / 2) continue;
                            const double dv = fabs(G[r * W + c] - G2[r * W + c]) / fmax(fabs(G2[r * W + c]), 1e-300);
                            if (dv > worst) { worst = dv; wr = r; wc = c; }
                        }
                    fprintf(stderr, "[carry check] hit %lld m %d worst rel %.3e at (%d,%d): carried %.17g sweep %.17g (max |G| %.3e)\n",
                            (long long)ws->carry_hits, m, worst, wr, wc, wr >= 0 ? G[wr * W + wc] : 0.0,
                            wr >= 0 ? G2[wr * W + wc] : 0.0, scale);
                }
            }
        }
    }
    if (!carried) {
        rc = chain_guard(ws);
        if (rc) return rc;
        if (ws->carry_valid) ws->carry_misses += 1;
        ws->carry_valid = 0;
        rc = vmlmb_ensure_p(ws);
        if (rc) return rc;
        const void* Sp[OPKB_MEM_MAX];
        const void* Yp[OPKB_MEM_MAX];
        for (int j = 0; j < m; ++j) {
            Sp[j] = ws->S[slots[j]]->data;
            Yp[j] = ws->Y[slots[j]]->data;
        }
        // v0 = p for VMLMB (mask p != 0), g otherwise; the gradient difference uses
        // p for BLMVM (src/quasinewton.jl:513)
        const void* v0 = (ws->method == 2 ? ws->p->data : ws->g->data);
        const void* gcur = (ws->method == 1 ? ws->p->data : ws->g->data);
        if (ws->eltype == OPKB_DOUBLE)
            rc = gram2_run<double>(ws->ctx, ws->n, m, Sp, Yp, v0, ws->x->data, gcur, ws->method == 2, G, ws->hist == 1);
        else
            rc = gram2_run<float>(ws->ctx, ws->n, m, Sp, Yp, v0, ws->x->data, gcur, ws->method == 2, G, ws->hist == 1);
        if (rc) {
            ws->Gc_valid = 0;
            return rc;
        }
    }
    memcpy(ws->Gc, G, sizeof(double) * 2 * m * (m + 1));
    ws->Gc_m = m;
    for (int j = 0; j < m; ++j) ws->Gc_slots[j] = slots[j];
    ws->Gc_valid = 1;
    return 0;
}

extern "C" int opkb_vmlmb_last_gram(const opkb_vmlmb* ws, int* m, int* slots, double* G)
{
    if (!ws || !m) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    *m = ws->Gc_valid ? ws->Gc_m : 0;
    if (!ws->Gc_valid) return 0;
    if (slots) for (int j = 0; j < ws->Gc_m; ++j) slots[j] = ws->Gc_slots[j];
    if (G) memcpy(G, ws->Gc, sizeof(double) * 2 * ws->Gc_m * (ws->Gc_m + 1));
    return 0;
}

extern "C" int opkb_vmlmb_history(const opkb_vmlmb* ws) { return ws ? ws->hist : -1; }

extern "C" int opkb_vmlmb_carry_stats(const opkb_vmlmb* ws, int64_t* hits, int64_t* misses)
{
    if (!ws) return opkb_set_error(OPKB_EINVAL, "NULL workspace");
    if (hits) *hits = ws->carry_hits;
    if (misses) *misses = ws->carry_misses;
    return 0;
}

// ---------------------------------------------------------------------------
// P4: direction assembly + projection + reductions + save
template <typename T> struct DirArgs {
    const T* S[OPKB_MEM_MAX];
    const T* Y[OPKB_MEM_MAX];
    T ma[OPKB_MEM_MAX];   // -alpha_j in the element type
    T cc[OPKB_MEM_MAX];   // alpha_j - beta_j in the element type
    unsigned use;         // bit j: pair j takes part (rho_j > 0)
    int m;
    T gamma;
    const T* v0;          // p (VMLMB) or g
    const T* x;
    const T* g;
    const T* ysrc;        // what is saved in Y[mark_next]: g, or p for BLMVM
    T* d;
    T* Snext;
    T* Ynext;
    Bound<T> lo, hi;
    int64_t n;
    int masked, bounded;
    int steepest;         // 0: two-loop assembly; 1: d = p|g (:554); 2: d given, finish only
    const T* pu;          // finish only: pending last axpy d += pa*pu on the free set {pw != 0}
    const T* pw;
    T pa;
};

template <typename T>
__device__ __forceinline__ T dir_elem(const DirArgs<T>& A, T v0, const T* yv, const T* sv)
{
    // apply_lbfgs! :716-745 / :747-779 with the scalars solved by the host;
    // same sequence of roundings as the reference's vupdate!/vscale! calls
    T v = v0;
    const bool upd = !A.masked || (v0 != T(0));
    if (upd) {
        for (int j = A.m - 1; j >= 0; --j)
            if ((A.use >> j) & 1u) v = v + A.ma[j] * yv[j];
    }
    v = A.gamma * v;
    if (upd) {
        for (int j = 0; j < A.m; ++j)
            if ((A.use >> j) & 1u) v = v + A.cc[j] * sv[j];
    }
    return v;
}

template <typename T, int MB>
__global__ void __launch_bounds__(OPKB_BLOCK) k_direction(DirArgs<T> A, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = A.n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // acc: 0 g.d; 1 |d|^2; 2 smin; 3 smax
    T smin = T(INFINITY), smax = T(0);
    double acc[4] = {0, 0, 0, 0};
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> pv = ld_pack<T>(A.v0, iv), px = ld_pack<T>(A.x, iv), pg = ld_pack<T>(A.g, iv);
        Pack<T> pys = (A.ysrc == A.g) ? pg : ld_pack<T>(A.ysrc, iv);
        Pack<T> pl, ph, pd;
        if (A.bounded) {
            pl = A.lo.pack(iv);
            ph = A.hi.pack(iv);
        } else {
            pl = splat_pack<T>(T(0));
            ph = pl;
        }
        if (A.steepest) {
            pd = pv;
            if (A.pu) {  // the last vupdate! of the sequential two-loop recursion (:738, :772)
                const Pack<T> u = ld_pack<T>(A.pu, iv);
                Pack<T> w = u;
                if (A.pw) w = ld_pack<T>(A.pw, iv);
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    if (!A.pw || w.v[e] != T(0)) pd.v[e] = pd.v[e] + A.pa * u.v[e];
            }
        } else {
            Pack<T> py[MB], ps[MB];
#pragma unroll
            for (int j = 0; j < MB; ++j) {
                if (j < A.m) {
                    py[j] = ld_pack<T>(A.Y[j], iv);
                    ps[j] = ld_pack<T>(A.S[j], iv);
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                T v = pv.v[e];
                const bool upd = !A.masked || (v != T(0));
#pragma unroll
                for (int j = MB - 1; j >= 0; --j)
                    if (j < A.m && upd && ((A.use >> j) & 1u)) v = v + A.ma[j] * py[j].v[e];
                v = A.gamma * v;
#pragma unroll
                for (int j = 0; j < MB; ++j)
                    if (j < A.m && upd && ((A.use >> j) & 1u)) v = v + A.cc[j] * ps[j].v[e];
                pd.v[e] = v;
            }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            T di = pd.v[e];
            // project_direction!(d, x, lo, hi, -, d) :535 (skipped for the
            // steepest descent: p is already projected, :554)
            if (A.bounded && A.steepest != 1)
                di = projdir(px.v[e], pl.v[e], ph.v[e], A.lo.has, A.hi.has, -1, di);
            pd.v[e] = di;
            acc[0] = dacc(pg.v[e], di, acc[0]);   // vdot(g, d) :537
            acc[1] = dacc(di, di, acc[1]);        // vnorm2(d) :629
            if (A.bounded)                        // step_limits(x, lo, hi, -1, d) :577
                step_limit_update(px.v[e], pl.v[e], ph.v[e], A.lo.has, A.hi.has, -di, smin, smax);
        }
        st_pack<T>(A.d, iv, pd);
        if (A.Snext) {                      // (the workspace path aliases instead of copying)
            st_pack<T>(A.Snext, iv, px);    // vcopy!(S[mark], x) :562
            st_pack<T>(A.Ynext, iv, pys);   // vcopy!(Y[mark], g|p) :563
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < A.n; ++i) {
            T yv[OPKB_MEM_MAX], sv[OPKB_MEM_MAX];
            T xi = A.x[i], gi = A.g[i], ys = A.ysrc[i], v0 = A.v0[i];
            T lo = A.bounded ? A.lo.at(i) : T(0), hi = A.bounded ? A.hi.at(i) : T(0);
            T di;
            if (A.steepest) {
                di = v0;
                if (A.pu && (!A.pw || A.pw[i] != T(0))) di = di + A.pa * A.pu[i];
            } else {
                for (int j = 0; j < A.m; ++j) {
                    yv[j] = A.Y[j][i];
                    sv[j] = A.S[j][i];
                }
                di = dir_elem(A, v0, yv, sv);
            }
            if (A.bounded && A.steepest != 1) di = projdir(xi, lo, hi, A.lo.has, A.hi.has, -1, di);
            acc[0] = dacc(gi, di, acc[0]);
            acc[1] = dacc(di, di, acc[1]);
            if (A.bounded) step_limit_update(xi, lo, hi, A.lo.has, A.hi.has, -di, smin, smax);
            A.d[i] = di;
            if (A.Snext) {
                A.Snext[i] = xi;
                A.Ynext[i] = ys;
            }
        }
    }
    acc[2] = (double)smin;
    acc[3] = (double)smax;
    block_reduce_publish<4, 8u, 4u>(acc, rs);  // slot 2: min, slot 3: max
}

// How a call of direction_run takes part in a chain of launches (opkb_chain.cuh; native driver, one
// rank).  HEAD: first launch of a chain -- the caller has sent state and coefficients to `dev`;
// ADOPT: the launch is already in flight (ws->pend), only its results are collected; PRELAUNCH:
// enqueue the launch and return (no result is awaited, the workspace is left as it was).
enum { DIR_CHAIN_HEAD = 1, DIR_CHAIN_ADOPT = 2, DIR_CHAIN_PRELAUNCH = 3 };
struct DirChain {
    int mode;
    ChainDev* dev;
    int prelaunch_next;            // HEAD / ADOPT: enqueue the next iteration's launch before waiting
    double rec[OPKB_CHAIN_NREC];   // HEAD / ADOPT: decision record of this launch (out)
    int chained;                   // out: the launch was a chained one
};

template <typename T>
static int chain_prelaunch_next(opkb_vmlmb* ws, int m, const int* slots, int mark_next, ChainDev* dev);

template <typename T>
static int direction_run(opkb_vmlmb* ws, int m, const int* slots, const double* alpha,
                         const double* c, double gamma, int mark_next, int steepest, double out[4],
                         const opkb_vector* pend_u = NULL, double pend_a = 0.0, double* out8 = NULL,
                         DirChain* ch = NULL)
{
    opkb_context* ctx = ws->ctx;
    int chmode = ch ? ch->mode : 0;
    if (ch) ch->chained = 0;
    if (chmode == 0) {
        int rcg = chain_guard(ws);
        if (rcg) return rcg;
    }
    if (chmode != DIR_CHAIN_PRELAUNCH) {
        ws->carry_valid = 0;
        ws->maskb_valid = 0;
        ws->xspec_mark = -1;
    }
    int xspec = 0;
    DirArgs<T> A;
    memset(&A, 0, sizeof(A));
    A.m = steepest ? 0 : m;
    A.use = 0;
    for (int j = 0; j < A.m; ++j) {
        A.S[j] = (const T*)ws->S[slots[j]]->data;
        A.Y[j] = (const T*)ws->Y[slots[j]]->data;
        A.ma[j] = (T)(-alpha[j]);
        A.cc[j] = (T)(c[j]);
        // NaN marks a skipped pair (rho <= 0, src/quasinewton.jl:727, :759)
        if (alpha[j] == alpha[j]) A.use |= (1u << j);
    }
    A.gamma = (T)gamma;
    A.v0 = (const T*)(ws->method > 0 ? ws->p->data : ws->g->data);
    if (!steepest && ws->method == 1) A.v0 = (const T*)ws->g->data;  // vcopy!(d, g) :525
    if (steepest == 2) A.v0 = (const T*)ws->d->data;                 // d already holds H*v
    A.x = (const T*)ws->x->data;
    A.g = (const T*)ws->g->data;
    A.ysrc = (const T*)(ws->method == 1 ? ws->p->data : ws->g->data);
    A.d = (T*)ws->d->data;
    A.Snext = NULL;   // saved by aliasing after the launch (vmlmb_alias)
    A.Ynext = NULL;
    A.lo = make_bound<T>(ws->lo, ws->lo_val, true);
    A.hi = make_bound<T>(ws->hi, ws->hi_val, false);
    A.n = ws->n;
    A.masked = (ws->method == 2);
    A.bounded = (ws->method > 0);
    A.steepest = steepest;
    if (pend_u) {
        A.pu = (const T*)pend_u->data;
        A.pa = (T)pend_a;
        A.pw = (ws->method == 2) ? (const T*)ws->p->data : NULL;
    }
    static thread_local int kinds[OPKB_NRED_MAX];
    for (int k = 0; k < OPKB_NRED_MAX; ++k) kinds[k] = OPKB_RSUM;
    kinds[2] = OPKB_RMIN;
    kinds[3] = OPKB_RMAX;
    static thread_local double part[2][OPKB_NRED_MAX];
    for (int q = 0; q < 2; ++q) {
        part[q][0] = part[q][1] = part[q][3] = 0.0;
        part[q][2] = INFINITY;
    }
    // speculative first trial of the next line search (stp = 1) in the same pass?
    const int fuse = (out8 && !steepest && A.m >= 1 && ws->method != 1 &&
                      (ws->obj == OPKB_OBJ_ROSENBROCK || ws->obj == OPKB_OBJ_QUADRATIC))
                         ? ws->obj : 0;
    if (out8) out8[7] = 0.0;
    int nparts = 0;
    int ran_lazy_p = 0;
    int64_t done = 0;

    // ---- whole tiles through the bulk-async ring (m >= 1 pairs) ---------------------
    if (!steepest && A.m >= 1) {
        Dir2Args<T> B;
        memset(&B, 0, sizeof(B));
        int nr = 0;
        B.rows[nr++] = A.v0;
        B.v0_from_g = (ws->method == 2);   // v0 = p: recomputed in the kernel, not read
        B.rows[nr++] = A.x;
        B.rows[nr++] = A.g;
        B.r_lo = B.r_hi = B.r_ysrc = -1;
        if (A.bounded && A.lo.arr) { B.r_lo = nr; B.rows[nr++] = A.lo.arr; }
        if (A.bounded && A.hi.arr) { B.r_hi = nr; B.rows[nr++] = A.hi.arr; }
        if (A.ysrc != A.g) { B.r_ysrc = nr; B.rows[nr++] = A.ysrc; }
        // incremental Gram (CARRY): only when the caller solves the two-loop recursion
        // from the Gram block of exactly these pairs (opkb_vmlmb_gram just before)
        int carry = 0;
        // user objective (VMLMB / L-BFGS): the trial point of the step the driver takes after a
        // successful update, stp = 1, is formed in the same pass (one more write, no more reads)
        if (!fuse && ws->opt_xspec && ws->obj == OPKB_OBJ_USER && ws->method != 1 &&
            ws->aliased_slot != mark_next) {
            B.xout = (T*)ws->S[mark_next]->data;
            xspec = 1;
        }
        if (fuse) {
            // x', g' go to the buffers of the recycled slot, which become x and g
            B.xout = (T*)ws->S[mark_next]->data;
            B.gout = (T*)ws->Y[mark_next]->data;
            B.pout = ws->p ? (T*)ws->p->data : NULL;
            // VMLMB: p' is not written (see p_stale); OPKB_LAZY_P=0 restores the write
            if (ws->opt_lazy_p && ws->method == 2) {
                B.pout = NULL;
                ran_lazy_p = 1;
            }
            carry = (A.m <= 12 && ws->Gc_valid && ws->Gc_m == A.m);
            for (int j = 0; carry && j < A.m; ++j) carry = (ws->Gc_slots[j] == slots[j]);
            // (enqueued ahead: the Gram block of exactly these pairs is what opkb_vmlmb_gram will
            // have assembled from the carry if the launch is ever adopted)
            if (chmode == DIR_CHAIN_PRELAUNCH) carry = (A.m <= 12);
            // mark_next is the oldest pair when the memory is full, a free slot otherwise
            if (carry && A.m == ws->mem && slots[0] != mark_next) carry = 0;
            if (carry && A.m < ws->mem)
                for (int j = 0; j < A.m; ++j) if (slots[j] == mark_next) carry = 0;
            if (ws->hist == 1) {
                // iterate history: nothing to write, no spare buffers; same rule for the element types
                if (ws->opt_carry == 0 || (sizeof(T) == 4 && ws->opt_carry != 2 && ws->mem > 8)) carry = 0;
            } else if (!ws->carry_s || !ws->carry_y) {
                // the spare buffers exist iff the workspace was created for it: same on every rank
                carry = 0;
            }
            if (carry && ws->hist != 1) {
                B.sout = (T*)ws->carry_s->data;
                B.yout = (T*)ws->carry_y->data;
            }
        }
        for (int j = 0; j < A.m; ++j) {
            B.rows[nr++] = A.Y[j];
            B.rows[nr++] = A.S[j];
            B.ma[j] = A.ma[j];
            B.cc[j] = A.cc[j];
        }
        if (fuse == OPKB_OBJ_QUADRATIC) {   // objective data: after the pair rows
            B.r_a = nr;
            B.rows[nr++] = (const T*)ws->qa->data;
            B.r_c = nr;
            B.rows[nr++] = (const T*)ws->qc->data;
        }
        B.use = A.use;
        B.m = A.m;
        B.hist = (ws->hist == 1);
        B.nrows = nr;
        B.gamma = A.gamma;
        B.lo_val = A.lo.val;
        B.hi_val = A.hi.val;
        B.has_lo = A.lo.has;
        B.has_hi = A.hi.has;
        B.masked = A.masked;
        B.bounded = A.bounded;
        B.d = A.d;
        B.Snext = A.Snext;
        B.Ynext = A.Ynext;
        // ring geometry: 4 KB rows when two stages fit in ~220 KB, else 2 KB rows; the
        // CARRY instances use 3.5 KB rows (7 consumer warps) and need two stages of them
        const size_t budget = 220 * 1024;
        const int carry_mb = (A.m <= 5) ? 5 : ((A.m <= 8) ? 8 : ((A.m == 10) ? 10 : 12));
        int carry_rowb = (carry_mb == 5) ? DIR2_CARRY_ROWB5 : DIR2_CARRY_ROWB;
        // Float32, 6..8 pairs: the SPLIT instances (10 consumer warps on 5 KB rows, two stages)
        static const int split_on = env_int("OPKB_DIR_SPLIT", 1);
        if (sizeof(T) == 4 && carry_mb == 8 && split_on && (size_t)nr * DIR2_SPLIT_ROWB * 2 <= budget)
            carry_rowb = DIR2_SPLIT_ROWB;
        if (carry && (size_t)nr * carry_rowb * 2 > budget) carry = 0;
        int tile = (int)(4096 / sizeof(T));
        if ((size_t)nr * tile * sizeof(T) * 2 > budget || A.m > 12) tile = (int)(2048 / sizeof(T));
        if (const char* e = getenv("OPKB_DIR_ROWB")) {   // tuning knob: 2048 | 4096 bytes per row
            const int rb = atoi(e);
            if (rb == 2048 || (rb == 4096 && A.m <= 12)) tile = rb / (int)sizeof(T);
        }
        if (carry) tile = carry_rowb / (int)sizeof(T);
        size_t stage_bytes = (size_t)nr * tile * sizeof(T);
        int nstage = (int)(budget / stage_bytes);
        if (const char* e = getenv("OPKB_DIR_NSTAGE")) nstage = atoi(e);
        if (nstage > DIR2_MAXSTAGE) nstage = DIR2_MAXSTAGE;
        const int64_t ntiles = (ws->n + tile - 1) / tile;
        // (an EMPTY shard still runs the ring kernel -- one CTA, no tile: neutral partials -- so
        // that every rank of a job reduces the same slots and takes the same carry decision)
        if (nstage >= 2) {
            B.tile = tile;
            B.nstage = nstage;
            B.n = ws->n;
            const size_t smem = stage_bytes * nstage;
            int mb;
            if (carry) mb = carry_mb;
            else if (tile * sizeof(T) == 4096) mb = (A.m <= 8) ? 8 : 12;
            else mb = (A.m <= 12) ? 12 : 20;
            const int nd = carry ? 4 * mb + 4 : 0;
            const int ncorr = carry ? 32 * ((mb * (mb + 1) + 31) / 32) : 0;
            int64_t grid = ctx->sm_count;
            if (grid > ntiles) grid = ntiles;
            if (grid < 1) grid = 1;
            // chained launch?  (Float64 carry instances of up to OPKB_CHAIN_MB pairs, one rank with the
            // polled completion flag; anything else runs as it always did)
            bool chained = false;
            unsigned int seq = 0;
            int resoff = 0;
            if (chmode) {
                chained = carry && fuse && sizeof(T) == 8 && (mb == 5 || mb == 8) && ws->mem <= OPKB_CHAIN_MB &&
                          opkb_chain_supported(ctx);
                if (!chained) {
                    if (chmode == DIR_CHAIN_ADOPT)
                        return opkb_set_error(OPKB_EINVAL, "internal error: the launch in flight is not the one the "
                                              "workspace would enqueue now");
                    if (chmode == DIR_CHAIN_PRELAUNCH) return 0;
                    chmode = 0;     // (a head launch that cannot be chained is an ordinary launch)
                }
            }
            int rc = 0;
            if (chmode == DIR_CHAIN_ADOPT) {
                if (ws->pend.fuse != fuse || ws->pend.mb != mb || ws->pend.nd != nd || ws->pend.ncorr != ncorr)
                    return opkb_set_error(OPKB_EINVAL, "internal error: geometry of the launch in flight");
                seq = ws->pend.seq;
                resoff = ws->pend.resoff;
                ran_lazy_p = ws->pend.ran_lazy_p;
                ws->pend.valid = 0;
                ws->chain_adopted += 1;
            } else {
                if (chained) {
                    B.chain = ch->dev;
                    resoff = ((ctx->seq + 1u) & 1u) ? OPKB_NRED_MAX / 2 : 0;   // results by sequence parity
                }
                rc = carry ? dir2_launch_carry<T>(ctx, B, fuse, mb, smem, (int)grid, resoff)
                           : dir2_launch_plain<T>(ctx, B, fuse, mb, smem, (int)grid);
                if (rc) return rc;
                seq = ctx->seq;
                if (chmode == DIR_CHAIN_HEAD) ws->chain_heads += 1;
            }
            if (chmode == DIR_CHAIN_PRELAUNCH) {
                ws->pend.valid = 1;
                ws->pend.go = -1;
                ws->pend.fuse = fuse;
                ws->pend.mb = mb;
                ws->pend.nd = nd;
                ws->pend.ncorr = ncorr;
                ws->pend.m = A.m;
                ws->pend.mark_next = mark_next;
                ws->pend.ran_lazy_p = ran_lazy_p;
                ws->pend.resoff = resoff;
                ws->pend.seq = seq;
                for (int j = 0; j < A.m; ++j) ws->pend.slots[j] = slots[j];
                ws->pend.x = ws->x->data;
                ws->pend.g = ws->g->data;
                return 0;
            }
            if (chained) {
                ch->chained = 1;
                if (ch->prelaunch_next) {
                    rc = chain_prelaunch_next<T>(ws, A.m, slots, mark_next, ch->dev);
                    if (rc) return rc;
                }
                rc = opkb_finish_reduction_seq(ctx, seq, resoff, 12 + nd + ncorr + OPKB_CHAIN_NREC, part[nparts++]);
                if (rc) return rc;
                // what the last CTA decided about the next launch
                memcpy(ch->rec, part[0] + 12 + nd + ncorr, sizeof(double) * OPKB_CHAIN_NREC);
                memcpy(ws->last_rec, ch->rec, sizeof(ch->rec));
                ws->last_rec_valid = 1;
                ws->last_rec_m = (A.m + 1 < ws->mem ? A.m + 1 : ws->mem);
                if (ws->pend.valid) {
                    ws->pend.go = (ch->rec[0] != 0.0) ? 1 : 0;
                    memcpy(ws->pend.rec, ch->rec, sizeof(ch->rec));
                    if (!ws->pend.go) {     // it returns at once, nothing is touched: forgotten
                        ws->pend.valid = 0;
                        ws->chain_nogo += 1;
                    }
                }
            } else {
                rc = opkb_finish_reduction(ctx, (fuse ? 12 : 4) + nd + ncorr, kinds, part[nparts++]);
                if (rc) return rc;
            }
            done = B.n;
            if (carry) {
                const double* r = part[0] + 12;
                memcpy(ws->carry_dense, r, sizeof(double) * nd);
                memcpy(ws->carry_corr, r + nd, sizeof(double) * mb * (mb + 1));
                ws->carry_m = A.m;
                ws->carry_mb = mb;
                ws->carry_mark = mark_next;
                for (int j = 0; j < A.m; ++j) ws->carry_slots[j] = slots[j];
                // the last correction slot counts the warps that gave up on their corrections (too
                // many variables change their status at this trial point, DIR2_CORR_MAX): the block
                // cannot be assembled then.  The count is combined over the ranks like every other
                // slot: the same decision everywhere.
                ws->carry_valid = (r[nd + ncorr - 1] == 0.0) ? 1 : 0;
                if (!ws->carry_valid) ws->carry_skips += 1;
            }
            if (fuse) {
                const double* r = part[0];
                out8[0] = (fuse == OPKB_OBJ_ROSENBROCK) ? r[4] + r[5] : 0.5 * r[4];
                out8[1] = r[6];
                out8[2] = r[7];
                out8[3] = r[8];
                out8[4] = r[9];
                out8[5] = r[10];
                out8[6] = r[11];
                out8[7] = 1.0;
            }
        }
    }

    if (chmode == DIR_CHAIN_PRELAUNCH) return 0;   // (the ring did not fit: nothing enqueued)
    if (chmode == DIR_CHAIN_ADOPT && done < ws->n)
        return opkb_set_error(OPKB_EINVAL, "internal error: the launch in flight did not cover the vector");
    // ---- the rest (tail of the ring path, or everything) with the register kernel -----
    if (done < ws->n && ws->hist == 1 && !steepest && A.m >= 1)
        return opkb_set_error(OPKB_EINVAL, "the direction ring does not fit (m = %d): set OPKB_HISTORY=0", A.m);
    // (nparts == 0: the ring did not run -- steepest descent, the tail of the step-by-step recursion --:
    // this launch then runs on EVERY rank, also on one whose shard is empty, so that all ranks
    // reduce the same slots in the same phases)
    if (done < ws->n || nparts == 0) {
        {
            int rc = vmlmb_ensure_p(ws);   // v0 / the mask are read from p here
            if (rc) return rc;
        }
        const int64_t off = done;
        for (int j = 0; j < A.m; ++j) {
            A.S[j] += off;
            A.Y[j] += off;
        }
        A.v0 += off;
        A.x += off;
        A.g += off;
        A.ysrc += off;
        A.d += off;
        if (A.pu) A.pu += off;
        if (A.pw) A.pw += off;
        if (A.lo.arr) A.lo.arr += off;
        if (A.hi.arr) A.hi.arr += off;
        A.n = ws->n - off;
        ReduceScratch rs = scratch(ctx);
        int grid = opkb_grid_for(ctx, A.n / VecOf<T>::N, 2);
        if (A.m <= 4)
            LAUNCHB(ctx, OPKB_K_DIRECTION, (k_direction<T, 4>), grid, OPKB_BLOCK, 0, A, rs);
        else if (A.m <= 8)
            LAUNCHB(ctx, OPKB_K_DIRECTION, (k_direction<T, 8>), grid, OPKB_BLOCK, 0, A, rs);
        else if (A.m <= 12)
            LAUNCHB(ctx, OPKB_K_DIRECTION, (k_direction<T, 12>), grid, OPKB_BLOCK, 0, A, rs);
        else if (A.m <= 16)
            LAUNCHB(ctx, OPKB_K_DIRECTION, (k_direction<T, 16>), grid, OPKB_BLOCK, 0, A, rs);
        else
            LAUNCHB(ctx, OPKB_K_DIRECTION, (k_direction<T, OPKB_MEM_MAX>), grid, OPKB_BLOCK, 0, A, rs);
        int rc = opkb_finish_reduction(ctx, 4, kinds, part[nparts++]);
        if (rc) return rc;
    }
    out[0] = part[0][0];
    out[1] = part[0][1];
    out[2] = part[0][2];
    out[3] = part[0][3];
    if (nparts == 2) {  // fixed order: ring part, then tail
        out[0] = out[0] + part[1][0];
        out[1] = out[1] + part[1][1];
        out[2] = (part[1][2] < out[2]) ? part[1][2] : out[2];
        out[3] = (part[1][3] > out[3]) ? part[1][3] : out[3];
    }
    if (!A.bounded) {  // no step limits for L-BFGS (:581-584)
        out[2] = INFINITY;
        out[3] = INFINITY;
    } else if (!A.lo.has && !A.hi.has) {
        out[2] = INFINITY;
        out[3] = INFINITY;
    }
    if (fuse) {
        // S[mark'] <- x, Y[mark'] <- g (:562-563) and x, g <- the trial point just
        // written into the recycled buffers: two pointer swaps
        if (ws->method == 2) ws->p_stale = ran_lazy_p;
        vmlmb_unalias(ws);
        void* t = ws->S[mark_next]->data;
        ws->S[mark_next]->data = ws->x->data;
        ws->x->data = t;
        t = ws->Y[mark_next]->data;
        ws->Y[mark_next]->data = ws->g->data;
        ws->g->data = t;
        return 0;
    }
    // vcopy!(S[mark], x); vcopy!(Y[mark], g|p) :562-563, without a copy
    vmlmb_alias(ws, mark_next);
    if (xspec && done == ws->n) ws->xspec_mark = mark_next;
    return 0;
}

// Enqueue the fused launch of the iteration AFTER the one whose launch has just been issued (pairs
// `slots`, saving into `mark_next`), as the workspace will look once that trial point has been
// accepted: x, g <-> S, Y[mark_next] (the swaps at the end of direction_run) and, with stored
// differences, S, Y[mark_next] <-> the spare buffers that received the new pair (gram_from_carry).
// The pointers are rotated for the build and rotated back: the workspace itself is untouched.
template <typename T>
static int chain_prelaunch_next(opkb_vmlmb* ws, int m, const int* slots, int mark_next, ChainDev* dev)
{
    if (ws->aliased_slot >= 0 || ws->pend.valid) return 0;     // (not in steady state yet)
    if (ws->hist != 1 && (!ws->carry_s || !ws->carry_y)) return 0;
    const int mem = ws->mem, drop = (m == mem) ? 1 : 0;
    const int m2 = (m + 1 < mem ? m + 1 : mem);
    int slots2[OPKB_MEM_MAX];
    for (int j = 0; j < m2 - 1; ++j) slots2[j] = slots[j + drop];
    slots2[m2 - 1] = mark_next;
    const int mn2 = (mark_next < mem - 1 ? mark_next + 1 : 0);
    auto swp = [](void*& a, void*& b) { void* t = a; a = b; b = t; };
    auto rotate = [&](bool undo) {
        if (!undo) {
            swp(ws->S[mark_next]->data, ws->x->data);
            swp(ws->Y[mark_next]->data, ws->g->data);
            if (ws->hist != 1) {
                swp(ws->S[mark_next]->data, ws->carry_s->data);
                swp(ws->Y[mark_next]->data, ws->carry_y->data);
            }
        } else {
            if (ws->hist != 1) {
                swp(ws->Y[mark_next]->data, ws->carry_y->data);
                swp(ws->S[mark_next]->data, ws->carry_s->data);
            }
            swp(ws->Y[mark_next]->data, ws->g->data);
            swp(ws->S[mark_next]->data, ws->x->data);
        }
    };
    double zero[OPKB_MEM_MAX], o4[4], o8[8];
    for (int j = 0; j < OPKB_MEM_MAX; ++j) zero[j] = 0.0;
    DirChain pc;
    memset(&pc, 0, sizeof(pc));
    pc.mode = DIR_CHAIN_PRELAUNCH;
    pc.dev = dev;
    rotate(false);
    const int rc = direction_run<T>(ws, m2, slots2, zero, zero, 1.0, mn2, 0, o4, NULL, 0.0, o8, &pc);
    rotate(true);
    return rc;
}

// opkb_vmlmb_direction_trial for the native driver with device-resident decisions (opkb_chain.cuh).
// `state`: the solver's scalars as chain_decide() needs them (everything but the coefficients, the
// pairs and the Gram block, which are filled in here); `dev`: the device copy.  If the launch this
// call asks for is already in flight (enqueued ahead by the previous call, go = 1 from its last
// CTA) it is adopted after a bit-for-bit comparison of the coefficients; otherwise state and
// coefficients go to the device and the launch is the head of a new chain.  `prelaunch_next`: also
// enqueue the launch of the following iteration before waiting.  Falls back to the ordinary launch
// when this shape has no chained instance.
int opkb_vmlmb_direction_trial_chain(opkb_vmlmb* ws, int m, const int* slots, const double* alpha,
                                     const double* c, double gamma, int mark_next, double out[4],
                                     double trial[8], const ChainDev* state, ChainDev* dev, ChainDev* stage,
                                     int prelaunch_next)
{
    // (stage: one ChainDev of pinned host memory -- a head launch is awaited before this function
    // returns, so the copy that precedes it in the stream has completed when the buffer is written again)
    if (!ws || !state || !dev || !stage) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (mark_next < 0 || mark_next >= ws->mem) return opkb_set_error(OPKB_EINVAL, "mark out of range");
    if (ws->eltype != OPKB_DOUBLE || m < 1 || m > OPKB_CHAIN_MB || ws->mem > OPKB_CHAIN_MB ||
        !(ws->Gc_valid && ws->Gc_m == m)) {
        int rc = chain_guard(ws);
        if (rc) return rc;
        return opkb_vmlmb_direction_trial(ws, m, slots, alpha, c, gamma, mark_next, out, trial);
    }
    DirChain ch;
    memset(&ch, 0, sizeof(ch));
    ch.mode = DIR_CHAIN_HEAD;
    ch.dev = dev;
    ch.prelaunch_next = prelaunch_next;
    auto same_bits = [](double a, double b) { return (a != a && b != b) || memcmp(&a, &b, sizeof(double)) == 0; };
    {
        if (ws->opt_chain_check && ws->last_rec_valid) {
            // the previous chained launch and nothing else since: its last CTA must have foreseen this call
            const double* r = ws->last_rec;
            bool ok = (r[0] == 1.0 && ws->last_rec_m == m && same_bits(gamma, r[2]));
            for (int j = 0; ok && j < m; ++j) ok = same_bits(alpha[j], r[3 + j]) && same_bits(c[j], r[3 + OPKB_CHAIN_MB + j]);
            ws->chain_checked += 1;
            if (!ok) {
                ws->chain_check_failed += 1;
                fprintf(stderr, "[chain check] device go=%g why=%g gamma=%.17g | host m=%d gamma=%.17g", r[0], r[1], r[2], m, gamma);
                for (int j = 0; j < m; ++j) fprintf(stderr, " a%d %.17g/%.17g c%d %.17g/%.17g", j, r[3 + j], alpha[j], j, r[3 + OPKB_CHAIN_MB + j], c[j]);
                fprintf(stderr, "\n");
            }
        }
        ws->last_rec_valid = 0;
    }
    if (ws->pend.valid) {
        bool same = (ws->pend.go == 1 && ws->pend.m == m && ws->pend.mark_next == mark_next &&
                     ws->pend.x == ws->x->data && ws->pend.g == ws->g->data && same_bits(gamma, ws->pend.rec[2]));
        for (int j = 0; same && j < m; ++j)
            same = (ws->pend.slots[j] == slots[j] && same_bits(alpha[j], ws->pend.rec[3 + j]) &&
                    same_bits(c[j], ws->pend.rec[3 + OPKB_CHAIN_MB + j]));
        if (!same) {
            ws->pend.valid = 0;
            cudaStreamSynchronize(ws->ctx->stream);
            return opkb_set_error(OPKB_EINVAL, "internal error: the chained launch in flight does not match the "
                                  "host's decision (m %d/%d, mark %d/%d, gamma %.17g/%.17g)", ws->pend.m, m,
                                  ws->pend.mark_next, mark_next, ws->pend.rec[2], gamma);
        }
        ch.mode = DIR_CHAIN_ADOPT;
    } else {
        ChainDev& hs = *stage;
        hs = *state;
        hs.go = 1;
        hs.m = m;
        for (int j = 0; j < m; ++j) {
            if (ws->Gc_slots[j] != slots[j]) return opkb_set_error(OPKB_EINVAL, "the Gram block is not that of these pairs");
            hs.alpha[j] = alpha[j];
            hs.c[j] = c[j];
            hs.slots[j] = slots[j];
        }
        hs.gamma_dir = gamma;
        hs.mark_next = mark_next;
        hs.mem = ws->mem;
        hs.method = ws->method;
        hs.obj = ws->obj;
        const Bound<double> lo = make_bound<double>(ws->lo, ws->lo_val, true), hi = make_bound<double>(ws->hi, ws->hi_val, false);
        hs.limits_inf = (ws->method == 0 || (!lo.has && !hi.has)) ? 1 : 0;
        hs.pad = ws->opt_chain_dbg;
        memcpy(hs.G, ws->Gc, sizeof(double) * 2 * m * (m + 1));
        opkb_bind_device(ws->ctx);
        OPKB_CUDA(cudaMemcpyAsync(dev, stage, sizeof(ChainDev), cudaMemcpyHostToDevice, ws->ctx->stream));
    }
    return direction_run<double>(ws, m, slots, alpha, c, gamma, mark_next, 0, out, NULL, 0.0, trial, &ch);
}

int opkb_vmlmb_chain_possible(const opkb_vmlmb* ws)
{
    return ws && ws->eltype == OPKB_DOUBLE && ws->mem <= OPKB_CHAIN_MB && ws->method != 1 && ws->opt_carry != 0 &&
           !ws->opt_carry_check &&
           (ws->obj == OPKB_OBJ_ROSENBROCK || ws->obj == OPKB_OBJ_QUADRATIC) && opkb_chain_supported(ws->ctx);
}

int opkb_vmlmb_chain_pending(const opkb_vmlmb* ws) { return ws ? ws->pend.valid : 0; }

extern "C" int opkb_vmlmb_chain_check_stats(const opkb_vmlmb* ws, int64_t* checked, int64_t* failed)
{
    if (!ws) return opkb_set_error(OPKB_EINVAL, "NULL workspace");
    if (checked) *checked = ws->chain_checked;
    if (failed) *failed = ws->chain_check_failed;
    return 0;
}

extern "C" int opkb_vmlmb_chain_stats(const opkb_vmlmb* ws, int64_t* heads, int64_t* adopted, int64_t* nogo)
{
    if (!ws) return opkb_set_error(OPKB_EINVAL, "NULL workspace");
    if (heads) *heads = ws->chain_heads;
    if (adopted) *adopted = ws->chain_adopted;
    if (nogo) *nogo = ws->chain_nogo;
    return 0;
}

extern "C" int opkb_vmlmb_direction(opkb_vmlmb* ws, int m, const int* slots, const double* alpha,
                                    const double* c, double gamma, int mark_next, double out[4])
{
    int rc = check_mark(ws, mark_next);
    if (rc) return rc;
    if (m < 1 || m > ws->mem || m > OPKB_MEM_MAX || !slots || !alpha || !c || !out)
        return opkb_set_error(OPKB_EINVAL, "invalid argument (the Gram-form direction supports m <= 20)");
    for (int j = 0; j < m; ++j)
        if (slots[j] < 0 || slots[j] >= ws->mem) return opkb_set_error(OPKB_EINVAL, "slot out of range");
    return (ws->eltype == OPKB_DOUBLE)
               ? direction_run<double>(ws, m, slots, alpha, c, gamma, mark_next, 0, out)
               : direction_run<float>(ws, m, slots, alpha, c, gamma, mark_next, 0, out);
}

extern "C" int opkb_vmlmb_direction_trial(opkb_vmlmb* ws, int m, const int* slots, const double* alpha,
                                          const double* c, double gamma, int mark_next, double out[4],
                                          double trial[8])
{
    int rc = check_mark(ws, mark_next);
    if (rc) return rc;
    if (m < 1 || m > ws->mem || m > OPKB_MEM_MAX || !slots || !alpha || !c || !out || !trial)
        return opkb_set_error(OPKB_EINVAL, "invalid argument (the Gram-form direction supports m <= 20)");
    for (int j = 0; j < m; ++j)
        if (slots[j] < 0 || slots[j] >= ws->mem) return opkb_set_error(OPKB_EINVAL, "slot out of range");
    return (ws->eltype == OPKB_DOUBLE)
               ? direction_run<double>(ws, m, slots, alpha, c, gamma, mark_next, 0, out, NULL, 0.0, trial)
               : direction_run<float>(ws, m, slots, alpha, c, gamma, mark_next, 0, out, NULL, 0.0, trial);
}

extern "C" int opkb_vmlmb_discard_trial(opkb_vmlmb* ws, int mark_next)
{
    // undo the two pointer swaps of a fused trial that the host does not use
    // because the direction was rejected (:542-546): x, g are the current
    // iterate again, S[mark'], Y[mark'] alias them, p is recomputed (:396)
    int rc = check_mark(ws, mark_next);
    if (rc) return rc;
    rc = chain_guard(ws);
    if (rc) return rc;
    ws->carry_valid = 0;
    ws->maskb_valid = 0;
    ws->xspec_mark = -1;
    void* t = ws->S[mark_next]->data;
    ws->S[mark_next]->data = ws->x->data;
    ws->x->data = t;
    t = ws->Y[mark_next]->data;
    ws->Y[mark_next]->data = ws->g->data;
    ws->g->data = t;
    if (ws->p) {
        rc = opkb_project_direction(ws->ctx, ws->p, ws->x, ws->lo, ws->lo_val, ws->hi, ws->hi_val,
                                    OPKB_BACKWARD, ws->g);
        if (rc) return rc;
        ws->p_stale = 0;
    }
    ws->aliased_slot = -1;
    vmlmb_alias(ws, mark_next);
    return 0;
}

extern "C" int opkb_vmlmb_steepest(opkb_vmlmb* ws, int mark_next, double out[4])
{
    int rc = check_mark(ws, mark_next);
    if (rc) return rc;
    if (!out) return opkb_set_error(OPKB_EINVAL, "NULL result");
    return (ws->eltype == OPKB_DOUBLE)
               ? direction_run<double>(ws, 0, NULL, NULL, NULL, 1.0, mark_next, 1, out)
               : direction_run<float>(ws, 0, NULL, NULL, NULL, 1.0, mark_next, 1, out);
}

static opkb_vector* pair_vector(opkb_vmlmb* ws, int which, int slot)
{
    if (slot < 0 || slot >= ws->mem) return NULL;
    if (which == 'S') return ws->S[slot];
    if (which == 'Y') return ws->Y[slot];
    return NULL;
}

extern "C" int opkb_vmlmb_finish_direction(opkb_vmlmb* ws, int mark_next, int u_which, int u_slot,
                                           double a, double out[4])
{
    int rc = check_mark(ws, mark_next);
    if (rc) return rc;
    if (!out) return opkb_set_error(OPKB_EINVAL, "NULL result");
    const opkb_vector* u = NULL;
    if (u_which) {
        u = pair_vector(ws, u_which, u_slot);
        if (!u) return opkb_set_error(OPKB_EINVAL, "invalid pending vector");
    }
    return (ws->eltype == OPKB_DOUBLE)
               ? direction_run<double>(ws, 0, NULL, NULL, NULL, 1.0, mark_next, 2, out, u, a)
               : direction_run<float>(ws, 0, NULL, NULL, NULL, 1.0, mark_next, 2, out, u, a);
}

// ---------------------------------------------------------------------------
// Sequential two-loop recursion, fused step by step (parity-grade: the
// intermediate vector is rounded in the element type after every vupdate!,
// exactly as the reference does; used for Float32).
template <typename T> struct StepArgs {
    const T* vsrc;
    T* vdst;          // NULL: v is not modified by this step
    const T* u;       // pending axpy vector or NULL
    const T* z;       // out[0] = <z, v>
    const T* z2;      // out[1] = <z2, z3>, out[2] = <z2, z2>, or NULL
    const T* z3;
    const T* w;       // free set {w != 0} or NULL
    // the same set, one byte per 16-byte pack (bit e = element e of the pack is free): the
    // first step of a two-loop recursion writes it (wb_out) while it reads p, the 2m - 1
    // following steps read n/VEC bytes instead of a whole vector
    const unsigned char* wb;
    unsigned char* wb_out;
    T a, gamma;
    int has_scale;
    int64_t n;
};

template <typename T>
__device__ __forceinline__ T step_elem(const StepArgs<T>& A, T v, T u, bool fr)
{
    if (A.u && fr) v = v + A.a * u;     // vupdate!(v, [sel,] a, u) :729, :738, :761, :772
    if (A.has_scale) v = A.gamma * v;   // vscale!(v, gamma) :707, :768 (all elements)
    return v;
}

template <typename T>
__global__ void __launch_bounds__(OPKB_BLOCK) k_twoloop_step(StepArgs<T> A, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = A.n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc[3] = {0, 0, 0};
    // (two packs per thread in flight: the loads of both iterations are issued before the first use)
#pragma unroll 2
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> v = ld_pack<T>(A.vsrc, iv), z = ld_pack<T>(A.z, iv), u = v, z2 = v, z3 = v;
        if (A.u) u = ld_pack<T>(A.u, iv);
        unsigned mb = 0xffu;
        if (A.wb) {
            mb = A.wb[iv];
        } else if (A.w) {
            const Pack<T> w = ld_pack<T>(A.w, iv);
            mb = 0u;
#pragma unroll
            for (int e = 0; e < VEC; ++e) mb |= (w.v[e] != T(0)) ? (1u << e) : 0u;
            if (A.wb_out) A.wb_out[iv] = (unsigned char)mb;
        }
        if (A.z2) {
            z2 = ld_pack<T>(A.z2, iv);
            z3 = ld_pack<T>(A.z3, iv);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const bool fr = (mb >> e) & 1u;
            v.v[e] = step_elem(A, v.v[e], u.v[e], fr);
            if (fr) {
                acc[0] = dacc(z.v[e], v.v[e], acc[0]);
                if (A.z2) {
                    acc[1] = dacc(z2.v[e], z3.v[e], acc[1]);
                    acc[2] = dacc(z2.v[e], z2.v[e], acc[2]);
                }
            }
        }
        if (A.vdst) st_pack<T>(A.vdst, iv, v);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < A.n; ++i) {
            const bool fr = !A.w || (A.w[i] != T(0));
            T v = step_elem(A, A.vsrc[i], A.u ? A.u[i] : T(0), fr);
            if (fr) {
                acc[0] = dacc(A.z[i], v, acc[0]);
                if (A.z2) {
                    acc[1] = dacc(A.z2[i], A.z3[i], acc[1]);
                    acc[2] = dacc(A.z2[i], A.z2[i], acc[2]);
                }
            }
            if (A.vdst) A.vdst[i] = v;
        }
    }
    block_reduce_publish<3, 0u, 0u>(acc, rs);
}

template <typename T>
static int step_run(opkb_vmlmb* ws, int first, const opkb_vector* u, double a, int has_scale,
                    double gamma, const opkb_vector* z, const opkb_vector* z2, const opkb_vector* z3,
                    double out[3])
{
    opkb_context* ctx = ws->ctx;
    if (ws->hist == 1)
        return opkb_set_error(OPKB_EINVAL, "this workspace keeps the iterate history of the Gram path: the "
                              "step-by-step recursion needs its own workspace");
    {
        int rc = vmlmb_ensure_p(ws);
        if (rc) return rc;
    }
    StepArgs<T> A;
    memset(&A, 0, sizeof(A));
    const opkb_vector* src = first ? (ws->method == 2 ? ws->p : ws->g) : ws->d;   // vcopy!(d, p|g) :525/:528
    A.vsrc = (const T*)src->data;
    A.vdst = (u || has_scale || first) ? (T*)ws->d->data : NULL;
    if (first && !u && !has_scale) A.vdst = NULL;   // nothing to write yet: d stays "= p|g" virtually
    A.u = u ? (const T*)u->data : NULL;
    A.z = (const T*)z->data;
    A.z2 = z2 ? (const T*)z2->data : NULL;
    A.z3 = z3 ? (const T*)z3->data : NULL;
    A.w = (ws->method == 2) ? (const T*)ws->p->data : NULL;
    if (A.w) {
        if (!ws->opt_maskb) ws->maskb_failed = 1;   // OPKB_MASKB=0: A/B switch
        if (!ws->maskb && !ws->maskb_failed) {
            if (cudaMalloc(&ws->maskb, (size_t)(ws->n / VecOf<T>::N) + 16) != cudaSuccess) {
                cudaGetLastError();
                ws->maskb = NULL;
                ws->maskb_failed = 1;      // no memory for it: every step reads p
            }
        }
        if (ws->maskb && first) {
            A.wb_out = (unsigned char*)ws->maskb;
            ws->maskb_valid = 1;
        } else if (ws->maskb && ws->maskb_valid) {
            A.wb = (const unsigned char*)ws->maskb;
        }
    }
    A.a = (T)a;
    A.gamma = (T)gamma;
    A.has_scale = has_scale;
    A.n = ws->n;
    int grid = opkb_grid_for(ctx, ws->n / VecOf<T>::N, 4);
    LAUNCHB(ctx, OPKB_K_REDUCE, (k_twoloop_step<T>), grid, OPKB_BLOCK, 0, A, scratch(ctx));
    return opkb_finish_reduction(ctx, 3, NULL, out);
}

extern "C" int opkb_vmlmb_twoloop_step(opkb_vmlmb* ws, int first, int u_which, int u_slot, double a,
                                       int has_scale, double gamma, int z_which, int z_slot,
                                       int want_rho, double out[3])
{
    if (!ws || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    const opkb_vector* u = NULL;
    if (u_which) {
        u = pair_vector(ws, u_which, u_slot);
        if (!u) return opkb_set_error(OPKB_EINVAL, "invalid pending vector");
    }
    const opkb_vector* z = pair_vector(ws, z_which, z_slot);
    if (!z) return opkb_set_error(OPKB_EINVAL, "invalid dot vector");
    const opkb_vector* z2 = want_rho ? ws->Y[z_slot] : NULL;
    const opkb_vector* z3 = want_rho ? ws->S[z_slot] : NULL;
    return (ws->eltype == OPKB_DOUBLE)
               ? step_run<double>(ws, first, u, a, has_scale, gamma, z, z2, z3, out)
               : step_run<float>(ws, first, u, a, has_scale, gamma, z, z2, z3, out);
}

// :512-518 fused: S[mark] -= x; Y[mark] -= g|p; out = { <y, s>, <y, y> } (unrestricted)
template <typename T>
__global__ void __launch_bounds__(OPKB_BLOCK)
k_pair_update(T* s, T* y, const T* x, const T* g, int64_t n, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc[2] = {0, 0};
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> ps = ld_pack<T>(s, iv), py = ld_pack<T>(y, iv), px = ld_pack<T>(x, iv), pg = ld_pack<T>(g, iv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            ps.v[e] = ps.v[e] - px.v[e];   // vupdate!(S[mark], -1, x) :512
            py.v[e] = py.v[e] - pg.v[e];   // vupdate!(Y[mark], -1, g|p) :513
            acc[0] = dacc(py.v[e], ps.v[e], acc[0]);
            acc[1] = dacc(py.v[e], py.v[e], acc[1]);
        }
        st_pack<T>(s, iv, ps);
        st_pack<T>(y, iv, py);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < n; ++i) {
            T sv = s[i] - x[i], yv = y[i] - g[i];
            s[i] = sv;
            y[i] = yv;
            acc[0] = dacc(yv, sv, acc[0]);
            acc[1] = dacc(yv, yv, acc[1]);
        }
    }
    block_reduce_publish<2, 0u, 0u>(acc, rs);
}

extern "C" int opkb_vmlmb_pair_update(opkb_vmlmb* ws, int mark, double out[2])
{
    int rc = check_mark(ws, mark);
    if (rc) return rc;
    if (!out) return opkb_set_error(OPKB_EINVAL, "NULL result");
    opkb_context* ctx = ws->ctx;
    if (ws->hist == 1)
        return opkb_set_error(OPKB_EINVAL, "this workspace keeps the iterate history of the Gram path "
                              "(opkb_vmlmb_gram was used): the step-by-step recursion needs its own workspace");
    ws->hist = 0;
    ws->carry_valid = 0;
    ws->Gc_valid = 0;    // the pairs change outside opkb_vmlmb_gram
    const opkb_vector* gs = (ws->method == 1 ? ws->p : ws->g);
    int grid = opkb_grid_for(ctx, ws->n / (ws->eltype == OPKB_DOUBLE ? 2 : 4), 4);
    if (ws->eltype == OPKB_DOUBLE)
        LAUNCHB(ctx, OPKB_K_ELEMENTWISE, k_pair_update<double>, grid, OPKB_BLOCK, 0,
                (double*)ws->S[mark]->data, (double*)ws->Y[mark]->data, (const double*)ws->x->data,
                (const double*)gs->data, ws->n, scratch(ctx));
    else
        LAUNCHB(ctx, OPKB_K_ELEMENTWISE, k_pair_update<float>, grid, OPKB_BLOCK, 0,
                (float*)ws->S[mark]->data, (float*)ws->Y[mark]->data, (const float*)ws->x->data,
                (const float*)gs->data, ws->n, scratch(ctx));
    return opkb_finish_reduction(ctx, 2, NULL, out);
}

// ---------------------------------------------------------------------------
// P0: revert to the best point, :490-492
extern "C" int opkb_vmlmb_revert(opkb_vmlmb* ws, int mark, double stp)
{
    int rc = check_mark(ws, mark);
    if (rc) return rc;
    return (ws->eltype == OPKB_DOUBLE) ? trial_launch<double>(ws, mark, stp, 0, 1)
                                       : trial_launch<float>(ws, mark, stp, 0, 1);
}
