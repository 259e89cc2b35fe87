// opkb_spg.cu -- Tier 2: the SPG iteration (src/spg.jl:245-380) with the box
// projection as ONE pass per function evaluation.  x, g, x0, g0 never need the
// vcopy!'s of :322-323 / :344 (buffer rotation), pg / d / s / y are never
// stored (:224-226 aliases): they are recomputed in registers.
#include <math.h>
#include <string.h>

#include "opkb_internal.cuh"

struct opkb_spg_ {
    opkb_context* ctx;
    int eltype;
    int64_t n;
    const opkb_vector* lo;
    double lo_val;
    const opkb_vector* hi;
    double hi_val;
    int obj;
    const opkb_vector* qa;
    const opkb_vector* qc;
    opkb_vector* xb[4];
    opkb_vector* gb[2];
    int ix, ix0, ibest, ig, ig0;
};

template <typename T>
static Bound<T> make_bound(const opkb_vector* arr, double val, bool lower)
{
    Bound<T> b;
    b.arr = arr ? (const T*)arr->data : NULL;
    b.val = (T)val;
    b.has = arr ? 1 : (lower ? (b.val > -(T)INFINITY) : (b.val < (T)INFINITY));
    return b;
}

extern "C" int opkb_spg_create(opkb_context* ctx, int eltype, int64_t n, const opkb_vector* lo,
                               double lo_val, const opkb_vector* hi, double hi_val, opkb_spg** out)
{
    if (!ctx || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    *out = NULL;
    if (eltype != OPKB_FLOAT && eltype != OPKB_DOUBLE)
        return opkb_set_error(OPKB_EINVAL, "invalid element type");
    if ((lo && (lo->n != n || lo->eltype != eltype)) || (hi && (hi->n != n || hi->eltype != eltype)))
        return opkb_set_error(OPKB_EINVAL, "bounds of a different space");
    opkb_spg* ws = (opkb_spg*)calloc(1, sizeof(opkb_spg));
    if (!ws) return opkb_set_error(OPKB_ENOMEM, "out of host memory");
    ws->ctx = ctx;
    ws->eltype = eltype;
    ws->n = n;
    ws->lo = lo;
    ws->lo_val = lo_val;
    ws->hi = hi;
    ws->hi_val = hi_val;
    int rc = 0;
    for (int k = 0; k < 4 && !rc; ++k) rc = opkb_vector_create(ctx, eltype, n, &ws->xb[k]);
    for (int k = 0; k < 2 && !rc; ++k) rc = opkb_vector_create(ctx, eltype, n, &ws->gb[k]);
    if (rc) {
        opkb_spg_destroy(ws);
        return rc;
    }
    ws->ix = 0;
    ws->ix0 = 1;
    ws->ibest = 0;
    ws->ig = 0;
    ws->ig0 = 1;
    *out = ws;
    return 0;
}

extern "C" int opkb_spg_destroy(opkb_spg* ws)
{
    if (!ws) return 0;
    for (int k = 0; k < 4; ++k) opkb_vector_destroy(ws->xb[k]);
    for (int k = 0; k < 2; ++k) opkb_vector_destroy(ws->gb[k]);
    free(ws);
    return 0;
}

extern "C" int opkb_spg_set_objective(opkb_spg* ws, int kind, const opkb_vector* a, const opkb_vector* c)
{
    if (!ws) return opkb_set_error(OPKB_EINVAL, "NULL workspace");
    if (kind == OPKB_OBJ_QUADRATIC) {
        if (!a || !c || a->n != ws->n || c->n != ws->n || a->eltype != ws->eltype ||
            c->eltype != ws->eltype)
            return opkb_set_error(OPKB_EINVAL, "quadratic objective needs a, c of the same space");
    } else if (kind == OPKB_OBJ_ROSENBROCK) {
        if (ws->n % 2)
            return opkb_set_error(OPKB_EINVAL, "Rosenbrock needs an even number of (local) variables");
    } else {
        return opkb_set_error(OPKB_EINVAL, "the fused SPG path needs a built-in objective");
    }
    ws->obj = kind;
    ws->qa = a;
    ws->qc = c;
    return 0;
}

extern "C" opkb_vector* opkb_spg_vector(opkb_spg* ws, int which)
{
    if (!ws) return NULL;
    switch (which) {
    case 'x': return ws->xb[ws->ix];
    case '0': return ws->xb[ws->ix0];
    case 'b': return ws->xb[ws->ibest];
    case 'g': return ws->gb[ws->ig];
    case 'G': return ws->gb[ws->ig0];
    }
    return NULL;
}

// MODE 0: start (x = prj(x)); 1: step (x = prj(x0 - lambda g0));
//      2: backtrack (x = x0 + stp d, d = prj(x0 - lambda g0) - x0)
template <typename T> struct SpgArgs {
    const T* x0;
    const T* g0;
    T* x;
    T* g;
    const T* qa;
    const T* qc;
    Bound<T> lo, hi;
    T mlambda, stp, meta, ieta, mieta;  // -lambda, stp, -eta, 1/eta, -1/eta
    int eta_is_one;
    int64_t n;
};

template <typename T, int MODE>
__device__ __forceinline__ T spg_point(const SpgArgs<T>& A, T xin, T x0, T g0, T lo, T hi, T& d)
{
    if (MODE == 0) {
        d = T(0);
        return fastclamp(xin, lo, hi);  // prj!(x, x) :230
    }
    // prj!(x, vcombine!(x, 1, x0, -lambda, g0)) :327; d = x - x0 :329
    T xp = fastclamp(x0 + A.mlambda * g0, lo, hi);
    d = xp - x0;
    if (MODE == 1) return xp;
    return x0 + A.stp * d;  // vcombine!(x, 1, x0, stp, d) :368
}

template <typename T, int MODE>
__device__ __forceinline__ void spg_reduce(const SpgArgs<T>& A, T xi, T gi, T x0, T g0, T d, T lo, T hi,
                                           double (&acc)[8])
{
    // pg = (x - prj(x - eta*g))/eta :259
    T t = fastclamp(xi + A.meta * gi, lo, hi);
    T pg = A.eta_is_one ? (xi - t) : (A.ieta * xi + A.mieta * t);
    acc[3] = dacc(pg, pg, acc[3]);                       // vnorm2(pg) :261
    double apg = fabs((double)pg);
    acc[4] = (apg > acc[4] ? apg : acc[4]);              // vnorminf(pg) :262
    if (MODE != 0) {
        acc[2] = dacc(g0, d, acc[2]);                    // delta = vdot(g0, d) :330
        T s = xi - x0, y = gi - g0;                      // :309-310
        acc[5] = dacc(s, y, acc[5]);                     // sty :311
        acc[6] = dacc(s, s, acc[6]);                     // sts :314
    }
}

template <typename T, int OBJ, int MODE>
__global__ void __launch_bounds__(OPKB_BLOCK) k_spg(SpgArgs<T> A, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = A.n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // acc: 0,1 objective sums; 2 delta; 3 |pg|^2; 4 |pg|_inf; 5 s.y; 6 s.s; 7 unused
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        Pack<T> px0, pg0, px, pg, pd, pl = A.lo.pack(iv), ph = A.hi.pack(iv);
        if (MODE == 0) {
            px0 = ld_pack<T>(A.x, iv);
            pg0 = splat_pack<T>(T(0));
        } else {
            px0 = ld_pack<T>(A.x0, iv);
            pg0 = ld_pack<T>(A.g0, iv);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            px.v[e] = spg_point<T, MODE>(A, px0.v[e], px0.v[e], pg0.v[e], pl.v[e], ph.v[e], pd.v[e]);
        st_pack<T>(A.x, iv, px);
        if (OBJ == OPKB_OBJ_ROSENBROCK) {
#pragma unroll
            for (int e = 0; e < VEC; e += 2)
                rosenbrock_pair(px.v[e], px.v[e + 1], pg.v[e], pg.v[e + 1], acc[0], acc[1]);
        } else {
            Pack<T> pa = ld_pack<T>(A.qa, iv), pc = ld_pack<T>(A.qc, iv);
#pragma unroll
            for (int e = 0; e < VEC; ++e) quadratic_elem(px.v[e], pa.v[e], pc.v[e], pg.v[e], acc[0]);
        }
        st_pack<T>(A.g, iv, pg);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            spg_reduce<T, MODE>(A, px.v[e], pg.v[e], px0.v[e], pg0.v[e], pd.v[e], pl.v[e], ph.v[e], acc);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t t0 = nvec * VEC;
        for (int64_t i = t0; i < A.n; ++i) {
            T x0 = (MODE == 0) ? A.x[i] : A.x0[i], g0 = (MODE == 0) ? T(0) : A.g0[i], d;
            A.x[i] = spg_point<T, MODE>(A, x0, x0, g0, A.lo.at(i), A.hi.at(i), d);
        }
        if (OBJ == OPKB_OBJ_ROSENBROCK) {
            for (int64_t i = t0; i + 1 < A.n; i += 2)
                rosenbrock_pair(A.x[i], A.x[i + 1], A.g[i], A.g[i + 1], acc[0], acc[1]);
        } else {
            for (int64_t i = t0; i < A.n; ++i) quadratic_elem(A.x[i], A.qa[i], A.qc[i], A.g[i], acc[0]);
        }
        for (int64_t i = t0; i < A.n; ++i) {
            // MODE 0 never reads x0/g0/d in spg_reduce
            T x0 = (MODE == 0) ? T(0) : A.x0[i], g0 = (MODE == 0) ? T(0) : A.g0[i], d = T(0);
            if (MODE != 0) {
                T xp = fastclamp(x0 + A.mlambda * g0, A.lo.at(i), A.hi.at(i));
                d = xp - x0;
            }
            spg_reduce<T, MODE>(A, A.x[i], A.g[i], x0, g0, d, A.lo.at(i), A.hi.at(i), acc);
        }
    }
    block_reduce_publish<8, 16u, 0u>(acc, rs);  // slot 4: max
}

template <typename T>
static int spg_run(opkb_spg* ws, int mode, double lambda, double stp, double eta, double out[8])
{
    opkb_context* ctx = ws->ctx;
    SpgArgs<T> A;
    A.x0 = (const T*)ws->xb[ws->ix0]->data;
    A.g0 = (const T*)ws->gb[ws->ig0]->data;
    A.x = (T*)ws->xb[ws->ix]->data;
    A.g = (T*)ws->gb[ws->ig]->data;
    A.qa = ws->qa ? (const T*)ws->qa->data : NULL;
    A.qc = ws->qc ? (const T*)ws->qc->data : NULL;
    A.lo = make_bound<T>(ws->lo, ws->lo_val, true);
    A.hi = make_bound<T>(ws->hi, ws->hi_val, false);
    A.mlambda = (T)(-lambda);
    A.stp = (T)stp;
    A.meta = (T)(-eta);
    A.ieta = (T)(1 / eta);
    A.mieta = (T)(-1 / eta);
    A.eta_is_one = (eta == 1.0);
    A.n = ws->n;
    int grid = opkb_grid_for(ctx, ws->n / VecOf<T>::N, 4);
    ReduceScratch rs = opkb_make_scratch(ctx);
#define SPG_LAUNCH(OBJ, MODE)                                                     \
    do {                                                                          \
        opkb_prof_begin(ctx, OPKB_K_SPG);                                         \
        k_spg<T, OBJ, MODE><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);         \
        opkb_prof_end(ctx);                                                       \
        ctx->launches += 1;                                                       \
        OPKB_CUDA(cudaGetLastError());                                            \
    } while (0)
    if (ws->obj == OPKB_OBJ_ROSENBROCK) {
        if (mode == 0) SPG_LAUNCH(OPKB_OBJ_ROSENBROCK, 0);
        else if (mode == 1) SPG_LAUNCH(OPKB_OBJ_ROSENBROCK, 1);
        else SPG_LAUNCH(OPKB_OBJ_ROSENBROCK, 2);
    } else {
        if (mode == 0) SPG_LAUNCH(OPKB_OBJ_QUADRATIC, 0);
        else if (mode == 1) SPG_LAUNCH(OPKB_OBJ_QUADRATIC, 1);
        else SPG_LAUNCH(OPKB_OBJ_QUADRATIC, 2);
    }
#undef SPG_LAUNCH
    int kinds[8] = {OPKB_RSUM, OPKB_RSUM, OPKB_RSUM, OPKB_RSUM, OPKB_RMAX, OPKB_RSUM, OPKB_RSUM, OPKB_RSUM};
    double r[8];
    int rc = opkb_finish_reduction(ctx, 8, kinds, r);
    if (rc) return rc;
    out[0] = (ws->obj == OPKB_OBJ_ROSENBROCK) ? r[0] + r[1] : 0.5 * r[0];
    out[1] = r[2];
    out[2] = r[3];
    out[3] = r[4];
    out[4] = r[5];
    out[5] = r[6];
    out[6] = 0;
    out[7] = 0;
    return 0;
}

static int spg_check(opkb_spg* ws, const double* out)
{
    if (!ws || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (ws->obj != OPKB_OBJ_ROSENBROCK && ws->obj != OPKB_OBJ_QUADRATIC)
        return opkb_set_error(OPKB_EINVAL, "no objective set");
    return 0;
}

static int free_x_buffer(const opkb_spg* ws)
{
    for (int k = 0; k < 4; ++k)
        if (k != ws->ix0 && k != ws->ibest && k != ws->ix) return k;
    return -1;
}

extern "C" int opkb_spg_start(opkb_spg* ws, double eta, double out[8])
{
    int rc = spg_check(ws, out);
    if (rc) return rc;
    if (!(eta > 0)) return opkb_set_error(OPKB_EINVAL, "eta must be positive");
    rc = (ws->eltype == OPKB_DOUBLE) ? spg_run<double>(ws, 0, 0, 0, eta, out)
                                     : spg_run<float>(ws, 0, 0, 0, eta, out);
    ws->ibest = ws->ix;  // vcopy!(xbest, x) :241
    return rc;
}

extern "C" int opkb_spg_step(opkb_spg* ws, double lambda, double eta, double out[8])
{
    int rc = spg_check(ws, out);
    if (rc) return rc;
    // vcopy!(x0, x); vcopy!(g0, g) :322-323 by rotation
    ws->ix0 = ws->ix;
    int nx = -1;
    for (int k = 0; k < 4; ++k)
        if (k != ws->ix0 && k != ws->ibest) { nx = k; break; }
    ws->ix = nx;
    int t = ws->ig0;
    ws->ig0 = ws->ig;
    ws->ig = t;
    return (ws->eltype == OPKB_DOUBLE) ? spg_run<double>(ws, 1, lambda, 1.0, eta, out)
                                       : spg_run<float>(ws, 1, lambda, 1.0, eta, out);
}

extern "C" int opkb_spg_backtrack(opkb_spg* ws, double lambda, double stp, double eta, double out[8])
{
    int rc = spg_check(ws, out);
    if (rc) return rc;
    if (ws->ix == ws->ibest) {  // the rejected trial is the best point so far: keep it
        int nx = free_x_buffer(ws);
        if (nx < 0) return opkb_set_error(OPKB_EINVAL, "no free buffer");
        ws->ix = nx;
    }
    return (ws->eltype == OPKB_DOUBLE) ? spg_run<double>(ws, 2, lambda, stp, eta, out)
                                       : spg_run<float>(ws, 2, lambda, stp, eta, out);
}

extern "C" int opkb_spg_mark_best(opkb_spg* ws)
{
    if (!ws) return opkb_set_error(OPKB_EINVAL, "NULL workspace");
    ws->ibest = ws->ix;  // vcopy!(xbest, x) :344 without a copy
    return 0;
}

// ---------------------------------------------------------------------------
// Tier 3: the loop of `_spg!` (src/spg.jl:245-380) inside the library, for the built-in
// objectives and the box projection: one call runs `niter` iterations (or to termination),
// the host only reads the counters.  Same stage logic as the Python / Julia hosts, line by
// line (spg.py: SPG.iterate); all scalars double.
struct opkb_spg_solver_ {
    opkb_spg* ws;
    int m;
    int64_t maxit, maxfc;
    double eps1, eps2, eta;
    double lmin, lmax, ftol, amin, amax;      // :201-205
    int64_t iter, fcnt, pcnt;
    int pg_valid;          // pgtwon / pginfn / pcnt already belong to the current iterate
    int status, started, finished;
    double f, f0, fbest, pgtwon, pginfn, lambda;
    double* lastfv;
    double out[8];
};

static double jl_clamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

extern "C" int opkb_spg_solver_create(opkb_spg* ws, int m, int64_t maxit, int64_t maxfc, double eps1,
                                      double eps2, double eta, opkb_spg_solver** out)
{
    if (!ws || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    *out = NULL;
    if (!(m >= 1 && eps1 >= 0 && eps2 >= 0 && eta > 0))                       // :197-200
        return opkb_set_error(OPKB_EINVAL, "spg: m >= 1, eps1 >= 0, eps2 >= 0, eta > 0 are required");
    if (ws->obj != OPKB_OBJ_ROSENBROCK && ws->obj != OPKB_OBJ_QUADRATIC)
        return opkb_set_error(OPKB_EINVAL, "the native SPG loop needs a built-in objective");
    opkb_spg_solver* s = (opkb_spg_solver*)calloc(1, sizeof(opkb_spg_solver));
    if (!s) return opkb_set_error(OPKB_ENOMEM, "out of host memory");
    s->lastfv = (double*)malloc(sizeof(double) * (size_t)m);
    if (!s->lastfv) {
        free(s);
        return opkb_set_error(OPKB_ENOMEM, "out of host memory");
    }
    for (int i = 0; i < m; ++i) s->lastfv[i] = -INFINITY;                     // :222
    s->ws = ws;
    s->m = m;
    s->maxit = maxit;
    s->maxfc = maxfc;
    s->eps1 = eps1;
    s->eps2 = eps2;
    s->eta = eta;
    s->lmin = 1e-30;
    s->lmax = 1e+30;
    s->ftol = 1e-4;
    s->amin = 0.1;
    s->amax = 0.9;
    s->pgtwon = s->pginfn = INFINITY;
    *out = s;
    return 0;
}

extern "C" int opkb_spg_solver_destroy(opkb_spg_solver* s)
{
    if (!s) return 0;
    free(s->lastfv);
    free(s);
    return 0;
}

// one trip of the main loop; returns 1 while searching, 0 once finished, < 0 on error
static int spg_solver_iterate(opkb_spg_solver* s)
{
    opkb_spg* ws = s->ws;
    int rc;
    if (s->finished) return 0;
    if (!s->started) {                                                        // :230-241
        rc = opkb_spg_start(ws, s->eta, s->out);
        if (rc) return rc;
        s->f = s->out[0];
        s->pcnt += 1;
        s->fcnt += 1;
        s->fbest = s->f;
        s->started = 1;
    }
    double fmin, fmax;
    int fconst = 0;
    if (s->m > 1) {                                                           // :249-255
        s->lastfv[s->iter % s->m] = s->f;
        fmin = fmax = s->lastfv[0];
        for (int i = 1; i < s->m; ++i) {
            const double v = s->lastfv[i];       // Julia's min / max propagate NaN
            fmin = (fmin != fmin) ? fmin : ((v != v) ? v : (v < fmin ? v : fmin));
            fmax = (fmax != fmax) ? fmax : ((v != v) ? v : (v > fmax ? v : fmax));
        }
        fconst = !(fmin < fmax);
    } else {
        fmin = fmax = s->f;
    }
    if (!s->pg_valid) {
        s->pgtwon = sqrt(s->out[2]);                                          // :259-262
        s->pginfn = s->out[3];
        s->pcnt += 1;
    }
    s->pg_valid = 0;
    // :278-302
    if (s->pginfn <= s->eps1) { s->status = 1; s->finished = 1; return 0; }
    if (s->pgtwon <= s->eps2) { s->status = 2; s->finished = 1; return 0; }
    if (fconst) { s->status = 3; s->finished = 1; return 0; }
    if (s->iter >= s->maxit) { s->status = -1; s->finished = 1; return 0; }
    if (s->fcnt >= s->maxfc) { s->status = -2; s->finished = 1; return 0; }
    double lam;                                                               // :305-319
    if (s->iter == 0) {
        lam = jl_clamp(1 / s->pginfn, s->lmin, s->lmax);
    } else {
        const double sty = s->out[4], sts = s->out[5];
        lam = (sty > 0) ? jl_clamp(sts / sty, s->lmin, s->lmax) : s->lmax;
    }
    s->f0 = s->f;                                                             // :324
    s->lambda = lam;
    rc = opkb_spg_step(ws, lam, s->eta, s->out);                              // :327-336
    if (rc) return rc;
    s->f = s->out[0];
    const double delta = s->out[1];
    s->pcnt += 1;
    double stp = 1.0;                                                         // :333
    for (;;) {
        s->fcnt += 1;                                                         // :337
        if (s->f < s->fbest) {                                                // :342-345
            s->fbest = s->f;
            opkb_spg_mark_best(ws);
        }
        if (s->f <= fmax + stp * s->ftol * delta) break;                      // :348
        if (s->fcnt >= s->maxfc) {                                            // :352-356
            s->status = -2;
            break;
        }
        const double q = -delta * (stp * stp);                                // :359-365
        const double r = (s->f - s->f0 - stp * delta) * 2;
        if (r > 0 && s->amin * r <= q && q <= s->amax * stp * r) stp = q / r;
        else stp /= 2;
        rc = opkb_spg_backtrack(ws, lam, stp, s->eta, s->out);                // :368, :336
        if (rc) return rc;
        s->f = s->out[0];
    }
    if (s->status != 0) {                                                     // :371-375
        s->finished = 1;
        return 0;
    }
    s->iter += 1;                                                             // :378
    // the projected gradient of the new iterate (:259-262 of the next trip) came out of the same
    // pass: what `opkb_spg_solver_info` reports at an iteration boundary is then the state the
    // reference's `printer` sees (:265-275) -- f, both norms and `pcnt` of the SAME point
    s->pgtwon = sqrt(s->out[2]);
    s->pginfn = s->out[3];
    s->pcnt += 1;
    s->pg_valid = 1;
    return 1;
}

extern "C" int opkb_spg_solver_run(opkb_spg_solver* s, int64_t niter, int* searching)
{
    if (!s) return opkb_set_error(OPKB_EINVAL, "NULL solver");
    int alive = s->finished ? 0 : 1;
    const int64_t target = (niter > INT64_MAX - s->iter) ? INT64_MAX : s->iter + niter;
    while (alive == 1 && s->iter < target) {
        alive = spg_solver_iterate(s);
        if (alive < 0) return alive;
    }
    if (searching) *searching = alive;
    return 0;
}

extern "C" int opkb_spg_solver_info(const opkb_spg_solver* s, double* f, double* fbest, double* pginfn,
                                    double* pgtwon, int64_t* iter, int64_t* fcnt, int64_t* pcnt, int* status,
                                    int* best_is_x)
{
    if (!s) return opkb_set_error(OPKB_EINVAL, "NULL solver");
    if (f) *f = s->f;
    if (fbest) *fbest = s->fbest;
    if (pginfn) *pginfn = s->pginfn;
    if (pgtwon) *pgtwon = s->pgtwon;
    if (iter) *iter = s->iter;
    if (fcnt) *fcnt = s->fcnt;
    if (pcnt) *pcnt = s->pcnt;
    if (status) *status = s->status;
    if (best_is_x) *best_is_x = !(s->fbest < s->f);       // :401: the solution is 'b' when fbest < f
    return 0;
}
