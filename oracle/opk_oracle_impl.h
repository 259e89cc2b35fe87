/*
 * opk_oracle_impl.h -- type-generic body of the CPU oracle (TEST
 * INFRASTRUCTURE ONLY; PARITY UNPINNED -- see opk_oracle.h).  Included twice
 * by opk_oracle.c with REAL = double / float and FN(x) = opko_d_x / opko_f_x.
 *
 * One memory pass per reference call ("reference-faithful unfused" mode): the
 * pass counts of SURVEY.md 8(a) hold for this code, which is why it also
 * serves as the CPU baseline of bench.py.
 */

#define SUM_BLOCK 128     /* leaf of the pairwise sums */
#define SUM_CHUNK 8192    /* unit of OpenMP work; fixed => deterministic */

/* ---- blocked pairwise sums --------------------------------------------- */
/* kind 0: x[i]*y[i]; kind 1: via index list sel.  Leaves are sequential in
 * the accumulation type, halves are combined pairwise (like Julia's
 * mapreduce), chunks are combined pairwise in index order, so the result
 * does not depend on the number of threads. */

static double FN(leaf_acc)(const REAL* x, const REAL* y, const int64_t* sel,
                           int64_t i0, int64_t i1)
{
    double s = 0.0;
    if (sel) {
        for (int64_t j = i0; j < i1; ++j) {
            int64_t i = sel[j];
            s += (double)x[i] * (double)y[i];
        }
    } else {
        for (int64_t i = i0; i < i1; ++i) s += (double)x[i] * (double)y[i];
    }
    return s;
}

static REAL FN(leaf_nat)(const REAL* x, const REAL* y, const int64_t* sel,
                         int64_t i0, int64_t i1)
{
    REAL s = 0;
    if (sel) {
        for (int64_t j = i0; j < i1; ++j) {
            int64_t i = sel[j];
            s += x[i] * y[i];
        }
    } else {
        for (int64_t i = i0; i < i1; ++i) s += x[i] * y[i];
    }
    return s;
}

static double FN(pair_acc)(const REAL* x, const REAL* y, const int64_t* sel,
                           int64_t i0, int64_t i1)
{
    if (i1 - i0 <= SUM_BLOCK) return FN(leaf_acc)(x, y, sel, i0, i1);
    int64_t mid = i0 + ((i1 - i0) >> 1);
    return FN(pair_acc)(x, y, sel, i0, mid) + FN(pair_acc)(x, y, sel, mid, i1);
}

static REAL FN(pair_nat)(const REAL* x, const REAL* y, const int64_t* sel,
                         int64_t i0, int64_t i1)
{
    if (i1 - i0 <= SUM_BLOCK) return FN(leaf_nat)(x, y, sel, i0, i1);
    int64_t mid = i0 + ((i1 - i0) >> 1);
    return FN(pair_nat)(x, y, sel, i0, mid) + FN(pair_nat)(x, y, sel, mid, i1);
}

static double FN(pair_dbl)(const double* v, int64_t i0, int64_t i1)
{
    if (i1 - i0 <= 8) {
        double s = 0.0;
        for (int64_t i = i0; i < i1; ++i) s += v[i];
        return s;
    }
    int64_t mid = i0 + ((i1 - i0) >> 1);
    return FN(pair_dbl)(v, i0, mid) + FN(pair_dbl)(v, mid, i1);
}

/* sum of x[i]*y[i] over i (sel == NULL) or over i in sel[0..n) */
static double FN(sum_prod)(int64_t n, const int64_t* sel, const REAL* x,
                           const REAL* y)
{
    if (g_sum_mode == OPKO_SUM_SEQUENTIAL)
        return (double)FN(leaf_nat)(x, y, sel, 0, n);
    if (g_sum_mode == OPKO_SUM_PAIRWISE)
        return (double)FN(pair_nat)(x, y, sel, 0, n);
    if (n <= SUM_CHUNK) return FN(pair_acc)(x, y, sel, 0, n);
    int64_t nchunks = (n + SUM_CHUNK - 1) / SUM_CHUNK;
    double* part = (double*)malloc((size_t)nchunks * sizeof(double));
    if (!part) return NAN;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nchunks; ++c) {
        int64_t i0 = c * SUM_CHUNK;
        int64_t i1 = (i0 + SUM_CHUNK < n ? i0 + SUM_CHUNK : n);
        part[c] = FN(pair_acc)(x, y, sel, i0, i1);
    }
    double s = FN(pair_dbl)(part, 0, nchunks);
    free(part);
    return s;
}

/* ---- LazyAlgebra vector operations ------------------------------------- */
/* Call sites: src/quasinewton.jl:45-69 (vdot/vnorm2 shims returning Float64)
 * and the list in SURVEY.md 8(c).  Semantics: doc/algebra.md:84-116,
 * test/vectors-tests.jl:15-81 (multipliers 0 and +-1 are special, other
 * multipliers are converted to the element type first). */

double FN(vdot)(int64_t n, const REAL* x, const REAL* y)
{
    return FN(sum_prod)(n, NULL, x, y);
}

double FN(vdot_sel)(int64_t nsel, const int64_t* sel, const REAL* x,
                    const REAL* y)
{
    /* vdot(sel, x, y): src/quasinewton.jl:55-65 */
    return FN(sum_prod)(nsel, sel, x, y);
}

double FN(vnorm2)(int64_t n, const REAL* x)
{
    /* src/quasinewton.jl:68-69 */
    double s = FN(sum_prod)(n, NULL, x, x);
    if (g_sum_mode == OPKO_SUM_ACCURATE) return sqrt(s);
#if REAL_IS_DOUBLE
    return sqrt(s);
#else
    return (double)sqrtf((float)s);
#endif
}

double FN(vnorminf)(int64_t n, const REAL* x)
{
    /* vnorminf: src/spg.jl:262 */
    REAL m = 0;
#pragma omp parallel for schedule(static) reduction(max : m)
    for (int64_t i = 0; i < n; ++i) {
        REAL a = (x[i] < 0 ? -x[i] : x[i]);
        if (a > m) m = a;
    }
    return (double)m;
}

void FN(vcopy)(int64_t n, REAL* dst, const REAL* src)
{
    if (dst == src) return;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) dst[i] = src[i];
}

void FN(vscale)(int64_t n, REAL* x, double alpha)
{
    /* vscale!(v, gamma): src/quasinewton.jl:707, :768 */
    if (alpha == 1) return;
    if (alpha == 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) x[i] = 0;
    } else if (alpha == -1) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) x[i] = -x[i];
    } else {
        const REAL a = (REAL)alpha;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) x[i] = a * x[i];
    }
}

void FN(vupdate)(int64_t n, REAL* y, double alpha, const REAL* x)
{
    /* vupdate!(y, alpha, x): y += alpha*x, src/quasinewton.jl:512-513, :729 */
    if (alpha == 0) return;
    if (alpha == 1) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) y[i] = y[i] + x[i];
    } else if (alpha == -1) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) y[i] = y[i] - x[i];
    } else {
        const REAL a = (REAL)alpha;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) y[i] = y[i] + a * x[i];
    }
}

void FN(vupdate_sel)(REAL* y, int64_t nsel, const int64_t* sel, double alpha,
                     const REAL* x)
{
    /* vupdate!(y, sel, alpha, x): src/quasinewton.jl:761, :772 */
    if (alpha == 0) return;
    const REAL a = (REAL)alpha;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < nsel; ++j) {
        int64_t i = sel[j];
        y[i] = y[i] + a * x[i];
    }
}

void FN(vcombine)(int64_t n, REAL* dst, double alpha, const REAL* x,
                  double beta, const REAL* y)
{
    /* vcombine!(dst, alpha, x, beta, y): src/quasinewton.jl:427, :490, :599;
     * src/spg.jl:259, :309-310, :327-329, :368 */
    const REAL a = (REAL)alpha, b = (REAL)beta;
    if (alpha == 0 && beta == 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) dst[i] = 0;
    } else if (alpha == 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) dst[i] = b * y[i];
    } else if (beta == 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) dst[i] = a * x[i];
    } else {
        /* with a or b equal to +-1 the product is exact, so this also equals
         * x + b*y, x - y, ... bit for bit */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) dst[i] = a * x[i] + b * y[i];
    }
}

void FN(vproduct)(int64_t n, REAL* dst, const REAL* x, const REAL* y)
{
    /* vproduct!(dst, x, y): src/quasinewton.jl:713 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) dst[i] = x[i] * y[i];
}

/* ---- SimpleBounds (src/bounds.jl) --------------------------------------- */

#define REAL_INF ((REAL)INFINITY)

void FN(project_variables)(int64_t n, REAL* dst, const REAL* src,
                           const REAL* lo, double lo_val, const REAL* hi,
                           double hi_val)
{
    /* src/bounds.jl:137-213; fastclamp :65-67 = fastmax(fastmin(x, hi), lo)
     * with fastmin/fastmax :41, :53 (NaN in x is kept).  An absent bound is
     * +-Inf, for which the comparison is always false, i.e. the `nothing`
     * variants. */
    const REAL slo = (REAL)lo_val, shi = (REAL)hi_val;
    if (!lo && !hi && !(slo > -REAL_INF) && !(shi < REAL_INF)) {
        FN(vcopy)(n, dst, src); /* :165-167 */
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        REAL l = (lo ? lo[i] : slo), h = (hi ? hi[i] : shi);
        REAL t = src[i];
        t = (t > h ? h : t);
        t = (t < l ? l : t);
        dst[i] = t;
    }
}

int FN(project_direction)(int64_t n, REAL* dst, const REAL* x, const REAL* lo,
                          double lo_val, const REAL* hi, double hi_val,
                          int orient, const REAL* d)
{
    /* src/bounds.jl:321-409; projdir :306-311; can_change :638-648.
     * orient > 0 is `+` (Forward), orient < 0 is `-` (Backward). */
    const REAL slo = (REAL)lo_val, shi = (REAL)hi_val;
    if (orient == 0) return -1; /* "invalid orientation" :223-225 */
    /* a scalar bound only counts when finite (:340-341, :367, :386) */
    const int has_lo = (lo != NULL) || (slo > -REAL_INF);
    const int has_hi = (hi != NULL) || (shi < REAL_INF);
    if (!lo && !hi && !(slo <= shi)) return -2; /* "invalid bounds" :339 */
    if (!has_lo && !has_hi) {
        FN(vcopy)(n, dst, d); /* :354-356 */
        return 0;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        REAL l = (lo ? lo[i] : slo), h = (hi ? hi[i] : shi);
        REAL di = d[i], xi = x[i];
        int neg = (orient > 0 ? di < 0 : di > 0); /* is_negative(o, d) */
        int pos = (orient > 0 ? di > 0 : di < 0); /* is_positive(o, d) */
        int ok = (neg & (!has_lo | (xi > l))) | (pos & (!has_hi | (xi < h)));
        dst[i] = (ok ? di : (REAL)0);
    }
    return 0;
}

int FN(step_limits)(int64_t n, const REAL* x, const REAL* lo, double lo_val,
                    const REAL* hi, double hi_val, int orient, const REAL* d,
                    double* smin_out, double* smax_out)
{
    /* src/bounds.jl:431-624.  The scalar-lo/array-hi method is misnamed in
     * the reference (:555); its intended semantics is implemented. */
    const REAL slo = (REAL)lo_val, shi = (REAL)hi_val;
    if (orient == 0) return -1;
    const REAL s = (orient > 0 ? (REAL)1 : (REAL)-1);
    if (!lo && !hi && !(slo <= shi)) return -2; /* :450 */
    const int has_lo = (lo != NULL) || (slo > -REAL_INF);
    const int has_hi = (hi != NULL) || (shi < REAL_INF);
    REAL smin = REAL_INF, smax = 0;
    if (!has_lo && !has_hi) {
        smax = REAL_INF; /* :503-505 */
    } else {
        int unbounded = 0;
#pragma omp parallel for schedule(static) reduction(min : smin) \
    reduction(max : smax) reduction(| : unbounded)
        for (int64_t i = 0; i < n; ++i) {
            REAL p = s * d[i];
            if (p > 0) {
                if (has_hi) {
                    REAL a = ((hi ? hi[i] : shi) - x[i]) / p;
                    if (0 < a && a < smin) smin = a;
                    if (a > smax) smax = a;
                } else {
                    unbounded = 1; /* :475-476, :538-539 */
                }
            } else if (p < 0) {
                if (has_lo) {
                    REAL a = ((lo ? lo[i] : slo) - x[i]) / p;
                    if (0 < a && a < smin) smin = a;
                    if (a > smax) smax = a;
                } else {
                    unbounded = 1; /* :499-500, :591-592 */
                }
            }
        }
        if (unbounded) smax = REAL_INF;
    }
    *smin_out = (double)smin;
    *smax_out = (double)smax;
    return 0;
}

int64_t FN(get_free_variables)(int64_t n, int64_t* sel, const REAL* gp)
{
    /* get_free_variables!(sel, gp): src/bounds.jl:701-714 (0-based here) */
    int64_t j = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (gp[i] != 0) sel[j++] = i;
    }
    return j;
}

int64_t FN(get_free_variables_dir)(int64_t n, int64_t* sel, const REAL* x,
                                   const REAL* lo, double lo_val,
                                   const REAL* hi, double hi_val, int orient,
                                   const REAL* d)
{
    /* get_free_variables!(sel, x, lo, hi, orient, d): src/bounds.jl:716-844,
     * can_change :638-648 (0-based indices); -1 / -2 on invalid arguments */
    const REAL slo = (REAL)lo_val, shi = (REAL)hi_val;
    if (orient == 0) return -1;
    if (!lo && !hi && !(slo <= shi)) return -2; /* :735 */
    const int has_lo = (lo != NULL) || (slo > -REAL_INF);
    const int has_hi = (hi != NULL) || (shi < REAL_INF);
    int64_t j = 0;
    for (int64_t i = 0; i < n; ++i) {
        REAL l = (lo ? lo[i] : slo), h = (hi ? hi[i] : shi);
        REAL di = d[i], xi = x[i];
        int neg = (orient > 0 ? di < 0 : di > 0);
        int pos = (orient > 0 ? di > 0 : di < 0);
        int ok;
        if (!has_lo && !has_hi)
            ok = (di != 0); /* is_non_zero :638-639; the scalar/scalar method lists all (:763-766) */
        else
            ok = (neg & (!has_lo | (xi > l))) | (pos & (!has_hi | (xi < h)));
        if (!lo && !hi && !has_lo && !has_hi) ok = 1; /* :762-767: sel = 1:n */
        if (ok) sel[j++] = i;
    }
    return j;
}

/* ---- apply_lbfgs! (src/quasinewton.jl:703-779) --------------------------- */

int FN(apply_lbfgs)(int64_t n, int64_t mem, REAL* const* S, REAL* const* Y,
                    const double* rho, double gamma, const REAL* diag,
                    int64_t m, int64_t mark, REAL* v, double* alpha)
{
    /* :716-745 with H0 = gamma*I (:703-708) or diag (:710-714).
     * Slots are 0-based: k runs mark, mark-1, ... (mod mem). */
    int modif = 0;
    int64_t k = mark + 1;
    for (int64_t i = 0; i < m; ++i) {
        k = (k > 0 ? k - 1 : mem - 1);
        if (rho[k] > 0) {
            alpha[k] = FN(vdot)(n, S[k], v) / rho[k];
            FN(vupdate)(n, v, -alpha[k], Y[k]);
            modif = 1;
        }
    }
    if (modif) {
        if (diag)
            FN(vproduct)(n, v, diag, v);
        else
            FN(vscale)(n, v, gamma);
        for (int64_t i = 0; i < m; ++i) {
            if (rho[k] > 0) {
                double beta = FN(vdot)(n, Y[k], v) / rho[k];
                FN(vupdate)(n, v, alpha[k] - beta, S[k]);
            }
            k = (k < mem - 1 ? k + 1 : 0);
        }
    }
    return modif;
}

int FN(apply_lbfgs_sel)(int64_t n, int64_t mem, REAL* const* S,
                        REAL* const* Y, double* rho, int64_t m, int64_t mark,
                        REAL* v, double* alpha, int64_t nsel,
                        const int64_t* sel)
{
    /* :747-779, subspace version: rho and gamma recomputed on `sel` */
    double gamma = 0;
    int64_t k = mark + 1;
    for (int64_t i = 0; i < m; ++i) {
        k = (k > 0 ? k - 1 : mem - 1);
        rho[k] = FN(vdot_sel)(nsel, sel, Y[k], S[k]);
        if (rho[k] > 0) {
            alpha[k] = FN(vdot_sel)(nsel, sel, S[k], v) / rho[k];
            FN(vupdate_sel)(v, nsel, sel, -alpha[k], Y[k]);
            if (gamma == 0)
                gamma = rho[k] / FN(vdot_sel)(nsel, sel, Y[k], Y[k]);
        }
    }
    if (gamma != 0) {
        FN(vscale)(n, v, gamma);
        for (int64_t i = 0; i < m; ++i) {
            if (rho[k] > 0) {
                double beta = FN(vdot_sel)(nsel, sel, Y[k], v) / rho[k];
                FN(vupdate_sel)(v, nsel, sel, alpha[k] - beta, S[k]);
            }
            k = (k < mem - 1 ? k + 1 : 0);
        }
    }
    return (gamma != 0);
}

/* ---- objectives --------------------------------------------------------- */

void FN(rosenbrock_init)(int64_t n, REAL* x0)
{
    /* test/rosenbrock.jl:10-14 */
    for (int64_t i = 0; i < n; ++i) x0[i] = (i % 2 == 0 ? (REAL)-1.2 : (REAL)1.0);
}

double FN(rosenbrock_fg)(void* ctx, int64_t n, const REAL* x, REAL* g)
{
    /* test/rosenbrock.jl:16-29.  The reference sums t1.*t1 and t2.*t2
     * separately in the element type; in ACCURATE mode the two sums are
     * accumulated in double (terms still rounded to the element type). */
    (void)ctx;
    const int64_t np = n / 2;
    const REAL c1 = 1, c2 = 2, c10 = 10, c200 = 200;
    double f1 = 0, f2 = 0;
    if (g_sum_mode != OPKO_SUM_ACCURATE) {
        REAL s1 = 0, s2 = 0;
        for (int64_t j = 0; j < np; ++j) {
            REAL x1 = x[2 * j], x2 = x[2 * j + 1];
            REAL t1 = c1 - x1;
            REAL u = x2 - x1 * x1;
            REAL t2 = c10 * u;
            REAL g2 = c200 * u;
            g[2 * j] = -c2 * (x1 * g2 + t1);
            g[2 * j + 1] = g2;
            s1 += t1 * t1;
            s2 += t2 * t2;
        }
        return (double)(REAL)(s1 + s2);
    }
    int64_t nchunks = (np + SUM_CHUNK - 1) / SUM_CHUNK;
    double* part = (double*)malloc((size_t)(2 * nchunks + 2) * sizeof(double));
    if (!part) return NAN;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nchunks; ++c) {
        int64_t j0 = c * SUM_CHUNK;
        int64_t j1 = (j0 + SUM_CHUNK < np ? j0 + SUM_CHUNK : np);
        double a1 = 0, a2 = 0;
        for (int64_t b0 = j0; b0 < j1; b0 += SUM_BLOCK) {
            int64_t b1 = (b0 + SUM_BLOCK < j1 ? b0 + SUM_BLOCK : j1);
            double s1 = 0, s2 = 0;
            for (int64_t j = b0; j < b1; ++j) {
                REAL x1 = x[2 * j], x2 = x[2 * j + 1];
                REAL t1 = c1 - x1;
                REAL u = x2 - x1 * x1;
                REAL t2 = c10 * u;
                REAL g2 = c200 * u;
                g[2 * j] = -c2 * (x1 * g2 + t1);
                g[2 * j + 1] = g2;
                s1 += (double)(REAL)(t1 * t1);
                s2 += (double)(REAL)(t2 * t2);
            }
            a1 += s1;
            a2 += s2;
        }
        part[c] = a1;
        part[nchunks + c] = a2;
    }
    f1 = FN(pair_dbl)(part, 0, nchunks);
    f2 = FN(pair_dbl)(part + nchunks, 0, nchunks);
    free(part);
    return f1 + f2;
}

double FN(quadratic_fg)(void* ctx, int64_t n, const REAL* x, REAL* g)
{
    /* Synthetic objective of BASELINE.json (not in the reference):
     * f = 1/2 sum a_i (x_i - c_i)^2, g_i = a_i (x_i - c_i).  Elementwise in
     * the element type: r = x - c; g = a*r; term = g*r; f = 0.5*sum(term). */
    const FN(quadratic)* q = (const FN(quadratic)*)ctx;
    const REAL* a = q->a;
    const REAL* c = q->c;
    if (g_sum_mode != OPKO_SUM_ACCURATE) {
        REAL s = 0;
        for (int64_t i = 0; i < n; ++i) {
            REAL r = x[i] - c[i];
            REAL gi = a[i] * r;
            g[i] = gi;
            s += gi * r;
        }
        return 0.5 * (double)s;
    }
    int64_t nchunks = (n + SUM_CHUNK - 1) / SUM_CHUNK;
    double* part = (double*)malloc((size_t)(nchunks + 1) * sizeof(double));
    if (!part) return NAN;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nchunks; ++k) {
        int64_t i0 = k * SUM_CHUNK;
        int64_t i1 = (i0 + SUM_CHUNK < n ? i0 + SUM_CHUNK : n);
        double acc = 0;
        for (int64_t b0 = i0; b0 < i1; b0 += SUM_BLOCK) {
            int64_t b1 = (b0 + SUM_BLOCK < i1 ? b0 + SUM_BLOCK : i1);
            double s = 0;
            for (int64_t i = b0; i < b1; ++i) {
                REAL r = x[i] - c[i];
                REAL gi = a[i] * r;
                g[i] = gi;
                s += (double)(REAL)(gi * r);
            }
            acc += s;
        }
        part[k] = acc;
    }
    double f = FN(pair_dbl)(part, 0, nchunks);
    free(part);
    return 0.5 * f;
}

void FN(box_prj)(void* ctx, int64_t n, REAL* dst, const REAL* src)
{
    /* the box `prj!` of SPG; for lo = 0, hi = +Inf this is `nonnegative!`
     * (test/rosenbrock.jl:50-58) */
    const FN(box)* b = (const FN(box)*)ctx;
    FN(project_variables)(n, dst, src, b->lo, b->lo_val, b->hi, b->hi_val);
}

/* ---- VMLMB driver (src/quasinewton.jl:226-612) -------------------------- */

int FN(vmlmb)(FN(fg_fn) fg, void* fg_ctx, int64_t n, REAL* x, const REAL* lo,
              double lo_val, const REAL* hi, double hi_val,
              const opko_vmlmb_options* opt, opko_vmlmb_result* res,
              FN(trace_fn) trace, void* trace_ctx)
{
    /* vmlmb! :243-293 */
    unsigned bounds = 0;
    if (lo || lo_val > -INFINITY) bounds |= 1; /* :248-258 */
    if (hi || hi_val < INFINITY) bounds |= 2;  /* :259-269 */
    const int method = (bounds == 0 ? 0 : (opt->blmvm ? 1 : 2)); /* :273 */

    opko_linesearch ls; /* :276-284 */
    int lnsrch = opt->lnsrch;
    if (lnsrch == 0)
        lnsrch = (method == 0 ? OPKO_LS_MORE_THUENTE : OPKO_LS_MORE_TORALDO);
    if (lnsrch == OPKO_LS_MORE_THUENTE)
        opko_ls_more_thuente(&ls, opt->ls_ftol, opt->ls_gtol, opt->ls_xtol);
    else if (lnsrch == OPKO_LS_MORE_TORALDO)
        opko_ls_more_toraldo(&ls, opt->ls_ftol, opt->ls_gamma1, opt->ls_gamma2,
                             opt->ls_strict);
    else
        opko_ls_armijo(&ls, opt->ls_ftol);

    /* _vmlmb! :296-612 */
    int64_t mem = opt->mem;
    const double fmin = opt->fmin;
    const int64_t maxiter = opt->maxiter, maxeval = opt->maxeval;
    const double xatol = opt->xatol, xrtol = opt->xrtol;
    const double fatol = opt->fatol, frtol = opt->frtol;
    const double gatol = opt->gatol, grtol = opt->grtol;
    const double epsilon = opt->epsilon;
    if (!(mem >= 1 && maxiter >= 0 && maxeval >= 1 && xatol >= 0 &&
          xrtol >= 0 && fatol >= 0 && frtol >= 0 && gatol >= 0 && grtol >= 0 &&
          epsilon >= 0 && epsilon < 1))
        return -1; /* the @assert's :310-319 */

    const double STPMIN = 1e-20, STPMAX = 1e+20; /* :321-322 */
    if (mem > n) mem = n;                        /* :328 */
    const char* reason = "";
    const int fminset = (!(fmin != fmin) && fmin > -INFINITY); /* :330 */

    int64_t mark = 0, m = 0, iter = 0, eval = 0, rejects = 0; /* :332-336 */
    double f = 0, f0 = 0, gd = 0, gd0 = 0, stp = 0, stpmin = 0, stpmax = 0;
    double smin = 0, smax = 0, gamma = 1, gnorm = 0, gtest = 0;
    double beststp = 0, bestf = 0, bestgnorm = 0; /* :352-354 */
    int stage = 0, status = 0;

    /* :358-371 */
    size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(REAL);
    REAL* g = (REAL*)malloc(bytes);
    REAL* p = (method > 0 ? (REAL*)malloc(bytes) : NULL);
    REAL* d = (REAL*)malloc(bytes);
    REAL* s = (REAL*)malloc(bytes);
    REAL** S = (REAL**)calloc((size_t)mem, sizeof(REAL*));
    REAL** Y = (REAL**)calloc((size_t)mem, sizeof(REAL*));
    double* rho = (double*)calloc((size_t)mem, sizeof(double));
    double* alpha = (double*)calloc((size_t)mem, sizeof(double));
    int64_t* sel = (method == 2 ? (int64_t*)malloc((size_t)(n > 0 ? n : 1) *
                                                   sizeof(int64_t))
                                : NULL);
    for (int64_t k = 0; k < mem; ++k) {
        S[k] = (REAL*)malloc(bytes);
        Y[k] = (REAL*)malloc(bytes);
    }

    for (;;) {
        /* :385-392 */
        if (method > 0) FN(project_variables)(n, x, x, lo, lo_val, hi, hi_val);
        f = fg(fg_ctx, n, x, g);
        eval += 1;
        if (method > 0) /* :395-397 */
            FN(project_direction)(n, p, x, lo, lo_val, hi, hi_val, -1, g);
        if (eval == 1 || f < bestf) { /* :398-419 */
            gnorm = FN(vnorm2)(n, (method > 0 ? p : g));
            beststp = stp;
            bestf = f;
            bestgnorm = gnorm;
            if (eval == 1) gtest = jl_max(gatol, grtol * gnorm);
            if (gnorm <= gtest) {
                stage = 3;
                if (gnorm == 0)
                    reason = "a stationary point has been found";
                else if (method > 0)
                    reason = "projected gradient norm sufficiently small";
                else
                    reason = "gradient norm sufficiently small";
                if (eval > 1) iter += 1;
            }
        }
        if (fminset && f < fmin) { /* :420-423 */
            stage = 4;
            reason = "f < fmin";
        }

        if (stage == 1) { /* :425-474 */
            FN(vcombine)(n, s, 1, x, -1, S[mark]);
            if (opko_ls_usederivatives(&ls)) {
                if (method > 0)
                    gd = FN(vdot)(n, g, s) / stp;
                else
                    gd = -FN(vdot)(n, g, d);
            }
            int task = opko_ls_iterate(&ls, stp, f, gd);
            if (task == OPKO_TASK_SEARCH) {
                stp = ls.stp;
            } else if (task == OPKO_TASK_CONVERGENCE) {
                iter += 1;
                double xtest = jl_max(xatol, 0.0);
                if (xrtol > 0) xtest = jl_max(xtest, xrtol * FN(vnorm2)(n, x));
                if (FN(vnorm2)(n, s) <= xtest) {
                    stage = 3;
                    reason = "X test satisfied";
                } else {
                    double delta = jl_max(fabs(f - f0), stp * fabs(gd0));
                    if (delta <= fatol) {
                        stage = 3;
                        reason = "fatol test satisfied";
                    } else if (delta <= frtol * fabs(f0)) {
                        stage = 3;
                        reason = "frtol test satisfied";
                    } else if (iter >= maxiter) {
                        stage = 4;
                        reason = "too many iterations";
                    } else if (eval >= maxeval) {
                        stage = 4;
                        reason = "too many evaluations";
                    } else {
                        stage = 2;
                    }
                }
            } else if (task == OPKO_TASK_ERROR) {
                status = -3; /* a thrown exception in the reference */
                stage = 4;
                reason = opko_ls_reason(&ls);
                break;
            } else {
                stage = 4; /* :469-473 */
                reason = opko_ls_reason(&ls);
            }
        }
        if (stage < 2 && eval >= maxeval) { /* :475-478 */
            stage = 4;
            reason = "too many evaluations";
        }

        if (stage != 1) { /* :480-596 */
            if (stp != beststp) { /* :485-497 */
                if (stage >= 3) {
                    stp = beststp;
                    f = bestf;
                    gnorm = bestgnorm;
                    FN(vcombine)(n, x, 1, S[mark], -stp, d);
                    if (method > 0)
                        FN(project_variables)(n, x, x, lo, lo_val, hi, hi_val);
                } else {
                    gnorm = FN(vnorm2)(n, (method > 0 ? p : g));
                }
            }
            if (trace) /* the `printer` call site :501-503 */
                trace(trace_ctx, iter, eval, rejects, f, gnorm, stp, n, x, g, p);
            if (stage >= 3) break; /* :504-506 */

            if (stage == 2) { /* :509-550 */
                FN(vupdate)(n, S[mark], -1, x);
                FN(vupdate)(n, Y[mark], -1, (method == 1 ? p : g));
                if (method < 2) {
                    rho[mark] = FN(vdot)(n, Y[mark], S[mark]);
                    if (rho[mark] > 0)
                        gamma = rho[mark] / FN(vdot)(n, Y[mark], Y[mark]);
                }
                m = (m + 1 < mem ? m + 1 : mem);
                int change;
                if (method < 2) {
                    FN(vcopy)(n, d, g);
                    change = FN(apply_lbfgs)(n, mem, S, Y, rho, gamma, NULL, m,
                                             mark, d, alpha);
                } else {
                    FN(vcopy)(n, d, p);
                    int64_t nsel = FN(get_free_variables)(n, sel, d);
                    change = FN(apply_lbfgs_sel)(n, mem, S, Y, rho, m, mark, d,
                                                 alpha, nsel, sel);
                }
                int reject;
                if (change) { /* :533-541 */
                    if (method > 0)
                        FN(project_direction)(n, d, x, lo, lo_val, hi, hi_val,
                                              -1, d);
                    gd = -FN(vdot)(n, g, d);
                    /* sufficient_descent :628-632: vnorm2(d) is evaluated even
                     * when epsilon = 0 */
                    double dnorm = FN(vnorm2)(n, d);
                    reject = !(epsilon > 0 ? (gd <= -epsilon * dnorm * gnorm)
                                           : (gd < 0));
                } else {
                    reject = 1;
                }
                if (reject) { /* :542-546 */
                    stage = 0;
                    rejects += 1;
                }
                mark = (mark < mem - 1 ? mark + 1 : 0); /* :549 */
            }
            if (stage == 0) { /* :551-556 */
                FN(vcopy)(n, d, (method > 0 ? p : g));
                gd = -gnorm * gnorm;
            }

            /* :560-563 */
            f0 = f;
            gd0 = gd;
            FN(vcopy)(n, S[mark], x);
            FN(vcopy)(n, Y[mark], (method == 1 ? p : g));

            /* :566-584 */
            if (stage == 2) {
                stp = 1;
            } else if (fminset && fmin < f0) {
                stp = 2 * (fmin - f0) / gd0;
            } else {
                stp = 1 / gnorm;
            }
            if (method > 0) {
                FN(step_limits)(n, x, lo, lo_val, hi, hi_val, -1, d, &smin,
                                &smax);
                stp = jl_min(stp, smax);
                stpmin = STPMIN * stp;
                stpmax = jl_min(STPMAX * stp, smax);
            } else {
                stpmin = STPMIN * stp;
                stpmax = STPMAX * stp;
            }

            /* :587-594 */
            int task = opko_ls_start(&ls, f0, gd0, stp, stpmin, stpmax);
            if (task != OPKO_TASK_SEARCH) {
                /* the reference throws from start! (ArgumentError) */
                status = -3;
                stage = 4;
                reason = opko_ls_reason(&ls);
                break;
            }
            stage = 1;
        }

        /* :599 */
        FN(vcombine)(n, x, 1, S[mark], -stp, d);
    }

    if (res) {
        res->iter = iter;
        res->eval = eval;
        res->rejects = rejects;
        res->f = f;
        res->gnorm = gnorm;
        res->stp = stp;
        res->stage = stage;
        strncpy(res->reason, reason, sizeof(res->reason) - 1);
        res->reason[sizeof(res->reason) - 1] = 0;
    }
    (void)smin;
    for (int64_t k = 0; k < mem; ++k) {
        free(S[k]);
        free(Y[k]);
    }
    free(S); free(Y); free(rho); free(alpha); free(sel);
    free(g); free(p); free(d); free(s);
    return status;
}

/* ---- SPG driver (src/spg.jl:193-403) ------------------------------------ */

int FN(spg)(FN(fg_fn) fg, void* fg_ctx, FN(prj_fn) prj, void* prj_ctx,
            int64_t n, REAL* x, int64_t m, int64_t maxit, int64_t maxfc,
            double eps1, double eps2, double eta, opko_spg_info* ws,
            FN(spg_trace_fn) trace, void* trace_ctx)
{
    if (!(m >= 1 && eps1 >= 0 && eps2 >= 0 && eta > 0)) return -1; /* :197-200 */
    const double lmin = 1e-30, lmax = 1e+30, ftol = 1e-4; /* :201-205 */
    const double amin = 0.1, amax = 0.9;
    int64_t iter = 0, fcnt = 0, pcnt = 0;
    int status = 0; /* SEARCHING */
    double pgtwon = INFINITY, pginfn = INFINITY;
    double f, f0 = 0, fbest, fmin, fmax, lambda, delta, stp;

    /* :222-227; d = pg, s = x0, y = g0 are aliases in the reference too */
    double* lastfv = (double*)malloc((size_t)m * sizeof(double));
    for (int64_t i = 0; i < m; ++i) lastfv[i] = -INFINITY;
    size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(REAL);
    REAL* g = (REAL*)malloc(bytes);
    REAL* pg = (REAL*)malloc(bytes);
    REAL* x0 = (REAL*)malloc(bytes);
    REAL* g0 = (REAL*)malloc(bytes);
    REAL* xbest = (REAL*)malloc(bytes);
    REAL* d = pg;
    REAL* s = x0;
    REAL* y = g0;

    prj(prj_ctx, n, x, x); /* :230-232 */
    FN(vcopy)(n, x0, x);
    pcnt += 1;
    f = fg(fg_ctx, n, x, g); /* :236-237 */
    fcnt += 1;
    fbest = f; /* :240-241 */
    FN(vcopy)(n, xbest, x);

    int fconst = 0;
    for (;;) {
        if (m > 1) { /* :249-255 */
            lastfv[iter % m] = f;
            fmin = fmax = lastfv[0];
            for (int64_t i = 1; i < m; ++i) {
                fmin = jl_min(fmin, lastfv[i]);
                fmax = jl_max(fmax, lastfv[i]);
            }
            fconst = !(fmin < fmax);
        } else {
            fmin = fmax = f;
        }

        /* :259-262 */
        FN(vcombine)(n, pg, 1, x, -eta, g);
        prj(prj_ctx, n, pg, pg);
        FN(vcombine)(n, pg, 1 / eta, x, -1 / eta, pg);
        pcnt += 1;
        pgtwon = FN(vnorm2)(n, pg);
        pginfn = FN(vnorminf)(n, pg);

        if (trace) { /* the `printer` call site :265-275 */
            opko_spg_info nfo;
            nfo.f = f; nfo.fbest = fbest; nfo.pgtwon = pgtwon;
            nfo.pginfn = pginfn; nfo.iter = iter; nfo.fcnt = fcnt;
            nfo.pcnt = pcnt; nfo.status = status;
            trace(trace_ctx, &nfo, n, x, pg);
        }

        /* :278-302 */
        if (pginfn <= eps1) { status = 1; break; }
        if (pgtwon <= eps2) { status = 2; break; }
        if (fconst) { status = 3; break; }
        if (iter >= maxit) { status = -1; break; }
        if (fcnt >= maxfc) { status = -2; break; }

        if (iter == 0) { /* :305-319 */
#if !REAL_IS_DOUBLE
            if (g_sum_mode != OPKO_SUM_ACCURATE)
                /* `1/pginfn` is evaluated in Float32 when vnorminf returns the
                 * element type */
                lambda = jl_clamp((double)(1.0f / (float)pginfn), lmin, lmax);
            else
#endif
                lambda = jl_clamp(1 / pginfn, lmin, lmax);
        } else {
            FN(vcombine)(n, s, 1, x, -1, x0);
            FN(vcombine)(n, y, 1, g, -1, g0);
            double sty = FN(vdot)(n, s, y);
            if (sty > 0) {
                double sts = FN(vdot)(n, s, s);
                lambda = jl_clamp(sts / sty, lmin, lmax);
            } else {
                lambda = lmax;
            }
        }

        FN(vcopy)(n, x0, x); /* :322-324 */
        FN(vcopy)(n, g0, g);
        f0 = f;

        FN(vcombine)(n, x, 1, x0, -lambda, g0); /* :327-330 */
        prj(prj_ctx, n, x, x);
        pcnt += 1;
        FN(vcombine)(n, d, 1, x, -1, x0);
        delta = FN(vdot)(n, g0, d);

        stp = 1.0; /* :333-369 */
        for (;;) {
            f = fg(fg_ctx, n, x, g);
            fcnt += 1;
            if (f < fbest) {
                fbest = f;
                FN(vcopy)(n, xbest, x);
            }
            if (f <= fmax + stp * ftol * delta) break;
            if (fcnt >= maxfc) { status = -2; break; }
            double q = -delta * (stp * stp);
            double r = (f - f0 - stp * delta) * 2;
            if (r > 0 && amin * r <= q && q <= amax * stp * r)
                stp = q / r;
            else
                stp /= 2;
            FN(vcombine)(n, x, 1, x0, stp, d);
        }
        if (status != 0) break; /* :371-375 */
        iter += 1;
    }

    if (ws) { /* :383-390 */
        ws->f = fbest; ws->fbest = fbest; ws->pgtwon = pgtwon;
        ws->pginfn = pginfn; ws->iter = iter; ws->fcnt = fcnt;
        ws->pcnt = pcnt; ws->status = status;
    }
    if (fbest < f) FN(vcopy)(n, x, xbest); /* :401 */
    (void)fmin;
    free(lastfv); free(g); free(pg); free(x0); free(g0); free(xbest);
    return 0;
}

/* ---- linear conjugate gradient (LazyAlgebra `conjgrad!`) ----------------- */
/* `conjgrad` / `conjgrad!` are re-exported by the reference from LazyAlgebra.jl
 * (src/OptimPackNextGen.jl:18-19, :36); LazyAlgebra is neither vendored nor version
 * pinned (Project.toml:8, no [compat] entry), so this is a restatement of its
 * PUBLISHED algorithm (LazyAlgebra src/conjgrad.jl, 0.2.x series: Hestenes-Stiefel
 * recurrences with beta = rho/oldrho, keywords ftol, gtol = (gatol, grtol), xtol,
 * maxiter, restart, strict), anchored on the reference's call site
 * test/conjgrad-tests.jl:16-36 (`conjgrad(A!, b; verb, ftol=1e-16,
 * maxiter=length(b))`, max|x - A\b| <= 1e-7 in Float64, 1e-5 in Float32).
 * PARITY UNPINNED for the same reasons as the rest of this file.
 *
 * Minimises phi(x) = x'.A.x/2 - b'.x for a symmetric positive definite A given as
 * `apply(ctx, n, dst, src)`; x holds the starting point on entry.  status:
 * 1 = gtest, 2 = ftest, 3 = xtest, -1 = too many iterations, -2 = A is not
 * positive definite. */
int FN(conjgrad)(FN(op_fn) apply, void* op_ctx, int64_t n, REAL* x,
                 const REAL* b, double ftol, double gatol, double grtol,
                 double xtol, int64_t maxiter, int64_t restart,
                 opko_cg_info* info, FN(cg_trace_fn) trace, void* trace_ctx)
{
    if (!(ftol >= 0 && gatol >= 0 && grtol >= 0 && xtol >= 0 && restart >= 1))
        return -1;
    size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(REAL);
    REAL* p = (REAL*)malloc(bytes);
    REAL* q = (REAL*)malloc(bytes);
    REAL* r = (REAL*)malloc(bytes);
    const int xtest = (xtol > 0);
    int status = 0;
    int64_t k = 0;
    double rho, oldrho = 0, phi = 0, psi = 0, psimax = 0;

    if (FN(vdot)(n, x, x) > 0) {       /* r = b - A.x (x0 may be non-zero) */
        apply(op_ctx, n, r, x);
        FN(vcombine)(n, r, 1, b, -1, r);
    } else {                            /* spare one product when x0 = 0 */
        FN(vcopy)(n, r, b);
    }
    rho = FN(vdot)(n, r, r);
    const double gtest = jl_max(gatol, grtol * sqrt(rho));

    for (;;) {
        if (trace) trace(trace_ctx, k, phi, rho, n, x, r);
        if (sqrt(rho) <= gtest) { status = 1; break; }
        if (k >= maxiter) { status = -1; break; }

        /* search direction */
        if (k % restart == 0) {
            if (k > 0) {                /* restart: true residuals */
                apply(op_ctx, n, r, x);
                FN(vcombine)(n, r, 1, b, -1, r);
            }
            FN(vcopy)(n, p, r);
        } else {
            double beta = rho / oldrho;
            FN(vcombine)(n, p, beta, p, 1, r);
        }

        /* optimal step */
        apply(op_ctx, n, q, p);
        double gamma = FN(vdot)(n, p, q);
        if (!(gamma > 0)) { status = -2; break; }
        double alpha = rho / gamma;

        /* variables, convergence in phi and in x */
        FN(vupdate)(n, x, alpha, p);
        psi = alpha * rho / 2;          /* phi(x_k) - phi(x_{k+1}) */
        phi -= psi;
        psimax = jl_max(psi, psimax);
        if (psi <= ftol * psimax) { status = 2; k += 1; break; }
        if (xtest && alpha * FN(vnorm2)(n, p) <= xtol * FN(vnorm2)(n, x)) {
            status = 3; k += 1; break;
        }

        /* residuals */
        FN(vupdate)(n, r, -alpha, q);
        oldrho = rho;
        rho = FN(vdot)(n, r, r);
        k += 1;
    }
    if (info) {
        info->iter = k; info->status = status; info->rho = rho;
        info->phi = phi; info->psi = psi;
    }
    free(p); free(q); free(r);
    return 0;
}

/* A = diag(a): `vproduct!` */
void FN(diag_apply)(void* ctx, int64_t n, REAL* dst, const REAL* src)
{
    FN(vproduct)(n, dst, (const REAL*)ctx, src);
}

#undef REAL_INF
#undef SUM_BLOCK
#undef SUM_CHUNK
