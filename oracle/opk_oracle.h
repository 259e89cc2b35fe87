/*
 * opk_oracle.h -- CPU oracle for the VMLMB / SPG hot path of OptimPackNextGen.jl.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (Julia, /root/reference, v0.4.3).  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py
 * may load it; the product (libopkb200.so and the host package) never does.
 *
 * PARITY UNPINNED.  The reference holds no golden vectors for this path, the
 * per-element arithmetic lives in LazyAlgebra.jl (not vendored, not version
 * pinned: Project.toml:8, no [compat] entry) and no Julia runtime exists in
 * this environment, so the oracle cannot be checked against outputs of the
 * reference itself.  It is pinned only against (i) the reference's own test
 * assertions (test/rosenbrock.jl:79,92,104: max|x-1| < 1e-3 for n=20) and
 * (ii) analytic known answers (tests/test_oracle.py), and two of its parts are
 * pinned against independent published implementations shipped with SciPy
 * (tests/test_pins_scipy.py): the More-Thuente search + cstep give bit-identical
 * step sequences to SciPy's port of MINPACK-2 dcsrch/dcstep, and apply_lbfgs
 * equals scipy.optimize.LbfgsInvHessProduct to rounding, and the unconstrained
 * driver (method 0) follows SciPy's L-BFGS-B (Nocedal's code; same line-search
 * constants, first step and H0 scaling) to 1e-10 over the first 20 iterations.
 * The bound-constrained and SPG trajectories remain unpinned.  `conjgrad`
 * (LazyAlgebra's, re-exported by the reference) is a restatement of the published
 * algorithm anchored on test/conjgrad-tests.jl and scipy.sparse.linalg.cg.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Two instances of the type-generic part exist: prefix
 * opko_d_ (Float64 data) and opko_f_ (Float32 data).  All scalars are double,
 * as in the reference (`Float = Cdouble`, src/OptimPackNextGen.jl:42).
 */
#ifndef OPK_ORACLE_H
#define OPK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reduction modes ---------------------------------------------------- */
/* 0: accumulate in the element type, index order (a generic Julia loop).
 * 1: pairwise in the element type (what Julia's `sum` does).
 * 2: exact products converted to double, blocked pairwise sum in double
 *    (canonical parity mode: accuracy ~1e-16 whatever the element type). */
#define OPKO_SUM_SEQUENTIAL 0
#define OPKO_SUM_PAIRWISE   1
#define OPKO_SUM_ACCURATE   2
void opko_set_sum_mode(int mode);
int  opko_get_sum_mode(void);
void opko_set_num_threads(int nthreads); /* OpenMP threads (0 = all) */
int  opko_get_max_threads(void);

/* ---- line searches (src/linesearches.jl) -------------------------------- */
#define OPKO_LS_ARMIJO        1
#define OPKO_LS_MORE_TORALDO  2
#define OPKO_LS_MORE_THUENTE  3

#define OPKO_TASK_START       0
#define OPKO_TASK_SEARCH      1
#define OPKO_TASK_CONVERGENCE 2
#define OPKO_TASK_WARNING     3
#define OPKO_TASK_ERROR       4   /* stands for a thrown ArgumentError */

/* status codes, in the order of REASONS (src/linesearches.jl:33-51) */
enum {
    OPKO_NOT_STARTED = 0, OPKO_F_LE_FMIN, OPKO_COMPUTE_FG, OPKO_NEW_X,
    OPKO_FINAL_X, OPKO_FATOL_TEST_SATISFIED, OPKO_FRTOL_TEST_SATISFIED,
    OPKO_GATOL_TEST_SATISFIED, OPKO_GRTOL_TEST_SATISFIED,
    OPKO_FIRST_WOLFE_SATISFIED, OPKO_STRONG_WOLFE_SATISFIED,
    OPKO_STP_EQ_STPMIN, OPKO_STP_EQ_STPMAX, OPKO_XTOL_TEST_SATISFIED,
    OPKO_ROUNDING_ERRORS, OPKO_NOT_DESCENT, OPKO_INIT_STP_LE_ZERO,
    /* errors thrown by start!/cstep, mapped to statuses */
    OPKO_ERR_STPMAX_LT_STPMIN, OPKO_ERR_STPMIN_LT_0, OPKO_ERR_STP_GT_STPMAX,
    OPKO_ERR_STP_LT_STPMIN, OPKO_ERR_XTOL_LT_0, OPKO_ERR_FTOL_LE_0,
    OPKO_ERR_GTOL_LE_0, OPKO_ERR_OUTSIDE_BRACKET, OPKO_ERR_DESCENT_VIOLATED
};

typedef struct {
    int kind, task, status;
    /* common */
    double ftol, finit, ginit, stp, stpmin;
    /* More & Toraldo */
    double gamma1, gamma2; int strict;
    /* More & Thuente */
    double gtol, xtol, stpmax;
    double stx, fx, gx, sty, fy, gy, lower, upper, width, oldwidth;
    int stage, brackt;
} opko_linesearch;

void opko_ls_armijo(opko_linesearch* ls, double ftol);
void opko_ls_more_toraldo(opko_linesearch* ls, double ftol, double gamma1,
                          double gamma2, int strict);
void opko_ls_more_thuente(opko_linesearch* ls, double ftol, double gtol,
                          double xtol);
int  opko_ls_usederivatives(const opko_linesearch* ls);
int  opko_ls_start(opko_linesearch* ls, double f0, double g0, double stp,
                   double stpmin, double stpmax);
int  opko_ls_iterate(opko_linesearch* ls, double stp, double f, double g);
const char* opko_ls_reason(const opko_linesearch* ls);
/* cstep (src/linesearches.jl:709-870); v = {stx,fx,dx,sty,fy,dy,stp}, returns
 * info (1..4) or a negative value for a thrown error. */
int  opko_cstep(int* brackt, double stpmin, double stpmax, double v[7],
                double fp, double dp);

/* ---- VMLMB driver options / result (src/quasinewton.jl:226-242) ---------- */
typedef struct {
    int64_t mem;        /* default min(5, n) */
    int     blmvm;      /* EMULATE_BLMVM */
    double  fmin;       /* -Inf = unset */
    int64_t maxiter, maxeval;
    double  xatol, xrtol, fatol, frtol, gatol, grtol, epsilon;
    int     lnsrch;     /* 0 = default for the method, else OPKO_LS_* */
    double  ls_ftol, ls_gtol, ls_xtol, ls_gamma1, ls_gamma2;
    int     ls_strict;
} opko_vmlmb_options;
void opko_vmlmb_default_options(opko_vmlmb_options* o);

typedef struct {
    int64_t iter, eval, rejects;
    double  f, gnorm, stp;
    int     stage;      /* 3 = convergence, 4 = warning */
    char    reason[96];
} opko_vmlmb_result;

/* ---- SPG (src/spg.jl:31-49) --------------------------------------------- */
typedef struct {
    double  f, fbest, pginfn, pgtwon;
    int64_t iter, fcnt, pcnt;
    int     status;
} opko_spg_info;

/* ---- conjgrad (LazyAlgebra, re-exported: src/OptimPackNextGen.jl:18-19) --- */
typedef struct {
    double  rho, phi, psi;
    int64_t iter;
    int     status;
} opko_cg_info;

/* ---- type-generic part: declared twice ---------------------------------- */
#define OPKO_DECLARE(P, T)                                                     \
    typedef double (*P##fg_fn)(void* ctx, int64_t n, const T* x, T* g);        \
    typedef void (*P##prj_fn)(void* ctx, int64_t n, T* dst, const T* src);     \
    /* called where the reference calls `printer` (src/quasinewton.jl:501);    \
     * p is NULL for the unconstrained method */                               \
    typedef void (*P##trace_fn)(void* ctx, int64_t iter, int64_t eval,         \
                                int64_t rejects, double f, double gnorm,       \
                                double stp, int64_t n, const T* x, const T* g, \
                                const T* p);                                   \
    typedef void (*P##spg_trace_fn)(void* ctx, const opko_spg_info* info,      \
                                    int64_t n, const T* x, const T* pg);       \
    typedef void (*P##op_fn)(void* ctx, int64_t n, T* dst, const T* src);      \
    typedef void (*P##cg_trace_fn)(void* ctx, int64_t iter, double phi,        \
                                   double rho, int64_t n, const T* x,          \
                                   const T* r);                                \
    /* LazyAlgebra vector ops (doc/algebra.md:84-116) */                       \
    double P##vdot(int64_t n, const T* x, const T* y);                         \
    double P##vdot_sel(int64_t nsel, const int64_t* sel, const T* x,           \
                       const T* y);                                            \
    double P##vnorm2(int64_t n, const T* x);                                   \
    double P##vnorminf(int64_t n, const T* x);                                 \
    void   P##vcopy(int64_t n, T* dst, const T* src);                          \
    void   P##vscale(int64_t n, T* x, double alpha);                           \
    void   P##vupdate(int64_t n, T* y, double alpha, const T* x);              \
    void   P##vupdate_sel(T* y, int64_t nsel, const int64_t* sel,              \
                          double alpha, const T* x);                           \
    void   P##vcombine(int64_t n, T* dst, double alpha, const T* x,            \
                       double beta, const T* y);                               \
    void   P##vproduct(int64_t n, T* dst, const T* x, const T* y);             \
    /* SimpleBounds (src/bounds.jl).  A bound is an array when the pointer is  \
     * non-NULL, else the scalar (+-Inf = none). */                            \
    void   P##project_variables(int64_t n, T* dst, const T* src,               \
                                const T* lo, double lo_val, const T* hi,       \
                                double hi_val);                                \
    int    P##project_direction(int64_t n, T* dst, const T* x, const T* lo,    \
                                double lo_val, const T* hi, double hi_val,     \
                                int orient, const T* d);                       \
    int    P##step_limits(int64_t n, const T* x, const T* lo, double lo_val,   \
                          const T* hi, double hi_val, int orient, const T* d,  \
                          double* smin, double* smax);                         \
    int64_t P##get_free_variables(int64_t n, int64_t* sel, const T* gp);       \
    int64_t P##get_free_variables_dir(int64_t n, int64_t* sel, const T* x,     \
                                      const T* lo, double lo_val, const T* hi, \
                                      double hi_val, int orient, const T* d);  \
    /* apply_lbfgs! (src/quasinewton.jl:703-779); slots are 0-based here */    \
    int    P##apply_lbfgs(int64_t n, int64_t mem, T* const* S, T* const* Y,    \
                          const double* rho, double gamma, const T* diag,      \
                          int64_t m, int64_t mark, T* v, double* alpha);       \
    int    P##apply_lbfgs_sel(int64_t n, int64_t mem, T* const* S,             \
                              T* const* Y, double* rho, int64_t m,             \
                              int64_t mark, T* v, double* alpha,               \
                              int64_t nsel, const int64_t* sel);               \
    /* objectives */                                                           \
    double P##rosenbrock_fg(void* ctx, int64_t n, const T* x, T* g);           \
    void   P##rosenbrock_init(int64_t n, T* x0);                               \
    typedef struct { const T* a; const T* c; } P##quadratic;                   \
    double P##quadratic_fg(void* ctx, int64_t n, const T* x, T* g);            \
    typedef struct { const T* lo; double lo_val; const T* hi; double hi_val; } \
        P##box;                                                                \
    void   P##box_prj(void* ctx, int64_t n, T* dst, const T* src);             \
    /* drivers */                                                              \
    int    P##vmlmb(P##fg_fn fg, void* fg_ctx, int64_t n, T* x, const T* lo,   \
                    double lo_val, const T* hi, double hi_val,                 \
                    const opko_vmlmb_options* opt, opko_vmlmb_result* res,     \
                    P##trace_fn trace, void* trace_ctx);                       \
    int    P##spg(P##fg_fn fg, void* fg_ctx, P##prj_fn prj, void* prj_ctx,     \
                  int64_t n, T* x, int64_t m, int64_t maxit, int64_t maxfc,    \
                  double eps1, double eps2, double eta, opko_spg_info* ws,     \
                  P##spg_trace_fn trace, void* trace_ctx);                     \
    int    P##conjgrad(P##op_fn apply, void* op_ctx, int64_t n, T* x,          \
                       const T* b, double ftol, double gatol, double grtol,    \
                       double xtol, int64_t maxiter, int64_t restart,          \
                       opko_cg_info* info, P##cg_trace_fn trace,               \
                       void* trace_ctx);                                       \
    void   P##diag_apply(void* ctx, int64_t n, T* dst, const T* src);

OPKO_DECLARE(opko_d_, double)
OPKO_DECLARE(opko_f_, float)

#ifdef __cplusplus
}
#endif
#endif /* OPK_ORACLE_H */
