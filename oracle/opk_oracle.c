/*
 * opk_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see opk_oracle.h).
 *
 * Scalar state machines restated from /root/reference/src/linesearches.jl,
 * then the type-generic vector part (opk_oracle_impl.h) instantiated for
 * double and float.  PARITY UNPINNED (no reference golden vectors, no Julia
 * runtime here; see the header).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp; contraction is
 * off so that a*x + b*y is two roundings, as in the reference's Julia loops).
 */
#include "opk_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
static int g_sum_mode = OPKO_SUM_ACCURATE;
void opko_set_sum_mode(int mode) { g_sum_mode = mode; }
int opko_get_sum_mode(void) { return g_sum_mode; }

void opko_set_num_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_num_procs();
    omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

int opko_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Julia's max/min propagate NaN; clamp(x,lo,hi) = x>hi ? hi : x<lo ? lo : x. */
static double jl_max(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    return (a > b ? a : b);
}
static double jl_min(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    return (a < b ? a : b);
}
static double jl_clamp(double x, double lo, double hi)
{
    return (x > hi ? hi : (x < lo ? lo : x));
}

/* ------------------------------------------------------------------------ */
/* Line searches.                                                            */

static const char* const REASONS[] = {
    /* src/linesearches.jl:33-51 */
    "Algorithm not yet started",
    "F(X) \xE2\x89\xA4 FMIN",
    "Caller must compute f(x) and g(x)",
    "A new iterate is available for examination",
    "A solution has been found within tolerances",
    "FATOL test satisfied",
    "FRTOL test satisfied",
    "GATOL test satisfied",
    "GRTOL test satisfied",
    "First Wolfe condition satisfied",
    "Strong Wolfe conditions both satisfied",
    "Step at lower bound",
    "Step at upper bound",
    "XTOL test satisfied",
    "Rounding errors prevent progress",
    "Search direction is not a descent direction",
    "Initial step \xE2\x89\xA4 0",
    /* ArgumentError messages (src/linesearches.jl:334-349, 476-499, 724-732) */
    "STPMAX < STPMIN",
    "STPMIN < 0",
    "STP > STPMAX",
    "STP < STPMIN",
    "XTOL < 0",
    "FTOL \xE2\x89\xA4 0",
    "GTOL \xE2\x89\xA4 0",
    "STP outside bracket (STX,STY)",
    "descent condition violated",
};

static int report(opko_linesearch* ls, int task, int status)
{
    /* src/linesearches.jl:121-125 */
    ls->status = status;
    ls->task = task;
    return task;
}

const char* opko_ls_reason(const opko_linesearch* ls)
{
    /* src/linesearches.jl:133.  Note: the reference reports the key
     * :FIRST_WOLFE_CONDITION_SATISFIED (:239, :364) which is absent from
     * REASONS, so getreason would throw there; the oracle returns the text of
     * :FIRST_WOLFE_SATISFIED instead. */
    int n = (int)(sizeof(REASONS) / sizeof(REASONS[0]));
    if (ls->status < 0 || ls->status >= n) return "unknown status";
    return REASONS[ls->status];
}

void opko_ls_armijo(opko_linesearch* ls, double ftol)
{
    /* src/linesearches.jl:222-226 */
    memset(ls, 0, sizeof(*ls));
    ls->kind = OPKO_LS_ARMIJO;
    ls->task = OPKO_TASK_START;
    ls->status = OPKO_NOT_STARTED;
    ls->ftol = ftol;
}

void opko_ls_more_toraldo(opko_linesearch* ls, double ftol, double gamma1,
                          double gamma2, int strict)
{
    /* src/linesearches.jl:316-325 */
    memset(ls, 0, sizeof(*ls));
    ls->kind = OPKO_LS_MORE_TORALDO;
    ls->task = OPKO_TASK_START;
    ls->status = OPKO_NOT_STARTED;
    ls->ftol = ftol;
    ls->gamma1 = gamma1;
    ls->gamma2 = gamma2;
    ls->strict = strict;
}

void opko_ls_more_thuente(opko_linesearch* ls, double ftol, double gtol,
                          double xtol)
{
    /* src/linesearches.jl:438-449 */
    memset(ls, 0, sizeof(*ls));
    ls->kind = OPKO_LS_MORE_THUENTE;
    ls->task = OPKO_TASK_START;
    ls->status = OPKO_NOT_STARTED;
    ls->ftol = ftol;
    ls->gtol = gtol;
    ls->xtol = xtol;
}

int opko_ls_usederivatives(const opko_linesearch* ls)
{
    /* src/linesearches.jl:230, :329, :453 */
    return ls->kind == OPKO_LS_MORE_THUENTE;
}

#define XTRAPL 1.1 /* src/linesearches.jl:456 */
#define XTRAPU 4.0 /* src/linesearches.jl:457 */

int opko_ls_start(opko_linesearch* ls, double f0, double g0, double stp,
                  double stpmin, double stpmax)
{
    if (ls->kind == OPKO_LS_MORE_THUENTE) {
        /* src/linesearches.jl:470-529 */
        if (stpmax < stpmin)
            return report(ls, OPKO_TASK_ERROR, OPKO_ERR_STPMAX_LT_STPMIN);
        if (stpmin < 0)
            return report(ls, OPKO_TASK_ERROR, OPKO_ERR_STPMIN_LT_0);
        if (ls->xtol < 0)
            return report(ls, OPKO_TASK_ERROR, OPKO_ERR_XTOL_LT_0);
        if (ls->ftol <= 0)
            return report(ls, OPKO_TASK_ERROR, OPKO_ERR_FTOL_LE_0);
        if (ls->gtol <= 0)
            return report(ls, OPKO_TASK_ERROR, OPKO_ERR_GTOL_LE_0);
        if (g0 >= 0) return report(ls, OPKO_TASK_ERROR, OPKO_NOT_DESCENT);
        if (stp > stpmax)
            return report(ls, OPKO_TASK_ERROR, OPKO_ERR_STP_GT_STPMAX);
        if (stp < stpmin)
            return report(ls, OPKO_TASK_ERROR, OPKO_ERR_STP_LT_STPMIN);
        ls->stpmin = stpmin;
        ls->stpmax = stpmax;
        ls->finit = f0;
        ls->ginit = g0;
        ls->stp = stp;
        ls->stx = 0;
        ls->fx = f0;
        ls->gx = g0;
        ls->sty = 0;
        ls->fy = f0;
        ls->gy = g0;
        ls->lower = 0;
        ls->upper = stp + stp * XTRAPU;
        ls->width = stpmax - stpmin;
        ls->oldwidth = 2 * (stpmax - stpmin);
        ls->stage = 1;
        ls->brackt = 0;
        return report(ls, OPKO_TASK_SEARCH, OPKO_COMPUTE_FG);
    }
    /* src/linesearches.jl:331-358 (Armijo and More & Toraldo) */
    if (stpmax < stpmin)
        return report(ls, OPKO_TASK_ERROR, OPKO_ERR_STPMAX_LT_STPMIN);
    if (stpmin < 0) return report(ls, OPKO_TASK_ERROR, OPKO_ERR_STPMIN_LT_0);
    if (stp > stpmax) return report(ls, OPKO_TASK_ERROR, OPKO_ERR_STP_GT_STPMAX);
    if (stp < stpmin) return report(ls, OPKO_TASK_ERROR, OPKO_ERR_STP_LT_STPMIN);
    if (g0 >= 0) return report(ls, OPKO_TASK_ERROR, OPKO_NOT_DESCENT);
    ls->finit = f0;
    ls->ginit = g0;
    ls->stp = stp;
    ls->stpmin = stpmin;
    return report(ls, OPKO_TASK_SEARCH, OPKO_COMPUTE_FG);
}

int opko_cstep(int* brackt_io, double stpmin, double stpmax, double v[7],
               double fp, double dp)
{
    /* src/linesearches.jl:709-870 */
    double stx = v[0], fx = v[1], dx = v[2];
    double sty = v[3], fy = v[4], dy = v[5], stp = v[6];
    double theta, gamma, p, q, r, s, t, stpc, stpq, stpf;
    int brackt = *brackt_io, info, opposite;

    if (brackt && (stp <= jl_min(stx, sty) || stp >= jl_max(stx, sty)))
        return -OPKO_ERR_OUTSIDE_BRACKET;
    if (dx * (stp - stx) >= 0) return -OPKO_ERR_DESCENT_VIOLATED;
    if (stpmax < stpmin) return -OPKO_ERR_STPMAX_LT_STPMIN;

    opposite = ((dx < 0 && dp > 0) || (dx > 0 && dp < 0));

    if (fp > fx) {
        info = 1; /* :737-759 */
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        s = jl_max(jl_max(fabs(theta), fabs(dx)), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2) * (stp - stx);
        if (fabs(stpc - stx) < fabs(stpq - stx))
            stpf = stpc;
        else
            stpf = stpc + (stpq - stpc) / 2;
        brackt = 1;
    } else if (opposite) {
        info = 2; /* :760-782 */
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        s = jl_max(jl_max(fabs(theta), fabs(dx)), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp))
            stpf = stpc;
        else
            stpf = stpq;
        brackt = 1;
    } else if (fabs(dp) < fabs(dx)) {
        info = 3; /* :783-826 */
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        s = jl_max(jl_max(fabs(theta), fabs(dx)), fabs(dp));
        t = (theta / s) * (theta / s) - (dx / s) * (dp / s);
        gamma = (t > 0 ? s * sqrt(t) : 0.0);
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0 && gamma != 0)
            stpc = stp + r * (stx - stp);
        else if (stp > stx)
            stpc = stpmax;
        else
            stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = (fabs(stpc - stp) < fabs(stpq - stp) ? stpc : stpq);
            t = stp + 0.66 * (sty - stp);
            if (stp > stx ? stpf > t : stpf < t) stpf = t;
        } else {
            stpf = (fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq);
            stpf = jl_max(stpmin, jl_min(stpf, stpmax));
        }
    } else {
        info = 4; /* :827-851 */
        if (brackt) {
            theta = 3 * (fp - fy) / (sty - stp) + dy + dp;
            s = jl_max(jl_max(fabs(theta), fabs(dy)), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }

    /* :853-869 */
    if (fp > fx) {
        sty = stp;
        fy = fp;
        dy = dp;
    } else {
        if (opposite) {
            sty = stx;
            fy = fx;
            dy = dx;
        }
        stx = stp;
        fx = fp;
        dx = dp;
    }
    stp = stpf;
    v[0] = stx; v[1] = fx; v[2] = dx;
    v[3] = sty; v[4] = fy; v[5] = dy; v[6] = stp;
    *brackt_io = brackt;
    return info;
}

int opko_ls_iterate(opko_linesearch* ls, double stp, double f, double g)
{
    /* the @assert's of the reference (:236-237, :361-362, :539-540) */
    if (ls->task != OPKO_TASK_SEARCH || stp != ls->stp)
        return report(ls, OPKO_TASK_ERROR, OPKO_NOT_STARTED);

    if (ls->kind == OPKO_LS_ARMIJO) {
        /* src/linesearches.jl:235-247 */
        if (f <= ls->finit + ls->ftol * stp * ls->ginit) {
            return report(ls, OPKO_TASK_CONVERGENCE, OPKO_FIRST_WOLFE_SATISFIED);
        } else if (stp > ls->stpmin) {
            ls->stp = jl_max(stp / 2, ls->stpmin);
            return report(ls, OPKO_TASK_SEARCH, OPKO_COMPUTE_FG);
        } else {
            ls->stp = ls->stpmin;
            return report(ls, OPKO_TASK_WARNING, OPKO_STP_EQ_STPMIN);
        }
    }

    if (ls->kind == OPKO_LS_MORE_TORALDO) {
        /* src/linesearches.jl:360-381 */
        if (f <= ls->finit + ls->ftol * stp * ls->ginit) {
            return report(ls, OPKO_TASK_CONVERGENCE, OPKO_FIRST_WOLFE_SATISFIED);
        } else if (stp > ls->stpmin) {
            double q = -ls->ginit * stp;
            double r = f - (ls->finit - q);
            if (ls->strict ? (r > 0 && q > 0) : (r != 0)) {
                stp *= jl_clamp(q / (r + r), ls->gamma1, ls->gamma2);
            } else {
                stp /= 2;
            }
            ls->stp = jl_max(stp, ls->stpmin);
            return report(ls, OPKO_TASK_SEARCH, OPKO_COMPUTE_FG);
        } else {
            ls->stp = ls->stpmin;
            return report(ls, OPKO_TASK_WARNING, OPKO_STP_EQ_STPMIN);
        }
    }

    /* More & Thuente: src/linesearches.jl:536-647 */
    {
        double gtest = ls->ftol * ls->ginit;
        double ftest = ls->finit + stp * gtest;
        double v[7];
        int info, brackt;

        if (ls->stage == 1 && f <= ftest && g >= 0) ls->stage = 2;

        if (f <= ftest && fabs(g) <= -ls->gtol * ls->ginit)
            return report(ls, OPKO_TASK_CONVERGENCE, OPKO_STRONG_WOLFE_SATISFIED);
        if (stp == ls->stpmin && (f > ftest || g >= gtest))
            return report(ls, OPKO_TASK_WARNING, OPKO_STP_EQ_STPMIN);
        if (stp == ls->stpmax && f <= ftest && g <= gtest)
            return report(ls, OPKO_TASK_WARNING, OPKO_STP_EQ_STPMAX);
        if (ls->brackt) {
            if (ls->upper - ls->lower <= ls->xtol * ls->upper)
                return report(ls, OPKO_TASK_WARNING, OPKO_XTOL_TEST_SATISFIED);
            if (stp <= ls->lower || stp >= ls->upper)
                return report(ls, OPKO_TASK_WARNING, OPKO_ROUNDING_ERRORS);
        }

        brackt = ls->brackt;
        if (ls->stage == 1 && f <= ls->fx && f > ftest) {
            /* :577-593, modified function */
            v[0] = ls->stx; v[1] = ls->fx - ls->stx * gtest; v[2] = ls->gx - gtest;
            v[3] = ls->sty; v[4] = ls->fy - ls->sty * gtest; v[5] = ls->gy - gtest;
            v[6] = stp;
            info = opko_cstep(&brackt, ls->lower, ls->upper, v,
                              f - stp * gtest, g - gtest);
            if (info < 0) return report(ls, OPKO_TASK_ERROR, -info);
            ls->brackt = brackt;
            ls->stx = v[0];
            ls->sty = v[3];
            stp = v[6];
            ls->fx = v[1] + ls->stx * gtest;
            ls->fy = v[4] + ls->sty * gtest;
            ls->gx = v[2] + gtest;
            ls->gy = v[5] + gtest;
        } else {
            /* :597-604 */
            v[0] = ls->stx; v[1] = ls->fx; v[2] = ls->gx;
            v[3] = ls->sty; v[4] = ls->fy; v[5] = ls->gy;
            v[6] = stp;
            info = opko_cstep(&brackt, ls->lower, ls->upper, v, f, g);
            if (info < 0) return report(ls, OPKO_TASK_ERROR, -info);
            ls->brackt = brackt;
            ls->stx = v[0]; ls->fx = v[1]; ls->gx = v[2];
            ls->sty = v[3]; ls->fy = v[4]; ls->gy = v[5];
            stp = v[6];
        }

        /* :608-616 */
        if (ls->brackt) {
            double wcur = fabs(ls->sty - ls->stx);
            if (wcur >= 0.66 * ls->oldwidth)
                stp = ls->stx + 0.5 * (ls->sty - ls->stx);
            ls->oldwidth = ls->width;
            ls->width = wcur;
        }

        /* :618-630 */
        if (ls->brackt) {
            if (ls->stx <= ls->sty) {
                ls->lower = ls->stx;
                ls->upper = ls->sty;
            } else {
                ls->lower = ls->sty;
                ls->upper = ls->stx;
            }
        } else {
            ls->lower = stp + XTRAPL * (stp - ls->stx);
            ls->upper = stp + XTRAPU * (stp - ls->stx);
        }

        /* :632-646 */
        stp = jl_clamp(stp, ls->stpmin, ls->stpmax);
        if (ls->brackt && (stp <= ls->lower || stp >= ls->upper ||
                           ls->upper - ls->lower <= ls->xtol * ls->upper))
            stp = ls->stx;
        ls->stp = stp;
        return report(ls, OPKO_TASK_SEARCH, OPKO_COMPUTE_FG);
    }
}

/* ------------------------------------------------------------------------ */
void opko_vmlmb_default_options(opko_vmlmb_options* o)
{
    /* src/quasinewton.jl:227-242 */
    o->mem = 5;
    o->blmvm = 0;
    o->fmin = -INFINITY;
    o->maxiter = INT64_MAX;
    o->maxeval = INT64_MAX;
    o->xatol = 0.0;
    o->xrtol = 1e-7;
    o->fatol = 0.0;
    o->frtol = 1e-8;
    o->gatol = 0.0;
    o->grtol = 1e-6;
    o->epsilon = 0.0;
    o->lnsrch = 0;
    o->ls_ftol = 1e-3;
    o->ls_gtol = 0.9;
    o->ls_xtol = 0.1;
    o->ls_gamma1 = 0.1;
    o->ls_gamma2 = 0.5;
    o->ls_strict = 0;
}

/* ------------------------------------------------------------------------ */
/* Type-generic part, instantiated twice.                                    */

#define REAL double
#define FN(name) opko_d_##name
#define REAL_IS_DOUBLE 1
#include "opk_oracle_impl.h"
#undef REAL
#undef FN
#undef REAL_IS_DOUBLE

#define REAL float
#define FN(name) opko_f_##name
#define REAL_IS_DOUBLE 0
#include "opk_oracle_impl.h"
#undef REAL
#undef FN
#undef REAL_IS_DOUBLE
