// opkb_driver.cu -- native reverse-communication driver of VMLMB:
// the stage machine of `_vmlmb!` (src/quasinewton.jl:296-612), the three line
// searches (src/linesearches.jl) and the coefficient-space two-loop solver, in
// host C++ on top of the fused phases of opkb_vmlmb.cu.  It is the same logic
// as the Python twin (optimpacknextgen.jl_b200/vmlmb.py) and the Julia module;
// a C, Julia or Fortran host needs one call per objective evaluation, in the
// style of the reference's own C FFI (`opk_iterate_vmlmb`, `OPK_TASK_*`,
// src/wrappers.jl:758-765, :1831-1859), or ONE call for a built-in objective.
#include <math.h>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "opkb_internal.cuh"
#include "opkb_chain.cuh"

// ---------------------------------------------------------------------------
// line searches (scalar, Float64): src/linesearches.jl
namespace {

enum { LS_ARMIJO = 1, LS_MORE_TORALDO = 2, LS_MORE_THUENTE = 3 };
enum { T_START = 0, T_SEARCH = 1, T_CONVERGENCE = 2, T_WARNING = 3, T_ERROR = 4 };

inline double jl_max(double a, double b) { return a != a ? a : (b != b ? b : (a > b ? a : b)); }
inline double jl_min(double a, double b) { return a != a ? a : (b != b ? b : (a < b ? a : b)); }
inline double jl_clamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

struct LineSearch {
    int kind = LS_MORE_TORALDO, task = T_START;
    const char* reason = "Algorithm not yet started";   // REASONS, :33-51
    double ftol = 1e-3, gtol = 0.9, xtol = 0.1, gamma1 = 0.1, gamma2 = 0.5;
    int strict = 0;
    double finit = 0, ginit = 0, stp = 0, stpmin = 0, stpmax = 0;
    double stx = 0, fx = 0, gx = 0, sty = 0, fy = 0, gy = 0, lower = 0, upper = 0, width = 0, oldwidth = 0;
    int stage = 0, brackt = 0;

    bool usederivatives() const { return kind == LS_MORE_THUENTE; }   // :230, :329, :453
    int report(int t, const char* r) { task = t; reason = r; return t; }

    int start(double f0, double g0, double stp_, double stpmin_, double stpmax_)
    {
        // :331-358 and :470-529; a thrown ArgumentError is T_ERROR here
        if (stpmax_ < stpmin_) return report(T_ERROR, "STPMAX < STPMIN");
        if (stpmin_ < 0) return report(T_ERROR, "STPMIN < 0");
        if (kind == LS_MORE_THUENTE) {
            if (xtol < 0) return report(T_ERROR, "XTOL < 0");
            if (ftol <= 0) return report(T_ERROR, "FTOL \xE2\x89\xA4 0");
            if (gtol <= 0) return report(T_ERROR, "GTOL \xE2\x89\xA4 0");
            if (g0 >= 0) return report(T_ERROR, "Not a descent direction");
            if (stp_ > stpmax_) return report(T_ERROR, "STP > STPMAX");
            if (stp_ < stpmin_) return report(T_ERROR, "STP < STPMIN");
            stpmin = stpmin_; stpmax = stpmax_; finit = f0; ginit = g0; stp = stp_;
            stx = 0; fx = f0; gx = g0; sty = 0; fy = f0; gy = g0;
            lower = 0; upper = stp_ + stp_ * 4.0;                      // XTRAPU :457
            width = stpmax_ - stpmin_; oldwidth = 2 * (stpmax_ - stpmin_);
            stage = 1; brackt = 0;
            return report(T_SEARCH, "Caller must compute f(x) and g(x)");
        }
        if (stp_ > stpmax_) return report(T_ERROR, "STP > STPMAX");
        if (stp_ < stpmin_) return report(T_ERROR, "STP < STPMIN");
        if (g0 >= 0) return report(T_ERROR, "Not a descent direction");
        finit = f0; ginit = g0; stp = stp_; stpmin = stpmin_;
        return report(T_SEARCH, "Caller must compute f(x) and g(x)");
    }

    // cstep :709-870; returns false for a thrown error
    static bool cstep(int& brackt, double stpmin, double stpmax, double& stx, double& fx, double& dx,
                      double& sty, double& fy, double& dy, double& stp, double fp, double dp,
                      const char*& err)
    {
        double theta, gamma, p, q, r, s, t, stpc, stpq, stpf;
        if (brackt && (stp <= jl_min(stx, sty) || stp >= jl_max(stx, sty))) {
            err = "STP outside bracket (STX,STY)";
            return false;
        }
        if (dx * (stp - stx) >= 0) { err = "descent condition violated"; return false; }
        if (stpmax < stpmin) { err = "STPMAX < STPMIN"; return false; }
        const bool opposite = ((dx < 0 && dp > 0) || (dx > 0 && dp < 0));
        if (fp > fx) {
            theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
            s = jl_max(jl_max(fabs(theta), fabs(dx)), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp < stx) gamma = -gamma;
            p = (gamma - dx) + theta;
            q = ((gamma - dx) + gamma) + dp;
            r = p / q;
            stpc = stx + r * (stp - stx);
            stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2) * (stp - stx);
            stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2;
            brackt = 1;
        } else if (opposite) {
            theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
            s = jl_max(jl_max(fabs(theta), fabs(dx)), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp > stx) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dx;
            r = p / q;
            stpc = stp + r * (stx - stp);
            stpq = stp + (dp / (dp - dx)) * (stx - stp);
            stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
            brackt = 1;
        } else if (fabs(dp) < fabs(dx)) {
            theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
            s = jl_max(jl_max(fabs(theta), fabs(dx)), fabs(dp));
            t = (theta / s) * (theta / s) - (dx / s) * (dp / s);
            gamma = (t > 0 ? s * sqrt(t) : 0.0);
            if (stp > stx) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = (gamma + (dx - dp)) + gamma;
            r = p / q;
            if (r < 0 && gamma != 0) stpc = stp + r * (stx - stp);
            else if (stp > stx) stpc = stpmax;
            else stpc = stpmin;
            stpq = stp + (dp / (dp - dx)) * (stx - stp);
            if (brackt) {
                stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
                t = stp + 0.66 * (sty - stp);
                if (stp > stx ? stpf > t : stpf < t) stpf = t;
            } else {
                stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
                stpf = jl_max(stpmin, jl_min(stpf, stpmax));
            }
        } else {
            if (brackt) {
                theta = 3 * (fp - fy) / (sty - stp) + dy + dp;
                s = jl_max(jl_max(fabs(theta), fabs(dy)), fabs(dp));
                gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
                if (stp > sty) gamma = -gamma;
                p = (gamma - dp) + theta;
                q = ((gamma - dp) + gamma) + dy;
                r = p / q;
                stpf = stp + r * (sty - stp);
            } else if (stp > stx) {
                stpf = stpmax;
            } else {
                stpf = stpmin;
            }
        }
        if (fp > fx) {
            sty = stp; fy = fp; dy = dp;
        } else {
            if (opposite) { sty = stx; fy = fx; dy = dx; }
            stx = stp; fx = fp; dx = dp;
        }
        stp = stpf;
        return true;
    }

    int iterate(double stp_, double f, double g)
    {
        if (task != T_SEARCH || stp_ != stp) return report(T_ERROR, "iterate! called out of sequence");
        if (kind != LS_MORE_THUENTE) {
            // Armijo :235-247, More & Toraldo :360-381
            if (f <= finit + ftol * stp_ * ginit)
                return report(T_CONVERGENCE, "First Wolfe condition satisfied");
            if (stp_ > stpmin) {
                double nstp = stp_;
                if (kind == LS_ARMIJO) {
                    nstp = stp_ / 2;
                } else {
                    const double q = -ginit * stp_, r = f - (finit - q);
                    if (strict ? (r > 0 && q > 0) : (r != 0)) nstp = stp_ * jl_clamp(q / (r + r), gamma1, gamma2);
                    else nstp = stp_ / 2;
                }
                stp = jl_max(nstp, stpmin);
                return report(T_SEARCH, "Caller must compute f(x) and g(x)");
            }
            stp = stpmin;
            return report(T_WARNING, "Step at lower bound");
        }
        // More & Thuente :536-647
        const double gtest = ftol * ginit, ftest = finit + stp_ * gtest;
        double nstp = stp_;
        const char* err = "";
        if (stage == 1 && f <= ftest && g >= 0) stage = 2;
        if (f <= ftest && fabs(g) <= -gtol * ginit)
            return report(T_CONVERGENCE, "Strong Wolfe conditions both satisfied");
        if (stp_ == stpmin && (f > ftest || g >= gtest)) return report(T_WARNING, "Step at lower bound");
        if (stp_ == stpmax && f <= ftest && g <= gtest) return report(T_WARNING, "Step at upper bound");
        if (brackt) {
            if (upper - lower <= xtol * upper) return report(T_WARNING, "XTOL test satisfied");
            if (stp_ <= lower || stp_ >= upper) return report(T_WARNING, "Rounding errors prevent progress");
        }
        if (stage == 1 && f <= fx && f > ftest) {
            double fxm = fx - stx * gtest, gxm = gx - gtest, fym = fy - sty * gtest, gym = gy - gtest;
            if (!cstep(brackt, lower, upper, stx, fxm, gxm, sty, fym, gym, nstp, f - stp_ * gtest, g - gtest, err))
                return report(T_ERROR, err);
            fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
        } else {
            if (!cstep(brackt, lower, upper, stx, fx, gx, sty, fy, gy, nstp, f, g, err)) return report(T_ERROR, err);
        }
        if (brackt) {
            const double wcur = fabs(sty - stx);
            if (wcur >= 0.66 * oldwidth) nstp = stx + 0.5 * (sty - stx);
            oldwidth = width;
            width = wcur;
        }
        if (brackt) {
            if (stx <= sty) { lower = stx; upper = sty; } else { lower = sty; upper = stx; }
        } else {
            lower = nstp + 1.1 * (nstp - stx);   // XTRAPL :456
            upper = nstp + 4.0 * (nstp - stx);
        }
        nstp = jl_clamp(nstp, stpmin, stpmax);
        if (brackt && (nstp <= lower || nstp >= upper || upper - lower <= xtol * upper)) nstp = stx;
        stp = nstp;
        return report(T_SEARCH, "Caller must compute f(x) and g(x)");
    }
};

}  // namespace

// ---------------------------------------------------------------------------
struct opkb_solver_ {
    opkb_vmlmb* ws;
    opkb_vmlmb_options opt;
    LineSearch ls;
    int mem, method, eltype, builtin, twoloop;   // twoloop: 0 gram, 1 sequential
    int fminset;
    // _vmlmb! state (:329-379)
    int mark, m, stage;
    int64_t iter, eval, rejects;
    double f, f0, gd, gd0, stp, stpmin, stpmax, smin, smax, gamma, gnorm, gtest;
    double beststp, bestf, bestgnorm;
    char reason[96];
    int finished, awaiting_f, new_x;
    // fused-phase cache
    double r[8], limits[2], spec[8];
    int saved_for, need_steepest, spec_valid, spec_mark;
    double rho[OPKB_SLOTS_MAX], alpha[OPKB_SLOTS_MAX];
    // device-callback objective (opkb_solver_set_fg_callback)
    opkb_fg_callback fg;
    void* fg_user;
    int fg_partial;
    // device-resident decisions (opkb_chain.cuh): the steady-state launches of opkb_solver_run are
    // enqueued one iteration ahead and take their coefficients from `chain_dev`
    int chain_on;
    ChainDev* chain_dev;
    ChainDev* chain_stage;      // pinned: what a head launch sends to chain_dev
    int64_t run_target;
};

extern "C" void opkb_vmlmb_default_options(opkb_vmlmb_options* o)
{
    if (!o) return;
    memset(o, 0, sizeof(*o));   // src/quasinewton.jl:227-242
    o->blmvm = 0;
    o->fmin = -INFINITY;
    o->maxiter = INT64_MAX;
    o->maxeval = INT64_MAX;
    o->xatol = 0.0;  o->xrtol = 1e-7;
    o->fatol = 0.0;  o->frtol = 1e-8;
    o->gatol = 0.0;  o->grtol = 1e-6;
    o->epsilon = 0.0;
    o->lnsrch = 0;
    o->ls_ftol = 1e-3;  o->ls_gtol = 0.9;  o->ls_xtol = 0.1;
    o->ls_gamma1 = 0.1;  o->ls_gamma2 = 0.5;  o->ls_strict = 0;
    o->twoloop = -1;
    o->speculate = 1;
}

// accessors of the workspace needed here (opkb_vmlmb.cu)
int opkb_vmlmb_info(const opkb_vmlmb* ws, int* mem, int* method, int* eltype, int* obj);
int opkb_vmlmb_chain_possible(const opkb_vmlmb* ws);
int opkb_vmlmb_chain_pending(const opkb_vmlmb* ws);
int opkb_vmlmb_direction_trial_chain(opkb_vmlmb* ws, int m, const int* slots, const double* alpha,
                                     const double* c, double gamma, int mark_next, double out[4],
                                     double trial[8], const ChainDev* state, ChainDev* dev, ChainDev* stage,
                                     int prelaunch_next);
int opkb_vmlmb_trial_end_mode(opkb_vmlmb* ws, int mark, double f, int initial, int f_mode, double out[8]);

extern "C" int opkb_solver_create(opkb_vmlmb* ws, const opkb_vmlmb_options* opt, opkb_solver** out)
{
    if (!ws || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    *out = NULL;
    opkb_vmlmb_options o;
    if (opt) o = *opt; else opkb_vmlmb_default_options(&o);
    // the @assert's of :310-319
    if (!(o.maxiter >= 0 && o.maxeval >= 1 && o.xatol >= 0 && o.xrtol >= 0 && o.fatol >= 0 && o.frtol >= 0 &&
          o.gatol >= 0 && o.grtol >= 0 && o.epsilon >= 0 && o.epsilon < 1))
        return opkb_set_error(OPKB_EINVAL, "invalid option");
    opkb_solver* s = (opkb_solver*)calloc(1, sizeof(opkb_solver));
    if (!s) return opkb_set_error(OPKB_ENOMEM, "out of host memory");
    new (&s->ls) LineSearch();
    s->ws = ws;
    s->opt = o;
    int obj = 0;
    opkb_vmlmb_info(ws, &s->mem, &s->method, &s->eltype, &obj);
    s->builtin = (obj != OPKB_OBJ_USER);
    s->twoloop = (o.twoloop < 0) ? (s->eltype == OPKB_DOUBLE ? 0 : 1) : (o.twoloop ? 1 : 0);
    // the reference accepts any `mem` (src/quasinewton.jl:227, :328); the Gram-form kernels hold
    // OPKB_MEM_MAX pairs, beyond that the step-by-step recursion (no such limit) runs
    if (s->mem > OPKB_MEM_MAX) {
        if (o.twoloop == 0) { free(s); return opkb_set_error(OPKB_EINVAL, "twoloop = Gram needs mem <= %d", OPKB_MEM_MAX); }
        s->twoloop = 1;
    }
    int kind = o.lnsrch;                                        // :276-284
    if (kind == 0) kind = (s->method == 0 ? LS_MORE_THUENTE : LS_MORE_TORALDO);
    if (kind < 1 || kind > 3) { free(s); return opkb_set_error(OPKB_EINVAL, "invalid line search"); }
    s->ls.kind = kind;
    s->ls.ftol = o.ls_ftol;  s->ls.gtol = o.ls_gtol;  s->ls.xtol = o.ls_xtol;
    s->ls.gamma1 = o.ls_gamma1;  s->ls.gamma2 = o.ls_gamma2;  s->ls.strict = o.ls_strict;
    s->fminset = (!(o.fmin != o.fmin) && o.fmin > -INFINITY);   // :330
    s->gamma = 1;
    s->saved_for = -1;
    s->limits[0] = s->limits[1] = INFINITY;
    s->reason[0] = 0;
    s->run_target = INT64_MAX;
    // Device-resident decisions (opkb_chain.cuh): opt-in, opkb_solver_set_chain or OPKB_CHAIN=1 in the
    // environment (read per solver).  Measured on a B200 (profiles/r02u_*): the host round trip it removes
    // is only ~8 us per iteration (polled pinned flag, results written straight to host memory), while
    // the chained kernel is ~12 us longer: C1 18.9 k -> 18.2 k it/s (profiles/r02c2_*).  Same bits either way.
    if (getenv("OPKB_CHAIN") && atoi(getenv("OPKB_CHAIN")) > 0) (void)opkb_solver_set_chain(s, 1);
    *out = s;
    return 0;
}

// Switch the device-resident decisions of opkb_solver_run on or off (include/opkb200.h).  Returns 0 also
// when the solver cannot use them (user objective, Float32, more than 8 pairs, BLMVM, the step-by-step
// two-loop, several ranks): `*enabled` (may be NULL) tells what is in effect.
extern "C" int opkb_solver_set_chain(opkb_solver* s, int on)
{
    if (!s) return opkb_set_error(OPKB_EINVAL, "NULL solver");
    if (opkb_vmlmb_chain_pending(s->ws)) return opkb_set_error(OPKB_EINVAL, "a chained launch is in flight");
    if (!on) {
        s->chain_on = 0;
        return 0;
    }
    const opkb_vmlmb_options& o = s->opt;
    const bool ls_ok = (s->ls.kind != LS_MORE_THUENTE) || (o.ls_ftol > 0 && o.ls_gtol > 0 && o.ls_xtol >= 0);
    if (!(ls_ok && s->builtin && o.speculate && s->twoloop == 0 && opkb_vmlmb_chain_possible(s->ws))) return 0;
    if (!s->chain_dev) {
        opkb_bind_device(opkb_vmlmb_vector(s->ws, 'x', 0)->ctx);     // (the structure lives on the workspace's GPU)
        if (cudaMalloc((void**)&s->chain_dev, sizeof(ChainDev)) != cudaSuccess ||
            cudaMallocHost((void**)&s->chain_stage, sizeof(ChainDev)) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
    }
    s->chain_on = 1;
    return 0;
}

extern "C" int opkb_solver_chain_enabled(const opkb_solver* s) { return s ? s->chain_on : 0; }

extern "C" int opkb_solver_destroy(opkb_solver* s)
{
    if (s && s->chain_dev) cudaFree(s->chain_dev);   // (synchronises the device: no launch reads it any more)
    if (s && s->chain_stage) cudaFreeHost(s->chain_stage);
    free(s);
    return 0;
}

static void set_reason(opkb_solver* s, const char* r)
{
    strncpy(s->reason, r, sizeof(s->reason) - 1);
    s->reason[sizeof(s->reason) - 1] = 0;
}

// two-loop recursion in coefficient space from the Gram block (apply_lbfgs!, :716-779)
static bool solve_two_loop(opkb_solver* s, const double* G, int m, const int* slots, double* alpha,
                           double* c, double* gamma_out)
{
    const int W = m + 1;
    auto sv = [&](int a) { return G[2 * a * W]; };
    auto yv = [&](int a) { return G[(2 * a + 1) * W]; };
    auto SY = [&](int a, int b) { return G[2 * a * W + 1 + b]; };
    auto YY = [&](int a, int b) { return b >= a ? G[(2 * a + 1) * W + 1 + b] : G[(2 * b + 1) * W + 1 + a]; };
    double rho[OPKB_MEM_MAX];
    if (s->method == 2) {
        for (int j = 0; j < m; ++j) rho[j] = SY(j, j);                      // :758
    } else {
        const int nw = m - 1;
        s->rho[slots[nw]] = SY(nw, nw);                                     // :515
        if (s->rho[slots[nw]] > 0) s->gamma = s->rho[slots[nw]] / YY(nw, nw);  // :518
        for (int j = 0; j < m; ++j) rho[j] = s->rho[slots[j]];
    }
    double gamma = 0.0;
    bool modif = false;
    for (int k = m - 1; k >= 0; --k) {                                      // :725-732 / :756-766
        alpha[k] = NAN;
        c[k] = 0.0;
        if (rho[k] > 0) {
            double acc = sv(k);
            for (int j = m - 1; j > k; --j)
                if (rho[j] > 0) acc -= alpha[j] * SY(k, j);
            alpha[k] = acc / rho[k];
            if (gamma == 0) gamma = rho[k] / YY(k, k);                      // :762-764
            modif = true;
        }
    }
    if (s->method < 2) gamma = s->gamma;                                    // :526
    *gamma_out = gamma;
    if (!modif || gamma == 0) return false;
    for (int k = 0; k < m; ++k) {                                           // :735-741 / :769-775
        if (rho[k] > 0) {
            double acc = yv(k);
            for (int j = m - 1; j >= 0; --j)
                if (rho[j] > 0) acc -= alpha[j] * YY(k, j);
            acc *= gamma;
            for (int j = 0; j < k; ++j)
                if (rho[j] > 0) acc += c[j] * SY(j, k);
            c[k] = alpha[k] - acc / rho[k];
        }
    }
    return true;
}

static int new_direction_gram(opkb_solver* s, bool* change, double* gd, double* dnorm)
{
    const int mem = s->mem, mark = s->mark;
    const int m = (s->m + 1 < mem ? s->m + 1 : mem);                        // :521
    int slots[OPKB_MEM_MAX];
    for (int j = 0; j < m; ++j) slots[j] = ((mark - (m - 1) + j) % mem + mem) % mem;
    static thread_local double G[2 * OPKB_MEM_MAX * (OPKB_MEM_MAX + 1)];
    int rc = opkb_vmlmb_gram(s->ws, mark, m, slots, G);
    if (rc) return rc;
    s->m = m;
    double alpha[OPKB_MEM_MAX], c[OPKB_MEM_MAX], gamma = 0;
    *change = solve_two_loop(s, G, m, slots, alpha, c, &gamma);
    s->saved_for = -1;
    if (!*change) return 0;
    const int mark_next = (mark < mem - 1 ? mark + 1 : 0);
    double out4[4];
    if (s->opt.speculate && s->builtin && s->method != 1) {
        if (s->chain_on) {
            // the solver's scalars as the last CTA needs them to take this launch's decisions
            ChainDev st;
            memset(&st, 0, sizeof(st));
            st.ls_kind = s->ls.kind;
            st.fminset = s->fminset;
            st.epsilon = s->opt.epsilon;
            st.fmin = s->opt.fmin;
            st.ls_ftol = s->ls.ftol;
            st.ls_gtol = s->ls.gtol;
            st.xatol = s->opt.xatol;
            st.xrtol = s->opt.xrtol;
            st.fatol = s->opt.fatol;
            st.frtol = s->opt.frtol;
            st.maxiter = s->opt.maxiter;
            st.maxeval = s->opt.maxeval;
            st.iter = s->iter;
            st.eval = s->eval;
            st.f = s->f;
            st.gnorm = s->gnorm;
            st.bestf = s->bestf;
            st.gtest = s->gtest;
            st.gamma = s->gamma;
            for (int k = 0; k < mem && k < OPKB_CHAIN_MB; ++k) st.rho[k] = s->rho[k];
            // the launch of the iteration after this one is enqueued ahead if this run goes that far
            static const int dbg = getenv("OPKB_CHAIN_DBG") ? atoi(getenv("OPKB_CHAIN_DBG")) : 0;   // 4: never ahead
            const int ahead = (s->iter < s->run_target && !(dbg & 4)) ? 1 : 0;
            rc = opkb_vmlmb_direction_trial_chain(s->ws, m, slots, alpha, c, gamma, mark_next, out4, s->spec,
                                                  &st, s->chain_dev, s->chain_stage, ahead);
        } else {
            rc = opkb_vmlmb_direction_trial(s->ws, m, slots, alpha, c, gamma, mark_next, out4, s->spec);
        }
        if (rc) return rc;
        s->spec_valid = (s->spec[7] == 1.0);
        s->spec_mark = mark_next;
    } else {
        rc = opkb_vmlmb_direction(s->ws, m, slots, alpha, c, gamma, mark_next, out4);
        if (rc) return rc;
    }
    *gd = -out4[0];
    *dnorm = sqrt(out4[1]);
    s->limits[0] = out4[2];
    s->limits[1] = out4[3];
    s->saved_for = mark_next;
    s->need_steepest = 0;
    return 0;
}

static int new_direction_sequential(opkb_solver* s, bool* change, double* gd, double* dnorm)
{
    const int mem = s->mem, mark = s->mark, method = s->method;
    double out2[2], out3[3], out4[4];
    int rc = opkb_vmlmb_pair_update(s->ws, mark, out2);                     // :512-513
    if (rc) return rc;
    double* rho = s->rho;
    double* alpha = s->alpha;
    if (method < 2) {
        rho[mark] = out2[0];                                                // :515
        if (rho[mark] > 0) s->gamma = rho[mark] / out2[1];                  // :518
    }
    s->m = (s->m + 1 < mem ? s->m + 1 : mem);
    const int m = s->m;
    int first = 1, pw = 0, ps = 0;
    double pa = 0.0, gamma = 0.0;
    bool modif = false;
    int k = mark + 1;
    for (int i = 0; i < m; ++i) {                                           // :725-732 / :756-766
        k = (k > 0 ? k - 1 : mem - 1);
        if (method < 2 && !(rho[k] > 0)) continue;
        rc = opkb_vmlmb_twoloop_step(s->ws, first, pw, ps, pa, 0, 1.0, 'S', k, method == 2, out3);
        if (rc) return rc;
        if (pw) first = 0;
        pw = 0;
        if (method == 2) rho[k] = out3[1];                                  // :758
        if (rho[k] > 0) {
            alpha[k] = out3[0] / rho[k];                                    // :728 / :760
            pw = 'Y'; ps = k; pa = -alpha[k];                               // :729 / :761
            modif = true;
            if (method == 2 && gamma == 0) gamma = rho[k] / out3[2];        // :762-764
        }
    }
    s->saved_for = -1;
    if (method < 2) { gamma = s->gamma; *change = modif; } else { *change = (gamma != 0); }
    if (!*change) return 0;
    int scale = 1;
    for (int i = 0; i < m; ++i) {                                           // :735-741 / :769-775
        if (rho[k] > 0) {
            rc = opkb_vmlmb_twoloop_step(s->ws, first, pw, ps, pa, scale, gamma, 'Y', k, 0, out3);
            if (rc) return rc;
            first = 0;
            scale = 0;
            const double beta = out3[0] / rho[k];                           // :737 / :771
            pw = 'S'; ps = k; pa = alpha[k] - beta;                         // :738 / :772
        }
        k = (k < mem - 1 ? k + 1 : 0);
    }
    const int mark_next = (mark < mem - 1 ? mark + 1 : 0);
    rc = opkb_vmlmb_finish_direction(s->ws, mark_next, pw, ps, pa, out4);
    if (rc) return rc;
    *gd = -out4[0];
    *dnorm = sqrt(out4[1]);
    s->limits[0] = out4[2];
    s->limits[1] = out4[3];
    s->saved_for = mark_next;
    s->need_steepest = 0;
    return 0;
}

// the part of a loop trip AFTER the objective has been evaluated (:394-601);
// on return either the solver is finished or s->stp is the next step to try
static int after_evaluation(opkb_solver* s)
{
    const opkb_vmlmb_options& o = s->opt;
    LineSearch& ls = s->ls;
    s->f = s->r[0];
    s->eval += 1;
    s->new_x = 0;
    if (s->eval == 1 || s->f < s->bestf) {                                  // :398-419
        s->gnorm = sqrt(s->r[1]);
        s->beststp = s->stp;
        s->bestf = s->f;
        s->bestgnorm = s->gnorm;
        if (s->eval == 1) s->gtest = jl_max(o.gatol, o.grtol * s->gnorm);
        if (s->gnorm <= s->gtest) {
            s->stage = 3;
            set_reason(s, s->gnorm == 0 ? "a stationary point has been found"
                          : (s->method > 0 ? "projected gradient norm sufficiently small"
                                           : "gradient norm sufficiently small"));
            if (s->eval > 1) s->iter += 1;
        }
    }
    if (s->fminset && s->f < o.fmin) {                                      // :420-423
        s->stage = 4;
        set_reason(s, "f < fmin");
    }
    if (s->stage == 1) {                                                    // :425-474
        if (ls.usederivatives()) s->gd = (s->method > 0 ? s->r[2] / s->stp : -s->r[3]);
        const int task = ls.iterate(s->stp, s->f, s->gd);
        if (task == T_SEARCH) {
            s->stp = ls.stp;
        } else if (task == T_CONVERGENCE) {
            s->iter += 1;
            double xtest = jl_max(o.xatol, 0.0);
            if (o.xrtol > 0) xtest = jl_max(xtest, o.xrtol * sqrt(s->r[4]));
            if (sqrt(s->r[5]) <= xtest) {
                s->stage = 3;
                set_reason(s, "X test satisfied");
            } else {
                const double delta = jl_max(fabs(s->f - s->f0), s->stp * fabs(s->gd0));
                if (delta <= o.fatol) { s->stage = 3; set_reason(s, "fatol test satisfied"); }
                else if (delta <= o.frtol * fabs(s->f0)) { s->stage = 3; set_reason(s, "frtol test satisfied"); }
                else if (s->iter >= o.maxiter) { s->stage = 4; set_reason(s, "too many iterations"); }
                else if (s->eval >= o.maxeval) { s->stage = 4; set_reason(s, "too many evaluations"); }
                else s->stage = 2;
            }
        } else if (task == T_ERROR) {
            s->stage = 4;
            set_reason(s, ls.reason);
            s->finished = 1;
            return opkb_set_error(OPKB_EINVAL, "line search: %s", ls.reason);
        } else {
            s->stage = 4;
            set_reason(s, ls.reason);
        }
    }
    if (s->stage < 2 && s->eval >= o.maxeval) {                             // :475-478
        s->stage = 4;
        set_reason(s, "too many evaluations");
    }
    if (s->stage == 1) return 0;

    // :480-596
    if (s->stp != s->beststp) {
        if (s->stage >= 3) {
            s->stp = s->beststp;
            s->f = s->bestf;
            s->gnorm = s->bestgnorm;
            int rc = opkb_vmlmb_revert(s->ws, s->mark, s->stp);             // :490-492
            if (rc) return rc;
        } else {
            s->gnorm = sqrt(s->r[1]);
        }
    }
    s->new_x = 1;                                                           // the `printer` site :501
    if (s->stage >= 3) {
        s->finished = 1;
        return 0;
    }
    if (s->stage == 2) {                                                    // :509-550
        bool change = false;
        double gd = 0, dnorm = 0;
        int rc = s->twoloop ? new_direction_sequential(s, &change, &gd, &dnorm)
                            : new_direction_gram(s, &change, &gd, &dnorm);
        if (rc) return rc;
        bool reject = true;
        if (change) {
            s->gd = gd;
            reject = !(o.epsilon > 0 ? (s->gd <= -o.epsilon * dnorm * s->gnorm) : (s->gd < 0));  // :631-632
        }
        if (reject) {
            s->stage = 0;
            s->rejects += 1;
        }
        s->mark = (s->mark < s->mem - 1 ? s->mark + 1 : 0);                 // :549
    }
    if (s->stage == 0) {                                                    // :551-556
        if (s->spec_valid) {   // the rejected direction's speculative trial is dropped
            int rc = opkb_vmlmb_discard_trial(s->ws, s->spec_mark);
            if (rc) return rc;
            s->spec_valid = 0;
        }
        s->need_steepest = 1;
        s->gd = -s->gnorm * s->gnorm;
    }
    s->f0 = s->f;                                                           // :560-563
    s->gd0 = s->gd;
    if (s->need_steepest || s->saved_for != s->mark) {
        double out4[4];
        int rc = opkb_vmlmb_steepest(s->ws, s->mark, out4);                 // :554, :562-563, :577
        if (rc) return rc;
        s->limits[0] = out4[2];
        s->limits[1] = out4[3];
        s->need_steepest = 0;
    }
    s->saved_for = -1;
    s->smin = s->limits[0];
    s->smax = s->limits[1];
    if (s->stage == 2) s->stp = 1.0;                                        // :566-574
    else if (s->fminset && o.fmin < s->f0) s->stp = 2 * (o.fmin - s->f0) / s->gd0;
    else s->stp = 1 / s->gnorm;
    if (s->method > 0) {                                                    // :575-584
        s->stp = jl_min(s->stp, s->smax);
        s->stpmin = 1e-20 * s->stp;
        s->stpmax = jl_min(1e+20 * s->stp, s->smax);
    } else {
        s->stpmin = 1e-20 * s->stp;
        s->stpmax = 1e+20 * s->stp;
    }
    const int task = ls.start(s->f0, s->gd0, s->stp, s->stpmin, s->stpmax);  // :587-594
    if (task != T_SEARCH) {
        s->stage = 4;
        set_reason(s, ls.reason);
        s->finished = 1;
        return opkb_set_error(OPKB_EINVAL, "line search: %s", ls.reason);
    }
    s->stage = 1;
    return 0;
}

static int task_of(const opkb_solver* s)
{
    if (s->finished) return s->stage > 3 ? OPKB_TASK_WARNING : OPKB_TASK_FINAL_X;
    return OPKB_TASK_COMPUTE_FG;
}

// Form the next point to evaluate: x = project(S[mark] - stp*d) (:599, :386).
extern "C" int opkb_solver_start(opkb_solver* s, int* task)
{
    if (!s || !task) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (s->eval != 0 || s->awaiting_f) return opkb_set_error(OPKB_EINVAL, "solver already started");
    int rc = opkb_vmlmb_trial_begin(s->ws, s->mark, 0.0, 1);
    if (rc) return rc;
    s->awaiting_f = 1;
    *task = OPKB_TASK_COMPUTE_FG;
    return 0;
}

// The caller has written g (workspace vector 'g') and passes f = f(x).
static int solver_iterate(opkb_solver* s, double f, int f_mode, int* task)
{
    if (!s || !task) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (s->finished) { *task = task_of(s); return 0; }
    if (!s->awaiting_f) return opkb_set_error(OPKB_EINVAL, "call opkb_solver_start first");
    int rc = opkb_vmlmb_trial_end_mode(s->ws, s->mark, f, s->eval == 0, f_mode, s->r);
    if (rc) return rc;
    s->awaiting_f = 0;
    rc = after_evaluation(s);
    if (rc) { *task = OPKB_TASK_ERROR; return rc; }
    if (!s->finished) {
        rc = opkb_vmlmb_trial_begin(s->ws, s->mark, s->stp, 0);
        if (rc) { *task = OPKB_TASK_ERROR; return rc; }
        s->awaiting_f = 1;
    }
    *task = task_of(s);
    return 0;
}

extern "C" int opkb_solver_iterate(opkb_solver* s, double f, int* task)
{
    return solver_iterate(s, f, 1, task);
}

extern "C" int opkb_solver_set_fg_callback(opkb_solver* s, opkb_fg_callback fn, void* user, int f_is_partial)
{
    if (!s) return opkb_set_error(OPKB_EINVAL, "NULL solver");
    if (s->builtin) return opkb_set_error(OPKB_EINVAL, "the workspace has a built-in objective");
    s->fg = fn;
    s->fg_user = user;
    s->fg_partial = f_is_partial ? 1 : 0;
    return 0;
}

// user objective through the device callback: the whole reverse-communication loop in here
static int solver_run_callback(opkb_solver* s, int64_t target, int* task)
{
    int rc = 0;
    if (s->eval == 0 && !s->awaiting_f) {
        rc = opkb_solver_start(s, task);
        if (rc) return rc;
    }
    opkb_context* ctx = opkb_vmlmb_vector(s->ws, 'x', 0)->ctx;
    while (!s->finished && s->iter < target) {
        // the data pointers move between calls (buffers are traded, not copied): asked for every time
        opkb_vector* x = opkb_vmlmb_vector(s->ws, 'x', 0);
        opkb_vector* g = opkb_vmlmb_vector(s->ws, 'g', 0);
        double f = 0.0;
        if (s->fg(s->fg_user, x->n, x->data, g->data, (void*)ctx->stream, &f) != 0) {
            *task = OPKB_TASK_ERROR;
            return opkb_set_error(OPKB_EINVAL, "the objective callback failed");
        }
        rc = solver_iterate(s, f, s->fg_partial ? 2 : 1, task);
        if (rc) return rc;
    }
    *task = task_of(s);
    return 0;
}

// Built-in objective (or device callback): run until `iter` has advanced by niter (or termination).
extern "C" int opkb_solver_run(opkb_solver* s, int64_t niter, int* task)
{
    if (!s || !task) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    const int64_t target = (niter > INT64_MAX - s->iter) ? INT64_MAX : s->iter + niter;
    if (!s->builtin) {
        if (!s->fg)
            return opkb_set_error(OPKB_EINVAL, "opkb_solver_run needs a built-in objective or a device callback "
                                  "(opkb_solver_set_fg_callback)");
        return solver_run_callback(s, target, task);
    }
    s->run_target = target;
    while (!s->finished && s->iter < target) {
        int rc = 0;
        if (s->spec_valid && s->eval > 0 && s->spec_mark == s->mark && s->stp == 1.0) {
            memcpy(s->r, s->spec, sizeof(s->r));   // evaluated by the direction kernel
            s->spec_valid = 0;
        } else {
            s->spec_valid = 0;
            rc = opkb_vmlmb_trial(s->ws, s->mark, s->stp, s->eval == 0, s->r);
        }
        if (!rc) rc = after_evaluation(s);
        if (rc) { *task = OPKB_TASK_ERROR; return rc; }
    }
    if (s->chain_on && opkb_vmlmb_chain_pending(s->ws)) {
        *task = OPKB_TASK_ERROR;
        return opkb_set_error(OPKB_EINVAL, "internal error: a chained launch is still in flight at the end of the run");
    }
    *task = task_of(s);
    return 0;
}

extern "C" int opkb_solver_status(const opkb_solver* s, int64_t* iter, int64_t* eval, int64_t* rejects,
                                  double* f, double* gnorm, double* stp, int* stage, int* new_x)
{
    if (!s) return opkb_set_error(OPKB_EINVAL, "NULL solver");
    if (iter) *iter = s->iter;
    if (eval) *eval = s->eval;
    if (rejects) *rejects = s->rejects;
    if (f) *f = s->f;
    if (gnorm) *gnorm = s->gnorm;
    if (stp) *stp = s->stp;
    if (stage) *stage = s->stage;
    if (new_x) *new_x = s->new_x;
    return 0;
}

extern "C" const char* opkb_solver_reason(const opkb_solver* s) { return s ? s->reason : ""; }
extern "C" int opkb_solver_mark(const opkb_solver* s) { return s ? s->mark : -1; }
