// opkb_chain.cuh -- device-resident decisions of the steady-state VMLMB / L-BFGS iteration.
//
// In steady state an iteration of the native driver (opkb_driver.cu) is ONE launch of the fused
// kernel k_direction2<.., CARRY> (direction + first trial point of the next line search + the next
// Gram block) followed by a few hundred host flops: is the direction a descent direction
// (src/quasinewton.jl:542-546), is the step the driver tries exactly 1 (:566-584), does the line
// search accept it at once (src/linesearches.jl:235-247, :360-381, :536-547), do the convergence
// tests let the loop go on (:398-423, :437-468), and what are the two-loop coefficients of the next
// direction (apply_lbfgs!, :716-779) -- and then the next launch.  For small problems that round
// trip (completion flag over PCIe, host, launch) is a third of the iteration (C1: 20 of 59 us).
//
// CHAIN instances of the kernel take it out of the critical path: the last CTA, once it has combined
// the partials, runs chain_decide() below -- the very same tests and the same coefficient-space
// two-loop recursion, in the same order, on the same Float64 scalars -- and leaves `go` and the
// coefficients of the NEXT direction in a small device structure.  The host has ALREADY enqueued
// that next launch (pointers rotated as an accepted step rotates them); its prologue reads the
// structure: go = 0 -> every CTA returns at once and nothing has been touched, the host goes on as
// it always did (another step length, a rejected direction, convergence ...).  The host remains the
// authority: it receives the reduced scalars of every launch as before, runs its own stage machine
// on them, and ADOPTS the launch that is already in flight only if its own decision and its own
// coefficients are bit-identical to what the device published (anything else is an internal error,
// reported loudly).  Trajectories are therefore the same bits with and without the chain.
#pragma once
#include <math.h>

#define OPKB_CHAIN_MB 8    // pairs (and slots) the chained launches support

struct ChainDev {
    // ---- what the next chained launch reads in its prologue
    int go;                           // 1: run; 0: leave at once
    int m;                            // pairs of that launch (oldest first)
    double alpha[OPKB_CHAIN_MB];      // NaN: the pair does not take part (rho <= 0, :727 / :759)
    double c[OPKB_CHAIN_MB];          // alpha_j - beta_j
    double gamma_dir;
    // ---- configuration (written by the host with every head launch)
    int method, mem, obj, ls_kind;    // ls_kind: 1 Armijo, 2 More-Toraldo, 3 More-Thuente
    int fminset, limits_inf;          // limits_inf: the step limits are +Inf whatever the kernel reduced
    double epsilon, fmin, ls_ftol, ls_gtol, xatol, xrtol, fatol, frtol;
    long long maxiter, maxeval;
    // ---- state of the solver when that launch's results arrive
    long long iter, eval;
    double f, gnorm, bestf, gtest, gamma;   // gamma: the scaling kept between iterations (methods 0, 1; :518)
    double rho[OPKB_CHAIN_MB];        // by SLOT (methods 0, 1; :515)
    int slots[OPKB_CHAIN_MB];         // slots of that launch's pairs
    int mark_next;                    // slot that launch saves (x, g) into = the newest pair of the next one
    int pad;
    double G[2 * OPKB_CHAIN_MB * (OPKB_CHAIN_MB + 1)];   // Gram block of that launch's pairs (opkb_vmlmb_gram)
};

// decision record appended to the reduced scalars of a chained launch (host side: opkb_vmlmb.cu)
//   rec[0] go | rec[1] reason of a no-go | rec[2] gamma | rec[3 .. 3+MB) alpha | rec[3+MB .. 3+2MB) c
#define OPKB_CHAIN_NREC (3 + 2 * OPKB_CHAIN_MB)

#ifdef __CUDACC__
#include <stddef.h>

__device__ __forceinline__ double chain_max(double a, double b) { return a != a ? a : (b != b ? b : (a > b ? a : b)); }
__device__ __forceinline__ double chain_min(double a, double b) { return a != a ? a : (b != b ? b : (a < b ? a : b)); }

// The scalar part, one thread, everything in shared memory.  cs: this launch's copy of the
// structure (updated in place for the next launch, G excepted); R: the combined scalars of this launch
// (k_direction2: 12 + dense + corrections); Gn: the next Gram block, already assembled (below).
// Mirrors, statement by statement, what the host does between two fused launches:
// new_direction_gram (tail) -> after_evaluation (tail) -> opkb_solver_run -> after_evaluation (head)
// -> opkb_vmlmb_gram (gram_from_carry) -> solve_two_loop of opkb_driver.cu / opkb_vmlmb.cu.
template <int MB>
__device__ __forceinline__ void chain_decide_serial(ChainDev* cs, const double* R, const double* G, double* rec)
{
    constexpr int ND = 4 * MB + 4, NENT = MB * (MB + 1), NCORR = 32 * ((NENT + 31) / 32);
    // the NEXT launch works on up to MB + 1 pairs (this one held m <= MB)
    constexpr int MN = (MB + 1 <= OPKB_CHAIN_MB) ? MB + 1 : OPKB_CHAIN_MB;
    const int m = cs->m, mem = cs->mem, method = cs->method;
    int why = 0;
    double alpha[MN], c[MN], rho[MN];
    int slots[MN];
    double gamma_dir = 0.0, gamma_keep = cs->gamma, f = 0.0, gnorm = 0.0;
    long long iter = cs->iter, eval = cs->eval;
    const int mn = (m + 1 < mem ? m + 1 : mem);
    do {
        // -- the direction this launch assembled: descent test (:542-546, :631-632)
        const double gd = -R[0];
        const double dnorm = sqrt(R[1]);
        const double smax = cs->limits_inf ? INFINITY : R[3];
        const bool reject = !(cs->epsilon > 0 ? (gd <= -cs->epsilon * dnorm * cs->gnorm) : (gd < 0));
        if (reject) { why = 1; break; }
        // -- first step of the line search (:566-584): the trial point of this launch is x - 1*d
        const double f0 = cs->f, gd0 = gd;
        double stp = 1.0, stpmin, stpmax;
        if (method > 0) {
            stp = chain_min(stp, smax);
            stpmin = 1e-20 * stp;
            stpmax = chain_min(1e+20 * stp, smax);
        } else {
            stpmin = 1e-20 * stp;
            stpmax = 1e+20 * stp;
        }
        if (stp != 1.0) { why = 2; break; }
        // start! (src/linesearches.jl:331-358, :470-529): anything it would throw is the host's business
        if (stpmax < stpmin || stpmin < 0 || !(gd0 < 0) || stp > stpmax || stp < stpmin) { why = 3; break; }
        // -- the trial point (k_direction2, FUSE): f, |p'|^2, g'.s, g'.d, |x'|^2, |s|^2
        f = (cs->obj == OPKB_OBJ_ROSENBROCK) ? R[4] + R[5] : 0.5 * R[4];
        const double r1 = R[6], r2 = R[7], r3 = R[8], r4 = R[9], r5 = R[10];
        eval += 1;
        if (!(f < cs->bestf)) { why = 4; break; }             // (:398; a trial that does not improve: host)
        gnorm = sqrt(r1);
        if (gnorm <= cs->gtest) { why = 5; break; }           // converged (:409-417)
        if (cs->fminset && f < cs->fmin) { why = 6; break; }  // (:420-423)
        // iterate! : convergence at the first trial, or the host goes on with the search
        bool conv;
        if (cs->ls_kind == 3) {
            const double g_ls = (method > 0 ? r2 / stp : -r3);                // (:432 / :434)
            const double gtest = cs->ls_ftol * gd0, ftest = f0 + stp * gtest; // src/linesearches.jl:536-547
            conv = (f <= ftest && fabs(g_ls) <= -cs->ls_gtol * gd0);
        } else {
            conv = (f <= f0 + cs->ls_ftol * stp * gd0);                       // :235-237, :360-362
        }
        if (!conv) { why = 7; break; }
        iter += 1;                                                            // (:443)
        double xtest = chain_max(cs->xatol, 0.0);                             // (:444-468)
        if (cs->xrtol > 0) xtest = chain_max(xtest, cs->xrtol * sqrt(r4));
        if (sqrt(r5) <= xtest) { why = 8; break; }
        const double delta = chain_max(fabs(f - f0), stp * fabs(gd0));
        if (delta <= cs->fatol || delta <= cs->frtol * fabs(f0)) { why = 9; break; }
        if (iter >= cs->maxiter || eval >= cs->maxeval) { why = 10; break; }
        // -- the next Gram block comes from what this launch carried (gram_from_carry, opkb_vmlmb.cu)
        if (R[12 + ND + NCORR - 1] != 0.0) { why = 11; break; }              // a warp gave its corrections up
        const int drop = (m == mem) ? 1 : 0;
        const int W = mn + 1, nn = mn - 1;
#pragma unroll
        for (int a2 = 0; a2 < MN; ++a2) slots[a2] = (a2 < nn) ? cs->slots[a2 + drop] : ((a2 == nn) ? cs->mark_next : 0);
        // -- two-loop recursion in coefficient space (solve_two_loop, opkb_driver.cu); unrolled over
        // MB with guards, so that alpha, c, rho live in registers (one thread: every trip through
        // local memory would be on the critical path of the whole grid)
#define CH_SV(a) G[2 * (a) * W]
#define CH_YV(a) G[(2 * (a) + 1) * W]
#define CH_SY(a, b) G[2 * (a) * W + 1 + (b)]
#define CH_YY(a, b) ((b) >= (a) ? G[(2 * (a) + 1) * W + 1 + (b)] : G[(2 * (b) + 1) * W + 1 + (a)])
        if (method == 2) {
#pragma unroll
            for (int j = 0; j < MN; ++j) rho[j] = (j < mn) ? CH_SY(j, j) : 0.0;
        } else {
            const double rnew = CH_SY(nn, nn);
            cs->rho[cs->mark_next] = rnew;     // (slots[nn])
            if (rnew > 0) gamma_keep = rnew / CH_YY(nn, nn);
#pragma unroll
            for (int j = 0; j < MN; ++j) rho[j] = (j < mn) ? cs->rho[slots[j]] : 0.0;
        }
        double gamma = 0.0;
        bool modif = false;
#pragma unroll
        for (int k = MN - 1; k >= 0; --k) {
            alpha[k] = NAN;
            c[k] = 0.0;
            if (k < mn && rho[k] > 0) {
                double acc = CH_SV(k);
#pragma unroll
                for (int j = MN - 1; j > k; --j)
                    if (j < mn && rho[j] > 0) acc -= alpha[j] * CH_SY(k, j);
                alpha[k] = acc / rho[k];
                if (gamma == 0) gamma = rho[k] / CH_YY(k, k);
                modif = true;
            }
        }
        if (method < 2) gamma = gamma_keep;
        if (!modif || gamma == 0) { why = 12; break; }
#pragma unroll
        for (int k = 0; k < MN; ++k) {
            if (k < mn && rho[k] > 0) {
                double acc = CH_YV(k);
#pragma unroll
                for (int j = MN - 1; j >= 0; --j)
                    if (j < mn && rho[j] > 0) acc -= alpha[j] * CH_YY(k, j);
                acc *= gamma;
#pragma unroll
                for (int j = 0; j < MN; ++j)
                    if (j < k && rho[j] > 0) acc += c[j] * CH_SY(j, k);
                c[k] = alpha[k] - acc / rho[k];
            }
        }
#undef CH_SV
#undef CH_YV
#undef CH_SY
#undef CH_YY
        gamma_dir = gamma;
    } while (0);
    for (int i = 0; i < OPKB_CHAIN_NREC; ++i) rec[i] = 0.0;
    rec[1] = (double)why;
    if (why != 0) {
        cs->go = 0;
        return;
    }
#pragma unroll
    for (int j = 0; j < MN; ++j) {
        if (j < mn) {
            cs->alpha[j] = alpha[j];
            cs->c[j] = c[j];
            cs->slots[j] = slots[j];
            rec[3 + j] = alpha[j];
            rec[3 + OPKB_CHAIN_MB + j] = c[j];
        }
    }
    cs->gamma_dir = gamma_dir;
    cs->gamma = gamma_keep;
    cs->m = mn;
    cs->mark_next = (cs->mark_next < mem - 1) ? cs->mark_next + 1 : 0;
    cs->iter = iter;
    cs->eval = eval;
    cs->f = f;
    cs->bestf = f;
    cs->gnorm = gnorm;
    cs->go = 1;
    rec[0] = 1.0;
    rec[2] = gamma_dir;
}

// All threads of the last CTA.  cs_g: the structure in device memory; R: the combined scalars of
// this launch, in shared memory; rec_g: where the decision record goes (pinned host memory, after
// the scalars); work: shared-memory scratch of at least CHAIN_WORK doubles.  The structure is
// staged through shared memory (dependent loads from device memory by one thread cost ~20 us) and
// the next Gram block -- the previous one plus the carried corrections, and the dense sums of this
// launch (gram_from_carry, opkb_vmlmb.cu) -- is assembled by all threads, one entry each.
#define OPKB_CHAIN_NW ((int)(sizeof(ChainDev) / 8))
#define OPKB_CHAIN_WORK (OPKB_CHAIN_NW + 2 * OPKB_CHAIN_MB * (OPKB_CHAIN_MB + 1) + OPKB_CHAIN_NREC + 1)
template <int MB>
__device__ __forceinline__ void chain_decide_block(ChainDev* cs_g, const double* R, double* rec_g, double* work,
                                                   int tid, int nthreads)
{
    static_assert(sizeof(ChainDev) % 8 == 0 && offsetof(ChainDev, G) % 8 == 0, "ChainDev is copied word by word");
    constexpr int ND = 4 * MB + 4, HALF = MB * (MB + 1) / 2;
    constexpr int G0 = (int)(offsetof(ChainDev, G) / 8), GN = 2 * OPKB_CHAIN_MB * (OPKB_CHAIN_MB + 1);
    ChainDev* cs = reinterpret_cast<ChainDev*>(work);
    double* Gn = work + OPKB_CHAIN_NW;
    double* rec = Gn + GN;
    const double* src = reinterpret_cast<const double*>(cs_g);
    double* dst = reinterpret_cast<double*>(cs_g);
    const int dbg = cs_g->pad;      // development aid (OPKB_CHAIN_DBG): 2 = no decision at all, 1 = no scalar part
    if (dbg & 2) {
        if (tid == 0) cs_g->go = 0;
        for (int i = tid; i < OPKB_CHAIN_NREC; i += nthreads) rec_g[i] = 0.0;
        return;
    }
    for (int i = tid; i < OPKB_CHAIN_NW; i += nthreads) work[i] = src[i];
    __syncthreads();
    {
        const int mo = cs->m, mem = cs->mem, drop = (mo == mem) ? 1 : 0;
        const int mn = (mo + 1 < mem ? mo + 1 : mem);
        const int W = mn + 1, Wo = mo + 1, nn = mn - 1;
        const double* D = R + 12;
        const double* Cq = R + 12 + ND;
        const double* Go = cs->G;
        for (int idx = tid; idx < 2 * mn * W; idx += nthreads) {
            const int row = idx / W, col = idx - row * W;
            const int a2 = row >> 1, isy = row & 1;
            const int a = a2 + drop;
            double v = 0.0;
            if (col == 0) {
                v = (a2 < nn) ? D[4 * a + isy] : D[4 * MB + isy];
            } else {
                const int b2 = col - 1, b = b2 + drop;
                if (b2 == nn) v = (a2 < nn) ? D[4 * a + 2 + isy] : D[4 * MB + 2 + isy];
                else if (b2 >= a2 && a2 < nn)
                    v = Go[(2 * a + isy) * Wo + 1 + b] + Cq[(isy ? HALF : 0) + a * MB - a * (a - 1) / 2 + (b - a)];
            }
            Gn[idx] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (dbg & 1) {
            cs->go = 0;
            for (int i = 0; i < OPKB_CHAIN_NREC; ++i) rec[i] = 0.0;
        } else {
            chain_decide_serial<MB>(cs, R, Gn, rec);
        }
    }
    __syncthreads();
    if (cs->go) {
        for (int i = tid; i < OPKB_CHAIN_NW; i += nthreads)
            if (i < G0 || i >= G0 + GN) dst[i] = work[i];
        const int mn = cs->m;
        for (int i = tid; i < 2 * mn * (mn + 1); i += nthreads) dst[G0 + i] = Gn[i];
    } else if (tid == 0) {
        cs_g->go = 0;
    }
    for (int i = tid; i < OPKB_CHAIN_NREC; i += nthreads) rec_g[i] = rec[i];
}
#endif
