// opkb_chain.cuh -- device-resident decisions of the steady-state VMLMB / L-BFGS iteration.
//
// In steady state an iteration of the native driver (opkb_driver.cu) is ONE launch of the fused
// kernel k_direction2<.., CARRY> (direction + first trial point of the next line search + the next
// Gram block) followed by a few hundred host flops: is the direction a descent direction
// (src/quasinewton.jl:542-546), is the step the driver tries exactly 1 (:566-584), does the line
// search accept it at once (src/linesearches.jl:235-247, :360-381, :536-547), do the convergence
// tests let the loop go on (:398-423, :437-468), and what are the two-loop coefficients of the next
// direction (apply_lbfgs!, :716-779) -- and then the next launch.  That round trip (completion flag
// over PCIe, host, launch) measures ~8 us on a B200 (C1: 53 us per iteration).
//
// CHAIN instances of the kernel take it out of the critical path: the last CTA, once it has combined
// the partials, runs chain_decide_block() below -- the very same tests and the same coefficient-space
// two-loop recursion, in the same order, on the same Float64 scalars -- and leaves `go` and the
// coefficients of the NEXT direction in a small device structure.  The host has ALREADY enqueued
// that next launch (pointers rotated as an accepted step rotates them); its prologue reads the
// structure: go = 0 -> every CTA returns at once and nothing has been touched, the host goes on as
// it always did (another step length, a rejected direction, convergence ...).  The host remains the
// authority: it receives the reduced scalars of every launch as before, runs its own stage machine
// on them, and ADOPTS the launch that is already in flight only if its own decision and its own
// coefficients are bit-identical to what the device published (anything else is an internal error,
// reported loudly).  Trajectories are therefore the same bits with and without the chain.
// Opt-in (opkb_solver_set_chain): the decision's 2m + 1 dependent Float64 divisions cost ~6.5 us on the
// GPU, the prologue and the staging ~4.5 us more -- about what the round trip costs.
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

// The scalar part, by ONE WARP (all 32 lanes call it), everything in shared memory.  cs: this launch's
// copy of the structure (updated in place for the next launch, G excepted); R: the combined scalars
// of this launch (k_direction2: 12 + dense + corrections); G: the next Gram block, already assembled
// (below).  Mirrors, statement by statement, what the host does between two fused launches:
// new_direction_gram (tail) -> after_evaluation (tail) -> opkb_solver_run -> after_evaluation (head)
// -> opkb_vmlmb_gram (gram_from_carry) -> solve_two_loop of opkb_driver.cu / opkb_vmlmb.cu.
// One thread doing all of it costs ~9.5 us (~1500 dependent instructions: measured, profiles/r02u_*).
// Here lane 0 takes the decisions (the four square roots come from lanes 1..4), and the two
// recurrences of the two-loop solve run with lane k owning pair k: every accumulator sees its
// terms in the host's order (alpha: j = m-1 down to k+1; c: all alpha_j, j = m-1 down to 0, then c_j,
// j = 0 up to k-1), each product rounded before it is added -- the same bits, ~40 dependent
// operations and 2m divisions on the critical path instead of ~1500.
template <int MB>
__device__ __forceinline__ void chain_decide_warp(ChainDev* cs, const double* R, const double* G, double* rec,
                                                  int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int ND = 4 * MB + 4, NENT = MB * (MB + 1), NCORR = 32 * ((NENT + 31) / 32);
    const int m = cs->m, mem = cs->mem, method = cs->method;
    const int mn = (m + 1 < mem ? m + 1 : mem);
    // square roots the decisions need, one lane each: 1: |d|, 2: |p'| (gnorm), 3: |x'|, 4: |s|
    double sq = 0.0;
    if (lane >= 1 && lane <= 4) sq = sqrt(R[lane == 1 ? 1 : (lane == 2 ? 6 : (lane == 3 ? 9 : 10))]);
    const double dnorm = __shfl_sync(FULL, sq, 1), gnorm = __shfl_sync(FULL, sq, 2);
    const double xnorm = __shfl_sync(FULL, sq, 3), snorm = __shfl_sync(FULL, sq, 4);
    int why = 0;
    double f = 0.0;
    long long iter = cs->iter, eval = cs->eval;
    if (lane == 0) {
        do {
            // -- the direction this launch assembled: descent test (:542-546, :631-632)
            const double gd = -R[0];
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
            const double r2 = R[7], r3 = R[8];
            eval += 1;
            if (!(f < cs->bestf)) { why = 4; break; }             // (:398; a trial that does not improve: host)
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
            if (cs->xrtol > 0) xtest = chain_max(xtest, cs->xrtol * xnorm);
            if (snorm <= xtest) { why = 8; break; }
            const double delta = chain_max(fabs(f - f0), stp * fabs(gd0));
            if (delta <= cs->fatol || delta <= cs->frtol * fabs(f0)) { why = 9; break; }
            if (iter >= cs->maxiter || eval >= cs->maxeval) { why = 10; break; }
            // -- the next Gram block comes from what this launch carried (gram_from_carry, opkb_vmlmb.cu)
            if (R[12 + ND + NCORR - 1] != 0.0) { why = 11; break; }              // a warp gave its corrections up
        } while (0);
    }
    why = __shfl_sync(FULL, why, 0);
    // -- two-loop recursion in coefficient space (solve_two_loop, opkb_driver.cu): lane k = pair k
    const int drop = (m == mem) ? 1 : 0;
    const int W = mn + 1, nn = mn - 1;
    const int k = lane;
    const bool mine = (k < mn);
    const int kk = mine ? k : 0;                       // (clamped: lanes beyond the pairs read valid memory)
    const int slot = (k < nn) ? cs->slots[kk + drop] : cs->mark_next;
    double gamma_keep = cs->gamma, gamma = 0.0, alpha = NAN, c = 0.0;
    bool change = false;
    if (why == 0) {
#define CH_SV(a) G[2 * (a) * W]
#define CH_YV(a) G[(2 * (a) + 1) * W]
#define CH_SY(a, b) G[2 * (a) * W + 1 + (b)]
#define CH_YY(a, b) ((b) >= (a) ? G[(2 * (a) + 1) * W + 1 + (b)] : G[(2 * (b) + 1) * W + 1 + (a)])
        double rho;
        if (method == 2) {
            rho = mine ? CH_SY(kk, kk) : 0.0;                                    // (:758)
        } else {
            const double rnew = CH_SY(nn, nn);                                    // (:515)
            if (rnew > 0) gamma_keep = rnew / CH_YY(nn, nn);                      // (:518)
            rho = !mine ? 0.0 : ((k == nn) ? rnew : cs->rho[slot]);
            if (k == nn) cs->rho[slot] = rnew;
        }
        const bool use = mine && (rho > 0);
        const unsigned usemask = __ballot_sync(FULL, use);
        // first loop (:725-732 / :756-766): alpha_t newest -> oldest, every older pair takes its term
        double acc = mine ? CH_SV(kk) : 0.0;
        for (int t = mn - 1; t >= 0; --t) {
            const bool ut = (usemask >> t) & 1u;                 // (uniform)
            double at = (k == t && use) ? acc / rho : 0.0;
            at = __shfl_sync(FULL, at, t);
            if (ut) {
                if (k == t) alpha = at;
                if (mine && k < t) acc -= at * CH_SY(kk, t);
            }
        }
        // gamma: from the newest pair that takes part (:762-764), or the one kept between iterations
        {
            const double q = use ? rho / CH_YY(kk, kk) : 0.0;
            const int top = usemask ? (31 - __clz((int)usemask)) : 0;
            gamma = usemask ? __shfl_sync(FULL, q, top) : 0.0;
        }
        if (method < 2) gamma = gamma_keep;
        change = (usemask != 0u) && (gamma != 0);
        if (change) {
            // second loop (:735-741 / :769-775): every alpha_j first (newest -> oldest), the scaling, then
            // the c_j of the older pairs oldest -> newest
            double b = mine ? CH_YV(kk) : 0.0;
            for (int j = mn - 1; j >= 0; --j) {
                const double aj = __shfl_sync(FULL, alpha, j);
                if (((usemask >> j) & 1u) && mine) b -= aj * CH_YY(kk, j);
            }
            b *= gamma;
            for (int t = 0; t < mn; ++t) {
                const bool ut = (usemask >> t) & 1u;             // (uniform)
                double ct = (k == t && use) ? alpha - b / rho : 0.0;
                ct = __shfl_sync(FULL, ct, t);
                if (ut) {
                    if (k == t) c = ct;
                    if (mine && k > t) b += ct * CH_SY(t, kk);
                }
            }
        } else if (lane == 0) {
            why = 12;
        }
#undef CH_SV
#undef CH_YV
#undef CH_SY
#undef CH_YY
    }
    why = __shfl_sync(FULL, why, 0);
    if (lane < OPKB_CHAIN_NREC) rec[lane] = 0.0;
    __syncwarp();
    if (why != 0) {
        if (lane == 0) {
            rec[1] = (double)why;
            cs->go = 0;
        }
        return;
    }
    if (mine) {
        cs->alpha[k] = alpha;
        cs->c[k] = c;
        rec[3 + k] = alpha;
        rec[3 + OPKB_CHAIN_MB + k] = c;
    }
    __syncwarp();                 // (every lane has read cs->slots / cs->mark_next before they change)
    if (mine) cs->slots[k] = slot;
    if (lane == 0) {
        cs->gamma_dir = gamma;
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
        rec[2] = gamma;
    }
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
__device__ __noinline__ void chain_decide_block(ChainDev* cs_g, const double* R, double* rec_g, double* work,
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
    if (tid < 32) {
        if (dbg & 1) {
            if (tid == 0) cs->go = 0;
            if (tid < OPKB_CHAIN_NREC) rec[tid] = 0.0;
        } else {
            chain_decide_warp<MB>(cs, R, Gn, rec, tid);
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
