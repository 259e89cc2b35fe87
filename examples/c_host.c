/*
 * c_host.c -- a plain C host of libopkb200.so (no Python, no torch): what a binding in any
 * language does through the C ABI of include/opkb200.h.
 *
 *   gcc -O2 -I include examples/c_host.c -o c_host -L optimpacknextgen.jl_b200 -lopkb200 \
 *       -Wl,-rpath,$PWD/optimpacknextgen.jl_b200 -lm
 *   ./c_host [n]
 *
 * 1. vmlmb on the bound-constrained quadratic (built-in objective): the native driver runs the
 *    whole loop (`opkb_solver_run`), as `opk_vmlmb`-style C callers of the reference would.
 * 2. vmlmb with a USER objective by reverse communication (`opkb_solver_start / _iterate`, the
 *    scheme of the reference's own C API, src/wrappers.jl:1831-1859): f = 1/2 |x - c|^2, g = x - c,
 *    evaluated by the caller with the Tier 1 operators (`vcombine!`, `vdot`) on device vectors.
 * 3. spg with the native loop (`opkb_spg_solver_*`).
 * 4. run 1 again from this ONE process on a GROUP of ranks (`opkb_group_*`): one context per GPU
 *    (both ranks on device 0 when the box has a single GPU -- the code path is the same), the
 *    vectors sharded in contiguous blocks, the native drivers of all ranks on the library's threads.
 * Prints one line per solver and returns 0 when they agree with the known solutions.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "opkb200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != 0) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, opkb_last_error());   \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char** argv)
{
    const int64_t n = (argc > 1) ? atoll(argv[1]) : 100000;
    opkb_context* ctx = NULL;
    CHECK(opkb_context_create(&ctx, -1));
    opkb_vector *a, *c, *lo, *hi, *tmp;
    CHECK(opkb_vector_create(ctx, OPKB_DOUBLE, n, &a));
    CHECK(opkb_vector_create(ctx, OPKB_DOUBLE, n, &c));
    CHECK(opkb_vector_create(ctx, OPKB_DOUBLE, n, &lo));
    CHECK(opkb_vector_create(ctx, OPKB_DOUBLE, n, &hi));
    CHECK(opkb_vector_create(ctx, OPKB_DOUBLE, n, &tmp));
    CHECK(opkb_vector_fill_random(a, 3, 0, 1.0, 1e3, 1));     /* a_i = 1e3^u  */
    CHECK(opkb_vector_fill_random(c, 5, 0, -1.0, 2.0, 0));    /* c_i in [-1, 2) */
    CHECK(opkb_vector_fill(lo, 0.0));
    CHECK(opkb_vector_fill(hi, 1.0));
    int bad = 0;
    double f_vmlmb = 0.0;

    /* ---- 1. built-in objective, native loop ------------------------------------------- */
    {
        opkb_vmlmb* ws = NULL;
        CHECK(opkb_vmlmb_create(ctx, OPKB_DOUBLE, n, 10, OPKB_METHOD_VMLMB, lo, -INFINITY, hi, INFINITY, &ws));
        CHECK(opkb_vmlmb_set_objective(ws, OPKB_OBJ_QUADRATIC, a, c));
        CHECK(opkb_vector_fill(opkb_vmlmb_vector(ws, 'x', 0), 0.5));
        opkb_vmlmb_options opt;
        opkb_vmlmb_default_options(&opt);
        opt.maxiter = 500;
        opt.frtol = 0.0;                 /* stop on the projected gradient (grtol = 1e-6), not on f */
        opt.xrtol = 0.0;
        opkb_solver* sv = NULL;
        CHECK(opkb_solver_create(ws, &opt, &sv));
        int task = OPKB_TASK_START;
        CHECK(opkb_solver_run(sv, INT64_MAX / 2, &task));
        int64_t iter, eval, rej;
        double f, gnorm, stp;
        int stage, newx;
        CHECK(opkb_solver_status(sv, &iter, &eval, &rej, &f, &gnorm, &stp, &stage, &newx));
        /* the solution of the separable problem is clamp(c, 0, 1): x - clamp(c) must vanish */
        CHECK(opkb_project_variables(ctx, tmp, c, lo, -INFINITY, hi, INFINITY));
        CHECK(opkb_vcombine(ctx, tmp, 1.0, opkb_vmlmb_vector(ws, 'x', 0), -1.0, tmp));
        double err;
        CHECK(opkb_vnorminf(ctx, tmp, &err));
        printf("vmlmb  built-in quadratic  n=%lld: %lld iterations, %lld evaluations, f=%.6e, |x - clamp(c)|_inf=%.2e (%s)\n",
               (long long)n, (long long)iter, (long long)eval, f, err, opkb_solver_reason(sv));
        f_vmlmb = f;
        bad |= !(err < 5e-2);   /* (sanity of an example: the parity tests live in tests/) */
        CHECK(opkb_solver_destroy(sv));
        CHECK(opkb_vmlmb_destroy(ws));
    }

    /* ---- 2. user objective, reverse communication ---------------------------------------- */
    {
        opkb_vmlmb* ws = NULL;
        CHECK(opkb_vmlmb_create(ctx, OPKB_DOUBLE, n, 5, OPKB_METHOD_VMLMB, lo, -INFINITY, hi, INFINITY, &ws));
        CHECK(opkb_vmlmb_set_objective(ws, OPKB_OBJ_USER, NULL, NULL));
        CHECK(opkb_vector_fill(opkb_vmlmb_vector(ws, 'x', 0), 0.5));
        opkb_vmlmb_options opt;
        opkb_vmlmb_default_options(&opt);
        opt.maxiter = 100;
        opkb_solver* sv = NULL;
        CHECK(opkb_solver_create(ws, &opt, &sv));
        int task = OPKB_TASK_START;
        CHECK(opkb_solver_start(sv, &task));
        int64_t nfg = 0;
        while (task == OPKB_TASK_COMPUTE_FG) {
            opkb_vector* x = opkb_vmlmb_vector(ws, 'x', 0);     /* the handles are stable; their DATA POINTERS move between calls */
            opkb_vector* g = opkb_vmlmb_vector(ws, 'g', 0);
            double gg;
            CHECK(opkb_vcombine(ctx, g, 1.0, x, -1.0, c));      /* g = x - c */
            CHECK(opkb_vdot(ctx, g, g, &gg));
            CHECK(opkb_solver_iterate(sv, 0.5 * gg, &task));    /* f = |x - c|^2 / 2 */
            ++nfg;
        }
        int64_t iter, eval, rej;
        double f, gnorm, stp;
        int stage, newx;
        CHECK(opkb_solver_status(sv, &iter, &eval, &rej, &f, &gnorm, &stp, &stage, &newx));
        CHECK(opkb_project_variables(ctx, tmp, c, lo, -INFINITY, hi, INFINITY));
        CHECK(opkb_vcombine(ctx, tmp, 1.0, opkb_vmlmb_vector(ws, 'x', 0), -1.0, tmp));
        double err;
        CHECK(opkb_vnorminf(ctx, tmp, &err));
        printf("vmlmb  user objective      n=%lld: %lld iterations, %lld evaluations (%lld callbacks), f=%.6e, |x - clamp(c)|_inf=%.2e (%s)\n",
               (long long)n, (long long)iter, (long long)eval, (long long)nfg, f, err, opkb_solver_reason(sv));
        bad |= !(err < 1e-5) || (task != OPKB_TASK_FINAL_X && task != OPKB_TASK_WARNING);
        CHECK(opkb_solver_destroy(sv));
        CHECK(opkb_vmlmb_destroy(ws));
    }

    /* ---- 3. spg, native loop ------------------------------------------------------------- */
    {
        opkb_spg* ws = NULL;
        CHECK(opkb_spg_create(ctx, OPKB_DOUBLE, n, lo, -INFINITY, hi, INFINITY, &ws));
        CHECK(opkb_spg_set_objective(ws, OPKB_OBJ_QUADRATIC, a, c));
        CHECK(opkb_vector_fill(opkb_spg_vector(ws, 'x'), 0.5));
        opkb_spg_solver* sv = NULL;
        CHECK(opkb_spg_solver_create(ws, 10, 2000, 100000, 1e-7, 1e-7, 1.0, &sv));
        int searching = 1;
        CHECK(opkb_spg_solver_run(sv, INT64_MAX / 2, &searching));
        double f, fbest, pginfn, pgtwon;
        int64_t iter, fcnt, pcnt;
        int status, best_is_x;
        CHECK(opkb_spg_solver_info(sv, &f, &fbest, &pginfn, &pgtwon, &iter, &fcnt, &pcnt, &status, &best_is_x));
        CHECK(opkb_project_variables(ctx, tmp, c, lo, -INFINITY, hi, INFINITY));
        CHECK(opkb_vcombine(ctx, tmp, 1.0, opkb_spg_vector(ws, best_is_x ? 'x' : 'b'), -1.0, tmp));
        double err;
        CHECK(opkb_vnorminf(ctx, tmp, &err));
        printf("spg    built-in quadratic  n=%lld: %lld iterations, %lld evaluations, status %d, fbest=%.6e, |x - clamp(c)|_inf=%.2e\n",
               (long long)n, (long long)iter, (long long)fcnt, status, fbest, err);
        /* two different algorithms, one minimum */
        bad |= !(err < 1e-4) || status < 1 || !(fabs(fbest - f_vmlmb) <= 1e-5 * fabs(fbest));
        CHECK(opkb_spg_solver_destroy(sv));
        CHECK(opkb_spg_destroy(ws));
    }

    /* ---- 4. one process, a group of ranks ------------------------------------------------- */
    {
        enum { G = 2 };
        int have = 0, devices[G] = {0, 0};
        opkb_group* grp = NULL;
        {   /* (a second GPU if there is one) */
            opkb_group* probe = NULL;
            if (opkb_group_create(&probe, 0, NULL) == 0) {
                have = opkb_group_size(probe);
                opkb_group_destroy(probe);
            }
            if (have >= 2) devices[1] = 1;
        }
        CHECK(opkb_group_create(&grp, G, devices));
        opkb_vector *ga[G], *gc[G], *glo[G], *ghi[G];
        opkb_vmlmb* gws[G];
        opkb_solver* gsv[G];
        int tasks[G];
        const int64_t per = ((n + G - 1) / G + 3) / 4 * 4;       /* blocks of a multiple of 4 elements */
        for (int r = 0; r < G; ++r) {
            opkb_context* cr = opkb_group_context(grp, r);
            const int64_t first = (r * per < n) ? r * per : n;
            const int64_t cnt = (first + per <= n) ? per : n - first;
            CHECK(opkb_vector_create(cr, OPKB_DOUBLE, cnt, &ga[r]));
            CHECK(opkb_vector_create(cr, OPKB_DOUBLE, cnt, &gc[r]));
            CHECK(opkb_vector_create(cr, OPKB_DOUBLE, cnt, &glo[r]));
            CHECK(opkb_vector_create(cr, OPKB_DOUBLE, cnt, &ghi[r]));
            CHECK(opkb_vector_fill_random(ga[r], 3, first, 1.0, 1e3, 1));     /* the same a, c: `first` */
            CHECK(opkb_vector_fill_random(gc[r], 5, first, -1.0, 2.0, 0));
            CHECK(opkb_vector_fill(glo[r], 0.0));
            CHECK(opkb_vector_fill(ghi[r], 1.0));
            CHECK(opkb_vmlmb_create(cr, OPKB_DOUBLE, cnt, 10, OPKB_METHOD_VMLMB, glo[r], -INFINITY, ghi[r], INFINITY,
                                    &gws[r]));
            CHECK(opkb_vmlmb_set_objective(gws[r], OPKB_OBJ_QUADRATIC, ga[r], gc[r]));
            CHECK(opkb_vector_fill(opkb_vmlmb_vector(gws[r], 'x', 0), 0.5));
            opkb_vmlmb_options opt;
            opkb_vmlmb_default_options(&opt);
            opt.maxiter = 500;
            opt.frtol = 0.0;
            opt.xrtol = 0.0;
            CHECK(opkb_solver_create(gws[r], &opt, &gsv[r]));
            tasks[r] = OPKB_TASK_START;
        }
        CHECK(opkb_group_solver_run(grp, gsv, INT64_MAX / 2, tasks));
        int64_t iter, eval, rej;
        double f, gnorm, stp;
        int stage, newx;
        CHECK(opkb_solver_status(gsv[0], &iter, &eval, &rej, &f, &gnorm, &stp, &stage, &newx));
        printf("vmlmb  group of %d ranks    n=%lld: %lld iterations, %lld evaluations, f=%.6e (%s)\n", G,
               (long long)n, (long long)iter, (long long)eval, f, opkb_solver_reason(gsv[0]));
        /* same problem, same algorithm: the minimum of run 1 (the sums are combined in another order) */
        bad |= !(fabs(f - f_vmlmb) <= 1e-9 * fabs(f_vmlmb));
        for (int r = 0; r < G; ++r) {
            CHECK(opkb_solver_destroy(gsv[r]));
            CHECK(opkb_vmlmb_destroy(gws[r]));
            opkb_vector_destroy(ga[r]);
            opkb_vector_destroy(gc[r]);
            opkb_vector_destroy(glo[r]);
            opkb_vector_destroy(ghi[r]);
        }
        CHECK(opkb_group_destroy(grp));
    }

    opkb_vector_destroy(a);
    opkb_vector_destroy(c);
    opkb_vector_destroy(lo);
    opkb_vector_destroy(hi);
    opkb_vector_destroy(tmp);
    const long long launches = (long long)opkb_context_launch_count(ctx);
    CHECK(opkb_context_destroy(ctx));
    printf("%lld kernels launched; result: %s\n", launches, bad ? "MISMATCH" : "ok");
    return bad;
}
