/*
 * opkb200.h -- C ABI of libopkb200.so: the B200 (sm_100a) implementation of
 * the per-iteration vector algebra of OptimPackNextGen.jl's `vmlmb`/`vmlmb!`
 * and `spg`/`spg!`.
 *
 * This is the drop-in boundary.  A Julia host (`ccall`) or the Python twin in
 * optimpacknextgen.jl_b200/ runs the reference's driver loop, line searches and
 * convergence tests on the scalars these entry points return; every n-vector
 * lives in device memory behind an opaque handle.  Citations `file:line` are
 * relative to the reference tree (emmt/OptimPackNextGen.jl v0.4.3); the style
 * (opaque handles, `double` scalars, out-pointers, int status) follows the
 * reference's own C FFI, src/wrappers.jl:52-141, :560-628, :1238-1299.
 *
 * Conventions
 *  - every function returns 0 on success, a negative OPKB_E* code otherwise;
 *    opkb_last_error() gives the text (thread-local).  NaN/Inf in the data are
 *    not errors (IEEE propagation, src/bounds.jl:41,53).
 *  - all scalars are `double` (`Float = Cdouble`, src/OptimPackNextGen.jl:42);
 *    scalars that multiply a vector are converted to the element type first
 *    (test/vectors-tests.jl:33,46,79).
 *  - a bound is an array when its vector handle is non-NULL, else the scalar
 *    value (+-Inf = no bound), as `lower=`/`upper=` of src/quasinewton.jl:228-229.
 *  - reductions are deterministic (fixed grid, fixed tree, fixed rank order),
 *    accumulate exact products in double whatever the element type, and are
 *    GLOBAL over all ranks once opkb_comm_init() has been called: each rank
 *    holds a contiguous shard and one small packed all-gather per call
 *    combines the partials (no other data-path communication exists).
 *  - calls are synchronous w.r.t. returned scalars and asynchronous (stream
 *    ordered) otherwise; a context is not re-entrant.  All kernels of a context
 *    run on ONE private non-blocking stream, opkb_context_stream().
 *  - there is no CPU fallback: every entry point needs a CUDA device.
 *
 * Raw device pointers (callers that run their own kernels on the vectors, e.g. a
 * user objective written with CUDA.jl / CUDA C / torch):
 *  - a HANDLE is stable for the life of its object, its DATA POINTER is not: the
 *    vectors of a VMLMB or SPG workspace (`x`, `g`, `p`, `S`, `Y`, ...) trade
 *    buffers instead of copying (the saves of src/quasinewton.jl:562-563, the
 *    speculative trial point, the pair of the incremental Gram).  Call
 *    opkb_vector_data() again after EVERY solver / workspace call; never cache it.
 *  - ordering: either submit your kernels to opkb_context_stream(ctx), or
 *      * read x only after opkb_vmlmb_trial_begin / opkb_solver_start /
 *        opkb_solver_iterate have returned (with a user objective these calls
 *        return once x is complete in device memory), and
 *      * before opkb_vmlmb_trial_end / opkb_solver_iterate, make your writes to g
 *        visible to the library: synchronise your stream, or call
 *        opkb_context_wait_stream(ctx, your_stream).
 *  - every entry point makes the device of its context current (cudaSetDevice) and
 *    leaves it so.
 */
#ifndef OPKB200_H
#define OPKB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPKB_VERSION 100

/* element types: OPK_FLOAT / OPK_DOUBLE of src/wrappers.jl:226-229 */
#define OPKB_FLOAT  0
#define OPKB_DOUBLE 1

/* status codes (cf. opk_status, src/wrappers.jl:110-141) */
#define OPKB_SUCCESS          0
#define OPKB_EINVAL          -1  /* invalid argument */
#define OPKB_ECUDA           -2  /* CUDA runtime error */
#define OPKB_ENCCL           -3  /* NCCL error / NCCL not available */
#define OPKB_ENOMEM          -4
#define OPKB_EBOUNDS         -5  /* "invalid bounds", src/bounds.jl:339, :450 */
#define OPKB_EORIENT         -6  /* "invalid orientation", src/bounds.jl:223-225 */

/* orientation of project_direction / step_limits: `+` or `-`
 * (src/bounds.jl:218-229; OPK_ASCENT/OPK_DESCENT of src/wrappers.jl:1296-1299) */
#define OPKB_FORWARD   1
#define OPKB_BACKWARD -1

/* optimisation method, src/quasinewton.jl:273 */
#define OPKB_METHOD_LBFGS 0
#define OPKB_METHOD_BLMVM 1
#define OPKB_METHOD_VMLMB 2

/* built-in objectives with a fused f + gradient kernel */
#define OPKB_OBJ_USER       0  /* gradient written by the caller between trial_begin/trial_end */
#define OPKB_OBJ_ROSENBROCK 1  /* test/rosenbrock.jl:16-29 */
#define OPKB_OBJ_QUADRATIC  2  /* f = 1/2 sum a_i (x_i - c_i)^2 (BASELINE.json north_star) */

typedef struct opkb_context_ opkb_context;
typedef struct opkb_vector_  opkb_vector;
typedef struct opkb_vmlmb_   opkb_vmlmb;
typedef struct opkb_spg_     opkb_spg;
typedef struct opkb_spg_solver_ opkb_spg_solver;
typedef struct opkb_solver_  opkb_solver;
typedef struct opkb_group_   opkb_group;

const char* opkb_last_error(void);
int         opkb_version(void);

/* ---- Tier 0: lifecycle --------------------------------------------------- */
/* One context per process and GPU (device < 0: current device). */
int opkb_context_create(opkb_context** ctx, int device);
int opkb_context_destroy(opkb_context* ctx);
int opkb_context_synchronize(opkb_context* ctx);
int opkb_context_device(const opkb_context* ctx);
int opkb_context_sm_count(const opkb_context* ctx);
/* the context's stream (a cudaStream_t) and a one-way dependency on another stream of the
 * caller: everything submitted to `user_stream` so far completes before anything the library
 * launches afterwards (event record + cudaStreamWaitEvent, no host synchronisation) */
void* opkb_context_stream(const opkb_context* ctx);
int   opkb_context_wait_stream(opkb_context* ctx, void* user_stream);
/* Device memory of destroyed vectors is kept by the context and handed out again to vectors
 * of the same size (a second solve in the same process pays no cudaMalloc); trim releases it
 * to the driver (also done automatically when an allocation fails); OPKB_POOL=0 disables. */
int     opkb_context_trim(opkb_context* ctx);
int64_t opkb_context_cached_bytes(const opkb_context* ctx);
/* number of kernels launched by this context so far (bench `gpu_launches`) */
int64_t opkb_context_launch_count(const opkb_context* ctx);
/* CUDA-event timer on the context's stream */
int opkb_timer_start(opkb_context* ctx);
int opkb_timer_stop(opkb_context* ctx, double* milliseconds);
/* per-kernel profile: with the profile enabled every launch is bracketed by a
 * CUDA-event pair on the launching stream; kernel_class: 0 other, 1 trial (P1),
 * 2 gram (P3), 3 direction (P4/P4'), 4 spg, 5 reductions, 6 element-wise,
 * 7 bounds, 8 stand-alone objectives */
int opkb_profile_enable(opkb_context* ctx, int on);
int opkb_profile_reset(opkb_context* ctx);
int opkb_profile_get(opkb_context* ctx, int kernel_class, double* ms_total, int64_t* count);
/* write 256 MiB of scratch memory (> L2) to flush the L2 between timed runs */
int opkb_flush_l2(opkb_context* ctx);

/* Multi-GPU: rank 0 obtains a 128-byte id, the host broadcasts it (e.g.
 * torch.distributed / MPI.jl), every rank calls opkb_comm_init. */
int opkb_comm_unique_id(void* id128);
int opkb_comm_init(opkb_context* ctx, const void* id128, int rank, int nranks);
/* 1 if the reduced scalars of a phase travel through the shared-memory mailbox of the node
 * (the last CTA of every rank writes its partials into a segment all processes have mapped and
 * registered with CUDA, every process combines them in rank order), 0 if through the packed
 * ncclAllGather (OPKB_MAILBOX=0, or a rank that cannot map the segment). */
int opkb_comm_uses_mailbox(const opkb_context* ctx);
int opkb_comm_rank(const opkb_context* ctx);
int opkb_comm_size(const opkb_context* ctx);

/* ONE process, several GPUs.  A group owns one context per device; they are the
 * ranks 0..G-1 of an in-process communicator (same mailbox layout and rank-order combination as
 * the multi-process job: bit-identical results).  The host creates this rank's shard of every
 * vector / workspace / solver on opkb_group_context(g, r) -- none of those calls blocks -- and
 * then runs the native solvers of all ranks concurrently, one host thread per GPU inside the
 * library: a single-process `vmlmb(fg!, x0; lower, upper)` (src/quasinewton.jl:215, :226-242)
 * uses the whole box.  devices == NULL: devices 0..ndev-1 (ndev <= 0: all).  opkb_group_run is
 * the generic form: fn(rank, user) on the group's threads, for hosts that drive the ranks with
 * Tier 1 / Tier 2 calls themselves (any call that returns reduced scalars blocks until every rank
 * has made it).  The halo exchange of opkb_smooth2d_* needs the multi-process mode (NCCL). */
int           opkb_group_create(opkb_group** group, int ndev, const int* devices);
int           opkb_group_destroy(opkb_group* group);
int           opkb_group_size(const opkb_group* group);
opkb_context* opkb_group_context(opkb_group* group, int rank);
int           opkb_group_solver_run(opkb_group* group, opkb_solver* const* solvers, int64_t niter, int* tasks);
int           opkb_group_spg_solver_run(opkb_group* group, opkb_spg_solver* const* solvers, int64_t niter,
                                        int* searching);
int           opkb_group_run(opkb_group* group, int (*fn)(int rank, void* user), void* user);

/* vcreate (doc/algebra.md:27): a vector of `n` local elements (this rank's
 * shard).  The library owns the memory; host arrays are only borrowed during
 * upload/download. */
int     opkb_vector_create(opkb_context* ctx, int eltype, int64_t n, opkb_vector** v);
int     opkb_vector_destroy(opkb_vector* v);
int64_t opkb_vector_length(const opkb_vector* v);
int     opkb_vector_eltype(const opkb_vector* v);
void*   opkb_vector_data(const opkb_vector* v); /* device pointer */
int     opkb_vector_upload(opkb_vector* v, const void* host, int64_t offset, int64_t count);
int     opkb_vector_download(const opkb_vector* v, void* host, int64_t offset, int64_t count);
int     opkb_vector_fill(opkb_vector* v, double value);              /* vfill! */
/* x[i] = lo + (hi - lo) * u(seed, first + i), u = splitmix64-based uniform in [0,1)
 * (SURVEY 8(d) synthetic inputs; bit-identical to the host generator in the package);
 * if `log_scale` != 0: x[i] = lo * (hi/lo)^u. */
int     opkb_vector_fill_random(opkb_vector* v, uint64_t seed, int64_t first,
                                double lo, double hi, int log_scale);

/* ---- Tier 1: LazyAlgebra / SimpleBounds primitives, one-for-one ----------- */
/* vdot(x,y), src/quasinewton.jl:45-53 */
int opkb_vdot(opkb_context* ctx, const opkb_vector* x, const opkb_vector* y, double* result);
/* vdot(sel,x,y) with sel = {i : w[i] != 0}, src/quasinewton.jl:55-65, src/bounds.jl:701-714 */
int opkb_vdot_masked(opkb_context* ctx, const opkb_vector* w, const opkb_vector* x,
                     const opkb_vector* y, double* result);
/* vnorm2(x), src/quasinewton.jl:68-69; vnorminf(x), src/spg.jl:262 */
int opkb_vnorm2(opkb_context* ctx, const opkb_vector* x, double* result);
int opkb_vnorminf(opkb_context* ctx, const opkb_vector* x, double* result);
/* vcopy!(dst,src); vscale!(x,alpha); vupdate!(y,alpha,x); vupdate!(y,sel,alpha,x);
 * vcombine!(dst,alpha,x,beta,y); vproduct!(dst,x,y)  (doc/algebra.md:84-105) */
int opkb_vcopy(opkb_context* ctx, opkb_vector* dst, const opkb_vector* src);
int opkb_vscale(opkb_context* ctx, opkb_vector* x, double alpha);
int opkb_vupdate(opkb_context* ctx, opkb_vector* y, double alpha, const opkb_vector* x);
int opkb_vupdate_masked(opkb_context* ctx, opkb_vector* y, const opkb_vector* w,
                        double alpha, const opkb_vector* x);
int opkb_vcombine(opkb_context* ctx, opkb_vector* dst, double alpha, const opkb_vector* x,
                  double beta, const opkb_vector* y);
int opkb_vproduct(opkb_context* ctx, opkb_vector* dst, const opkb_vector* x, const opkb_vector* y);
/* The rest of the vector-space contract (doc/algebra.md:31 vswap!, :50 vnorm1, :60-67 vcombine!
 * with three terms = vcombine! then update!, same roundings; :116 vproduct!(dst, sel, x, y) with
 * sel = {w != 0}, the other elements of dst are left alone).  Not called by vmlmb / spg. */
int opkb_vswap(opkb_context* ctx, opkb_vector* x, opkb_vector* y);
int opkb_vnorm1(opkb_context* ctx, const opkb_vector* x, double* result);
int opkb_vcombine3(opkb_context* ctx, opkb_vector* dst, double alpha, const opkb_vector* x,
                   double beta, const opkb_vector* y, double gamma, const opkb_vector* z);
int opkb_vproduct_masked(opkb_context* ctx, opkb_vector* dst, const opkb_vector* w,
                         const opkb_vector* x, const opkb_vector* y);
/* project_variables!(dst,src,lo,hi), src/bounds.jl:137-213 */
int opkb_project_variables(opkb_context* ctx, opkb_vector* dst, const opkb_vector* src,
                           const opkb_vector* lo, double lo_val,
                           const opkb_vector* hi, double hi_val);
/* project_direction!(dst,x,lo,hi,orient,d), src/bounds.jl:321-409 */
int opkb_project_direction(opkb_context* ctx, opkb_vector* dst, const opkb_vector* x,
                           const opkb_vector* lo, double lo_val,
                           const opkb_vector* hi, double hi_val,
                           int orient, const opkb_vector* d);
/* step_limits(x,lo,hi,orient,d) -> (smin, smax), src/bounds.jl:431-624 */
int opkb_step_limits(opkb_context* ctx, const opkb_vector* x,
                     const opkb_vector* lo, double lo_val,
                     const opkb_vector* hi, double hi_val,
                     int orient, const opkb_vector* d, double* smin, double* smax);
/* get_free_variables!(sel,gp), src/bounds.jl:701-714: no index list is built;
 * the free set is the mask gp != 0.  count = |sel| (global); if `mask_bytes`
 * is non-NULL it receives this rank's mask as one byte per element (host). */
int opkb_free_variables(opkb_context* ctx, const opkb_vector* gp, int64_t* count,
                        uint8_t* mask_bytes);
/* built-in objectives as stand-alone `fg!(x, g) -> f` (test/rosenbrock.jl:16-29) */
int opkb_rosenbrock_fg(opkb_context* ctx, const opkb_vector* x, opkb_vector* g, double* f);
int opkb_quadratic_fg(opkb_context* ctx, const opkb_vector* a, const opkb_vector* c,
                      const opkb_vector* x, opkb_vector* g, double* f);

/* A non-separable objective with a neighbour exchange (no reference
 * code: north_star "bound-constrained convex quadratic", here the smoothness-
 * regularised weighted least squares typical of the reference's users):
 *   f(x) = 1/2 sum w_i (x_i - y_i)^2 + mu/2 sum over horizontal and vertical
 *   neighbour pairs (x_i - x_j)^2 on a rows x cols grid (row-major, free boundary).
 * A stand-alone `fg!(x, g) -> f` for the user-objective path of vmlmb / spg.  With
 * several ranks the grid is sharded in blocks of whole rows (x, w, y, g hold this
 * rank's rows); the last row of the previous rank and the first row of the next
 * one are fetched into `up` / `dn` (vectors of `cols` elements, may be NULL on one
 * rank) by ONE ncclSend/ncclRecv group over NVLink on a second stream while the
 * rows that need no halo are processed; *f is the global value (every rank). */
int opkb_smooth2d_fg(opkb_context* ctx, const opkb_vector* w, const opkb_vector* y, double mu,
                     int64_t cols, const opkb_vector* x, opkb_vector* up, opkb_vector* dn,
                     opkb_vector* g, double* f);

/* The same stencil as a linear operator: q = A.p, A = diag(w) + mu*L (the Hessian of
 * the objective above; symmetric positive definite for w > 0), the `A` of `conjgrad`
 * for this workload.  Same row sharding and halo exchange.  out[0] = p.q, out[1] = p.p
 * (global values) are reduced in the same pass. */
int opkb_smooth2d_apply(opkb_context* ctx, const opkb_vector* w, double mu, int64_t cols,
                        const opkb_vector* p, opkb_vector* up, opkb_vector* dn, opkb_vector* q,
                        double out[2]);

/* One pass of `conjgrad!` with that operator: p <- r (restart != 0) | beta*p + r
 * (`vcombine!(p, beta, p, 1, r)`), q = A.p, out[0] = p.q, out[1] = p.p -- read p, r, w; write
 * p, q: 5 passes instead of 3 + 3 (+ 2).  `pscratch` receives the new direction and the two
 * buffers are swapped before returning (p is updated in place for the caller; both must be
 * vectors created by opkb_vector_create).  With several ranks only the boundary rows of r
 * travel (received in up_r / dn_r); up / dn keep the neighbours' boundary rows of the
 * direction between calls (all four: `cols` elements, may be NULL on one rank). */
int opkb_cg_smooth2d_step(opkb_context* ctx, const opkb_vector* w, double mu, int64_t cols,
                          opkb_vector* p, opkb_vector* pscratch, const opkb_vector* r, opkb_vector* q,
                          opkb_vector* up, opkb_vector* dn, opkb_vector* up_r, opkb_vector* dn_r,
                          double beta, int restart, double out[2]);

/* ---- Tier 2: fused passes of `conjgrad!` ------------------------
 * `conjgrad` / `conjgrad!` are re-exported by the reference from LazyAlgebra
 * (src/OptimPackNextGen.jl:18-19, :36; call site test/conjgrad-tests.jl:30).  One
 * iteration = direction (`vcombine!(p, beta, p, 1, r)`: opkb_vcombine), operator,
 * curvature gamma = vdot(p, q), update.  All sums: exact products accumulated in double,
 * fixed order, combined over the ranks. */
/* x += alpha*p; r -= alpha*q (two `vupdate!`) and, from the same pass, out[0] = vdot(r, r),
 * out[1] = vdot(x, x): 6 passes instead of 9 (11 with the x test). */
int opkb_cg_update(opkb_context* ctx, opkb_vector* x, opkb_vector* r, const opkb_vector* p,
                   const opkb_vector* q, double alpha, double out[2]);
/* out[0] = vdot(p, q), out[1] = vdot(p, p) in one pass (user operators). */
int opkb_cg_curvature(opkb_context* ctx, const opkb_vector* p, const opkb_vector* q, double out[2]);
/* A = diag(a) (`vproduct!`): p = r (restart != 0) or beta*p + r; q = a.*p; out[0] = p.q,
 * out[1] = p.p: direction + operator + curvature in 5 passes. */
int opkb_cg_diag_step(opkb_context* ctx, const opkb_vector* a, opkb_vector* p, const opkb_vector* r,
                      opkb_vector* q, double beta, int restart, double out[2]);

/* ---- Tier 2: fused phases of `_vmlmb!` (src/quasinewton.jl:381-601) ------- */
/* The workspace owns x, g, p, d and the mem pairs S[k], Y[k] (src/quasinewton.jl:358-369).
 * lo/hi handles must outlive the workspace. */
/* mem: any value in 1..256 (the reference accepts any, src/quasinewton.jl:227, :328).  The
 * Gram-form entry points (opkb_vmlmb_gram, _direction, _direction_trial) work on m <= 20 pairs;
 * beyond that the hosts use the step-by-step form (opkb_vmlmb_twoloop_step), which has no limit. */
int opkb_vmlmb_create(opkb_context* ctx, int eltype, int64_t n, int mem, int method,
                      const opkb_vector* lo, double lo_val,
                      const opkb_vector* hi, double hi_val, opkb_vmlmb** ws);
int opkb_vmlmb_destroy(opkb_vmlmb* ws);
int opkb_vmlmb_set_objective(opkb_vmlmb* ws, int kind, const opkb_vector* a, const opkb_vector* c);
/* borrowed handles to the workspace vectors: which = 'x','g','p','d','S','Y' */
opkb_vector* opkb_vmlmb_vector(opkb_vmlmb* ws, int which, int slot);

/* P1 "trial" = :599, :386, :391 (built-in objective), :396, :399, :427, :432/:434,
 * :446, :448.  If `initial` the point is x itself (first evaluation), else
 * x = S[mark] - stp*d.  out[0..7] = { f, |p|^2 (|g|^2 for L-BFGS), g.s, g.d,
 * |x|^2, |s|^2 (s = x - S[mark]), number of free variables, 0 }. */
int opkb_vmlmb_trial(opkb_vmlmb* ws, int mark, double stp, int initial, double out[8]);
/* same, for a user objective: begin computes x; the caller writes g (device)
 * and passes f to end, which computes p and the reductions. */
int opkb_vmlmb_trial_begin(opkb_vmlmb* ws, int mark, double stp, int initial);
int opkb_vmlmb_trial_end(opkb_vmlmb* ws, int mark, double f, int initial, double out[8]);

/* P3 "pair update + sweep" = :512-513 then every inner product apply_lbfgs!
 * needs (:728, :737, :758-764, :771) in ONE pass: slot `mark` is turned in
 * place into (S[mark]-x, Y[mark]-g) [Y[mark]-p for BLMVM] (not with the iterate history, see
 * opkb_vmlmb_history below: the pair is then formed on the fly), then with the m
 * pairs ordered oldest..newest, slots[j] (j = 0 oldest), the Gram block
 *   G[r][c], r in {s_0,y_0,s_1,y_1,...} (2m rows), c in {v0,y_0,...,y_{m-1}} (m+1 cols)
 * restricted to the free set {i : p[i] != 0} for VMLMB (unrestricted otherwise,
 * v0 = p for VMLMB, g otherwise) is returned row-major in G[2m*(m+1)].  Only
 * the entries with age(col) >= age(row) and the v0 column are defined. */
int opkb_vmlmb_gram(opkb_vmlmb* ws, int mark, int m, const int* slots, double* G);
/* Representation of the pairs.  A workspace whose pairs are formed by opkb_vmlmb_gram keeps the
 * ITERATE HISTORY instead of the differences: slot k holds (x_k, g_k) as saved when line search k
 * started (:562-563), it is NOT turned in place into the difference of :512-513, and pair k is
 * S[k] - S[k'] , Y[k] - Y[k'] with k' the slot saved next (the current x, g for the newest pair),
 * formed on the fly by the kernels (the same bits: one rounding of the same operands).  Nothing is
 * written back, the fused direction kernel moves 2m + 9 passes instead of 2m + 11, the sweep 2m + 3
 * instead of 2m + 5.  `slots` must then list consecutive line searches (as `_vmlmb!` produces them).
 * opkb_vmlmb_vector(ws, 'S' | 'Y', k) of such a workspace is the iterate, not the pair.  A workspace
 * whose pairs are formed by opkb_vmlmb_pair_update (the step-by-step recursion), a BLMVM workspace,
 * a Float32 workspace or one with more than 8 pairs (unless OPKB_HISTORY=2: forming the differences
 * costs 2m operations per element, the two passes saved do not depend on m -- measured: +5 % at m = 5,
 * -1 % at m = 10, Float32 always loses) or any workspace created with OPKB_HISTORY=0 keeps the reference's stored
 * differences (and opkb_vmlmb_gram then works on those).  1: history, 0: differences, -1: not yet
 * decided (no pair formed). */
int opkb_vmlmb_history(const opkb_vmlmb* ws);

/* P4 "direction" = :525/:528 + the axpys of :729/:738/:761/:768/:772, :535,
 * :537, :629, :562-563, :577.  d = v0; for j newest..oldest d -= alpha[j]*y_j;
 * d *= gamma; for j oldest..newest d += c[j]*s_j (free set only for VMLMB; the
 * rounding sequence of the reference's axpys is kept); d = project_direction(d);
 * S[mark_next] = x; Y[mark_next] = g (p for BLMVM).  A pair with alpha[j] = c[j] = 0
 * exactly is skipped (rho <= 0).  out[0..3] = { g.d, |d|^2, smin, smax }. */
int opkb_vmlmb_direction(opkb_vmlmb* ws, int m, const int* slots, const double* alpha,
                         const double* c, double gamma, int mark_next, double out[4]);
/* P4 + P1 in ONE pass: as opkb_vmlmb_direction, and when the workspace has a
 * built-in objective (and the method is not BLMVM) the first trial point of the
 * next line search, x' = project(x - 1*d) -- the step the driver takes after a
 * successful update, :567 -- is evaluated speculatively in the same kernel:
 * trial[0..6] as out[] of opkb_vmlmb_trial, trial[7] = 1 (0: nothing was fused,
 * call opkb_vmlmb_trial as usual).  The caller uses trial[] iff the direction
 * is accepted and min(1, smax) == 1; if the direction is rejected it must call
 * opkb_vmlmb_discard_trial before the steepest-descent phase; a different first
 * step needs no repair (opkb_vmlmb_trial simply overwrites x, g, p). */
int opkb_vmlmb_direction_trial(opkb_vmlmb* ws, int m, const int* slots, const double* alpha,
                               const double* c, double gamma, int mark_next, double out[4],
                               double trial[8]);
int opkb_vmlmb_discard_trial(opkb_vmlmb* ws, int mark_next);
/* Incremental Gram.  When opkb_vmlmb_direction_trial follows opkb_vmlmb_gram for
 * the same pairs (m <= 12), the fused kernel also writes the pair the trial
 * point would form, (x - x', g - g') (:512-513), to two spare buffers and reduces
 * every inner product of the NEXT Gram block that involves it or the new v0;
 * the products between old pairs are the previous block plus signed corrections
 * over the variables that enter or leave the free set.  If the caller then takes
 * the speculated point and calls opkb_vmlmb_gram for the foreseen pairs, the
 * block is assembled on the host and the spare buffers are swapped in: no kernel
 * launch, S and Y cross HBM once per iteration instead of twice.  Any other use
 * of the workspace falls back to the sweep.  Same values up to the summation
 * order.  OPKB_CARRY=0 in the environment (read when the workspace is created) disables
 * it; Float32 workspaces use it for mem <= 8 (beyond: only with OPKB_CARRY=2, those
 * instances are issue-bound).  A launch whose trial point changes the status of more than
 * a quarter of the variables gives the incremental block up (the sweep runs if that point
 * is accepted).  carry_stats: how many opkb_vmlmb_gram calls were served that way / fell
 * back although armed. */
int opkb_vmlmb_carry_stats(const opkb_vmlmb* ws, int64_t* hits, int64_t* misses);
/* Device-resident decisions of the native driver (opkb_solver_run with a built-in objective,
 * Float64, mem <= 8, one rank; csrc/opkb_chain.cuh).  In steady state the last CTA of the fused
 * launch takes the decisions the host takes between two launches -- descent test
 * (src/quasinewton.jl:542-546), first step = 1 (:566-584), acceptance by the line search at the first
 * trial (src/linesearches.jl:235-247, :360-381, :536-547), convergence tests (:398-423, :437-468) --
 * and solves the two-loop coefficients of the next direction (:716-779), which the launch of the next
 * iteration, enqueued ahead by the host, reads from device memory: the host round trip leaves the
 * critical path.  The host still runs its own stage machine on the same scalars and adopts a launch
 * only if its decision and coefficients are bit-identical: same trajectories with and without.
 * Opt-in (opkb_solver_set_chain, or OPKB_CHAIN=1 in the environment when the solver is created): on a B200 the round trip it
 * removes (~8 us) is shorter than what the one-thread decision adds to the kernel (~12 us), see
 * DESIGN.md.  chain_stats: launches
 * that started a chain / launches adopted while already in flight / launches enqueued ahead that the
 * device cancelled (another step length, a rejected direction, convergence: the host took over). */
int opkb_vmlmb_chain_stats(const opkb_vmlmb* ws, int64_t* heads, int64_t* adopted, int64_t* nogo);
/* development aid (OPKB_CHAIN_CHECK=1 when the workspace is created): how many times the host asked
 * for the next direction right after a chained launch, and how many times the decision record of
 * that launch's last CTA did NOT hold the same decision and coefficients bit for bit (must be 0) */
int opkb_vmlmb_chain_check_stats(const opkb_vmlmb* ws, int64_t* checked, int64_t* failed);
/* the Gram block the last opkb_vmlmb_gram call returned (*m = 0: none), for checks */
int opkb_vmlmb_last_gram(const opkb_vmlmb* ws, int* m, int* slots, double* G);
/* Sequential two-loop recursion, fused step by step (parity-grade: the
 * intermediate vector is rounded in the element type after every vupdate!, as
 * the reference does).  v is the workspace vector d.
 * pair_update = :512-518: S[mark] -= x; Y[mark] -= g (p for BLMVM);
 *   out = { <y, s>, <y, y> } (unrestricted: rho and gamma of L-BFGS / BLMVM).
 * twoloop_step = one iteration of the loops :725-741 / :756-775:
 *   [v += a*u on the free set]  u = S[u_slot] if u_which == 'S', Y[u_slot] if 'Y', nothing if 0
 *   [v *= gamma]                if has_scale (:707, :768)
 *   out[0] = <z, v>, z = S|Y[z_slot];  if want_rho: out[1] = <Y[z_slot], S[z_slot]>,
 *   out[2] = <Y[z_slot], Y[z_slot]>, all on the free set {p != 0} for VMLMB.
 *   `first` != 0: v is read from p (g for L-BFGS / BLMVM): the copy of :525/:528
 *   is fused into the first step that writes v.
 * finish_direction = P4 tail for a direction held in d, after the pending last
 *   axpy d += a*u (u_which = 0: none): :535, :537, :629, :562-563, :577; same
 *   outputs as opkb_vmlmb_direction. */
int opkb_vmlmb_pair_update(opkb_vmlmb* ws, int mark, double out[2]);
int opkb_vmlmb_twoloop_step(opkb_vmlmb* ws, int first, int u_which, int u_slot, double a,
                            int has_scale, double gamma, int z_which, int z_slot,
                            int want_rho, double out[3]);
int opkb_vmlmb_finish_direction(opkb_vmlmb* ws, int mark_next, int u_which, int u_slot,
                                double a, double out[4]);
/* P4' "steepest descent" = :554, :562-563, :577: d = p (g for L-BFGS). */
int opkb_vmlmb_steepest(opkb_vmlmb* ws, int mark_next, double out[4]);
/* P0 "revert to best" = :490-492: x = project(S[mark] - stp*d). */
int opkb_vmlmb_revert(opkb_vmlmb* ws, int mark, double stp);

/* ---- Tier 3: native reverse-communication driver of VMLMB ------
 * The stage machine of `_vmlmb!` (src/quasinewton.jl:296-612), the line searches
 * (src/linesearches.jl) and the two-loop solver in host C++ inside the library,
 * over the fused phases above: one call per objective evaluation, as the
 * reference's own C API does (`opk_iterate_vmlmb`, src/wrappers.jl:1831-1859;
 * tasks as OPK_TASK_*, :758-765).  The Julia/Python hosts may keep their own
 * driver loop (north_star) or use this one. */
#define OPKB_TASK_ERROR      -1
#define OPKB_TASK_START       0
#define OPKB_TASK_COMPUTE_FG  1   /* evaluate f(x), write g (workspace vectors 'x', 'g') */
#define OPKB_TASK_NEW_X       2   /* (reported through `new_x` of opkb_solver_status) */
#define OPKB_TASK_FINAL_X     3   /* convergence: x holds the solution */
#define OPKB_TASK_WARNING     4   /* stopped with a warning: x holds the best point */

#define OPKB_LS_DEFAULT       0   /* More & Thuente without bounds, More & Toraldo with (:276-284) */
#define OPKB_LS_ARMIJO        1
#define OPKB_LS_MORE_TORALDO  2
#define OPKB_LS_MORE_THUENTE  3

typedef struct opkb_vmlmb_options {   /* keywords of vmlmb!, src/quasinewton.jl:227-242 */
    int     blmvm;                    /* informative: the method is fixed by opkb_vmlmb_create */
    double  fmin;
    int64_t maxiter, maxeval;
    double  xatol, xrtol, fatol, frtol, gatol, grtol, epsilon;
    int     lnsrch;                   /* OPKB_LS_* */
    double  ls_ftol, ls_gtol, ls_xtol, ls_gamma1, ls_gamma2;
    int     ls_strict;
    int     twoloop;                  /* -1 default (Gram for Float64, sequential for Float32), 0 Gram, 1 sequential */
    int     speculate;                /* fuse the first trial of a line search into P4 (default 1) */
} opkb_vmlmb_options;

void opkb_vmlmb_default_options(opkb_vmlmb_options* opt);
/* The workspace must hold the starting point in 'x' and have its objective set. */
int  opkb_solver_create(opkb_vmlmb* ws, const opkb_vmlmb_options* opt, opkb_solver** solver);
int  opkb_solver_destroy(opkb_solver* solver);
/* reverse communication: start -> COMPUTE_FG; then iterate(f) after every evaluation */
int  opkb_solver_start(opkb_solver* solver, int* task);
int  opkb_solver_iterate(opkb_solver* solver, double f, int* task);
/* built-in objective, or a user objective with a device callback (below): run until `iter`
 * has advanced by niter or termination */
int  opkb_solver_run(opkb_solver* solver, int64_t niter, int* task);
/* Device-callback objective (cf. the callback style of the reference's C
 * solvers, src/cobyla.jl:184-193): `fn(user, n, x_dev, g_dev, stream, &f)` must compute
 * f(x) and write the gradient to g_dev (n local elements of the workspace's element type),
 * with kernels submitted to `stream` (the context's cudaStream_t: ordered with the library's
 * own work) or otherwise completed before it returns; it returns 0, or non-zero to abort
 * (opkb_solver_run then fails with OPKB_EINVAL).  With several ranks *f is this rank's value
 * of a separable objective when `f_is_partial` != 0 (the library sums it over the ranks in rank
 * order), the global value otherwise.  Then opkb_solver_run drives the whole reverse-
 * communication loop inside the library: no host round trip between trial_begin and trial_end. */
typedef int (*opkb_fg_callback)(void* user, int64_t n, const void* x_dev, void* g_dev, void* stream,
                                double* f);
int  opkb_solver_set_fg_callback(opkb_solver* solver, opkb_fg_callback fn, void* user, int f_is_partial);
/* Device-resident decisions of opkb_solver_run (see opkb_vmlmb_chain_stats above): on = 1 asks for them, 0
 * switches them off (the default; OPKB_CHAIN=1 in the environment asks for them at creation).  Not an
 * error when the solver cannot use them (user objective, Float32, mem > 8, BLMVM, the step-by-step
 * two-loop recursion, several ranks): opkb_solver_chain_enabled tells what is in effect.  Results do not
 * depend on the switch. */
int  opkb_solver_set_chain(opkb_solver* solver, int on);
int  opkb_solver_chain_enabled(const opkb_solver* solver);
int  opkb_solver_status(const opkb_solver* solver, int64_t* iter, int64_t* eval, int64_t* rejects,
                        double* f, double* gnorm, double* stp, int* stage, int* new_x);
const char* opkb_solver_reason(const opkb_solver* solver);
/* slot of S, Y that holds (x0, g0) of the line search in progress: S[mark] is the current iterate */
int  opkb_solver_mark(const opkb_solver* solver);

/* ---- Tier 2: fused passes of `_spg!` with the box projection (src/spg.jl:245-380) */
/* The workspace owns four x buffers and two g buffers: x0/g0 (:322-323) and
 * xbest (:344) are kept by rotating buffers, never by copying; pg, d, s, y
 * (:224-226) are recomputed in registers and never stored. */
int opkb_spg_create(opkb_context* ctx, int eltype, int64_t n,
                    const opkb_vector* lo, double lo_val,
                    const opkb_vector* hi, double hi_val, opkb_spg** ws);
int opkb_spg_destroy(opkb_spg* ws);
int opkb_spg_set_objective(opkb_spg* ws, int kind, const opkb_vector* a, const opkb_vector* c);
/* borrowed handles: which = 'x', '0' (x0), 'b' (xbest), 'g', 'G' (g0); they
 * change at every step/backtrack/mark_best call */
opkb_vector* opkb_spg_vector(opkb_spg* ws, int which);
/* Every pass returns out[0..5] = { f, delta = g0.d, |pg|^2, |pg|_inf, s.y, s.s }
 * with pg = (x - prj(x - eta*g))/eta (:259-262), d = prj(x0 - lambda*g0) - x0
 * (:327-330), s = x - x0, y = g - g0 (:309-314), all at the NEW point x.
 * start     :230-241  x = prj(x); f = fg(x, g); xbest = x   (delta, s.y, s.s = 0)
 * step      :322-330 + :336  x0 = x, g0 = g; x = prj(x0 - lambda*g0); f = fg(x, g)
 * backtrack :368 + :336      x = x0 + stp*d; f = fg(x, g) */
int opkb_spg_start(opkb_spg* ws, double eta, double out[8]);
int opkb_spg_step(opkb_spg* ws, double lambda, double eta, double out[8]);
int opkb_spg_backtrack(opkb_spg* ws, double lambda, double stp, double eta, double out[8]);
/* :342-345: xbest = x (no copy; the buffer is simply not reused) */
int opkb_spg_mark_best(opkb_spg* ws);

/* ---- Tier 3: the loop of `_spg!` (src/spg.jl:245-380) inside the library, for a workspace with a
 * built-in objective and the box projection; keywords of `spg!` (src/spg.jl:168-178).  `run` performs
 * `niter` iterations (or stops at termination): *searching = 1 while the status is SEARCHING.
 * `info` = the fields of SPG.Info (:38-47) + whether the solution is 'x' or 'b' (xbest, :401). */
int opkb_spg_solver_create(opkb_spg* ws, int m, int64_t maxit, int64_t maxfc, double eps1, double eps2,
                           double eta, opkb_spg_solver** solver);
int opkb_spg_solver_destroy(opkb_spg_solver* solver);
int opkb_spg_solver_run(opkb_spg_solver* solver, int64_t niter, int* searching);
int opkb_spg_solver_info(const opkb_spg_solver* solver, double* f, double* fbest, double* pginfn,
                         double* pgtwon, int64_t* iter, int64_t* fcnt, int64_t* pcnt, int* status,
                         int* best_is_x);

#ifdef __cplusplus
}
#endif
#endif /* OPKB200_H */
