// opkb_stencil.cu -- a NON-separable objective: smoothness-
// regularised weighted least squares on a rows x cols grid,
//     f(x) = 1/2 sum_i w_i (x_i - y_i)^2 + mu/2 sum_{edges (i,j)} (x_i - x_j)^2,
//     g_i  = w_i (x_i - y_i) + mu sum_{j in N(i)} (x_i - x_j),
// (edges = horizontal and vertical neighbours, free boundary): the first workload of
// the library whose objective needs neighbour data across shards.  The grid is
// sharded in contiguous blocks of rows; a rank needs the last row of the previous
// rank and the first row of the next one.  Those two rows travel point to point
// (ncclSend / ncclRecv, one group, over NVLink) on a second stream WHILE the rows
// that need no halo are being processed; only the first and the last local rows
// wait for them.  The vertical edge between two ranks is counted by the lower
// one (each edge once), all reductions are deterministic, and f is combined over
// the ranks by the usual packed all-gather.
#include "opkb_internal.cuh"

template <typename T> struct SmoothArgs {
    const T* x;
    const T* w;
    const T* y;
    const T* up;     // row above the block (previous rank) or NULL
    const T* dn;     // row below the block (next rank) or NULL
    T* g;
    T mu;
    int64_t rows, cols;       // of this rank's block
    int64_t r0, r1;           // rows [r0, r1) are processed by this launch
};

// One CTA = 256 x VEC consecutive columns of one row (VEC = the 16-byte pack when the
// row length allows aligned packs, else 1); the neighbours in the row come from the
// same pack or cache line, those of the rows above / below from L2 (each x element is
// fetched from HBM once).  Arithmetic in T without contraction, in the fixed order up,
// down, left, right, so that g is bit-identical to the obvious array expression.
// HESS: the same stencil as a linear operator, q = A.p with A = diag(w) + mu*L (the
// Hessian of f, y plays no role), and the two sums a conjugate-gradient step needs from
// that pass: acc[0] = p.q, acc[1] = p.p (exact products accumulated in double).
template <typename T, bool HESS> __device__ __forceinline__ void smooth_elem(
    T xi, T wi, T yi, bool hu, T xu, bool hd, T xd, bool hl, T xl, bool hr, T xr, T mu, T& gi, double (&acc)[2])
{
    const T r = HESS ? xi : xi - yi;
    const T t = wi * r;
    if (!HESS) acc[0] += (double)(T)(t * r);
    T s = T(0);
    if (hu) {
        const T d = xi - xu;
        s = s + d;
        if (!HESS) acc[1] += (double)(T)(d * d);      // vertical edge: owned by the lower row
    }
    if (hd) s = s + (xi - xd);
    if (hl) {
        const T d = xi - xl;
        s = s + d;
        if (!HESS) acc[1] += (double)(T)(d * d);      // horizontal edge: owned by the right element
    }
    if (hr) s = s + (xi - xr);
    gi = t + mu * s;
    if (HESS) {
        acc[0] = dacc(xi, gi, acc[0]);
        acc[1] = dacc(xi, xi, acc[1]);
    }
}

template <typename T, int VEC, bool HESS>
__global__ void __launch_bounds__(OPKB_BLOCK) k_smooth2d(SmoothArgs<T> A, ReduceScratch rs)
{
    const int64_t ncb = (A.cols + OPKB_BLOCK * VEC - 1) / (OPKB_BLOCK * VEC);
    const int64_t nitems = (A.r1 - A.r0) * ncb;
    double acc[2] = {0.0, 0.0};   // 0: sum w r^2; 1: sum of squared differences over the owned edges
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int64_t row = A.r0 + item / ncb;
        const int64_t col = ((item % ncb) * OPKB_BLOCK + threadIdx.x) * VEC;
        if (col >= A.cols) continue;
        const int64_t i = row * A.cols + col;
        const bool hu = (row > 0) || (A.up != NULL), hd = (row + 1 < A.rows) || (A.dn != NULL);
        const T* xu = (row > 0) ? A.x + i - A.cols : A.up + col;
        const T* xd = (row + 1 < A.rows) ? A.x + i + A.cols : A.dn + col;
        if (VEC == 1) {
            T gi;
            smooth_elem<T, HESS>(A.x[i], A.w[i], HESS ? T(0) : A.y[i], hu, hu ? xu[0] : T(0), hd, hd ? xd[0] : T(0), col > 0,
                           col > 0 ? A.x[i - 1] : T(0), col + 1 < A.cols, col + 1 < A.cols ? A.x[i + 1] : T(0),
                           A.mu, gi, acc);
            A.g[i] = gi;
        } else {
            constexpr int V = VecOf<T>::N;
            const Pack<T> px = ld_pack<T>(A.x + i, 0), pw = ld_pack<T>(A.w + i, 0);
            const Pack<T> py = HESS ? px : ld_pack<T>(A.y + i, 0);
            const Pack<T> pu = hu ? ld_pack<T>(xu, 0) : px, pd = hd ? ld_pack<T>(xd, 0) : px;
            const T xl = (col > 0) ? A.x[i - 1] : T(0);
            const T xr = (col + V < A.cols) ? A.x[i + V] : T(0);
            Pack<T> pg;
#pragma unroll
            for (int q = 0; q < V; ++q)
                smooth_elem<T, HESS>(px.v[q], pw.v[q], py.v[q], hu, pu.v[q], hd, pd.v[q], col + q > 0,
                               q > 0 ? px.v[q > 0 ? q - 1 : 0] : xl, col + q + 1 < A.cols,
                               q + 1 < V ? px.v[q + 1 < V ? q + 1 : 0] : xr, A.mu, pg.v[q], acc);
            st_pack<T>(A.g + i, 0, pg);
        }
    }
    block_reduce_publish<2, 0u, 0u>(acc, rs);
}

template <typename T>
static int smooth_launch(opkb_context* ctx, SmoothArgs<T> A, int slot)
{
    if (A.r1 <= A.r0) return 0;
    // 16-byte packs need rows that start on a 16-byte boundary
    constexpr int V = VecOf<T>::N;
    const bool vec = (A.cols % V == 0);
    const int64_t per = (int64_t)OPKB_BLOCK * (vec ? V : 1);
    const int64_t ncb = (A.cols + per - 1) / per;
    // (6 CTAs of 256 threads per SM: room is left for the NCCL send/recv kernel to run alongside)
    int grid = opkb_grid_for(ctx, (A.r1 - A.r0) * ncb * OPKB_BLOCK, 6);
    ReduceScratch rs;
    rs.part = ctx->d_part;
    rs.ticket = ctx->d_ticket;
    rs.res = ctx->d_res + 2 * slot;
    opkb_prof_begin(ctx, OPKB_K_OBJECTIVE);
    const bool hess = (A.y == NULL);
    if (vec && hess) k_smooth2d<T, V, true><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);
    else if (vec) k_smooth2d<T, V, false><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);
    else if (hess) k_smooth2d<T, 1, true><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);
    else k_smooth2d<T, 1, false><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);
    opkb_prof_end(ctx);
    ctx->launches += 1;
    OPKB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int smooth_run(opkb_context* ctx, const opkb_vector* w, const opkb_vector* y, double mu, int64_t cols,
                      const opkb_vector* x, opkb_vector* up, opkb_vector* dn, opkb_vector* g, double* f)
{
    // y == NULL: operator mode, f[0] = x.g, f[1] = x.x (opkb_smooth2d_apply)
    const int64_t rows = x->n / cols;
    const bool has_up = ctx->nranks > 1 && ctx->rank > 0;
    const bool has_dn = ctx->nranks > 1 && ctx->rank < ctx->nranks - 1;
    if ((has_up && (!up || up->n != cols)) || (has_dn && (!dn || dn->n != cols)))
        return opkb_set_error(OPKB_EINVAL, "halo rows of %lld elements are needed", (long long)cols);
    OPKB_CUDA(cudaMemsetAsync(ctx->d_res, 0, 6 * sizeof(double), ctx->stream));
    if (ctx->nranks > 1) {
        if (!ctx->halo_stream) {
            OPKB_CUDA(cudaStreamCreateWithFlags(&ctx->halo_stream, cudaStreamNonBlocking));
            OPKB_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
            OPKB_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
        }
        // x is ready on the main stream -> exchange the boundary rows on the halo stream
        OPKB_CUDA(cudaEventRecord(ctx->ev_ready, ctx->stream));
        OPKB_CUDA(cudaStreamWaitEvent(ctx->halo_stream, ctx->ev_ready, 0));
        const T* xd = (const T*)x->data;
        int rc = opkb_nccl_halo(ctx, x->eltype, (size_t)cols, xd, has_up ? up->data : NULL,
                                xd + (rows - 1) * cols, has_dn ? dn->data : NULL, ctx->halo_stream);
        if (rc) return rc;
        OPKB_CUDA(cudaEventRecord(ctx->ev_halo, ctx->halo_stream));
    }
    SmoothArgs<T> A;
    A.x = (const T*)x->data;
    A.w = (const T*)w->data;
    A.y = y ? (const T*)y->data : NULL;
    A.up = has_up ? (const T*)up->data : NULL;
    A.dn = has_dn ? (const T*)dn->data : NULL;
    A.g = (T*)g->data;
    A.mu = (T)mu;
    A.rows = rows;
    A.cols = cols;
    // rows that need no halo: meanwhile the boundary rows are in flight
    A.r0 = has_up ? 1 : 0;
    A.r1 = has_dn ? rows - 1 : rows;
    if (A.r1 < A.r0) A.r1 = A.r0;
    int rc = smooth_launch<T>(ctx, A, 0);
    if (rc) return rc;
    if (has_up || has_dn) {
        OPKB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
        if (has_up) {
            A.r0 = 0;
            A.r1 = 1;
            rc = smooth_launch<T>(ctx, A, 1);
            if (rc) return rc;
        }
        if (has_dn && !(has_up && rows == 1)) {
            A.r0 = rows - 1;
            A.r1 = rows;
            rc = smooth_launch<T>(ctx, A, 2);
            if (rc) return rc;
        }
    }
    double r[6];
    rc = opkb_finish_reduction(ctx, 6, NULL, r);
    if (rc) return rc;
    // fixed order: interior, first row, last row
    const double fd = (r[0] + r[2]) + r[4], fe = (r[1] + r[3]) + r[5];
    if (y) {
        *f = 0.5 * fd + (0.5 * mu) * fe;
    } else {
        f[0] = fd;
        f[1] = fe;
    }
    return 0;
}

extern "C" int opkb_smooth2d_fg(opkb_context* ctx, const opkb_vector* w, const opkb_vector* y, double mu,
                                int64_t cols, const opkb_vector* x, opkb_vector* up, opkb_vector* dn,
                                opkb_vector* g, double* f)
{
    if (!ctx || !w || !y || !x || !g || !f) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (cols < 1 || x->n < cols || x->n % cols != 0)
        return opkb_set_error(OPKB_EINVAL, "the local block must be a whole number of rows of %lld elements",
                              (long long)cols);
    if (w->n != x->n || y->n != x->n || g->n != x->n || w->eltype != x->eltype || y->eltype != x->eltype ||
        g->eltype != x->eltype || (up && up->eltype != x->eltype) || (dn && dn->eltype != x->eltype))
        return opkb_set_error(OPKB_EINVAL, "w, y, x, g must belong to the same space");
    if (x == g) return opkb_set_error(OPKB_EINVAL, "g must not alias x");
    return (x->eltype == OPKB_DOUBLE) ? smooth_run<double>(ctx, w, y, mu, cols, x, up, dn, g, f)
                                      : smooth_run<float>(ctx, w, y, mu, cols, x, up, dn, g, f);
}

// q = A.p with A = diag(w) + mu*L, the Hessian of the objective above (symmetric positive
// definite when w > 0): the linear operator of `conjgrad` for this workload, same row
// sharding and halo exchange; out[0] = p.q and out[1] = p.p come out of the same pass
// (the `vdot(p, q)` of a conjugate-gradient iteration costs no extra sweep).
extern "C" int opkb_smooth2d_apply(opkb_context* ctx, const opkb_vector* w, double mu, int64_t cols,
                                   const opkb_vector* p, opkb_vector* up, opkb_vector* dn, opkb_vector* q,
                                   double out[2])
{
    if (!ctx || !w || !p || !q || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (cols < 1 || p->n < cols || p->n % cols != 0)
        return opkb_set_error(OPKB_EINVAL, "the local block must be a whole number of rows of %lld elements",
                              (long long)cols);
    if (w->n != p->n || q->n != p->n || w->eltype != p->eltype || q->eltype != p->eltype ||
        (up && up->eltype != p->eltype) || (dn && dn->eltype != p->eltype))
        return opkb_set_error(OPKB_EINVAL, "w, p, q must belong to the same space");
    if (p == q || p->data == q->data) return opkb_set_error(OPKB_EINVAL, "q must not alias p");
    return (p->eltype == OPKB_DOUBLE) ? smooth_run<double>(ctx, w, NULL, mu, cols, p, up, dn, q, out)
                                      : smooth_run<float>(ctx, w, NULL, mu, cols, p, up, dn, q, out);
}
