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
    // conjugate-gradient step (k_stencil_strip<.., PRE = 1>): the field the stencil is
    // applied to is u = r (restart) or beta*x + r with x = the old direction; u is also
    // stored (the new direction); up / dn then hold rows of u
    const T* r;
    T* uout;
    T beta;
    int restart;
    int64_t seg_rows;         // k_stencil_strip: rows per segment (one CTA per segment and column block)
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


// ---------------------------------------------------------------------------
// Row-strip version.  One item = STRIP_ROWS consecutive rows x (256 x V) columns; a thread
// owns one pack of columns and holds the STRIP_ROWS + 2 rows of the field it needs in
// registers (every load of the strip is in flight before the first use), so every element
// of the field is fetched ONCE per strip (+ the two rows around it) instead of three times (as centre, as "above", as "below": the L2 -> SM traffic of k_smooth2d), the
// horizontal neighbours come from the adjacent lanes (shuffles; only the first and the last
// lane of a warp fetch theirs).  Same arithmetic, same order: bit-identical results.
// PRE = 1: the field is not stored but formed on the fly from two vectors, u = beta*x + r
// (`vcombine!(p, beta, p, 1, r)` of conjgrad; u = r on a restart), written to `uout`: the
// direction update, the operator and p.q, p.p of a CG iteration in ONE pass.
#define STRIP_ROWS 4

template <typename T, int V> struct StripVec { T v[V]; };

template <typename T, int V>
__device__ __forceinline__ StripVec<T, V> strip_ld(const T* p)
{
    StripVec<T, V> o;
    if (V == VecOf<T>::N) {
        const Pack<T> pk = ld_pack<T>(p, 0);
#pragma unroll
        for (int q = 0; q < V; ++q) o.v[q] = pk.v[q];
    } else {
        o.v[0] = p[0];
    }
    return o;
}

// A CTA owns one column block (256 x V columns) of one SEGMENT of rows and walks down it strip
// by strip (RB rows at a time), carrying the last two rows of the field in registers from one
// strip to the next: every element of the field is loaded ONCE (plus two rows per segment).
// (With independent strips scheduled grid-stride, the rows shared by neighbouring strips had
// left L2 -- ~20 us of streaming -- by the time the other strip asked for them: +27 % DRAM
// reads, ncu.)  Per strip, phase 1 issues every load from clamped, always valid addresses with
// no dependent instruction and no branch in between (rows that do not exist and threads beyond
// the last column load something valid that is never used), a __syncwarp() keeps ptxas from
// sinking each load next to its consumer (which had the warp wait for DRAM once per row),
// then phase 2 forms the field and applies the stencil.
template <typename T, int V, bool HESS, int PRE, int RB>
__global__ void __launch_bounds__(OPKB_BLOCK, 2) k_stencil_strip(SmoothArgs<T> A, ReduceScratch rs)
{
    const int64_t ncb = (A.cols + OPKB_BLOCK * V - 1) / (OPKB_BLOCK * V);
    const int lane = threadIdx.x & 31;
    const bool restart = (PRE == 1) && A.restart;
    // PRE = 1: u = beta*a + b with (a, b) = (old direction, r); on a restart u = b
    const T* const srcA = restart ? A.r : A.x;
    const T* const srcB = (PRE == 1) ? A.r : A.x;
    double acc[2] = {0.0, 0.0};
    const int64_t seg = blockIdx.x / ncb;
    const int64_t col = ((blockIdx.x - seg * ncb) * OPKB_BLOCK + threadIdx.x) * V;
    const bool active = col < A.cols;               // (whole warps stay in the loop: shuffles)
    const int64_t colc = active ? col : 0;
    const int64_t rs0 = A.r0 + seg * A.seg_rows;
    const int64_t rs1 = (rs0 + A.seg_rows < A.r1) ? rs0 + A.seg_rows : A.r1;

    // the field on row `row` (own rows, or the halo rows -1 and A.rows), raw loads
    auto row_ptrs = [&](int64_t row, const T*& pa, const T*& pb, bool& halo) {
        const bool above = (row < 0), below = (row >= A.rows);
        halo = (above && A.up != NULL) || (below && A.dn != NULL);
        const int64_t rowc = above ? 0 : (below ? A.rows - 1 : row);
        pa = srcA + rowc * A.cols + colc;
        pb = srcB + rowc * A.cols + colc;
        if (above && A.up != NULL) pa = pb = A.up + colc;          // (pointer selects, no branch)
        if (below && A.dn != NULL) pa = pb = A.dn + colc;
    };
    auto field = [&](const StripVec<T, V>& a, const StripVec<T, V>& b, bool halo, T (&u)[V]) {
#pragma unroll
        for (int q = 0; q < V; ++q) {
            T t = a.v[q];
            if (PRE == 1 && !restart && !halo) t = A.beta * a.v[q] + b.v[q];
            u[q] = t;
        }
    };

    T f[RB + 2][V];                                  // rows ra-1 .. ra+RB of the field
    {
        const T *pa, *pb;
        bool h0, h1;
        row_ptrs(rs0 - 1, pa, pb, h0);
        const StripVec<T, V> a0 = strip_ld<T, V>(pa), b0 = (PRE == 1) ? strip_ld<T, V>(pb) : a0;
        row_ptrs(rs0, pa, pb, h1);
        const StripVec<T, V> a1 = strip_ld<T, V>(pa), b1 = (PRE == 1) ? strip_ld<T, V>(pb) : a1;
        field(a0, b0, h0, f[0]);
        field(a1, b1, h1, f[1]);
    }
    for (int64_t ra = rs0; ra < rs1; ra += RB) {
        const int64_t rb = (ra + RB < rs1) ? ra + RB : rs1;
        // ---- phase 1: raw loads of rows ra+1 .. ra+RB, of w, y and of the warp's outer neighbours
        StripVec<T, V> fa[RB], fb[RB], wv[RB], yv[RB];
        T la[RB], lb[RB], rta[RB], rtb[RB];
        bool halo[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k) {
            const T *pa, *pb;
            row_ptrs(ra + 1 + k, pa, pb, halo[k]);
            fa[k] = strip_ld<T, V>(pa);
            if (PRE == 1) fb[k] = strip_ld<T, V>(pb);
            const int64_t row = ra + k;
            const int64_t rowc = (row < A.rows) ? row : A.rows - 1;
            const int64_t i = rowc * A.cols + colc;
            wv[k] = strip_ld<T, V>(A.w + i);
            if (!HESS) yv[k] = strip_ld<T, V>(A.y + i);
            const int64_t il = rowc * A.cols + (colc > 0 ? colc - 1 : 0);
            const int64_t ir = rowc * A.cols + (colc + V < A.cols ? colc + V : colc);
            la[k] = (lane == 0) ? srcA[il] : T(0);
            rta[k] = (lane == 31) ? srcA[ir] : T(0);
            if (PRE == 1) {
                lb[k] = (lane == 0) ? srcB[il] : T(0);
                rtb[k] = (lane == 31) ? srcB[ir] : T(0);
            }
        }
        __syncwarp();
        // ---- phase 2: the field, then the stencil row by row -------------------------------------
#pragma unroll
        for (int k = 0; k < RB; ++k) field(fa[k], (PRE == 1) ? fb[k] : fa[k], halo[k], f[k + 2]);
#pragma unroll
        for (int k = 0; k < RB; ++k) {
            const int64_t row = ra + k;
            if (row < rb) {                          // (uniform over the CTA)
                const bool hu = (row > 0) || (A.up != NULL), hd = (row + 1 < A.rows) || (A.dn != NULL);
                T lf = __shfl_up_sync(0xffffffffu, f[k + 1][V - 1], 1);
                T rt = __shfl_down_sync(0xffffffffu, f[k + 1][0], 1);
                if (lane == 0) lf = (PRE == 1 && !restart) ? A.beta * la[k] + lb[k] : la[k];
                if (lane == 31) rt = (PRE == 1 && !restart) ? A.beta * rta[k] + rtb[k] : rta[k];
                if (active) {
                    const int64_t i = row * A.cols + col;
                    T gv[V];
#pragma unroll
                    for (int q = 0; q < V; ++q)
                        smooth_elem<T, HESS>(f[k + 1][q], wv[k].v[q], HESS ? T(0) : yv[k].v[q], hu, f[k][q], hd,
                                             f[k + 2][q], col + q > 0, q > 0 ? f[k + 1][q > 0 ? q - 1 : 0] : lf,
                                             col + q + 1 < A.cols, q + 1 < V ? f[k + 1][q + 1 < V ? q + 1 : 0] : rt,
                                             A.mu, gv[q], acc);
                    if (V == VecOf<T>::N) {
                        Pack<T> pg;
#pragma unroll
                        for (int q = 0; q < V; ++q) pg.v[q] = gv[q];
                        st_pack<T>(A.g + i, 0, pg);
                        if (PRE == 1) {
                            Pack<T> pu;
#pragma unroll
                            for (int q = 0; q < V; ++q) pu.v[q] = f[k + 1][q];
                            st_pack<T>(A.uout + i, 0, pu);
                        }
                    } else {
                        A.g[i] = gv[0];
                        if (PRE == 1) A.uout[i] = f[k + 1][0];
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < V; ++q) {                // the window moves down by RB rows
            f[0][q] = f[RB][q];
            f[1][q] = f[RB + 1][q];
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
    const bool hess = (A.y == NULL), cgstep = (A.r != NULL);
    static const int use_strip = getenv("OPKB_STENCIL_STRIP") ? atoi(getenv("OPKB_STENCIL_STRIP")) : 1;
    // (one CTA per column block: rows of more than ~5e8 elements keep the grid-stride kernel)
    const bool fits = ncb * 2 <= (int64_t)OPKB_MAX_GRID * OPKB_NRED_MAX / 2;
    if (cgstep && !fits) return opkb_set_error(OPKB_EINVAL, "too many columns for the stencil step kernel");
    const bool strip = cgstep || (use_strip && fits);
    static const int srows_env = getenv("OPKB_STRIP_ROWS") ? atoi(getenv("OPKB_STRIP_ROWS")) : STRIP_ROWS;
    const int srows = (srows_env == 2) ? 2 : 4;
    int grid;
    if (strip) {
        // one CTA per (segment of rows, column block); as many segments as keep every SM's
        // resident CTAs (2 per SM: 128 registers) busy in ONE wave
        const int64_t nrows = A.r1 - A.r0;
        static const int per_sm = getenv("OPKB_STRIP_CTAS") ? atoi(getenv("OPKB_STRIP_CTAS")) : 2;
        int64_t nseg = ((int64_t)ctx->sm_count * per_sm) / ncb;
        if (nseg < 1) nseg = 1;
        int64_t seg_rows = (nrows + nseg - 1) / nseg;
        seg_rows = ((seg_rows + srows - 1) / srows) * srows;
        nseg = (nrows + seg_rows - 1) / seg_rows;
        A.seg_rows = seg_rows;
        grid = (int)(nseg * ncb);
    } else {
        // (6 CTAs of 256 threads per SM: room is left for the NCCL send/recv kernel to run alongside)
        grid = opkb_grid_for(ctx, (A.r1 - A.r0) * ncb * OPKB_BLOCK, 6);
    }
    ReduceScratch rs = opkb_make_scratch(ctx, 2 * slot);
    opkb_prof_begin(ctx, OPKB_K_OBJECTIVE);
    if (strip) {
#define STRIP_GO(VV, HH, PP)                                                                      \
    do {                                                                                          \
        if (srows == 2) k_stencil_strip<T, VV, HH, PP, 2><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs); \
        else k_stencil_strip<T, VV, HH, PP, 4><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);      \
    } while (0)
        if (cgstep && vec) STRIP_GO(V, true, 1);
        else if (cgstep) STRIP_GO(1, true, 1);
        else if (vec && hess) STRIP_GO(V, true, 0);
        else if (vec) STRIP_GO(V, false, 0);
        else if (hess) STRIP_GO(1, true, 0);
        else STRIP_GO(1, false, 0);
#undef STRIP_GO
    } else {
        if (vec && hess) k_smooth2d<T, V, true><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);
        else if (vec) k_smooth2d<T, V, false><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);
        else if (hess) k_smooth2d<T, 1, true><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);
        else k_smooth2d<T, 1, false><<<grid, OPKB_BLOCK, 0, ctx->stream>>>(A, rs);
    }
    opkb_prof_end(ctx);
    ctx->launches += 1;
    OPKB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int smooth_run(opkb_context* ctx, const opkb_vector* w, const opkb_vector* y, double mu, int64_t cols,
                      const opkb_vector* x, opkb_vector* up, opkb_vector* dn, opkb_vector* g, double* f,
                      const opkb_vector* cg_r = NULL, opkb_vector* cg_out = NULL, double beta = 0.0,
                      int restart = 0, opkb_vector* up_r = NULL, opkb_vector* dn_r = NULL)
{
    // cg_r != NULL: conjugate-gradient step (opkb_cg_smooth2d_step): x is the old direction,
    // the field is u = cg_r | beta*x + cg_r, written to cg_out; what travels between the ranks
    // are the boundary rows of cg_r (into up_r / dn_r) and up / dn, which hold the neighbours'
    // rows of the direction, are advanced with the same formula
    // y == NULL: operator mode, f[0] = x.g, f[1] = x.x (opkb_smooth2d_apply)
    const int64_t rows = x->n / cols;
    const bool has_up = ctx->nranks > 1 && ctx->rank > 0;
    const bool has_dn = ctx->nranks > 1 && ctx->rank < ctx->nranks - 1;
    if ((has_up && (!up || up->n != cols)) || (has_dn && (!dn || dn->n != cols)))
        return opkb_set_error(OPKB_EINVAL, "halo rows of %lld elements are needed", (long long)cols);
    OPKB_CUDA(cudaMemsetAsync(opkb_res_ptr(ctx), 0, 6 * sizeof(double), ctx->stream));
    if (ctx->nranks > 1) {
        if (!ctx->halo_stream) {
            OPKB_CUDA(cudaStreamCreateWithFlags(&ctx->halo_stream, cudaStreamNonBlocking));
            OPKB_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
            OPKB_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
        }
        // x is ready on the main stream -> exchange the boundary rows on the halo stream
        OPKB_CUDA(cudaEventRecord(ctx->ev_ready, ctx->stream));
        OPKB_CUDA(cudaStreamWaitEvent(ctx->halo_stream, ctx->ev_ready, 0));
        const T* xd = (const T*)(cg_r ? cg_r->data : x->data);
        if (cg_r && ((has_up && (!up_r || up_r->n != cols)) || (has_dn && (!dn_r || dn_r->n != cols))))
            return opkb_set_error(OPKB_EINVAL, "halo rows of %lld elements are needed", (long long)cols);
        int rc = opkb_nccl_halo(ctx, x->eltype, (size_t)cols, xd,
                                has_up ? (cg_r ? up_r->data : up->data) : NULL, xd + (rows - 1) * cols,
                                has_dn ? (cg_r ? dn_r->data : dn->data) : NULL, ctx->halo_stream);
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
    A.r = cg_r ? (const T*)cg_r->data : NULL;
    A.uout = cg_out ? (T*)cg_out->data : NULL;
    A.beta = (T)beta;
    A.restart = restart;
    // rows that need no halo: meanwhile the boundary rows are in flight
    A.r0 = has_up ? 1 : 0;
    A.r1 = has_dn ? rows - 1 : rows;
    if (A.r1 < A.r0) A.r1 = A.r0;
    int rc = smooth_launch<T>(ctx, A, 0);
    if (rc) return rc;
    if (has_up || has_dn) {
        OPKB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
        if (cg_r) {
            // the neighbours' rows of the new direction, by the formula they use themselves
            if (has_up) rc = restart ? opkb_vcopy(ctx, up, up_r) : opkb_vcombine(ctx, up, beta, up, 1.0, up_r);
            if (rc) return rc;
            if (has_dn) rc = restart ? opkb_vcopy(ctx, dn, dn_r) : opkb_vcombine(ctx, dn, beta, dn, 1.0, dn_r);
            if (rc) return rc;
        }
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

// One pass of `conjgrad!` for the stencil operator: p <- r (restart) | beta*p + r
// (`vcombine!(p, beta, p, 1, r)`), q = A.p, out[0] = p.q, out[1] = p.p: read p, r, w; write p, q
// = 5 passes (k_stencil_strip<.., PRE = 1>).  The new direction is written to `pscratch` and
// the two buffers are swapped, so that p is updated "in place" for the caller while the
// neighbours of an element still see the old direction during the pass.  With several ranks
// only the boundary rows of r travel (into up_r / dn_r); up / dn keep the neighbours'
// boundary rows of the direction from one call to the next.
extern "C" int opkb_cg_smooth2d_step(opkb_context* ctx, const opkb_vector* w, double mu, int64_t cols,
                                     opkb_vector* p, opkb_vector* pscratch, const opkb_vector* r,
                                     opkb_vector* q, opkb_vector* up, opkb_vector* dn, opkb_vector* up_r,
                                     opkb_vector* dn_r, double beta, int restart, double out[2])
{
    if (!ctx || !w || !p || !pscratch || !r || !q || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (cols < 1 || p->n < cols || p->n % cols != 0)
        return opkb_set_error(OPKB_EINVAL, "the local block must be a whole number of rows of %lld elements",
                              (long long)cols);
    const opkb_vector* all[4] = {w, pscratch, r, q};
    for (int k = 0; k < 4; ++k)
        if (all[k]->n != p->n || all[k]->eltype != p->eltype)
            return opkb_set_error(OPKB_EINVAL, "w, p, pscratch, r, q must belong to the same space");
    if ((up && up->eltype != p->eltype) || (dn && dn->eltype != p->eltype) ||
        (up_r && up_r->eltype != p->eltype) || (dn_r && dn_r->eltype != p->eltype))
        return opkb_set_error(OPKB_EINVAL, "halo rows of another element type");
    if (p->data == pscratch->data || p->data == q->data || pscratch->data == q->data || r->data == q->data ||
        r->data == pscratch->data || r->data == p->data || !p->owned || !pscratch->owned)
        return opkb_set_error(OPKB_EINVAL, "p, pscratch, r, q must be distinct vectors (p, pscratch owned)");
    int rc = (p->eltype == OPKB_DOUBLE)
                 ? smooth_run<double>(ctx, w, NULL, mu, cols, p, up, dn, q, out, r, pscratch, beta, restart, up_r, dn_r)
                 : smooth_run<float>(ctx, w, NULL, mu, cols, p, up, dn, q, out, r, pscratch, beta, restart, up_r, dn_r);
    if (rc) return rc;
    void* t = p->data;
    p->data = pscratch->data;
    pscratch->data = t;
    return 0;
}
