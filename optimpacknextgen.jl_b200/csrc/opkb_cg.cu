// opkb_cg.cu -- fused passes of the linear conjugate gradient `conjgrad!` that the
// reference re-exports from LazyAlgebra (src/OptimPackNextGen.jl:18-19, :36; call site
// test/conjgrad-tests.jl:30): SURVEY 8(f)-4.  One iteration of the reference is
//     vcombine!(p, beta, p, 1, r)       3 passes      |  direction
//     A(q, p)                           operator      |
//     gamma = vdot(p, q)                2 passes      |  curvature
//     vupdate!(x, alpha, p)             3 passes      |
//     [vnorm2(p), vnorm2(x)             2 passes]     |  update
//     vupdate!(r, -alpha, q)            3 passes      |
//     rho = vdot(r, r)                  1 pass        |
// Here the update is ONE kernel (read x, p, r, q; write x, r; rho and |x|^2 reduced in the
// same pass: 6 passes instead of 9), the curvature rides on the operator when the operator
// is built in (opkb_cg_diag_step: direction + A.p + p.q + p.p in 5 passes;
// opkb_smooth2d_apply in opkb_stencil.cu: A.p + p.q + p.p in 3), and a user operator costs
// one extra 2-pass kernel (opkb_cg_curvature).  Element-wise arithmetic in the element type
// without contraction (same bits as the vupdate!/vcombine! sequence), sums of exact
// products in double, fixed reduction order, ranks combined by the packed all-gather.
#include "opkb_internal.cuh"

static ReduceScratch cg_scratch(opkb_context* ctx)
{
    return opkb_make_scratch(ctx);
}

#define CG_LAUNCH(ctx, cls, kernel, grid, ...)                               \
    do {                                                                     \
        opkb_prof_begin((ctx), (cls));                                       \
        kernel<<<(grid), OPKB_BLOCK, 0, (ctx)->stream>>>(__VA_ARGS__);       \
        opkb_prof_end((ctx));                                                \
        (ctx)->launches += 1;                                                \
        OPKB_CUDA(cudaGetLastError());                                       \
    } while (0)

static bool cg_same(const opkb_vector* a, const opkb_vector* b)
{
    return a && b && a->ctx == b->ctx && a->eltype == b->eltype && a->n == b->n;
}

// x += alpha*p; r -= alpha*q; acc[0] = r.r; acc[1] = x.x     (two packs in flight per stream)
template <typename T>
__global__ void __launch_bounds__(OPKB_BLOCK) k_cg_update(T* __restrict__ x, T* __restrict__ r,
                                                          const T* __restrict__ p, const T* __restrict__ q,
                                                          T alpha, int64_t n, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc[2] = {0.0, 0.0};
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += 2 * stride) {
        const int64_t iv2 = iv + stride;
        const bool two = iv2 < nvec;
        Pack<T> x0 = ld_pack<T>(x, iv), p0 = ld_pack<T>(p, iv), r0 = ld_pack<T>(r, iv), q0 = ld_pack<T>(q, iv);
        Pack<T> x1 = x0, p1 = p0, r1 = r0, q1 = q0;
        if (two) {
            x1 = ld_pack<T>(x, iv2);
            p1 = ld_pack<T>(p, iv2);
            r1 = ld_pack<T>(r, iv2);
            q1 = ld_pack<T>(q, iv2);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            x0.v[e] = x0.v[e] + alpha * p0.v[e];
            r0.v[e] = r0.v[e] - alpha * q0.v[e];     // y + (-a)*x: same bits
            acc[0] = dacc(r0.v[e], r0.v[e], acc[0]);
            acc[1] = dacc(x0.v[e], x0.v[e], acc[1]);
        }
        st_pack<T>(x, iv, x0);
        st_pack<T>(r, iv, r0);
        if (two) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                x1.v[e] = x1.v[e] + alpha * p1.v[e];
                r1.v[e] = r1.v[e] - alpha * q1.v[e];
                acc[0] = dacc(r1.v[e], r1.v[e], acc[0]);
                acc[1] = dacc(x1.v[e], x1.v[e], acc[1]);
            }
            st_pack<T>(x, iv2, x1);
            st_pack<T>(r, iv2, r1);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < n; ++i) {
            const T xn = x[i] + alpha * p[i], rn = r[i] - alpha * q[i];
            x[i] = xn;
            r[i] = rn;
            acc[0] = dacc(rn, rn, acc[0]);
            acc[1] = dacc(xn, xn, acc[1]);
        }
    }
    block_reduce_publish<2, 0u, 0u>(acc, rs);
}

// acc[0] = p.q; acc[1] = p.p
template <typename T>
__global__ void __launch_bounds__(OPKB_BLOCK) k_cg_curvature(const T* __restrict__ p, const T* __restrict__ q,
                                                             int64_t n, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc[2] = {0.0, 0.0};
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += 2 * stride) {
        const int64_t iv2 = iv + stride;
        const bool two = iv2 < nvec;
        const Pack<T> p0 = ld_pack<T>(p, iv), q0 = ld_pack<T>(q, iv);
        Pack<T> p1 = p0, q1 = q0;
        if (two) {
            p1 = ld_pack<T>(p, iv2);
            q1 = ld_pack<T>(q, iv2);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            acc[0] = dacc(p0.v[e], q0.v[e], acc[0]);
            acc[1] = dacc(p0.v[e], p0.v[e], acc[1]);
        }
        if (two) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                acc[0] = dacc(p1.v[e], q1.v[e], acc[0]);
                acc[1] = dacc(p1.v[e], p1.v[e], acc[1]);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < n; ++i) {
            acc[0] = dacc(p[i], q[i], acc[0]);
            acc[1] = dacc(p[i], p[i], acc[1]);
        }
    }
    block_reduce_publish<2, 0u, 0u>(acc, rs);
}

// p = r (RESTART) or beta*p + r; q = a.*p; acc[0] = p.q; acc[1] = p.p
template <typename T, bool RESTART>
__global__ void __launch_bounds__(OPKB_BLOCK) k_cg_diag_step(const T* __restrict__ a, T* __restrict__ p,
                                                             const T* __restrict__ r, T* __restrict__ q, T beta,
                                                             int64_t n, ReduceScratch rs)
{
    constexpr int VEC = VecOf<T>::N;
    const int64_t nvec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc[2] = {0.0, 0.0};
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
        const Pack<T> a0 = ld_pack<T>(a, iv), r0 = ld_pack<T>(r, iv);
        Pack<T> p0 = r0, q0;
        if (!RESTART) p0 = ld_pack<T>(p, iv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (!RESTART) p0.v[e] = beta * p0.v[e] + r0.v[e];
            q0.v[e] = a0.v[e] * p0.v[e];
            acc[0] = dacc(p0.v[e], q0.v[e], acc[0]);
            acc[1] = dacc(p0.v[e], p0.v[e], acc[1]);
        }
        st_pack<T>(p, iv, p0);
        st_pack<T>(q, iv, q0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * VEC; i < n; ++i) {
            const T pn = RESTART ? r[i] : beta * p[i] + r[i];
            const T qn = a[i] * pn;
            p[i] = pn;
            q[i] = qn;
            acc[0] = dacc(pn, qn, acc[0]);
            acc[1] = dacc(pn, pn, acc[1]);
        }
    }
    block_reduce_publish<2, 0u, 0u>(acc, rs);
}

extern "C" int opkb_cg_update(opkb_context* ctx, opkb_vector* x, opkb_vector* r, const opkb_vector* p,
                              const opkb_vector* q, double alpha, double out[2])
{
    if (!ctx || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (!cg_same(x, r) || !cg_same(x, p) || !cg_same(x, q))
        return opkb_set_error(OPKB_EINVAL, "vectors of different spaces");
    if (x->data == r->data || x->data == p->data || x->data == q->data || r->data == p->data ||
        r->data == q->data)
        return opkb_set_error(OPKB_EINVAL, "x, r, p, q must be distinct vectors");
    const int grid = opkb_grid_for(ctx, x->n / (x->eltype == OPKB_DOUBLE ? 2 : 4) / 2, 4);
    if (x->eltype == OPKB_DOUBLE)
        CG_LAUNCH(ctx, OPKB_K_ELEMENTWISE, k_cg_update<double>, grid, (double*)x->data, (double*)r->data,
                  (const double*)p->data, (const double*)q->data, alpha, x->n, cg_scratch(ctx));
    else
        CG_LAUNCH(ctx, OPKB_K_ELEMENTWISE, k_cg_update<float>, grid, (float*)x->data, (float*)r->data,
                  (const float*)p->data, (const float*)q->data, (float)alpha, x->n, cg_scratch(ctx));
    return opkb_finish_reduction(ctx, 2, NULL, out);
}

extern "C" int opkb_cg_curvature(opkb_context* ctx, const opkb_vector* p, const opkb_vector* q, double out[2])
{
    if (!ctx || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (!cg_same(p, q)) return opkb_set_error(OPKB_EINVAL, "vectors of different spaces");
    const int grid = opkb_grid_for(ctx, p->n / (p->eltype == OPKB_DOUBLE ? 2 : 4) / 2, 4);
    if (p->eltype == OPKB_DOUBLE)
        CG_LAUNCH(ctx, OPKB_K_REDUCE, k_cg_curvature<double>, grid, (const double*)p->data,
                  (const double*)q->data, p->n, cg_scratch(ctx));
    else
        CG_LAUNCH(ctx, OPKB_K_REDUCE, k_cg_curvature<float>, grid, (const float*)p->data, (const float*)q->data,
                  p->n, cg_scratch(ctx));
    return opkb_finish_reduction(ctx, 2, NULL, out);
}

extern "C" int opkb_cg_diag_step(opkb_context* ctx, const opkb_vector* a, opkb_vector* p, const opkb_vector* r,
                                 opkb_vector* q, double beta, int restart, double out[2])
{
    if (!ctx || !out) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (!cg_same(a, p) || !cg_same(a, r) || !cg_same(a, q))
        return opkb_set_error(OPKB_EINVAL, "vectors of different spaces");
    if (p->data == r->data || p->data == q->data || q->data == r->data || q->data == a->data ||
        p->data == a->data)
        return opkb_set_error(OPKB_EINVAL, "a, p, r, q must be distinct vectors");
    const int grid = opkb_grid_for(ctx, p->n / (p->eltype == OPKB_DOUBLE ? 2 : 4), 8);
    if (p->eltype == OPKB_DOUBLE) {
        if (restart)
            CG_LAUNCH(ctx, OPKB_K_ELEMENTWISE, (k_cg_diag_step<double, true>), grid, (const double*)a->data,
                      (double*)p->data, (const double*)r->data, (double*)q->data, beta, p->n, cg_scratch(ctx));
        else
            CG_LAUNCH(ctx, OPKB_K_ELEMENTWISE, (k_cg_diag_step<double, false>), grid, (const double*)a->data,
                      (double*)p->data, (const double*)r->data, (double*)q->data, beta, p->n, cg_scratch(ctx));
    } else {
        if (restart)
            CG_LAUNCH(ctx, OPKB_K_ELEMENTWISE, (k_cg_diag_step<float, true>), grid, (const float*)a->data,
                      (float*)p->data, (const float*)r->data, (float*)q->data, (float)beta, p->n,
                      cg_scratch(ctx));
        else
            CG_LAUNCH(ctx, OPKB_K_ELEMENTWISE, (k_cg_diag_step<float, false>), grid, (const float*)a->data,
                      (float*)p->data, (const float*)r->data, (float*)q->data, (float)beta, p->n,
                      cg_scratch(ctx));
    }
    return opkb_finish_reduction(ctx, 2, NULL, out);
}
