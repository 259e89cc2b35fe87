// opkb_gram2.cuh -- P3, the tall-skinny Gram sweep, pipelined version.
//
// Same mathematics and role decomposition as opkb_gram.cuh (see there), other
// data movement: one persistent CTA per SM; a producer warp streams raw tiles of
// the 2m+3 vectors (v0, s_q, y_q, x, g) into a ring of NSTAGE shared-memory
// stages with bulk asynchronous copies (cp.async.bulk -> UBLKCP, completion on
// an mbarrier), so that several stages (>= 100 KB per SM) are always in flight;
// the consumer warps (one register tile of the Gram block each) wait on the
// "full" barrier of a stage, turn the newest slot into the pair (x0 - x, g0 - g)
// in place (written back to HBM once), accumulate, and release the stage
// through its "empty" barrier.  Tiles stay in the element type in shared
// memory (Float32 is widened when read), the free-set mask is p != 0 read from
// the staged v0 row.
#pragma once

#define GRAM2_MAXSTAGE 8

template <typename T> struct Gram2Args {
    const T* rows[2 * OPKB_MEM_MAX + 3];  // 0: v0; 1+2q: S_q; 2+2q: Y_q (q oldest..newest); then x, gcur
    T* Snew;
    T* Ynew;
    int64_t n;
    int m, nb, kshare, masked, tile, nstage;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <typename T, int BS>
__global__ void __launch_bounds__(352, 1) k_gram2(Gram2Args<T> A, ReduceScratch rs)
{
    constexpr int RT = 2 * BS, CT = BS + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[GRAM2_MAXSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[GRAM2_MAXSTAGE];
    __shared__ unsigned int s_last;

    const int nb = A.nb, m = A.m, k = A.kshare, TILE = A.tile, NSTAGE = A.nstage;
    const int nroles = nb * (nb + 1) / 2;
    const int cwarps = nroles * k;            // consumer warps; warp `cwarps` is the producer
    const int cthreads = cwarps * 32;
    const int NRL = 2 * m + 3;                // rows loaded per stage
    const int zrow = NRL;                     // an all-zero row per stage (pairs that do not exist)
    const int srows = NRL + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T* const stage0 = reinterpret_cast<T*>(smem_raw);
    const size_t stage_elems = (size_t)srows * TILE;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], cwarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int s = 0; s < NSTAGE; ++s)
        for (int e = tid; e < TILE; e += blockDim.x) stage0[s * stage_elems + (size_t)zrow * TILE + e] = T(0);
    fence_proxy_async();
    __syncthreads();

    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    const uint32_t rowbytes = (uint32_t)(TILE * sizeof(T));

    double acc[RT][CT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[r][c] = 0.0;

    if (warp == cwarps) {
        // ------------------------------ producer ---------------------------------
        if (lane == 0) {
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int s = it % NSTAGE;
                const uint32_t ph = (uint32_t)((it / NSTAGE) & 1);
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const int64_t base = tile * TILE;
                uint32_t bytes = rowbytes;
                if (base + TILE > A.n) bytes = (uint32_t)(((A.n - base) * sizeof(T)) & ~(size_t)15);
                mbar_expect_tx(&full_bar[s], bytes * (uint32_t)NRL);
                if (bytes > 0) {
                    T* dst = stage0 + s * stage_elems;
                    for (int r = 0; r < NRL; ++r)
                        bulk_g2s(dst + (size_t)r * TILE, A.rows[r] + base, bytes, &full_bar[s]);
                }
            }
        }
    } else if (warp < cwarps) {
        // ------------------------------ consumers --------------------------------
        const int role = warp / k, sub = warp % k;
        int bi = 0, bj = 0;
        {
            int r = role;
            for (bi = 0; bi < nb; ++bi) {
                int cnt = nb - bi;
                if (r < cnt) { bj = bi + r; break; }
                r -= cnt;
            }
        }
        // row / column -> staged row index (zrow for pairs beyond m)
        int rowoff[RT], coloff[CT];   // element offsets of the staged rows inside a stage
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int q = bi * BS + (r >> 1);
            rowoff[r] = ((q < m) ? 1 + 2 * q + (r & 1) : zrow) * TILE;
        }
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            if (bi == bj) {
                const int q = bi * BS + c - 1;
                coloff[c] = ((c == 0) ? 0 : ((q < m) ? 2 + 2 * q : zrow)) * TILE;
            } else {
                const int q = bj * BS + c;
                coloff[c] = ((c < BS && q < m) ? 2 + 2 * q : zrow) * TILE;
            }
        }
        const int rs_new = 1 + 2 * (m - 1), ry_new = 2 + 2 * (m - 1), rx = 2 * m + 1, rg = 2 * m + 2;

        int it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)((it / NSTAGE) & 1);
            T* st = stage0 + s * stage_elems;
            const int64_t base = tile * TILE;
            mbar_wait(&full_bar[s], ph);
            if (base + TILE > A.n) {
                // last, partial tile: elements the 16-byte granular bulk copy left out, zero padding
                const int valid = (int)(A.n - base);
                const int copied = (int)((((size_t)valid * sizeof(T)) & ~(size_t)15) / sizeof(T));
                for (int idx = tid; idx < NRL * TILE; idx += cthreads) {
                    const int r = idx / TILE, e = idx - r * TILE;
                    if (e >= copied) st[(size_t)r * TILE + e] = (e < valid) ? A.rows[r][base + e] : T(0);
                }
                named_bar_sync(1, cthreads);
            }
            // vupdate!(S[mark], -1, x); vupdate!(Y[mark], -1, g|p)  (src/quasinewton.jl:512-513)
            for (int e = tid; e < TILE; e += cthreads) {
                if (base + e < A.n) {
                    const T sn = st[(size_t)rs_new * TILE + e] - st[(size_t)rx * TILE + e];
                    const T yn = st[(size_t)ry_new * TILE + e] - st[(size_t)rg * TILE + e];
                    st[(size_t)rs_new * TILE + e] = sn;
                    st[(size_t)ry_new * TILE + e] = yn;
                    A.Snew[base + e] = sn;
                    A.Ynew[base + e] = yn;
                }
            }
            named_bar_sync(1, cthreads);
            // accumulate this role's register tile
#pragma unroll 2
            for (int e = sub * 32 + lane; e < TILE; e += 32 * k) {
                double rv[RT], cv[CT];
#pragma unroll
                for (int r = 0; r < RT; ++r) rv[r] = (double)st[rowoff[r] + e];
#pragma unroll
                for (int c = 0; c < CT; ++c) cv[c] = (double)st[coloff[c] + e];
                if (A.masked) {
                    const bool fr = (st[e] != T(0));  // free variable: p[i] != 0 (row 0 is v0 = p)
#pragma unroll
                    for (int c = 0; c < CT; ++c) cv[c] = fr ? cv[c] : 0.0;
                }
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int c = 0; c < CT; ++c) acc[r][c] = fma(rv[r], cv[c], acc[r][c]);
            }
            // release the stage: generic-proxy writes above must be ordered before the next bulk copy
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    }
    __syncthreads();

    // ---- CTA reduction: butterfly per warp, fixed order over the k sharing warps
    const int nslot = nroles * RT * CT;
    double* red = reinterpret_cast<double*>(smem_raw);  // [warp][RT*CT]; the ring is drained
    if (warp < cwarps) {
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                double a = acc[r][c];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
                if (lane == 0) red[warp * (RT * CT) + r * CT + c] = a;
            }
    }
    __syncthreads();
    for (int s = tid; s < nslot; s += blockDim.x) {
        const int ro = s / (RT * CT), idx = s - ro * (RT * CT);
        double a = red[(ro * k) * (RT * CT) + idx];
        for (int j = 1; j < k; ++j) a += red[(ro * k + j) * (RT * CT) + idx];
        rs.part[(size_t)blockIdx.x * nslot + s] = a;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int t = atomicAdd(rs.ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int s = tid; s < nslot; s += blockDim.x) {
        double a = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) a += __ldcg(&rs.part[(size_t)b * nslot + s]);
        rs.res[s] = a;
    }
    if (tid == 0) *rs.ticket = 0u;
}

template <typename T, int BS>
static int gram2_launch(opkb_context* ctx, Gram2Args<T>& A, int nroles, int* nslot_out)
{
    const int warps = nroles * A.kshare + 1;
    const int block = warps * 32;
    // tile / ring geometry: >= 3 stages within ~200 KB of shared memory
    const int srows = 2 * A.m + 4;
    const size_t budget = 200 * 1024;
    int tile = 512;
    while (tile > 64 && (size_t)srows * tile * sizeof(T) * 3 > budget) tile >>= 1;
    size_t stage_bytes = (size_t)srows * tile * sizeof(T);
    int nstage = (int)(budget / stage_bytes);
    if (nstage > GRAM2_MAXSTAGE) nstage = GRAM2_MAXSTAGE;
    if (nstage < 2) return opkb_set_error(OPKB_EINVAL, "Gram ring does not fit in shared memory");
    size_t smem = stage_bytes * nstage;
    const size_t red_bytes = (size_t)warps * (2 * BS) * (BS + 1) * sizeof(double);
    if (smem < red_bytes) smem = red_bytes;
    A.tile = tile;
    A.nstage = nstage;
    auto kern = k_gram2<T, BS>;
    OPKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t ntiles = (A.n + tile - 1) / tile;
    int64_t grid = ctx->sm_count;   // one persistent CTA per SM
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    ReduceScratch rs;
    rs.part = ctx->d_part;
    rs.ticket = ctx->d_ticket;
    rs.res = ctx->d_res;
    opkb_prof_begin(ctx, OPKB_K_GRAM);
    kern<<<(int)grid, block, smem, ctx->stream>>>(A, rs);
    opkb_prof_end(ctx);
    ctx->launches += 1;
    OPKB_CUDA(cudaGetLastError());
    *nslot_out = nroles * (2 * BS) * (BS + 1);
    return 0;
}

template <typename T>
static int gram2_run(opkb_context* ctx, int64_t n, int m, const void* const* Sp, const void* const* Yp,
                     const void* v0, const void* x, const void* gcur, int masked, double* G)
{
    Gram2Args<T> A;
    memset(&A, 0, sizeof(A));
    A.rows[0] = (const T*)v0;
    for (int j = 0; j < m; ++j) {
        A.rows[1 + 2 * j] = (const T*)Sp[j];
        A.rows[2 + 2 * j] = (const T*)Yp[j];
    }
    A.rows[2 * m + 1] = (const T*)x;
    A.rows[2 * m + 2] = (const T*)gcur;
    A.Snew = (T*)Sp[m - 1];
    A.Ynew = (T*)Yp[m - 1];
    A.n = n;
    A.m = m;
    A.masked = masked;
    const int nb = (m + 4) / 5;
    const int bs = (m + nb - 1) / nb;
    const int nroles = nb * (nb + 1) / 2;
    A.nb = nb;
    // consumer warps per role: 1 role: 8; 3 roles: 3 each; 6 or 10 roles: 1 each
    A.kshare = (nroles == 1) ? 8 : ((nroles == 3) ? 3 : 1);
    int nslot = 0, rc = 0;
    switch (bs) {
    case 1: rc = gram2_launch<T, 1>(ctx, A, nroles, &nslot); break;
    case 2: rc = gram2_launch<T, 2>(ctx, A, nroles, &nslot); break;
    case 3: rc = gram2_launch<T, 3>(ctx, A, nroles, &nslot); break;
    case 4: rc = gram2_launch<T, 4>(ctx, A, nroles, &nslot); break;
    default: rc = gram2_launch<T, 5>(ctx, A, nroles, &nslot); break;
    }
    if (rc) return rc;
    static thread_local double res[OPKB_NRED_MAX];
    rc = opkb_finish_reduction(ctx, nslot, NULL, res);
    if (rc) return rc;
    gram_scatter(res, m, nb, bs, G);
    return 0;
}
