// opkb_gram.cuh -- P3: the tall-skinny Gram sweep of the L-BFGS two-loop
// recursion (north_star: "the m dot products batched as one S^T.g / Y^T.g
// sweep").  One pass over S[1..m], Y[1..m], v0 (and x, g for the in-place pair
// update of the newest slot) yields every inner product apply_lbfgs!
// (src/quasinewton.jl:716-779) needs:
//     s_a.v0, y_a.v0, s_a.y_b, y_a.y_b   for age(b) >= age(a),
// restricted to the free set {i : p[i] != 0} for VMLMB (:758-771).
//
// Layout.  A CTA stages a tile of TILE elements of the 2m+1 vectors in shared
// memory as doubles (Float32 data is widened there, so products are exact).
// The pairs are split in NB blocks of BS pairs; warp "roles" own one register
// tile each: rows {s_q, y_q : q in block bi} (2*BS rows) times columns
// {v0, y_q : q in block bi} (diagonal role) or {y_q : q in block bj > bi}.
// Only the NB(NB+1)/2 upper-triangular roles exist (age(col) >= age(row)).
// Each lane owns the elements e = lane (mod 32) of the tile, so shared-memory
// reads are conflict free, and keeps 2*BS*(BS+1) double accumulators for the
// whole sweep.  Partials go through the same deterministic CTA -> last-CTA
// scheme as every other reduction of the library.
#pragma once

#define GRAM_TILE 256

template <typename T> struct GramArgs {
    const T* S[OPKB_MEM_MAX];   // oldest .. newest
    const T* Y[OPKB_MEM_MAX];
    const T* v0;
    const T* x;      // for the in-place update of the newest pair
    const T* gcur;
    T* Snew;         // = S[m-1], written
    T* Ynew;
    int64_t n;
    int m, nb, kshare, masked;
};

// number of vector rows kept in shared memory: v0, (s,y) x nb*BS, mask
template <int BS> __host__ __device__ constexpr int gram_rows(int nb) { return 2 + 2 * nb * BS; }

template <typename T, int BS>
__global__ void __launch_bounds__(320) k_gram(GramArgs<T> A, ReduceScratch rs)
{
    constexpr int RT = 2 * BS, CT = BS + 1, TILE = GRAM_TILE, VEC = VecOf<T>::N;
    constexpr int PPT = TILE / VEC;  // packs per tile row
    extern __shared__ double smem[];
    __shared__ unsigned int s_last;
    const int nb = A.nb, m = A.m, k = A.kshare;
    const int nrows = gram_rows<BS>(nb);
    const int maskrow = nrows - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int role = warp / k, sub = warp % k;
    const int nroles = nb * (nb + 1) / 2;

    // decode the role -> (bi, bj), bj >= bi
    int bi = 0, bj = 0;
    {
        int r = role;
        for (bi = 0; bi < nb; ++bi) {
            int cnt = nb - bi;
            if (r < cnt) { bj = bi + r; break; }
            r -= cnt;
        }
    }
    const bool active = role < nroles;
    const int rowbase = 1 + 2 * bi * BS;
    int colidx[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        if (bi == bj) colidx[c] = (c == 0) ? 0 : 2 + 2 * (bi * BS + c - 1);
        else colidx[c] = (c < BS) ? 2 + 2 * (bj * BS + c) : 0;
    }

    double acc[RT][CT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[r][c] = 0.0;

    // rows of pairs that do not exist (m < nb*BS) stay zero for the whole sweep
    for (int i = tid; i < nrows * TILE; i += blockDim.x) smem[i] = 0.0;
    __syncthreads();

    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    const int nv = 2 * m + 1;  // live rows: 0 .. 2m
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        const bool full = (base + TILE <= A.n);
        // ---- stage the tile -------------------------------------------------
        if (full) {
            for (int flat = tid; flat < nv * PPT; flat += blockDim.x) {
                const int row = flat / PPT, pk = flat - row * PPT;
                const int64_t iv = base / VEC + pk;
                Pack<T> v;
                if (row == 0) {
                    v = ld_pack<T>(A.v0, iv);
                    if (A.masked) {
#pragma unroll
                        for (int e = 0; e < VEC; ++e)
                            smem[maskrow * TILE + pk * VEC + e] = (v.v[e] != T(0)) ? 1.0 : 0.0;
                    }
                } else {
                    const int q = (row - 1) >> 1;
                    const bool isy = ((row - 1) & 1) != 0;
                    const T* src = isy ? A.Y[q] : A.S[q];
                    v = ld_pack<T>(src, iv);
                    if (q == m - 1) {
                        // vupdate!(S[mark], -1, x); vupdate!(Y[mark], -1, g|p) :512-513
                        Pack<T> cur = ld_pack<T>(isy ? A.gcur : A.x, iv);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) v.v[e] = v.v[e] - cur.v[e];
                        st_pack<T>(isy ? A.Ynew : A.Snew, iv, v);
                    }
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) smem[row * TILE + pk * VEC + e] = (double)v.v[e];
            }
        } else {
            for (int flat = tid; flat < nv * TILE; flat += blockDim.x) {
                const int row = flat / TILE, e = flat - row * TILE;
                const int64_t i = base + e;
                T v = T(0);
                if (i < A.n) {
                    if (row == 0) {
                        v = A.v0[i];
                    } else {
                        const int q = (row - 1) >> 1;
                        const bool isy = ((row - 1) & 1) != 0;
                        v = isy ? A.Y[q][i] : A.S[q][i];
                        if (q == m - 1) {
                            v = v - (isy ? A.gcur[i] : A.x[i]);
                            if (isy) A.Ynew[i] = v; else A.Snew[i] = v;
                        }
                    }
                }
                if (row == 0 && A.masked) smem[maskrow * TILE + e] = (v != T(0)) ? 1.0 : 0.0;
                smem[row * TILE + e] = (double)v;
            }
        }
        __syncthreads();
        // ---- accumulate this role's register tile ---------------------------
        if (active) {
            for (int e = sub * 32 + lane; e < TILE; e += 32 * k) {
                double rv[RT], cv[CT];
#pragma unroll
                for (int r = 0; r < RT; ++r) rv[r] = smem[(rowbase + r) * TILE + e];
                if (A.masked) {
                    const double mk = smem[maskrow * TILE + e];
#pragma unroll
                    for (int c = 0; c < CT; ++c) cv[c] = smem[colidx[c] * TILE + e] * mk;
                } else {
#pragma unroll
                    for (int c = 0; c < CT; ++c) cv[c] = smem[colidx[c] * TILE + e];
                }
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int c = 0; c < CT; ++c) acc[r][c] = fma(rv[r], cv[c], acc[r][c]);
            }
        }
        __syncthreads();
    }

    // ---- CTA reduction: butterfly per warp, fixed order over the k sharing warps
    const int nslot = nroles * RT * CT;
    double* red = smem;  // [warp][RT*CT], tiles are done
    if (active) {
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
static int gram_launch(opkb_context* ctx, const GramArgs<T>& A, int nroles, int* nslot_out)
{
    const int nb = A.nb;
    const int warps = nroles * A.kshare;
    const int block = warps * 32;
    const size_t smem = (size_t)gram_rows<BS>(nb) * GRAM_TILE * sizeof(double);
    auto kern = k_gram<T, BS>;
    OPKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    OPKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem));
    if (per_sm < 1) return opkb_set_error(OPKB_ECUDA, "Gram kernel does not fit on an SM");
    if (per_sm > 3) per_sm = 3;
    int64_t ntiles = (A.n + GRAM_TILE - 1) / GRAM_TILE;
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    if (grid > OPKB_MAX_GRID) grid = OPKB_MAX_GRID;
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

// scatter the role tiles into G[2m][m+1] (row 2q: s_q, 2q+1: y_q; col 0: v0, 1+q: y_q)
static void gram_scatter(const double* res, int m, int nb, int bs, double* G)
{
    const int RT = 2 * bs, CT = bs + 1, W = m + 1;
    for (int i = 0; i < 2 * m * W; ++i) G[i] = 0.0;
    int role = 0;
    for (int bi = 0; bi < nb; ++bi) {
        for (int bj = bi; bj < nb; ++bj, ++role) {
            for (int r = 0; r < RT; ++r) {
                const int qrow = bi * bs + r / 2;
                if (qrow >= m) continue;
                for (int c = 0; c < CT; ++c) {
                    int col;
                    if (bi == bj) {
                        col = (c == 0) ? 0 : 1 + bi * bs + (c - 1);
                    } else {
                        if (c >= bs) continue;
                        col = 1 + bj * bs + c;
                    }
                    if (col > m) continue;
                    G[(2 * qrow + (r & 1)) * W + col] = res[role * RT * CT + r * CT + c];
                }
            }
        }
    }
}

template <typename T>
static int gram_run(opkb_context* ctx, int64_t n, int m, const void* const* Sp, const void* const* Yp,
                    const void* v0, const void* x, const void* gcur, int masked, double* G)
{
    GramArgs<T> A;
    memset(&A, 0, sizeof(A));
    for (int j = 0; j < m; ++j) {
        A.S[j] = (const T*)Sp[j];
        A.Y[j] = (const T*)Yp[j];
    }
    A.v0 = (const T*)v0;
    A.x = (const T*)x;
    A.gcur = (const T*)gcur;
    A.Snew = (T*)Sp[m - 1];
    A.Ynew = (T*)Yp[m - 1];
    A.n = n;
    A.m = m;
    A.masked = masked;
    const int nb = (m + 4) / 5;
    const int bs = (m + nb - 1) / nb;
    const int nroles = nb * (nb + 1) / 2;
    A.nb = nb;
    A.kshare = (nroles >= 8) ? 1 : (8 / nroles);  // 1 role: 8 warps; 3: 2 each; 6, 10: 1 each
    int nslot = 0, rc = 0;
    switch (bs) {
    case 1: rc = gram_launch<T, 1>(ctx, A, nroles, &nslot); break;
    case 2: rc = gram_launch<T, 2>(ctx, A, nroles, &nslot); break;
    case 3: rc = gram_launch<T, 3>(ctx, A, nroles, &nslot); break;
    case 4: rc = gram_launch<T, 4>(ctx, A, nroles, &nslot); break;
    default: rc = gram_launch<T, 5>(ctx, A, nroles, &nslot); break;
    }
    if (rc) return rc;
    static thread_local double res[OPKB_NRED_MAX];
    rc = opkb_finish_reduction(ctx, nslot, NULL, res);
    if (rc) return rc;
    gram_scatter(res, m, nb, bs, G);
    return 0;
}
