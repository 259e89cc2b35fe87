// opkb_gram2.cuh -- P3: the tall-skinny Gram sweep of the L-BFGS two-loop
// recursion (north_star: "the m dot products batched as one S^T.g / Y^T.g
// sweep").  One pass over S[1..m], Y[1..m], v0 (and x, g for the in-place pair
// update of the newest slot) yields every inner product apply_lbfgs!
// (src/quasinewton.jl:716-779) needs:
//     s_a.v0, y_a.v0, s_a.y_b, y_a.y_b   for age(b) >= age(a),
// restricted to the free set {i : p[i] != 0} for VMLMB (:758-771).
//
// Decomposition.  The pairs are split in NB blocks of BS <= 5 pairs; warp
// "roles" own one register tile each: rows {s_q, y_q : q in block bi} (2*BS
// rows) times columns {v0, y_q : q in block bi} (diagonal role) or
// {y_q : q in block bj > bi}.  Only the NB(NB+1)/2 upper-triangular roles exist
// (age(col) >= age(row)).  Each lane owns the elements e = lane (mod 32) of a
// tile, so shared-memory reads are conflict free, and keeps 2*BS*(BS+1) double
// accumulators for the whole sweep (exact products: Float32 is widened when
// read).  Partials go through the same deterministic CTA -> last-CTA scheme as
// every other reduction of the library.
//
// Data movement.  One persistent CTA per SM; a producer warp streams raw tiles
// of the 2m+3 vectors (v0, x, g, s_q, y_q) into a ring of NSTAGE shared-memory
// stages with bulk asynchronous copies (cp.async.bulk -> UBLKCP, completion on
// an mbarrier), so that ~100 KB per SM are always in flight; the consumer warps
// wait on the "full" barrier of a stage, accumulate, and release the stage
// through its "empty" barrier.  The free-set mask is p != 0 read from the staged
// v0 row.
#pragma once
#include <stdlib.h>

#define GRAM2_MAXSTAGE 8

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

template <typename T> struct Gram2Args {
    const T* rows[2 * OPKB_MEM_MAX + 3];  // 0: v0; 1: x; 2: g|p; 3+2q: S_q; 4+2q: Y_q (q oldest..newest)
    T* Snew;
    T* Ynew;
    int64_t n;
    int m, nb, kshare, masked, tile, nstage;
    // HISTORY: the pair rows hold the iterates (x_q, g_q), pair q = (x_q - x_{q+1}, g_q - g_{q+1}) with
    // (x, g) as the successor of the newest one; the consumers turn a stage into differences before
    // they accumulate, nothing is written back (2m + 3 passes, all reads)
    int hist;
    int coop;     // several roles: the newest pair is formed by all consumer threads (OPKB_GRAM_COOP=0: by the writer role)
};

// Stage layout (rows of TILE elements of T): 0: v0, 1: x, 2: g|p, then the pairs
// 3+2q: s_q, 4+2q: y_q for q = 0 (oldest) .. nb*BS-1 (rows of pairs q >= m are
// zero and never overwritten), then 2 pad rows.  With TILE a compile-time
// constant every shared-memory read of the inner loop is `LDS [base + imm]`.
//
// MAXT: 256 threads (<= 7 consumer warps + producer: 2 warps per SM sub-partition,
// so up to 255 registers per thread and no spills) or 352 (3 warps per
// sub-partition, 168 registers: the 10-role case m >= 16)
template <typename T, int BS, int TILE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_gram2(Gram2Args<T> A, ReduceScratch rs)
{
    constexpr int RT = 2 * BS, CT = BS + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[GRAM2_MAXSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[GRAM2_MAXSTAGE];
    __shared__ __align__(8) uint64_t pair_bar[GRAM2_MAXSTAGE];
    __shared__ unsigned int s_last;

    const int nb = A.nb, m = A.m, k = A.kshare, NSTAGE = A.nstage;
    const int nroles = nb * (nb + 1) / 2;
    const int cwarps = nroles * k;            // consumer warps; warp `cwarps` is the producer
    const int cthreads = cwarps * 32;
    const int NRL = 2 * m + 3;                // rows loaded per stage
    const int srows = 3 + 2 * nb * BS + 2;    // rows per stage
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T* const stage0 = reinterpret_cast<T*>(smem_raw);
    const int stage_elems = srows * TILE;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], cwarps);
            mbar_init(&pair_bar[s], k);   // the k warps of the writer role
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // rows the bulk copies never write: zero once
    for (int s = 0; s < NSTAGE; ++s)
        for (int i = NRL * TILE + tid; i < stage_elems; i += blockDim.x) stage0[s * stage_elems + i] = T(0);
    fence_proxy_async();
    __syncthreads();

    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    constexpr uint32_t rowbytes = (uint32_t)(TILE * sizeof(T));

    double acc[RT][CT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[r][c] = 0.0;

    const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), pair_a = smem_u32(pair_bar);
    if (warp == cwarps) {
        // ------------------------------ producer ---------------------------------
        // lane 0 waits for the stage and arms its barrier, then every lane issues
        // the bulk copies of its rows (row r -> lane r mod 32)
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)((it / NSTAGE) & 1);
            const int64_t base = tile * TILE;
            uint32_t bytes = rowbytes;
            if (base + TILE > A.n) bytes = (uint32_t)(((A.n - base) * sizeof(T)) & ~(size_t)15);
            if (lane == 0) {
                mbar_wait_a(empty_a + 8u * s, ph ^ 1u);
                mbar_expect_tx_a(full_a + 8u * s, bytes * (uint32_t)NRL);
            }
            __syncwarp();
            if (bytes > 0) {
                T* dst = stage0 + s * stage_elems;
                for (int r = lane; r < NRL; r += 32)
                    bulk_g2s(dst + r * TILE, A.rows[r] + base, bytes, &full_bar[s]);
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
        const bool diag = (bi == bj);
        // element offsets inside a stage of: the first row of the register tile,
        // its column 0 and its column 1 (columns c >= 1 are 2 rows apart)
        const int rowbase = (3 + 2 * bi * BS) * TILE;
        const int col0 = diag ? 0 : (4 + 2 * bj * BS) * TILE;
        const int col1 = diag ? (4 + 2 * bi * BS) * TILE : (6 + 2 * bj * BS) * TILE;
        // The newest slot still holds (x0, g0).  The warps of the last diagonal
        // role (the "writer") turn it in place, in the stage, into the pair
        // (x0 - x, g0 - g) -- vupdate!(S[mark], -1, x); vupdate!(Y[mark], -1, g|p),
        // src/quasinewton.jl:512-513 -- write it back to HBM once, and signal
        // pair_bar; only the roles that read the newest pair wait for it.
        const int qn = m - 1;
        const bool writer = (bi == nb - 1 && bj == nb - 1);
        const bool needs_pair = !writer && (bj == nb - 1);   // its columns include y_new
        const int snew = (3 + 2 * qn) * TILE, ynew = (4 + 2 * qn) * TILE;
        const bool masked = A.masked != 0;
        const bool coop = (nroles > 1) && A.coop;    // all consumers form the newest pair (see below)

        int it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)((it / NSTAGE) & 1);
            T* st = stage0 + s * stage_elems;
            const int64_t base = tile * TILE;
            mbar_wait_a(full_a + 8u * s, ph);
            const bool partial = (base + TILE > A.n);
            if (partial) {
                // last tile: elements the 16-byte granular bulk copy left out, zero padding
                const int valid = (int)(A.n - base);
                const int copied = (int)((((size_t)valid * sizeof(T)) & ~(size_t)15) / sizeof(T));
                for (int idx = tid; idx < NRL * TILE; idx += cthreads) {
                    const int r = idx / TILE, e = idx - r * TILE;
                    if (e >= copied) st[r * TILE + e] = (e < valid) ? A.rows[r][base + e] : T(0);
                }
                named_bar_sync(1, cthreads);
            }
            if (A.hist) {
                // every consumer thread converts its packs of all 2m pair rows, newest -> oldest
                // (the original row is kept in registers as the successor of the next older pair)
                constexpr int VEC = VecOf<T>::N;
                for (int pk = tid; pk < TILE / VEC; pk += cthreads) {
                    Pack<T> xn = ld_pack<T>(st + TILE, pk), gn = ld_pack<T>(st + 2 * TILE, pk);
                    for (int q = m - 1; q >= 0; --q) {
                        T* sx = st + (3 + 2 * q) * TILE;
                        T* sg = st + (4 + 2 * q) * TILE;
                        const Pack<T> xq = ld_pack<T>(sx, pk), gq = ld_pack<T>(sg, pk);
                        Pack<T> ds, dy;
#pragma unroll
                        for (int u = 0; u < VEC; ++u) {
                            ds.v[u] = xq.v[u] - xn.v[u];
                            dy.v[u] = gq.v[u] - gn.v[u];
                        }
                        st_pack<T>(sx, pk, ds);
                        st_pack<T>(sg, pk, dy);
                        xn = xq;
                        gn = gq;
                    }
                }
                named_bar_sync(2, cthreads);
            } else if (coop) {
                // Several roles: EVERY consumer thread turns its packs of the newest slot into the pair
                // (x0 - x, g0 - g), in the stage and in HBM, and the consumer warps meet at a named
                // barrier.  (With a writer role doing it alone -- two warps -- the warps of the role
                // whose columns hold y_new waited for it 40 % of their time: ncu source page of the
                // Float32 sweep at m = 7, r02e_gram2_f32.)
                constexpr int VEC = VecOf<T>::N;
                for (int pk = tid; pk < TILE / VEC; pk += cthreads) {
                    Pack<T> ps = ld_pack<T>(st + snew, pk), py = ld_pack<T>(st + ynew, pk);
                    const Pack<T> px = ld_pack<T>(st + TILE, pk), pg = ld_pack<T>(st + 2 * TILE, pk);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        ps.v[q] = ps.v[q] - px.v[q];       // s_new = x0 - x (element type)
                        py.v[q] = py.v[q] - pg.v[q];       // y_new = g0 - g
                    }
                    st_pack<T>(st + snew, pk, ps);
                    st_pack<T>(st + ynew, pk, py);
                    const int64_t e0 = base + (int64_t)pk * VEC;
                    if (e0 + VEC <= A.n) {
                        st_pack<T>(A.Snew, e0 / VEC, ps);
                        st_pack<T>(A.Ynew, e0 / VEC, py);
                    } else {
                        for (int q = 0; q < VEC && e0 + q < A.n; ++q) {
                            A.Snew[e0 + q] = ps.v[q];
                            A.Ynew[e0 + q] = py.v[q];
                        }
                    }
                }
                named_bar_sync(2, cthreads);
            } else if (writer && sizeof(T) == 4) {
                // Float32: the same work on 16-byte packs (lane l owns the packs l, l + 32 k, ... of
                // its warp's share: the partition of the accumulation loop below)
                constexpr int VEC = VecOf<T>::N;
                for (int pk = sub * 32 + lane; pk < TILE / VEC; pk += 32 * k) {
                    Pack<T> ps = ld_pack<T>(st + snew, pk), py = ld_pack<T>(st + ynew, pk);
                    const Pack<T> px = ld_pack<T>(st + TILE, pk), pg = ld_pack<T>(st + 2 * TILE, pk);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        ps.v[q] = ps.v[q] - px.v[q];       // s_new = x0 - x (element type)
                        py.v[q] = py.v[q] - pg.v[q];       // y_new = g0 - g
                    }
                    st_pack<T>(st + snew, pk, ps);
                    st_pack<T>(st + ynew, pk, py);
                    const int64_t e0 = base + (int64_t)pk * VEC;
                    if (e0 + VEC <= A.n) {
                        st_pack<T>(A.Snew, e0 / VEC, ps);
                        st_pack<T>(A.Ynew, e0 / VEC, py);
                    } else {
                        for (int q = 0; q < VEC && e0 + q < A.n; ++q) {
                            A.Snew[e0 + q] = ps.v[q];
                            A.Ynew[e0 + q] = py.v[q];
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(pair_a + 8u * s);
            } else if (writer) {
                for (int e = sub * 32 + lane; e < TILE; e += 32 * k) {
                    const T sn = st[snew + e] - st[TILE + e];       // s_new = x0 - x (element type)
                    const T yn = st[ynew + e] - st[2 * TILE + e];   // y_new = g0 - g
                    st[snew + e] = sn;
                    st[ynew + e] = yn;
                    if (base + e < A.n) {
                        A.Snew[base + e] = sn;
                        A.Ynew[base + e] = yn;
                    }
                }
                // a writer warp only reads back the elements it has just updated (same
                // element partition in the loop below): no wait on the other writers
                __syncwarp();
                if (lane == 0) mbar_arrive_a(pair_a + 8u * s);
            } else if (needs_pair) {
                mbar_wait_a(pair_a + 8u * s, ph);
            }
            // (the chain of else-ifs: with HISTORY there is no writer role and nobody waits for it)
            if constexpr (sizeof(T) == 8) {
#pragma unroll 2
                for (int e = sub * 32 + lane; e < TILE; e += 32 * k) {
                    const T* rp = st + rowbase + e;
                    const T* c0 = st + col0 + e;
                    const T* c1 = st + col1 + e;
                    double rv[RT], cv[CT];
#pragma unroll
                    for (int r = 0; r < RT; ++r) rv[r] = rp[r * TILE];
                    const bool fr = !masked || (st[e] != T(0));  // free variable: p[i] != 0 (row 0 = p)
                    cv[0] = fr ? c0[0] : 0.0;
#pragma unroll
                    for (int c = 1; c < CT; ++c) cv[c] = fr ? c1[(c - 1) * 2 * TILE] : 0.0;
#pragma unroll
                    for (int r = 0; r < RT; ++r)
#pragma unroll
                        for (int c = 0; c < CT; ++c) acc[r][c] = fma(rv[r], cv[c], acc[r][c]);
                }
            } else {
                // Float32: widening every operand to Float64 is conversion-bound (the F2F pipe is
                // 1/4 of the FP64 rate), so the products of ONE stage (<= 32 per lane and entry) are
                // accumulated with FFMA and each stage sum is then added to the Float64 accumulators.
                // Every lane works on 16-byte packs (4 elements per LDS.128: a quarter of the
                // shared-memory instructions of an element-per-lane loop).
                constexpr int VEC = VecOf<T>::N;
                float af[RT][CT];
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int c = 0; c < CT; ++c) af[r][c] = 0.0f;
                for (int pk = sub * 32 + lane; pk < TILE / VEC; pk += 32 * k) {
                    const T* rp = st + rowbase;
                    Pack<T> cv[CT];
                    cv[0] = ld_pack<T>(st + col0, pk);
#pragma unroll
                    for (int c = 1; c < CT; ++c) cv[c] = ld_pack<T>(st + col1 + (c - 1) * 2 * TILE, pk);
                    if (masked) {
                        // branch-free mask: all-ones / all-zeros bit pattern ANDed with the columns
                        const Pack<T> p0 = ld_pack<T>(st, pk);       // row 0 = p: free <=> p != 0
#pragma unroll
                        for (int q = 0; q < VEC; ++q) {
                            const int fm = (p0.v[q] != T(0)) ? -1 : 0;
#pragma unroll
                            for (int c = 0; c < CT; ++c)
                                cv[c].v[q] = (T)__int_as_float(__float_as_int((float)cv[c].v[q]) & fm);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const Pack<T> rv = ld_pack<T>(rp + r * TILE, pk);
#pragma unroll
                        for (int q = 0; q < VEC; ++q)
#pragma unroll
                            for (int c = 0; c < CT; ++c)
                                af[r][c] = fmaf((float)rv.v[q], (float)cv[c].v[q], af[r][c]);
                    }
                }
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int c = 0; c < CT; ++c) acc[r][c] += (double)af[r][c];
            }
            // release the stage (the writer's generic-proxy stores into it must be
            // ordered before the next bulk copy)
            if (partial || writer || A.hist || coop) fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty_a + 8u * s);
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
    last_cta_combine(rs.part, rs.res, nslot, gridDim.x, [](int) { return (int)OPKB_RSUM; });
    reduce_done(rs);
}

template <typename T, int BS, int TILE, int MAXT>
static int gram2_launch(opkb_context* ctx, Gram2Args<T>& A, int nroles, int* nslot_out)
{
    const int warps = nroles * A.kshare + 1;
    const int block = warps * 32;
    const int srows = 3 + 2 * A.nb * BS + 2;
    const size_t budget = 200 * 1024;
    const size_t stage_bytes = (size_t)srows * TILE * sizeof(T);
    int nstage = (int)(budget / stage_bytes);
    if (const char* e = getenv("OPKB_GRAM_NSTAGE")) nstage = atoi(e);   // tuning knob
    if (nstage > GRAM2_MAXSTAGE) nstage = GRAM2_MAXSTAGE;
    if (nstage < 2) return opkb_set_error(OPKB_EINVAL, "Gram ring does not fit in shared memory");
    size_t smem = stage_bytes * nstage;
    const size_t red_bytes = (size_t)warps * (2 * BS) * (BS + 1) * sizeof(double);
    if (smem < red_bytes) smem = red_bytes;
    A.tile = TILE;
    A.nstage = nstage;
    auto kern = k_gram2<T, BS, TILE, MAXT>;
    if (block > MAXT) return opkb_set_error(OPKB_EINVAL, "Gram launch: %d threads > %d", block, MAXT);
    OPKB_CUDA(opkb_ensure_smem((const void*)kern, smem));
    int64_t ntiles = (A.n + TILE - 1) / TILE;
    int64_t grid = ctx->sm_count;   // one persistent CTA per SM
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    ReduceScratch rs = opkb_make_scratch(ctx);
    opkb_prof_begin(ctx, OPKB_K_GRAM);
    kern<<<(int)grid, block, smem, ctx->stream>>>(A, rs);
    opkb_prof_end(ctx);
    ctx->launches += 1;
    {
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) {
            cudaFuncAttributes fa;
            memset(&fa, 0, sizeof(fa));
            cudaFuncGetAttributes(&fa, kern);
            return opkb_set_error(OPKB_ECUDA,
                                  "k_gram2 launch failed: %s (grid %d, block %d, smem %zu, regs %d, "
                                  "maxThreadsPerBlock %d, m %d, BS %d, tile %d, nstage %d)",
                                  cudaGetErrorString(err), (int)grid, block, smem, fa.numRegs,
                                  fa.maxThreadsPerBlock, A.m, BS, TILE, nstage);
        }
    }
    *nslot_out = nroles * (2 * BS) * (BS + 1);
    return 0;
}

// row size of a stage: 4 KB rows for up to 10 pairs (nb <= 2), 2 KB beyond;
// OPKB_GRAM_ROWB = 2048 | 3072 | 4096 overrides (tuning knob)
template <typename T, int BS>
static int gram2_dispatch(opkb_context* ctx, Gram2Args<T>& A, int nroles, int* nslot)
{
    constexpr int R4 = 4096 / (int)sizeof(T), R3 = 3072 / (int)sizeof(T), R2 = 2048 / (int)sizeof(T);
    int rowb = (A.nb <= 2) ? 4096 : 2048;
    if (const char* e = getenv("OPKB_GRAM_ROWB")) rowb = atoi(e);
    const bool small_block = (nroles * A.kshare + 1 <= 8);
    if (!small_block && rowb == 4096) return gram2_launch<T, BS, R4, 352>(ctx, A, nroles, nslot);
    if (!small_block) return gram2_launch<T, BS, R2, 352>(ctx, A, nroles, nslot);
    if (rowb == 4096) return gram2_launch<T, BS, R4, 256>(ctx, A, nroles, nslot);
    if (rowb == 3072) return gram2_launch<T, BS, R3, 256>(ctx, A, nroles, nslot);
    return gram2_launch<T, BS, R2, 256>(ctx, A, nroles, nslot);
}

template <typename T>
static int gram2_run(opkb_context* ctx, int64_t n, int m, const void* const* Sp, const void* const* Yp,
                     const void* v0, const void* x, const void* gcur, int masked, double* G, int hist = 0)
{
    Gram2Args<T> A;
    memset(&A, 0, sizeof(A));
    A.hist = hist;
    A.rows[0] = (const T*)v0;
    A.rows[1] = (const T*)x;
    A.rows[2] = (const T*)gcur;
    for (int j = 0; j < m; ++j) {
        A.rows[3 + 2 * j] = (const T*)Sp[j];
        A.rows[4 + 2 * j] = (const T*)Yp[j];
    }
    A.Snew = (T*)Sp[m - 1];
    A.Ynew = (T*)Yp[m - 1];
    A.n = n;
    A.m = m;
    A.masked = masked;
    int nb = (m + 4) / 5;
    if (const char* e = getenv("OPKB_GRAM_NB")) {   // tuning knob: more, smaller blocks (more roles, fewer accumulators each)
        const int v = atoi(e);
        if (v >= nb && v <= 4 && v <= m) nb = v;
    }
    const int bs = (m + nb - 1) / nb;
    const int nroles = nb * (nb + 1) / 2;
    A.nb = nb;
    {
        // measured (profiles/r02g2_gram_coop_ab.txt): Float64 m = 10: 10.33 -> 10.08 ms; Float32 m = 7: 4.16 -> 4.27 ms
        static const int coop_env = getenv("OPKB_GRAM_COOP") ? atoi(getenv("OPKB_GRAM_COOP")) : -1;
        A.coop = (coop_env < 0) ? (sizeof(T) == 8 ? 1 : 0) : coop_env;
    }
    // consumer warps per role: 1 role: 7; 3 roles: 2 each; 6 or 10 roles: 1 each
    A.kshare = (nroles == 1) ? 7 : ((nroles == 3) ? 2 : 1);
    if (const char* e = getenv("OPKB_GRAM_KSHARE")) {   // tuning knob
        const int kk = atoi(e);
        if (kk >= 1 && (nroles * kk + 1) * 32 <= 352) A.kshare = kk;
    }
    int nslot = 0, rc = 0;
    switch (bs) {
    case 1: rc = gram2_dispatch<T, 1>(ctx, A, nroles, &nslot); break;
    case 2: rc = gram2_dispatch<T, 2>(ctx, A, nroles, &nslot); break;
    case 3: rc = gram2_dispatch<T, 3>(ctx, A, nroles, &nslot); break;
    case 4: rc = gram2_dispatch<T, 4>(ctx, A, nroles, &nslot); break;
    default: rc = gram2_dispatch<T, 5>(ctx, A, nroles, &nslot); break;
    }
    if (rc) return rc;
    static thread_local double res[OPKB_NRED_MAX];
    rc = opkb_finish_reduction(ctx, nslot, NULL, res);
    if (rc) return rc;
    gram_scatter(res, m, nb, bs, G);
    return 0;
}
