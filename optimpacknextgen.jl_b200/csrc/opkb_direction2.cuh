// opkb_direction2.cuh -- P4 "direction", pipelined version (m >= 1 pairs).
//
// Same arithmetic as k_direction (opkb_vmlmb.cu): d = H*v0 from the host-solved
// two-loop coefficients with the reference's rounding sequence, projection,
// g.d, |d|^2, step limits, save (x, g|p) for the next line search.  Other data
// movement: one persistent CTA per SM; a producer warp streams rows of the
// 2m+3..2m+6 input vectors into a shared-memory ring with bulk asynchronous
// copies (cp.async.bulk, mbarrier completion) so that the 25+ concurrent input
// streams reach HBM as multi-KB bursts with >= 100 KB per SM in flight; the
// consumer warps do the element-wise work out of shared memory (128-bit LDS)
// and write d with 128-bit stores.  The last, partial tile is copied with
// 16-byte granularity, completed and zero-padded by the consumers, and its
// results are stored element by element.
#pragma once
#include <type_traits>

#include "opkb_chain.cuh"

#define DIR2_MAXSTAGE 8
#define DIR2_CORR_SHARE 8    // a warp gives up its corrections beyond 8 / 32 changed elements on average ...
#define DIR2_CORR_WARMUP 64  // ... over at least 64 steps (2048 elements) of its stream
// 8 consumer warps + 1 producer warp; the CARRY instances run 7 + 1 (two warps per
// SM sub-partition: up to 255 registers per thread for the extra accumulators)
#define DIR2_NTHREADS(CARRY) ((CARRY) ? 256 : 288)

template <typename T> struct Dir2Args {
    const T* rows[2 * OPKB_MEM_MAX + 8];  // v0 | -, x, g, [lo], [hi], [ysrc], then (y_j, s_j) oldest..newest
    int v0_from_g;                        // VMLMB: v0 = p is recomputed from (x, g, lo, hi), row 0 unused
    T ma[OPKB_MEM_MAX];                   // -alpha_j
    T cc[OPKB_MEM_MAX];                   // alpha_j - beta_j
    unsigned use;
    int m, nrows;
    int r_lo, r_hi, r_ysrc;               // row index or -1 (scalar bound / ysrc == g)
    T gamma, lo_val, hi_val;
    int has_lo, has_hi, masked, bounded;
    T* d;
    T* Snext;
    T* Ynext;
    int64_t n;
    int tile, nstage;
    // FUSE != 0: the first trial point of the next line search is evaluated in the
    // same pass with the step the driver takes after a successful update, stp = 1
    // (src/quasinewton.jl:567, :599, :386-399, :427-448); the host discards it if
    // the direction is rejected or min(1, smax) < 1.
    int r_a, r_c;                         // rows of the quadratic's a and c
    T* xout;
    T* gout;
    T* pout;                              // NULL for L-BFGS
    // CARRY: the pair (s', y') = (x - x', g - g') that this trial point would
    // form (src/quasinewton.jl:512-513) goes to two spare buffers, and every inner
    // product the next two-loop recursion needs that involves it or the new v0
    // is reduced in the same pass (see "incremental Gram" below)
    T* sout;
    T* yout;
    // HISTORY: the pair rows hold the ITERATES (g_j, x_j) at the start of the line search that
    // formed pair j, not the differences: s_j = x_j - x_{j+1}, y_j = g_j - g_{j+1}, where the
    // successor of the newest pair is the current (x, g).  The differences are formed in the
    // stage (each thread for its own pack, while it walks the chain) and nothing is ever written
    // back: with the incremental Gram the iteration moves 2m + 9 passes instead of 2m + 11.
    int hist;
    // CHAIN instances (opkb_chain.cuh): the two-loop coefficients, the `use` mask and the go / no-go
    // of this launch are read from this device structure (left there by the last CTA of the previous
    // launch, or by the host for the first launch of a chain) instead of ma / cc / gamma / use above,
    // and the last CTA leaves those of the next launch there.  NULL otherwise.
    ChainDev* chain;
};

// FUSE: 0 direction only; OPKB_OBJ_ROSENBROCK / OPKB_OBJ_QUADRATIC: + speculative trial.
// MB >= m (unrolled pair loops, coefficients as constant-bank operands), TILE =
// elements per staged row (compile time: every shared read is `LDS [base + imm]`).
//
// CARRY (with FUSE): incremental Gram.  If the host accepts the speculative trial
// point, the next iteration's Gram block (opkb_gram2.cuh) is known without another
// sweep over the pairs: with F' = {p' != 0} the new free set (everything for L-BFGS)
//   * s_j.v0', y_j.v0', s_j.y', y_j.y' on F' for every stored pair j, and the four
//     products of the new pair, are "dense" sums accumulated here (4m+4 numbers)
//     while the rows of S, Y sit in shared memory;
//   * s_a.y_b, y_a.y_b (b >= a) between OLD pairs only change where the free set
//     changes: the previous Gram block plus the signed products over the elements
//     that enter or leave the free set ("corrections", m(m+1) numbers).  Each
//     consumer warp walks its changed elements in a fixed order (ballot), lane r
//     fetches row r of that element and lane l owns the entries l, l+32, ...
// The host assembles the block in opkb_vmlmb_gram (no kernel launch at all).
template <int MB> struct Dir2Carry {
    static constexpr int ND = 4 * MB + 4;           // dense sums
    static constexpr int NENT = MB * (MB + 1);      // corrections: SY(a<=b) then YY(a<=b)
    static constexpr int CPL = (NENT + 31) / 32;    // correction entries per lane
};

// EXACT: m == MB (and every pair takes part is still checked at run time): the
// unrolled pair loops lose their guards.
// CW: consumer warps (+ 1 producer warp).  8 without CARRY; the CARRY instances run 7
// (two warps per SM sub-partition: up to 255 registers for the 4 MB + 4 accumulators)
// or, for few pairs (MB <= 5, where the fixed per-element work makes the kernel
// issue-bound rather than memory-bound), 11 (three per sub-partition, <= 168 registers).
//
// SPLIT (with CARRY; the Float32 instances): the 4 MB dense sums that involve the STORED pairs are
// not accumulated by the thread that owns the elements (4 MB + 4 Float64 accumulators per thread:
// 245 registers, 8 warps per SM, issue-latency bound at half the bytes per element) but in a
// second phase of the same tile by a warp that owns the PAIR: every thread leaves v0' and the
// masked y' of its pack in two rows of the stage that are free by then (row 0: v0 / unused, row 2:
// g, already in registers), the consumer warps meet at a named barrier, and warp w then sweeps
// the items w, w + CW, ... -- item = (pair, one fifth of the tile) -- with FFMA into 4 Float32
// partial sums (8 products each) that are added to its 4 Float64 accumulators of that item, the
// accumulation scheme of the Float32 Gram sweep (opkb_gram2.cuh).  4 ceil(5 MB / CW) + 4
// accumulators per thread instead of 4 MB + 4: 10 consumer warps instead of 7, no Float32 ->
// Float64 conversion of the pair rows.  Fixed item -> warp map and fixed part order: deterministic.
//
// SPEC = 3: L-BFGS without bounds on the iterate history (C1), same idea: see the flags in the kernel.
// SPEC = 1 (the headline instances; 2: the same with stored differences): the grid-uniform flags of the common case -- VMLMB on a box
// given by two arrays (bounded, masked, both bounds present as rows 3 and 4, v0 recomputed from g,
// y = g differences), pairs kept as the iterate history -- are compile-time constants, and the tile
// body is instantiated twice: without the per-element "is this a real element" predicates for the
// whole tiles, with them for the last, partial one.  The generic instances (SPEC = 0) read the
// flags from the arguments.
struct Dir2RtBool { bool value; };
//
// CHAIN = 1 (with CARRY, Float64, MB <= 8): device-resident decisions, see opkb_chain.cuh.  The
// coefficients live in shared memory (read by the prologue from A.chain), a launch whose `go` is 0
// returns at once, and the last CTA runs chain_decide() on the combined scalars before it
// publishes them, followed by its decision record.
template <typename T, int FUSE, int MB, int TILE, int CARRY, int EXACT = 0, int CW = (CARRY ? 7 : 8), int SPLIT = 0,
          int SPEC = 0, int CHAIN = 0>
__global__ void __launch_bounds__((CW + 1) * 32, 1) k_direction2(Dir2Args<T> A, ReduceScratch rs)
{
    constexpr int DIR2_THREADS = (CW + 1) * 32;
    static_assert(!CARRY || (FUSE != 0 && 2 * MB <= 32), "CARRY needs the fused trial and 2*MB <= 32 rows");
    static_assert(!CHAIN || (CARRY && !SPLIT && MB <= OPKB_CHAIN_MB && sizeof(T) == 8),
                  "CHAIN: Float64 carry instances of up to OPKB_CHAIN_MB pairs");
    static_assert(!SPLIT || (CARRY && TILE == CW * 32 * VecOf<T>::N && (TILE / VecOf<T>::N) % (5 * 32) == 0),
                  "SPLIT: one pack per thread and tile, five parts of whole warps of packs");
    constexpr int NSPLIT = 5;                                         // parts of a tile
    constexpr int IPW = SPLIT ? (MB * NSPLIT + CW - 1) / CW : 1;      // items per consumer warp
    constexpr int VEC = VecOf<T>::N;
    constexpr int CWARPS = DIR2_THREADS / 32 - 1;
    constexpr int CTHREADS = CWARPS * 32;
    static_assert(TILE <= CTHREADS * VEC && TILE % (32 * VEC) == 0,
                  "one element pack per thread and tile, whole warps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[DIR2_MAXSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[DIR2_MAXSTAGE];

    const int NSTAGE = A.nstage, NR = A.nrows, m = EXACT ? MB : A.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T* const stage0 = reinterpret_cast<T*>(smem_raw);
    const int stage_elems = NR * TILE;
    constexpr uint32_t rowbytes = (uint32_t)(TILE * sizeof(T));
    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    // CHAIN: coefficients, mask and go / no-go of this launch from the device structure
    __shared__ T s_ma[CHAIN ? MB : 1], s_cc[CHAIN ? MB : 1];
    __shared__ T s_gamma;
    __shared__ unsigned s_use;
    __shared__ int s_go;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (CHAIN != 0) {
        if (tid >= 32 && tid < 32 + MB) {
            const int j = tid - 32;
            s_ma[j] = (T)(-A.chain->alpha[j]);
            s_cc[j] = (T)(A.chain->c[j]);
        }
        if (tid == 64) {
            unsigned u = 0u;
            for (int j = 0; j < m; ++j) {
                const double a = A.chain->alpha[j];
                if (a == a) u |= (1u << j);
            }
            s_use = u;
            s_gamma = (T)A.chain->gamma_dir;
            s_go = A.chain->go;
        }
    }
    __syncthreads();
    if constexpr (CHAIN != 0) {
        if (!s_go) return;     // the step this launch was enqueued for is not taken: nothing is touched
    }
    const unsigned use_mask = CHAIN ? s_use : A.use;
    const bool all_used = (use_mask == (m >= 32 ? 0xffffffffu : ((1u << m) - 1u)));
    auto MA = [&](int j) -> T { if constexpr (CHAIN != 0) return s_ma[j]; else return A.ma[j]; };
    auto CC = [&](int j) -> T { if constexpr (CHAIN != 0) return s_cc[j]; else return A.cc[j]; };
    const T dgamma = CHAIN ? s_gamma : A.gamma;

    // acc: 0 g.d; 1 |d|^2; 2 smin; 3 smax; with FUSE: 4,5 objective sums; 6 |p'|^2; 7 g'.s;
    // 8 g'.d; 9 |x'|^2; 10 |s|^2; 11 number of free variables at x'
    constexpr int NACC = FUSE ? 12 : 4;
    T smin = T(INFINITY), smax = T(0);
    int nfree_i = 0;
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

    // CARRY: cd[4j..4j+3] = s_j.v0', y_j.v0', s_j.y', y_j.y' (y' restricted to F');
    // cd[4MB..] = s'.v0', y'.v0', s'.y', y'.y'; cq[t] = correction entry lane + 32 t
    typedef Dir2Carry<MB> CY;
    constexpr int ND = CARRY ? CY::ND : 1;
    constexpr int CPL = CARRY ? CY::CPL : 1;
    double cd[ND];
    double ci[4 * IPW];   // SPLIT: the 4 sums of each item of this warp
#pragma unroll
    for (int i = 0; i < 4 * IPW; ++i) ci[i] = 0.0;
    double cq[CPL];
    // Corrections are serial work per changed element (~15 warp instructions each).  A trial point
    // that throws a large part of the variables onto / off their bounds -- typically a step that
    // the line search is going to reject anyway -- would make the launch several times longer
    // (measured: 12.4 instead of 4.6 ms at C5).  Every warp therefore keeps the running share of
    // changed elements in its own stream of tiles; once that share exceeds DIR2_CORR_SHARE / 32
    // (after DIR2_CORR_WARMUP steps of 32 elements) it gives up for the rest of the launch and says
    // so in the LAST correction slot (never a real entry: MB (MB + 1) < 32 CPL); the host drops the
    // carry of a launch whose count is not zero (the Gram sweep runs if that point is accepted after
    // all).  Below the threshold the extra work is bounded by ~45 % of the kernel.
    static_assert(!CARRY || CY::NENT < 32 * CY::CPL, "the last correction slot must be free");
    double cskip = 0.0;
    int csteps = 0, cchanged = 0;
    int cl[CPL];   // source lanes (= pair rows) of the two factors of entry lane + 32 t: a | b << 8
#pragma unroll
    for (int i = 0; i < ND; ++i) cd[i] = 0.0;
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
        cq[t] = 0.0;
        cl[t] = 0;
    }
    if (CARRY) {
        // entry idx: first the MB(MB+1)/2 products s_a.y_b (a <= b), then y_a.y_b (a <= b);
        // stage rows of pair j: y_j at prow + 2j, s_j at prow + 2j + 1
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            int idx = lane + 32 * t;
            const int half = MB * (MB + 1) / 2;
            const bool yy = idx >= half;
            if (yy) idx -= half;
            int a = 0;
            while (a < MB && idx >= MB - a) {
                idx -= MB - a;
                ++a;
            }
            const int b = a + idx;
            if (a < MB && b < MB) cl[t] = (yy ? 2 * a : 2 * a + 1) | ((2 * b) << 8);
            else cl[t] = 31 | (31 << 8);   // no such entry (the host ignores the slot)
        }
    }

    const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
    if (warp == CWARPS) {
        const uint32_t stage_a = smem_u32(stage0);
        int s = 0;
        uint32_t ph = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ph ^= (uint32_t)(++s == NSTAGE), s = (s == NSTAGE) ? 0 : s) {
            const int64_t base = tile * TILE;
            uint32_t bytes = rowbytes;
            if (base + TILE > A.n) bytes = (uint32_t)(((A.n - base) * sizeof(T)) & ~(size_t)15);
            if (lane == 0) {
                mbar_wait_a(empty_a + 8u * s, ph ^ 1u);
                mbar_expect_tx_a(full_a + 8u * s, bytes * (uint32_t)(NR - A.v0_from_g));
            }
            __syncwarp();
            const uint32_t dst = stage_a + (uint32_t)(s * stage_elems) * (uint32_t)sizeof(T);
            if (bytes > 0) {
                for (int r = lane + A.v0_from_g; r < NR; r += 32)
                    bulk_g2s_a(dst + (uint32_t)r * rowbytes, A.rows[r] + base, bytes, full_a + 8u * s);
            }
        }
    } else {
        // grid-uniform flags: constants in the SPEC instances, arguments otherwise
        // (SPEC = 3: L-BFGS without bounds on the iterate history -- no projection, no mask, no step
        // limits, v0 = g taken from the g row: row 0 is not even loaded)
        constexpr bool UNB = (SPEC == 3);
        const bool f_bounded = UNB ? false : (SPEC ? true : (A.bounded != 0));
        const bool f_masked = UNB ? false : (SPEC ? true : (A.masked != 0));
        const int f_has_lo = UNB ? 0 : (SPEC ? 1 : A.has_lo), f_has_hi = UNB ? 0 : (SPEC ? 1 : A.has_hi);
        const int r_lo = UNB ? -1 : (SPEC ? 3 : A.r_lo), r_hi = UNB ? -1 : (SPEC ? 4 : A.r_hi);
        const int r_ysrc = SPEC ? -1 : A.r_ysrc;
        const int f_v0g = SPEC ? 1 : A.v0_from_g;       // (row 0 is not in the stage)
        const bool f_hist = (SPEC == 1 || SPEC == 3) ? true : ((SPEC == 2) ? false : (A.hist != 0));   // SPEC 2: stored differences
        const int prow = 3 + (r_lo >= 0) + (r_hi >= 0) + (r_ysrc >= 0);  // first pair row
        int s = 0;
        uint32_t ph = 0;
        T* st = stage0;
        int64_t base = 0;
        auto tile_body = [&](auto partial_c) {
            const bool partial = partial_c.value;
            if (partial) {
                // elements the 16-byte granular bulk copy left out, zero padding
                const int valid = (int)(A.n - base);
                const int copied = (int)((((size_t)valid * sizeof(T)) & ~(size_t)15) / sizeof(T));
                for (int idx = tid; idx < NR * TILE; idx += CTHREADS) {
                    const int r = idx / TILE, e = idx - r * TILE;
                    if (e >= copied && r >= f_v0g)
                        st[r * TILE + e] = (e < valid) ? A.rows[r][base + e] : T(0);
                }
                named_bar_sync(1, CTHREADS);
            }
            // One pack per thread and tile; the warps beyond the tile idle.  The zero
            // padding of a partial tile runs through the arithmetic (it yields d = 0
            // and contributes nothing) so that whole warps stay converged; every
            // reduction of the trial point and every store is guarded per element.
            const int e = tid * VEC;
            if (e < TILE) {
                bool ok[VEC];
#pragma unroll
                for (int q = 0; q < VEC; ++q) ok[q] = !partial || (base + e + q < A.n);
                Pack<T> px = ld_pack<T>(st + TILE, e / VEC);
                Pack<T> pg = ld_pack<T>(st + 2 * TILE, e / VEC);
                Pack<T> pl = (r_lo >= 0) ? ld_pack<T>(st + r_lo * TILE, e / VEC) : splat_pack<T>(A.lo_val);
                Pack<T> ph_ = (r_hi >= 0) ? ld_pack<T>(st + r_hi * TILE, e / VEC) : splat_pack<T>(A.hi_val);
                Pack<T> pys = (r_ysrc >= 0) ? ld_pack<T>(st + r_ysrc * TILE, e / VEC) : pg;
                Pack<T> pv;
                bool flo[VEC], fhi[VEC];   // x may decrease / increase (can_change, src/bounds.jl:638-648)
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    flo[q] = !f_has_lo | (px.v[q] > pl.v[q]);
                    fhi[q] = !f_has_hi | (px.v[q] < ph_.v[q]);
                }
                if constexpr (UNB) {
                    pv = pg;                        // vcopy!(d, g) (src/quasinewton.jl:528)
                } else if (f_v0g) {
                    // p = project_direction!(p, x, lo, hi, -, g) (src/quasinewton.jl:396),
                    // bit-identical to the stored vector, without reading it
#pragma unroll
                    for (int q = 0; q < VEC; ++q) pv.v[q] = projdir_m1(flo[q], fhi[q], pg.v[q]);
                } else {
                    pv = ld_pack<T>(st, e / VEC);
                }
                Pack<T> v = pv;
                bool upd[VEC];
#pragma unroll
                for (int q = 0; q < VEC; ++q) upd[q] = !f_masked || (pv.v[q] != T(0));
                const T* pb = st + prow * TILE + e;
                // apply_lbfgs! :716-745 / :747-779, rounding sequence of the reference: unrolled and
                // branch-free per element.  A pair that does not take part (rho <= 0, :727 / :759 -- it
                // stays in the memory for `mem` iterations) is skipped by a test that is uniform over
                // the grid.  The chain also runs on the masked elements, whose result is then replaced
                // by what the reference gives them, gamma*v0 (they are never updated).
                if (!f_hist) {
#pragma unroll
                    for (int jj = 0; jj < MB; ++jj) {
                        const int j = MB - 1 - jj;
                        if ((EXACT || j < m) && (all_used || ((use_mask >> j) & 1u))) {
                            const Pack<T> y = ld_pack<T>(pb + 2 * j * TILE, 0);
#pragma unroll
                            for (int q = 0; q < VEC; ++q) v.v[q] = v.v[q] + MA(j) * y.v[q];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < VEC; ++q) v.v[q] = dgamma * v.v[q];
#pragma unroll
                    for (int j = 0; j < MB; ++j) {
                        if ((EXACT || j < m) && (all_used || ((use_mask >> j) & 1u))) {
                            const Pack<T> sv = ld_pack<T>(pb + (2 * j + 1) * TILE, 0);
#pragma unroll
                            for (int q = 0; q < VEC; ++q) v.v[q] = v.v[q] + CC(j) * sv.v[q];
                        }
                    }
                } else {
                    // HISTORY: y_j = g_j - g_{j+1} (the same bits as the stored difference of
                    // src/quasinewton.jl:513), formed while the chain walks newest -> oldest; likewise
                    // s_j = x_j - x_{j+1} oldest -> newest.  One subtraction per row and pack, the
                    // predecessor kept in registers.  The dense sums and the corrections below form
                    // them again from the same rows (registers are cheaper than shared-memory stores +
                    // a proxy fence per tile); only the SPLIT instances, whose second phase reads the
                    // rows of other warps, leave the differences in the stage.
                    T* wb = const_cast<T*>(pb);
                    Pack<T> gnext = pg;
#pragma unroll
                    for (int jj = 0; jj < MB; ++jj) {
                        const int j = MB - 1 - jj;
                        if (EXACT || j < m) {
                            const Pack<T> gj = ld_pack<T>(pb + 2 * j * TILE, 0);
                            Pack<T> y;
#pragma unroll
                            for (int q = 0; q < VEC; ++q) y.v[q] = gj.v[q] - gnext.v[q];
                            if (SPLIT) st_pack<T>(wb + 2 * j * TILE, 0, y);
                            gnext = gj;
                            if (all_used || ((use_mask >> j) & 1u)) {
#pragma unroll
                                for (int q = 0; q < VEC; ++q) v.v[q] = v.v[q] + MA(j) * y.v[q];
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < VEC; ++q) v.v[q] = dgamma * v.v[q];
                    Pack<T> xj = ld_pack<T>(pb + TILE, 0);          // x_0 (m >= 1)
#pragma unroll
                    for (int j = 0; j < MB; ++j) {
                        if (EXACT || j < m) {
                            Pack<T> xn = px;                            // successor of the newest pair
                            if (j + 1 < MB && (EXACT || j + 1 < m)) xn = ld_pack<T>(pb + (2 * j + 3) * TILE, 0);
                            Pack<T> sv;
#pragma unroll
                            for (int q = 0; q < VEC; ++q) sv.v[q] = xj.v[q] - xn.v[q];
                            if (SPLIT) st_pack<T>(wb + (2 * j + 1) * TILE, 0, sv);
                            xj = xn;
                            if (all_used || ((use_mask >> j) & 1u)) {
#pragma unroll
                                for (int q = 0; q < VEC; ++q) v.v[q] = v.v[q] + CC(j) * sv.v[q];
                            }
                        }
                    }
                    if (SPLIT) __syncwarp();   // the corrections below read the packs of the other lanes of this warp
                }
#pragma unroll
                for (int q = 0; q < VEC; ++q)
                    if (!upd[q]) v.v[q] = dgamma * pv.v[q];
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    T di = v.v[q];
                    if (f_bounded)  // project_direction!(d, x, lo, hi, -, d) :535
                        di = projdir_m1(flo[q], fhi[q], di);
                    v.v[q] = di;
                    acc[0] = dacc(pg.v[q], di, acc[0]);   // vdot(g, d) :537
                    acc[1] = dacc(di, di, acc[1]);        // vnorm2(d) :629
                    if (f_bounded)                        // step_limits(x, lo, hi, -1, d) :577
                        step_limit_update_fast(px.v[q], pl.v[q], ph_.v[q], f_has_lo, f_has_hi, -di, smin, smax);
                }
                Pack<T> xt, gt, pt;
                if (!FUSE && A.xout) {
                    // user objective: only the trial point x' = project(x - 1*d) (:599, :386) is
                    // formed here; the caller evaluates f, g' there (opkb_vmlmb_trial_begin with
                    // stp = 1 then has nothing left to do)
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        T xi = px.v[q] + T(-1) * v.v[q];
                        if (f_bounded) xi = fastclamp(xi, pl.v[q], ph_.v[q]);
                        xt.v[q] = xi;
                    }
                }
                if (FUSE) {
                    // x' = project(x - 1*d) (:599, :386); f, g' = fg(x') (:391); p' (:396)
                    Pack<T> pa, pc;
                    if (FUSE == OPKB_OBJ_QUADRATIC) {
                        pa = ld_pack<T>(st + A.r_a * TILE, e / VEC);
                        pc = ld_pack<T>(st + A.r_c * TILE, e / VEC);
                    }
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        T xi = px.v[q] + T(-1) * v.v[q];
                        if (f_bounded) xi = fastclamp(xi, pl.v[q], ph_.v[q]);
                        xt.v[q] = xi;
                    }
                    if (FUSE == OPKB_OBJ_ROSENBROCK) {
#pragma unroll
                        for (int q = 0; q < VEC; q += 2) {
                            gt.v[q] = gt.v[q + 1] = T(0);
                            if (ok[q])   // zero padding is not a variable
                                rosenbrock_pair(xt.v[q], xt.v[q + 1], gt.v[q], gt.v[q + 1], acc[4], acc[5]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < VEC; ++q)
                            quadratic_elem(xt.v[q], pa.v[q], pc.v[q], gt.v[q], acc[4]);
                    }
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        const T xi = xt.v[q], gi = gt.v[q];
                        const T pi = f_bounded ? projdir_m1(!f_has_lo | (xi > pl.v[q]), !f_has_hi | (xi < ph_.v[q]), gi)
                                               : gi;
                        pt.v[q] = pi;
                        if (ok[q]) {
                            const T si = xi - px.v[q];              // s = x' - x0 (:427)
                            acc[6] = dacc(pi, pi, acc[6]);          // |p'|^2 (:399)
                            acc[7] = dacc(gi, si, acc[7]);          // g'.s (:432)
                            acc[8] = dacc(gi, v.v[q], acc[8]);      // g'.d (:434)
                            acc[9] = dacc(xi, xi, acc[9]);          // |x'|^2 (:446)
                            acc[10] = dacc(si, si, acc[10]);        // |s|^2 (:448)
                            nfree_i += (pi != T(0)) ? 1 : 0;        // |sel| of the trial point (an integer add)
                        }
                    }
                }
                Pack<T> sn, yn;
                if (CARRY) {
                    // the pair this trial point would form: S[mark'] - x', Y[mark'] - g' (:512-513)
                    T v0n[VEC], ynm[VEC];
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        sn.v[q] = px.v[q] - xt.v[q];
                        yn.v[q] = pg.v[q] - gt.v[q];
                        const bool frn = !f_masked || (pt.v[q] != T(0));
                        v0n[q] = ok[q] ? pt.v[q] : T(0);                 // v0' = p' (g' for L-BFGS)
                        ynm[q] = (ok[q] && frn) ? yn.v[q] : T(0);        // y' on the new free set
                    }
                    // exact products accumulated in Float64 whatever T (as every reduction
                    // of the library; for Float32 each operand is widened once per element)
                    double v0d[VEC], ynd[VEC];
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        v0d[q] = (double)v0n[q];
                        ynd[q] = (double)ynm[q];
                    }
                    if constexpr (!SPLIT) {
                        Pack<T> hg = ld_pack<T>(pb, 0), hx = ld_pack<T>(pb + TILE, 0);   // (HISTORY: g_0, x_0)
#pragma unroll
                        for (int j = 0; j < MB; ++j) {
                            if (EXACT || j < m) {
                                Pack<T> y = hg, sv = hx;
                                if (f_hist) {
                                    Pack<T> gn = pg, xn = px;          // successor of the newest pair
                                    if (j + 1 < MB && (EXACT || j + 1 < m)) {
                                        gn = ld_pack<T>(pb + (2 * j + 2) * TILE, 0);
                                        xn = ld_pack<T>(pb + (2 * j + 3) * TILE, 0);
                                    }
#pragma unroll
                                    for (int q = 0; q < VEC; ++q) {
                                        y.v[q] = hg.v[q] - gn.v[q];
                                        sv.v[q] = hx.v[q] - xn.v[q];
                                    }
                                    hg = gn;
                                    hx = xn;
                                } else {
                                    y = ld_pack<T>(pb + 2 * j * TILE, 0);
                                    sv = ld_pack<T>(pb + (2 * j + 1) * TILE, 0);
                                }
#pragma unroll
                                for (int q = 0; q < VEC; ++q) {
                                    const double sd = (double)sv.v[q], yd = (double)y.v[q];
                                    cd[4 * j + 0] = fma(sd, v0d[q], cd[4 * j + 0]);
                                    cd[4 * j + 1] = fma(yd, v0d[q], cd[4 * j + 1]);
                                    cd[4 * j + 2] = fma(sd, ynd[q], cd[4 * j + 2]);
                                    cd[4 * j + 3] = fma(yd, ynd[q], cd[4 * j + 3]);
                                }
                            }
                        }
                    } else {
                        // v0' and the masked y' of this pack for the second phase: rows 0 and 2 of
                        // the stage (this thread has read its packs of them; nobody else reads them)
                        Pack<T> pv0, pym;
#pragma unroll
                        for (int q = 0; q < VEC; ++q) {
                            pv0.v[q] = v0n[q];
                            pym.v[q] = ynm[q];
                        }
                        st_pack<T>(st, e / VEC, pv0);
                        st_pack<T>(st + 2 * TILE, e / VEC, pym);
                    }
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        const double sd = (double)sn.v[q], yd = (double)yn.v[q];
                        cd[4 * MB + 0] = fma(sd, v0d[q], cd[4 * MB + 0]);
                        cd[4 * MB + 1] = fma(yd, v0d[q], cd[4 * MB + 1]);
                        cd[4 * MB + 2] = fma(sd, ynd[q], cd[4 * MB + 2]);
                        cd[4 * MB + 3] = fma(yd, ynd[q], cd[4 * MB + 3]);
                    }
                    if (f_masked) {
                        // elements that enter (+) or leave (-) the free set: signed products
                        // between the stored pairs, one element at a time, whole warp
#pragma unroll
                        for (int q = 0; q < VEC; ++q) {
                            const bool nf = (pt.v[q] != T(0));
                            const bool chg = ok[q] && (nf != (pv.v[q] != T(0)));
                            unsigned bal = __ballot_sync(0xffffffffu, chg);
                            const unsigned nfb = __ballot_sync(0xffffffffu, nf);
                            csteps += 1;
                            cchanged += __popc(bal);
                            if (cskip == 0.0 && csteps >= DIR2_CORR_WARMUP && cchanged > DIR2_CORR_SHARE * csteps)
                                cskip = 1.0;
                            if (cskip != 0.0) bal = 0u;
                            while (bal) {
                                const int src = __ffs(bal) - 1;
                                bal &= bal - 1;
                                const int ec = (e - lane * VEC) + src * VEC + q;
                                double val = 0.0;
                                if (lane < 2 * m) {
                                    T rv = st[(prow + lane) * TILE + ec];
                                    if (f_hist && !SPLIT) {
                                        // row `lane` is g_j (even) or x_j (odd), j = lane / 2: minus its successor
                                        const T nx = (lane + 2 < 2 * m) ? st[(prow + lane + 2) * TILE + ec]
                                                                        : st[((lane & 1) ? 1 : 2) * TILE + ec];
                                        rv = rv - nx;
                                    }
                                    val = (double)rv;
                                }
                                const double sg = ((nfb >> src) & 1u) ? 1.0 : -1.0;
#pragma unroll
                                for (int t = 0; t < CPL; ++t) {
                                    const double fa = __shfl_sync(0xffffffffu, val, cl[t] & 0xff);
                                    const double fb = __shfl_sync(0xffffffffu, val, cl[t] >> 8);
                                    cq[t] = fma(sg * fa, fb, cq[t]);
                                }
                            }
                        }
                    }
                }
                if (!partial || base + e + VEC <= A.n) {
                    const int64_t iv = (base + e) / VEC;
                    st_pack<T>(A.d, iv, v);
                    if (FUSE) {
                        st_pack<T>(A.xout, iv, xt);
                        st_pack<T>(A.gout, iv, gt);
                        if (A.pout) st_pack<T>(A.pout, iv, pt);
                    } else if (A.xout) {
                        st_pack<T>(A.xout, iv, xt);
                    }
                    if (CARRY && A.sout) {              // (HISTORY: the pair is never materialised)
                        st_pack<T>(A.sout, iv, sn);
                        st_pack<T>(A.yout, iv, yn);
                    }
                    if (A.Snext) {                      // (the workspace path aliases instead)
                        st_pack<T>(A.Snext, iv, px);    // vcopy!(S[mark], x) :562
                        st_pack<T>(A.Ynext, iv, pys);   // vcopy!(Y[mark], g|p) :563
                    }
                } else {
                    for (int q = 0; q < VEC && base + e + q < A.n; ++q) {
                        A.d[base + e + q] = v.v[q];
                        if (FUSE) {
                            A.xout[base + e + q] = xt.v[q];
                            A.gout[base + e + q] = gt.v[q];
                            if (A.pout) A.pout[base + e + q] = pt.v[q];
                        } else if (A.xout) {
                            A.xout[base + e + q] = xt.v[q];
                        }
                        if (CARRY && A.sout) {
                            A.sout[base + e + q] = sn.v[q];
                            A.yout[base + e + q] = yn.v[q];
                        }
                        if (A.Snext) {
                            A.Snext[base + e + q] = px.v[q];
                            A.Ynext[base + e + q] = pys.v[q];
                        }
                    }
                }
            }
            if constexpr (SPLIT) {
                named_bar_sync(2, CTHREADS);      // every pack's v0', y' are in the stage
                constexpr int PP = TILE / VEC / NSPLIT;   // packs per part
#pragma unroll
                for (int t = 0; t < IPW; ++t) {
                    const int item = warp + CWARPS * t;
                    const int pj = item / NSPLIT, part = item - pj * NSPLIT;
                    if (item < MB * NSPLIT && pj < m) {
                        const T* ys = st + (prow + 2 * pj) * TILE;    // y_j, then s_j
                        float af0 = 0.0f, af1 = 0.0f, af2 = 0.0f, af3 = 0.0f;
#pragma unroll
                        for (int u = 0; u < PP / 32; ++u) {
                            const int pk = part * PP + u * 32 + lane;
                            const Pack<T> y = ld_pack<T>(ys, pk), sv = ld_pack<T>(ys + TILE, pk);
                            const Pack<T> w0 = ld_pack<T>(st, pk), wy = ld_pack<T>(st + 2 * TILE, pk);
#pragma unroll
                            for (int q = 0; q < VEC; ++q) {
                                af0 = fmaf((float)sv.v[q], (float)w0.v[q], af0);   // s_j . v0'
                                af1 = fmaf((float)y.v[q], (float)w0.v[q], af1);    // y_j . v0'
                                af2 = fmaf((float)sv.v[q], (float)wy.v[q], af2);   // s_j . y'
                                af3 = fmaf((float)y.v[q], (float)wy.v[q], af3);    // y_j . y'
                            }
                        }
                        ci[4 * t + 0] += (double)af0;
                        ci[4 * t + 1] += (double)af1;
                        ci[4 * t + 2] += (double)af2;
                        ci[4 * t + 3] += (double)af3;
                    }
                }
            }
            // (SPLIT: rows 0 and 2, HISTORY: the pair rows were written through the generic proxy)
            if (partial || SPLIT) fence_proxy_async();
        };
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ph ^= (uint32_t)(++s == NSTAGE), s = (s == NSTAGE) ? 0 : s) {
            st = stage0 + s * stage_elems;
            base = tile * TILE;
            mbar_wait_a(full_a + 8u * s, ph);
            if constexpr (SPEC != 0) {
                if (base + TILE > A.n) tile_body(std::true_type{});
                else tile_body(std::false_type{});
            } else {
                tile_body(Dir2RtBool{base + TILE > A.n});
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty_a + 8u * s);
        }
    }
    acc[2] = (double)smin;
    acc[3] = (double)smax;
    if (FUSE) acc[NACC - 1] = (double)nfree_i;
    __syncthreads();
    if constexpr (!CARRY) {
        block_reduce_publish<NACC, 8u, 4u>(acc, rs);  // slot 2: min, slot 3: max
    } else {
        // NACC + ND sums over all threads (slot 2: min, slot 3: max) and 32*CPL per-lane
        // entries summed over the warps; the drained ring is the scratch.  Fixed order
        // everywhere: butterfly, warps 0.., CTAs 0.. (last CTA by ticket).
        constexpr int NT = NACC + ND;
        constexpr int NTOT = NT + 32 * CPL;
        constexpr int NWARPS = DIR2_THREADS / 32;
        double* red = reinterpret_cast<double*>(smem_raw);   // [NWARPS][NTOT] (+ SPLIT: [MB * NSPLIT][4])
        double* itm = red + NWARPS * NTOT;
        __shared__ unsigned int s_last2;
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            if (SPLIT && k >= NACC && k < NACC + 4 * MB) continue;   // reduced per item below
            const int kind = (k == 2) ? OPKB_RMIN : ((k == 3) ? OPKB_RMAX : OPKB_RSUM);
            double a = (k < NACC) ? acc[k < NACC ? k : 0] : cd[k >= NACC ? k - NACC : 0];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double b = __shfl_xor_sync(0xffffffffu, a, off);
                a = rcombine_dyn(kind, a, b);
            }
            if (lane == 0) red[warp * NTOT + k] = a;
        }
        if constexpr (SPLIT) {
#pragma unroll
            for (int t = 0; t < IPW; ++t) {
                const int item = warp + CWARPS * t;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    double a = ci[4 * t + kk];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
                    if (lane == 0 && warp < CWARPS && item < MB * NSPLIT) itm[item * 4 + kk] = a;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < CPL; ++t) red[warp * NTOT + NT + 32 * t + lane] = cq[t];
        if (lane == 31) red[warp * NTOT + NT + 32 * (CPL - 1) + 31] = cskip;   // (warp-uniform count)
        __syncthreads();
        for (int k = tid; k < NTOT; k += DIR2_THREADS) {
            const int kind = (k == 2) ? OPKB_RMIN : ((k == 3) ? OPKB_RMAX : OPKB_RSUM);
            double a;
            if (SPLIT && k >= NACC && k < NACC + 4 * MB) {
                // stored pair j: its NSPLIT parts in part order (pairs the launch does not hold: 0)
                const int j = (k - NACC) >> 2, kk = (k - NACC) & 3;
                a = 0.0;
                if (j < m)
                    for (int part = 0; part < NSPLIT; ++part) a += itm[(j * NSPLIT + part) * 4 + kk];
            } else {
                a = red[k];
                for (int w = 1; w < NWARPS; ++w) a = rcombine_dyn(kind, a, red[w * NTOT + k]);
            }
            rs.part[(size_t)blockIdx.x * NTOT + k] = a;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int t = atomicAdd(rs.ticket, 1u);
            s_last2 = (t == gridDim.x - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (!s_last2) return;
        __threadfence();
        if constexpr (CHAIN != 0) {
            // combined scalars into shared memory (the drained scratch), from there to the host, and
            // the decision about the next launch (one thread) before the completion flag
            double* comb = red;
            last_cta_combine(rs.part, comb, NTOT, gridDim.x,
                             [](int k) { return (k == 2) ? OPKB_RMIN : ((k == 3) ? OPKB_RMAX : OPKB_RSUM); });
            __syncthreads();
            for (int k = tid; k < NTOT; k += DIR2_THREADS) rs.res[k] = comb[k];
            chain_decide_block<MB>(A.chain, comb, rs.res + NTOT, comb + ((NTOT + 1) & ~1), tid, DIR2_THREADS);
        } else {
            last_cta_combine(rs.part, rs.res, NTOT, gridDim.x,
                             [](int k) { return (k == 2) ? OPKB_RMIN : ((k == 3) ? OPKB_RMAX : OPKB_RSUM); });
        }
        reduce_done(rs);
    }
}

// ---------------------------------------------------------------------------
// Launchers (instantiated in opkb_direction2_plain.cu / opkb_direction2_carry.cu so
// that the translation units build in parallel).  `fuse`: 0 | OPKB_OBJ_ROSENBROCK |
// OPKB_OBJ_QUADRATIC; (mb, tile) must be one of the instances listed there.
#define DIR2_CARRY_ROWB 3584    // bytes per staged row of the CARRY instances (7 consumer warps x 512 B)
#define DIR2_CARRY_ROWB5 5632   // same for the MB = 5 instances (11 consumer warps)
#define DIR2_SPLIT_ROWB 5120    // the SPLIT instances (Float32, MB = 8: 10 consumer warps x 512 B)
#define DIR2_SPLIT_CW 10
template <typename T>
int dir2_launch_plain(opkb_context* ctx, const Dir2Args<T>& B, int fuse, int mb, size_t smem, int grid);
template <typename T>
int dir2_launch_carry(opkb_context* ctx, const Dir2Args<T>& B, int fuse, int mb, size_t smem, int grid,
                      int res_offset = 0);

template <typename T>
static int dir2_do_launch(opkb_context* ctx, void (*kern)(Dir2Args<T>, ReduceScratch), const Dir2Args<T>& B,
                          int threads, size_t smem, int grid, int res_offset = 0)
{
    if (!kern) return opkb_set_error(OPKB_EINVAL, "no direction kernel instance for this shape");
    OPKB_CUDA(opkb_ensure_smem((const void*)kern, smem));
    ReduceScratch rs = opkb_make_scratch(ctx, res_offset);
    opkb_prof_begin(ctx, OPKB_K_DIRECTION);
    kern<<<grid, threads, smem, ctx->stream>>>(B, rs);
    opkb_prof_end(ctx);
    ctx->launches += 1;
    OPKB_CUDA(cudaGetLastError());
    return 0;
}
