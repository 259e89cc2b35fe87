// opkb_direction2_carry.cu -- instances of k_direction2 with the incremental Gram
// (direction + speculative first trial + the next Gram block, see opkb_direction2.cuh).
#include <stdlib.h>

#include "opkb_internal.cuh"
#include "opkb_direction2.cuh"

template <typename T>
int dir2_launch_carry(opkb_context* ctx, const Dir2Args<T>& B, int fuse, int mb, size_t smem, int grid, int res_offset)
{
    constexpr int TC = DIR2_CARRY_ROWB / (int)sizeof(T), TC5 = DIR2_CARRY_ROWB5 / (int)sizeof(T);
    constexpr int TS = DIR2_SPLIT_ROWB / (int)sizeof(T);
    void (*kern)(Dir2Args<T>, ReduceScratch) = NULL;
    if constexpr (sizeof(T) == 4) {
        // Float32, up to 8 pairs: the dense sums of the stored pairs in a second phase by the warp
        // that owns the pair (SPLIT, opkb_direction2.cuh): 10 consumer warps on 5 KB rows
        if (mb == 8 && B.tile == TS) {
            // SPEC = 2: VMLMB on a box given by two arrays, stored differences: flags as constants,
            // whole tiles without the partial-tile predicates
            static const int spec_on32 = getenv("OPKB_DIR_SPEC") ? atoi(getenv("OPKB_DIR_SPEC")) : 1;
            const bool box = spec_on32 && B.bounded && B.masked && B.has_lo && B.has_hi && B.r_lo == 3 &&
                             B.r_hi == 4 && B.r_ysrc < 0 && B.v0_from_g;
            const int spec = !box ? 0 : (B.hist ? 1 : 2);
            if (fuse == OPKB_OBJ_ROSENBROCK)
                kern = spec == 1 ? k_direction2<T, OPKB_OBJ_ROSENBROCK, 8, TS, 1, 0, DIR2_SPLIT_CW, 1, 1>
                     : spec == 2 ? k_direction2<T, OPKB_OBJ_ROSENBROCK, 8, TS, 1, 0, DIR2_SPLIT_CW, 1, 2>
                                 : k_direction2<T, OPKB_OBJ_ROSENBROCK, 8, TS, 1, 0, DIR2_SPLIT_CW, 1, 0>;
            else if (fuse == OPKB_OBJ_QUADRATIC)
                kern = spec == 1 ? k_direction2<T, OPKB_OBJ_QUADRATIC, 8, TS, 1, 0, DIR2_SPLIT_CW, 1, 1>
                     : spec == 2 ? k_direction2<T, OPKB_OBJ_QUADRATIC, 8, TS, 1, 0, DIR2_SPLIT_CW, 1, 2>
                                 : k_direction2<T, OPKB_OBJ_QUADRATIC, 8, TS, 1, 0, DIR2_SPLIT_CW, 1, 0>;
            return dir2_do_launch<T>(ctx, kern, B, (DIR2_SPLIT_CW + 1) * 32, smem, grid);
        }
    }
    if (B.tile != ((mb == 5) ? TC5 : TC))
        return opkb_set_error(OPKB_EINVAL, "carry instances use %d / %d-byte rows", DIR2_CARRY_ROWB5, DIR2_CARRY_ROWB);
    // (MB, EXACT) instances: m < 5: (5, 0); m = 5: (5, 1); m <= 8: (8, 0); m = 10: (10, 1); else (12, 0)
#define DIR2_PICK(F, SPEC)                                                                  \
    do {                                                                                    \
        if (mb == 5 && B.m == 5) kern = k_direction2<T, F, 5, TC5, 1, 1, 11, 0, SPEC>;      \
        else if (mb == 5) kern = k_direction2<T, F, 5, TC5, 1, 0, 11, 0, SPEC>;             \
        else if (mb == 8) kern = k_direction2<T, F, 8, TC, 1, 0, 7, 0, SPEC>;               \
        else if (mb == 10 && B.m == 10) kern = k_direction2<T, F, 10, TC, 1, 1, 7, 0, SPEC>; \
        else if (mb == 12) kern = k_direction2<T, F, 12, TC, 1, 0, 7, 0, SPEC>;             \
    } while (0)
    // CHAIN instances (device-resident decisions, opkb_chain.cuh): Float64, up to 8 pairs
#define DIR2_PICK_CHAIN(F, SPEC)                                                            \
    do {                                                                                    \
        if (mb == 5 && B.m == 5) kern = k_direction2<T, F, 5, TC5, 1, 1, 11, 0, SPEC, 1>;   \
        else if (mb == 5) kern = k_direction2<T, F, 5, TC5, 1, 0, 11, 0, SPEC, 1>;          \
        else if (mb == 8) kern = k_direction2<T, F, 8, TC, 1, 0, 7, 0, SPEC, 1>;            \
    } while (0)
    // SPEC = 1 (Float64): VMLMB on a box given by two arrays with the iterate history -- the flags are
    // compile-time constants and the whole tiles run a body without the partial-tile predicates
    static const int spec_on = getenv("OPKB_DIR_SPEC") ? atoi(getenv("OPKB_DIR_SPEC")) : 1;
    const bool spec = sizeof(T) == 8 && spec_on && B.bounded && B.masked && B.has_lo && B.has_hi &&
                      B.r_lo == 3 && B.r_hi == 4 && B.r_ysrc < 0 && B.v0_from_g;
    if constexpr (sizeof(T) == 8) {
        // SPEC = 3: L-BFGS without bounds, iterate history, v0 = g (the g row serves as v0: row 0 is not loaded)
        if (!B.chain && spec_on && !B.bounded && !B.masked && B.r_ysrc < 0 && !B.v0_from_g && B.hist &&
            B.rows[0] == B.rows[2]) {
            if (fuse == OPKB_OBJ_ROSENBROCK) DIR2_PICK(OPKB_OBJ_ROSENBROCK, 3);
            else if (fuse == OPKB_OBJ_QUADRATIC) DIR2_PICK(OPKB_OBJ_QUADRATIC, 3);
            if (kern) {
                Dir2Args<T> U = B;
                U.v0_from_g = 1;
                return dir2_do_launch<T>(ctx, kern, U, (mb == 5) ? 384 : DIR2_NTHREADS(1), smem, grid);
            }
        }
        if (B.chain) {
            const bool unb = spec_on && !B.bounded && !B.masked && B.r_ysrc < 0 && !B.v0_from_g && B.hist &&
                             B.rows[0] == B.rows[2];
            if (unb) {      // L-BFGS without bounds: SPEC = 3 (row 0 is not loaded)
                if (fuse == OPKB_OBJ_ROSENBROCK) DIR2_PICK_CHAIN(OPKB_OBJ_ROSENBROCK, 3);
                else if (fuse == OPKB_OBJ_QUADRATIC) DIR2_PICK_CHAIN(OPKB_OBJ_QUADRATIC, 3);
                if (kern) {
                    Dir2Args<T> U = B;
                    U.v0_from_g = 1;
                    return dir2_do_launch<T>(ctx, kern, U, (mb == 5) ? 384 : DIR2_NTHREADS(1), smem, grid, res_offset);
                }
            }
            const int sp = !spec ? 0 : (B.hist ? 1 : 2);
            if (fuse == OPKB_OBJ_ROSENBROCK) {
                if (sp == 1) DIR2_PICK_CHAIN(OPKB_OBJ_ROSENBROCK, 1);
                else if (sp == 2) DIR2_PICK_CHAIN(OPKB_OBJ_ROSENBROCK, 2);
                else DIR2_PICK_CHAIN(OPKB_OBJ_ROSENBROCK, 0);
            } else if (fuse == OPKB_OBJ_QUADRATIC) {
                if (sp == 1) DIR2_PICK_CHAIN(OPKB_OBJ_QUADRATIC, 1);
                else if (sp == 2) DIR2_PICK_CHAIN(OPKB_OBJ_QUADRATIC, 2);
                else DIR2_PICK_CHAIN(OPKB_OBJ_QUADRATIC, 0);
            }
            if (!kern) return opkb_set_error(OPKB_EINVAL, "no chained instance of the direction kernel for this shape");
            return dir2_do_launch<T>(ctx, kern, B, (mb == 5) ? 384 : DIR2_NTHREADS(1), smem, grid, res_offset);
        }
        if (spec && B.hist) {
            if (fuse == OPKB_OBJ_ROSENBROCK) DIR2_PICK(OPKB_OBJ_ROSENBROCK, 1);
            else if (fuse == OPKB_OBJ_QUADRATIC) DIR2_PICK(OPKB_OBJ_QUADRATIC, 1);
        } else if (spec) {      // stored differences (OPKB_HISTORY=0)
            if (fuse == OPKB_OBJ_ROSENBROCK) DIR2_PICK(OPKB_OBJ_ROSENBROCK, 2);
            else if (fuse == OPKB_OBJ_QUADRATIC) DIR2_PICK(OPKB_OBJ_QUADRATIC, 2);
        }
    }
    if (B.chain) return opkb_set_error(OPKB_EINVAL, "chained launches are Float64 only");
    if (!kern) {
        if (fuse == OPKB_OBJ_ROSENBROCK) DIR2_PICK(OPKB_OBJ_ROSENBROCK, 0);
        else if (fuse == OPKB_OBJ_QUADRATIC) DIR2_PICK(OPKB_OBJ_QUADRATIC, 0);
    }
#undef DIR2_PICK
#undef DIR2_PICK_CHAIN
    return dir2_do_launch<T>(ctx, kern, B, (mb == 5) ? 384 : DIR2_NTHREADS(1), smem, grid);
}

template int dir2_launch_carry<double>(opkb_context*, const Dir2Args<double>&, int, int, size_t, int, int);
template int dir2_launch_carry<float>(opkb_context*, const Dir2Args<float>&, int, int, size_t, int, int);
