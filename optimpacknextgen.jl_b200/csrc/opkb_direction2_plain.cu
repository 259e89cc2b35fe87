// opkb_direction2_plain.cu -- instances of k_direction2 without the incremental Gram
// (direction only, or direction + speculative first trial).
#include "opkb_internal.cuh"
#include "opkb_direction2.cuh"

template <typename T>
int dir2_launch_plain(opkb_context* ctx, const Dir2Args<T>& B, int fuse, int mb, size_t smem, int grid)
{
    // (MB, TILE) instances: 4 KB rows with up to 8 / 12 pairs, 2 KB rows with up to 12 / 20
    constexpr int TB = 4096 / (int)sizeof(T), TS = 2048 / (int)sizeof(T);
    void (*kern)(Dir2Args<T>, ReduceScratch) = NULL;
#define DIR2_PICK(F)                                                             \
    do {                                                                         \
        if (B.tile == TB && mb == 8) kern = k_direction2<T, F, 8, TB, 0>;        \
        else if (B.tile == TB && mb == 12) kern = k_direction2<T, F, 12, TB, 0>; \
        else if (B.tile == TS && mb == 12) kern = k_direction2<T, F, 12, TS, 0>; \
        else if (B.tile == TS && mb == 20) kern = k_direction2<T, F, 20, TS, 0>; \
    } while (0)
    if (fuse == OPKB_OBJ_ROSENBROCK) DIR2_PICK(OPKB_OBJ_ROSENBROCK);
    else if (fuse == OPKB_OBJ_QUADRATIC) DIR2_PICK(OPKB_OBJ_QUADRATIC);
    else DIR2_PICK(0);
#undef DIR2_PICK
    return dir2_do_launch<T>(ctx, kern, B, DIR2_NTHREADS(0), smem, grid);
}

template int dir2_launch_plain<double>(opkb_context*, const Dir2Args<double>&, int, int, size_t, int);
template int dir2_launch_plain<float>(opkb_context*, const Dir2Args<float>&, int, int, size_t, int);
