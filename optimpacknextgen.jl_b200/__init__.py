"""opk-b200: the VMLMB / SPG hot path of OptimPackNextGen.jl on B200 (sm_100a).

Python twin of the Julia host: same names, keywords and error behaviour as the
reference (`vmlmb`, `vmlmb!` -> `vmlmb_`, `spg`, `spg!` -> `spg_`, the three
line searches, `SimpleBounds`), every n-vector resident on the GPU behind the C
ABI of `include/opkb200.h`.  No CPU fallback: importing is cheap, the first
call loads `libopkb200.so` and fails loudly if it is missing.
"""
from ._lib import OpkbError, SO_PATH, EXPORTED_SYMBOLS
from .vectors import (Context, Vector, default_context, comm_unique_id, vcreate, vcopy, vcopy_,
                      vfill_, vdot, vnorm1, vnorm2, vnorminf, vswap_, vscale_, vupdate_, vcombine_, vproduct_,
                      project_variables_, project_direction_, project_gradient_, step_limits,
                      get_free_variables, fastmin, fastmax, fastclamp)
from .linesearches import (LineSearch, ArmijoLineSearch, MoreToraldoLineSearch,
                           MoreThuenteLineSearch, cstep, start_, iterate_, getstep, gettask,
                           getstatus, usederivatives)
from .linesearches import getreason as _ls_getreason
from .objectives import Rosenbrock, Quadratic, Smooth2D, DeviceObjective, splitmix64, uniform
from .vmlmb import (VMLMB, NativeVMLMB, vmlmb, vmlmb_, BoundedSet, EMULATE_BLMVM, apply_lbfgs_,
                    apply_lbfgs_masked_, sufficient_descent, print_iteration, verbose)
from .spg import SPG, NativeSPG, spg, spg_, Info, BoxProjection, default_printer
from .spg import getreason as _spg_getreason
from .spg import (SEARCHING, INFNORM_CONVERGENCE, TWONORM_CONVERGENCE, FUNCTION_CONVERGENCE,
                  TOO_MANY_ITERATIONS, TOO_MANY_EVALUATIONS)
from .conjgrad import (ConjGrad, conjgrad, conjgrad_, DiagonalOperator, Smooth2DOperator, CG_REASON,
                       CG_SEARCHING, CG_GTEST, CG_FTEST, CG_XTEST, CG_TOO_MANY_ITERATIONS,
                       CG_NOT_POSITIVE_DEFINITE)
from . import dist
from .multi import Group, MultiVMLMB


def getreason(obj):
    """`getreason(ls)` (src/linesearches.jl:133) or `getreason(ws::SPG.Info)` (src/spg.jl:166)."""
    if isinstance(obj, Info):
        return _spg_getreason(obj)
    return _ls_getreason(obj)


__all__ = [n for n in dir() if not n.startswith("_")]
