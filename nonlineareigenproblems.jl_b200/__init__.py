"""nepb200: B200-native (sm_100a) hot path for NEP-PACK style nonlinear eigensolvers.

The product is libnepb200.so (csrc/, C ABI in include/nepb200.h) plus the Julia shim in julia/.
This Python package is the in-container mirror of that shim: the same plugin interface
(NEP compute contract, LinSolver / LinSolverCreator, contour integrator, orthogonalisation) bound
with ctypes, so the parity tests read like the reference's own tests.  Importing it requires the
built library; there is no CPU fallback.

The directory name contains a dot, so it is imported under the alias `nepb200` (see /nepb200.py).
"""
from . import _lib  # noqa: F401  (raises ImportError when libnepb200.so is missing)
from ._lib import NepbError, SingularException, device_count, LIB_PATH  # noqa: F401
from .functions import ScalarFunction, Monomial, Exp, PowShift, Callable, ONE, IDENTITY  # noqa: F401
from .neptypes import SPMF_NEP, PEP, DEP, SumNEP, LowRankFactorizedNEP, B200SPMF, Block, B200ProjSPMF, create_proj_NEP  # noqa: F401
from .linsolve import (B200LU, B200FactorizeLinSolver, B200BackslashLinSolver, B200LinSolverCreator, matching,  # noqa: F401
                       B200BackslashLinSolverCreator, DefaultLinSolverCreator, LinSolverCache, LinSolver, LinSolverCreator,
                       symbolic_info, symbolic_get, analyse_pattern, GMRESLinSolver, GMRESLinSolverCreator, gmres)
from .solvers import (contour_block_SS, block_ss_quadrature, block_ss_extract, contour_beyn, ContourIntegrator, beyn_extract, beyn_quadrature, iar, tiar, resinv, infbilanczos, ilan, compute_rf, dgks_host,  # noqa: F401
                      ResidualErrmeasure, StandardSPMFErrmeasure, DefaultErrmeasure, NoConvergenceException,
                      LostOrthogonalityException)
from .dense import (dgks, block_gemm, copy_cols, colnorms, solve_block, mlincomb_block, residual_errors,  # noqa: F401
                    tiar_device, iar_device, iar_chebyshev_device)
from .nleigs import (nleigs, nleigs_backslash, backslash_coefficients, DeviceLinSolverCache, nleigs_lowrank,  # noqa: F401
                     lowrank_backslash, LowRankStructure)
from .deflation import (DeflatedGenericNEP, deflate_eigpair, get_deflated_eigpairs, normalize_schur_pair,  # noqa: F401
                        DeflatedNEPLinSolver, DeflatedNEPLinSolverCreator)
from .wep import (WEP_FD, nep_gallery_WEP, WEPLinSolverCreator, WEPFactorizedLinSolver, WEPBackslashLinSolver,  # noqa: F401
                  WEPGMRESLinSolver, SchurMatVec, construct_WEP_schur_complement, SqrtQuadratic)
from . import rk_helper  # noqa: F401
