"""Host-side mirror of the reference's LinSolver / LinSolverCreator plugin pair, backed by the device LU.

Mirrors (reference file:line, relative to src/):
  abstract LinSolver, lin_solve(solver, b; tol)            LinSolvers.jl:100,124-137
  FactorizeLinSolver(nep, lambda, umfpack_refinements)     LinSolvers.jl:109-122
  BackslashLinSolver                                        LinSolvers.jl:147-159
  LinSolverCreator, create_linsolver(creator, nep, lambda)  LinSolverCreators.jl:11,24-37
  FactorizeLinSolverCreator(umfpack_refinements, max_factorizations, nep, precomp_values)   LinSolverCreators.jl:62-122
  LinSolverCache / solve(cache, shift, y, add_to_cache)    rk_helper/linsolvercache.jl:7-26
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, ptr
from .neptypes import B200SPMF


class B200LU:
    """Numeric factors of one or several shifts of a B200SPMF (handle wrapper with a finaliser)."""

    def __init__(self, nep: B200SPMF, lams):
        lams = np.atleast_1d(np.asarray(lams, dtype=np.complex128))
        coef = np.ascontiguousarray(np.stack([nep.coefficients(l) for l in lams]))  # nshift x p
        h = C.c_void_p()
        check(lib.nepb_lu_create(nep._h, len(lams), ptr(coef), C.byref(h)))
        self._h = h
        self.nep = nep
        self.lams = lams
        self.nshift = len(lams)

    def status(self, shift=0):
        flags, npert, ratio = C.c_int(), C.c_int(), C.c_double()
        check(lib.nepb_lu_status(self._h, shift, C.byref(flags), C.byref(npert), C.byref(ratio)))
        return {"flags": flags.value, "nperturbed": npert.value, "min_pivot_ratio": ratio.value}

    def solve(self, b, shift=0, refine_steps=0, want_berr=False):
        b = np.asarray(b)
        vec = b.ndim == 1
        B = _lib.as_c128_f(b.reshape(self.nep.n, -1, order="F") if vec else b)
        if B.shape[0] != self.nep.n:
            raise ValueError("right-hand side has %d rows, the NEP has size %d" % (B.shape[0], self.nep.n))
        X = np.empty_like(B, order="F")
        berr = C.c_double(0.0)
        check(lib.nepb_lu_solve(self._h, shift, B.shape[1], ptr(B), B.shape[0], ptr(X), X.shape[0], refine_steps,
                                C.byref(berr) if (want_berr or refine_steps > 0) else None))
        self.last_berr = berr.value
        return X[:, 0].copy() if vec else X

    def close(self):
        if getattr(self, "_h", None):
            lib.nepb_lu_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def symbolic_info(nep: B200SPMF):
    nnzf, fe, nfr, nlv, mf, fl = C.c_int64(), C.c_int64(), C.c_int(), C.c_int(), C.c_int(), C.c_double()
    check(lib.nepb_lu_symbolic_info(nep._h, C.byref(nnzf), C.byref(fe), C.byref(nfr), C.byref(nlv), C.byref(mf), C.byref(fl)))
    return {"nnz_factor": nnzf.value, "front_entries": fe.value, "nfronts": nfr.value, "nlevels": nlv.value,
            "max_front": mf.value, "flops": fl.value}


def symbolic_get(nep: B200SPMF):
    info = symbolic_info(nep)
    perm = np.empty(nep.n, np.int32)
    parent = np.empty(nep.n, np.int32)
    sn_ptr = np.empty(info["nfronts"] + 1, np.int32)
    sn_parent = np.empty(info["nfronts"], np.int32)
    check(lib.nepb_lu_symbolic_get(nep._h, ptr(perm), ptr(parent), ptr(sn_ptr), ptr(sn_parent), None, None))
    return perm, parent, sn_ptr, sn_parent


def symbolic_fronts(nep: B200SPMF):
    """(pivot columns, front order, level) of every front."""
    info = symbolic_info(nep)
    ns = info["nfronts"]
    sn_ptr, rows, level = np.empty(ns + 1, np.int32), np.empty(ns, np.int32), np.empty(ns, np.int32)
    check(lib.nepb_lu_symbolic_get(nep._h, None, None, ptr(sn_ptr), None, ptr(rows), ptr(level)))
    return np.diff(sn_ptr), rows, level


def analyse_pattern(A, ordering=0, relax_leaf=0, max_np=0, fronts=False):
    """Host-only symbolic analysis of the pattern of a scipy sparse matrix (no device needed)."""
    A = A.tocsc()
    n = A.shape[0]
    cp, rv = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    perm, parent, cc, st = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32), np.zeros(8)
    sn_ptr, sn_rows, sn_level = np.zeros(n + 1, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    check(lib.nepb_lu_analyse_pattern(n, ptr(cp), ptr(rv), 0, ordering, relax_leaf, max_np, ptr(perm), ptr(parent), ptr(cc), ptr(st),
                                      ptr(sn_ptr), ptr(sn_rows), ptr(sn_level)))
    keys = ("nnz_factor", "front_entries", "nfronts", "nlevels", "max_front", "max_np", "flops", "solve_rows")
    info = dict(zip(keys, st))
    if fronts:
        ns = int(info["nfronts"])
        info["np"] = np.diff(sn_ptr[:ns + 1])
        info["nf"] = sn_rows[:ns].copy()
        info["level"] = sn_level[:ns].copy()
    return perm, parent, cc, info


def matching(A):
    """Maximum-product row matching and I-matrix scalings of a scipy sparse matrix (host only): row_of_col, dr, dc with
    |dr_i a_ij dc_j| <= 1 and = 1 on the matching (the static-pivoting preprocessing of the device LU)."""
    A = A.tocsc()
    n = A.shape[0]
    cp, rv = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    av = np.ascontiguousarray(np.abs(A.data), dtype=np.float64)
    roc, dr, dc = np.empty(n, np.int32), np.empty(n), np.empty(n)
    check(lib.nepb_lu_matching(n, ptr(cp), ptr(rv), 0, ptr(av), ptr(roc), ptr(dr), ptr(dc)))
    return roc, dr, dc


# ---------------------------------------------------------------------------------------------
# LinSolver objects
# ---------------------------------------------------------------------------------------------
class LinSolver:
    def lin_solve(self, b, tol=0):
        raise NotImplementedError


class B200FactorizeLinSolver(LinSolver):
    """FactorizeLinSolver: factorise once at construction (LinSolvers.jl:113-121), solve many times."""

    def __init__(self, nep: B200SPMF, lam, umfpack_refinements=10):
        self.lu = B200LU(nep, [lam])
        self.refinements = umfpack_refinements
        st = self.lu.status(0)
        # UMFPACK throws SingularException for an exactly singular matrix (`lu` inside `factorize`, LinSolvers.jl:116).  Here:
        # a non-finite pivot, or a zero pivot that survived the static-pivoting fallback (bit 3 = row matching in use).  Pivots
        # that were only lifted (nperturbed > 0) are not fatal by themselves: every solve with such factors is verified through
        # its backward error in the library and fails loudly there (nepb_lu_solve / nepb_lu_solve_block_ex).
        if st["flags"] & 2:
            raise _lib.SingularException(_lib.NEPB_E_SINGULAR, "non-finite pivot while factorising M(%s)" % (lam,))
        if st["flags"] & 1:
            raise _lib.SingularException(_lib.NEPB_E_SINGULAR, "M(%s) is singular to working precision (zero pivot%s)"
                                         % (lam, " after row matching" if st["flags"] & 8 else ""))
        self.status = st

    def lin_solve(self, b, tol=0):
        # `tol` is ignored by direct solvers in the reference as well (LinSolvers.jl:135)
        return self.lu.solve(b, 0, self.refinements)


class B200BackslashLinSolver(LinSolver):
    """BackslashLinSolver: keeps the shift, factorises at every lin_solve (LinSolvers.jl:152-159)."""

    def __init__(self, nep: B200SPMF, lam):
        self.nep, self.lam = nep, lam

    def lin_solve(self, b, tol=0):
        lu = B200LU(self.nep, [self.lam])
        try:
            return lu.solve(b, 0, 2)  # UMFPACK's default of two refinement steps
        finally:
            lu.close()


def gmres(matvec, b, tol, restart=20, maxiter=None, Pl=None, log=False):
    """Restarted GMRES with modified Gram-Schmidt and Givens rotations -- the algorithm behind IterativeSolvers.gmres! 0.9
    (restart = min(20, n), maxiter = n, left preconditioner Pl, stopping when the preconditioned residual has dropped by `tol`
    relative to the initial one).  `matvec` is the LinearMap of LinSolvers.jl:176-178."""
    b = np.asarray(b, dtype=np.complex128)
    n = b.shape[0]
    restart = min(restart, n)
    maxiter = n if maxiter is None else maxiter
    if Pl is None:
        prec = lambda r: r
    elif callable(Pl):
        prec = Pl
    elif hasattr(Pl, "solve"):
        prec = Pl.solve
    else:
        Pm = np.asarray(Pl.todense() if hasattr(Pl, "todense") else Pl)
        diag_only = np.count_nonzero(Pm - np.diag(np.diag(Pm))) == 0
        prec = (lambda r, d=np.diag(Pm): r / d) if diag_only else (lambda r: np.linalg.solve(Pm, r))
    x = np.zeros(n, dtype=np.complex128)
    r = prec(b - matvec(x))
    beta0 = np.linalg.norm(r)
    history = [beta0]
    if beta0 == 0:
        return (x, history) if log else x
    it = 0
    while it < maxiter:
        beta = np.linalg.norm(r)
        V = np.zeros((n, restart + 1), dtype=np.complex128)
        H = np.zeros((restart + 1, restart), dtype=np.complex128)
        cs, sn = np.zeros(restart, dtype=np.complex128), np.zeros(restart, dtype=np.complex128)
        g = np.zeros(restart + 1, dtype=np.complex128)
        g[0] = beta
        V[:, 0] = r / beta
        j_used = 0
        done = False
        for j in range(restart):
            w = prec(matvec(V[:, j]))
            for i in range(j + 1):  # modified Gram-Schmidt
                H[i, j] = np.vdot(V[:, i], w)
                w = w - H[i, j] * V[:, i]
            H[j + 1, j] = np.linalg.norm(w)
            if H[j + 1, j] != 0:
                V[:, j + 1] = w / H[j + 1, j]
            for i in range(j):  # earlier rotations
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -np.conj(sn[i]) * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            a, c = H[j, j], H[j + 1, j]
            d = np.sqrt(abs(a) ** 2 + abs(c) ** 2)
            cs[j], sn[j] = (abs(a) / d, (a / abs(a)) * np.conj(c) / d) if a != 0 else (0.0, 1.0)
            H[j, j] = cs[j] * a + sn[j] * c
            H[j + 1, j] = 0.0
            g[j + 1] = -np.conj(sn[j]) * g[j]
            g[j] = cs[j] * g[j]
            j_used = j + 1
            it += 1
            history.append(abs(g[j + 1]))
            if abs(g[j + 1]) <= tol * beta0 or it >= maxiter:
                done = abs(g[j + 1]) <= tol * beta0
                break
        y = np.linalg.solve(np.triu(H[:j_used, :j_used]), g[:j_used])
        x = x + V[:, :j_used] @ y
        if done:
            break
        r = prec(b - matvec(x))
    return (x, history) if log else x


class GMRESLinSolver(LinSolver):
    """GMRESLinSolver (LinSolvers.jl:171-188): GMRES on the linear map v -> compute_Mlincomb(nep, lambda, v); every product is one
    fused device SpMM.  `lin_solve(b; tol=eps)` takes vector right-hand sides only, as in the reference."""

    def __init__(self, nep, lam, kwargs=None):
        self.nep, self.lam = nep, lam
        self.kwargs = dict(kwargs or {})
        self.A = lambda v: nep.compute_Mlincomb(lam, v)

    def lin_solve(self, b, tol=np.finfo(float).eps):
        b = np.asarray(b)
        if b.ndim != 1:
            raise TypeError("GMRESLinSolver.lin_solve takes a vector right-hand side (LinSolvers.jl:183)")
        kw = dict(self.kwargs)
        tol = kw.pop("tol", tol)  # the creator's keyword wins, as `solver.kwargs...` is splatted last in the reference
        tol = kw.pop("reltol", tol)
        kw.pop("log", None)
        return gmres(self.A, b, tol, restart=kw.pop("restart", 20), maxiter=kw.pop("maxiter", None), Pl=kw.pop("Pl", None))


# ---------------------------------------------------------------------------------------------
# creators
# ---------------------------------------------------------------------------------------------
class LinSolverCreator:
    def create_linsolver(self, nep, lam):
        raise NotImplementedError


class B200LinSolverCreator(LinSolverCreator):
    """FactorizeLinSolverCreator semantics (LinSolverCreators.jl:62-122): optional recycling of up to
    `max_factorizations` factorisations keyed by the shift, optional precomputation at `precomp_values`."""

    def __init__(self, umfpack_refinements=10, max_factorizations=0, nep=None, precomp_values=()):
        self.umfpack_refinements = umfpack_refinements
        self.max_factorizations = max_factorizations
        self.recycled_factorizations = {}
        if len(precomp_values) > 0:
            if nep is None:
                raise ValueError("When you want to precompute factorizations you need to supply the keyword argument `nep`")
            if len(precomp_values) > max_factorizations:
                self.max_factorizations = len(precomp_values)
            for lam in precomp_values:
                self.recycled_factorizations[complex(lam)] = B200FactorizeLinSolver(nep, lam, umfpack_refinements)

    def create_linsolver(self, nep, lam):
        key = complex(lam)
        if key in self.recycled_factorizations:
            return self.recycled_factorizations[key]
        solver = B200FactorizeLinSolver(nep, lam, self.umfpack_refinements)
        if len(self.recycled_factorizations) < self.max_factorizations:
            self.recycled_factorizations[key] = solver
        return solver


class B200BackslashLinSolverCreator(LinSolverCreator):
    def create_linsolver(self, nep, lam):
        return B200BackslashLinSolver(nep, lam)


class GMRESLinSolverCreator(LinSolverCreator):
    """GMRESLinSolverCreator(;kwargs...) (LinSolverCreators.jl:124-145): the keywords are stored and handed to gmres."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs

    def create_linsolver(self, nep, lam):
        return GMRESLinSolver(nep, lam, self.kwargs)


DefaultLinSolverCreator = B200LinSolverCreator


class LinSolverCache:
    """rk_helper/linsolvercache.jl:7-26 (nleigs): factorisations keyed by shift."""

    def __init__(self, nep, creator=None):
        self.nep = nep
        self.creator = creator or B200LinSolverCreator()
        self.solvers = {}

    def solve(self, shift, y, add_to_cache=True):
        key = complex(shift)
        solver = self.solvers.get(key)
        if solver is None:
            solver = self.creator.create_linsolver(self.nep, shift)
            if add_to_cache:
                self.solvers[key] = solver
        return solver.lin_solve(y)
