"""The waveguide eigenvalue problem in its native format on the device: host-side mirror of the
reference's `GalleryWaveguide` module for the FD discretisation,
    nep_gallery(WEP; nx, nz, benchmark_problem, neptype, delta)      src/gallery_extra/GalleryWaveguide.jl:60-92
    WEP_FD, compute_Mlincomb(::WEP_FD, ...)                          src/gallery_extra/waveguide/Waveguide.jl:203-240, 324-379
    SchurMatVec, WEP{Backslash,Factorized,GMRES}LinSolver, lin_solve Waveguide.jl:393-567
    WEPLinSolverCreator                                              Waveguide.jl:491-521
over libnepb200's `nepb_wep_*` entry points (csrc/wep.cu).  As everywhere in this package scalar functions stay on the host
(here: the Gegenbauer recurrence `sqrt_derivative`, Waveguide.jl:574-616, vectorised over the 2 nz boundary modes) and
everything that touches an n-vector runs on the device: the Sylvester-form interior as one stencil pass, the boundary
operator R diag(.) R^-1 as direct odd-length transforms, the Schur-complement solve on the device multifrontal LU."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import lib, check, ptr
from .functions import ScalarFunction, ONE, IDENTITY, Monomial
from .neptypes import SPMF_NEP, B200SPMF, Block
from .linsolve import LinSolver, LinSolverCreator, B200FactorizeLinSolver, gmres


# ---- scalar helpers ------------------------------------------------------------------------------------------------------
def sqrt_pos_imag(a):
    """Square root on the branch with non-negative imaginary part (Waveguide.jl:130-144); elementwise."""
    a = np.asarray(a, dtype=np.complex128)
    s = np.sign(a.imag)
    return np.where(s == 0, 1.0, s) * np.sqrt(a)


def sqrt_derivative(a, b, c, d=0, x=0.0):
    """Derivatives 0..d of sqrt(a z^2 + b z + c) at z = x (Waveguide.jl:574-616) for arrays b, c: returns (len(b), d + 1)."""
    b = np.atleast_1d(np.asarray(b, dtype=np.complex128))
    c = np.atleast_1d(np.asarray(c, dtype=np.complex128))
    if d < 0:
        raise ValueError("Cannot take negative derivative. d = %d" % d)
    aa, bb, cc = a, b + 2 * a * x, c + a * x ** 2 + b * x
    out = np.zeros((len(b), d + 1), dtype=np.complex128)
    yi = sqrt_pos_imag(cc)
    out[:, 0] = yi
    if d == 0:
        return out
    yip1 = bb / (2 * yi)
    out[:, 1] = yip1
    fact = 1.0
    for i in range(2, d + 1):
        m = i - 2
        yip2 = -(2 * aa * (m - 1) * yi + bb * (1 + 2 * m) * yip1) / (2 * cc * (2 + m))
        fact *= i
        yi, yip1 = yip1, yip2
        out[:, i] = yip2 * fact
    return out


class SqrtQuadratic(ScalarFunction):
    """lambda -> 1im*sqrt(lambda^2 + b lambda + c) + d0 on the branch of sqrt_pos_imag: the boundary functions S(lambda, j) of
    the SPMF format (generate_S_function, Waveguide.jl:70-113).  Scalar arguments only (Taylor coefficients through the
    Gegenbauer recurrence); the matrix version of the reference is a Schur square root and is not needed by the device path."""

    def __init__(self, b, c, d0):
        self.b, self.c, self.d0 = complex(b), complex(c), float(d0)

    def __call__(self, S):
        if isinstance(S, np.ndarray) and S.ndim == 2:
            raise NotImplementedError("SqrtQuadratic takes scalars; use the WEP format for matrix arguments")
        lam = complex(S)
        return complex(1j * sqrt_pos_imag(lam * lam + self.b * lam + self.c) + self.d0)

    def taylor(self, lam, m):
        der = 1j * sqrt_derivative(1.0, [self.b], [self.c], m - 1, complex(lam))[0]
        der[0] += self.d0
        return der / np.array([math.factorial(j) for j in range(m)], dtype=np.float64)


# ---- discretisation (waveguide_FD.jl) -----------------------------------------------------------------------------------
def generate_fd_interior_mat(nx, nz, hx, hz):
    """Dxx (nx x nx), periodic Dzz and Dz (nz x nz) (waveguide_FD.jl:8-33)."""
    def tri(n, lo, di, up, wrap_lo=None, wrap_up=None):
        M = sp.diags([np.full(n - 1, lo), np.full(n, di), np.full(n - 1, up)], [-1, 0, 1], format="lil")
        if wrap_lo is not None:
            M[0, n - 1], M[n - 1, 0] = wrap_lo, wrap_up
        return sp.csc_matrix(M)
    return (tri(nx, 1.0, -2.0, 1.0) / hx ** 2, tri(nz, 1.0, -2.0, 1.0, 1.0, 1.0) / hz ** 2,
            tri(nz, -1.0, 0.0, 1.0, -1.0, 1.0) / (2 * hz))


def generate_fd_boundary_mat(nx, nz, hx, hz):
    """C1 (nx nz x 2 nz) and C2T (2 nz x nx nz) (waveguide_FD.jl:41-63)."""
    m = nx * nz
    z = np.arange(nz)
    C1 = sp.csc_matrix((np.full(2 * nz, 1 / hx ** 2), (np.concatenate([z, z + nz * (nx - 1)]), np.arange(2 * nz))), shape=(m, 2 * nz))
    d1, d2 = 2 / hx, -1 / (2 * hx)
    rows = np.concatenate([z, z, nz + z, nz + z])
    cols = np.concatenate([z, z + nz, z + nz * (nx - 1), z + nz * (nx - 2)])
    vals = np.concatenate([np.full(nz, d1), np.full(nz, d2), np.full(nz, d1), np.full(nz, d2)])
    return C1, sp.csc_matrix((vals, (rows, cols)), shape=(2 * nz, m))


def generate_wavenumber_fd(nx, nz, wg, delta):
    """Squared wavenumber on the grid and at the two ends (waveguide_FD.jl:74-182): K (nz x nx), hx, hz, Km, Kp."""
    pi = math.pi
    if wg == "TAUSCH":
        xm, xp = -delta, 2 / pi + 0.4 + delta
    elif wg == "JARLEBRING":
        xm, xp = -1 - delta, 1 + delta
    else:
        raise ValueError("No wavenumber loaded: The given Waveguide '%s' is not supported in 'FD' discretization." % wg)
    X = np.linspace(xm, xp, nx + 2)[1:-1][None, :] * np.ones((nz, 1))
    Z = np.linspace(0.0, 1.0, nz + 1)[1:][:, None] * np.ones((1, nx))
    hx, hz = (xp - xm) / (nx + 1), 1.0 / nz
    k = np.zeros((nz, nx))
    if wg == "TAUSCH":
        k1, k2, k3 = math.sqrt(2.3) * pi, math.sqrt(3) * pi, pi
        mid = (X > 2 / pi) & (X <= 2 / pi + 0.4)
        k[X <= 0] = k1
        k[(X > 0) & (X <= 2 / pi)] = k2
        k[mid & (Z > 0.5)] = k2
        k[mid & (Z <= 0.5)] = k3
        k[X > 2 / pi + 0.4] = k3
        return k ** 2, hx, hz, k1, k3
    k1, k2, k3, k4 = math.sqrt(2.3) * pi, 2 * math.sqrt(3) * pi, 4 * math.sqrt(3) * pi, pi
    left = (X > -1) & (X <= 0)
    k[X <= -1] = k1
    k[X > 1] = k4
    k[(X > 0.5) & (X <= 1) & (Z <= 0.4)] = k4
    k[(X > 0) & (X <= 0.5)] = k3
    k[(X > 0.5) & (X <= 1) & (Z > 0.4)] = k3
    k[left & (Z > 0.5) & (Z - X / 2 <= 1)] = k3
    k[left & (Z > 0.5) & (Z - X / 2 > 1)] = k2
    k[left & (Z <= 0.5) & (Z + X / 2 > 0)] = k3
    k[left & (Z <= 0.5) & (Z + X / 2 <= 0)] = k2
    return k ** 2, hx, hz, k1, k4


# ---- WEP_FD ---------------------------------------------------------------------------------------------------------------
class WEP_FD:
    """Device-resident WEP_FD (Waveguide.jl:203-240).  n = nx*nz + 2*nz; not an SPMF: no compute_Mder (as the reference,
    Waveguide.jl:384-386) -- linear systems go through WEPLinSolverCreator."""

    def __init__(self, nx, nz, hx, hz, Dxx, Dzz, Dz, C1, C2T, K, Km, Kp):
        self.nx, self.nz, self.hx, self.hz = int(nx), int(nz), float(hx), float(hz)
        self.Dxx, self.Dzz, self.Dz, self.C1, self.C2T = Dxx, Dzz, Dz, C1, C2T
        self.n = self.nx * self.nz + 2 * self.nz
        self.k_bar = complex(np.mean(K))
        self.K = np.asfortranarray(np.asarray(K, dtype=np.complex128) - self.k_bar)
        p = (nz - 1) / 2
        self.p = p
        self.d0, self.d1, self.d2 = -3 / (2 * hx), 2 / hx, -1 / (2 * hx)
        modes = np.arange(nz) - p
        self.b = 4 * math.pi * 1j * modes
        self.cM = (Km ** 2 - 4 * math.pi ** 2 * modes ** 2).astype(np.complex128)
        self.cP = (Kp ** 2 - 4 * math.pi ** 2 * modes ** 2).astype(np.complex128)
        self.bb = np.exp(-2j * math.pi * np.arange(nz) * (-p) / nz)
        self.bbinv = 1 / self.bb
        h = C.c_void_p()
        kb = np.array([self.k_bar], dtype=np.complex128)
        check(lib.nepb_wep_create(self.nx, self.nz, self.hx, self.hz, ptr(self.K), ptr(kb), ptr(np.ascontiguousarray(self.bb)), C.byref(h)))
        self._h = h
        self._table = (None, 0)  # (lambda, columns) of the derivative table resident on the device

    def close(self):
        if getattr(self, "_h", None):
            lib.nepb_wep_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def size(self, dim=None):
        return (self.n, self.n) if dim is None else self.n

    # -- host-side scalars ------------------------------------------------------------------------------------------------
    def s_values(self, lam):
        """[sM(lambda); sP(lambda)] (Waveguide.jl:181-189)."""
        lam = complex(lam)
        beta = np.concatenate([lam * lam + self.b * lam + self.cM, lam * lam + self.b * lam + self.cP])
        return 1j * np.sign(beta.imag) * np.sqrt(beta) + self.d0

    def derivative_table(self, lam, ncols):
        """D[m, :] = 1im * sqrt_derivative(1, b_m, c_m, ncols - 1, lambda), + d0 in column 0 (Waveguide.jl:351-373); (2 nz, ncols)."""
        D = 1j * sqrt_derivative(1.0, np.concatenate([self.b, self.b]), np.concatenate([self.cM, self.cP]), ncols - 1, complex(lam))
        D[:, 0] += self.d0
        return np.ascontiguousarray(D)

    def boundary_coefficients(self, lam, a):
        """coef[m, j] = a_j D[m, j]; row-major (2 nz, na): the explicit coefficient block of nepb_wep_mlincomb_block."""
        return np.ascontiguousarray(self.derivative_table(lam, len(a)) * np.asarray(a, dtype=np.complex128)[None, :])

    def _ensure_table(self, lam, na):
        """The derivative table of `lam` with >= na columns on the device: a solver loop at a fixed shift (iar, tiar, resinv's
        Rayleigh functional) computes and uploads it once and grows it geometrically."""
        tl, tc = self._table
        if tl != complex(lam) or tc < na:
            ncols = na if tl != complex(lam) else max(na, 2 * tc)
            check(lib.nepb_wep_set_table(self._h, ncols, ptr(self.derivative_table(lam, ncols))))
            self._table = (complex(lam), ncols)

    # -- compute_Mlincomb -------------------------------------------------------------------------------------------------
    def mlincomb_block(self, lam, Vb: Block, vcol0, na, a, Zb: Block, zcol, use_table=True):
        """Z[:, zcol] = sum_j a_j M^{(j)}(lambda) V[:, vcol0 + j] with all operands in HBM.  use_table: keep the derivative table
        of `lam` on the device (solver loops at a fixed shift); False: one explicit coefficient block for this call."""
        a = np.ascontiguousarray(np.asarray(a, dtype=np.complex128))
        if len(a) != na:
            raise ValueError("Incompatible sizes: Number of coefficients = %d, number of vectors = %d." % (len(a), na))
        lam_ = np.array([complex(lam)], dtype=np.complex128)
        coef = None
        if use_table:
            self._ensure_table(lam, na)
        else:
            coef = self.boundary_coefficients(lam, a)
        check(lib.nepb_wep_mlincomb_block(self._h, ptr(lam_), Vb._h, vcol0, na, ptr(a), None if coef is None else ptr(coef), Zb._h, zcol))

    def compute_Mlincomb(self, lam, V, a=None, startder=None):
        V = np.asarray(V, dtype=np.complex128)
        Vm = V.reshape(-1, 1) if V.ndim == 1 else V
        if Vm.shape[0] != self.n:
            raise ValueError("Incompatible sizes: Length of vectors = %d, size of NEP = %d." % (Vm.shape[0], self.n))
        k = Vm.shape[1]
        a = np.ones(k, dtype=np.complex128) if a is None else np.asarray(a, dtype=np.complex128)
        if len(a) != k:
            raise ValueError("Incompatible sizes: Number of coefficients = %d, number of vectors = %d." % (len(a), k))
        if startder:  # NEPCore.jl:156-160
            a = np.concatenate([np.zeros(startder, dtype=np.complex128), a])
            Vm = np.concatenate([np.zeros((self.n, startder), dtype=np.complex128), Vm], axis=1)
        Vb = Block.from_host(Vm)
        Zb = Block(self.n, 1)
        self.mlincomb_block(lam, Vb, 0, Vm.shape[1], a, Zb, 0)
        z = Zb.download()[:, 0].copy()
        Vb.close()
        Zb.close()
        return z

    def compute_Mder(self, lam, i=0):
        raise NotImplementedError("The WEP does not implement this function. If this was called in a situation where you want to "
                                  "solve linear systems please look at `WEPLinSolverCreator`")

    def residual_norms(self, lams, Q):
        """||M(lambda_s) q_s|| / ||q_s|| for all columns (the ResidualErrmeasure of errmeasure.jl:128-130), one upload."""
        Q = np.asarray(Q, dtype=np.complex128)
        Qb, Zb = Block.from_host(Q), Block(self.n, Q.shape[1])
        for s, lam in enumerate(lams):
            self.mlincomb_block(lam, Qb, s, 1, [1.0], Zb, s, use_table=False)
        out = np.empty(Q.shape[1])
        check(lib.nepb_block_colnorms(Zb._h, 0, Q.shape[1], self.n, ptr(out)))
        Qb.close()
        Zb.close()
        return out / np.linalg.norm(Q, axis=0)

    def residual_block(self, lams, Qb: Block, k, Rb: Block):
        """The same with Q already in HBM (the device-resident iar / tiar loops)."""
        for s, lam in enumerate(lams[:k]):
            self.mlincomb_block(lam, Qb, s, 1, [1.0], Rb, s, use_table=False)
        num, den = np.empty(k), np.empty(k)
        check(lib.nepb_block_colnorms(Rb._h, 0, k, self.n, ptr(num)))
        check(lib.nepb_block_colnorms(Qb._h, 0, k, self.n, ptr(den)))
        return num / den

    # -- boundary operator and Schur complement ----------------------------------------------------------------------------
    def Pinv(self, lam, x):
        """[R(Rinv(x1) ./ sM); R(Rinv(x2) ./ sP)] (Waveguide.jl:160-163) on the device."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.complex128))
        coef = np.ascontiguousarray(1.0 / self.s_values(lam))
        y = np.empty(2 * self.nz, dtype=np.complex128)
        check(lib.nepb_wep_pinv(self._h, ptr(coef), ptr(x), ptr(y)))
        return y

    def schur_matvec_block(self, lam, Xb: Block, xcol, Yb: Block, ycol):
        lam_ = np.array([complex(lam)], dtype=np.complex128)
        coef = np.ascontiguousarray(1.0 / self.s_values(lam))
        check(lib.nepb_wep_schur_matvec_block(self._h, ptr(lam_), ptr(coef), Xb._h, xcol, Yb._h, ycol))

    def A(self, lam, d=0):
        Iz = sp.identity(self.nz, format="csc", dtype=np.complex128)
        if d == 0:
            return (self.Dzz + 2 * lam * self.Dz + (lam ** 2 + self.k_bar) * Iz).tocsc()
        if d == 1:
            return (2 * self.Dz + 2 * lam * Iz).tocsc()
        return 2 * Iz if d == 2 else sp.csc_matrix((self.nz, self.nz), dtype=np.complex128)


class SchurMatVec:
    """v -> (A(lambda) X + X B + K .* X) - C1 Pinv(lambda, C2T v) (Waveguide.jl:388-420), one device pass per product."""

    def __init__(self, nep: WEP_FD, lam):
        self.nep, self.lam = nep, complex(lam)
        m = nep.nx * nep.nz
        self._x, self._y = Block(m, 1), Block(m, 1)

    def __call__(self, v):
        self._x.upload(np.asarray(v, dtype=np.complex128))
        self.nep.schur_matvec_block(self.lam, self._x, 0, self._y, 0)
        return self._y.download()[:, 0].copy()

    __mul__ = __call__

    def size(self, dim=None):
        m = self.nep.nx * self.nep.nz
        return (m, m) if dim is None else m


def construct_WEP_schur_complement(nep: WEP_FD, lam):
    """Kronecker form of Ringh, Proposition 3.1 (Waveguide.jl:523-549): the five-point interior operator plus the four dense
    nz x nz blocks of the boundary operators in the first and last block row.  Host assembly of the matrix the device
    factorises (the reference assembles it in Julia the same way); the columns of Pinv_minus / Pinv_plus come from the
    device boundary operator applied to the unit vectors, as in the reference's loop."""
    nx, nz = nep.nx, nep.nz
    Pm = np.empty((nz, nz), dtype=np.complex128)
    Pp = np.empty((nz, nz), dtype=np.complex128)
    e = np.zeros(2 * nz, dtype=np.complex128)
    for i in range(nz):  # columns P_inv_m(nep, lambda, e_i), P_inv_p(nep, lambda, e_i) (:531-537), both halves in one device call
        e[i] = e[nz + i] = 1
        col = nep.Pinv(lam, e)
        Pm[:, i], Pp[:, i] = col[:nz], col[nz:]
        e[i] = e[nz + i] = 0
    E = sp.csc_matrix(([nep.d1 / nep.hx ** 2, nep.d2 / nep.hx ** 2], ([0, 0], [0, 1])), shape=(nx, nx))
    EE = sp.csc_matrix(([nep.d1 / nep.hx ** 2, nep.d2 / nep.hx ** 2], ([nx - 1, nx - 1], [nx - 1, nx - 2])), shape=(nx, nx))
    Inz = sp.identity(nz, format="csc", dtype=np.complex128)
    Inx = sp.identity(nx, format="csc", dtype=np.complex128)
    S = (sp.kron(nep.Dxx.T, Inz) + sp.kron(Inx, nep.A(lam)) + sp.diags(nep.K.reshape(-1, order="F"))
         - sp.kron(E, sp.csc_matrix(Pm)) - sp.kron(EE, sp.csc_matrix(Pp)))
    return sp.csc_matrix(S, dtype=np.complex128)


class _WEPLinSolver(LinSolver):
    """lin_solve of the WEP solvers (Ringh, Proposition 2.1; Waveguide.jl:555-567): eliminate the boundary unknowns, solve
    with the Schur complement, substitute back."""

    def __init__(self, nep: WEP_FD, lam):
        self.nep, self.lam = nep, complex(lam)

    def inner_solve(self, rhs, tol):
        raise NotImplementedError

    def lin_solve(self, x, tol=np.finfo(float).eps):
        nep, lam = self.nep, self.lam
        x = np.asarray(x, dtype=np.complex128).reshape(-1)
        m = nep.nx * nep.nz
        x_int, x_ext = x[:m], x[m:]
        rhs = x_int - nep.C1 @ nep.Pinv(lam, x_ext)
        q = self.inner_solve(rhs, tol)
        return np.concatenate([q, nep.Pinv(lam, -(nep.C2T @ q) + x_ext)])


class WEPFactorizedLinSolver(_WEPLinSolver):
    """The Schur complement factorised once by the device multifrontal LU (Waveguide.jl:476-489)."""

    def __init__(self, nep, lam, umfpack_refinements=2):
        super().__init__(nep, lam)
        self.schur = B200SPMF([construct_WEP_schur_complement(nep, lam)], [ONE])
        self.fact = B200FactorizeLinSolver(self.schur, 0.0, umfpack_refinements)

    def inner_solve(self, rhs, tol):
        return self.fact.lin_solve(rhs)


class WEPBackslashLinSolver(_WEPLinSolver):
    """Keeps the assembled Schur complement and factorises at every solve (Waveguide.jl:459-473)."""

    def __init__(self, nep, lam):
        super().__init__(nep, lam)
        self.schur = B200SPMF([construct_WEP_schur_complement(nep, lam)], [ONE])

    def inner_solve(self, rhs, tol):
        f = B200FactorizeLinSolver(self.schur, 0.0, 2)
        q = f.lin_solve(rhs)
        f.lu.close()
        return q


class WEPGMRESLinSolver(_WEPLinSolver):
    """GMRES on the matrix-free Schur complement (Waveguide.jl:424-456); kwargs as IterativeSolvers' gmres: Pl (a callable
    applying the left preconditioner), restart, maxiter, reltol, log."""

    def __init__(self, nep, lam, kwargs=()):
        super().__init__(nep, lam)
        self.matvec = SchurMatVec(nep, lam)
        self.kwargs = dict(kwargs)

    def inner_solve(self, rhs, tol):
        kw = dict(self.kwargs)
        tol = kw.pop("reltol", tol)
        out = gmres(self.matvec, rhs, tol, restart=kw.get("restart", min(20, len(rhs))), maxiter=kw.get("maxiter"), Pl=kw.get("Pl"),
                    log=kw.get("log", False))
        return out[0] if isinstance(out, tuple) else out


class WEPLinSolverCreator(LinSolverCreator):
    """WEPLinSolverCreator(; solver_type = :factorized, kwargs = ()) (Waveguide.jl:491-521)."""

    def __init__(self, solver_type="factorized", kwargs=()):
        self.solver_type, self.kwargs = solver_type, kwargs

    def create_linsolver(self, nep, lam):
        if not isinstance(nep, WEP_FD):
            raise TypeError("WEPLinSolver can only be used in combination with WEPs: type(nep)=%s" % type(nep).__name__)
        if self.solver_type == "backslash":
            return WEPBackslashLinSolver(nep, lam)
        if self.solver_type == "gmres":
            return WEPGMRESLinSolver(nep, lam, self.kwargs)
        if self.solver_type == "factorized":
            return WEPFactorizedLinSolver(nep, lam)
        raise ValueError("Unknown type of solver_type in linsolvercreator:%s" % self.solver_type)


# ---- gallery ----------------------------------------------------------------------------------------------------------------
def assemble_waveguide_spmf_fd(nx, nz, hx, Dxx, Dzz, Dz, C1, C2T, K, Km, Kp):
    """The SPMF format (Waveguide.jl:9-45): 3 polynomial terms and 2 nz rank-one boundary terms S_j(lambda) E_j."""
    Ix = sp.identity(nx, format="csc", dtype=np.complex128)
    Iz = sp.identity(nz, format="csc", dtype=np.complex128)
    m, e = nx * nz, 2 * nz
    Q0 = sp.kron(Ix, Dzz) + sp.kron(Dxx, Iz) + sp.diags(np.asarray(K, dtype=np.complex128).reshape(-1, order="F"))

    def embed(Q, with_c=False):
        return sp.bmat([[Q, C1 if with_c else sp.csc_matrix((m, e))], [C2T if with_c else sp.csc_matrix((e, m)), sp.csc_matrix((e, e))]],
                       format="csc", dtype=np.complex128)
    A = [embed(Q0, True), embed(sp.kron(Ix, 2 * Dz)), embed(sp.kron(Ix, Iz))]
    f = [ONE, IDENTITY, Monomial(2)]
    p = (nz - 1) / 2
    modes = np.arange(nz) - p
    bbv = np.exp(-2j * math.pi * np.arange(nz) * (-p) / nz)
    Rcols = (bbv[:, None] * np.fft.fft(np.eye(nz), axis=0))[::-1, :]  # column j = R(e_j)
    d0 = -3 / (2 * hx)
    b = 4 * math.pi * 1j * modes
    for half, kk in ((0, Km), (1, Kp)):
        cc = kk ** 2 - 4 * math.pi ** 2 * modes ** 2
        for j in range(nz):
            col = np.zeros(e, dtype=np.complex128)
            col[half * nz:(half + 1) * nz] = Rcols[:, j]
            Ej = sp.csc_matrix(np.outer(col, np.conj(col) / nz))
            A.append(sp.bmat([[sp.csc_matrix((m, m)), None], [None, Ej]], format="csc", dtype=np.complex128))
            f.append(SqrtQuadratic(b[j], cc[j], d0))
    return SPMF_NEP(A, f)


def nep_gallery_WEP(nx=3 * 5 * 7, nz=3 * 5 * 7, benchmark_problem="TAUSCH", neptype="WEP", delta=0.1):
    """nep_gallery(WEP; nx, nz, benchmark_problem, neptype, delta) (GalleryWaveguide.jl:60-92).  neptype "WEP": the native
    device format; "SPMF": the host descriptor of the 3 + 2 nz term SPMF (B200SPMF.from_nep puts it on the device)."""
    wg = benchmark_problem.upper()
    if nx < 3 or nz < 1:
        raise ValueError("nx >= 3 and nz >= 1 are required")
    K, hx, hz, Km, Kp = generate_wavenumber_fd(nx, nz, wg, delta)
    Dxx, Dzz, Dz = generate_fd_interior_mat(nx, nz, hx, hz)
    C1, C2T = generate_fd_boundary_mat(nx, nz, hx, hz)
    if neptype == "SPMF":
        return assemble_waveguide_spmf_fd(nx, nz, hx, Dxx, Dzz, Dz, C1, C2T, K, Km, Kp)
    if neptype == "WEP":
        return WEP_FD(nx, nz, hx, hz, Dxx, Dzz, Dz, C1, C2T, K, Km, Kp)
    raise ValueError("The NEP-type '%s' is not supported for the waveguide eigenvalue problem." % neptype)
