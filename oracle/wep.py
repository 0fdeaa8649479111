"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product) for the waveguide eigenvalue problem, SURVEY 8(f)
rank 3: a NumPy / SciPy restatement of the reference's `GalleryWaveguide` FD path.

Follows (reference file:line)
    src/gallery_extra/waveguide/waveguide_FD.jl:8-33      generate_fd_interior_mat
    src/gallery_extra/waveguide/waveguide_FD.jl:41-63     generate_fd_boundary_mat
    src/gallery_extra/waveguide/waveguide_FD.jl:90-182    generate_wavenumber_fd (TAUSCH, JARLEBRING)
    src/gallery_extra/waveguide/Waveguide.jl:9-45         assemble_waveguide_spmf_fd (the SPMF format, 3 + 2 nz terms)
    src/gallery_extra/waveguide/Waveguide.jl:52-101       generate_R_matvecs, generate_S_function
    src/gallery_extra/waveguide/Waveguide.jl:146-189      sqrt_pos_imag, Pinv, R, Rinv, sM, sP
    src/gallery_extra/waveguide/Waveguide.jl:203-240      WEP_FD
    src/gallery_extra/waveguide/Waveguide.jl:324-379      compute_Mlincomb(::WEP_FD)
    src/gallery_extra/waveguide/Waveguide.jl:393-402      SchurMatVec
    src/gallery_extra/waveguide/Waveguide.jl:523-549      construct_WEP_schur_complement
    src/gallery_extra/waveguide/Waveguide.jl:555-567      lin_solve (Ringh, Proposition 2.1)
    src/gallery_extra/waveguide/Waveguide.jl:574-616      sqrt_derivative
    src/gallery_extra/GalleryWaveguide.jl:60-92           nep_gallery(WEP, ...)

Pinned (tests/test_wep_oracle.py) to the reference's own tests: the SPMF and the native format agree to 1e-14 at nx = 11,
nz = 7 and lambda = -1.3 - 0.31im (test/wep_small.jl:17-26); resinv with the Schur-complement solver on JARLEBRING,
nx = 109, nz = 105, from lambda0 = -3 - 3.5im converges to the reference eigenvalue
-2.743228671961724 - 3.1439375599649972im (test/wep_small.jl:33-47).  FFTW's butterflies are third-party arithmetic
(unpinned); numpy.fft stands in, agreement is at the 1e-14 level of the reference's own format-equivalence test."""
import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla


# ---- discretisation (waveguide_FD.jl) ----------------------------------------------------------------------------------------
def generate_fd_interior_mat(nx, nz, hx, hz):
    ex, ez = np.ones(nx), np.ones(nz)
    Dxx = sp.diags([ex[:-1], -2 * ex, ex[:-1]], [-1, 0, 1], format="lil")
    Dzz = sp.diags([ez[:-1], -2 * ez, ez[:-1]], [-1, 0, 1], format="lil")
    Dzz[0, nz - 1] = 1   # periodicity in z (:17-18)
    Dzz[nz - 1, 0] = 1
    Dz = sp.diags([-ez[:-1], ez[:-1]], [-1, 1], format="lil")
    Dz[0, nz - 1] = -1   # (:28-29)
    Dz[nz - 1, 0] = 1
    return sp.csc_matrix(Dxx) / hx ** 2, sp.csc_matrix(Dzz) / hz ** 2, sp.csc_matrix(Dz) / (2 * hz)


def generate_fd_boundary_mat(nx, nz, hx, hz):
    e1 = sp.csc_matrix(([1.0], ([0], [0])), shape=(nx, 1))
    en = sp.csc_matrix(([1.0], ([nx - 1], [0])), shape=(nx, 1))
    Iz = sp.identity(nz, format="csc")
    C1 = sp.hstack([sp.kron(e1, Iz), sp.kron(en, Iz)]).tocsc() / hx ** 2
    d1, d2 = 2 / hx, -1 / (2 * hx)
    vm = sp.csc_matrix(([d1, d2], ([0, 0], [0, 1])), shape=(1, nx))
    vp = sp.csc_matrix(([d1, d2], ([0, 0], [nx - 1, nx - 2])), shape=(1, nx))
    C2T = sp.vstack([sp.kron(vm, Iz), sp.kron(vp, Iz)]).tocsc()
    return C1, C2T


def _grid(nx, nz, xm, xp):
    X = np.linspace(xm, xp, nx + 2)
    hx = (xp - xm) / (nx + 1)
    Z = np.linspace(0.0, 1.0, nz + 1)
    hz = 1.0 / nz
    return X[1:-1][None, :], Z[1:][:, None], hx, hz


def generate_wavenumber_fd(nx, nz, wg, delta):
    """Returns K (nz x nx, squared wavenumber), hx, hz, Km, Kp (waveguide_FD.jl:74-182)."""
    if wg == "TAUSCH":
        x, z, hx, hz = _grid(nx, nz, 0 - delta, 2 / math.pi + 0.4 + delta)
        k1, k2, k3 = math.sqrt(2.3) * math.pi, math.sqrt(3) * math.pi, math.pi
        one = np.ones_like(z)
        k = (k1 * (x <= 0) * one + k2 * (x > 0) * (x <= 2 / math.pi) * one
             + k2 * (x > 2 / math.pi) * (x <= 2 / math.pi + 0.4) * (z > 0.5)
             + k3 * (x > 2 / math.pi) * (z <= 0.5) * (x <= 2 / math.pi + 0.4)
             + k3 * (x > 2 / math.pi + 0.4) * one)
        return (k ** 2).astype(np.float64), hx, hz, k1, k3
    if wg == "JARLEBRING":
        x, z, hx, hz = _grid(nx, nz, -1 - delta, 1 + delta)
        k1, k2, k3, k4 = math.sqrt(2.3) * math.pi, 2 * math.sqrt(3) * math.pi, 4 * math.sqrt(3) * math.pi, math.pi
        one, xone = np.ones_like(z), np.ones_like(x)
        zz, xx = z * xone, x * one
        k = (k1 * (x <= -1) * one + k4 * (x > 1) * one
             + k4 * (x > 0.5) * (x <= 1) * (z <= 0.4)
             + k3 * (x > 0) * (x <= 0.5) * one
             + k3 * (x > 0.5) * (x <= 1) * (z > 0.4)
             + k3 * (x > -1) * (x <= 0) * (z > 0.5) * (zz - xx / 2 <= 1)
             + k2 * (x > -1) * (x <= 0) * (z > 0.5) * (zz - xx / 2 > 1)
             + k3 * (x > -1) * (x <= 0) * (z <= 0.5) * (zz + xx / 2 > 0)
             + k2 * (x > -1) * (x <= 0) * (z <= 0.5) * (zz + xx / 2 <= 0))
        return (k ** 2).astype(np.float64), hx, hz, k1, k4
    raise ValueError("No wavenumber loaded: The given Waveguide '%s' is not supported in 'FD' discretization." % wg)


# ---- scalar helpers (Waveguide.jl:127-144, 574-616) -------------------------------------------------------------------------
def sqrt_pos_imag(a):
    a = complex(a)
    s = np.sign(a.imag)
    return np.sqrt(a) if s == 0 else s * np.sqrt(a)


def sqrt_derivative(a, b, c, d=0, x=0):
    """All derivatives 0..d of sqrt(a z^2 + b z + c) at z = x on the branch with positive imaginary part (Gegenbauer recurrence,
    :574-616).  Returns a scalar for d == 0, else an array of d + 1 values."""
    if d < 0:
        raise ValueError("Cannot take negative derivative. d = %d" % d)
    aa, bb, cc = a, b + 2 * a * x, c + a * x ** 2 + b * x
    der = np.zeros(d + 1, dtype=np.complex128)
    yi = sqrt_pos_imag(cc)
    der[0] = yi
    if d == 0:
        return der[0]
    yip1 = bb / (2 * sqrt_pos_imag(cc))
    fact = 1.0
    der[1] = yip1
    for i in range(2, d + 1):
        m = i - 2
        yip2 = -(2 * aa * (m - 1) * yi + bb * (1 + 2 * m) * yip1) / (2 * cc * (2 + m))
        fact *= i
        yi, yip1 = yip1, yip2
        der[i] = yip2 * fact
    return der


# ---- WEP_FD (Waveguide.jl:203-379) -------------------------------------------------------------------------------------------
class WEP_FD:
    def __init__(self, nx, nz, hx, hz, Dxx, Dzz, Dz, C1, C2T, K, Km, Kp):
        self.nx, self.nz, self.hx, self.hz = nx, nz, hx, hz
        self.Dxx, self.Dzz, self.Dz, self.C1, self.C2T = Dxx, Dzz, Dz, C1, C2T
        self.n = nx * nz + 2 * nz
        self.k_bar = complex(np.mean(K))
        self.K = K.astype(np.complex128) - self.k_bar
        p = (nz - 1) / 2
        self.p = p
        self.d0, self.d1, self.d2 = -3 / (2 * hx), 2 / hx, -1 / (2 * hx)
        rng = np.arange(nz) - p
        self.b = 4 * math.pi * 1j * rng
        self.cM = Km ** 2 - 4 * math.pi ** 2 * rng ** 2 + 0j
        self.cP = Kp ** 2 - 4 * math.pi ** 2 * rng ** 2 + 0j
        self.bb = np.exp(-2j * math.pi * np.arange(nz) * (-p) / nz)
        self.bbinv = 1 / self.bb
        self.Iz = sp.identity(nz, format="csc", dtype=np.complex128)

    # boundary transforms (:165-171)
    def R(self, x):
        return (self.bb * np.fft.fft(np.asarray(x).ravel()))[::-1]

    def Rinv(self, x):
        return np.fft.ifft(self.bbinv * np.asarray(x).ravel()[::-1])

    def sM(self, lam):
        beta = lam ** 2 + self.b * lam + self.cM
        return 1j * np.sign(beta.imag) * np.sqrt(beta) + self.d0

    def sP(self, lam):
        beta = lam ** 2 + self.b * lam + self.cP
        return 1j * np.sign(beta.imag) * np.sqrt(beta) + self.d0

    def Pinv(self, lam, x):
        h = len(x) // 2
        return np.concatenate([self.R(self.Rinv(x[:h]) / self.sM(lam)), self.R(self.Rinv(x[h:]) / self.sP(lam))])

    def A(self, lam, d=0):
        if d == 0:
            return self.Dzz + 2 * lam * self.Dz + (lam ** 2 + self.k_bar) * self.Iz
        if d == 1:
            return 2 * self.Dz + 2 * lam * self.Iz
        if d == 2:
            return 2 * self.Iz
        return sp.csc_matrix((self.nz, self.nz), dtype=np.complex128)

    def B(self, lam, d=0):
        return self.Dxx if d == 0 else sp.csc_matrix((self.nx, self.nx))

    def native_Mlincomb(self, lam, V, a=None):  # hook of oracle.nep.compute_Mlincomb
        return wep_compute_Mlincomb(self, lam, V, a)


def wep_compute_Mlincomb(nep, lam, V, a=None):
    """compute_Mlincomb(::WEP_FD, lambda, V, a) (:324-379): sum_i a_i M^{(i)}(lambda) v_i."""
    V = np.asarray(V, dtype=np.complex128)
    V = V.reshape(-1, 1) if V.ndim == 1 else V
    na = V.shape[1]
    a = np.ones(na, dtype=np.complex128) if a is None else np.asarray(a, dtype=np.complex128)
    if len(a) != na:
        raise ValueError("Incompatible sizes: Number of coefficients = %d, number of vectors = %d." % (len(a), na))
    if V.shape[0] != nep.n:
        raise ValueError("Incompatible sizes: Length of vectors = %d, size of NEP = %d." % (V.shape[0], nep.n))
    nx, nz = nep.nx, nep.nz
    lam = complex(lam)
    max_d = na - 1
    V1 = V[:nx * nz, :]
    V1m = [V1[:, j].reshape(nz, nx, order="F") for j in range(na)]
    V2 = V[nx * nz:, :]
    y1m = (nep.A(lam) @ V1m[0] + V1m[0] @ nep.B(lam).toarray() + nep.K * V1m[0]) * a[0]   # (:343)
    for d in range(1, min(max_d, 3) + 1):
        y1m = y1m + (nep.A(lam, d) @ V1m[d]) * a[d]
    y1 = y1m.reshape(-1, order="F") + (nep.C1 @ V2[:, 0]) * a[0]
    D = np.zeros((2 * nz, na), dtype=np.complex128)    # (:351-361)
    cMP = np.concatenate([nep.cM, nep.cP])
    for j in range(2 * nz):
        der = 1j * np.atleast_1d(sqrt_derivative(1, nep.b[j % nz], cMP[j], max_d, lam))
        D[j, :] = der[:na]
    y2t = (D[:, 0] + nep.d0) * np.concatenate([nep.Rinv(V2[:nz, 0]), nep.Rinv(V2[nz:, 0])]) * a[0]
    for jj in range(1, na):
        y2t = y2t + D[:, jj] * np.concatenate([nep.Rinv(V2[:nz, jj]), nep.Rinv(V2[nz:, jj])]) * a[jj]
    y2 = np.concatenate([nep.R(y2t[:nz]), nep.R(y2t[nz:])])
    y2 = y2 + (nep.C2T @ V1[:, 0]) * a[0]
    return np.concatenate([y1, y2])


def schur_matvec(nep, lam, v):
    """SchurMatVec (:393-402)."""
    X = np.asarray(v, dtype=np.complex128).reshape(nep.nz, nep.nx, order="F")
    top = (nep.A(lam) @ X + X @ nep.B(lam).toarray() + nep.K * X).reshape(-1, order="F")
    return top - nep.C1 @ nep.Pinv(lam, nep.C2T @ v)


def construct_WEP_schur_complement(nep, lam):
    """(:523-549) Kronecker form of Ringh, Proposition 3.1."""
    nx, nz = nep.nx, nep.nz
    Pm = np.zeros((nz, nz), dtype=np.complex128)
    Pp = np.zeros((nz, nz), dtype=np.complex128)
    sMi, sPi = 1 / nep.sM(lam), 1 / nep.sP(lam)
    for i in range(nz):
        e = np.zeros(nz, dtype=np.complex128)
        e[i] = 1
        Pm[:, i] = nep.R(nep.Rinv(e) * sMi)
        Pp[:, i] = nep.R(nep.Rinv(e) * sPi)
    E = sp.lil_matrix((nx, nx))
    E[0, 0] = nep.d1 / nep.hx ** 2
    E[0, 1] = nep.d2 / nep.hx ** 2
    EE = sp.lil_matrix((nx, nx))
    EE[nx - 1, nx - 1] = nep.d1 / nep.hx ** 2
    EE[nx - 1, nx - 2] = nep.d2 / nep.hx ** 2
    Inz = sp.identity(nz, format="csc", dtype=np.complex128)
    Inx = sp.identity(nx, format="csc", dtype=np.complex128)
    return (sp.kron(nep.B(lam).T, Inz) + sp.kron(Inx, nep.A(lam)) + sp.diags(nep.K.reshape(-1, order="F"))
            - sp.kron(E, Pm) - sp.kron(EE, Pp)).tocsc()


class WEPLinSolverCreator:
    """WEPLinSolverCreator(solver_type = :factorized) (:491-521); :backslash solves with the same Schur complement."""

    def create_linsolver(self, nep, lam):
        if not isinstance(nep, WEP_FD):
            raise TypeError("WEPLinSolver can only be used in combination with WEPs")
        return WEPFactorizedLinSolver(nep, lam)


class WEPFactorizedLinSolver:
    """WEPFactorizedLinSolver + lin_solve (:479-489, 555-567)."""

    def __init__(self, nep, lam):
        self.nep, self.lam = nep, complex(lam)
        self.fact = sla.splu(construct_WEP_schur_complement(nep, self.lam))

    def lin_solve(self, x, tol=0):
        nep, lam = self.nep, self.lam
        m = nep.nx * nep.nz
        x = np.asarray(x, dtype=np.complex128)
        x_int, x_ext = x[:m], x[m:]
        rhs = x_int - nep.C1 @ nep.Pinv(lam, x_ext)
        q = self.fact.solve(rhs)
        return np.concatenate([q, nep.Pinv(lam, -(nep.C2T @ q) + x_ext)])


# ---- SPMF format (Waveguide.jl:9-45) -----------------------------------------------------------------------------------------
def assemble_waveguide_spmf_fd(nx, nz, hx, Dxx, Dzz, Dz, C1, C2T, K, Km, Kp):
    """Returns (A list, f list of scalar callables): 3 polynomial terms and 2 nz boundary terms S_j(lambda) E_j."""
    Ix = sp.identity(nx, format="csc", dtype=np.complex128)
    Iz = sp.identity(nz, format="csc", dtype=np.complex128)
    Q0 = sp.kron(Ix, Dzz) + sp.kron(Dxx, Iz) + sp.diags(K.reshape(-1, order="F").astype(np.complex128))
    Q1 = sp.kron(Ix, 2 * Dz)
    Q2 = sp.kron(Ix, Iz)
    m, e = nx * nz, 2 * nz
    Z = sp.csc_matrix
    A = [sp.bmat([[Q0, C1], [C2T, Z((e, e))]]).tocsc().astype(np.complex128),
         sp.bmat([[Q1, Z((m, e))], [Z((e, m)), Z((e, e))]]).tocsc().astype(np.complex128),
         sp.bmat([[Q2, Z((m, e))], [Z((e, m)), Z((e, e))]]).tocsc().astype(np.complex128)]
    f = [lambda l: 1.0 + 0j, lambda l: complex(l), lambda l: complex(l) ** 2]
    p = (nz - 1) / 2
    bbv = np.exp(-2j * math.pi * np.arange(nz) * (-p) / nz)
    d0 = -3 / (2 * hx)
    rng = np.arange(nz) - p
    b = 4 * math.pi * 1j * rng
    cM = Km ** 2 - 4 * math.pi ** 2 * rng ** 2
    cP = Kp ** 2 - 4 * math.pi ** 2 * rng ** 2

    def R(x):
        return (bbv * np.fft.fft(x))[::-1]
    for half, cc in ((0, cM), (1, cP)):
        for j in range(nz):
            ej = np.zeros(nz)
            ej[j] = 1
            col = np.zeros(e, dtype=np.complex128)
            col[half * nz:(half + 1) * nz] = R(ej)
            Ej = np.outer(col, np.conj(col / nz))
            A.append(sp.bmat([[Z((m, m)), Z((m, e))], [Z((e, m)), sp.csc_matrix(Ej)]]).tocsc())
            f.append((lambda bj, cj: (lambda l: 1j * sqrt_pos_imag(complex(l) ** 2 + bj * complex(l) + cj) + d0))(b[j], cc[j]))
    return A, f


def spmf_compute_Mlincomb(A, f, lam, v):
    """compute_Mlincomb of the SPMF format for one vector: sum_i f_i(lambda) A_i v."""
    return sum(fi(lam) * (Ai @ v) for Ai, fi in zip(A, f))


def nep_gallery_wep(nx=3 * 5 * 7, nz=3 * 5 * 7, benchmark_problem="TAUSCH", neptype="WEP", delta=0.1):
    """nep_gallery(WEP; ...) (GalleryWaveguide.jl:60-92)."""
    wg = benchmark_problem.upper()
    K, hx, hz, Km, Kp = generate_wavenumber_fd(nx, nz, wg, delta)
    Dxx, Dzz, Dz = generate_fd_interior_mat(nx, nz, hx, hz)
    C1, C2T = generate_fd_boundary_mat(nx, nz, hx, hz)
    if neptype == "SPMF":
        return assemble_waveguide_spmf_fd(nx, nz, hx, Dxx, Dzz, Dz, C1, C2T, K, Km, Kp)
    if neptype == "WEP":
        return WEP_FD(nx, nz, hx, hz, Dxx, Dzz, Dz, C1, C2T, K, Km, Kp)
    raise ValueError("The NEP-type '%s' is not supported for the waveguide eigenvalue problem." % neptype)


# ---- Sylvester-SMW preconditioner (waveguide_preconditioner.jl) -----------------------------------------------------------
# Test infrastructure only: the product ships the matrix-free Schur-complement product and a GMRES solver that takes any left
# preconditioner as a callable (`Pl`, as IterativeSolvers' gmres); this restatement is what the GPU test hands to it, and it is
# pinned to test/wep_small.jl:28-31 (with as many domains as grid lines the preconditioner is the exact inverse, 1e-14).
def _F(v, n1):
    """F (:190-199): zero-padded FFT of length 2 (n1 + 1) along the first axis, rows 2..n of the result."""
    n = n1 + 1
    pad = np.zeros((2 * n, v.shape[1]), dtype=np.complex128)
    pad[1:n, :] = v
    return np.fft.fft(pad, axis=0)[1:n, :]


def _Fh(v, n1):
    """Fh (:201-211)."""
    n = n1 + 1
    pad = np.zeros((2 * n, v.shape[1]), dtype=np.complex128)
    pad[1:n, :] = v
    return np.fft.ifft(pad, axis=0)[1:n, :] * 2 * n


def _W(X):
    """W = Wh (:160-187): the (real, symmetric) sine-transform matrix of the eigenvectors of Dxx, through two FFTs."""
    n1 = X.shape[0]
    return (_F(X, n1) - _Fh(X, n1)) * ((1j / 2) / math.sqrt((n1 + 1) / 2.0))


def solve_wg_sylvester_fft(C, lam, k_bar, hx, hz):
    """solve_wg_sylvester_fft! (:114-158): A X + X B = C with A = Dzz + 2 lam Dz + (lam^2 + k_bar) I (circulant) and B = Dxx."""
    C = np.array(C, dtype=np.complex128)
    nz, nx = C.shape
    alpha = lam ** 2 + k_bar
    v = np.zeros(nz, dtype=np.complex128)
    v[0], v[1], v[nz - 1] = -2, 1, 1
    v = v / hz ** 2
    w = np.zeros(nz, dtype=np.complex128)
    w[1], w[nz - 1] = 1, -1
    w = w * (lam / hz)
    D = np.fft.fft(v + w) + alpha
    S = -(4 / hx ** 2) * np.sin(math.pi * np.arange(1, nx + 1) / (2 * (nx + 1))) ** 2
    T = C.conj().T
    C = _W(T).conj().T
    C = np.fft.ifft(C, axis=0) * math.sqrt(nx)   # Vh! (:170-175)
    Z = C / (D[:, None] + S[None, :])
    T = Z.conj().T
    C = _W(T).conj().T
    return np.fft.fft(C, axis=0) / math.sqrt(nx)  # V! (:163-168)


class WEPPreconditioner:
    """wep_generate_preconditioner / generate_smw_matrix / solve_smw / ldiv! (:9-111, 218-421) for a WEP_FD with nx = nz + 4
    and N domains in the z direction (nz / N an integer)."""

    def __init__(self, nep, N, sigma):
        import scipy.linalg as sl
        if nep.nz + 4 != nep.nx:
            raise ValueError("This implementation requires nx = nz + 4. Provided NEP has nz = %d and nx = %d" % (nep.nz, nep.nx))
        if nep.nz % N:
            raise ValueError("This implementation is uniform in the blocking and therefore requires nz/N to be an integer.")
        self.nep, self.N, self.sigma = nep, int(N), complex(sigma)
        n = nep.nz
        self.L = n // self.N
        self.dd1, self.dd2 = nep.d1 / nep.hx ** 2, nep.d2 / nep.hx ** 2
        mm = self.N ** 2 + 4 * self.N
        M = np.zeros((mm, mm), dtype=np.complex128)
        for k in range(1, mm + 1):
            E = self._Ek(k, 1.0, np.zeros((n, nep.nx), dtype=np.complex128))
            Fk = self._Linv(E)
            M[:, k - 1] = self._functionals(Fk)
        self.lu = sl.lu_factor(M + np.eye(mm))

    # index helpers (:236-255); 1-based i, j as in the reference
    def _II(self, i):
        return slice((i - 1) * self.L, i * self.L)

    def _JJ(self, j):
        return slice((j - 3) * self.L + 2, (j - 2) * self.L + 2)

    def _JJ2(self, j):
        n, N = self.nep.nz, self.N
        return {1: 0, 2: 1, N + 3: n + 2, N + 4: n + 3}[j]

    def _k2ij(self, k):
        N = self.N
        j = k % (N + 4) + (k % (N + 4) == 0) * (N + 4)
        return (k - j) // (N + 4) + 1, j

    def _Linv(self, rhs):
        nep = self.nep
        return solve_wg_sylvester_fft(rhs, self.sigma, nep.k_bar, nep.hx, nep.hz)

    def _Pm(self, v):
        nep = self.nep
        return -nep.R(nep.Rinv(v) / nep.sM(self.sigma))

    def _Pp(self, v):
        nep = self.nep
        return -nep.R(nep.Rinv(v) / nep.sP(self.sigma))

    def _Ek(self, k, a, Y, quirk=False):
        """Adds a * E~_k to Y (:265-293 with a = 1; :376-404 in solve_smw, whose j == 2 branch reads
        `Y[II(i), 2] += Y[II(i), 2] + alpha[k]*K[II(i), 2]`, i.e. it doubles what is already there: kept, quirk=True)."""
        nep, N = self.nep, self.N
        n, nx, K = nep.nz, nep.nx, nep.K
        i, j = self._k2ij(k)
        II = self._II(i)
        ek = np.zeros(n, dtype=np.complex128)
        if j in (1, 2, N + 3, N + 4):
            c = self._JJ2(j)
            if quirk and j == 2:
                Y[II, c] += Y[II, c] + a * K[II, c]
            else:
                Y[II, c] += a * K[II, c]
            ek[II] = self.dd1 if j in (1, N + 4) else self.dd2
            if j <= 2:
                Y[:, 0] += a * self._Pm(ek)
            else:
                Y[:, nx - 1] += a * self._Pp(ek)
        else:
            JJ = self._JJ(j)
            Y[II, JJ] += a * K[II, JJ]
        return Y

    def _functionals(self, X):
        mm, N, L = self.N ** 2 + 4 * self.N, self.N, self.L
        b = np.zeros(mm, dtype=np.complex128)
        for k in range(1, mm + 1):
            i, j = self._k2ij(k)
            if j in (1, 2, N + 3, N + 4):
                b[k - 1] = X[self._II(i), self._JJ2(j)].sum() / L
            else:
                b[k - 1] = X[self._II(i), self._JJ(j)].sum() / (L * L)
        return b

    def ldiv(self, B):
        """ldiv!(precond, B) (:27-30): B <- vec(solve_smw(reshape(B, nz, nx)))."""
        import scipy.linalg as sl
        nep = self.nep
        C = self._Linv(np.asarray(B, dtype=np.complex128).reshape(nep.nz, nep.nx, order="F"))
        alpha = sl.lu_solve(self.lu, self._functionals(C))
        Y = np.zeros((nep.nz, nep.nx), dtype=np.complex128)
        for k in range(1, len(alpha) + 1):
            self._Ek(k, alpha[k - 1], Y, quirk=True)
        return (C - self._Linv(Y)).reshape(-1, order="F")

    __call__ = ldiv


def wep_generate_preconditioner(nep, N, sigma):
    return WEPPreconditioner(nep, N, sigma)
