"""Host-side scalar / matrix functions f_i of an SPMF (they never cross the C ABI).

The reference passes Julia closures that accept a scalar or a square matrix (src/NEPTypes.jl:140-160)
and evaluates f_i(S)[:,1] of the scaled bidiagonal matrix S = bidiag(lam; s_1..s_{k-1}) on the host
(src/NEPTypes.jl:993-1004).  Because S = lam*I + N with N nilpotent, that column is
    f(S)[j,0] = f^{(j)}(lam)/j! * prod_{l<=j} s_l ,
so every function class below provides `bidiag_column(lam, s)` from its Taylor *ratios*
t_j/t_{j-1} -- products are formed as prod (ratio_l * s_l), which stays finite where the raw
derivatives under/overflow (gun: lam^(1/2-100), gamma^100 * 100!).  Arbitrary callables fall back to
scipy.linalg.funm on S, which is what the reference does.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla


def _is_mat(S):
    return isinstance(S, np.ndarray) and S.ndim == 2


class ScalarFunction:
    """f: C -> C, extended to square matrices."""

    def __call__(self, S):
        raise NotImplementedError

    def taylor(self, lam, m):
        """t_j = f^{(j)}(lam)/j!, j = 0..m-1 (may under/overflow for large m; see bidiag_column)."""
        raise NotImplementedError

    def bidiag_column(self, lam, s):
        """f(S)[:,0] for S = diag(lam) + subdiag(s); len(s) = k-1."""
        k = len(s) + 1
        t = self.taylor(lam, k)
        out = np.empty(k, dtype=np.complex128)
        prod = 1.0 + 0j
        out[0] = t[0]
        for j in range(1, k):
            prod = prod * s[j - 1]
            # a vanishing Taylor coefficient (polynomial terms beyond their degree) must give exactly zero even when the
            # running product of the scalings has overflowed (gamma^j * j! for deep Krylov spaces)
            out[j] = t[j] * prod if t[j] != 0 else 0.0
        return out

    def derivative(self, lam, j):
        """f^{(j)}(lam)."""
        return self.taylor(lam, j + 1)[j] * math.factorial(j)


class Monomial(ScalarFunction):
    """S -> c * S^d  (d = 0: c*I; the PEP basis of src/types_poly.jl:83-98, DEP's -S)."""

    def __init__(self, d, c=1.0):
        self.d, self.c = int(d), c

    def __call__(self, S):
        if _is_mat(S):
            return self.c * np.linalg.matrix_power(S.astype(np.complex128), self.d)
        return self.c * S ** self.d

    def taylor(self, lam, m):
        t = np.zeros(m, dtype=np.complex128)
        for j in range(min(m, self.d + 1)):
            t[j] = self.c * math.comb(self.d, j) * complex(lam) ** (self.d - j)
        return t


class Exp(ScalarFunction):
    """S -> c * exp(a*S)  (DEP terms exp(-tau*S), src/NEPTypes.jl:495-513)."""

    def __init__(self, a, c=1.0):
        self.a, self.c = a, c

    def __call__(self, S):
        if _is_mat(S):
            return self.c * sla.expm(self.a * S.astype(np.complex128))
        return self.c * np.exp(self.a * S)

    def taylor(self, lam, m):
        t = np.empty(m, dtype=np.complex128)
        t[0] = self.c * np.exp(self.a * complex(lam))
        for j in range(1, m):
            t[j] = t[j - 1] * self.a / j
        return t

    def bidiag_column(self, lam, s):
        k = len(s) + 1
        out = np.empty(k, dtype=np.complex128)
        out[0] = self.c * np.exp(self.a * complex(lam))
        for j in range(1, k):
            out[j] = out[j - 1] * (self.a / j) * s[j - 1]
        return out


class PowShift(ScalarFunction):
    """S -> c * (S - shift)^alpha, principal branch (gun: 1im*sqrt(S - sigma_c^2),
    src/gallery_extra/NLEVP_native.jl:13-14)."""

    def __init__(self, alpha, shift=0.0, c=1.0):
        self.alpha, self.shift, self.c = alpha, shift, c

    def __call__(self, S):
        if _is_mat(S):
            n = S.shape[0]
            B = S.astype(np.complex128) - self.shift * np.eye(n)
            if self.alpha == 0.5:
                return self.c * sla.sqrtm(B)
            return self.c * sla.fractional_matrix_power(B, self.alpha)
        return self.c * (complex(S) - self.shift) ** self.alpha

    def _ratios(self, lam, k):
        z = complex(lam) - self.shift
        return [(self.alpha - j + 1) / (j * z) for j in range(1, k)]

    def taylor(self, lam, m):
        t = np.empty(m, dtype=np.complex128)
        t[0] = self.c * (complex(lam) - self.shift) ** self.alpha
        for j, r in enumerate(self._ratios(lam, m), start=1):
            t[j] = t[j - 1] * r
        return t

    def bidiag_column(self, lam, s):
        k = len(s) + 1
        out = np.empty(k, dtype=np.complex128)
        out[0] = self.c * (complex(lam) - self.shift) ** self.alpha
        for j, r in enumerate(self._ratios(lam, k), start=1):
            out[j] = out[j - 1] * r * s[j - 1]
        return out


class Callable(ScalarFunction):
    """Arbitrary user function valid for scalars and square matrices, as in the reference."""

    def __init__(self, f):
        self.f = f

    def __call__(self, S):
        return self.f(S)

    def bidiag_column(self, lam, s):
        k = len(s) + 1
        S = np.diag(np.full(k, complex(lam))) + (np.diag(np.asarray(s, dtype=np.complex128), -1) if k > 1 else 0)
        return np.asarray(self.f(S), dtype=np.complex128)[:, 0]

    def taylor(self, lam, m):
        col = self.bidiag_column(lam, np.arange(1, m, dtype=np.complex128))  # Jordan trick, NEPTypes.jl:376-386
        return np.array([col[j] / math.factorial(j) for j in range(m)])


def as_function(f) -> ScalarFunction:
    return f if isinstance(f, ScalarFunction) else Callable(f)


ONE = Monomial(0)
IDENTITY = Monomial(1)
