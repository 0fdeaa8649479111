"""Oracle: gallery problems and matrix I/O (test infrastructure, see oracle/__init__.py).

Restates, in NumPy/SciPy:
  * the release-stable Middle-Square-Weyl-Sequence generator and the random-matrix helpers
    of src/gallery_extra/basic_random_examples.jl:71-128,
  * nep_gallery("dep0") (basic_random_examples.jl:1-9),
  * the sparse-matrix text format of src/utils/Serialization.jl:8-31,
  * nep_gallery("nlevp_native_gun") (src/gallery_extra/NLEVP_native.jl:4-18) and
    nep_gallery("qdep0") (src/gallery_extra/gallery_examples.jl:75-88),
  * the synthetic degree-3 stencil PEP of SURVEY.md section 8(d) (config C4) -- this one is our own
    benchmark generator, there is no reference counterpart.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp

_M128 = (1 << 128) - 1
_M64 = (1 << 64) - 1


class MSWS_RNG:
    """basic_random_examples.jl:73-84 -- UInt128 state, wrap-around arithmetic."""

    def __init__(self, seed: int = 0):
        base = 0x9EF09A97AC0F9ECAEF01C4F2DB0958C9
        self.s = ((seed << 1) + base) & _M128
        self.x = 0x1DE568E1A1CA1B593CBF13F7407CF43E
        self.w = 0xD4AC5C288559E14A5FAFC1B7DF9F9E0E


def gen_rng_int(rng: MSWS_RNG) -> int:
    """basic_random_examples.jl:86-91."""
    rng.x = (rng.x * rng.x) & _M128
    rng.w = (rng.w + rng.s) & _M128
    rng.x = (rng.x + rng.w) & _M128
    rng.x = ((rng.x >> 64) | (rng.x << 64)) & _M128
    return rng.x & _M64


def gen_rng_float(rng: MSWS_RNG) -> float:
    """basic_random_examples.jl:93-95: Float64(UInt64 / typemax(UInt64)); Julia converts both
    operands to Float64 first, and Float64(typemax(UInt64)) == 2.0^64."""
    return float(gen_rng_int(rng)) / 18446744073709551616.0


def gen_rng_mat(rng: MSWS_RNG, n: int, m: int) -> np.ndarray:
    """basic_random_examples.jl:97-105 -- column-major fill with 1-2u."""
    A = np.zeros((n, m))
    for c in range(m):
        for r in range(n):
            A[r, c] = 1 - 2 * gen_rng_float(rng)
    return A


def dep0_matrices(n: int = 5):
    """basic_random_examples.jl:2-9 -> (A0, A1, tauv)."""
    rng = MSWS_RNG()
    A0 = gen_rng_mat(rng, n, n)
    A1 = gen_rng_mat(rng, n, n)
    return A0, A1, np.array([0.0, 1.0])


def dep0_tridiag_matrices(n: int = 100):
    """basic_random_examples.jl:24-32: sparse tridiagonal DEP from the MSWS stream -> (A0, A1, tauv)."""
    rng = MSWS_RNG()
    K = np.concatenate([np.arange(n), np.arange(1, n), np.arange(n - 1)])
    J = np.concatenate([np.arange(n), np.arange(n - 1), np.arange(1, n)])
    A0 = sp.csc_matrix((gen_rng_mat(rng, 3 * n - 2, 1).ravel(), (K, J)), shape=(n, n))
    A1 = sp.csc_matrix((gen_rng_mat(rng, 3 * n - 2, 1).ravel(), (K, J)), shape=(n, n))
    return A0, A1, np.array([0.0, 1.0])


def neuron0_matrices():
    """gallery_examples.jl:122-139: the 2 x 2 neuron DEP with four delays (trivial stationary solution)."""
    kappa, beta, a12, a21 = 0.5, -1.0, 1.0, 2.34
    tauv = np.array([0.0, 0.2, 0.2, 1.5])
    A0 = -kappa * np.eye(2)
    A1 = a21 * np.array([[0.0, 0.0], [1.0, 0.0]])
    A2 = a12 * np.array([[0.0, 1.0], [0.0, 0.0]])
    A3 = beta * np.eye(2)
    return [A0, A1, A2, A3], tauv


def read_sparse_matrix(filename: str) -> sp.csc_matrix:
    """utils/Serialization.jl:19-31: line1 m, line2 n, then c row indices, c column indices,
    c values (1-based); Julia's sparse(I,J,V,m,n) sums duplicates and keeps explicit zeros."""
    with open(filename) as f:
        data = f.read().split()
    m = int(data[0])
    n = int(data[1])
    c = (len(data) - 2) // 3
    I = np.array(data[2:2 + c], dtype=np.int64) - 1
    J = np.array(data[2 + c:2 + 2 * c], dtype=np.int64) - 1
    V = np.array(data[2 + 2 * c:2 + 3 * c], dtype=np.float64)
    A = sp.coo_matrix((V, (I, J)), shape=(m, n)).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    return A


_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_gun_matrices():
    """K, M, W1, W2 of the gun problem as CSC float64 (NLEVP_native.jl:4-11).

    Reads the committed binary fixture tests/golden/gun.npz (made from the reference's text files by
    tests/golden/make_fixtures.py, because /root/reference does not exist on the GPU box)."""
    z = np.load(os.path.join(_GOLDEN, "gun.npz"))
    out = []
    for name in ("K", "M", "W1", "W2"):
        n = int(z["n"])
        out.append(sp.csc_matrix((z[name + "_data"], z[name + "_indices"], z[name + "_indptr"]), shape=(n, n)))
    return out


def load_qdep0_matrices():
    """A0, A1 of qdep0 (gallery_examples.jl:75-88), from tests/golden/qdep0.npz."""
    z = np.load(os.path.join(_GOLDEN, "qdep0.npz"))
    n = int(z["n"])
    return [sp.csc_matrix((z[k + "_data"], z[k + "_indices"], z[k + "_indptr"]), shape=(n, n)) for k in ("A0", "A1")]


# ----------------------------------------------------------------------------------------------
# Synthetic config C4: degree-3 PEP on a 2-D grid, 21-point stencil (our own generator)
# ----------------------------------------------------------------------------------------------
STENCIL = [(di, dj) for di in range(-2, 3) for dj in range(-2, 3) if not (abs(di) == 2 and abs(dj) == 2)]


def stencil_pattern(g: int):
    """CSR pattern (indptr, indices) of the g x g grid, row-major node numbering r = i*g + j, no
    wrap-around; per row the columns are visited in STENCIL order, which is ascending column order."""
    indptr = np.zeros(g * g + 1, dtype=np.int64)
    indices = []
    kinds = []  # 0 diagonal, 1 direct neighbour, 2 other
    for i in range(g):
        for j in range(g):
            for (di, dj) in STENCIL:
                ii, jj = i + di, j + dj
                if 0 <= ii < g and 0 <= jj < g:
                    indices.append(ii * g + jj)
                    kinds.append(0 if (di == 0 and dj == 0) else (1 if abs(di) + abs(dj) == 1 else 2))
            indptr[i * g + j + 1] = len(indices)
    return indptr, np.array(indices, dtype=np.int64), np.array(kinds, dtype=np.int8)


def stencil_pep(g: int, seed: int = 0):
    """A0..A3 (CSR, shared pattern) of the synthetic PEP: one MSWS stream, drawn term after term in CSR
    order: A0 = -(5-point Laplacian) + 0.1*u, A_i = s_i*(1-2u), s = 1, 0.1, 0.01."""
    indptr, indices, kinds = stencil_pattern(g)
    nnz = len(indices)
    rng = MSWS_RNG(seed)
    lap = np.where(kinds == 0, 4.0, np.where(kinds == 1, -1.0, 0.0))
    vals = []
    u = np.array([gen_rng_float(rng) for _ in range(nnz)])
    vals.append(lap + 0.1 * u)
    for s in (1.0, 0.1, 0.01):
        u = np.array([gen_rng_float(rng) for _ in range(nnz)])
        vals.append(s * (1 - 2 * u))
    n = g * g
    return [sp.csr_matrix((v, indices.copy(), indptr.copy()), shape=(n, n)) for v in vals], rng


def stencil_block(rng: MSWS_RNG, n: int, k: int) -> np.ndarray:
    """Complex n x k block, column-major draw order, re then im per entry, each 1-2u."""
    V = np.zeros((n, k), dtype=np.complex128)
    for c in range(k):
        for r in range(n):
            re = 1 - 2 * gen_rng_float(rng)
            im = 1 - 2 * gen_rng_float(rng)
            V[r, c] = complex(re, im)
    return V


def dep_symm_double_matrices(n=100):
    """nep_gallery("dep_symm_double", n) (src/gallery_extra/gallery_examples.jl:13-30): symmetric DEP of size n^2 with double
    eigenvalues, A = kron(L, L) + diag(8 sin x sin y), B = diag(-100 |sin(x + y)|), delays (0, 2)."""
    import scipy.sparse as sp
    LL = -sp.diags(2 * np.ones(n)) + sp.diags(np.ones(n - 1), -1) + sp.diags(np.ones(n - 1), 1)
    x = np.linspace(0, np.pi, n)
    h = x[1] - x[0]
    LL = LL / h ** 2
    LL = sp.kron(LL, LL)
    b = -100 * np.abs(np.sin(x[:, None] + x[None, :]))
    a = 8 * np.sin(x[:, None]) * np.sin(x[None, :])
    B = sp.diags(b.reshape(-1, order="F"))
    A = LL + sp.diags(a.reshape(-1, order="F"))
    return sp.csc_matrix(A), sp.csc_matrix(B), [0.0, 2.0]
