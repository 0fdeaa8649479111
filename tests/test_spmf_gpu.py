"""Parity of the fused multi-term SpMM (through the C ABI) with the CPU oracle.  Runs on the B200.

Mirrors test/spmf.jl:26-63,94-178,266-283, test/core.jl:16-128 and test/spmf_stability.jl of the reference.
Floating-point tolerance: 1e-12 relative to ||Z|| (north star: 1e-10 on eigenresiduals); integer work bit-exact.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import nepb200
from nepb200 import B200SPMF, Monomial, Exp, PowShift, ONE, IDENTITY
from oracle import gallery as g
from oracle import nep as o

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def gun_pair():
    K, M, W1, W2 = g.load_gun_matrices()
    onep = o.nep_gallery("nlevp_native_gun")
    dnep = B200SPMF([K, -M, W1, W2], [ONE, IDENTITY, PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
    return onep, dnep


def test_union_pattern_bit_exact():
    K, M, W1, W2 = g.load_gun_matrices()
    aligned = o.form_aligned_sparsity_patterns([K, -M, W1, W2])
    dnep = B200SPMF([K, -M, W1, W2], [ONE] * 4)
    colptr, rowval = dnep.pattern()
    assert dnep.nnz_union == 148318
    assert np.array_equal(colptr, aligned[0].indptr) and np.array_equal(rowval, aligned[0].indices)
    rowptr, colind, perm = dnep.pattern_csr()
    R = sp.csc_matrix((np.arange(1, dnep.nnz_union + 1), rowval, colptr), shape=K.shape).tocsr()
    R.sort_indices()
    assert np.array_equal(rowptr, R.indptr) and np.array_equal(colind, R.indices)
    assert np.array_equal(perm, np.argsort(R.data))  # CSC position -> CSR position


def test_mder_gun():
    onep, dnep = gun_pair()
    for lam in (250.0 ** 2 + 1j, 22345.116783765 + 0.644998598j):
        Mo = sp.csc_matrix(o.compute_Mder(onep, lam))
        Md = dnep.compute_Mder(lam)
        assert abs(Mo - Md).max() <= 1e-13 * abs(Mo).max()
    M1o = sp.csc_matrix(o.compute_Mder(onep, 300.0 ** 2 + 2j, 1))
    M1d = dnep.compute_Mder(300.0 ** 2 + 2j, 1)
    assert abs(M1o - M1d).max() <= 1e-13 * abs(M1o).max()


@pytest.mark.parametrize("k", [1, 2, 3, 8, 20, 33])
def test_apply_M_gun(k):
    onep, dnep = gun_pair()
    rng = np.random.default_rng(k)
    V = rng.standard_normal((dnep.n, k)) + 1j * rng.standard_normal((dnep.n, k))
    lam = 250.0 ** 2 + 1j
    Z = dnep.compute_MM(lam * np.eye(k), V)
    Zo = sp.csc_matrix(o.compute_Mder(onep, lam)) @ V
    assert relerr(Z, Zo) < RTOL


def test_MM_diag_and_general_gun():
    onep, dnep = gun_pair()
    rng = np.random.default_rng(5)
    k = 6
    V = rng.standard_normal((dnep.n, k)) + 1j * rng.standard_normal((dnep.n, k))
    lams = 200.0 ** 2 + 1e4 * rng.standard_normal(k) + 1j * rng.standard_normal(k)
    Z = dnep.compute_MM(np.diag(lams), V)
    Zo = o.compute_MM(onep, np.diag(lams), V)
    assert relerr(Z, Zo) < RTOL
    r = dnep.residual_norms(lams, V)
    ro = np.array([np.linalg.norm(o.compute_Mlincomb(onep, lams[s], V[:, s])) / np.linalg.norm(V[:, s]) for s in range(k)])
    assert np.allclose(r, ro, rtol=1e-12)
    # non-diagonal S: general coefficient blocks f_i(S)
    S = np.diag(lams) + 10.0 * rng.standard_normal((k, k))
    Z = dnep.compute_MM(S, V)
    Zo = o.compute_MM(onep, S, V)
    assert relerr(Z, Zo) < 1e-10  # sqrtm of a non-normal S: host matrix-function accuracy dominates


@pytest.mark.parametrize("k", [1, 2, 5, 17, 40])
def test_mlincomb_gun(k):
    onep, dnep = gun_pair()
    rng = np.random.default_rng(10 + k)
    V = rng.standard_normal((dnep.n, k)) + 1j * rng.standard_normal((dnep.n, k))
    lam = 250.0 ** 2 + 3j
    gamma = 300.0 ** 2 - 200.0 ** 2
    a = gamma ** np.arange(k, dtype=float)
    z = dnep.compute_Mlincomb(lam, V if k > 1 else V[:, 0], a)
    zo = o.compute_Mlincomb_from_Mder(onep, lam, V, a)
    assert relerr(z, zo) < 1e-11
    if k > 2:  # iar's convention: a[0] = 0 (method_iar.jl:79) -> first column ignored
        a[0] = 0
        z = dnep.compute_Mlincomb(lam, V, a)
        zo = o.compute_Mlincomb_from_Mder(onep, lam, V, a)
        assert relerr(z, zo) < 1e-11


def test_mlincomb_does_not_modify_inputs():
    # test/spmf.jl:26-34
    onep, dnep = gun_pair()
    V = np.ones((dnep.n, 3), dtype=complex)
    a = np.array([1.0, 0.0, 2.0])
    V0, a0 = V.copy(), a.copy()
    dnep.compute_Mlincomb(1e4 + 1j, V, a)
    assert np.array_equal(V, V0) and np.array_equal(a, a0)


def test_dep0_dense_and_startder():
    # config C1 plumbing: dense 5x5 DEP through the same operator; test/core.jl:16-32
    A0, A1, tauv = g.dep0_matrices(5)
    onep = o.nep_gallery("dep0")
    dnep = B200SPMF.from_nep(nepb200.DEP([A0, A1], tauv))
    rng = np.random.default_rng(3)
    V = rng.standard_normal((5, 4))
    a = np.array([1.0, 2.0, 0.0, 0.5])
    lam = 0.3 - 0.2j
    assert relerr(dnep.compute_Mlincomb(lam, V, a), o.compute_Mlincomb(onep, lam, V, a)) < RTOL
    assert relerr(dnep.compute_Mlincomb(lam, V, a, startder=2), o.compute_Mlincomb(onep, lam, V, a, startder=2)) < RTOL
    assert abs(dnep.compute_Mder(3.0)[0, 0].real - (-2.942777908030041)) < 1e-15
    S = rng.standard_normal((4, 4))
    assert relerr(dnep.compute_MM(S, V), o.compute_MM(onep, S, V)) < RTOL
    z = dnep.compute_Mlincomb(1.0 + 1.0j, np.ones(5))
    assert relerr(z, o.compute_Mlincomb(onep, 1.0 + 1.0j, np.ones(5))) < RTOL


def test_dep0_100_literal():
    # src/Gallery.jl:174-176 through the device path
    A0, A1, tauv = g.dep0_matrices(100)
    dnep = B200SPMF.from_nep(nepb200.DEP([A0, A1], tauv))
    z = dnep.compute_Mlincomb(1.0 + 1.0j, np.ones(100))
    assert abs(np.linalg.norm(z) - 57.498446538064954) < 1e-12


def test_qdep0_sparse_spmf():
    A0, A1 = g.load_qdep0_matrices()
    n = A0.shape[0]
    onep = o.nep_gallery("qdep0")
    dnep = B200SPMF([-sp.identity(n, format="csc"), A0, A1], [Monomial(2), ONE, Exp(-1.0)])
    rng = np.random.default_rng(4)
    V = rng.standard_normal((n, 7)) + 1j * rng.standard_normal((n, 7))
    a = 1.0 / (1.0 + np.arange(7))
    lam = -1.0 + 0.1j
    assert relerr(dnep.compute_Mlincomb(lam, V, a), o.compute_Mlincomb(onep, lam, V, a)) < RTOL
    assert relerr(dnep.compute_MM(lam * np.eye(7), V), o.compute_MM(onep, lam * np.eye(7), V)) < RTOL


def test_complex_matrices_and_odd_term_counts():
    # test/spmf.jl:127-156 uses complex sparse A_i; also p = 1, 3, 5, 7 (generic kernel) and unequal patterns
    rng = np.random.default_rng(6)
    n = 257
    for p in (1, 3, 5, 7):
        for cplx in (False, True):
            A = []
            for i in range(p):
                M = sp.random(n, n, 0.03 + 0.01 * i, random_state=100 + i, format="csc")
                if cplx:
                    M = M + 1j * sp.random(n, n, 0.02, random_state=200 + i, format="csc")
                A.append(M.tocsc())
            fo = [o.f_pow(i) for i in range(p)]
            onep = o.SPMF_NEP(A, fo)
            dnep = B200SPMF(A, [Monomial(i) for i in range(p)])
            for k in (1, 4, 9):
                V = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
                lams = rng.standard_normal(k) + 1j * rng.standard_normal(k)
                assert relerr(dnep.compute_MM(np.diag(lams), V), o.compute_MM(onep, np.diag(lams), V)) < RTOL
                assert relerr(dnep.compute_MM(lams[0] * np.eye(k), V), o.compute_MM(onep, lams[0] * np.eye(k), V)) < RTOL
                a = rng.standard_normal(k)
                assert relerr(dnep.compute_Mlincomb(lams[0], V, a), o.compute_Mlincomb_from_Mder(onep, lams[0], V, a)) < RTOL


def test_stencil_pep_small_and_device_blocks():
    # config C4 generator at g=64 against the oracle, operands resident in HBM
    from nepb200 import Block, _lib
    mats, rng_state = g.stencil_pep(64)
    n = 64 * 64
    onep = o.PEP([m.tocsc() for m in mats])
    dnep = B200SPMF([m.tocsc() for m in mats], [Monomial(i) for i in range(4)])
    assert dnep.nnz_union == mats[0].nnz
    lam = 0.3 + 0.2j
    for k in (1, 8, 20):
        V = g.stencil_block(g.MSWS_RNG(7), n, k)
        Vb, Zb = Block.from_host(V), Block(n, k)
        dnep.apply_block(_lib.COEF_SCALAR, Vb, dnep.coefficients(lam), Zb)
        Z = Zb.download()
        Zo = sp.csc_matrix(o.compute_Mder(onep, lam)) @ V
        assert relerr(Z, Zo) < RTOL
        assert np.array_equal(Vb.download(), V)  # layout round trip is exact


def test_ragged_and_empty_rows():
    n = 50
    rows = np.array([0, 0, 0, 7, 49, 49])
    cols = np.array([0, 10, 49, 7, 0, 49])
    A = sp.csc_matrix((np.arange(1.0, 7.0), (rows, cols)), shape=(n, n))
    B = sp.csc_matrix(([2.0], ([3], [3])), shape=(n, n))
    dnep = B200SPMF([A, B], [ONE, IDENTITY])
    V = np.arange(n * 2, dtype=float).reshape(n, 2) + 1j
    Z = dnep.compute_MM(2.0 * np.eye(2), V)
    assert relerr(Z, (A + 2.0 * B) @ V) < RTOL


def test_error_behaviour():
    A = sp.identity(4, format="csc")
    with pytest.raises(ValueError):
        B200SPMF([A, A], [ONE])  # matrices / functions mismatch (NEPTypes.jl:197-199)
    d = B200SPMF([A], [ONE])
    with pytest.raises(ValueError):
        d.compute_MM(np.eye(2), np.ones((5, 2)))
    with pytest.raises(nepb200.NepbError):
        d.apply(0, np.ones((4, 2)), np.ones(1), 3)  # q != k in SCALAR mode


def test_row_tiles_of_the_multicolumn_kernel():
    """The tiled SpMM (V rows staged in shared memory) depends on host integer work: tiles of <= 32 consecutive rows with
    <= 192 distinct columns covering every row exactly once.  Checked through its invariants and, end to end, against the
    oracle on a matrix whose tiles close early (many distinct columns per row), one with empty rows, and one with a row too
    wide for a tile (which must fall back to the untiled kernels)."""
    mats, _ = g.stencil_pep(64)
    dnep = B200SPMF([m.tocsc() for m in mats], [Monomial(i) for i in range(4)])
    ntiles, distinct, mx = dnep.tiles_info()
    n = 64 * 64
    assert ntiles == n // 32 and mx <= 192
    assert distinct < dnep.nnz_union / 3  # the stencil shares most columns between neighbouring rows
    rng = np.random.default_rng(11)
    # ~40 random columns per row: 4-5 rows per tile; rows 100..163 empty
    n2 = 2000
    A = sp.random(n2, n2, 0.02, random_state=5, format="lil")
    A[100:164, :] = 0
    A = A.tocsc()
    B = sp.random(n2, n2, 0.004, random_state=6, format="csc")
    d2 = B200SPMF([A, B], [ONE, IDENTITY])
    nt2, dist2, mx2 = d2.tiles_info()
    assert nt2 > n2 // 32 and 0 < mx2 <= 192 and dist2 <= d2.nnz_union
    for k in (5, 8, 20, 32):
        V = rng.standard_normal((n2, k)) + 1j * rng.standard_normal((n2, k))
        lam = 0.7 - 0.2j
        assert relerr(d2.compute_MM(lam * np.eye(k), V), (A + lam * B) @ V) < RTOL
        lams = rng.standard_normal(k) + 1j * rng.standard_normal(k)
        assert relerr(d2.compute_MM(np.diag(lams), V), A @ V + (B @ V) * lams[None, :]) < RTOL
    # wide blocks use the untiled kernels by default; NEPB_SPMM_TILE_ROWS forces the tiled kernel (16- or 32-row tiles)
    import os
    for rows in ("16", "32"):
        os.environ["NEPB_SPMM_TILE_ROWS"] = rows
        try:
            for k in (8, 20, 32):
                V = rng.standard_normal((n2, k)) + 1j * rng.standard_normal((n2, k))
                assert relerr(d2.compute_MM((0.1 + 0.3j) * np.eye(k), V), (A + (0.1 + 0.3j) * B) @ V) < RTOL
        finally:
            del os.environ["NEPB_SPMM_TILE_ROWS"]
    # the other kernel variants behind the same call: the round-1 cp.async tiled kernel, 4 lanes per row (swizzled reads)
    for env in ({"NEPB_SPMM_TMA": "0"}, {"NEPB_SPMM_GC": "4"}, {"NEPB_SPMM_GC": "4", "NEPB_SPMM_TILE_ROWS": "32"}):
        os.environ.update(env)
        try:
            for k in (5, 8, 12, 20, 24, 32):
                V = rng.standard_normal((n2, k)) + 1j * rng.standard_normal((n2, k))
                assert relerr(d2.compute_MM((0.1 + 0.3j) * np.eye(k), V), (A + (0.1 + 0.3j) * B) @ V) < RTOL
                lams = rng.standard_normal(k) + 1j * rng.standard_normal(k)
                assert relerr(d2.compute_MM(np.diag(lams), V), A @ V + (B @ V) * lams[None, :]) < RTOL
        finally:
            for key in env:
                del os.environ[key]
    # one dense row: no tiling possible
    C = sp.lil_matrix((400, 400))
    C[7, :] = 1.0
    C.setdiag(2.0)
    d3 = B200SPMF([C.tocsc()], [ONE])
    assert d3.tiles_info() == (0, 0, 0)
    V = rng.standard_normal((400, 8)) + 0j
    assert relerr(d3.compute_MM(np.eye(8), V), C.tocsc() @ V) < RTOL


@pytest.mark.parametrize("k,q", [(40, 40), (12, 12), (33, 1), (7, 45), (45, 7), (64, 64), (101, 3)])
def test_general_mode_square_and_rectangular_blocks(k, q):
    """Z = sum_i A_i (V C_i) with dense k x q blocks straight through nepb_spmf_apply(COEF_GENERAL): the shapes of
    infbilanczos' Hankel blocks (q = k, p*q > 64; method_infbilanczos.jl:229-244), of Proj_SPMF_NEP's selector blocks
    (q = p*k; NEPTypes.jl:724-790) and of compute_Mlincomb at depth 101 (q = 1), against NumPy on gun (p = 4) and qdep0 (p = 3)."""
    import scipy.sparse as sp
    from nepb200 import _lib
    rng = np.random.default_rng(k * 100 + q)
    K, M, W1, W2 = g.load_gun_matrices()
    A0, A1 = g.load_qdep0_matrices()
    for mats in ([K, -M, W1, W2], [-sp.identity(A0.shape[0], format="csc"), A0, A1]):
        d = B200SPMF(mats, [ONE] * len(mats))
        n = d.n
        V = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
        Cs = [rng.standard_normal((k, q)) + 1j * rng.standard_normal((k, q)) for _ in mats]
        Cblk = np.stack([np.asfortranarray(c).T.copy() for c in Cs])  # p blocks, each column-major k x q
        Z = d.apply(_lib.COEF_GENERAL, V, Cblk, q)
        Zr = sum(m @ (V @ c) for m, c in zip(mats, Cs))
        assert relerr(Z, Zr) < 1e-13


def test_host_buffer_pipeline_roundtrip():
    """csrc/hostcopy.cu: column-major host arrays (pageable, with a leading dimension larger than n, and page-locked through
    nepb_host_register) through the pinned three-slot ring into row-major device blocks and back, bit for bit; sizes chosen so
    that several chunks (and several slot reuses) occur."""
    from nepb200 import Block, _lib
    rng = np.random.default_rng(17)
    for n, k in ((300_000, 23), (70_001, 9), (1_200_000, 3), (33, 5)):
        ld = n + 7
        buf = np.zeros((ld, k), dtype=np.complex128, order="F")
        buf[:n, :] = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
        b = Block(n, k + 2)
        _lib.check(_lib.lib.nepb_block_upload(b._h, 1, k, _lib.ptr(buf), ld))
        out = np.full((ld, k), np.nan + 0j, dtype=np.complex128, order="F")
        _lib.check(_lib.lib.nepb_block_download(b._h, 1, k, _lib.ptr(out), ld))
        assert np.array_equal(out[:n], buf[:n]) and np.all(np.isnan(out[n:].real))
        assert np.array_equal(b.download(1, k), buf[:n])
        # registered (page-locked) host memory: direct DMA
        _lib.check(_lib.lib.nepb_host_register(_lib.ptr(buf), buf.nbytes))
        _lib.check(_lib.lib.nepb_host_register(_lib.ptr(out), out.nbytes))
        try:
            buf[:n, :] *= 2.0
            out[:] = 0
            _lib.check(_lib.lib.nepb_block_upload(b._h, 2, k, _lib.ptr(buf), ld))
            _lib.check(_lib.lib.nepb_block_download(b._h, 2, k, _lib.ptr(out), ld))
            assert np.array_equal(out[:n], buf[:n]) and np.all(out[n:] == 0)
        finally:
            _lib.check(_lib.lib.nepb_host_unregister(_lib.ptr(buf)))
            _lib.check(_lib.lib.nepb_host_unregister(_lib.ptr(out)))
        b.close()


def test_two_dimensional_tiles_of_the_multicolumn_kernel():
    """The 2D tiles (S segments x R rows, segments one grid line apart; csrc/spmf.cu:spmf_build_tiles2d) depend on host integer
    work: the line length is found from the pattern, every row belongs to exactly one tile, the remainder rows and the short
    last segments of a line are handled.  Grid sides that are / are not multiples of the segment height, with / without a
    remainder of rows behind the last super-block; every width class of the kernel (1..4 pieces of 8 columns); against the
    oracle's products.  A pattern without a dominant line (random, gun) keeps the 1D tiles."""
    import os
    rng = np.random.default_rng(21)
    for grid in (37, 40, 50):
        mats, _ = g.stencil_pep(grid)
        Av = [m.tocsc() for m in mats]
        dnep = B200SPMF(Av, [Monomial(i) for i in range(4)])
        line, S, R, ntiles, staged = dnep.tiles2d_info()
        n = grid * grid
        assert (line, S, R) == (grid, 4, 8)
        nsb = n // (4 * grid)
        expect = nsb * -(-grid // 8) + -(-(n - nsb * 4 * grid) // 32)
        assert ntiles == expect
        _, staged1d, _ = dnep.tiles_info()
        assert staged < staged1d  # fewer staged rows of V than the 32-row strips
        lam = 0.3 + 0.2j
        Mo = sum(A * lam ** i for i, A in enumerate(Av))
        for k in (5, 8, 9, 16, 17, 20, 24, 25, 32):
            V = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
            Z = dnep.compute_MM(lam * np.eye(k), V)
            assert relerr(Z, Mo @ V) < RTOL
            os.environ["NEPB_SPMM_2D"] = "0"   # the 1D tiles give the same product
            try:
                assert relerr(dnep.compute_MM(lam * np.eye(k), V), Mo @ V) < RTOL
            finally:
                del os.environ["NEPB_SPMM_2D"]
            assert np.array_equal(Z, dnep.compute_MM(lam * np.eye(k), V))  # bitwise reproducible
        # a column window of a wider block (rows of V not adjacent in memory: one bulk copy per staged row)
        from nepb200 import Block
        Vw = rng.standard_normal((n, 24)) + 1j * rng.standard_normal((n, 24))
        Vb, Zb = Block.from_host(Vw), Block(n, 24)
        from nepb200 import _lib
        cf = np.ascontiguousarray(dnep.coefficients(lam), dtype=np.complex128)
        _lib.check(_lib.lib.nepb_spmf_apply_block_ex(dnep._h, 0, Vb._h, 3, 12, 12, _lib.ptr(cf), Zb._h, 5))
        assert relerr(Zb.download()[:, 5:17], Mo @ Vw[:, 3:15]) < RTOL
    A = sp.random(2000, 2000, 0.01, random_state=3, format="csc")
    assert B200SPMF([A], [ONE]).tiles2d_info()[0] == 0
    K, M, W1, W2 = g.load_gun_matrices()
    assert B200SPMF([K, -M, W1, W2], [ONE, IDENTITY, ONE, ONE]).tiles2d_info()[0] == 0
