"""Host side of the product's WEP module (nepb200/wep.py) against the oracle: discretisation, wavenumbers, the Gegenbauer
derivative table, the SPMF format -- everything that runs without a GPU.  The device kernels are covered by tests/test_wep_gpu.py."""
import numpy as np
import pytest

from nepb200 import wep as pw
from oracle import wep as ow


@pytest.mark.parametrize("wg", ["TAUSCH", "JARLEBRING"])
@pytest.mark.parametrize("nx,nz", [(11, 7), (109, 105), (20, 9)])
def test_discretisation_is_identical(wg, nx, nz):
    K, hx, hz, Km, Kp = pw.generate_wavenumber_fd(nx, nz, wg, 0.1)
    Ko, hxo, hzo, Kmo, Kpo = ow.generate_wavenumber_fd(nx, nz, wg, 0.1)
    assert abs(K - Ko).max() == 0 and abs(hx - hxo) < 1e-16 and (hz, Km, Kp) == (hzo, Kmo, Kpo)
    for a, b in zip(pw.generate_fd_interior_mat(nx, nz, hx, hz), ow.generate_fd_interior_mat(nx, nz, hx, hz)):
        assert abs(a - b).max() == 0
    for a, b in zip(pw.generate_fd_boundary_mat(nx, nz, hx, hz), ow.generate_fd_boundary_mat(nx, nz, hx, hz)):
        assert abs(a - b).max() == 0


def test_unknown_waveguide_and_format_raise():
    with pytest.raises(ValueError):
        pw.generate_wavenumber_fd(11, 7, "NOPE", 0.1)
    with pytest.raises(ValueError):
        pw.nep_gallery_WEP(nx=11, nz=7, neptype="XYZ")


def test_sqrt_derivative_table_and_spmf_format():
    lam = -1.3 - 0.31j
    pairs = ((0.3 + 2j, 2.0 - 1j), (1.0, -3.0), (-25.1j, 7.5))
    d = pw.sqrt_derivative(1.0, [p[0] for p in pairs], [p[1] for p in pairs], 7, lam)
    for i, (b, c) in enumerate(pairs):
        assert np.allclose(d[i], ow.sqrt_derivative(1, b, c, 7, lam), rtol=1e-14, atol=0)
    assert np.allclose(pw.sqrt_derivative(1.0, [1.0], [2.0], 0, lam)[0, 0], ow.sqrt_derivative(1, 1.0, 2.0, 0, lam))
    nep = pw.nep_gallery_WEP(nx=11, nz=7, neptype="SPMF")
    A, f = ow.nep_gallery_wep(nx=11, nz=7, neptype="SPMF")
    assert len(nep.A) == len(A) == 17
    for X, Y in zip(nep.A, A):
        assert abs(X - Y).max() < 1e-15
    for g, h in zip(nep.fi, f):
        assert abs(g(lam) - h(lam)) < 1e-14
    # Taylor coefficients of the boundary functions (used when the SPMF format runs through compute_Mlincomb with k > 1)
    fj = nep.fi[5]
    t = fj.taylor(lam, 4)
    h = 1e-4
    assert abs(t[1] - (fj(lam + h) - fj(lam - h)) / (2 * h)) < 1e-7 * abs(t[1])
    assert abs(2 * t[2] - (fj(lam + h) - 2 * fj(lam) + fj(lam - h)) / h ** 2) < 1e-5 * abs(t[2])
