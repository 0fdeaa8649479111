"""Quick device probe of the LU: correctness + timing of factor / solve for gun at several batch sizes."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, scipy.sparse as sp
import nepb200
from nepb200 import _lib, B200SPMF, ONE, IDENTITY, PowShift
from oracle import gallery as g, nep as o
lib = _lib.lib
K, M, W1, W2 = g.load_gun_matrices()
dnep = B200SPMF([K, -M, W1, W2], [ONE, IDENTITY, PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
onep = o.nep_gallery("nlevp_native_gun")
t0 = time.time(); info = nepb200.symbolic_info(dnep); print("symbolic", info, "%.2fs" % (time.time() - t0))
lam = 250.0 ** 2 + 1j
Mo = sp.csc_matrix(o.compute_Mder(onep, lam))
b = np.ones((dnep.n, 20), dtype=complex)
for nb in (1, 4, 16, 32):
    lams = lam + 100.0 * np.arange(nb)
    lu = nepb200.B200LU(dnep, lams)  # warm
    lib.nepb_synchronize(); t0 = time.perf_counter()
    lu2 = nepb200.B200LU(dnep, lams); lib.nepb_synchronize(); tf = time.perf_counter() - t0
    x = lu2.solve(b, 0)
    t0 = time.perf_counter(); x = lu2.solve(b, 0); ts = time.perf_counter() - t0
    r = np.linalg.norm(Mo @ x - b) / (abs(Mo).sum(axis=0).max() * np.linalg.norm(x))
    print("nb=%d factor %.2f ms (%.2f ms/shift, %.1f GFLOP/s real)  host solve k=20: %.2f ms  scaled resid %.2e status %s" %
          (nb, tf * 1e3, tf * 1e3 / nb, 8 * info["flops"] * nb / tf / 1e9, ts * 1e3, r, lu2.status(0)), flush=True)
    lu.close(); lu2.close()
