"""qdep0 at sigma = 0 (test/infbilanczos.jl configuration): device LU status / residuals for the operator and its transpose,
next to SuperLU, plus the isolated GENERAL-mode product with square Hankel-like blocks."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
import nepb200
from nepb200 import B200SPMF, Monomial, ONE, Exp
from oracle import gallery as g
A0, A1 = g.load_qdep0_matrices()
n = A0.shape[0]
mI = -sp.identity(n, format="csc")
fi = [Monomial(2), ONE, Exp(-1.0)]
rng = np.random.default_rng(1)
for name, mats in (("nep", [mI, A0, A1]), ("nept", [mI, sp.csc_matrix(A0.T), sp.csc_matrix(A1.T)])):
    d = B200SPMF(mats, fi)
    for lam in (0.0, -1 + 0.2j):
        Mo = sp.csc_matrix(-lam ** 2 * sp.identity(n) + mats[1] + np.exp(-lam) * mats[2]).astype(complex)
        b = rng.standard_normal((n, 3)) + 0j
        xs = spl.splu(Mo).solve(b)
        lu = nepb200.B200LU(d, [lam])
        for r in (0, 2, 10):
            try:
                x = lu.solve(b, 0, r, want_berr=True)
                print(name, lam, "refine", r, "status", lu.status(0), "berr %.2e" % lu.last_berr,
                      "relres %.2e" % (np.linalg.norm(Mo @ x - b) / np.linalg.norm(b)),
                      "vs superlu %.2e" % (np.linalg.norm(x - xs) / np.linalg.norm(xs)), flush=True)
            except Exception as e:
                print(name, lam, "refine", r, "EXC", e, flush=True)
        lu.close()
    # GENERAL mode with square blocks
    for k in (1, 4, 12, 40):
        V = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
        Cs = [rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k)) for _ in mats]
        Cblk = np.stack([np.asfortranarray(c).T.copy() for c in Cs])  # p blocks, each column-major
        Z = d.apply(nepb200._lib.COEF_GENERAL, V, Cblk, k)
        Zr = sum(m @ (V @ c) for m, c in zip(mats, Cs))
        print(name, "general k=q=%d relerr %.2e" % (k, np.linalg.norm(Z - Zr) / np.linalg.norm(Zr)), flush=True)
