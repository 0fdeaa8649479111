"""Host-buffer nepb_spmf_apply (pageable NumPy arrays in and out) on config C4 for k = 1, 8, 20: milliseconds per call."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib, synthetic
from bench import build_c4
dnep, mats, st = build_c4(1000)
coef = dnep.coefficients(0.3 + 0.2j)
for k in (1, 8, 20):
    V = synthetic.stencil_block(st, dnep.n, k)
    Z = np.empty((dnep.n, k), dtype=np.complex128, order="F")
    dnep.apply(_lib.COEF_SCALAR, V, coef, k, out=Z)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        dnep.apply(_lib.COEF_SCALAR, V, coef, k, out=Z)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("threads=%s k=%d: %s ms (min %.2f)" % (os.environ.get("NEPB_COPY_THREADS", "default"), k, " ".join("%.2f" % t for t in ts), min(ts)), flush=True)
