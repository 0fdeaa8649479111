"""Time the fused SpMM on config C4 under several environment settings (kernel variants), with a SciPy parity check.
usage: python tools/spmm_variants.py [grid] -- each variant is 'NAME=VAL,NAME=VAL' (or 'default'), from argv[2:]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib, Block, synthetic
from bench import build_c4, load_peaks

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
variants = sys.argv[2:] or ["default"]
dnep, mats, st = build_c4(grid)
lib = _lib.lib
peak, _ = load_peaks()
lam = 0.3 + 0.2j
coef = dnep.coefficients(lam)
Mo = sum(m * lam ** i for i, m in enumerate(mats)).tocsr()
ks = [int(x) for x in os.environ.get("KS", "8,20").split(",")]
for k in ks:
    V = synthetic.stencil_block(_lib.msws_state(1) if hasattr(_lib, "msws_state") else st, dnep.n, k)
    Vb, Zb = Block.from_host(V), Block(dnep.n, k)
    nbytes = dnep.apply_bytes(0, k, k)
    Zref = Mo @ V
    for var in variants:
        sets = [] if var == "default" else [kv.split("=") for kv in var.split(",")]
        for a, b in sets:
            os.environ[a] = b
        try:
            for _ in range(3):
                dnep.apply_block(0, Vb, coef, Zb)
            ms = C.c_float()
            lib.nepb_timer_start()
            for _ in range(20):
                dnep.apply_block(0, Vb, coef, Zb)
            lib.nepb_timer_stop(C.byref(ms))
            Z = Zb.download()
            err = np.linalg.norm(Z - Zref) / np.linalg.norm(Zref)
            t = ms.value / 20
            print("k=%2d %-40s %8.1f us  %7.0f GB/s  %5.1f%%  relerr %.1e" % (k, var, t * 1e3, nbytes / t / 1e6, 100 * nbytes / t / 1e6 / peak, err), flush=True)
        except Exception as e:
            print("k=%2d %-40s FAILED %s" % (k, var, e), flush=True)
        for a, b in sets:
            del os.environ[a]
    Vb.close(); Zb.close()
