"""GENERAL-mode products on config C4 (compute_Mlincomb with k = 20 / 100 columns, q = 1): what ncu wraps / CUDA-event timing."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib, Block, synthetic
from bench import build_c4, load_peaks
lib = _lib.lib
dnep, mats, st = build_c4(1000)
peak, _ = load_peaks()
rng = np.random.default_rng(0)
csr = [m.tocsr() for m in mats]
for k, q in ((20, 1), (100, 1), (7, 1), (8, 8)):
    V = synthetic.stencil_block(st, dnep.n, k)
    Vb, Zb = Block.from_host(V), Block(dnep.n, q)
    Cs = [rng.standard_normal((k, q)) + 1j * rng.standard_normal((k, q)) for _ in range(dnep.p)]
    Cblk = np.ascontiguousarray(np.stack([np.asfortranarray(c).T.copy() for c in Cs]))
    for _ in range(3):
        dnep.apply_block(_lib.COEF_GENERAL, Vb, Cblk, Zb)
    ms = C.c_float(); lib.nepb_timer_start()
    for _ in range(20):
        dnep.apply_block(_lib.COEF_GENERAL, Vb, Cblk, Zb)
    lib.nepb_timer_stop(C.byref(ms))
    t = ms.value / 20
    nbytes = dnep.apply_bytes(_lib.COEF_GENERAL, k, q)
    Zref = sum(m @ (V @ c) for m, c in zip(csr, Cs))
    err = np.linalg.norm(Zb.download() - Zref) / np.linalg.norm(Zref)
    print("general k=%d q=%d: %.1f us  %.0f GB/s  %.1f%%  relerr %.1e" % (k, q, t * 1e3, nbytes / t / 1e6, 100 * nbytes / t / 1e6 / peak, err), flush=True)
    Vb.close(); Zb.close()
