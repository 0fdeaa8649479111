"""On-device sweep of the fused SpMM row-ownership shapes (NEPB_SPMM_CFG=GC,GN,CPT,U) on config C4."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib, Block, synthetic
from bench import build_c4, load_peaks

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dnep, mats, st = build_c4(grid)
lib = _lib.lib
peak, _ = load_peaks()
coef = dnep.coefficients(0.3 + 0.2j)
CFGS = ["2,8,4,2", "4,4,2,4", "4,8,2,1", "8,4,1,2", "4,8,2,2", "2,16,4,1", "4,4,2,1", "8,4,1,1", "4,4,5,2", "4,8,5,1", "2,8,10,1", "4,2,5,1", "8,4,3,1", "8,2,3,1", "4,8,5,2",
        "4,4,2,2", "4,2,2,4", "4,2,5,2", "4,4,5,1"]
OLD = ["1,8,1,4", "1,4,1,6", "1,8,1,3", "1,16,1,2", "1,32,1,1", "1,4,1,8", "2,4,1,4", "4,2,1,4", "8,1,1,4", "8,2,1,4", "4,2,2,4", "4,4,2,2", "8,1,1,8",
        "4,2,4,2", "4,2,5,2", "4,1,5,4", "4,4,5,1", "4,2,5,4", "8,1,3,4", "8,2,3,2", "2,4,10,2", "8,1,4,4"]
for k in (8, 20):
    V = synthetic.stencil_block(_lib.msws_state(1), dnep.n, k)
    Vb, Zb = Block.from_host(V), Block(dnep.n, k)
    nbytes = dnep.apply_bytes(0, k, k)
    ref = None
    for cfg in CFGS:
        gc, gn, cpt, u = map(int, cfg.split(","))
        if gc * cpt < k:
            continue
        os.environ["NEPB_SPMM_CFG"] = cfg
        for _ in range(3):
            dnep.apply_block(0, Vb, coef, Zb)
        ms = C.c_float()
        lib.nepb_timer_start()
        for _ in range(20):
            dnep.apply_block(0, Vb, coef, Zb)
        lib.nepb_timer_stop(C.byref(ms))
        Z = Zb.download()
        if ref is None:
            ref = Z
        err = np.abs(Z - ref).max()
        t = ms.value / 20
        print("k=%2d cfg=%-9s %8.1f us  %7.0f GB/s  %5.1f%%  maxdiff %.1e" % (k, cfg, t * 1e3, nbytes / t / 1e6, 100 * nbytes / t / 1e6 / peak, err), flush=True)
    Vb.close(); Zb.close()
