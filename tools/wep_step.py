"""WEP compute_Mlincomb at config C5 size (nz = 945, nx = 949) for ncu: W warm-up + K timed calls with na columns.
Usage: wep_step.py [na] [W] [K]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import Block, _lib
na = int(sys.argv[1]) if len(sys.argv) > 1 else 1
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = _lib.lib
nep = nepb200.nep_gallery_WEP(nx=949, nz=945, benchmark_problem="JARLEBRING", neptype="WEP")
rng = np.random.default_rng(0)
V = rng.standard_normal((nep.n, na)) + 1j * rng.standard_normal((nep.n, na))
a = np.ones(na, dtype=np.complex128)
Vb, Zb = Block.from_host(V), Block(nep.n, 1)
lam = -2.7 - 3.1j
for _ in range(W):
    nep.mlincomb_block(lam, Vb, 0, na, a, Zb, 0)
lib.nepb_synchronize()
ms = C.c_float()
lib.nepb_timer_start()
for _ in range(K):
    nep.mlincomb_block(lam, Vb, 0, na, a, Zb, 0)
lib.nepb_timer_stop(C.byref(ms))
print("na=%d: %.1f us per call (CUDA events around %d calls issued from Python)" % (na, ms.value / K * 1e3, K))
