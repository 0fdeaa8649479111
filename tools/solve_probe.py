"""One factorisation of gun and a few device-resident 20-column solves (what ncu wraps to time the solve kernels)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ctypes as C
import numpy as np
import nepb200
from nepb200 import _lib, Block
from bench import gun_operator
lib = _lib.lib
dnep = gun_operator()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lu = nepb200.B200LU(dnep, [250.0 ** 2 + 1j])
B = Block.from_host(np.ones((dnep.n, k), dtype=complex)); X = Block(dnep.n, k)
from nepb200.dense import solve_block
for _ in range(3):
    solve_block(lu, B, 0, k, X, 0)
lib.nepb_synchronize()
ms = C.c_float(); lib.nepb_timer_start()
for _ in range(20):
    solve_block(lu, B, 0, k, X, 0)
lib.nepb_timer_stop(C.byref(ms))
print("device-resident solve k=%d: %.3f ms" % (k, ms.value / 20))
