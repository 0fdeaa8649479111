"""A few tall-skinny ZGEMMs and DGKS calls at config C5 size (what ncu wraps)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib, Block, block_gemm, dgks
nz = 945
n = (nz + 4) * nz + 2 * nz
m = 200
rng = np.random.default_rng(0)
Zb, Yb = Block(n, m + 1), Block(n, m + 1)
col = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(n)
for c in range(0, m + 1):
    Zb.upload(np.roll(col, c), c)
for k in (50, 100, 200):
    Cm = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
    for _ in range(2):
        block_gemm(Zb, 0, k, Cm, Yb, 0)
    for _ in range(2):
        dgks(Zb, k, Zb, k)
_lib.lib.nepb_synchronize()
