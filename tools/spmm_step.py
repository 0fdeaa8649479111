"""One fused SpMM launch per k on config C4 (what ncu wraps for the roofline `traffic` figure)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import nepb200
from nepb200 import _lib, Block, synthetic
from bench import build_c4
dnep, mats, st = build_c4(1000)
coef = dnep.coefficients(0.3 + 0.2j)
for k in (1, 8, 20):
    V = synthetic.stencil_block(st, dnep.n, k)
    Vb, Zb = Block.from_host(V), Block(dnep.n, k)
    for _ in range(3):
        dnep.apply_block(_lib.COEF_SCALAR, Vb, coef, Zb)
    _lib.lib.nepb_synchronize()
