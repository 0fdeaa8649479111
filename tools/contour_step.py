"""Run W warm-up + K timed contour steps (gun, N=128, k=20) -- the thing ncu wraps.  Usage: contour_step.py [W] [K] [batch] [nodes]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib
from bench import gun_operator, gun_probe, beyn_nodes, GUN_N, GUN_K, GUN_SIGMA, GUN_RADIUS
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
nodes = int(sys.argv[4]) if len(sys.argv) > 4 else GUN_N
lib = _lib.lib
dnep = gun_operator()
Vh = gun_probe(dnep.n, GUN_K)
lams, Wt = beyn_nodes(GUN_N, GUN_SIGMA, GUN_RADIUS)
lams, Wt = lams[:nodes], np.ascontiguousarray(Wt[:nodes])
integ = nepb200.ContourIntegrator(dnep, GUN_K, 2, min(batch, nodes))
coef = np.ascontiguousarray(np.stack([dnep.coefficients(l) for l in lams]))
Vf = _lib.as_c128_f(Vh)
_lib.check(lib.nepb_contour_set_probe(integ._h, _lib.ptr(Vf), dnep.n))
l0 = lib.nepb_launch_count()
for _ in range(W):
    _lib.check(lib.nepb_contour_integrate_dev(integ._h, nodes, _lib.ptr(coef), _lib.ptr(Wt), 0))
lib.nepb_synchronize()
l1 = lib.nepb_launch_count()
ms = C.c_float()
lib.nepb_timer_start()
for _ in range(K):
    _lib.check(lib.nepb_contour_integrate_dev(integ._h, nodes, _lib.ptr(coef), _lib.ptr(Wt), 0))
lib.nepb_timer_stop(C.byref(ms))
print("launches: setup+warmup %d, per step %d; %.2f ms/step (batch %d, %d nodes)" % (l1 - l0, (lib.nepb_launch_count() - l1) // max(K, 1), ms.value / max(K, 1), batch, nodes))
