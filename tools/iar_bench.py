"""Config C2 (BASELINE.json): gun SPMF, iar m=100 on one B200, everything resident in HBM; tiar beside it.
Usage: iar_bench.py [m] [cpu_m]  -- cpu_m > 0 also times the oracle (CPU restatement) at that depth."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib
from bench import gun_operator
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cpu_m = int(sys.argv[2]) if len(sys.argv) > 2 else 0
gamma = float(sys.argv[3]) if len(sys.argv) > 3 else 300.0 ** 2 - 200.0 ** 2  # gamma^m must stay below 1e308 (as in the reference)
dnep = gun_operator()
n = dnep.n
kw = dict(sigma=250.0 ** 2, gamma=gamma, neigs=np.inf, v=np.ones(n), tol=1e-10)
for name, fn in (("iar_device", nepb200.iar_device), ("tiar_device", nepb200.tiar_device)):
    fn(dnep, maxit=5, check_error_every=5, **kw)  # warm-up (symbolic analysis, allocations)
    _lib.lib.nepb_synchronize()
    l0 = _lib.lib.nepb_launch_count()
    t0 = time.perf_counter()
    out = fn(dnep, maxit=m, check_error_every=m, **kw)
    _lib.lib.nepb_synchronize()
    dt = time.perf_counter() - t0
    lam = out[0]
    res = dnep.residual_norms(lam, out[1]) if len(lam) else []
    print("gamma=%g" % gamma, "%s m=%d: %.3f s (%.2f ms/iteration), %d kernel launches, %d Ritz values with residual < 1e-10*|.|, max rel resid %.2e" %
          (name, m, dt, dt / m * 1e3, _lib.lib.nepb_launch_count() - l0, len(lam), max(res / np.abs(lam)) if len(lam) else 0), flush=True)
if cpu_m:
    from oracle import nep as o, solvers as osol
    onep = o.nep_gallery("nlevp_native_gun")
    t0 = time.perf_counter()
    lo, Qo, Vo = osol.iar(onep, maxit=cpu_m, check_error_every=cpu_m, **kw)
    dt = time.perf_counter() - t0
    print("oracle iar (CPU, NumPy/SuperLU) m=%d: %.2f s (%.1f ms/iteration), %d Ritz values" % (cpu_m, dt, dt / cpu_m * 1e3, len(lo)))
