"""Config C5 (BASELINE.json): the dense blocks of tiar at the size of a large waveguide instance (WEP nx = nz + 4,
nz = 3*5*7*9 = 945: n = nx*nz + 2*nz = 898 695), m = 200, with a synthetic basis Z (MSWS stream) -- SURVEY.md 8(d).
Times, with CUDA events on the library stream:
  * Z[:, :k] * C (k x k)   -- the tall-skinny ZGEMMs of src/method_tiar.jl:119,187-189 (nepb_block_gemm, FP64 tensor cores)
  * DGKS of one vector against Z[:, :k]  -- src/method_tiar.jl:128 (nepb_orth_dgks, HBM-bound)
and, as the FP64 ceiling measured in the same run, cuBLAS ZGEMM through torch.matmul on the same shapes and on a square
4096^3 product.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib, Block, block_gemm, dgks

lib = _lib.lib
nz = int(sys.argv[1]) if len(sys.argv) > 1 else 945
n = (nz + 4) * nz + 2 * nz
m = 200
peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
st = _lib.msws_state(5)
Zb, Yb = Block(n, m + 1), Block(n, m + 1)
for c0 in range(0, m + 1, 8):
    kc = min(8, m + 1 - c0)
    blk = (1 - 2 * _lib.msws_fill(st, n * kc)).reshape(n, kc, order="F") + 1j * (1 - 2 * _lib.msws_fill(st, n * kc)).reshape(n, kc, order="F")
    Zb.upload(blk / np.sqrt(n), c0)


def timed(fn, reps):
    fn()
    lib.nepb_synchronize()
    ms = C.c_float()
    lib.nepb_timer_start()
    for _ in range(reps):
        fn()
    lib.nepb_timer_stop(C.byref(ms))
    return ms.value / reps


out = {"workload": "C5 tiar dense blocks, n=%d (WEP nz=%d), m=%d" % (n, nz, m), "hbm_peak_gbs": peak, "gemm": {}, "dgks": {}}
rng = np.random.default_rng(0)
for k in (25, 50, 100, 200):
    Cm = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
    t = timed(lambda: block_gemm(Zb, 0, k, Cm, Yb, 0), 5)
    flops = 8.0 * n * k * k
    nbytes = 2.0 * 16 * n * k
    out["gemm"][k] = {"ms": t, "tflops": flops / t / 1e9, "gbs": nbytes / t / 1e6, "hbm_frac": nbytes / t / 1e6 / peak}
for k in (50, 100, 200):
    sweeps = []

    def one():
        h, nrm, sw = dgks(Zb, k, Zb, k)
        sweeps.append(sw)
    t = timed(one, 5)
    nbytes = 2.0 * 16 * n * k * np.mean(sweeps)  # each sweep reads the basis twice (h = V'w, w -= V h)
    out["dgks"][k] = {"ms": t, "sweeps": float(np.mean(sweeps)), "gbs": nbytes / t / 1e6, "hbm_frac": nbytes / t / 1e6 / peak}
try:
    import torch
    dev = torch.device("cuda", 0)
    ceil = {}
    for k in (50, 100, 200):
        A = torch.randn(n, k, dtype=torch.complex128, device=dev)
        B = torch.randn(k, k, dtype=torch.complex128, device=dev)
        for _ in range(2):
            torch.matmul(A, B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            torch.matmul(A, B)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5
        ceil[k] = {"ms": t, "tflops": 8.0 * n * k * k / t / 1e9}
        del A, B
    A = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
    for _ in range(2):
        torch.matmul(A, A)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        torch.matmul(A, A)
    e1.record()
    torch.cuda.synchronize()
    ceil["square_4096"] = {"ms": e0.elapsed_time(e1) / 3, "tflops": 8.0 * 4096 ** 3 / (e0.elapsed_time(e1) / 3) / 1e9}
    out["cublas_zgemm"] = ceil
except Exception as ex:  # torch is only the yardstick here
    out["cublas_zgemm"] = {"error": str(ex)}
print(json.dumps(out))
