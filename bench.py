#!/usr/bin/env python
"""bench.py -- the measurement contract of the nepb200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload auto|spmm|contour]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how every field is produced.
Workloads (BASELINE.json):
  contour : config C3 -- gun SPMF, contour_beyn moment integration, N=128 quadrature points, k=20 probe columns,
            points sharded over the ranks, one NCCL reduce of the moment block (strong scaling).
  spmm    : config C4 -- synthetic degree-3 PEP, n=10^6, 21-point stencil, fused multi-term SpMM M(lam)V for
            k in {1,8,20}; this is where `roofline` (HBM) comes from.
A step is one pass of the hot path over one batch: one full 128-point contour integration (contour) or one
fused SpMM launch per k (spmm).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# distributed plumbing (torch.distributed only for rendezvous / barrier / max-over-ranks)
# ------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.td = None
        if self.world > 1:
            import torch
            import torch.distributed as td
            torch.cuda.set_device(self.local_rank)
            td.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
            self.td, self.torch = td, torch

    def barrier(self):
        if self.td:
            self.td.barrier()

    def max(self, x: float) -> float:
        if not self.td:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if not self.td:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.td:
            self.td.destroy_process_group()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# workload: fused SPMF SpMM on config C4
# ------------------------------------------------------------------------------------------------
def build_c4(grid):
    import nepb200
    from nepb200 import synthetic, Monomial
    t0 = time.time()
    mats, st = synthetic.stencil_pep(grid)
    csc = [m.tocsc() for m in mats]
    for m in csc:
        m.sort_indices()
    dnep = nepb200.B200SPMF(csc, [Monomial(i) for i in range(4)])
    log("[bench] C4 operator: n=%d nnz_u=%d built in %.1fs" % (dnep.n, dnep.nnz_union, time.time() - t0))
    return dnep, mats, st


def bench_spmm(args, dist, ks=(1, 8, 20)):
    import nepb200
    from nepb200 import synthetic, Block, _lib
    lib = _lib.lib
    import ctypes as C
    grid = args.grid
    dnep, mats, st = build_c4(grid)
    n = dnep.n
    lam = 0.3 + 0.2j
    coef = dnep.coefficients(lam)
    peak, peak_src = load_peaks()
    out = {}
    for k in ks:
        V = synthetic.stencil_block(st, n, k)
        Vb, Zb = Block.from_host(V), Block(n, k)
        for _ in range(max(args.warmup, 3)):
            dnep.apply_block(_lib.COEF_SCALAR, Vb, coef, Zb)
        lib.nepb_synchronize()
        dist.barrier()
        l0 = lib.nepb_launch_count()
        ms = C.c_float()
        lib.nepb_timer_start()
        for _ in range(args.steps):
            dnep.apply_block(_lib.COEF_SCALAR, Vb, coef, Zb)
        lib.nepb_timer_stop(C.byref(ms))
        dist.barrier()
        launches = lib.nepb_launch_count() - l0
        t = dist.max(ms.value / args.steps)  # ms per launch, max over ranks
        nbytes = dnep.apply_bytes(_lib.COEF_SCALAR, k, k)
        # e2e: host buffers through nepb_spmf_apply (H2D of V, D2H of Z inside the timed region)
        Z = dnep.apply(_lib.COEF_SCALAR, V, coef, k)
        t0 = time.perf_counter()
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            Z = dnep.apply(_lib.COEF_SCALAR, V, coef, k)
        te = dist.max((time.perf_counter() - t0) / reps * 1e3)
        out[k] = {"k": k, "ms": t, "gbs": nbytes / t / 1e6, "bytes": int(nbytes), "frac": nbytes / t / 1e6 / peak,
                  "launches": int(launches), "e2e_ms": te, "e2e_gbs": nbytes / te / 1e6,
                  "h2d": int(n * k * 16), "d2h": int(n * k * 16), "checksum": float(np.abs(Z).sum())}
        log("[bench] spmm k=%d: %.1f us/launch, %.0f GB/s (%.1f%% of %s); e2e %.2f ms" %
            (k, t * 1e3, out[k]["gbs"], 100 * out[k]["frac"], peak_src, te))
        Vb.close()
        Zb.close()
    return dnep, mats, out, peak, peak_src


def cpu_spmm_baseline(mats, k, budget_s=10.0):
    """CPU port: sum_i c_i A_i V with SciPy CSR (one core), the reference's algorithm for compute_MM with S = lam*I
    (NEPTypes.jl:299-311: p separate SpMMs).  Bounded sample: as many full passes as fit in the budget (>= 1)."""
    lam = 0.3 + 0.2j
    n = mats[0].shape[0]
    rng = np.random.default_rng(0)
    V = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
    t0 = time.perf_counter()
    reps = 0
    while True:
        Z = np.zeros((n, k), dtype=complex)
        for i, A in enumerate(mats):
            Z += A @ (V * lam ** i)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 20:
            break
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "spmm", "contour"])
    ap.add_argument("--grid", type=int, default=1000, help="C4 grid side (n = grid^2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    dist = Dist()
    import nepb200
    from nepb200 import _lib
    if nepb200.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the nepb200 hot path has no CPU fallback")
    _lib.check(_lib.lib.nepb_set_device(dist.local_rank))

    sampler = ClockSampler(dist.local_rank)
    if dist.rank == 0:
        sampler.start()
    dnep, mats, spmm, peak, peak_src = bench_spmm(args, dist)
    clocks = sampler.stop() if dist.rank == 0 else None

    head = spmm[1]
    line = {
        "metric": "SPMF fused SpMM GB/s (algorithmic bytes / device time), config C4",
        "value": head["gbs"] * dist.world, "unit": "GB/s", "n_gpus": dist.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (complex128 V/Z, real f64 A_i)", "data": "synthetic",
        "config": {"workload": "C4 synthetic PEP deg 3, n=%d, nnz_u=%d, p=4, fused SpMM M(lam)V k=1" % (dnep.n, dnep.nnz_union),
                   "l2": "inputs (%.0f MB) larger than the 126 MB L2; no flush needed" % (head["bytes"] / 1e6),
                   "parallelism": "replicas only" if dist.world > 1 else "single GPU"},
        "roofline": {"bound": "hbm", "achieved": head["gbs"], "peak": peak, "unit": "GB/s", "frac": head["frac"],
                     "traffic": None, "peak_source": peak_src, "kernel": "spmm_fused_kernel<4,real,SCALAR> k=1",
                     "algorithmic_bytes_per_launch": head["bytes"]},
        "spmm": {str(k): {kk: v[kk] for kk in ("ms", "gbs", "frac", "bytes", "e2e_ms")} for k, v in spmm.items()},
        "e2e": {"value": head["e2e_gbs"] * dist.world, "unit": "GB/s", "h2d_bytes_per_step": head["h2d"],
                "d2h_bytes_per_step": head["d2h"]},
        "gpu_launches": int(sum(v["launches"] for v in spmm.values())),
        "clocks": clocks,
    }
    if dist.rank == 0 and not args.no_cpu_baseline:
        t = cpu_spmm_baseline(mats, 1)
        line["cpu_baseline"] = {"value": head["bytes"] / t / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                "sample": "full C4 SpMM passes (k=1) with SciPy CSR, p separate products, %.2f s each" % t}
    if dist.rank == 0:
        print(json.dumps(line), flush=True)
    dist.close()


if __name__ == "__main__":
    main()
