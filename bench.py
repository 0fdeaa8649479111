#!/usr/bin/env python
"""bench.py -- the measurement contract of the nepb200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One JSON line on stdout (rank 0).  DESIGN.md "Measurement" explains every field.

Headline (BASELINE.json config C3): contour_beyn on the gun SPMF, N=128 quadrature nodes, k=20 probe columns.
A step = one full moment integration S_j = sum_i w_ij M(lambda_i)^-1 Vh (128 batched factorisations + 128 x 20 solves
+ accumulate), nodes sharded round-robin over the ranks, one NCCL all-reduce of the moment block.  `value` is
quadrature-point solves per second of the whole job (strong scaling: the 128 nodes are split over the ranks).

Roofline (BASELINE.json config C4): the fused multi-term SPMF SpMM M(lam)V on the synthetic degree-3 PEP, n=10^6,
nnz_u=20 956 020, measured live in the same run (k=1 headline, k=8 and k=20 beside it) against the measured HBM peak.

--impl reference times the CPU restatement of the reference's own path (oracle/, SciPy SuperLU standing in for UMFPACK;
Julia is not installed in this image) on the host cores, one process per core over the quadrature nodes -- the
reference's documented `@distributed (+)` scheme (docs/src/tutorial_contour.md:205-231).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NumPy's BLAS worker threads keep spinning for tens of milliseconds after a call (the parity norms below) and then compete with
# the library's host-copy threads for the cores: the host-buffer timings of the next call went from 1.2 ms to 24 ms
# (gpurun_out/r2_hostcopy3.log).  The harness itself needs no threaded BLAS.
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import numpy as np  # noqa: E402

GUN_SIGMA = 150.0 ** 2
GUN_RADIUS = 500.0
GUN_N = 128
GUN_K = 20


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
            self.lines.append((time.time(), ln.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
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
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
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
    def __init__(self, use_cuda=True):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.td = None
        if self.world > 1 and use_cuda:
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

    def bcast_bytes(self, b: bytes, n: int) -> bytes:
        if not self.td:
            return b
        t = self.torch.zeros(n, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            t.copy_(self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8))
        self.td.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

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
# problem set-up
# ------------------------------------------------------------------------------------------------
def load_gun_csc():
    import scipy.sparse as sp
    z = np.load(os.path.join(ROOT, "tests", "golden", "gun.npz"))
    n = int(z["n"])
    return [sp.csc_matrix((z[k + "_data"], z[k + "_indices"], z[k + "_indptr"]), shape=(n, n)) for k in ("K", "M", "W1", "W2")]


def gun_operator():
    import nepb200
    from nepb200 import ONE, IDENTITY, PowShift
    K, M, W1, W2 = load_gun_csc()
    return nepb200.B200SPMF([K, -M, W1, W2], [ONE, IDENTITY, PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])


def gun_probe(n, k):
    """MSWS 1-2u probe matrix, column-major draw (reproducible in Julia, Python and C)."""
    from nepb200 import _lib
    st = _lib.msws_state(0)
    return np.asfortranarray((1 - 2 * _lib.msws_fill(st, n * k)).reshape(n, k, order="F")).astype(np.complex128)


def beyn_nodes(N, sigma, radius):
    h = 2 * np.pi / N
    t = h * np.arange(N)
    g = radius * (np.cos(t) + 1j * np.sin(t))
    gp = radius * (-np.sin(t) + 1j * np.cos(t))
    W = np.stack([gp * h / (2j * np.pi), gp * g * h / (2j * np.pi)], axis=1)
    return g + sigma, W


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


# ------------------------------------------------------------------------------------------------
# workload: fused SPMF SpMM on config C4 (roofline)
# ------------------------------------------------------------------------------------------------
def bench_spmm(args, ks=(1, 8, 20), with_cpu=True):
    """Config C4.  SCALAR mode (M(lambda) V, the north-star formula) for k = 1, 8, 20; GENERAL mode as the solver loops call it
    (compute_Mlincomb with k = 20 / 100 basis columns, NEPTypes.jl:972-1011, and the nleigs stacked product with N = 7 blocks,
    method_nleigs.jl:456-472).  Every timed product is compared with a SciPy CSR product on the same V (parity at the stated
    size: relative error <= 1e-12 or the run fails)."""
    from nepb200 import synthetic, Block, _lib
    lib = _lib.lib
    dnep, mats, st = build_c4(args.grid)
    n = dnep.n
    lam = 0.3 + 0.2j
    coef = dnep.coefficients(lam)
    peak, peak_src = load_peaks()
    csr = [m.tocsr() for m in mats]
    Mo = sum(m * lam ** i for i, m in enumerate(csr)).tocsr()
    out = {}

    def time_block(fn, reps):
        for _ in range(max(args.warmup, 3)):
            fn()
        lib.nepb_synchronize()
        l0 = lib.nepb_launch_count()
        ms = C.c_float()
        lib.nepb_timer_start()
        for _ in range(reps):
            fn()
        lib.nepb_timer_stop(C.byref(ms))
        return ms.value / reps, (lib.nepb_launch_count() - l0) // reps

    for k in ks:
        V = synthetic.stencil_block(st, n, k)
        Vb, Zb = Block.from_host(V), Block(n, k)
        t, launches = time_block(lambda: dnep.apply_block(_lib.COEF_SCALAR, Vb, coef, Zb), max(args.steps, 20))
        nbytes = dnep.apply_bytes(_lib.COEF_SCALAR, k, k)
        Zref = Mo @ V
        err = float(np.linalg.norm(Zb.download() - Zref) / np.linalg.norm(Zref))
        Zh = np.empty((n, k), dtype=np.complex128, order="F")  # the caller's result array, reused (no fresh pages per call)
        dnep.apply(_lib.COEF_SCALAR, V, coef, k, out=Zh)
        tes = []
        for _ in range(5):
            t0 = time.perf_counter()
            dnep.apply(_lib.COEF_SCALAR, V, coef, k, out=Zh)
            tes.append((time.perf_counter() - t0) * 1e3)
        te = float(np.median(tes))
        log("[bench]   host-buffer calls k=%d: %s ms" % (k, " ".join("%.2f" % x for x in tes)))
        err = max(err, float(np.linalg.norm(Zh - Zref) / np.linalg.norm(Zref)))
        out[k] = {"k": k, "ms": t, "gbs": nbytes / t / 1e6, "bytes": int(nbytes), "frac": nbytes / t / 1e6 / peak,
                  "launches": int(launches), "e2e_ms": te, "e2e_gbs": nbytes / te / 1e6, "parity_relerr": err}
        log("[bench] spmm k=%d: %.1f us/launch, %.0f GB/s (%.1f%% of %s); host-buffer call %.2f ms; parity vs SciPy %.1e" %
            (k, t * 1e3, out[k]["gbs"], 100 * out[k]["frac"], peak_src, te, err))
        if not err < 1e-12:
            raise SystemExit("bench: fused SpMM differs from the SciPy product at k=%d: %g" % (k, err))
        Vb.close()
        Zb.close()
    # GENERAL mode: panel product X = V [C_1..C_p] + stacked gather (two launches)
    general = {}
    rng = np.random.default_rng(0)
    for name, k, q in (("mlincomb_k20", 20, 1), ("mlincomb_k100", 100, 1), ("nleigs_stacked_N7", 7, 1)):
        V = synthetic.stencil_block(st, n, k)
        Vb, Zb = Block.from_host(V), Block(n, q)
        Cs = [rng.standard_normal((k, q)) + 1j * rng.standard_normal((k, q)) for _ in range(dnep.p)]
        Cblk = np.ascontiguousarray(np.stack([np.asfortranarray(c).T.copy() for c in Cs]))
        t, launches = time_block(lambda: dnep.apply_block(_lib.COEF_GENERAL, Vb, Cblk, Zb), max(args.steps, 20))
        nbytes = dnep.apply_bytes(_lib.COEF_GENERAL, k, q)
        Zref = sum(m @ (V @ c) for m, c in zip(csr, Cs))
        err = float(np.linalg.norm(Zb.download() - Zref) / np.linalg.norm(Zref))
        general[name] = {"k": k, "q": q, "ms": t, "gbs": nbytes / t / 1e6, "bytes": int(nbytes), "frac": nbytes / t / 1e6 / peak,
                         "launches": int(launches), "parity_relerr": err}
        log("[bench] general %s (k=%d q=%d): %.1f us, %.0f GB/s (%.1f%%), %d launches; parity %.1e" %
            (name, k, q, t * 1e3, general[name]["gbs"], 100 * general[name]["frac"], launches, err))
        if not err < 1e-12:
            raise SystemExit("bench: GENERAL-mode product %s differs from NumPy: %g" % (name, err))
        Vb.close()
        Zb.close()
    cpu = None
    if with_cpu and not args.no_cpu_baseline:
        # CPU side: the C restatement of the reference's compute_MM loop (oracle/csrc/spmf_mm.c) -- once in the reference's own
        # serial form (CSC, one thread) and once row-parallel over all host cores
        from oracle.cspmf import CSpmf
        cs = CSpmf(mats)
        cores = max(1, min(host_cores(), 64))
        f = [lam ** i for i in range(dnep.p)]
        v1 = synthetic.stencil_block(st, n, 1)
        res = {}
        for label, fn, th in (("serial_csc", cs.mm_csc, 1), ("allcore_csr", cs.mm_csr, cores)):
            fn(f, v1, th)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 2.0 or reps < 2:
                Zc = fn(f, v1, th)
                reps += 1
            res[label] = (time.perf_counter() - t0) / reps
        err = float(np.linalg.norm(Zc - Mo @ v1) / np.linalg.norm(Mo @ v1))
        cpu = {"value": out[1]["bytes"] / res["allcore_csr"] / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
               "serial_reference_loop_gbs": out[1]["bytes"] / res["serial_csc"] / 1e9,
               "sample": "full C4 SpMM passes (k=1) for 2 s each: C restatement of compute_MM (NEPTypes.jl:296-316), p separate "
                         "products; serial CSC loop %.3f s/pass, row-parallel CSR on %d threads %.4f s/pass; parity vs SciPy %.1e"
                         % (res["serial_csc"], cores, res["allcore_csr"], err)}
    dnep.close()
    return out, peak, peak_src, cpu, general


def bench_wep(args, peak, nz=945):
    """Config C5 (WEP, JARLEBRING, nz = 3*5*7*9 = 945, nx = nz + 4, n = 898 695): the native compute_Mlincomb
    (Waveguide.jl:324-379) on the device -- the Sylvester-form interior as one stencil pass plus the boundary transforms -- for
    1, 3 and 20 columns, against the HBM roofline (algorithmic bytes: nepb_wep_mlincomb_bytes), each result compared with a
    NumPy / SciPy evaluation of the reference's formula on the same operands (sparse Dzz / Dz / Dxx products and numpy.fft)."""
    import nepb200
    from nepb200 import Block, _lib
    from nepb200 import wep as pw
    lib = _lib.lib
    nep = nepb200.nep_gallery_WEP(nx=nz + 4, nz=nz, benchmark_problem="JARLEBRING", neptype="WEP")
    n, nx = nep.n, nep.nx
    lam = -2.7 - 3.1j
    rng = np.random.default_rng(5)
    out = {"n": int(n), "nx": int(nx), "nz": int(nz)}

    def reference(V, a):
        m = nx * nz
        X = [V[:m, j].reshape(nz, nx, order="F") for j in range(min(V.shape[1], 3))]
        y1 = (nep.A(lam) @ X[0] + (nep.Dxx.T @ X[0].T).T + nep.K * X[0]) * a[0]
        for d in range(1, len(X)):
            y1 = y1 + (nep.A(lam, d) @ X[d]) * a[d]
        y1 = y1.reshape(-1, order="F") + (nep.C1 @ V[m:, 0]) * a[0]
        D = 1j * pw.sqrt_derivative(1.0, np.concatenate([nep.b, nep.b]), np.concatenate([nep.cM, nep.cP]), V.shape[1] - 1, lam)
        D[:, 0] += nep.d0
        Rinv = lambda x: np.fft.ifft(nep.bbinv * x[::-1])  # noqa: E731
        R = lambda x: (nep.bb * np.fft.fft(x))[::-1]  # noqa: E731
        t = sum(D[:, j] * np.concatenate([Rinv(V[m:m + nz, j]), Rinv(V[m + nz:, j])]) * a[j] for j in range(V.shape[1]))
        return np.concatenate([y1, np.concatenate([R(t[:nz]), R(t[nz:])]) + (nep.C2T @ V[:m, 0]) * a[0]])

    for na in (1, 3, 20):
        V = rng.standard_normal((n, na)) + 1j * rng.standard_normal((n, na))
        a = (rng.standard_normal(na) + 1j * rng.standard_normal(na)) / np.array([math.factorial(min(j, 10)) for j in range(na)])
        Vb, Zb = Block.from_host(V), Block(n, 1)
        reps = max(args.steps, 20)
        for _ in range(max(args.warmup, 3)):
            nep.mlincomb_block(lam, Vb, 0, na, a, Zb, 0)
        lib.nepb_synchronize()
        l0 = lib.nepb_launch_count()
        ms = C.c_float()
        lib.nepb_timer_start()
        for _ in range(reps):
            nep.mlincomb_block(lam, Vb, 0, na, a, Zb, 0)
        lib.nepb_timer_stop(C.byref(ms))
        t = ms.value / reps
        launches = (lib.nepb_launch_count() - l0) // reps
        nbytes = int(lib.nepb_wep_mlincomb_bytes(nep._h, na))
        zref = reference(V, a)
        err = float(np.linalg.norm(Zb.download()[:, 0] - zref) / np.linalg.norm(zref))
        out[str(na)] = {"columns": na, "ms": t, "bytes": nbytes, "gbs": nbytes / t / 1e6, "frac": nbytes / t / 1e6 / peak, "launches": int(launches),
                        "parity_relerr": err}
        log("[bench] WEP compute_Mlincomb n=%d, %d column(s): %.1f us, %.0f GB/s (%.1f%% of the HBM peak), %d launches; parity %.1e" %
            (n, na, t * 1e3, out[str(na)]["gbs"], 100 * out[str(na)]["frac"], launches, err))
        if not err < 1e-12:
            raise SystemExit("bench: WEP compute_Mlincomb differs from the NumPy evaluation of the reference formula: %g" % err)
        Vb.close()
        Zb.close()
    # C5 end to end: tiar with the basis in HBM (src/method_tiar.jl) on this problem, Schur-complement solver (Waveguide.jl:476-567)
    # on the device multifrontal LU.  Reported, never fatal for the headline line.
    try:
        sigma = -3 - 3.5j
        t0 = time.perf_counter()
        solver = nepb200.WEPLinSolverCreator(solver_type="factorized").create_linsolver(nep, sigma)
        lib.nepb_synchronize()
        t_fact = time.perf_counter() - t0

        class _Creator:
            def create_linsolver(self, nep_, lam_):
                return solver
        l0 = lib.nepb_launch_count()
        t0 = time.perf_counter()
        lams, Q, _, _ = nepb200.tiar_device(nep, sigma=sigma, neigs=3, maxit=60, v=np.ones(n) / np.sqrt(n), tol=1e-8, linsolvercreator=_Creator())
        lib.nepb_synchronize()
        t_tiar = time.perf_counter() - t0
        res = [float(np.linalg.norm(nep.compute_Mlincomb(lams[i], Q[:, i])) / np.linalg.norm(Q[:, i])) for i in range(len(lams))]
        out["tiar_c5"] = {"schur_unknowns": int(nx * nz), "schur_lu": nepb200.symbolic_info(solver.schur), "assembly_analysis_factorisation_s": t_fact,
                          "tiar_s": t_tiar, "gpu_launches": int(lib.nepb_launch_count() - l0), "eigenvalues": [[float(x.real), float(x.imag)] for x in lams],
                          "residuals": res}
        log("[bench] C5 tiar on WEP n=%d: Schur-complement LU %.2f s, tiar %.2f s, %d eigenvalues, max residual %.1e" %
            (n, t_fact, t_tiar, len(lams), max(res) if res else float("nan")))
        solver.fact.lu.close()
        solver.schur.close()
    except Exception as exc:  # noqa: BLE001
        out["tiar_c5"] = {"error": repr(exc)}
        log("[bench] C5 tiar on WEP failed: %r" % (exc,))
    nep.close()
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle contour loop, one process per core over the nodes
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init():
    from oracle import nep as o
    _W["nep"] = o.nep_gallery("nlevp_native_gun")
    n = _W["nep"].n
    from oracle import gallery as g
    rng = g.MSWS_RNG(0)
    u = np.array([g.gen_rng_float(rng) for _ in range(n * GUN_K)])
    _W["Vh"] = np.asfortranarray((1 - 2 * u).reshape(n, GUN_K, order="F")).astype(np.complex128)


def _cpu_worker_nodes(arg):
    """One worker's share of the trapezoid sum: factor + 20-column solve + accumulate per node (oracle/solvers.py)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    from oracle import nep as o
    lams, W = arg
    nep, Vh = _W["nep"], _W["Vh"]
    S = np.zeros(Vh.shape + (2,), dtype=np.complex128)
    for lam, w in zip(lams, W):
        M = sp.csc_matrix(o.compute_Mder(nep, lam), dtype=np.complex128)
        X = sla.splu(M, permc_spec="MMD_AT_PLUS_A").solve(Vh)
        S[:, :, 0] += X * w[0]
        S[:, :, 1] += X * w[1]
    return S


def cpu_worker_main():
    """Child process of CpuContour: `bench.py --cpu-worker`.  Protocol on stdin/stdout: prints "ready", then for every
    line "<first> <count> <stride>" computes its nodes and answers "done <seconds> <checksum>"."""
    _cpu_worker_init()
    lams, W = beyn_nodes(GUN_N, GUN_SIGMA, GUN_RADIUS)
    print("ready", flush=True)
    for ln in sys.stdin:
        f = ln.split()
        if not f or f[0] == "quit":
            break
        first, count, stride = int(f[0]), int(f[1]), int(f[2])
        idx = (first + stride * np.arange(count)) % GUN_N
        t0 = time.perf_counter()
        S = _cpu_worker_nodes((lams[idx], W[idx]))
        dt = time.perf_counter() - t0
        if len(f) > 3:  # optional: where to leave this worker's partial moments (parity check of the GPU arm)
            np.save(f[3], S)
        print("done %.6f %.17g" % (dt, float(np.abs(S).sum())), flush=True)


class CpuContour:
    """One single-threaded worker process per host core over the quadrature nodes -- the reference's `julia -p <cores>`
    + `@distributed (+)` scheme (docs/src/tutorial_contour.md:205-231).  Start-up (imports, matrix load) is not timed."""

    def __init__(self, cores):
        self.cores = cores
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        self.procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker"], stdin=subprocess.PIPE,
                                       stdout=subprocess.PIPE, text=True, env=env, cwd=ROOT) for _ in range(cores)]
        for p in self.procs:
            ln = p.stdout.readline()
            if ln.strip() != "ready":
                raise RuntimeError("CPU worker failed to start: %r" % ln)

    def run(self, nodes_per_core, save_prefix=None):
        t0 = time.perf_counter()
        for c, p in enumerate(self.procs):
            p.stdin.write("%d %d %d%s\n" % (c, nodes_per_core, self.cores, (" %s.%d.npy" % (save_prefix, c)) if save_prefix else ""))
            p.stdin.flush()
        for p in self.procs:
            ln = p.stdout.readline().split()
            if not ln or ln[0] != "done":
                raise RuntimeError("CPU worker died")
        return self.cores * nodes_per_core, time.perf_counter() - t0

    def close(self):
        for p in self.procs:
            try:
                p.stdin.write("quit\n")
                p.stdin.flush()
                p.stdin.close()
            except Exception:
                pass
        for p in self.procs:
            try:
                p.wait(timeout=10)
            except Exception:
                p.kill()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


WORKLOAD = "C3 gun SPMF n=9956 p=4 nnz_u=148318, contour_beyn sigma=150^2 radius=500 N=128 k=20"


def run_reference(args):
    """The reference's CPU path on the host cores: one worker process per core over the quadrature nodes.  --steps / --warmup are
    honoured as given; one step is a bounded sample of the 128-node contour -- one node per worker process -- so that the run ends
    within minutes (a node takes ~2 s per core); the value is a rate (nodes per second), as on the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, min(host_cores(), 64))
    cpu = CpuContour(cores)
    for _ in range(args.warmup):
        cpu.run(1)
    tot_n, tot_t = 0, 0.0
    steps = max(1, args.steps)
    for _ in range(steps):
        nn, t = cpu.run(1)
        tot_n += nn
        tot_t += t
    cpu.close()
    v = tot_n / tot_t
    sample = ("each step = %d quadrature nodes (one per worker process) of the N=128 gun contour, k=20: factor + 20-column solve + "
              "accumulate per node; SciPy SuperLU (MMD_AT_PLUS_A) standing in for UMFPACK" % cores)
    line = {"impl": "reference", "metric": "contour_beyn quadrature-point solves/sec (gun, N=128, k=20)", "value": v,
            "unit": "solves/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": tot_t / steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128 factors / solves, real f64 A_i)",
            "data": "gun matrices (reference fixture) + synthetic MSWS probe",
            "config": {"workload": WORKLOAD, "parallelism": "%d host processes over nodes" % cores},
            "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of NEP-PACK's contour loop (Julia / UMFPACK are not installable here: kind = port; "
                    "bench/ref_cpu.jl times the real reference where Julia exists)"}
    print(json.dumps(line), flush=True)


def fp64_zgemm_peak():
    """Measured FP64 ceiling for the factorisation roofline: cuBLAS ZGEMM 4096^3 through torch (measurement plumbing only)."""
    try:
        import torch
        a = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
        b = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 8 * 4096.0 ** 3 / (best * 1e-3) / 1e12, "cuBLAS ZGEMM 4096^3 through torch.matmul, best of 3, same run"
    except Exception as e:  # noqa: BLE001
        return 36.8, "round-1 measurement (cuBLAS ZGEMM 4096^3); in-run measurement failed: %s" % e


# ------------------------------------------------------------------------------------------------
# main arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=1000, help="C4 grid side (n = grid^2)")
    ap.add_argument("--batch", type=int, default=128, help="quadrature nodes factorised concurrently per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spmm", action="store_true")
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:
        cpu_worker_main()
        return
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    dist = Dist()
    import nepb200
    from nepb200 import _lib
    lib = _lib.lib
    if nepb200.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the nepb200 hot path has no CPU fallback")
    _lib.check(lib.nepb_set_device(dist.local_rank))
    if dist.world > 1:
        idb = (C.c_char * 128)()
        if dist.rank == 0:
            _lib.check(lib.nepb_comm_unique_id(idb))
        raw = dist.bcast_bytes(bytes(idb.raw), 128)
        _lib.check(lib.nepb_comm_init(dist.world, dist.rank, raw))

    # ---- contour workload ---------------------------------------------------------------------------------
    dnep = gun_operator()
    n = dnep.n
    Vh = gun_probe(n, GUN_K)
    lams, W = beyn_nodes(GUN_N, GUN_SIGMA, GUN_RADIUS)
    mine = np.arange(dist.rank, GUN_N, dist.world)
    batch = min(args.batch, len(mine))
    integ = nepb200.ContourIntegrator(dnep, GUN_K, 2, batch)
    coef = np.ascontiguousarray(np.stack([dnep.coefficients(l) for l in lams[mine]]))
    Wm = np.ascontiguousarray(W[mine])
    Vf = _lib.as_c128_f(Vh)
    S = np.empty((n, GUN_K, 2), dtype=np.complex128, order="F")
    reduce = 1 if dist.world > 1 else 0
    reduce_e2e = 2 if dist.world > 1 else 0  # only rank 0 extracts (method_beyncontour.jl:114-184): the other ranks skip the download
    _lib.check(lib.nepb_contour_set_probe(integ._h, _lib.ptr(Vf), n))
    # the probe and the moment array are passed again every step: page-lock them once (nepb_host_register, the documented option
    # for buffers a caller re-uses) so that their transfers are direct DMA from / to pinned host memory
    pinned = []
    for arr in (Vf, S):
        if lib.nepb_host_register(_lib.ptr(arr), arr.nbytes) == 0:
            pinned.append(arr)

    def step_dev():
        _lib.check(lib.nepb_contour_integrate_dev(integ._h, len(mine), _lib.ptr(coef), _lib.ptr(Wm), reduce))

    def step_e2e():
        _lib.check(lib.nepb_contour_integrate(integ._h, len(mine), _lib.ptr(coef), _lib.ptr(Wm), _lib.ptr(Vf), n, reduce_e2e, _lib.ptr(S), None))

    sampler = ClockSampler(dist.local_rank)
    if dist.rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_dev()
    lib.nepb_synchronize()
    dist.barrier()
    l0 = lib.nepb_launch_count()
    ms = C.c_float()
    tmark0 = sampler.mark()
    lib.nepb_timer_start()
    for _ in range(args.steps):
        step_dev()
    lib.nepb_timer_stop(C.byref(ms))
    lib.nepb_synchronize()
    tmark1 = sampler.mark()
    dist.barrier()
    launches = lib.nepb_launch_count() - l0
    t_step = dist.max(ms.value / args.steps)  # ms per step, max over ranks
    launches_total = int(dist.sum(float(launches)))
    value = GUN_N / (t_step * 1e-3)
    # end to end through the host-buffer C-ABI call (H2D of the probe, D2H of the moments inside the timed region)
    step_e2e()
    dist.barrier()
    t0 = time.perf_counter()
    e2e_reps = max(3, min(args.steps, 10))
    for _ in range(e2e_reps):
        step_e2e()
    te = dist.max((time.perf_counter() - t0) / e2e_reps * 1e3)
    dist.barrier()
    clocks = sampler.stop(tmark0, tmark1) if dist.rank == 0 else None
    # sanity of the timed result: the moments must give the gun reference eigenvalue (test/gun_native.jl:9)
    lam_found = None
    if dist.rank == 0:
        lam, V, info = nepb200.beyn_extract(S[:, :, 0], S[:, :, 1], GUN_SIGMA, (GUN_RADIUS, GUN_RADIUS), GUN_K, 5, 1e-6, np.sqrt(np.finfo(float).eps),
                                            nepb200.DefaultErrmeasure(dnep), True)
        lam_found = [complex(x) for x in lam]
        log("[bench] contour: %.2f ms/step -> %.1f solves/s on %d GPU(s); e2e %.2f ms; eigenvalues inside: %s (p=%d)" %
            (t_step, value, dist.world, te, lam_found, info["p"]))
    sym = nepb200.symbolic_info(dnep)
    for arr in pinned:
        lib.nepb_host_unregister(_lib.ptr(arr))
    integ.close()
    dnep.close()

    # ---- SpMM roofline (every rank runs its own replica; rank 0 reports) -------------------------------------
    spmm, peak, peak_src, cpu_spmm, general = (None, None, None, None, None)
    if not args.no_spmm:
        spmm, peak, peak_src, cpu_spmm, general = bench_spmm(args, with_cpu=(dist.rank == 0 and dist.world == 1))

    line = {
        "metric": "contour_beyn quadrature-point solves/sec (gun, N=128, k=20)",
        "value": value, "unit": "solves/s", "n_gpus": dist.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64 (complex128 factors / solves, real f64 A_i)", "data": "gun matrices (reference fixture) + synthetic MSWS probe",
        "config": {"workload": WORKLOAD,
                   "parallelism": "quadrature nodes round-robin over %d rank(s), batch %d per GPU in 8 node groups (streams), forward solve pipelined beside the factorisation, one ncclAllReduce of 6.4 MB" % (dist.world, batch),
                   "l2": "factor storage per batch %.0f MB > 126 MB L2; SpMM roofline inputs 790 MB > L2; no flush needed" % (batch * sym["front_entries"] * 16e-6),
                   "lu": sym},
        "e2e": {"value": GUN_N / (te * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": int(n * GUN_K * 16 + coef.nbytes + Wm.nbytes),
                "d2h_bytes_per_step": int(n * GUN_K * 2 * 16), "ms_per_step": te,
                "note": "per rank: probe + coefficients up every step from page-locked host arrays (nepb_host_register); the moment block comes back on rank 0 only (reduce = 2)"},
        "gpu_launches": launches_total,
        "clocks": clocks,
        "eigenvalues_from_timed_moments": [[x.real, x.imag] for x in lam_found] if lam_found else None,
    }
    if spmm:
        head = spmm[1]
        traffic = None
        kernel_name = "spmm_fused_kernel<VW=4,real,SCALAR> (config C4, k=1)"
        try:
            # dram read+write per launch of THIS kernel from the committed ncu --set full capture; the file names the kernel it
            # was taken from, and a capture of another kernel is not reported
            with open(os.path.join(ROOT, "profiles", "r2_spmm_traffic.json")) as f:
                tj = json.load(f)
            if "spmm_fused_kernel" in tj.get("kernel", "spmm_fused_kernel"):
                traffic = tj["traffic_bytes_per_launch"]
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "achieved": head["gbs"], "peak": peak, "unit": "GB/s", "frac": head["frac"], "traffic": traffic,
                            "peak_source": peak_src, "kernel": kernel_name,
                            "algorithmic_bytes_per_launch": head["bytes"], "us_per_launch": head["ms"] * 1e3}
        line["spmm"] = {str(k): {kk: v[kk] for kk in ("ms", "gbs", "frac", "bytes", "e2e_ms", "parity_relerr")} for k, v in spmm.items()}
        line["spmm_kernels"] = {"1": "spmm_fused_kernel", "8": "spmm_tma2d_kernel<CPT=1> (2D tiles of 4 x 8 rows, TMA bulk staging)",
                                "20": "spmm_tma2d_kernel<CPT=3> (2D tiles of 4 x 8 rows, TMA bulk staging)"}
        line["spmm_general"] = general
        if dist.rank == 0 and dist.world == 1:  # auxiliary section: single-GPU runs only
            line["wep"] = bench_wep(args, peak)
    # factorisation roofline of the headline step: complex multiply-adds of the numeric LU (symbolic count) x 8 flops x nodes over
    # the step time, against a measured cuBLAS ZGEMM rate; the triangular solves and the assembly are in the time, not in the flops
    if dist.rank == 0:
        zpeak, zsrc = fp64_zgemm_peak()
        fl = 8.0 * sym["flops"] * GUN_N
        solve_bytes = 16.0 * sym["nnz_factor"] * GUN_N  # every factor entry streams once per 20-column solve
        line["roofline_contour"] = {"bound": "fp64", "achieved": fl / (t_step * 1e-3) / 1e12, "peak": zpeak * dist.world, "unit": "TFLOP/s",
                                    "frac": fl / (t_step * 1e-3) / 1e12 / (zpeak * dist.world), "peak_source": zsrc,
                                    "flops_per_step": fl, "factor_bytes_streamed_by_the_solves_per_step": solve_bytes,
                                    "note": "whole step time (factor + forward/backward solves + accumulate) against the factorisation flops only"}
    if dist.rank == 0 and dist.world == 1 and not args.no_cpu_baseline:  # the CPU baseline is timed at N = 1 only
        cores = max(1, min(host_cores(), 64))
        cpu = CpuContour(cores)
        import tempfile
        tmpd = tempfile.mkdtemp(prefix="nepb_bench_")
        nn, t = cpu.run(1, save_prefix=os.path.join(tmpd, "S"))
        cpu.close()
        line["cpu_baseline"] = {"value": nn / t, "unit": "solves/s", "cores": cores, "kind": "port",
                                "sample": "%d quadrature nodes of the same contour (one per worker process), k=20, SciPy SuperLU MMD_AT_PLUS_A, %.2f s" % (nn, t)}
        # parity at the stated size: the GPU path integrates exactly the nodes the CPU workers took (same probe, same weights)
        # and the two partial moment blocks must agree to 1e-10
        try:
            Scpu = sum(np.load(os.path.join(tmpd, "S.%d.npy" % c)) for c in range(cores))
            idx = np.arange(cores) % GUN_N
            dn2 = gun_operator()
            integ2 = nepb200.ContourIntegrator(dn2, GUN_K, 2, min(len(idx), args.batch))
            Sgpu, fl2 = integ2.integrate(lams[idx], W[idx], Vh, reduce=False)
            integ2.close()
            dn2.close()
            perr = float(np.linalg.norm(Sgpu - Scpu) / np.linalg.norm(Scpu))
            line["contour_parity"] = {"relerr_moments_vs_cpu_arm": perr, "nodes": int(len(idx)), "k": GUN_K, "tolerance": 1e-10}
            log("[bench] contour parity: %d-node partial moments (k=%d) GPU vs CPU arm: %.2e" % (len(idx), GUN_K, perr))
            if not perr < 1e-10:
                raise SystemExit("bench: contour moments differ from the CPU arm: %g" % perr)
        finally:
            import shutil
            shutil.rmtree(tmpd, ignore_errors=True)
        if cpu_spmm:
            line["cpu_baseline_spmm"] = cpu_spmm
    if dist.rank == 0:
        print(json.dumps(line), flush=True)
    if dist.world > 1:
        lib.nepb_comm_destroy()
    dist.close()


if __name__ == "__main__":
    main()
