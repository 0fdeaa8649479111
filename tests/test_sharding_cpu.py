"""The N>1 host logic of the contour path on CPU: world_size-2 gloo run.  Every rank integrates the quadrature nodes the
product's partition (`beyn_quadrature`) assigns to it -- with the oracle's CPU solver standing in for the device call --
and one all-reduce of the moment block must reproduce the serial oracle integral and the same eigenvalues."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import torch, torch.distributed as td
    import nepb200
    from oracle import nep as o, solvers as osol, gallery as g
    td.init_process_group(backend="gloo")
    rank, world = td.get_rank(), td.get_world_size()
    nep = o.nep_gallery("dep0")
    Vh = g.gen_rng_mat(g.MSWS_RNG(0), 5, 4)
    N, sigma, radius = 64, 0.2, 1.0
    mine, lams, W = nepb200.beyn_quadrature(N, radius, sigma, rank, world)
    S = np.zeros((5, 4, 2), dtype=complex)
    for lam, w in zip(lams, W):
        X = osol.FactorizeLinSolver(nep, lam).lin_solve(Vh)
        S[:, :, 0] += w[0] * X
        S[:, :, 1] += w[1] * X
    t = torch.from_numpy(np.ascontiguousarray(S.view(np.float64)))
    td.all_reduce(t)                      # the single reduce of the moment block
    S = t.numpy().view(np.complex128).reshape(5, 4, 2)
    lo, Vo, A0, A1, info = osol.contour_beyn(nep, Vh, sigma=sigma, radius=radius, N=N, neigs=3, k=4, return_moments=True, sanity_check=False)
    counts = torch.zeros(N, dtype=torch.int64); counts[torch.from_numpy(mine)] += 1
    td.all_reduce(counts)
    assert bool((counts == 1).all()), "every node must be owned by exactly one rank"
    assert np.linalg.norm(S[:, :, 0] - A0) < 1e-13 * np.linalg.norm(A0)
    assert np.linalg.norm(S[:, :, 1] - A1) < 1e-13 * np.linalg.norm(A1)
    lam, V, inf2 = nepb200.beyn_extract(S[:, :, 0], S[:, :, 1], sigma, (radius, radius), 4, 3, 1e-8, 1e-8, None, False)
    assert len(lam) == len(lo) and max(min(abs(lam - x)) for x in lo) < 1e-12
    open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rank" + str(rank) + ".ok"), "w").write(str(len(mine)))
    td.destroy_process_group()
""") % ROOT


def test_world_size_2_gloo_sharded_contour(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    import socket
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for attempt in range(3):  # the gloo rendezvous on a loaded CI box occasionally times out: retry on a fresh port
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
               "--master-port", str(port), str(script)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        if r.returncode == 0:
            break
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert [int((tmp_path / ("rank%d.ok" % k)).read_text()) for k in (0, 1)] == [32, 32]


def test_partition_properties():
    import nepb200
    for N, world in ((128, 1), (128, 8), (100, 3), (7, 8)):
        seen = np.zeros(N, dtype=int)
        Wsum = np.zeros(2, dtype=complex)
        for r in range(world):
            mine, lams, W = nepb200.beyn_quadrature(N, 2.0, 1.0 + 1j, r, world)
            seen[mine] += 1
            Wsum += W.sum(axis=0)
            assert len(lams) == len(mine) == W.shape[0]
        assert np.all(seen == 1)
        # trapezoid weights of a closed contour: sum_i gp_i h = 0 and sum_i gp_i g_i h = 0 (exactly integrable)
        assert abs(Wsum[0]) < 1e-13 and abs(Wsum[1]) < 1e-12
