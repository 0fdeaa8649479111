"""CPU oracle for the nepb200 hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the NEP-PACK (NonlinearEigenproblems.jl v1.1.1)
algorithms that sit on the hot path named in BASELINE.json: SPMF/PEP/DEP/SumNEP
``compute_Mder`` / ``compute_Mlincomb`` / ``compute_MM``, ``lin_solve`` behind the
``LinSolver`` / ``LinSolverCreator`` pair, the trapezoidal contour integrator and
``contour_beyn``, and the ``iar`` / ``tiar`` / ``resinv`` drivers.  Each function cites the
reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import anything from here, and there only as the checker or the timed CPU baseline.
The product (``nonlineareigenproblems.jl_b200``) never imports this package.

Parity pinning: the reference is 100 % Julia and Julia is not installed in the build container,
so the reference itself cannot be executed.  The oracle is pinned against the literal known
answers the reference's docs/tests hold (see ``tests/test_oracle_golden.py``): dep0 generator
values, gun 1-norms, the gun reference eigenvalue, dep0 eigenvalue counts, docstring values.
LU factors / pivot order / Gram-Schmidt coefficients live in third-party libraries (UMFPACK,
IterativeSolvers) and are pinned by the reference's own tests only at the solution level;
the oracle follows that: solution-level parity ("parity unpinned" for the factors themselves).
"""
