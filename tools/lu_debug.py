import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, scipy.sparse as sp
import nepb200
from nepb200 import B200SPMF, ONE
P = np.array([[0.0, 2.0, 0.0], [1.0, 0.0, 3.0], [0.0, 4.0, 1e-3]])
d = B200SPMF([sp.csc_matrix(P)], [ONE])
print(nepb200.symbolic_get(d))
s = nepb200.B200FactorizeLinSolver(d, 0.0, umfpack_refinements=0)
x = s.lin_solve(np.array([1.0, 2.0, 3.0]))
print("P:", x, np.linalg.solve(P, [1.0, 2.0, 3.0]), s.status)

for r in (0, 1, 10):
    s = nepb200.B200FactorizeLinSolver(d, 0.0, umfpack_refinements=r)
    x = s.lin_solve(np.array([1.0, 2.0, 3.0]))
    print("refine", r, x, "berr", s.lu.last_berr, "resid", np.linalg.norm(P @ x - [1.0, 2.0, 3.0]))
