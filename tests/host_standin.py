"""NumPy stand-ins with the contract of the device operator (`apply` in GENERAL mode), its solver creator and its residual
error measure: they let the CPU tests run the product's HOST recurrences (infbilanczos, projection) against the reference's
literals; the ABI calls behind the real objects are covered by the GPU tests.  Test infrastructure only."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from nepb200 import _lib, B200SPMF


class HostOperator:
    """NumPy stand-in for B200SPMF with the same `apply` contract (GENERAL mode only)."""

    def __init__(self, A, fi):
        self.A, self.fi, self.p, self.n = A, fi, len(A), A[0].shape[0]

    # host-side coefficient logic of the real operator (needs only fi / p / n / apply)
    lincomb_coefficients = B200SPMF.lincomb_coefficients
    compute_Mlincomb = B200SPMF.compute_Mlincomb
    compute_MM = B200SPMF.compute_MM

    def coefficients(self, lam, der=0):
        return np.array([complex(f.derivative(lam, der)) if der else complex(f(complex(lam))) for f in self.fi])

    def compute_Mder(self, lam, der=0):
        c = self.coefficients(lam, der)
        return sum(ci * A for ci, A in zip(c, self.A))

    def get_fv(self):
        return self.fi

    def get_Av(self):
        return self.A

    def apply(self, mode, V, blocks, q, out=None):
        V = np.asarray(V, dtype=np.complex128)
        V = V.reshape(-1, 1) if V.ndim == 1 else V
        k = V.shape[1]
        blocks = np.asarray(blocks, dtype=np.complex128)
        if mode == _lib.COEF_SCALAR:
            return sum(self.A[t] @ (V * blocks.ravel()[t]) for t in range(self.p))
        if mode == _lib.COEF_DIAG:  # p x k column-major: C[i + p*s]
            Cd = blocks.reshape(self.p, k, order="F")
            return sum(self.A[t] @ (V * Cd[t][None, :]) for t in range(self.p))
        blocks = blocks.reshape(self.p, -1)
        return sum(self.A[t] @ (V @ blocks[t].reshape(k, q, order="F")) for t in range(self.p))


class HostSolverCreator:
    def create_linsolver(self, op, lam):
        M = sum(complex(f(complex(lam))) * A for f, A in zip(op.fi, op.A))
        lu = sla.splu(sp.csc_matrix(M, dtype=np.complex128))

        class S:
            def lin_solve(self, b, tol=0):
                return lu.solve(np.asarray(b, dtype=np.complex128))
        return S()


class HostResidual:
    def __init__(self, op):
        self.op = op

    def estimate_error(self, lam, v):
        M = sum(complex(f(complex(lam))) * A for f, A in zip(self.op.fi, self.op.A))
        return float(np.linalg.norm(M @ v) / np.linalg.norm(v))
