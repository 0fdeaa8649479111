"""Deflation of computed eigenpairs (Effenberger) on top of the device operator, and the Schur-complement linear solver that
recycles the device factorisation -- host-side mirror of

  DeflatedGenericNEP, compute_Mlincomb / compute_Mder / compute_MM       src/nep_deflation.jl:46-52,65-147,183-197
  deflated_nep_compute_Q                                                  src/nep_deflation.jl:149-172
  deflate_eigpair, normalize_schur_pair!, get_deflated_eigpairs           src/nep_deflation.jl:278-287,369-440
  DeflatedNEPLinSolver / DeflatedNEPLinSolverCreator                      src/LinSolvers.jl:209-252, src/LinSolverCreators.jl:147-178

The deflated problem has size n + p (p deflated pairs).  Everything of size n goes through the original operator's device
calls: `compute_Mlincomb` of the original NEP (one fused SpMM each) and `lin_solve` of the original solver with the p columns of
U = Q(lambda) as extra right-hand sides in ONE block solve (the reference loops over the columns, LinSolvers.jl:241-243).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sl
import scipy.sparse as sp

from .linsolve import LinSolver, LinSolverCreator, B200LinSolverCreator


class DeflatedGenericNEP:
    """The deflated NEP  [M(lam) U(lam); X^H 0]  with U(lam) = M(lam) X (lam I - S)^-1 for the invariant pair (X, S) = (V0, S0)."""

    def __init__(self, orgnep, S0, V0):
        self.orgnep = orgnep
        self.S0 = np.atleast_2d(np.asarray(S0, dtype=np.complex128))
        self.V0 = np.asarray(V0, dtype=np.complex128).reshape(orgnep.n, -1)
        self.n = orgnep.n + self.V0.shape[1]

    def size(self, d=None):
        return (self.n, self.n) if d is None else self.n

    # deflated_nep_compute_Q (:149-172): Q = sum_{i<=der} (-1)^(der-i) der!/i! M^(i)(lam) X (lam I - S)^-(der-i+1)
    def compute_Q(self, lam, der=0):
        X, S = self.V0, self.S0
        p = S.shape[0]
        F = sl.lu_factor(lam * np.eye(p) - S)
        Q = np.zeros((self.orgnep.n, p), dtype=np.complex128)
        Vnew = X
        for i in range(der, -1, -1):
            Vnew = sl.lu_solve(F, Vnew.T, trans=1).T  # Vnew / F
            factor = (-1) ** (der - i) * (math.factorial(der) / math.factorial(i))
            for j in range(p):
                Q[:, j] += self.orgnep.compute_Mlincomb(lam, Vnew[:, j], np.array([factor], dtype=np.complex128), i)
        return Q

    # compute_Mlincomb (:65-108), binomial expansion of the derivatives of U(lam)
    def compute_Mlincomb(self, lam, V, a=None, startder=None):
        V = np.asarray(V, dtype=np.complex128)
        V = V.reshape(self.n, -1, order="F") if V.ndim == 1 else V
        k = V.shape[1]
        a = np.ones(k, dtype=np.complex128) if a is None else np.asarray(a, dtype=np.complex128)
        if startder:  # NEPCore.jl:156-160
            V = np.concatenate([np.zeros((self.n, startder), dtype=np.complex128), V], axis=1)
            a = np.concatenate([np.zeros(startder, dtype=np.complex128), a])
            k += startder
        X, S = self.V0, self.S0
        n0, p = self.orgnep.n, S.shape[0]
        F = sl.lu_factor(lam * np.eye(p) - S)
        Xhat = sl.lu_solve(F, X.T, trans=1).T
        Q = []
        for i in range(k):
            QQ = np.zeros((p, k), dtype=np.complex128)
            QQ[:, i] = V[n0:, i]
            for j in range(i - 1, -1, -1):
                QQ[:, j] = sl.lu_solve(F, QQ[:, j + 1])
            Q.append(QQ)
        Z = np.zeros((n0, k), dtype=np.complex128)
        for j in range(k):
            for i in range(j, k):
                factor = (-1) ** (i - j) * (a[i] * math.factorial(i) / math.factorial(j))
                Z[:, j] += factor * (Xhat @ Q[i][:, j])
        Vnew = V[:n0, :] * a[None, :] + Z
        z_top = self.orgnep.compute_Mlincomb(lam, Vnew)
        z_bottom = X.conj().T @ V[:n0, 0] * a[0]
        return np.concatenate([z_top, z_bottom])

    # compute_Mder (:110-147)
    def compute_Mder(self, lam, der=0):
        n0, p = self.orgnep.n, self.S0.shape[0]
        Q = self.compute_Q(lam, der)
        M0 = self.orgnep.compute_Mder(lam, der)
        low = self.V0.conj().T if der == 0 else np.zeros((p, n0), dtype=np.complex128)
        if sp.issparse(M0):
            return sp.bmat([[M0, sp.csc_matrix(Q)], [sp.csc_matrix(low), sp.csc_matrix((p, p), dtype=np.complex128)]], format="csc")
        return np.block([[np.asarray(M0), Q], [low, np.zeros((p, p), dtype=np.complex128)]])

    # compute_MM (:183-197)
    def compute_MM(self, S, V):
        S = np.atleast_2d(np.asarray(S, dtype=np.complex128))
        V = np.asarray(V, dtype=np.complex128)
        n0, p0, p = self.orgnep.n, self.S0.shape[0], S.shape[0]
        V1, V2 = V[:n0, :], V[n0:, :]
        Stilde = np.block([[self.S0, V2], [np.zeros((p, p0), dtype=np.complex128), S]])
        Vtilde = np.concatenate([self.V0, V1], axis=1)
        R = self.orgnep.compute_MM(Stilde, Vtilde)
        return np.concatenate([R[:n0, p0:], self.V0.conj().T @ V1], axis=0)


def normalize_schur_pair(S, V):
    """normalize_schur_pair! (:278-287): V orthonormal by a thin QR, S -> R S R^-1."""
    QQ, RR = np.linalg.qr(V)
    return RR @ S @ np.linalg.inv(RR), QQ


def deflate_eigpair(nep, lam, v):
    """deflate_eigpair (:369-398), mode :Generic -- for a plain NEP or an already deflated one (the partial Schur form grows)."""
    v = np.asarray(v, dtype=np.complex128)
    if isinstance(nep, DeflatedGenericNEP):
        n, p0 = nep.orgnep.n, nep.V0.shape[1]
        V1 = np.zeros((n, p0 + 1), dtype=np.complex128)
        S1 = np.zeros((p0 + 1, p0 + 1), dtype=np.complex128)
        V1[:, :p0] = nep.V0
        V1[:, p0] = v[:n]
        S1[:p0, :p0] = nep.S0
        S1[:, p0] = np.concatenate([v[n:], [lam]])
        S1, V1 = normalize_schur_pair(S1, V1)
        return DeflatedGenericNEP(nep.orgnep, S1, V1)
    S0 = np.array([[complex(lam)]])
    V0 = v.reshape(-1, 1)
    S0, V0 = normalize_schur_pair(S0, V0)
    return DeflatedGenericNEP(nep, S0, V0)


def get_deflated_eigpairs(dnep):
    """get_deflated_eigpairs (:433-438): eigenpairs of the original problem from the invariant pair."""
    D, X = np.linalg.eig(dnep.S0)
    return D, dnep.V0 @ X


class DeflatedNEPLinSolver(LinSolver):
    """[M U; X^H 0] [v1; v2] = [b1; b2] by the Schur complement S = -X^H M^-1 U, recycling the solver of the original NEP
    (LinSolvers.jl:221-252).  M^-1 [b1 U] is one block solve with p + 1 right-hand sides."""

    def __init__(self, deflated_nep: DeflatedGenericNEP, lam, orglinsolver):
        self.deflated_nep, self.lam, self.orglinsolver = deflated_nep, lam, orglinsolver

    def lin_solve(self, b, tol=0):
        dn = self.deflated_nep
        n, m = dn.orgnep.n, dn.S0.shape[0]
        b = np.asarray(b, dtype=np.complex128)
        b1, b2 = b[:n], b[n:]
        U = dn.compute_Q(self.lam, 0)
        sol = np.asarray(self.orglinsolver.lin_solve(np.concatenate([b1.reshape(n, 1), U], axis=1)))
        b1t, Z = sol[:, 0], sol[:, 1:]
        X = dn.V0
        Sc = -X.conj().T @ Z
        v2 = np.linalg.solve(Sc, b2 - X.conj().T @ b1t)
        return np.concatenate([b1t - Z @ v2, v2])


class DeflatedNEPLinSolverCreator(LinSolverCreator):
    """DeflatedNEPLinSolverCreator(orglinsolvercreator) (LinSolverCreators.jl:147-178)."""

    def __init__(self, orglinsolvercreator=None):
        self.orglinsolvercreator = orglinsolvercreator or B200LinSolverCreator()

    def create_linsolver(self, nep: DeflatedGenericNEP, lam):
        return DeflatedNEPLinSolver(nep, lam, self.orglinsolvercreator.create_linsolver(nep.orgnep, lam))
