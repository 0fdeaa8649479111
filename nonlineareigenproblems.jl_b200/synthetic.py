"""Deterministic synthetic SPMF problems of the benchmark shapes.

Vectorised twin of oracle.gallery.stencil_pep (the oracle's loop version is the checker at small
sizes): degree-3 PEP on a g x g grid, 21-point stencil (5x5 neighbourhood minus its corners), values
from the Middle-Square-Weyl stream (nepb_msws_fill), term after term in CSR order.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import _lib

STENCIL = [(di, dj) for di in range(-2, 3) for dj in range(-2, 3) if not (abs(di) == 2 and abs(dj) == 2)]


def stencil_pattern(g: int):
    """CSR (indptr, indices, kinds) -- kinds: 0 diagonal, 1 direct neighbour, 2 other."""
    i = np.repeat(np.arange(g, dtype=np.int64), g)
    j = np.tile(np.arange(g, dtype=np.int64), g)
    cols = np.empty((g * g, len(STENCIL)), dtype=np.int64)
    ok = np.empty((g * g, len(STENCIL)), dtype=bool)
    kinds = np.empty(len(STENCIL), dtype=np.int8)
    for s, (di, dj) in enumerate(STENCIL):
        ii, jj = i + di, j + dj
        ok[:, s] = (ii >= 0) & (ii < g) & (jj >= 0) & (jj < g)
        cols[:, s] = ii * g + jj
        kinds[s] = 0 if (di == 0 and dj == 0) else (1 if abs(di) + abs(dj) == 1 else 2)
    indptr = np.zeros(g * g + 1, dtype=np.int64)
    np.cumsum(ok.sum(axis=1), out=indptr[1:])
    indices = cols[ok]
    kind = np.broadcast_to(kinds, ok.shape)[ok]
    return indptr, indices, kind


def stencil_pep(g: int, seed: int = 0):
    """A0..A3 as CSR matrices with one shared pattern, plus the MSWS state after the draw."""
    indptr, indices, kind = stencil_pattern(g)
    nnz = len(indices)
    st = _lib.msws_state(seed)
    lap = np.where(kind == 0, 4.0, np.where(kind == 1, -1.0, 0.0))
    vals = [lap + 0.1 * _lib.msws_fill(st, nnz)]
    for s in (1.0, 0.1, 0.01):
        vals.append(s * (1 - 2 * _lib.msws_fill(st, nnz)))
    n = g * g
    return [sp.csr_matrix((v, indices, indptr), shape=(n, n)) for v in vals], st


def stencil_block(st, n: int, k: int) -> np.ndarray:
    """Complex n x k block (column-major draw order, re then im, each 1-2u)."""
    u = 1 - 2 * _lib.msws_fill(st, 2 * n * k)
    return np.asfortranarray((u[0::2] + 1j * u[1::2]).reshape(n, k, order="F"))
