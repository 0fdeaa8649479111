"""Regenerate the binary data fixtures from the reference's text matrices.

Run in the build container only (needs /root/reference):  python tests/golden/make_fixtures.py
The text format is that of src/utils/Serialization.jl:8-31; the conversion goes through
oracle.gallery.read_sparse_matrix.  Outputs (committed): gun.npz, qdep0.npz.
The GPU box has no /root/reference, so tests and bench read these instead.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.gallery import read_sparse_matrix  # noqa: E402

REF = "/root/reference/src/gallery_extra"


def pack(d, name, A):
    d[name + "_data"] = A.data.astype(np.float64)
    d[name + "_indices"] = A.indices.astype(np.int32)
    d[name + "_indptr"] = A.indptr.astype(np.int32)


def main():
    d = {}
    for name in ("K", "M", "W1", "W2"):
        A = read_sparse_matrix(os.path.join(REF, "converted_nlevp", "gun_%s.txt" % name))
        pack(d, name, A)
        d["n"] = np.int64(A.shape[0])
        print("gun", name, A.shape, A.nnz, "1-norm", abs(A).sum(axis=0).max())
    np.savez_compressed(os.path.join(HERE, "gun.npz"), **d)
    d = {}
    for name in ("A0", "A1"):
        A = read_sparse_matrix(os.path.join(REF, "converted_misc", "qdep_infbilanczos_%s.txt" % name))
        pack(d, name, A)
        d["n"] = np.int64(A.shape[0])
        print("qdep0", name, A.shape, A.nnz)
    np.savez_compressed(os.path.join(HERE, "qdep0.npz"), **d)


if __name__ == "__main__":
    main()
