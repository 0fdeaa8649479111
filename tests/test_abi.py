"""C-ABI library: loads, exports every symbol include/nepb200.h declares, and refuses to compute without
a device (no CPU fallback).  CPU only -- no kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nepb200
from nepb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nepb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nepb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(_lib.lib, s), "libnepb200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "%s has no ctypes signature" % s
    for s in _lib.SIGNATURES:
        assert s in syms, "%s bound but not declared in the header" % s


def test_version_and_msws_stream():
    assert b"sm_100a" in _lib.lib.nepb_version()
    from oracle import gallery as g
    st = _lib.msws_state(0)
    got = _lib.msws_fill(st, 1000)
    r = g.MSWS_RNG(0)
    ref = np.array([g.gen_rng_float(r) for _ in range(1000)])
    assert np.array_equal(got, ref)  # bit-exact integer recurrence
    st = _lib.msws_state(12345678901234567890123)
    r = g.MSWS_RNG(12345678901234567890123)
    assert _lib.msws_fill(st, 5)[4] == [g.gen_rng_float(r) for _ in range(5)][4]


@pytest.mark.skipif(nepb200.device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    import scipy.sparse as sp
    A = sp.identity(4, format="csc")
    with pytest.raises(nepb200.NepbError) as e:
        nepb200.B200SPMF([A], [nepb200.ONE])
    assert e.value.status == _lib.NEPB_E_CUDA
    assert "no CPU fallback" in str(e.value)


def test_argument_errors_do_not_need_a_device():
    h = C.c_void_p()
    assert _lib.lib.nepb_spmf_create(0, 1, None, None, None, 0, 1, C.byref(h)) == _lib.NEPB_E_INVALID
    assert "n must be positive" in _lib.last_error()


def test_synthetic_generator_matches_oracle():
    from nepb200 import synthetic
    from oracle import gallery as g
    mats, st = synthetic.stencil_pep(9)
    ref, rng = g.stencil_pep(9)
    for A, B in zip(mats, ref):
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        assert np.array_equal(A.data, B.data)
    assert np.array_equal(synthetic.stencil_block(st, 81, 2), g.stencil_block(rng, 81, 2))
    indptr, _, _ = synthetic.stencil_pattern(1000)
    assert indptr[-1] == 20956020  # nnz_u of config C4
