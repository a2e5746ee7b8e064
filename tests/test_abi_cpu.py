"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/bnmtf_b200.h
declares, and the Python classes validate their arguments like the reference (no compute calls without a GPU)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from bnmtf_b200 import _lib
    return _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bnmtf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bnmt?f_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    l = lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(l, n), "missing export %s" % n
        assert n in lib.SIGNATURES, "ctypes prototype missing for %s" % n
    assert set(lib.SIGNATURES) == set(names)


def test_plain_helpers(lib):
    l = lib.load()
    assert l.bnmtf_version() >= 100
    assert l.bnmtf_ld_for(80) == 128 and l.bnmtf_ld_for(32768) == 32768
    assert l.bnmtf_kp_for(20) == 24 and l.bnmtf_kp_for(7) == 8 and l.bnmtf_kp_for(8) == 16
    assert l.bnmtf_gram_len(20) == 6 * 64
    # scratch and row partitions of the S-phase reduction (host-side queries, through the same wrapper the engine uses)
    for K, L, vb in ((10, 10, 1), (10, 10, 0), (20, 20, 1), (3, 4, 0), (32, 32, 1), (63, 63, 1)):
        D = K * L
        assert lib.call("bnmtf_nmtf_sq_scratch_len", K, L, vb) >= D * D + 2 * D
        assert 1 <= lib.call("bnmtf_nmtf_sq_parts", 65536, K, L, vb) <= 148
        assert lib.call("bnmtf_nmtf_sq_parts", 1, K, L, vb) == 1


def test_constructor_validation_matches_reference_messages():
    """reference tests/code/test_bnmf_gibbs_optimised.py:24-65 (test_init)"""
    import bnmtf_b200
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
    for cls in (bnmtf_b200.bnmf_gibbs_optimised, bnmtf_b200.bnmf_vb_optimised, bnmtf_b200.nmf_icm):
        with pytest.raises(AssertionError) as e:
            cls(np.ones(3), np.ones((2, 3)), 3, pri)
        assert str(e.value) == "Input matrix R is not a two-dimensional array, but instead 1-dimensional."
        with pytest.raises(AssertionError) as e:
            cls(np.ones((4, 3, 2)), np.ones((3, 2)), 3, pri)
        assert str(e.value) == "Input matrix R is not a two-dimensional array, but instead 3-dimensional."
        with pytest.raises(AssertionError) as e:
            cls(np.ones((3, 2)), np.ones((2, 3)), 3, pri)
        assert str(e.value) == "Input matrix R is not of the same size as the indicator matrix M: (3, 2) and (2, 3) respectively."
        with pytest.raises(AssertionError) as e:
            cls(np.ones((2, 3)), np.ones((2, 3)), 3, {"alpha": 1., "beta": 1., "lambdaU": np.ones((2, 4)), "lambdaV": 0.1})
        assert str(e.value) == "Prior matrix lambdaU has the wrong shape: (2, 4) instead of (2, 3)."
        M = np.ones((2, 3))
        M[0] = 0
        with pytest.raises(AssertionError) as e:
            cls(np.ones((2, 3)), M, 3, pri)
        assert str(e.value) == "Fully unobserved row in R, row 0."
        M = np.ones((2, 3))
        M[:, 2] = 0
        with pytest.raises(AssertionError) as e:
            cls(np.ones((2, 3)), M, 3, pri)
        assert str(e.value) == "Fully unobserved column in R, column 2."
        m = cls(np.ones((2, 3)), np.ones((2, 3)), 3, pri)
        assert m.lambdaU.shape == (2, 3) and m.lambdaV.shape == (3, 3) and m.size_Omega == 6
        with pytest.raises(AssertionError):
            m.initialise("nope")


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bnmtf_b200
    from bnmtf_b200._lib import BnmtfError
    m = bnmtf_b200.nmf_icm(np.ones((2, 3)), np.ones((2, 3)), 2, {"alpha": 1., "beta": 1., "lambdaU": 0.1, "lambdaV": 0.1})
    with pytest.raises(BnmtfError):
        m.initialise("exp")     # needs beta_s() -> device


def test_product_code_never_imports_the_oracle():
    for root, _dirs, files in os.walk(os.path.join(ROOT, "bnmtf_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(root, f)).read(), f
